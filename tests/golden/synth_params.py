"""Deterministic parameter values keyed by (parameter name, shape): the golden generators fill the REFERENCE's modules with
them and the tests rebuild the identical values for the oracle / the CUDA path from the (name, shape) list stored in the
fixture, so a fixture holds inputs and outputs only."""
import json
import zlib

import numpy as np


def value(name: str, shape, rule=None) -> np.ndarray:
    """``rule`` "scale<f>": the default value times f (e.g. a final projection kept inside the clip range); "small": 0.1 N(0,1) (log-domain gains such as the Qwen3 tokenizer's SnakeBeta alpha/beta, kept near exp(0) = 1 so
    that the deep decoder stays well conditioned and float64 runs agree to ~1e-12)."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    shape = tuple(int(s) for s in shape)
    if rule == "small":
        return 0.1 * rng.standard_normal(shape)
    if rule and rule.startswith("scale"):
        return float(rule[5:]) * value(name, shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf.startswith("bias") or leaf.endswith("bias") or leaf in ("beta",):
        return 0.1 * rng.standard_normal(shape)
    if len(shape) <= 1 and leaf not in ("alpha",):
        return 1.0 + 0.1 * rng.standard_normal(shape)                 # norm gains, layer scales
    if leaf == "alpha":
        return 0.5 + rng.random(shape)                                # snake alphas: positive
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    return rng.standard_normal(shape) / np.sqrt(max(fan_in, 1))


def manifest(named_shapes) -> str:
    """entries (name, shape) or (name, shape, rule)"""
    return json.dumps([[e[0], [int(s) for s in e[1]]] + ([e[2]] if len(e) > 2 and e[2] else []) for e in named_shapes])


def from_manifest(text) -> dict:
    return {e[0]: value(e[0], e[1], e[2] if len(e) > 2 else None) for e in json.loads(str(text))}
