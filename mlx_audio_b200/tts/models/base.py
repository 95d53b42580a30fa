"""Return contracts shared by the TTS models (reference: tts/models/base.py:8-99).

The two result records carry exactly the reference's fields in the reference's order (checked against the reference's own dataclasses in
tests/test_host_cpu.py via tests/golden/config_golden.json); ``audio`` is a 1-D float32 torch tensor instead of an ``mx.array``."""
from __future__ import annotations

import inspect
from dataclasses import dataclass, make_dataclass

import torch


@dataclass
class BaseModelArgs:
    @classmethod
    def from_dict(cls, params):
        """Keep only the keys the dataclass declares (reference tts/models/base.py:10-18)."""
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})


def check_array_shape(arr) -> bool:
    """Layout heuristic of the reference's ``sanitize`` (tts/models/base.py:21-34): True when a 3-D
    conv weight already is (out, K, K')-shaped with out the largest and the last two equal."""
    shape = arr.shape
    if len(shape) != 3:
        return False
    out_channels, kh, kw = shape
    return (out_channels >= kh) and (out_channels >= kw) and (kh == kw)


def _record(name, required, flags=("is_streaming_chunk", "is_final_chunk")):
    return make_dataclass(name, [(n, t) for n, t in required] + [(f, bool, False) for f in flags])


_TIMING = [("processing_time_seconds", float), ("peak_memory_usage", float)]

# one segment of Model.generate (tts/models/base.py:71-86)
GenerationResult = _record("GenerationResult", [("audio", torch.Tensor), ("samples", int), ("sample_rate", int), ("segment_idx", int), ("token_count", int),
                                                ("audio_duration", str), ("real_time_factor", float), ("prompt", dict), ("audio_samples", dict)] + _TIMING)

# one sequence of Model.batch_generate (tts/models/base.py:89-99)
BatchGenerationResult = _record("BatchGenerationResult", [("audio", torch.Tensor), ("sequence_idx", int), ("samples", int), ("sample_rate", int),
                                                          ("token_count", int), ("audio_duration", str)] + _TIMING)
