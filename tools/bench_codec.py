#!/usr/bin/env python
"""BASELINE config 5: codec decode of a 10k-token stream (SNAC-24k, Mimi) on one GPU -- audio-s/s and achieved HBM GB/s of
the conv stack against the measured copy bandwidth (algorithmic bytes: SURVEY.md section 8d, bf16 convention)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import ops, synth
from mlx_audio_b200.codec import SNAC, Mimi, mimi_202407
from mlx_audio_b200.configs import MIMI_202407, SNAC_24K

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--which", default="snac,mimi")
args = ap.parse_args()
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}


def timed(fn, steps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, y


def profile(fn):
    prof = {}
    ops.PROFILE = prof
    fn()
    torch.cuda.synchronize()
    ops.PROFILE = None
    return {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in prof.items()}


if "snac" in args.which:
    T = args.frames // 4 * 4
    model = SNAC.from_config(SNAC_24K, device=dev).load_weights(synth.snac_weights(SNAC_24K))
    codes = [c.to(dev) for c in synth.snac_codes(SNAC_24K, T)]
    noises = [n.to(dev) for n in synth.snac_noises(SNAC_24K)]
    ms, y = timed(lambda: model.decode(codes, noises=noises), args.steps)
    secs = y.shape[1] / 24000
    alg = 27648.4e6 * T / 10000
    print(json.dumps({"codec": "snac-24k", "frames": T, "samples": y.shape[1], "ms": ms, "audio_s_per_s": secs / (ms / 1e3),
                      "algorithmic_GB": alg / 1e9, "achieved_GBps_bf16_convention": alg / (ms / 1e3) / 1e9,
                      "frac_of_measured_hbm": alg / (ms / 1e3) / 1e9 / peaks["hbm_gbs"], "tc_mode": ops.TC_MODE[0],
                      "ms_by_kind": profile(lambda: model.decode(codes, noises=noises))}))
if "mimi" in args.which:
    T = args.frames
    model = Mimi(mimi_202407(32), device=dev).load_weights(synth.mimi_weights(MIMI_202407))
    codes = synth.mimi_codes(MIMI_202407, T).to(dev)
    ms, y = timed(lambda: model.decode(codes), args.steps)
    secs = y.shape[2] / 24000
    alg = 21909.8e6 * T / 10000
    print(json.dumps({"codec": "mimi", "frames": T, "samples": y.shape[2], "ms": ms, "audio_s_per_s": secs / (ms / 1e3),
                      "algorithmic_GB_seanet": alg / 1e9, "achieved_GBps_bf16_convention": alg / (ms / 1e3) / 1e9,
                      "frac_of_measured_hbm": alg / (ms / 1e3) / 1e9 / peaks["hbm_gbs"], "tc_mode": ops.TC_MODE[0],
                      "ms_by_kind": profile(lambda: model.decode(codes))}))
