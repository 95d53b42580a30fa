"""Cached ORACLE outputs at the shapes the benchmarks time (VERDICT r01 "make the benched shapes the tested shapes").

The float64 oracle (oracle/, pinned to the reference's own code by the other fixtures in this directory) needs seconds to minutes on
these sizes, which is too slow for the GPU suite and out of bounds for bench.py's GPU arm -- so it is run HERE, once, and the result is
committed as ``bench_shapes_golden.npz``:

  kokoro_*   BASELINE config 2: 128 phonemes (T = 130), durations 3 -> F = 390, 234 000 samples.  Waveform with the oracle's own
             float32-rounded F0 / N curves injected (DESIGN.md "conditioning of the harmonic source"), the curves, the stage taps'
             RMS.
  snac_*     SNAC-24k, 2 048 fine frames (1 048 651 samples): three 16 384-sample windows (head / middle / tail) + global mean
             square, so that a [B, 1M]-sample decode is compared without committing 4 MB.
  mimi_*     Mimi, 2 000 frames (3 840 000 samples): windows likewise.
  whisper_*  Whisper-small TextDecoder with all 12 layers: logits of the first sampled position for 2 rows, 8 greedy tokens,
             sum of log-probabilities, no-speech probabilities.
  qwen3_*    Qwen3-TTS-0.6B talker + code predictor (28 + 5 layers): 25 frames x 16 code books with injected uniforms.

Run:  python tests/golden/make_bench_shape_golden.py [--only kokoro,snac,mimi,whisper,qwen3]   (about 10 minutes on 8 cores)
Inputs come from ``mlx_audio_b200.synth`` with the seeds written below, so tests rebuild them bit-identically.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "bench_shapes_golden.npz")

WIN = 16384


def windows(n):
    """Start offsets of the head / middle / tail comparison windows of a length-n signal."""
    return [0, (n // 2 // 128) * 128, n - WIN]


def kokoro(out):
    from mlx_audio_b200 import synth
    from oracle import kokoro as OK
    P64 = {k: v.double() for k, v in synth.kokoro_weights(OK.KOKORO_CONFIG, seed=0).items()}
    ids, ref_s = synth.kokoro_inputs(128, seed=1)
    T = ids.shape[1]
    nz = synth.kokoro_noise(T * 3 * 600, 3)[1].double()
    OK.TAP = {}
    free, _ = OK.forward(P64, ids, ref_s.double(), noise=nz, pred_dur_override=[3] * T)
    f0, n = OK.TAP["F0"].float().reshape(-1), OK.TAP["N"].float().reshape(-1)
    OK.TAP = {}
    audio, _ = OK.forward(P64, ids, ref_s.double(), noise=nz, pred_dur_override=[3] * T, f0n_override=(f0, n))
    tap, OK.TAP = OK.TAP, None
    out["kokoro_audio"] = audio.float().numpy()
    out["kokoro_f0"], out["kokoro_n"] = f0.numpy(), n.numpy()
    for k in ("gen_stage0", "gen_stage1", "xpost", "dec_out", "t_en", "en", "bert"):
        if k in tap:
            out[f"kokoro_rms_{k}"] = np.float64(torch.sqrt((tap[k].double() ** 2).mean()))


def snac(out):
    from mlx_audio_b200 import synth
    from oracle import codec as OC
    P64 = {k: v.double() for k, v in synth.snac_weights(OC.SNAC_24K).items()}
    codes = synth.snac_codes(OC.SNAC_24K, 2048, 1)
    noises = synth.snac_noises(OC.SNAC_24K, 1)
    y = OC.snac_decode(P64, codes, noises=[n.double() for n in noises]).reshape(-1)
    out["snac_len"] = np.int64(y.numel())
    out["snac_ms"] = np.float64((y ** 2).mean())
    for i, s in enumerate(windows(y.numel())):
        out[f"snac_win{i}"] = y[s:s + WIN].float().numpy()


def mimi(out):
    from mlx_audio_b200 import synth
    from oracle import codec as OC
    P64 = {k: v.double() for k, v in synth.mimi_weights(OC.MIMI_202407).items()}
    codes = synth.mimi_codes(OC.MIMI_202407, 2000, 1)
    y = OC.mimi_decode(P64, codes).reshape(-1)
    out["mimi_len"] = np.int64(y.numel())
    out["mimi_ms"] = np.float64((y ** 2).mean())
    for i, s in enumerate(windows(y.numel())):
        out[f"mimi_win{i}"] = y[s:s + WIN].float().numpy()


def whisper(out):
    from mlx_audio_b200 import synth
    from oracle import whisper as OW
    dims = dict(OW.WHISPER_SMALL)
    P64 = {k: v.double() for k, v in synth.whisper_decoder_weights(dims).items()}
    xa = torch.randn(2, 1500, 768, generator=torch.Generator().manual_seed(0))
    spec = OW.TokenizerSpec()
    tok0 = torch.tensor([list(spec.sot_sequence)] * 2)
    logits, _ = OW.decoder_forward(P64, tok0, xa.double(), None, dims)
    out["whisper_logits"] = logits[:, -1].float().numpy()
    tokens, lp, ns = OW.greedy_decode(P64, xa.double(), spec, sample_len=8, suppress=(11, 12), dims=dims)
    out["whisper_tokens"] = np.asarray(tokens, dtype=np.int64)
    out["whisper_sum_logprobs"] = lp.numpy()
    out["whisper_no_speech"] = ns.numpy()


def qwen3(out):
    from mlx_audio_b200 import synth
    from oracle import qwen3 as Q
    flat = dict(Q.TALKER)
    P = synth.qwen3_talker_weights(flat, seed=11)
    Pt = {k[len("talker."):]: v.double() for k, v in P.items()}
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 500, (12,), generator=g).tolist()
    cfg_ids = {"codec_nothink_id": 2155, "codec_think_id": 2154, "codec_think_bos_id": 2156, "codec_think_eos_id": 2157,
               "codec_pad_id": 2148, "codec_bos_id": 2149}
    ref = Q.prepare_generation_inputs_from_ids(Pt, ids, (501, 502, 500), cfg_ids, language_id=2050, speaker_id=2100)
    u = torch.rand(25, 16, generator=torch.Generator().manual_seed(2))
    trace = []
    want = Q.generate_codes(Pt, *ref, u.double(), 25, cfg=flat, trace=trace)
    out["qwen3_ids"] = np.asarray(ids, dtype=np.int64)
    out["qwen3_codes"] = want.numpy()
    out["qwen3_logits0"] = trace[0]["logits"].float().numpy()
    out["qwen3_cfg_ids"] = np.asarray([cfg_ids[k] for k in sorted(cfg_ids)], dtype=np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="kokoro,snac,mimi,whisper,qwen3")
    args = ap.parse_args()
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in args.only.split(","):
        t0 = time.time()
        globals()[name](out)
        print(f"{name}: {time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
