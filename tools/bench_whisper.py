#!/usr/bin/env python
"""BASELINE config 3: Whisper-small log-mel + encoder, batch 32 x 30 s, on one GPU -- audio-seconds transcribed-side
per second and achieved TFLOP/s of the encoder (11.0 TFLOP per batch of 32)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import ops, synth
from mlx_audio_b200.stt.models.whisper import Model, ModelDimensions
from mlx_audio_b200.configs import WHISPER_SMALL

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
model = Model(ModelDimensions.from_dict(WHISPER_SMALL), device=dev).load_weights(synth.whisper_encoder_weights(WHISPER_SMALL))
audio = synth.whisper_audio(args.batch).to(dev)
for _ in range(2):
    model.encode_audio(audio)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    y = model.encode_audio(audio)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
prof = {}
ops.PROFILE = prof
model.encode_audio(audio)
torch.cuda.synchronize()
ops.PROFILE = None
flops = 345e9 * args.batch
print(json.dumps({"workload": f"whisper-small mel+encoder batch {args.batch} x 30 s", "ms": ms, "audio_s_per_s": 30 * args.batch / (ms / 1e3),
                  "encoder_TFLOPs_achieved": flops / (ms / 1e3) / 1e12, "tc_mode": ops.TC_MODE[0],
                  "ms_by_kind": {k: round(sum(a.elapsed_time(b) for a, b in v), 2) for k, v in prof.items()}}))
