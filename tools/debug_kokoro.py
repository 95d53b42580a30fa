"""Diagnostic: where does the Kokoro generator deviate from the oracle?  (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlx_audio_b200 import synth
from mlx_audio_b200.tts.models.kokoro import Model, ModelConfig
from oracle import kokoro as OK

cfg = OK.KOKORO_CONFIG
P = synth.kokoro_weights(cfg, seed=0)
model = Model(ModelConfig.from_dict(cfg), device="cuda:0").load_weights(list(P.items()))
P64 = {k: v.double() for k, v in P.items()}
ids, ref_s = synth.kokoro_inputs(16, seed=1)
T = ids.shape[1]; pd = [3] * T; F = 3 * T
nz = synth.kokoro_noise(F * 600, 3)[1]
OK.TAP = {}
OK.forward(P64, ids, ref_s.double(), noise=nz.double(), pred_dur_override=pd)
f0n = (OK.TAP["F0"].float().reshape(-1), OK.TAP["N"].float().reshape(-1))
print("F0 stats", float(f0n[0].min()), float(f0n[0].max()), float(f0n[0].mean()), "voiced frames", int((f0n[0] > 10).sum()), "of", f0n[0].numel())
OK.TAP = {}
ref, _ = OK.forward(P64, ids, ref_s.double(), noise=nz.double(), pred_dur_override=pd, f0n_override=f0n)
tr = OK.TAP; OK.TAP = None
model.tap = {}
audio, _ = model.forward_ids(ids[0], ref_s, noise=nz.cuda().contiguous(), pred_dur=pd, f0n_override=f0n)
tp = model.tap
h, hr = tp["har"][0].double().cpu().numpy(), tr["har"][0].double().numpy()
print("har shapes", h.shape, hr.shape)
d = h - hr
for c in range(22):
    fl = np.abs(d[:, c]) > 1.0
    print(f"ch{c:2d} {'mag' if c < 11 else 'ph '} maxerr {np.abs(d[:, c]).max():.3e} flips {int(fl.sum())} at {np.nonzero(fl)[0][:8].tolist()} rms_noflip {np.sqrt((np.where(fl, 0, d[:, c])**2).mean()):.3e} ref_rms {np.sqrt((hr[:, c]**2).mean()):.3e}")
for k in ("gen_stage0", "gen_stage1", "xpost"):
    a, b = tp[k][0].double().cpu().numpy(), tr[k][0].double().numpy()
    e = np.abs(a - b)
    rowe = np.sqrt((e ** 2).mean(1)) / np.sqrt((b ** 2).mean())
    print(k, "rel rms", np.sqrt((e**2).mean()) / np.sqrt((b**2).mean()), "worst rows", np.argsort(-rowe)[:6].tolist(), "row err head/mid/tail", rowe[:3].round(4).tolist(), rowe[len(rowe)//2:len(rowe)//2+3].round(6).tolist(), rowe[-3:].round(4).tolist())
a, b = audio.double().cpu().numpy(), ref.numpy()
e = np.abs(a - b)
print("audio rel rms", np.sqrt((e**2).mean()) / np.sqrt((b**2).mean()), "interior (drop 2000 each side)", np.sqrt((e[2000:-2000]**2).mean()) / np.sqrt((b[2000:-2000]**2).mean()))
