#!/usr/bin/env python
"""Is the Kokoro forward bit-reproducible (eager vs eager, graph vs graph, graph vs eager), and if not, at which stage tap do the two paths
part?  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import ops, synth
from mlx_audio_b200.configs import KOKORO_82M
from mlx_audio_b200.tts.models.kokoro import Model, ModelConfig

dev = "cuda:0"
model = Model(ModelConfig.from_dict(KOKORO_82M), device=dev).load_weights(list(synth.kokoro_weights(KOKORO_82M).items()))
d = lambda a, b: float((a.float() - b.float()).abs().max())
cases = [(20, True), (20, False), (9, False)] if len(sys.argv) < 2 else [(int(sys.argv[1]), False)]
for n_ph, pinned in cases:
    ids, ref_s = synth.kokoro_inputs(n_ph, seed=3)
    T = ids.shape[1]
    dur = [2] * T if pinned else None
    F = 2 * T if pinned else 3 * T
    nz = ops.randn_(torch.empty(1, F * 600, 9, device=dev), 5, 0)
    model.tap = {}
    e1 = model.forward_ids(ids[0], ref_s, noise=nz, pred_dur=dur)[0].clone()
    te = model.tap
    model.tap = {}
    g1 = model.synthesize_ids(ids[0], ref_s, noise=nz, pred_dur=dur)[0].clone()      # the captures record the tap clones as graph nodes
    g2 = model.synthesize_ids(ids[0], ref_s, noise=nz, pred_dur=dur)[0].clone()
    tg = model.tap
    model.tap = None
    print(f"n_ph {n_ph} pinned {pinned}: graph-graph {d(g1, g2):.3e}  graph-eager {d(g1, e1):.3e}  scale {float(e1.abs().max()):.3f}")
    print("   taps graph-eager:", {k: f"{d(te[k], tg[k]):.1e}" for k in te if k in tg and te[k].shape == tg[k].shape})
