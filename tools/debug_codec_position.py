#!/usr/bin/env python
"""Position independence of the codec decoders (run on the GPU box): decode a stream whole and as prefix / spans, print the differences
under the dispatch toggles (ops.FUSED_DISPATCH, ops.TC_MODE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import synth, ops, configs as OC
from mlx_audio_b200.codec import SNAC, Mimi, mimi_202407

def run(tag):
    m = Mimi(mimi_202407(32), device="cuda:0").load_weights(synth.mimi_weights(OC.MIMI_202407))
    codes = synth.mimi_codes(OC.MIMI_202407, 63)
    y = m.decode(codes); y2 = m.decode(codes[:, :, :40])
    d = (y2 - y[:, :, :40 * 1920]).abs()
    print(tag, "mimi prefix: max abs diff %.3e at %d, ref max %.3e" % (float(d.max()), int(d.argmax()), float(y.abs().max())))
    s = SNAC.from_config(OC.SNAC_24K, device="cuda:0").load_weights(synth.snac_weights(OC.SNAC_24K))
    T = 236
    sc = synth.snac_codes(OC.SNAC_24K, T); nz = synth.snac_noises(OC.SNAC_24K)
    full = s.decode(sc, noises=nz)
    a = s.decode_span(sc, 0, 120, noises=nz); b = s.decode_span(sc, 120, T, noises=nz)
    got = torch.cat([a, b], dim=1)
    d = (got - full).abs()[0, :, 0]
    idx = torch.nonzero(d > 1e-5)[:, 0]
    print(tag, "snac span: max abs diff %.3e, ref max %.3e, bad samples %d, first %s last %s" % (float(d.max()), float(full.abs().max()), idx.numel(),
          idx[:3].tolist(), idx[-3:].tolist()))

run("default")
ops.FUSED_DISPATCH[0] = True
run("routed through the fused kernel")
ops.FUSED_DISPATCH[0] = False
ops.TC_MODE[0] = "off"
run("tensor cores off")
