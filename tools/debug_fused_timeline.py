#!/usr/bin/env python
"""Phase timeline of csrc/conv_fused.cu for a few Kokoro-shaped problems (run on the GPU box): %globaltimer stamps per CTA.
slots: 0 entry, 1 setup done, 2 first W TMA issued, 3 converter past the dependency wait, 4 MMA saw first W, 5 last MMA committed (tile 0),
6 first A chunk ready, 7 last A chunk ready, 8 epilogue saw tile 0, 9 split-K arrival decided, 10 tile 0 epilogue done, 11 last tile epilogue done,
12 before final sync, 13 after."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import ops, _lib

dev = "cuda:0"
ops.TC_MODE[0] = os.environ.get("TC", "x2")
ONLY = os.environ.get("ONLY")
print("tensor-core mode", ops.TC_MODE[0])
def w(cout, k, cin, seed):
    return (torch.randn(cout, k, cin, generator=torch.Generator().manual_seed(seed)) * 0.05).to(torch.bfloat16).float()

def run(name, probs_fn, reps=3):
    if ONLY and ONLY not in name:
        return
    if os.environ.get("NODBG"):                  # un-instrumented: %globaltimer reads inside the wait loops distort the per-role numbers
        for _ in range(3):
            ops.conv_fused(probs_fn())
        probs = [probs_fn() for _ in range(20)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for pr in probs:
                ops.conv_fused(pr)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(f"== {name}: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us per launch (20 launches in one graph, no instrumentation)")
        return
    stamps = torch.zeros(148, 32, dtype=torch.int64, device=dev)
    for _ in range(reps):
        ops.conv_fused(probs_fn())
    torch.cuda.synchronize()
    _lib.lib().b2a_conv1d_fused_debug(stamps.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv_fused(probs_fn()); e1.record()
    torch.cuda.synchronize()
    _lib.lib().b2a_conv1d_fused_debug(None)
    s = stamps.cpu()
    live = s[:, 0] > 0
    t0 = int(s[live, 0].min())
    rel = (s - t0).float() / 1e3
    rel[s == 0] = float("nan")
    print(f"== {name}: event {e0.elapsed_time(e1)*1e3:.1f} us, CTAs {int(live.sum())}")
    for c in sorted(set([0, int(live.sum()) // 2, int(live.sum()) - 1])):
        print(f"  cta {c}: " + " ".join(f"{i}:{rel[c, i]:.1f}" for i in range(14)))
    cyc = s[live][:, 16:26].float().mean(dim=0) / 1965.0          # SM cycles -> us at the 1965 MHz the B200 holds under load
    print(f"  mean us per CTA (clock64): producer loop {cyc[5]:.1f} (waiting on empty {cyc[6]:.1f}); MMA loop {cyc[7]:.1f} (waiting: tempty {cyc[0]:.1f}, A {cyc[1]:.1f}, W {cyc[2]:.1f}); "
          f"workers: convert {cyc[8]:.1f} (waiting a_empty {cyc[3]:.1f}), epilogue {cyc[9]:.1f} (waiting tfull {cyc[4]:.1f})")
    last = torch.nan_to_num(rel[live][:, :14], nan=0.0).max(dim=1).values
    print(f"  CTA end times: min {float(last.min()):.1f} median {float(last.median()):.1f} max {float(last.max()):.1f} us")

x130 = torch.randn(1, 130, 2048, device=dev)
cw_ffn_out = ops.pack_conv(w(768, 1, 2048, 1), None, 1, dev)
run("albert ffn_out M=130 K=2048 N=768 (split-K)", lambda: [ops.FusedProblem(x130, cw_ffn_out)])
x768 = torch.randn(1, 130, 768, device=dev)
cw_qkv = ops.pack_conv(w(2304, 1, 768, 2), None, 1, dev)
run("albert qkv M=130 K=768 N=2304", lambda: [ops.FusedProblem(x768, cw_qkv)])
old = ops.FUSED_WS_BYTES
xs = torch.randn(1, 46801, 128, device=dev)
cws = [ops.pack_conv(w(128, k, 128, 10 + k), None, 1, dev) for k in (3, 7, 11)]
a = torch.ones(128, device=dev)
run("generator stage 1 group k=3,7,11 d=1 snake (L=46801, C=128)",
    lambda: [ops.FusedProblem(xs, cw, pad_left=(cw.K - 1) // 2, pre=ops.Pre(act=ops.ACT["snake"], a=a, b=a), res=xs) for cw in cws])
run("generator stage 1 single k=7", lambda: [ops.FusedProblem(xs, cws[1], pad_left=3, pre=ops.Pre(act=ops.ACT["snake"], a=a, b=a), res=xs)])
st = [ops.new_stats(1, 128, dev) for _ in range(3)]
sin = ops.new_stats(1, 128, dev); ops.channel_stats(xs, sin)
gb = torch.randn(1, 256, device=dev) * 0.1
run("generator stage 1 group + stats in/out", lambda: [ops.FusedProblem(xs, cw, pad_left=(cw.K - 1) // 2, pre=ops.PreStats(sin, gb, 1e-5, ops.ACT["snake"], 0.0, a, a), res=xs, stats_out=s_) for cw, s_ in zip(cws, st)])
xd = torch.randn(1, 390, 1092, device=dev)[:, :, :1090]
cwd = ops.pack_conv(w(1024, 3, 1090, 5), None, 1, dev)
run("decoder conv 390 x 1090 -> 1024 k3 (split-K)", lambda: [ops.FusedProblem(xd, cwd, pad_left=1)])

# ---- is a row-shifted (tap) A descriptor slower than an aligned one?  Same MMA count per tile, one tile per CTA, no activation, no residual
xa = torch.randn(1, 148 * 128, 1024, device=dev)
cwa = ops.pack_conv(w(128, 1, 1024, 31), None, 1, dev)
run("mma probe: k=1 (aligned A), 16 K chunks = 128 MMAs per tile", lambda: [ops.FusedProblem(xa, cwa)])
xb = torch.randn(1, 148 * 128, 320, device=dev)
cwb = ops.pack_conv(w(128, 3, 320, 32), None, 1, dev)
run("mma probe: k=3 (row-shifted A), 5 K chunks x 3 taps = 120 MMAs per tile", lambda: [ops.FusedProblem(xb, cwb, pad_left=1)])
cwc = ops.pack_conv(w(128, 3, 320, 33), None, 1, dev)
run("mma probe: k=3 dilation 8 (shifts -8, 0, +8: multiples of the 8-row swizzle atom)", lambda: [ops.FusedProblem(xb, cwc, pad_left=8, dilation=8)])
