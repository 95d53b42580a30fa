"""Phase timing of conv_tc_kernel (clock64 stamps from CTA 0) for a few shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import ops, _lib
dev = torch.device("cuda:0")
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
ops.TC_MODE[0] = sys.argv[1] if len(sys.argv) > 1 else "x2"
for (L, Cin, Cout, K) in [(390, 1090, 1024, 3), (7800, 256, 256, 7), (46801, 128, 128, 11), (46801, 128, 128, 7), (46801, 128, 128, 3), (9600, 768, 768, 7), (48000, 768, 3072, 1)]:
    x = torch.randn(1, L, Cin, device=dev)
    w = (torch.randn(Cout, K, Cin) * 0.05).to(torch.bfloat16).float()
    cw = ops.pack_conv(w, torch.zeros(Cout), 1, dev)
    res = torch.randn(1, L, Cout, device=dev)
    for _ in range(3):
        y = ops.conv1d(x, cw, pad_left=(K - 1) // 2, res=res)
    torch.cuda.synchronize()
    _lib.lib().b2a_conv1d_tc_debug(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hi, lo = ops.prep_bf16(x, None, cw.cin_pad, 2 if ops.TC_MODE[0] == "x2" else 1)
    torch.cuda.synchronize()
    e0.record()
    y = ops.conv1d(x, cw, pad_left=(K - 1) // 2, res=res)
    e1.record()
    torch.cuda.synchronize()
    prof = {}
    ops.PROFILE = prof
    for _ in range(5):
        ops.conv1d(x, cw, pad_left=(K - 1) // 2, res=res)
    torch.cuda.synchronize()
    ops.PROFILE = None
    kern_us = {k: round(min(a.elapsed_time(b) for a, b in v) * 1e3, 1) for k, v in prof.items()}
    _lib.lib().b2a_conv1d_tc_debug(None)
    t = dbg.cpu().tolist()
    d = [(t[i] - t[0]) for i in range(7)]
    e = [t[i] - t[4] for i in (8, 9, 10, 11)]
    iters = K * cw.cin_pad // 64
    print(f"L={L} Cin={Cin} Cout={Cout} K={K} iters={iters} kernels(min of 5) {kern_us} us | cycles since entry: setup {d[1]} first_full {d[2]} last_full {d[3]} acc_ready {d[4]} epi_done {d[5]} | chunk0 since acc_ready: ldtm {e[0]} staged {e[1]} loads_issued {e[2]} stored {e[3]}")
