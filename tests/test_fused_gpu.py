"""csrc/conv_fused.cu beyond what test_tc_gpu.py covers through ops.conv1d: the statistics prologue / epilogue (InstanceNorm + AdaIN from
the producer's (sum, sumsq)), grouped problems on one grid, the folded branch average (x_add / in_scale), split-K for small-M problems,
polyphase scatter with statistics, and inputs shorter than one tile.  Reference: float64 torch on the CPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nn as ON

DEV = "cuda:0"


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def rel_err(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _adain_snake_ref(x, gb, a, eps=1e-5):
    """InstanceNorm over L (biased variance) -> (1 + gamma) * xhat + beta -> Snake: float64."""
    x = x.double()
    C = x.shape[2]
    mean, var = x.mean(dim=1, keepdim=True), x.var(dim=1, unbiased=False, keepdim=True)
    v = (x - mean) / torch.sqrt(var + eps)
    if gb is not None:
        v = v * (1 + gb.double()[:, None, :C]) + gb.double()[:, None, C:]
    if a is not None:
        v = v + (1.0 / a.double()) * torch.sin(a.double() * v) ** 2
    return v


def _w(cout, k, cin, seed):
    return _rand(cout, k, cin, seed=seed, scale=0.05).to(torch.bfloat16).float()


@pytest.mark.parametrize("B,L,C,K,dil", [(1, 1000, 128, 7, 3), (2, 390, 256, 3, 1), (1, 7801, 128, 11, 5), (1, 45, 64, 3, 1)])
def test_statistics_chain(B, L, C, K, dil):
    """Two chained layers: layer 1 accumulates the (sum, sumsq) of its output; layer 2 derives AdaIN + Snake from them in its prologue.
    Equals InstanceNorm computed from the float64 reference of layer 1's output."""
    from mlx_audio_b200 import ops
    x = _rand(B, L, C, seed=1)
    w1, w2 = _w(C, K, C, 2), _w(C, 3, C, 3)
    b1 = _rand(C, seed=4, scale=0.1)
    gb = _rand(B, 2 * C, seed=5, scale=0.3)
    a = (1 + 0.2 * _rand(C, seed=6)).abs() + 0.1
    pad = (K - 1) * dil // 2
    y1_ref = ON.conv1d(x.double(), w1.double(), 1, pad, dil, 1, b1.double())
    y2_ref = ON.conv1d(_adain_snake_ref(y1_ref, gb, a), w2.double(), 1, 1, 1, 1) + x.double()
    cw1, cw2 = ops.pack_conv(w1, b1, 1, DEV), ops.pack_conv(w2, None, 1, DEV)
    st = ops.new_stats(B, C, DEV)
    xd = x.to(DEV)
    y1 = ops.conv_fused(ops.FusedProblem(xd, cw1, dilation=dil, pad_left=pad, stats_out=st))[0]
    torch.cuda.synchronize()
    assert rel_err(y1, y1_ref) < 2e-5
    y1d = y1.double()                                  # the statistics are those of what the kernel WROTE (its fp32 output), to accumulation accuracy
    want = torch.stack([y1d.sum(dim=1), (y1d ** 2).sum(dim=1)], dim=-1)
    assert rel_err(ops.stats_value(st), want) < 1e-6
    pre = ops.PreStats(st, gb.to(DEV), 1e-5, ops.ACT["snake"], 0.0, a.to(DEV), (1.0 / a).to(DEV))
    y2 = ops.conv_fused(ops.FusedProblem(y1, cw2, pad_left=1, pre=pre, res=xd))[0]
    torch.cuda.synchronize()
    assert rel_err(y2, y2_ref) < 5e-5
    # the stand-alone statistics kernel writes the same format (two destinations at once, one a slice of a wider buffer)
    s2, wide = ops.new_stats(B, C, DEV), torch.zeros(B, C + 6, 2, ops.STAT_BINS, device=DEV, dtype=torch.int64)
    ops.channel_stats(y1, [s2, wide[:, 3:3 + C]])
    assert rel_err(ops.stats_value(s2), want) < 1e-12 and torch.equal(wide[:, 3:3 + C], s2) and int(wide[:, :3].abs().max()) == 0
    # integer bins: the accumulation is order-independent, so a second pass over the same data lands on the same bits
    st_b = ops.new_stats(B, C, DEV)
    ops.conv_fused(ops.FusedProblem(xd, cw1, dilation=dil, pad_left=pad, stats_out=st_b))
    assert torch.equal(st_b, st)
    sc, sh = ops.coeffs_from_stats(s2, L, gb.to(DEV))
    sc_ref, sh_ref = ops.adain_coeffs(y1, gb.to(DEV))
    assert rel_err(sc, sc_ref) < 1e-5 and rel_err(sh, sh_ref) < 1e-5


def test_grouped_problems_and_folded_average():
    """Three resblock-style problems (kernel sizes 3 / 7 / 11, dilation 3) in ONE launch equal three separate launches bit-for-bit, and a
    consumer that reads (o0 + o1 + o2) / 3 through x_add / in_scale equals the conv of the explicit average."""
    from mlx_audio_b200 import ops
    L, C = 2000, 128
    x = _rand(1, L, C, seed=1).to(DEV)
    cws = [ops.pack_conv(_w(C, k, C, 10 + k), _rand(C, seed=20 + k, scale=0.1), 1, DEV) for k in (3, 7, 11)]
    sts = [ops.new_stats(1, C, DEV) for _ in range(3)]
    probs = [ops.FusedProblem(x, cw, dilation=3, pad_left=(cw.K - 1) * 3 // 2, pre=ops.Pre(act=ops.ACT["lrelu"], p0=0.1), res=x, stats_out=s)
             for cw, s in zip(cws, sts)]
    outs = ops.conv_fused(probs)
    singles = [ops.conv_fused(ops.FusedProblem(x, cw, dilation=3, pad_left=(cw.K - 1) * 3 // 2, pre=ops.Pre(act=ops.ACT["lrelu"], p0=0.1), res=x))[0]
               for cw in cws]
    torch.cuda.synchronize()
    for o, s_, cw, st in zip(outs, singles, cws, sts):
        assert torch.equal(o, s_)
        ref = ON.conv1d(ON.leaky_relu(x.double().cpu(), 0.1), cw.w.permute(2, 0, 1).double().cpu(), 1, (cw.K - 1) * 3 // 2, 3, 1, cw.bias.double().cpu()) + x.double().cpu()
        assert rel_err(o, ref) < 2e-5
        assert rel_err(ops.stats_value(st)[0, :, 0], o.double().sum(dim=1)[0]) < 1e-6
    post = ops.pack_conv(_w(32, 7, C, 40), _rand(32, seed=41, scale=0.1), 1, DEV)
    y = ops.conv_fused(ops.FusedProblem(outs[0], post, pad_left=3, pre=ops.Pre(act=ops.ACT["lrelu"], p0=0.01), x_add=(outs[1], outs[2]), in_scale=1.0 / 3))[0]
    avg = (outs[0].double() + outs[1].double() + outs[2].double()).cpu() / 3
    ref = ON.conv1d(ON.leaky_relu(avg, 0.01), post.w.permute(2, 0, 1).double().cpu(), 1, 3, 1, 1, post.bias.double().cpu())
    assert rel_err(y, ref) < 2e-5


def test_grouped_launch_splits_k_per_problem():
    """A Kokoro decoder block as it is launched: the k=3 conv over 390 rows grouped with its 1x1 shortcut (64 output tiles for 148 SMs), so
    every problem of the group splits K with its own counters and partial tiles in the shared workspace.  Equal to the float64 reference,
    repeatable bit for bit, statistics from the reducing CTAs only."""
    from mlx_audio_b200 import ops
    L, Cin, Cout = 390, 1090, 1024
    ld = -(-Cin // 4) * 4
    x = _rand(1, L, ld, seed=1).to(DEV)[:, :, :Cin]
    cw3 = ops.pack_conv(_w(Cout, 3, Cin, 2), _rand(Cout, seed=3, scale=0.1), 1, DEV)
    cw1 = ops.pack_conv(_w(Cout, 1, Cin, 4), None, 1, DEV)

    def run():
        sts = [ops.new_stats(1, Cout, DEV) for _ in range(2)]
        outs = ops.conv_fused([ops.FusedProblem(x, cw3, pad_left=1, stats_out=sts[0]), ops.FusedProblem(x, cw1, stats_out=sts[1])])
        torch.cuda.synchronize()
        return outs, sts

    (y3, y1), sts = run()
    (z3, z1), sts2 = run()
    assert torch.equal(y3, z3) and torch.equal(y1, z1) and torch.equal(sts[0], sts2[0]) and torch.equal(sts[1], sts2[1])
    xd = x.double().cpu()
    assert rel_err(y3, ON.conv1d(xd, cw3.w.permute(2, 0, 1).double().cpu(), 1, 1, 1, 1, cw3.bias.double().cpu())) < 2e-5
    assert rel_err(y1, ON.conv1d(xd, cw1.w.permute(2, 0, 1).double().cpu(), 1, 0, 1, 1)) < 2e-5
    for y, st in ((y3, sts[0]), (y1, sts[1])):
        sv = ops.stats_value(st)
        assert rel_err(sv[0, :, 0], y.double().sum(dim=1)[0]) < 1e-6 and rel_err(sv[0, :, 1], (y.double() ** 2).sum(dim=1)[0]) < 1e-6


@pytest.mark.parametrize("L,Cin,Cout,K", [(390, 1090, 1024, 3), (130, 768, 2304, 1), (390, 514, 1024, 3), (20, 512, 512, 5)])
def test_split_k_small_m(L, Cin, Cout, K):
    """Decoder-sized problems (4 M tiles) split K across CTAs; the fixed-order reduction of the partial tiles is deterministic and equals
    the unsplit result to fp32 summation-order accuracy; statistics come from the reducing CTA only.  Cin = 514 / 1090: the last
    64-channel chunk is partial and the row stride is padded to a multiple of 4 floats."""
    from mlx_audio_b200 import ops
    ld = -(-Cin // 4) * 4
    xbuf = _rand(1, L, ld, seed=1).to(DEV)
    x = xbuf[:, :, :Cin]
    cw = ops.pack_conv(_w(Cout, K, Cin, 2), _rand(Cout, seed=3, scale=0.1), 1, DEV)
    ref = ON.conv1d(x.double().cpu(), cw.w.permute(2, 0, 1).double().cpu(), 1, (K - 1) // 2, 1, 1, cw.bias.double().cpu())
    st = ops.new_stats(1, Cout, DEV)
    y = ops.conv_fused(ops.FusedProblem(x, cw, pad_left=(K - 1) // 2, stats_out=st))[0]
    y2 = ops.conv_fused(ops.FusedProblem(x, cw, pad_left=(K - 1) // 2))[0]
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 2e-5 and torch.equal(y, y2)
    sv = ops.stats_value(st)
    assert rel_err(sv[0, :, 0], y.double().sum(dim=1)[0]) < 1e-6 and rel_err(sv[0, :, 1], (y.double() ** 2).sum(dim=1)[0]) < 1e-6


@pytest.mark.parametrize("L,Cin,Cout,K,stride", [(780, 512, 256, 20, 10), (7800, 256, 128, 12, 6), (37, 64, 32, 4, 2)])
def test_polyphase_with_statistics(L, Cin, Cout, K, stride):
    """Generator up-sampler: LeakyReLU prologue, scatter epilogue with the harmonic-source residual, statistics of the up-sampled output
    (each tile lies inside one phase; its column sums land on the phase's output channels)."""
    from mlx_audio_b200 import ops
    x = _rand(1, L, Cin, seed=1)
    w = _w(Cout, K, Cin, 2)
    bias = _rand(Cout, seed=3, scale=0.1)
    pad = (K - stride) // 2
    ref = ON.conv_transpose1d(ON.leaky_relu(x.double(), 0.1), w.double(), stride, pad, 1, 0, 1, bias.double())
    res = _rand(*ref.shape, seed=5)
    ref = ref + res.double()
    cw = ops.pack_conv(w, bias, 1, DEV)
    st = ops.new_stats(1, Cout, DEV)
    y = ops.conv_fused(ops.FusedProblem(x.to(DEV), cw, stride=stride, pad_left=pad, pre=ops.Pre(act=ops.ACT["lrelu"], p0=0.1), transpose=True,
                                        res=res.to(DEV), stats_out=st))[0]
    torch.cuda.synchronize()
    assert y.shape == ref.shape and rel_err(y, ref) < 2e-5
    sv = ops.stats_value(st)
    assert rel_err(sv[0, :, 0], y.double().sum(dim=1)[0]) < 1e-6 and rel_err(sv[0, :, 1], (y.double() ** 2).sum(dim=1)[0]) < 1e-6


def test_many_launches_back_to_back_share_the_workspace():
    """Programmatic dependent launch + the self-resetting split-K counters: 30 dependent launches in a row, each reading the previous
    output, match the float64 chain."""
    from mlx_audio_b200 import ops
    L, C = 300, 256
    x = _rand(1, L, C, seed=1)
    cw = ops.pack_conv(_w(C, 3, C, 2), None, 1, DEV)
    ref = x.double()
    y = x.to(DEV)
    for _ in range(30):
        ref = torch.tanh(ON.conv1d(ref, cw.w.permute(2, 0, 1).double().cpu(), 1, 1, 1, 1))
        y = ops.conv_fused(ops.FusedProblem(y, cw, pad_left=1, post_act=ops.ACT["tanh"]))[0]
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-3          # 30 layers deep: fp32-grade products, error grows with depth
