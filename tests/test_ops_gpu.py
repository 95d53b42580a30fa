"""GPU parity of every C-ABI op against the CPU oracle (float64) on seeded inputs.
Tolerances: fp32 kernels vs float64 oracle -> max-abs error relative to the output scale <= 2e-5
(DSP: 1e-4 abs on log-mel); integer/index work is bit-exact."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dsp as OD
from oracle import nn as ON


def _dev():
    return torch.device("cuda:0")


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _oracle_pre(x, scale, shift, act, p0, a, b):
    v = x
    if scale is not None:
        v = v * scale[:, None, :] + shift[:, None, :]
    if act == "lrelu":
        v = ON.leaky_relu(v, p0)
    elif act == "snake":
        v = v + b * torch.sin(a * v) ** 2
    elif act == "elu":
        v = ON.elu(v)
    return v


CONV_CASES = [
    # (B, L, Cin, Cout, K, stride, dil, pad, pre_act, norm, post_act, res, accumulate)
    (1, 130, 80, 96, 3, 1, 1, 1, None, False, "gelu", False, False),
    (2, 77, 22, 40, 12, 6, 1, 3, None, False, None, False, False),
    (1, 200, 64, 64, 11, 1, 5, 25, "snake", True, None, True, False),
    (1, 95, 130, 70, 7, 1, 3, 9, "lrelu", True, None, True, True),
    (2, 50, 33, 1, 7, 1, 1, 3, "elu", False, "tanh", False, False),
    (1, 64, 1, 1, 3, 2, 1, 1, None, False, None, False, False),
    (1, 1, 128, 300, 1, 1, 1, 0, None, False, None, False, False),
    (3, 41, 48, 48, 5, 2, 2, 4, None, False, None, False, False),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_dense(case):
    from mlx_audio_b200 import ops
    B, L, Cin, Cout, K, stride, dil, pad, pre_act, norm, post_act, use_res, accum = case
    dev = _dev()
    x = _rand(B, L, Cin, seed=1)
    w = _rand(Cout, K, Cin, seed=2, scale=0.1)
    bias = _rand(Cout, seed=3, scale=0.1)
    scale = (1 + 0.3 * _rand(B, Cin, seed=4)) if norm else None
    shift = 0.2 * _rand(B, Cin, seed=5) if norm else None
    a = (1 + 0.2 * _rand(Cin, seed=6)).abs() + 0.1
    bb = 1.0 / a
    xin = _oracle_pre(x.double(), None if scale is None else scale.double(), None if shift is None else shift.double(), pre_act, 0.15, a.double(), bb.double())
    ref = ON.conv1d(xin, w.double(), stride, pad, dil, 1, bias.double())
    if post_act == "gelu":
        ref = ON.gelu(ref)
    elif post_act == "tanh":
        ref = torch.tanh(ref)
    res = _rand(*ref.shape, seed=7) if use_res else None
    if res is not None:
        ref = ref + res.double()
    ref = ref * 0.7
    y0 = _rand(*ref.shape, seed=8) if accum else None
    if accum:
        ref = ref + y0.double()
    cw = ops.pack_conv(w, bias, 1, dev)
    pre = None
    if pre_act or norm:
        pre = ops.Pre(None if scale is None else scale.to(dev).contiguous(), None if shift is None else shift.to(dev).contiguous(),
                      ops.ACT[pre_act or "none"], 0.15, a.to(dev), bb.to(dev))
    out = y0.to(dev).clone() if accum else None
    y = ops.conv1d(x.to(dev), cw, stride=stride, dilation=dil, pad_left=pad, pre=pre, post_act=ops.ACT[post_act or "none"],
                   res=None if res is None else res.to(dev), out_scale=0.7, out=out, accumulate=accum)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 2e-5


def test_conv1d_strided_views_and_edge_pad():
    """Channel-slice input/output views (the no-concat layout) and pad_mode=1 (edge, mimi conv.py:334-347)."""
    from mlx_audio_b200 import ops
    dev = _dev()
    big_in = _rand(2, 60, 100, seed=1).to(dev)
    big_out = torch.zeros(2, 30, 90, device=dev)
    w = _rand(32, 4, 40, seed=2, scale=0.1)
    cw = ops.pack_conv(w, None, 1, dev)
    x = big_in[:, :, 10:50]
    ops.conv1d(x, cw, stride=2, pad_left=2, lout=30, pad_mode=1, out=big_out[:, :, 5:37])
    xp = torch.nn.functional.pad(x.cpu().double().transpose(1, 2), (2, 2), mode="replicate").transpose(1, 2)
    ref = ON.conv1d(xp, w.double(), 2, 0, 1, 1)[:, :30]
    assert rel_err(big_out[:, :, 5:37], ref) < 2e-5
    assert float(big_out[:, :, :5].abs().max()) == 0 and float(big_out[:, :, 37:].abs().max()) == 0


@pytest.mark.parametrize("B,L,C,K,stride,dil,pad", [(1, 100, 96, 7, 1, 1, 3), (2, 64, 33, 7, 1, 9, 27), (1, 50, 8, 4, 2, 1, 1),
    (2, 700, 64, 7, 1, 9, 27), (1, 515, 33, 7, 1, 3, 9), (1, 300, 40, 5, 1, 2, 4), (1, 129, 512, 7, 1, 1, 3), (2, 1000, 128, 7, 1, 3, 9),
    (1, 400, 256, 7, 1, 9, 54), (1, 260, 1024, 7, 1, 1, 6)])
def test_conv1d_depthwise(B, L, C, K, stride, dil, pad):
    from mlx_audio_b200 import ops
    dev = _dev()
    x, w, bias = _rand(B, L, C, seed=1), _rand(C, K, 1, seed=2, scale=0.3), _rand(C, seed=3)
    a = (1 + 0.2 * _rand(C, seed=6)).abs() + 0.1
    xin = x.double() + (1.0 / (a.double() + 1e-9)) * torch.sin(a.double() * x.double()) ** 2     # SNAC snake
    ref = ON.conv1d(xin, w.double(), stride, pad, dil, C, bias.double())
    y = ops.conv1d(x.to(dev), ops.pack_conv(w, bias, C, dev), stride=stride, dilation=dil, pad_left=pad,
                   pre=ops.Pre(act=ops.ACT["snake"], a=a.to(dev), b=(1.0 / (a + 1e-9)).to(dev)))
    assert rel_err(y, ref) < 2e-5


@pytest.mark.parametrize("B,L,Cin,Cout,K,stride,pad,opad", [
    (1, 78, 96, 64, 20, 10, 5, 0), (1, 50, 64, 32, 12, 6, 3, 0), (2, 33, 40, 24, 16, 8, 4, 1), (1, 20, 16, 8, 4, 2, 1, 1),
    (1, 40, 32, 16, 8, 4, 0, 0), (1, 17, 24, 100, 10, 5, 3, 1)])
def test_convtr1d_dense(B, L, Cin, Cout, K, stride, pad, opad):
    from mlx_audio_b200 import ops
    dev = _dev()
    x, w, bias = _rand(B, L, Cin, seed=1), _rand(Cout, K, Cin, seed=2, scale=0.1), _rand(Cout, seed=3)
    ref = ON.conv_transpose1d(ON.leaky_relu(x.double(), 0.1), w.double(), stride, pad, 1, opad, 1, bias.double())
    y = ops.conv1d(x.to(dev), ops.pack_conv(w, bias, 1, dev), stride=stride, pad_left=pad, lout=ref.shape[1],
                   pre=ops.Pre(act=ops.ACT["lrelu"], p0=0.1), transpose=True)
    assert y.shape == ref.shape and rel_err(y, ref) < 2e-5


def test_convtr1d_depthwise_kokoro_pool_pin():
    """The reference's own golden (tts/tests/test_istftnet_fidelity.py:18-31) through the CUDA path."""
    from mlx_audio_b200 import ops
    dev = _dev()
    x = torch.tensor([1.0, 2, 3, 4]).reshape(1, 4, 1)
    w = torch.tensor([1.0, 2, 3]).reshape(1, 3, 1)
    y = ops.conv1d(x.to(dev), ops.pack_conv(w, None, 1, dev), stride=2, pad_left=1, lout=8, transpose=True)
    assert y.reshape(-1).cpu().tolist() == [2, 5, 4, 9, 6, 13, 8, 12]
    C = 37
    x, w = _rand(2, 29, C, seed=1), _rand(C, 3, 1, seed=2)
    ref = ON.conv_transpose1d(x.double(), w.double(), 2, 0, 1, 0, C)[:, 1:]
    y = ops.conv1d(x.to(dev), ops.pack_conv(w, None, C, dev), stride=2, pad_left=1, lout=58, transpose=True)
    assert rel_err(y, ref) < 2e-5
    w4 = _rand(C, 4, 1, seed=3)                                      # Mimi ConvTrUpsample1d k4 s2 causal: trim right k-s
    ref = ON.conv_transpose1d(x.double(), w4.double(), 2, 0, 1, 0, C)[:, :58]
    y = ops.conv1d(x.to(dev), ops.pack_conv(w4, None, C, dev), stride=2, pad_left=0, lout=58, transpose=True)
    assert rel_err(y, ref) < 2e-5


def test_adain_coeffs_and_layernorm():
    from mlx_audio_b200 import ops
    dev = _dev()
    x = (_rand(2, 1000, 70, seed=1) * 3 + 50).to(dev)              # large mean: E[x^2]-mean^2 must stay exact
    gb = _rand(2, 140, seed=2).to(dev)
    sc, sh = ops.adain_coeffs(x, gb)
    xd = x.double().cpu()
    mu, var = xd.mean(1), xd.var(1, unbiased=False)
    g, b = 1 + gb.double().cpu()[:, :70], gb.double().cpu()[:, 70:]
    ref_sc = g / torch.sqrt(var + 1e-5)
    assert rel_err(sc, ref_sc) < 1e-5 and rel_err(sh, b - ref_sc * mu) < 1e-5
    big = _rand(1, 300, 200, seed=3).to(dev)
    v = big[:, :, 20:90]
    sc2, sh2 = ops.adain_coeffs(v, None)
    assert rel_err(sc2, 1 / torch.sqrt(v.double().cpu().var(1, unbiased=False) + 1e-5)) < 1e-5
    # layernorm: affine, residual, AdaLN, RMS
    h = _rand(37, 768, seed=4).to(dev)
    r = _rand(37, 768, seed=5).to(dev)
    w, b = _rand(768, seed=6).to(dev), _rand(768, seed=7).to(dev)
    y = ops.layernorm(h, w, b, eps=1e-12, res=r)
    assert rel_err(y, ON.layer_norm((h + r).double().cpu(), w.double().cpu(), b.double().cpu(), 1e-12)) < 1e-5
    ada = _rand(1536, seed=8).to(dev)
    y = ops.layernorm(h, eps=1e-5, ada=ada)
    ref = (1 + ada.double().cpu()[:768]) * ON.layer_norm(h.double().cpu(), eps=1e-5) + ada.double().cpu()[768:]
    assert rel_err(y, ref) < 1e-5
    y = ops.layernorm(h, w, eps=1e-6, rms=True)
    assert rel_err(y, ON.rms_norm(h.double().cpu(), w.double().cpu(), 1e-6)) < 1e-5


@pytest.mark.parametrize("Tq,Tk,causal,window,H,Hkv", [(130, 130, False, 0, 12, 12), (300, 300, True, 0, 8, 8), (700, 700, True, 250, 8, 8),
                                                       (5, 77, False, 0, 4, 2)])
def test_attention(Tq, Tk, causal, window, H, Hkv):
    from mlx_audio_b200 import ops
    dev = _dev()
    B, D = 2, 64
    q, k, v = _rand(B, Tq, H * D, seed=1), _rand(B, Tk, Hkv * D, seed=2), _rand(B, Tk, Hkv * D, seed=3)
    qh = q.double().reshape(B, Tq, H, D).transpose(1, 2)
    kh = k.double().reshape(B, Tk, Hkv, D).transpose(1, 2)
    vh = v.double().reshape(B, Tk, Hkv, D).transpose(1, 2)
    mask = None
    if causal:
        i, j = torch.arange(Tq)[:, None], torch.arange(Tk)[None, :]
        ok = j <= i
        if window:
            ok = ok & (i - j < window)
        mask = torch.where(ok, 0.0, -1e9).double()
    ref = ON.sdpa(qh, kh, vh, 0.125, mask).transpose(1, 2).reshape(B, Tq, H * D)
    y = ops.attention(q.to(dev), k.to(dev), v.to(dev), n_heads=H, n_kv_heads=Hkv, scale=0.125, causal=causal, window=window)
    assert rel_err(y, ref) < 2e-5


def test_rope_traditional():
    from mlx_audio_b200 import ops
    x = _rand(2, 50, 8 * 64, seed=1)
    ref = ON.rope_traditional(x.double().reshape(2, 50, 8, 64).transpose(1, 2), 7, 10000.0).transpose(1, 2).reshape(2, 50, 512)
    y = ops.rope_(x.to(_dev()).clone(), 8, offset=7, base=10000.0, traditional=True)
    assert rel_err(y, ref) < 1e-6


def test_lstm_bidir_matches_reference_recurrence():
    from mlx_audio_b200 import ops
    from oracle import kokoro as OK
    dev = _dev()
    T, In, H = 57, 640, 256
    P = {}
    for di, d in enumerate(("forward", "backward")):          # fixed seeds (str hash() is salted per process: the old seeds changed run to run)
        P[f"l.Wx_{d}"] = _rand(4 * H, In, seed=11 + 2 * di, scale=0.05).double()
        P[f"l.Wh_{d}"] = _rand(4 * H, H, seed=12 + 2 * di, scale=0.08).double()
        P[f"l.bias_ih_{d}"] = _rand(4 * H, seed=3, scale=0.1).double()
        P[f"l.bias_hh_{d}"] = _rand(4 * H, seed=4, scale=0.1).double()
    x = _rand(2, T, In, seed=9)
    ref = OK.lstm_bi(P, "l", x.double())
    wx = torch.cat([P["l.Wx_forward"], P["l.Wx_backward"]], 0).float()
    b = torch.cat([P["l.bias_ih_forward"] + P["l.bias_hh_forward"], P["l.bias_ih_backward"] + P["l.bias_hh_backward"]], 0).float()
    wh = torch.stack([P["l.Wh_forward"], P["l.Wh_backward"]], 0).float().contiguous().to(dev)
    xproj = ops.conv1d(x.to(dev), ops.pack_linear(wx, b, dev))
    y = ops.lstm_bidir(xproj, wh)
    assert rel_err(y, ref) < 4e-5          # fp32 recurrence over 57 steps vs float64


def test_gather_copy_durations():
    from mlx_audio_b200 import ops
    dev = _dev()
    src = _rand(13, 40, seed=1).to(dev)
    idx = torch.tensor([0, 0, 5, 12, 12, 3], device=dev)
    assert torch.equal(ops.gather_rows(src, idx), src[idx])
    dst = torch.zeros(13, 100, device=dev)
    ops.copy2d(src, dst[:, 30:70])
    assert torch.equal(dst[:, 30:70], src) and float(dst[:, :30].abs().max()) == 0
    # duration head: round-half-even, clip [1,100], nan/inf handling (kokoro.py:140-147)
    d = torch.tensor([0.2, 0.5, 1.5, 2.5, 3.49, 250.0, float("nan"), float("inf"), -float("inf"), 7.0], device=dev)
    pred, ix, total = ops.durations_to_index(d, 1000, 1.0)
    exp = [1, 1, 2, 2, 3, 100, 1, 100, 1, 7]
    assert pred.cpu().tolist() == exp and int(total.item()) == sum(exp)
    assert torch.equal(ix[: sum(exp)].cpu(), torch.repeat_interleave(torch.arange(10), torch.tensor(exp)))
    pred2, ix2, tot2 = ops.durations_to_index(torch.tensor([3, 0, 2], device=dev), 10)
    assert ix2[:5].cpu().tolist() == [0, 0, 0, 2, 2] and int(tot2.item()) == 5


@pytest.mark.parametrize("n_fft,hop,pad_mode,n", [(400, 160, 1, 16000), (1024, 256, 0, 12768), (20, 5, 1, 3000), (800, 200, 2, 5000)])
def test_stft(n_fft, hop, pad_mode, n):
    from mlx_audio_b200 import ops
    dev = _dev()
    x = _rand(2, n, seed=1)
    w = torch.as_tensor(OD.hanning(n_fft, periodic=(n_fft == 20)))
    modes = {0: dict(center=False), 1: dict(center=True, pad_mode="reflect"), 2: dict(center=True, pad_mode="constant")}
    ref = np.stack([OD.stft(r.numpy(), n_fft=n_fft, hop_length=hop, window=w.numpy(), **modes[pad_mode]) for r in x])
    re, im = ops.stft(x.to(dev), w.float().to(dev), n_fft, hop, pad_mode, ref.shape[1])
    scale = np.abs(ref).max()
    assert float((re.cpu().double().numpy() - ref.real).__abs__().max()) / scale < 2e-6
    assert float((im.cpu().double().numpy() - ref.imag).__abs__().max()) / scale < 2e-6


@pytest.mark.parametrize("n,padding", [(16000, 0), (16000, 480000), (48000, 480000)])
def test_whisper_logmel_config1(n, padding):
    """BASELINE config 1 (1 s 440 Hz sine, with and without the +30 s zero pad) and noise; tol 1e-4 abs."""
    from mlx_audio_b200 import ops
    dev = _dev()
    sine = np.sin(2 * np.pi * 440 * np.arange(n) / 16000).astype(np.float32)
    noise = (0.1 * _rand(n, seed=4)).numpy()
    x = torch.as_tensor(np.stack([sine, noise]))
    ref = np.stack([OD.whisper_log_mel(r, 80, padding) for r in x.numpy()])
    filt = torch.as_tensor(OD.mel_filters(16000, 400, 80, norm="slaney", mel_scale=None)).to(dev)
    win = torch.as_tensor(OD.hanning(400)).float().to(dev)
    y = ops.whisper_logmel(x.to(dev), padding, win, filt, ref.shape[1])
    assert y.shape == ref.shape
    assert float(np.abs(y.cpu().numpy() - ref).max()) < 1e-4


def test_istft_generic_matches_both_reference_variants():
    from mlx_audio_b200 import ops
    dev = _dev()
    n_fft, hop, T = 64, 16, 40
    re, im = _rand(2, 33, T, seed=1), _rand(2, 33, T, seed=2)
    w = OD.hanning(n_fft, periodic=True)
    wt = torch.as_tensor(w).float().to(dev)
    for norm_sq in (True, False):
        ref = np.stack([OD.istft(re[b].numpy() + 1j * im[b].numpy(), hop_length=hop, win_length=n_fft, window=w, normalized=norm_sq) for b in range(2)])
        y = ops.istft(re.to(dev), im.to(dev), n_fft, hop, wt, norm_sq=norm_sq, clamp_mode=0, trim=n_fft // 2, out_len=ref.shape[1])
        assert float(np.abs(y.cpu().numpy() - ref).max()) / np.abs(ref).max() < 1e-5
    ref = OD.istft_cache(re.numpy(), im.numpy(), n_fft, hop, n_fft, w, center=True)
    y = ops.istft(re.to(dev), im.to(dev), n_fft, hop, wt, norm_sq=True, clamp_mode=1, trim=n_fft // 2, out_len=ref.shape[1])
    assert float(np.abs(y.cpu().numpy() - ref)[:, : -n_fft].max()) / np.abs(ref[:, :-n_fft]).max() < 1e-5


@pytest.mark.parametrize("nF", [46, 14, 30, 124])
def test_kokoro_source_and_istft_head(nF):
    """nF = 14, 30, 124: the reference's ceil(float(L) * float(1/300)) down-sample length is nF + 1 there (parity quirk)."""
    from mlx_audio_b200 import ops
    from oracle import kokoro as OK
    dev = _dev()
    nv = nF - nF // 3
    f0 = torch.cat([torch.zeros(nF // 6), 80 + 300 * torch.rand(nv, generator=torch.Generator().manual_seed(1)), torch.zeros(nF - nv - nF // 6)])[None]
    f0 = torch.cat([f0, f0.flip(1) * 0.5], 0)                                                # B=2, voiced + unvoiced spans
    noise = _rand(2, nF * 300, 9, seed=2)
    lw, lb = _rand(1, 9, seed=3, scale=0.3), _rand(1, seed=4, scale=0.1)
    f0s = torch.repeat_interleave(f0.double()[:, :, None], 300, dim=1)
    sw, _, _ = OK.sinegen(f0s, rand_ini=torch.rand(2, 9).numpy(), noise=noise.double())
    src = torch.tanh(sw @ lw.double().T + lb.double())[:, :, 0]
    mag, ph = OK.mlxstft_transform(src.numpy(), 20, 5, 20)
    har = ops.kokoro_source(f0.to(dev), noise.to(dev), lw.reshape(-1).to(dev), lb.to(dev)).cpu().numpy()
    assert har.shape == (2, nF * 60 + 1, 22)
    assert np.abs(har[:, :, :11] - mag.transpose(0, 2, 1)).max() < 2e-6
    dphi = np.angle(np.exp(1j * (har[:, :, 11:] - ph.transpose(0, 2, 1))))                    # compare modulo 2 pi
    assert np.abs(dphi * (mag.transpose(0, 2, 1) > 1e-4)).max() < 1e-3
    # head
    x = _rand(2, 300, 22, seed=5, scale=0.7)
    spec, phase = np.exp(x[:, :, :11].double().numpy()).transpose(0, 2, 1), np.sin(x[:, :, 11:].double().numpy()).transpose(0, 2, 1)
    ref = OK.mlxstft_inverse(spec, phase, 20, 5, 20)[:, 0]
    y = ops.kokoro_istft_head(x.to(dev)).cpu().numpy()
    assert y.shape == ref.shape and np.abs(y - ref).max() / np.abs(ref).max() < 1e-5


def test_rvq_decode_bit_exact_and_bounds():
    from mlx_audio_b200 import ops
    dev = _dev()
    cb = _rand(5, 2048, 256, seed=1).to(dev)
    codes = torch.randint(0, 2048, (2, 5, 333), generator=torch.Generator().manual_seed(2)).to(dev)
    y = ops.rvq_decode(codes, cb)
    ref = sum(cb[q].double()[codes[:, q]] for q in range(5))
    assert rel_err(y, ref.cpu()) < 1e-6
    one = ops.rvq_decode(codes[:, :1].contiguous(), cb[:1].contiguous())
    assert torch.equal(one, cb[0][codes[:, 0]])                                              # a single gather is bit-exact
    codes[1, 2, 7] = 2048
    with pytest.raises(ValueError):
        ops.rvq_decode(codes, cb)


@pytest.mark.parametrize("Tq,Tk,causal,window,q_offset", [(130, 130, False, 0, 0), (1500, 1500, False, 0, 0), (200, 333, False, 0, 0),
                                                          (325, 325, True, 0, 0), (700, 700, True, 250, 0), (64, 200, True, 0, 136),
                                                          (1, 70, True, 0, 69)])
@pytest.mark.parametrize("mode", ["tc", "cuda"])
def test_attention_tensor_core_vs_cuda_core(Tq, Tk, causal, window, q_offset, mode):
    """b2a_attention_tc (tcgen05, fp16 hi/lo planes) and b2a_attention (CUDA cores) against float64 SDPA on strided q/k/v views
    of one fused qkv buffer, ragged tile tails, causal offset (decoder prefill with a cache) and Mimi's sliding window."""
    from mlx_audio_b200 import ops
    dev = _dev()
    B, H, D = 2, 3, 64
    T = max(Tq, Tk)
    qkv = _rand(B, T, 3 * H * D, seed=5, scale=1.5)
    q, k, v = qkv[:, :Tq, :H * D], qkv[:, :Tk, H * D:2 * H * D], qkv[:, :Tk, 2 * H * D:]
    qh, kh, vh = (t.double().reshape(B, -1, H, D).transpose(1, 2) for t in (q, k, v))
    mask = None
    if causal:
        i, j = torch.arange(Tq)[:, None] + q_offset, torch.arange(Tk)[None, :]
        ok = j <= i
        if window:
            ok = ok & (i - j < window)
        mask = torch.where(ok, 0.0, -1e9).double()
    ref = ON.sdpa(qh, kh, vh, 0.125, mask).transpose(1, 2).reshape(B, Tq, H * D)
    g = qkv.to(dev)
    old = ops.ATTN_MODE[0]
    ops.ATTN_MODE[0] = mode
    try:
        y = ops.attention(g[:, :Tq, :H * D], g[:, :Tk, H * D:2 * H * D], g[:, :Tk, 2 * H * D:], n_heads=H, scale=0.125, causal=causal,
                          window=window, q_offset=q_offset)
    finally:
        ops.ATTN_MODE[0] = old
    assert rel_err(y, ref) < 2e-5


@pytest.mark.parametrize("B,L,Cin,Cout,K,dil,pad_mode", [(2, 1000, 64, 1, 7, 1, 0), (1, 777, 96, 1, 7, 1, 0), (1, 300, 64, 1, 3, 1, 1), (1, 513, 40, 3, 5, 2, 0)])
def test_conv1d_narrow_output(B, L, Cin, Cout, K, dil, pad_mode):
    """Waveform heads (Mimi / SNAC / Qwen3 vocoder: C -> 1): the narrow-output kernel with Snake / ELU prologue, bias, clip epilogue."""
    from mlx_audio_b200 import ops
    dev = _dev()
    x, w, bias = _rand(B, L, Cin, seed=1), _rand(Cout, K, Cin, seed=2, scale=0.1), _rand(Cout, seed=3)
    a = (1 + 0.2 * _rand(Cin, seed=6)).abs() + 0.1
    xin = x.double() + (1.0 / (a.double() + 1e-9)) * torch.sin(a.double() * x.double()) ** 2
    padl = (K - 1) * dil                                               # causal
    if pad_mode == 1:
        xp = torch.nn.functional.pad(xin.transpose(1, 2), (padl, 0), mode="replicate").transpose(1, 2)
    else:
        xp = torch.nn.functional.pad(xin, (0, 0, padl, 0))
    ref = torch.clamp(ON.conv1d(xp, w.double(), 1, 0, dil, 1, bias.double()), -1.0, 1.0)
    y = ops.conv1d(x.to(dev), ops.pack_conv(w, bias, 1, dev), dilation=dil, pad_left=padl, lout=L, pad_mode=pad_mode,
                   pre=ops.Pre(act=ops.ACT["snake"], a=a.to(dev), b=(1.0 / (a + 1e-9)).to(dev)), post_act=ops.ACT["clip1"])
    assert y.shape == ref.shape and rel_err(y, ref) < 2e-5


def test_depthwise_conv_emits_next_layers_planes():
    """conv1d(..., emit=Pre): the vectorised depthwise kernel writes Snake(t) as the next tensor-core layer's bf16 hi/lo planes;
    they must equal what the separate prologue pass makes from the fp32 output, and the 1x1 conv fed with them the two-kernel result."""
    from mlx_audio_b200 import ops
    dev = _dev()
    B, L, C, K, d = 2, 700, 128, 7, 3
    x, w, bias = _rand(B, L, C, seed=1).to(dev), _rand(C, K, 1, seed=2, scale=0.3), _rand(C, seed=3)
    a1 = ((1 + 0.2 * _rand(C, seed=6)).abs() + 0.1).to(dev)
    a2 = ((1 + 0.2 * _rand(C, seed=7)).abs() + 0.1).to(dev)
    s1, s2 = ops.Pre(act=ops.ACT["snake"], a=a1, b=1.0 / (a1 + 1e-9)), ops.Pre(act=ops.ACT["snake"], a=a2, b=1.0 / (a2 + 1e-9))
    dw = ops.pack_conv(w, bias, C, dev)
    pw = ops.pack_conv(_rand(C, 1, C, seed=8, scale=0.1), _rand(C, seed=9), 1, dev)          # fp32 weights: split planes, as in SNAC
    assert ops.emit_eligible(dw, x, L, dilation=d)
    t = ops.conv1d(x, dw, dilation=d, pad_left=3 * d, pre=s1)
    hi, lo = ops.prep_bf16(t, s2, C, 2, False)
    pl = ops.conv1d(x, dw, dilation=d, pad_left=3 * d, pre=s1, emit=s2)
    assert torch.equal(pl.hi, hi) and torch.equal(pl.lo, lo)
    y_ref = ops.conv1d(t, pw, pre=s2, res=x)
    y = ops.conv1d(pl, pw, res=x)
    assert torch.equal(y, y_ref)
