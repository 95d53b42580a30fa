"""Oracle for Kokoro-82M ``Model.__call__`` (phoneme ids + style -> waveform).

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-CPU, dtype of the weights dict
(float64 for checking, float32 for the timed CPU baseline).  Follows
``tts/models/kokoro/{kokoro.py:111-177, modules.py, istftnet.py}`` of the reference;
parameter names are the reference's MLX parameter tree (post-``sanitize``), so the same
weights dict drives this oracle and the CUDA product.

Randomness (MLX PRNG, not reproducible outside MLX) is injected: ``rand_ini`` [B,9]
(istftnet.py:581) and ``noise`` [B, 600F, 9] (istftnet.py:649).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import dsp as D
from . import nn as N

TAP = None   # set to a dict to capture intermediates (tests only)
EFFECTIVE_WEIGHTS_BF16 = True   # weight_norm() result rounded to the checkpoint dtype; False = exact arithmetic


def _tap(name, value):
    if TAP is not None:
        TAP[name] = value

# ----------------------------------------------------------------------------- config

KOKORO_CONFIG = {   # tts/tests/test_models.py:143-173 (the public Kokoro-82M config)
    "istftnet": {
        "upsample_kernel_sizes": [20, 12], "upsample_rates": [10, 6], "gen_istft_hop_size": 5,
        "gen_istft_n_fft": 20, "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        "resblock_kernel_sizes": [3, 7, 11], "upsample_initial_channel": 512,
    },
    "dim_in": 64, "dropout": 0.2, "hidden_dim": 512, "max_conv_dim": 512, "max_dur": 50,
    "multispeaker": True, "n_layer": 3, "n_mels": 80, "n_token": 178, "style_dim": 128,
    "text_encoder_kernel_size": 5,
    "plbert": {"hidden_size": 768, "num_attention_heads": 12, "intermediate_size": 2048,
               "max_position_embeddings": 512, "num_hidden_layers": 12, "dropout": 0.1},
}

# ----------------------------------------------------------------------------- weight helpers


def weight_norm(v, g):
    """istftnet.py:53-93 with dim=0: w = g * v / (||v||_{axes != 0} + 1e-7).

    The reference evaluates this in the checkpoint dtype (bf16) every forward, so the
    effective weight is a bf16 tensor; we evaluate in float32 and round once to bf16
    (DESIGN.md "effective weights")."""
    if not EFFECTIVE_WEIGHTS_BF16:                  # algorithm pin against the reference code run in float64 (tests/golden)
        nrm = torch.sqrt((v * v).sum(dim=tuple(range(1, v.ndim)), keepdim=True))
        return (v / (nrm + 1e-7)) * g
    v32, g32 = v.to(torch.float32), g.to(torch.float32)
    nrm = torch.sqrt((v32 * v32).sum(dim=tuple(range(1, v.ndim)), keepdim=True))
    w = (v32 / (nrm + 1e-7)) * g32
    return w.to(torch.bfloat16).to(v.dtype)


def conv_weighted(P, pre, x, *, transpose=False, stride=1, padding=1, dilation=1, groups=1):
    """ConvWeighted.__call__ (istftnet.py:128-170) on NLC ``x``.

    Weight orientation rule (istftnet.py:159-166): use the weight as is when
    x.shape[-1] == weight.shape[-1] or groups > 1, otherwise ``weight.T`` (full axis reversal).
    """
    w = weight_norm(P[pre + ".weight_v"], P[pre + ".weight_g"])
    if not (x.shape[-1] == w.shape[-1] or groups > 1):
        w = w.permute(2, 1, 0)
    bias = P.get(pre + ".bias")
    fn = N.conv_transpose1d if transpose else N.conv1d
    y = fn(x, w, stride=stride, padding=padding, dilation=dilation, groups=groups)
    return y if bias is None else y + bias.to(y.dtype)

# ----------------------------------------------------------------------------- LSTM / AdaLN


def lstm_bi(P, pre, x):
    """modules.py:93-285: bidirectional LSTM, gate order i,f,g,o, both biases summed; x [B,T,In]."""
    outs = []
    for d in ("forward", "backward"):
        wx, wh = P[f"{pre}.Wx_{d}"].to(x.dtype), P[f"{pre}.Wh_{d}"].to(x.dtype)
        b = (P[f"{pre}.bias_ih_{d}"] + P[f"{pre}.bias_hh_{d}"]).to(x.dtype)
        xp = x @ wx.T + b
        hdim = wh.shape[1]
        h = x.new_zeros(x.shape[0], hdim)
        c = x.new_zeros(x.shape[0], hdim)
        seq = range(x.shape[1]) if d == "forward" else range(x.shape[1] - 1, -1, -1)
        hs = [None] * x.shape[1]
        for t in seq:
            i, f, g, o = torch.split(xp[:, t] + h @ wh.T, hdim, dim=-1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            hs[t] = h
        outs.append(torch.stack(hs, dim=1))
    return torch.cat(outs, dim=-1)


def ada_layer_norm(P, pre, x, s, eps=1e-5):
    """modules.py:71-90: x [B,T,C], s [B,style]; (1+gamma)*LN(x)+beta."""
    h = N.linear(s, P[pre + ".fc.weight"], P[pre + ".fc.bias"])
    c = x.shape[-1]
    gamma, beta = h[:, None, :c], h[:, None, c:]
    return (1 + gamma) * N.layer_norm(x, eps=eps) + beta

# ----------------------------------------------------------------------------- ALBERT


def albert(P, input_ids, attention_mask, cfg):
    """modules.py:434-645 (CustomAlbert): returns sequence output [B,T,768]."""
    pre = "bert."
    t = input_ids.shape[1]
    dt = P[pre + "embeddings.word_embeddings.weight"].dtype
    e = (P[pre + "embeddings.word_embeddings.weight"][input_ids]
         + P[pre + "embeddings.position_embeddings.weight"][torch.arange(t)][None]
         + P[pre + "embeddings.token_type_embeddings.weight"][torch.zeros_like(input_ids)])
    e = N.layer_norm(e, P[pre + "embeddings.LayerNorm.weight"], P[pre + "embeddings.LayerNorm.bias"], 1e-12)
    mask = (1.0 - attention_mask[:, None, None, :].to(dt)) * -10000.0
    h = N.linear(e, P[pre + "encoder.embedding_hidden_mapping_in.weight"], P[pre + "encoder.embedding_hidden_mapping_in.bias"])
    L = pre + "encoder.albert_layer_groups.0.albert_layers.0."
    nh = cfg["num_attention_heads"]
    hd = cfg["hidden_size"] // nh
    for _ in range(cfg["num_hidden_layers"]):
        b = h.shape[0]
        q = N.linear(h, P[L + "attention.query.weight"], P[L + "attention.query.bias"]).reshape(b, t, nh, hd).transpose(1, 2)
        k = N.linear(h, P[L + "attention.key.weight"], P[L + "attention.key.bias"]).reshape(b, t, nh, hd).transpose(1, 2)
        v = N.linear(h, P[L + "attention.value.weight"], P[L + "attention.value.bias"]).reshape(b, t, nh, hd).transpose(1, 2)
        sc = q @ k.transpose(-1, -2) / math.sqrt(hd) + mask
        ctx = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(b, t, nh * hd)
        a = N.linear(ctx, P[L + "attention.dense.weight"], P[L + "attention.dense.bias"])
        a = N.layer_norm(a + h, P[L + "attention.LayerNorm.weight"], P[L + "attention.LayerNorm.bias"], 1e-12)
        f = N.gelu(N.linear(a, P[L + "ffn.weight"], P[L + "ffn.bias"]))
        f = N.linear(f, P[L + "ffn_output.weight"], P[L + "ffn_output.bias"])
        h = N.layer_norm(f + a, P[L + "full_layer_layer_norm.weight"], P[L + "full_layer_layer_norm.bias"], 1e-12)
    return h

# ----------------------------------------------------------------------------- AdaIN blocks


def instance_norm(x, eps=1e-5):
    """istftnet.py:216-268 on NCL: per-(batch, channel) stats over L, biased variance."""
    mu = x.mean(dim=2, keepdim=True)
    var = x.var(dim=2, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def adain(P, pre, x, s):
    """istftnet.py:327-338 on NCL x, s [B,style]."""
    h = N.linear(s, P[pre + ".fc.weight"], P[pre + ".fc.bias"])[:, :, None]
    c = x.shape[1]
    return (1 + h[:, :c]) * instance_norm(x) + h[:, c:]


def adain_resblk1d(P, pre, x, s, upsample=False):
    """istftnet.py:853-933 on NCL x: (residual + shortcut) / sqrt(2)."""
    learned_sc = (pre + ".conv1x1.weight_v") in P
    # shortcut (istftnet.py:893-902)
    sc = x
    if upsample:
        sc = torch.repeat_interleave(sc, 2, dim=2)           # nn.Upsample(2, nearest)
    if learned_sc:
        sc = conv_weighted(P, pre + ".conv1x1", sc.transpose(1, 2), padding=0).transpose(1, 2)
    # residual (istftnet.py:904-928)
    r = N.leaky_relu(adain(P, pre + ".norm1", x, s), 0.2)
    r = r.transpose(1, 2)
    if upsample:
        r = conv_weighted(P, pre + ".pool", r, transpose=True, stride=2, padding=0, groups=r.shape[-1])[:, 1:, :]
    r = conv_weighted(P, pre + ".conv1", r, padding=1).transpose(1, 2)
    r = N.leaky_relu(adain(P, pre + ".norm2", r, s), 0.2)
    r = conv_weighted(P, pre + ".conv2", r.transpose(1, 2), padding=1).transpose(1, 2)
    return (r + sc) / math.sqrt(2.0)


def snake_kokoro(x, alpha):
    """istftnet.py:382,389: x + (1/a) * sin(a x)^2 (no epsilon)."""
    a = alpha.to(x.dtype)
    return x + (1.0 / a) * torch.sin(a * x) ** 2


def adain_resblock1(P, pre, x, s, kernel, dilations=(1, 3, 5)):
    """istftnet.py:341-396 on NCL x."""
    for j, d in enumerate(dilations):
        xt = snake_kokoro(adain(P, f"{pre}.adain1.{j}", x, s), P[f"{pre}.alpha1.{j}"])
        xt = conv_weighted(P, f"{pre}.convs1.{j}", xt.transpose(1, 2), padding=(kernel * d - d) // 2, dilation=d).transpose(1, 2)
        xt = snake_kokoro(adain(P, f"{pre}.adain2.{j}", xt, s), P[f"{pre}.alpha2.{j}"])
        xt = conv_weighted(P, f"{pre}.convs2.{j}", xt.transpose(1, 2), padding=(kernel - 1) // 2).transpose(1, 2)
        x = xt + x
    return x

# ----------------------------------------------------------------------------- source + STFT


def sinegen(f0, upsample_scale=300, harmonic_num=8, sine_amp=0.1, noise_std=0.003,
            voiced_threshold=10.0, sampling_rate=24000, rand_ini=None, noise=None):
    """istftnet.py:548-652.  f0 [B, L, 1]; rand_ini [B, H]; noise [B, L, H] (injected N(0,1))."""
    was_np = not isinstance(f0, torch.Tensor)
    f0 = torch.as_tensor(np.asarray(f0)) if was_np else f0
    dt = f0.dtype
    h = harmonic_num + 1
    fn = f0 * torch.arange(1, h + 1, dtype=dt)[None, None, :]
    rad = (fn / sampling_rate) % 1
    ri = torch.as_tensor(np.asarray(rand_ini), dtype=dt).clone() if rand_ini is not None else torch.zeros(f0.shape[0], h, dtype=dt)
    ri[:, 0] = 0
    rad = rad.clone()
    rad[:, 0, :] = rad[:, 0, :] + ri
    rad = torch.as_tensor(D.interpolate(rad.transpose(1, 2).numpy(), scale_factor=1 / upsample_scale, mode="linear")).to(dt).transpose(1, 2)
    phase = torch.cumsum(rad, dim=1) * 2 * math.pi
    phase = torch.as_tensor(D.interpolate((phase.transpose(1, 2) * upsample_scale).numpy(), scale_factor=upsample_scale, mode="linear")).to(dt).transpose(1, 2)
    sines = torch.sin(phase) * sine_amp
    tl = f0.shape[1]                                            # _match_f0_length, istftnet.py:624-632
    if sines.shape[1] > tl:
        sines = sines[:, :tl]
    elif sines.shape[1] < tl:
        sines = torch.nn.functional.pad(sines, (0, 0, 0, tl - sines.shape[1]))
    uv = (f0 > voiced_threshold).to(dt)
    nz = torch.as_tensor(np.asarray(noise), dtype=dt) if noise is not None else torch.zeros_like(sines)
    nz = (uv * noise_std + (1 - uv) * sine_amp / 3) * nz
    out = sines * uv + nz
    if was_np:
        return out.numpy(), uv.numpy(), nz.numpy()
    return out, uv, nz


def mlx_unwrap(p: np.ndarray, axis=-1) -> np.ndarray:
    """istftnet.py:417-450 with the default discont = period/2 = pi."""
    period = 2 * math.pi
    dd = np.diff(p, axis=axis)
    hi = period / 2
    ddmod = dd - period * np.floor((dd + hi) / period)
    ddmod = np.where((np.abs(dd - hi) < 1e-10) & (dd > 0), hi, ddmod)
    corr = np.where(np.abs(dd) < hi, 0.0, ddmod - dd)
    pad = [(0, 0)] * p.ndim
    pad[axis] = (1, 0)
    return p + np.cumsum(np.pad(corr, pad), axis=axis)


def mlxstft_transform(x, n_fft, hop, win):
    """istftnet.py:473-505: periodic Hann (istftnet.py:461-471), reflect centre -> (mag, phase) [B, n_freq, T]."""
    w = D.hanning(win, periodic=True)
    mags, phs = [], []
    for row in np.asarray(x, dtype=np.float64):
        sp = D.stft(row, n_fft=n_fft, hop_length=hop, win_length=win, window=w, center=True, pad_mode="reflect").T
        mags.append(np.abs(sp))
        # Exactly-real bins (DC, Nyquist, every bin of the reflect-symmetric first frame): the imaginary part is FFT rounding
        # noise whose sign would pick +pi or -pi at random -- in MLX's FFT as in NumPy's.  Canonicalised to +0 (angle 0 / +pi)
        # here and in the CUDA kernel (csrc/dsp.cu:ksrc_stft_kernel); everything else is the plain arctan2 of mlx_angle.
        im = np.where(np.abs(sp.imag) <= 1e-12 * np.abs(sp.real), 0.0, sp.imag)
        phs.append(np.arctan2(im, sp.real))
    return np.stack(mags), np.stack(phs)


def mlxstft_inverse(mag, phase, n_fft, hop, win):
    """istftnet.py:507-540: unwrap along time, mag*exp(j phase) -> istft(normalized=True) -> [B,1,samples]."""
    w = D.hanning(win, periodic=True)
    out = []
    for m, p in zip(np.asarray(mag, dtype=np.float64), np.asarray(phase, dtype=np.float64)):
        pc = mlx_unwrap(p, axis=1)
        out.append(D.istft(m * np.cos(pc) + 1j * m * np.sin(pc), hop_length=hop, win_length=win, window=w,
                           center=True, length=None, normalized=True))
    return np.stack(out)[:, None, :]

# ----------------------------------------------------------------------------- generator / decoder


def generator(P, pre, x, s, f0_curve, cfg, rand_ini, noise):
    """istftnet.py:725-835.  x [B,512,2F] NCL, f0_curve [B,2F] -> audio [B,1,600F]."""
    ist = cfg["istftnet"]
    rates, ksz = ist["upsample_rates"], ist["upsample_kernel_sizes"]
    rk, rd = ist["resblock_kernel_sizes"], ist["resblock_dilation_sizes"]
    n_fft, hop = ist["gen_istft_n_fft"], ist["gen_istft_hop_size"]
    total_up = math.prod(rates) * hop
    dt = x.dtype
    f0 = torch.repeat_interleave(f0_curve[:, :, None], total_up, dim=1)            # nn.Upsample nearest
    sw, _uv, _ = sinegen(f0, upsample_scale=total_up, harmonic_num=8, voiced_threshold=10.0, rand_ini=rand_ini, noise=noise)
    har_src = torch.tanh(N.linear(sw, P[pre + ".m_source.l_linear.weight"], P[pre + ".m_source.l_linear.bias"]))[:, :, 0]
    mag, ph = mlxstft_transform(har_src.numpy(), n_fft, hop, n_fft)
    har = torch.as_tensor(np.concatenate([mag, ph], axis=1)).to(dt).transpose(1, 2)        # NLC [B, T, 22]
    _tap("har", har)
    nk = len(rk)
    for i in range(len(rates)):
        x = N.leaky_relu(x, 0.1)
        if i + 1 < len(rates):
            sf0 = math.prod(rates[i + 1:])
            xs = N.conv1d(har, P[f"{pre}.noise_convs.{i}.weight"].to(dt), stride=sf0, padding=(sf0 + 1) // 2) + P[f"{pre}.noise_convs.{i}.bias"].to(dt)
            xs = adain_resblock1(P, f"{pre}.noise_res.{i}", xs.transpose(1, 2), s, 7)
        else:
            xs = N.conv1d(har, P[f"{pre}.noise_convs.{i}.weight"].to(dt)) + P[f"{pre}.noise_convs.{i}.bias"].to(dt)
            xs = adain_resblock1(P, f"{pre}.noise_res.{i}", xs.transpose(1, 2), s, 11)
        x = conv_weighted(P, f"{pre}.ups.{i}", x.transpose(1, 2), transpose=True, stride=rates[i],
                          padding=(ksz[i] - rates[i]) // 2).transpose(1, 2)
        if i == len(rates) - 1:
            x = torch.nn.functional.pad(x, (1, 0))             # "ReflectionPad1d" is a ZERO pad, istftnet.py:712-718
        x = x + xs
        acc = None
        for j in range(nk):
            r = adain_resblock1(P, f"{pre}.resblocks.{i * nk + j}", x, s, rk[j], rd[j])
            acc = r if acc is None else acc + r
        x = acc / nk
        _tap(f"gen_stage{i}", x.transpose(1, 2))
    x = N.leaky_relu(x, 0.01)
    _tap("gen_pre_post", x.transpose(1, 2))
    x = conv_weighted(P, pre + ".conv_post", x.transpose(1, 2), padding=3).transpose(1, 2)
    _tap("xpost", x.transpose(1, 2))
    nb = n_fft // 2 + 1
    spec = torch.exp(x[:, :nb])
    phase = torch.sin(x[:, nb:])
    return torch.as_tensor(mlxstft_inverse(spec.numpy(), phase.numpy(), n_fft, hop, n_fft)).to(dt)


def decoder(P, asr, f0_curve, n_curve, s, cfg, rand_ini, noise):
    """istftnet.py:936-997.  asr [B,512,F] NCL; F0/N [B,2F]; s [B,128] -> [B,1,600F]."""
    pre = "decoder"
    f0 = conv_weighted(P, pre + ".F0_conv", f0_curve[:, :, None], stride=2, padding=1).transpose(1, 2)
    nn_ = conv_weighted(P, pre + ".N_conv", n_curve[:, :, None], stride=2, padding=1).transpose(1, 2)
    x = torch.cat([asr, f0, nn_], dim=1)
    x = adain_resblk1d(P, pre + ".encode", x, s)
    _tap("dec_encode", x.transpose(1, 2))
    asr_res = conv_weighted(P, pre + ".asr_res.0", asr.transpose(1, 2), padding=0).transpose(1, 2)
    res = True
    for i in range(4):
        if res:
            x = torch.cat([x, asr_res, f0, nn_], dim=1)
        up = (pre + f".decode.{i}.pool.weight_v") in P
        x = adain_resblk1d(P, pre + f".decode.{i}", x, s, upsample=up)
        if up:
            res = False
    _tap("dec_out", x.transpose(1, 2))
    return generator(P, pre + ".generator", x, s, f0_curve, cfg, rand_ini, noise)

# ----------------------------------------------------------------------------- text side


def text_encoder(P, input_ids, cfg):
    """modules.py:21-68 (single unpadded utterance: the mask is all-False)."""
    pre = "text_encoder"
    x = P[pre + ".embedding.weight"][input_ids]                  # [B,T,C] NLC
    k = cfg["text_encoder_kernel_size"]
    for i in range(cfg["n_layer"]):
        x = conv_weighted(P, f"{pre}.cnn.{i}.0", x, padding=(k - 1) // 2)
        x = N.layer_norm(x, P[f"{pre}.cnn.{i}.1.weight"], P[f"{pre}.cnn.{i}.1.bias"])
        x = N.leaky_relu(x, 0.2)
    x = lstm_bi(P, pre + ".lstm", x)
    return x.transpose(1, 2)                                     # [B,512,T]


def duration_encoder(P, d_en, s, cfg):
    """modules.py:380-411.  d_en [B,512,T] NCL, s [B,128] -> d [B,T,640]."""
    pre = "predictor.text_encoder"
    x = d_en.transpose(1, 2)                                     # [B,T,512]
    sb = s[:, None, :].expand(x.shape[0], x.shape[1], s.shape[-1])
    x = torch.cat([x, sb], dim=-1)
    for i in range(cfg["n_layer"]):
        x = lstm_bi(P, f"{pre}.lstms.{2 * i}", x)
        x = ada_layer_norm(P, f"{pre}.lstms.{2 * i + 1}", x, s)
        x = torch.cat([x, sb], dim=-1)
    return x


def f0n_train(P, en, s):
    """modules.py:355-377.  en [B,640,F] -> F0, N [B,2F]."""
    x = lstm_bi(P, "predictor.shared", en.transpose(1, 2))      # [B,F,512]
    outs = []
    for name in ("F0", "N"):
        h = x.transpose(1, 2)
        h = adain_resblk1d(P, f"predictor.{name}.0", h, s)
        h = adain_resblk1d(P, f"predictor.{name}.1", h, s, upsample=True)
        h = adain_resblk1d(P, f"predictor.{name}.2", h, s)
        h = N.conv1d(h.transpose(1, 2), P[f"predictor.{name}_proj.weight"].to(h.dtype)) + P[f"predictor.{name}_proj.bias"].to(h.dtype)
        outs.append(h[:, :, 0])
    return outs


def round_half_even(x: torch.Tensor) -> torch.Tensor:
    return torch.round(x)                                        # torch.round is half-to-even like mx.round


def forward(P, input_ids, ref_s, cfg=KOKORO_CONFIG, speed=1.0, rand_ini=None, noise=None,
            pred_dur_override=None, return_intermediates=False, f0n_override=None):
    """kokoro.py:111-177 from token ids (``[0, *ids, 0]`` already applied by the caller).

    input_ids: LongTensor [1,T]; ref_s [1,256].  Returns (audio [samples], pred_dur [T]).
    ``noise`` may be a callable (n_samples) -> [1, n_samples, 9] since the length depends on pred_dur.
    """
    dt = ref_s.dtype
    t = input_ids.shape[1]
    assert t <= cfg["plbert"]["max_position_embeddings"], (t, cfg["plbert"]["max_position_embeddings"])
    attn_mask = torch.ones(1, t, dtype=torch.int64)
    bert_dur = albert(P, input_ids, attn_mask, cfg["plbert"])
    d_en = N.linear(bert_dur, P["bert_encoder.weight"], P["bert_encoder.bias"]).transpose(1, 2)
    s = ref_s[:, 128:]
    d = duration_encoder(P, d_en, s, cfg)
    x = lstm_bi(P, "predictor.lstm", d)
    dur = N.linear(x, P["predictor.duration_proj.linear_layer.weight"], P["predictor.duration_proj.linear_layer.bias"])
    dur = torch.sigmoid(dur).sum(-1) / speed
    dur = torch.nan_to_num(dur, nan=1.0, posinf=100.0, neginf=1.0)
    pred_dur = torch.clamp(round_half_even(dur), 1, 100).to(torch.int64)[0]
    if pred_dur_override is not None:
        pred_dur = torch.as_tensor(pred_dur_override, dtype=torch.int64)
    idx = torch.repeat_interleave(torch.arange(t), pred_dur)
    aln = torch.zeros(t, idx.shape[0], dtype=dt)
    aln[idx, torch.arange(idx.shape[0])] = 1
    en = d.transpose(1, 2) @ aln[None]
    f0_pred, n_pred = f0n_train(P, en, s)
    _tap("F0", f0_pred)
    _tap("N", n_pred)
    _tap("en", en.transpose(1, 2))
    if f0n_override is not None:
        f0_pred, n_pred = (torch.as_tensor(v).to(dt).reshape(1, -1) for v in f0n_override)
    t_en = text_encoder(P, input_ids, cfg)
    asr = t_en @ aln[None]
    _tap("asr", asr.transpose(1, 2))
    _tap("d", d)
    n_samples = idx.shape[0] * 600
    nz = noise(n_samples) if callable(noise) else noise
    audio = decoder(P, asr, f0_pred, n_pred, ref_s[:, :128], cfg, rand_ini, nz)[0, 0]
    if return_intermediates:
        return audio, pred_dur, {"bert": bert_dur, "d": d, "dur": dur, "en": en, "F0": f0_pred, "N": n_pred,
                                 "t_en": t_en, "asr": asr}
    return audio, pred_dur
