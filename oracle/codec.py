"""Oracle for the codec decode stacks: SNAC (codec/models/snac/*) and Mimi (codec/models/mimi/*).

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-CPU in the dtype of the weights dict.
Parameter names are the reference's MLX parameter trees.  MLX PRNG draws (SNAC NoiseBlock,
snac/layers.py:263) are injected.
"""
from __future__ import annotations

import math

import torch

from . import nn as N

# ============================================================================= SNAC

SNAC_24K = {   # codec/tests/test_snac.py:7-19
    "sampling_rate": 24000, "encoder_dim": 48, "encoder_rates": [2, 4, 8, 8], "decoder_dim": 1024,
    "decoder_rates": [8, 8, 4, 2], "attn_window_size": None, "codebook_size": 4096, "codebook_dim": 8,
    "vq_strides": [4, 2, 1], "noise": True, "depthwise": True,
}


def _wn(P, pre, except_dim=0):
    """snac/layers.py:9-14,57: w = g * v / ||v|| (norm over all axes but `except_dim`), no epsilon."""
    v, g = P[pre + ".weight_v"], P[pre + ".weight_g"]
    axes = tuple(i for i in range(v.ndim) if i != except_dim)
    return g * v / torch.sqrt((v * v).sum(dim=axes, keepdim=True))


def snac_wnconv(P, pre, x, stride=1, padding=0, dilation=1, groups=1):
    """WNConv1d (snac/layers.py:17-61) on NLC x."""
    return N.conv1d(x, _wn(P, pre), stride, padding, dilation, groups, P.get(pre + ".bias"))


def snac_wnconvtr(P, pre, x, stride, padding):
    """WNConvTranspose1d (snac/layers.py:64-121).  Quirk kept: the module passes
    ``groups`` (=1) in mx.conv_transpose1d's ``output_padding`` positional slot, so every decoder
    transposed conv runs with output_padding=1 and groups=1; the declared output_padding is ignored."""
    w = _wn(P, pre, except_dim=0).transpose(0, 2)            # stored (in, K, out) -> (out, K, in)
    return N.conv_transpose1d(x, w, stride, padding, 1, 1, 1, P.get(pre + ".bias"))


def snake_snac(x, alpha):
    """snac/layers.py:124-130 on NLC with alpha [1,1,C]: x + sin(a x)^2 / (a + 1e-9)."""
    return x + (1.0 / (alpha + 1e-9)) * torch.sin(alpha * x) ** 2


def snac_from_codes(P, codes, cfg):
    """ResidualVectorQuantize.from_codes (snac/vq.py:111-131): -> z_q [B, D, T]."""
    zq = 0.0
    for i, stride in enumerate(cfg["vq_strides"]):
        pre = f"quantizer.quantizers.{i}"
        zp = P[pre + ".codebook.weight"][codes[i]]                                   # [B,Tl,cd]
        z = snac_wnconv(P, pre + ".out_proj", zp)                                    # [B,Tl,D]
        if stride > 1:
            z = torch.repeat_interleave(z, stride, dim=1)
        zq = zq + z
    return zq.transpose(1, 2)


def snac_residual_unit(P, pre, x, dilation, groups):
    """ResidualUnit (snac/layers.py:208-231), kernel 7."""
    a1, a2 = P[pre + ".block.layers.0.alpha"].transpose(1, 2), P[pre + ".block.layers.2.alpha"].transpose(1, 2)
    y = snake_snac(x, a1)
    y = snac_wnconv(P, pre + ".block.layers.1", y, padding=3 * dilation, dilation=dilation, groups=groups)
    y = snake_snac(y, a2)
    y = snac_wnconv(P, pre + ".block.layers.3", y)
    return x + y


def snac_decoder(P, z, cfg, noises):
    """Decoder (snac/layers.py:159-205) on NLC z [B,T,latent]; noises[i] [B,1,C_i] injected.

    Quirk kept (snac/layers.py:261-267): NoiseBlock unpacks the NLC shape as (B, C, T), so its noise
    tensor has shape (B, 1, C) -- one Gaussian per CHANNEL, constant over time."""
    pre = "decoder.model.layers"
    li = 0
    c = z.shape[-1]
    if cfg["depthwise"]:
        x = snac_wnconv(P, f"{pre}.{li}", z, padding=3, groups=c); li += 1           # depthwise k7
        x = snac_wnconv(P, f"{pre}.{li}", x); li += 1                                 # pointwise
    else:
        x = snac_wnconv(P, f"{pre}.{li}", z, padding=3); li += 1                      # one dense k7 conv (layers.py:187-188)
    ch = cfg["decoder_dim"]
    for i, stride in enumerate(cfg["decoder_rates"]):
        bp = f"{pre}.{li}.block.layers"; li += 1
        out_dim = ch // (2 ** (i + 1))
        bi = 0
        x = snake_snac(x, P[f"{bp}.{bi}.alpha"].transpose(1, 2)); bi += 1
        x = snac_wnconvtr(P, f"{bp}.{bi}", x, stride, math.ceil(stride / 2)); bi += 1
        if cfg["noise"]:
            h = snac_wnconv(P, f"{bp}.{bi}.linear", x); bi += 1
            x = x + noises[i].to(x.dtype) * h
        for d in (1, 3, 9):
            x = snac_residual_unit(P, f"{bp}.{bi}", x, d, out_dim if cfg["depthwise"] else 1); bi += 1
    x = snake_snac(x, P[f"{pre}.{li}.alpha"].transpose(1, 2)); li += 1
    x = snac_wnconv(P, f"{pre}.{li}", x, padding=3)
    return torch.tanh(x)


def snac_encoder(P, x, cfg):
    """Encoder (snac/layers.py:133-158) on NLC x [B,n,1] -> [B,T,latent]: k7 conv, per stride {3 residual units, Snake, strided conv
    k = 2s, padding ceil(s/2)} with channels doubling, final (depthwise) k7 conv.  LocalMHA only when attn_window_size is set (44 kHz)."""
    if cfg.get("attn_window_size") is not None:
        raise NotImplementedError("LocalMHA (attn_window_size) is outside the 24 kHz path")
    pre = "encoder.block.layers"
    x = snac_wnconv(P, f"{pre}.0", x, padding=3)
    d, li = cfg["encoder_dim"], 1
    for stride in cfg["encoder_rates"]:
        d *= 2
        groups = d // 2 if cfg["depthwise"] else 1
        bp = f"{pre}.{li}.block.layers"
        for bi, dil in enumerate((1, 3, 9)):
            x = snac_residual_unit(P, f"{bp}.{bi}", x, dil, groups)
        x = snake_snac(x, P[f"{bp}.3.alpha"].transpose(1, 2))
        x = snac_wnconv(P, f"{bp}.4", x, stride=stride, padding=math.ceil(stride / 2))
        li += 1
    return snac_wnconv(P, f"{pre}.{li}", x, padding=3, groups=d if cfg["depthwise"] else 1)


def snac_quantize(P, z, cfg):
    """ResidualVectorQuantize.__call__ over VectorQuantize.__call__ / decode_latents (snac/vq.py:22-90,93-109) on z [B,D,T]:
    per level average-pool the residual by the level's stride, project to 8 dims, pick the nearest L2-normalised code
    (first index on ties), project back, repeat by the stride, subtract from the residual.  -> (z_q [B,D,T], [codes [B,T/stride]])."""
    zq, residual, codes = 0.0, z, []
    for i, stride in enumerate(cfg["vq_strides"]):
        pre = f"quantizer.quantizers.{i}"
        x = residual.transpose(1, 2)                                                   # NLC
        if stride > 1:
            t = (x.shape[1] - stride) // stride + 1
            x = x[:, : t * stride].reshape(x.shape[0], t, stride, x.shape[2]).sum(dim=2) / stride
        z_e = snac_wnconv(P, pre + ".in_proj", x)                                      # [B,T',cd]
        cb = P[pre + ".codebook.weight"]
        e = z_e.reshape(-1, z_e.shape[-1])
        en = e / torch.clamp(torch.sqrt((e.abs() ** 2).sum(1, keepdim=True)), min=1e-12)
        cn = cb / torch.clamp(torch.sqrt((cb.abs() ** 2).sum(1, keepdim=True)), min=1e-12)
        dist = (en ** 2).sum(1, keepdim=True) - 2 * en @ cn.T + (cn ** 2).sum(1, keepdim=True).T
        idx = (-dist).argmax(1).reshape(z_e.shape[0], z_e.shape[1])
        z_qi = z_e + (cb[idx] - z_e)                                                   # straight-through form, kept literally
        z_qi = snac_wnconv(P, pre + ".out_proj", z_qi).transpose(1, 2)                 # [B,D,T']
        if stride > 1:
            z_qi = torch.repeat_interleave(z_qi, stride, dim=2)
        zq = zq + z_qi
        residual = residual - z_qi
        codes.append(idx)
    return zq, codes


def snac_preprocess(audio, cfg):
    """SNAC.preprocess (snac/snac.py:67-84): right-pad [B,1,n] to a multiple of hop * lcm(vq_strides[, attn window])."""
    lcm = cfg["vq_strides"][0]
    for v in cfg["vq_strides"][1:]:
        lcm = abs(lcm * v) // math.gcd(lcm, v)
    if cfg.get("attn_window_size"):
        lcm = abs(lcm * cfg["attn_window_size"]) // math.gcd(lcm, cfg["attn_window_size"])
    pad_to = math.prod(cfg["encoder_rates"]) * lcm
    n = audio.shape[-1]
    return torch.nn.functional.pad(audio, (0, math.ceil(n / pad_to) * pad_to - n))


def snac_encode(P, audio, cfg=SNAC_24K):
    """SNAC.encode (snac/snac.py:95-99): audio [B,1,n] -> list of int64 codes, coarse to fine."""
    x = snac_preprocess(audio, cfg)
    z = snac_encoder(P, x.transpose(1, 2), cfg).transpose(1, 2)
    return snac_quantize(P, z, cfg)[1]


def snac_decode(P, codes, cfg=SNAC_24K, noises=None):
    """SNAC.decode (snac/snac.py:101-104): list of codes -> audio [B, T_out, 1]."""
    z = snac_from_codes(P, codes, cfg)
    if noises is None:
        noises = [torch.zeros(1, 1, 1, dtype=z.dtype)] * len(cfg["decoder_rates"])
    return snac_decoder(P, z.transpose(1, 2), cfg, noises)

def snac_decode_stream(P, codes, prev_codes=None, context_frames=8, cfg=SNAC_24K, noises=None):
    """SNAC.decode_stream (snac/snac.py:106-162), literally -- including the ``[..., context_samples:]`` slice that acts on the channel
    axis of the [B, T, 1] audio and therefore trims nothing.  -> (audio, new context)."""
    new_context = [c[:, -context_frames:] if c.shape[1] > context_frames else c for c in codes]
    if prev_codes is None:
        return snac_decode(P, codes, cfg, noises), new_context
    combined = []
    for stride, prev, new in zip(cfg["vq_strides"], prev_codes, codes):
        keep = max(1, context_frames // stride)
        prev = prev[:, -keep:] if prev.shape[1] > keep else prev
        combined.append(torch.cat([prev, new], dim=1))
    full = snac_decode(P, combined, cfg, noises)
    n = context_frames * math.prod(cfg["encoder_rates"])
    return (full[..., n:] if full.shape[-1] > n else full), new_context


# ============================================================================= Mimi

MIMI_202407 = {   # codec/models/mimi/mimi.py:47-96
    "dimension": 512, "nfilters": 64, "ratios": [8, 6, 5, 4], "ksize": 7, "residual_ksize": 3, "last_ksize": 3,
    "compress": 2, "d_model": 512, "num_heads": 8, "num_layers": 8, "dim_feedforward": 2048, "context": 250,
    "max_period": 10000, "layer_scale": 0.01, "nq": 32, "bins": 2048, "qdim": 256, "upsample_stride": 2,
}


def mimi_causal_conv(P, pre, x, ksize, stride=1, dilation=1, pad_mode="constant"):
    """StreamableConv1d.__call__ (mimi/modules/conv.py:224-243), causal, constant (or "edge" = replicate) pad, on NCL x."""
    k_eff = (ksize - 1) * dilation + 1
    pad_total = k_eff - stride
    ln = x.shape[-1]
    nframes = max(ln + pad_total - k_eff, 0) / stride + 1.0
    extra = max(0, (int(math.ceil(nframes)) - 1) * stride + k_eff - pad_total - ln)
    xp = torch.nn.functional.pad(x, (pad_total, extra), mode="replicate" if pad_mode == "edge" else "constant")
    return N.conv1d(xp.transpose(1, 2), P[pre + ".conv.conv.weight"].to(x.dtype), stride, 0, dilation, 1,
                    P.get(pre + ".conv.conv.bias")).transpose(1, 2)


def mimi_causal_convtr(P, pre, x, ksize, stride, groups=1):
    """StreamableConvTranspose1d.__call__ (conv.py:303-313): full scatter output, trim k - stride on the right."""
    y = N.conv_transpose1d(x.transpose(1, 2), P[pre + ".convtr.convtr.weight"].to(x.dtype), stride, 0, 1, 0, groups,
                           P.get(pre + ".convtr.convtr.bias")).transpose(1, 2)
    trim = max(ksize - stride, 0)
    return y[..., : y.shape[-1] - trim]


def mimi_quantizer_decode(P, codes, cfg):
    """SplitResidualVectorQuantizer.decode (quantization.py:187-191,144-149,103-108,47-49): codes [B,nq,T] -> [B,512,T]."""
    def emb(pre):
        usage = torch.clamp(P[pre + ".cluster_usage"], min=1e-5)[:, None]
        return P[pre + ".embedding_sum"] / usage
    out = None
    for name, qs in (("rvq_first", [0]), ("rvq_rest", list(range(1, codes.shape[1])))):
        if not qs:
            continue
        q = None
        for li, qi in enumerate(qs):
            e = emb(f"quantizer.{name}.vq.layers.{li}.codebook")[codes[:, qi]]        # [B,T,256]
            q = e if q is None else q + e
        q = q.transpose(1, 2)
        y = N.conv1d(q.transpose(1, 2), P[f"quantizer.{name}.output_proj.weight"].to(q.dtype)).transpose(1, 2)
        out = y if out is None else out + y
    return out


def mimi_transformer(P, pre, x, cfg, rope_traditional=True, full_causal=False):
    """ProjectedTransformer / Transformer (transformer.py:63-261) on NCL x (conv_layout), fresh cache (offset 0).  ``full_causal``: the
    caller passes its own causal mask, which replaces the context-window mask (speech_tokenizer.py:1052-1057); ``rope_traditional`` False
    rotates the half-split pairs (the Qwen3 tokenizer encoder's setting, :1012)."""
    x = x.transpose(1, 2)
    b, t, d = x.shape
    nh = cfg["num_heads"]
    hd = d // nh
    i, j = torch.arange(t)[:, None], torch.arange(t)[None, :]
    if full_causal:
        mask = torch.where(j <= i, 0.0, float("-inf")).to(x.dtype)
    else:
        allowed = (j <= i) & (i - j < cfg["context"])
        mask = torch.where(allowed, 0.0, -1e9).to(x.dtype)
    rope = N.rope_traditional if rope_traditional else N.rope_half
    for li in range(cfg["num_layers"]):
        L = f"{pre}.transformer.layers.{li}"
        n1 = N.layer_norm(x, P[L + ".norm1.weight"], P[L + ".norm1.bias"], 1e-5)
        qkv = N.linear(n1, P[L + ".self_attn.in_proj.weight"]).reshape(b, t, 3, nh, hd)
        q, k, v = (qkv[:, :, n].transpose(1, 2) for n in range(3))
        q = rope(q, 0, cfg["max_period"])
        k = rope(k, 0, cfg["max_period"])
        a = N.sdpa(q, k, v, hd ** -0.5, mask).transpose(1, 2).reshape(b, t, d)
        a = N.linear(a, P[L + ".self_attn.out_proj.weight"])
        x = x + a * P[L + ".layer_scale_1.scale"].to(x.dtype)
        n2 = N.layer_norm(x, P[L + ".norm2.weight"], P[L + ".norm2.bias"], 1e-5)
        m = N.linear(N.gelu_approx(N.linear(n2, P[L + ".gating.linear1.weight"])), P[L + ".gating.linear2.weight"])
        x = x + m * P[L + ".layer_scale_2.scale"].to(x.dtype)
    return x.transpose(1, 2)


def mimi_seanet_decoder(P, x, cfg):
    """SeanetDecoder.__call__ (seanet.py:257-300) on NCL x [B,512,T]."""
    pre = "decoder"
    x = mimi_causal_conv(P, pre + ".init_conv1d", x, cfg["ksize"])
    for li, ratio in enumerate(cfg["ratios"]):
        L = f"{pre}.layers.{li}"
        x = mimi_causal_convtr(P, L + ".upsample", N.elu(x), 2 * ratio, ratio)
        r = x                                                       # SeanetResnetBlock, true_skip (seanet.py:99-107)
        y = mimi_causal_conv(P, L + ".residuals.0.block.0", N.elu(x), cfg["residual_ksize"])
        y = mimi_causal_conv(P, L + ".residuals.0.block.1", N.elu(y), 1)
        x = y + r
    return mimi_causal_conv(P, pre + ".final_conv1d", N.elu(x), cfg["last_ksize"])


def mimi_seanet_encoder(P, x, cfg, root=""):
    """SeanetEncoder.__call__ (seanet.py:194-199) on NCL x [B,1,n]: init conv, per ratio (reversed) {resnet block, ELU, strided conv
    k = 2r}, ELU, final conv."""
    pre = root + "encoder"
    x = mimi_causal_conv(P, pre + ".init_conv1d", x, cfg["ksize"])
    for li, ratio in enumerate(reversed(cfg["ratios"])):
        L = f"{pre}.layers.{li}"
        y = mimi_causal_conv(P, L + ".residuals.0.block.0", N.elu(x), cfg["residual_ksize"])
        y = mimi_causal_conv(P, L + ".residuals.0.block.1", N.elu(y), 1)
        x = mimi_causal_conv(P, L + ".downsample", N.elu(y + x), 2 * ratio, stride=ratio)
    return mimi_causal_conv(P, pre + ".final_conv1d", N.elu(x), cfg["last_ksize"])


def mimi_quantizer_encode(P, x, cfg, root=""):
    """SplitResidualVectorQuantizer.encode (quantization.py:178-185,138-141,90-101,37-45): x [B,512,T] -> int64 codes [B,nq,T].
    Nearest code = argmin(|e|^2 / 2 - x.e) on the residual, which is then reduced by the chosen embedding."""
    codes = []
    for name, nq in (("rvq_first", 1), ("rvq_rest", cfg["nq"] - 1)):
        if nq <= 0:
            continue
        r = N.conv1d(x.transpose(1, 2), P[f"{root}quantizer.{name}.input_proj.weight"].to(x.dtype))    # [B,T,qdim]
        for li in range(nq):
            pre = f"{root}quantizer.{name}.vq.layers.{li}.codebook"
            emb = P[pre + ".embedding_sum"] / torch.clamp(P[pre + ".cluster_usage"], min=1e-5)[:, None]
            idx = ((emb * emb).sum(-1) / 2 - r @ emb.T).argmin(dim=-1)                                  # [B,T]
            r = r - emb[idx]
            codes.append(idx)
    return torch.stack(codes, dim=1)


def mimi_encode(P, pcm, cfg=MIMI_202407):
    """Mimi.encode (mimi.py:146-153): pcm [B,1,n] -> codes [B,nq,ceil(n/1920)] (SEANet encoder, encoder transformer, stride-2
    replicate-padded downsampling conv, split RVQ)."""
    x = mimi_seanet_encoder(P, pcm, cfg)
    x = mimi_transformer(P, "encoder_transformer", x, cfg)
    s = cfg["upsample_stride"]
    x = mimi_causal_conv(P, "downsample.conv", x, 2 * s, stride=s, pad_mode="edge")
    return mimi_quantizer_encode(P, x, cfg)


def mimi_decode(P, codes, cfg=MIMI_202407):
    """Mimi.decode (mimi.py:155-162): codes [B,nq,T] -> pcm [B,1,1920 T]."""
    x = mimi_quantizer_decode(P, codes, cfg)
    s = cfg["upsample_stride"]
    x = mimi_causal_convtr(P, "upsample.convtr", x, 2 * s, s, groups=x.shape[1])
    x = mimi_transformer(P, "decoder_transformer", x, cfg)
    return mimi_seanet_decoder(P, x, cfg)
