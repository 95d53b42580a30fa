"""Oracle building blocks: the MLX op semantics the reference's hot path leans on,
restated on torch-CPU tensors (float64 for checking, float32 for the timed CPU baseline).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Semantics per SURVEY.md appendix B:
activations are channels-last ``[N, L, C]``; conv weights are MLX-layout
``[Cout, K, Cin/groups]``; variance is biased; softmax accumulates in the working dtype.
NumPy arrays are accepted and returned as NumPy (used by the small pin tests).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x))


def _ret(y, was_np):
    return y.numpy() if was_np else y


def conv1d(x, w, stride=1, padding=0, dilation=1, groups=1, bias=None):
    """mx.conv1d(x[N,L,Cin], w[Cout,K,Cin/g]) -- cross-correlation, zero padding both sides."""
    was_np = not isinstance(x, torch.Tensor)
    x, w = _t(x), _t(w)
    y = F.conv1d(x.transpose(1, 2), w.permute(0, 2, 1).to(x.dtype), None, stride, padding, dilation, groups)
    y = y.transpose(1, 2)
    if bias is not None:
        y = y + _t(bias).to(y.dtype)
    return _ret(y, was_np)


def conv_transpose1d(x, w, stride=1, padding=0, dilation=1, output_padding=0, groups=1, bias=None):
    """mx.conv_transpose1d(x[N,L,Cin], w[Cout,K,Cin/g]): scatter form
    ``y[s*i + d*k - p, co] += x[i, ci] * w[co, k, ci]`` (no kernel flip);
    Lout = (L-1)s - 2p + d(K-1) + output_padding + 1.
    """
    was_np = not isinstance(x, torch.Tensor)
    x, w = _t(x), _t(w)
    cout, k, cin_g = w.shape
    # torch layout: [Cin, Cout/g, K]; MLX [Cout, K, Cin/g] with groups splitting Cout and Cin.
    wt = w.reshape(groups, cout // groups, k, cin_g).permute(0, 3, 1, 2).reshape(groups * cin_g, cout // groups, k)
    y = F.conv_transpose1d(x.transpose(1, 2), wt.to(x.dtype), None, stride, padding, output_padding, groups, dilation)
    y = y.transpose(1, 2)
    if bias is not None:
        y = y + _t(bias).to(y.dtype)
    return _ret(y, was_np)


def linear(x, w, b=None):
    """nn.Linear: x @ w.T + b with w [out, in]."""
    y = x @ w.to(x.dtype).T
    return y if b is None else y + b.to(x.dtype)


def layer_norm(x, w=None, b=None, eps=1e-5):
    """nn.LayerNorm over the last axis, biased variance."""
    mu = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    if w is not None:
        y = y * w.to(x.dtype)
    if b is not None:
        y = y + b.to(x.dtype)
    return y


def rms_norm(x, w, eps):
    """nn.RMSNorm: x * rsqrt(mean(x^2) + eps) * w."""
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.to(x.dtype)


def gelu(x):
    """nn.gelu -- exact erf form."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_approx(x):
    """nn.gelu_approx -- tanh form."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def leaky_relu(x, slope):
    return torch.where(x > 0, x, x * slope)


def softmax(x, dim=-1):
    return torch.softmax(x, dim=dim)


def sdpa(q, k, v, scale, mask=None):
    """mx.fast.scaled_dot_product_attention on [B,H,T,D] with additive mask; GQA by head repeat."""
    hq, hk = q.shape[1], k.shape[1]
    if hq != hk:
        k = k.repeat_interleave(hq // hk, dim=1)
        v = v.repeat_interleave(hq // hk, dim=1)
    s = (q @ k.transpose(-1, -2)) * scale
    if mask is not None:
        s = s + mask
    return torch.softmax(s, dim=-1) @ v


def rope_traditional(x, offset, base):
    """nn.RoPE(dims, traditional=True): rotate interleaved pairs (2i, 2i+1); x [B,H,T,D]."""
    b, h, t, d = x.shape
    pos = torch.arange(offset, offset + t, dtype=x.dtype)
    inv = torch.exp(-torch.arange(0, d // 2, dtype=x.dtype) * (math.log(base) / (d // 2)))
    ang = pos[:, None] * inv[None, :]
    c, s = torch.cos(ang), torch.sin(ang)
    x1, x2 = x[..., 0::2], x[..., 1::2]
    out = torch.stack([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1)
    return out.reshape(b, h, t, d)


def rope_half(x, offset, base):
    """nn.RoPE(dims, traditional=False): rotate the pairs (i, i + D/2); x [B,H,T,D]."""
    b, h, t, d = x.shape
    pos = torch.arange(offset, offset + t, dtype=x.dtype)
    inv = torch.exp(-torch.arange(0, d // 2, dtype=x.dtype) * (math.log(base) / (d // 2)))
    ang = pos[:, None] * inv[None, :]
    c, s = torch.cos(ang), torch.sin(ang)
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    return torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1)


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bfloat16 and back (the precision a bf16 checkpoint stores)."""
    return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)
