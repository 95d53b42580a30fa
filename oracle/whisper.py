"""Oracle for the Whisper audio encoder (stt/models/whisper/whisper.py:329-448) -- torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The log-mel frontend is in oracle/dsp.py
(``whisper_log_mel``).  Parameter names are the reference's MLX parameter tree.
"""
from __future__ import annotations

import math

import torch

from . import nn as N

WHISPER_SMALL = {"n_mels": 80, "n_audio_ctx": 1500, "n_audio_state": 768, "n_audio_head": 12, "n_audio_layer": 12,
                 "n_vocab": 51865, "n_text_ctx": 448, "n_text_state": 768, "n_text_head": 12, "n_text_layer": 12}


def sinusoids(length, channels, max_timescale=10000, dtype=torch.float64):
    """whisper.py:329-335."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=dtype))
    st = torch.arange(length, dtype=dtype)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def mha(P, pre, x, n_head, xa=None, mask=None):
    """MultiHeadAttention (whisper.py:338-385): q and k each scaled by d^-0.25, softmax in the working dtype."""
    q = N.linear(x, P[pre + ".query.weight"], P[pre + ".query.bias"])
    src = x if xa is None else xa
    k = N.linear(src, P[pre + ".key.weight"])
    v = N.linear(src, P[pre + ".value.weight"], P[pre + ".value.bias"])
    b, t, d = q.shape
    scale = (d // n_head) ** -0.25
    qh = q.reshape(b, t, n_head, -1).permute(0, 2, 1, 3) * scale
    kh = k.reshape(b, k.shape[1], n_head, -1).permute(0, 2, 3, 1) * scale
    vh = v.reshape(b, v.shape[1], n_head, -1).permute(0, 2, 1, 3)
    qk = qh @ kh
    if mask is not None:
        qk = qk + mask[:t, :t]
    out = (torch.softmax(qk, dim=-1) @ vh).permute(0, 2, 1, 3).reshape(b, t, d)
    return N.linear(out, P[pre + ".out.weight"], P[pre + ".out.bias"])


def encoder(P, mel, dims=WHISPER_SMALL):
    """AudioEncoder.__call__ (whisper.py:438-448): mel [B, 3000, n_mels] (frames-major) -> [B, 1500, d]."""
    dt = mel.dtype
    x = N.gelu(N.conv1d(mel, P["encoder.conv1.weight"].to(dt), 1, 1, 1, 1, P["encoder.conv1.bias"]))
    x = N.gelu(N.conv1d(x, P["encoder.conv2.weight"].to(dt), 2, 1, 1, 1, P["encoder.conv2.bias"]))
    assert x.shape[1:] == (dims["n_audio_ctx"], dims["n_audio_state"]), "incorrect audio shape"
    x = x + sinusoids(dims["n_audio_ctx"], dims["n_audio_state"], dtype=dt)
    for i in range(dims["n_audio_layer"]):
        L = f"encoder.blocks.{i}"
        x = x + mha(P, L + ".attn", N.layer_norm(x, P[L + ".attn_ln.weight"], P[L + ".attn_ln.bias"]), dims["n_audio_head"])
        h = N.layer_norm(x, P[L + ".mlp_ln.weight"], P[L + ".mlp_ln.bias"])
        x = x + N.linear(N.gelu(N.linear(h, P[L + ".mlp1.weight"], P[L + ".mlp1.bias"])), P[L + ".mlp2.weight"], P[L + ".mlp2.bias"])
    return N.layer_norm(x, P["encoder.ln_post.weight"], P["encoder.ln_post.bias"])
