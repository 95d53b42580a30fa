"""The C-ABI hot-path entry points surfaced as ``torch.library`` custom ops (namespace ``b200audio``), each with a fake (meta)
implementation so that they trace under FakeTensorMode / torch.export without touching the GPU (north_star: "a thin C-ABI layer
surfaced as torch custom ops").

The ops are thin: every one forwards to the wrapper in ``ops.py`` that calls ``libb200audio.so`` on torch's current stream; there is no
CPU implementation (calling one on CPU tensors raises).  Tensor layouts are those of ``include/b200audio.h``: activations fp32
channels-last ``[B, L, C]``, MLX-layout conv weights ``[Cout, K, Cin / groups]`` (packed once per weight tensor and cached).

    y = torch.ops.b200audio.conv1d_cl(x, w, bias, stride, dilation, pad_left, groups, transpose, pre_act, pre_p0, post_act)
    y = torch.ops.b200audio.linear(x, w, bias, post_act)
    o = torch.ops.b200audio.attention(q, k, v, n_heads, scale, causal, window)
    y = torch.ops.b200audio.lstm_bidir(xproj, wh)
    y = torch.ops.b200audio.layernorm(x, w, b, eps, rms)
    m = torch.ops.b200audio.whisper_logmel(audio, n_mels, padding)
    z = torch.ops.b200audio.rvq_decode(codes, codebooks)
    c = torch.ops.b200audio.rvq_encode(x, codebooks, c2, mode)
    re, im = torch.ops.b200audio.stft(x, window, n_fft, hop, pad_mode)
    a = torch.ops.b200audio.kokoro_istft_head(x)
    t = torch.ops.b200audio.sample_token(logits, u, temperature, top_k, top_p, min_p)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops

_PACKED = {}


def _packed(w: torch.Tensor, bias: Optional[torch.Tensor], groups: int):
    """Pack an MLX-layout weight once per (storage, version): folded layouts for the CUDA-core and tcgen05 paths (ops.pack_conv)."""
    key = (w.data_ptr(), w._version, tuple(w.shape), None if bias is None else bias.data_ptr(), groups, str(w.device))
    cw = _PACKED.get(key)
    if cw is None:
        if len(_PACKED) > 512:
            _PACKED.clear()
        cw = _PACKED[key] = ops.pack_conv(w.detach().float().cpu(), None if bias is None else bias.detach().float().cpu(), groups, w.device)
    return cw


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b200audio ops run on CUDA tensors only: the hot path has no CPU fallback")


@torch.library.custom_op("b200audio::conv1d_cl", mutates_args=())
def conv1d_cl(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], stride: int, dilation: int, pad_left: int, groups: int,
              transpose: bool, pre_act: int, pre_p0: float, post_act: int) -> torch.Tensor:
    _need_cuda(x, w)
    pre = ops.Pre(act=pre_act, p0=pre_p0) if pre_act else None
    return ops.conv1d(x, _packed(w, bias, groups), stride=stride, dilation=dilation, pad_left=pad_left, pre=pre, post_act=post_act, transpose=transpose)


@conv1d_cl.register_fake
def _(x, w, bias, stride, dilation, pad_left, groups, transpose, pre_act, pre_p0, post_act):
    B, L, _ = x.shape
    cout, K = w.shape[0], w.shape[1]
    lout = (L - 1) * stride + K - 2 * pad_left if transpose else (L + 2 * pad_left - dilation * (K - 1) - 1) // stride + 1
    return x.new_empty(B, lout, cout)


@torch.library.custom_op("b200audio::linear", mutates_args=())
def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], post_act: int) -> torch.Tensor:
    _need_cuda(x, w)
    shp = x.shape
    y = ops.linear(x.reshape(-1, shp[-1]), _packed(w[:, None, :], bias, 1), post_act=post_act)
    return y.reshape(*shp[:-1], w.shape[0])


@linear.register_fake
def _(x, w, bias, post_act):
    return x.new_empty(*x.shape[:-1], w.shape[0])


@torch.library.custom_op("b200audio::attention", mutates_args=())
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_heads: int, scale: float, causal: bool, window: int) -> torch.Tensor:
    _need_cuda(q, k, v)
    return ops.attention(q, k, v, n_heads=n_heads, scale=scale, causal=causal, window=window)


@attention.register_fake
def _(q, k, v, n_heads, scale, causal, window):
    return q.new_empty(q.shape)


@torch.library.custom_op("b200audio::lstm_bidir", mutates_args=())
def lstm_bidir(xproj: torch.Tensor, wh: torch.Tensor) -> torch.Tensor:
    _need_cuda(xproj, wh)
    return ops.lstm_bidir(xproj.contiguous(), wh.contiguous())


@lstm_bidir.register_fake
def _(xproj, wh):
    B, T, g8 = xproj.shape
    return xproj.new_empty(B, T, g8 // 4)


@torch.library.custom_op("b200audio::layernorm", mutates_args=())
def layernorm(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], eps: float, rms: bool) -> torch.Tensor:
    _need_cuda(x)
    return ops.layernorm(x, w, b, eps=eps, rms=rms)


@layernorm.register_fake
def _(x, w, b, eps, rms):
    return x.new_empty(x.shape)


@torch.library.custom_op("b200audio::whisper_logmel", mutates_args=())
def whisper_logmel(audio: torch.Tensor, n_mels: int, padding: int) -> torch.Tensor:
    _need_cuda(audio)
    from .stt.models.whisper.audio import log_mel_spectrogram
    return log_mel_spectrogram(audio, n_mels, padding=padding, device=audio.device)


@whisper_logmel.register_fake
def _(audio, n_mels, padding):
    frames = (audio.shape[-1] + padding) // 160
    return audio.new_empty(*audio.shape[:-1], frames, n_mels)


@torch.library.custom_op("b200audio::rvq_decode", mutates_args=())
def rvq_decode(codes: torch.Tensor, codebooks: torch.Tensor) -> torch.Tensor:
    _need_cuda(codes, codebooks)
    return ops.rvq_decode(codes, codebooks)


@rvq_decode.register_fake
def _(codes, codebooks):
    B, _, T = codes.shape
    return codebooks.new_empty(B, T, codebooks.shape[2])


@torch.library.custom_op("b200audio::rvq_encode", mutates_args=())
def rvq_encode(x: torch.Tensor, codebooks: torch.Tensor, c2: torch.Tensor, mode: int) -> torch.Tensor:
    _need_cuda(x, codebooks, c2)
    return ops.rvq_encode(x, codebooks, c2, mode=mode)


@rvq_encode.register_fake
def _(x, codebooks, c2, mode):
    return x.new_empty(x.shape[0], codebooks.shape[0], dtype=torch.int64)


@torch.library.custom_op("b200audio::stft", mutates_args=())
def stft(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int, pad_mode: int) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(x, window)
    n = x.shape[1]
    frames = n // hop + 1 if pad_mode else (n - n_fft) // hop + 1
    return ops.stft(x, window, n_fft, hop, pad_mode, frames)


@stft.register_fake
def _(x, window, n_fft, hop, pad_mode):
    n = x.shape[1]
    frames = n // hop + 1 if pad_mode else (n - n_fft) // hop + 1
    return x.new_empty(x.shape[0], frames, n_fft // 2 + 1), x.new_empty(x.shape[0], frames, n_fft // 2 + 1)


@torch.library.custom_op("b200audio::kokoro_istft_head", mutates_args=())
def kokoro_istft_head(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    return ops.kokoro_istft_head(x)


@kokoro_istft_head.register_fake
def _(x):
    return x.new_empty(x.shape[0], (x.shape[1] - 1) * 5)


@torch.library.custom_op("b200audio::sample_token", mutates_args=())
def sample_token(logits: torch.Tensor, u: torch.Tensor, temperature: float, top_k: int, top_p: float, min_p: float) -> torch.Tensor:
    _need_cuda(logits, u)
    return ops.sample_token(logits, temperature=temperature, top_k=top_k, top_p=top_p, min_p=min_p, u=u)


@sample_token.register_fake
def _(logits, u, temperature, top_k, top_p, min_p):
    return logits.new_empty(logits.shape[0], dtype=torch.int64)


OPS = ("conv1d_cl", "linear", "attention", "lstm_bidir", "layernorm", "whisper_logmel", "rvq_decode", "rvq_encode", "stft", "kokoro_istft_head",
       "sample_token")
