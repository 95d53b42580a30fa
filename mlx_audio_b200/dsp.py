"""``mlx_audio.dsp`` surface on B200: windows, STFT, iSTFT, ISTFTCache, mel filterbank.

Same names, defaults and error messages as the reference (dsp.py:39-94,385-752); arrays are torch
tensors and the transforms run in our kernels (csrc/dsp.cu).  Window tables and the filterbank are
one-time host computations (the reference builds them from Python floats / a cached float32 graph too).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Optional

import numpy as np
import torch

from . import ops

__all__ = ["hanning", "hamming", "blackman", "bartlett", "STR_TO_WINDOW_FN", "stft", "istft", "ISTFTCache", "mel_filters"]


def _window(size, periodic, coeffs):
    denom = size if periodic else size - 1
    vals = [sum(a * math.cos(2 * math.pi * order * n / denom) for order, a in enumerate(coeffs)) for n in range(size)]
    return torch.tensor(vals, dtype=torch.float32)


@lru_cache(maxsize=None)
def hanning(size, periodic=False):
    """dsp.py:39-50."""
    return _window(size, periodic, (0.5, -0.5))


@lru_cache(maxsize=None)
def hamming(size, periodic=False):
    """dsp.py:53-64."""
    return _window(size, periodic, (0.54, -0.46))


@lru_cache(maxsize=None)
def blackman(size, periodic=False):
    """dsp.py:67-78."""
    return _window(size, periodic, (0.42, -0.5, 0.08))


@lru_cache(maxsize=None)
def bartlett(size, periodic=False):
    """dsp.py:81-85."""
    denom = size if periodic else size - 1
    return torch.tensor([1 - 2 * abs(n - denom / 2) / denom for n in range(size)], dtype=torch.float32)


STR_TO_WINDOW_FN = {"hann": hanning, "hanning": hanning, "hamming": hamming, "blackman": blackman, "bartlett": bartlett}


def _resolve_window(window, length, periodic_from_plus_one=False):
    if isinstance(window, str):
        fn = STR_TO_WINDOW_FN.get(window.lower())
        if fn is None:
            raise ValueError(f"Unknown window function: {window}")
        return fn(length + 1)[:-1] if periodic_from_plus_one else fn(length)
    return torch.as_tensor(window, dtype=torch.float32)


def stft(x, n_fft=800, hop_length=None, win_length=None, window="hann", center=True, pad_mode="reflect", device="cuda"):
    """dsp.py:385-433: 1-D signal -> complex64 [num_frames, n_fft//2+1] (a [B,n] batch -> [B,frames,freq])."""
    hop_length = n_fft // 4 if hop_length is None else hop_length
    win_length = n_fft if win_length is None else win_length
    w = _resolve_window(window, win_length)
    if w.shape[0] < n_fft:
        w = torch.cat([w.cpu(), torch.zeros(n_fft - w.shape[0])])
    if center and pad_mode not in ("constant", "reflect"):
        raise ValueError(f"Invalid pad_mode {pad_mode}")
    xt = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(device=device, dtype=torch.float32)
    squeeze = xt.dim() == 1
    xt = (xt[None] if squeeze else xt).contiguous()
    n = xt.shape[1]
    padded = n + (2 * (n_fft // 2) if center else 0)
    num_frames = 1 + (padded - n_fft) // hop_length
    if num_frames <= 0:
        raise ValueError(f"Input is too short (length={padded}) for n_fft={n_fft} with hop_length={hop_length} and center={center}.")
    mode = 0 if not center else (1 if pad_mode == "reflect" else 2)
    re, im = ops.stft(xt, w.to(xt.device), n_fft, hop_length, mode, num_frames)
    out = torch.complex(re, im)
    return out[0] if squeeze else out


def istft(x, hop_length=None, win_length=None, window="hann", center=True, length=None, normalized=False):
    """dsp.py:436-513: complex [n_freq, num_frames] -> real signal; string windows are periodic here (dsp.py:472)."""
    if win_length is None:
        win_length = (x.shape[1] - 1) * 2          # the reference reads the FRAME count here (dsp.py:465-466)
    if hop_length is None:
        hop_length = win_length // 4
    w = _resolve_window(window, win_length, periodic_from_plus_one=True)
    if w.shape[0] < win_length:
        w = torch.cat([w.cpu(), torch.zeros(win_length - w.shape[0])])
    n_freq, T = x.shape
    n_fft = (n_freq - 1) * 2
    if n_fft != win_length:
        raise ValueError(f"istft: n_freq={n_freq} implies n_fft={n_fft} but win_length={win_length}")
    re = x.real.to(torch.float32)[None].contiguous()
    im = x.imag.to(torch.float32)[None].contiguous()
    t = (T - 1) * hop_length + win_length
    trim = win_length // 2 if (center and length is None) else 0
    out_len = t - 2 * trim if (center and length is None) else t
    if length is not None:
        out_len = min(length, t)
    return ops.istft(re, im, n_fft, hop_length, w.to(re.device), norm_sq=normalized, clamp_mode=0, trim=trim, out_len=out_len)[0]


class ISTFTCache:
    """dsp.py:612-752: batched iSTFT with w^2 normalisation.  The reference caches index/normalisation buffers because
    its overlap-add is a scatter; ours is a gather kernel that needs neither, so the cache is a no-op kept for API parity."""

    def __init__(self):
        self.norm_buffer_cache = {}
        self.position_cache = {}

    def istft(self, real_part, imag_part, n_fft, hop_length, win_length, window, center=True, audio_length=None,
              constrain_value_range=False):
        if constrain_value_range:
            raise NotImplementedError("constrain_value_range is not on the accelerated path")
        w = torch.as_tensor(window, dtype=torch.float32)
        if w.shape[0] < n_fft:
            w = torch.cat([w.cpu(), torch.zeros(n_fft - w.shape[0])])
        re = real_part.to(torch.float32).contiguous()
        im = imag_part.to(torch.float32).contiguous()
        B, _, T = re.shape
        ola = (T - 1) * hop_length + n_fft
        trim = n_fft // 2 if center else 0
        out_len = ola - trim
        if audio_length is not None:
            out_len = min(out_len, audio_length)
        return ops.istft(re, im, n_fft, hop_length, w.to(re.device), norm_sq=True, clamp_mode=1, trim=trim, out_len=out_len)

    def clear_cache(self):
        self.norm_buffer_cache.clear()
        self.position_cache.clear()

    def cache_info(self):
        return {"norm_buffers": 0, "position_indices": 0, "total_cached_items": 0}


def _hz_to_mel(freq, mel_scale):
    if mel_scale == "htk":
        return 2595.0 * math.log10(1.0 + freq / 700.0)
    f_sp = 200.0 / 3
    if freq >= 1000.0:
        return 1000.0 / f_sp + math.log(freq / 1000.0) / (math.log(6.4) / 27.0)
    return freq / f_sp


@lru_cache(maxsize=None)
def mel_filters(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0, f_max: Optional[float] = None,
                norm: Optional[str] = None, mel_scale: str = "htk", precise: bool = False) -> torch.Tensor:
    """dsp.py:519-609 -> float32 [n_mels, n_fft//2+1]; any mel_scale other than "htk" is Slaney; the bin axis ends at
    ``sample_rate // 2``; float32 arithmetic unless ``precise``."""
    dt = torch.float64 if precise else torch.float32
    f_max = f_max or sample_rate / 2
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64).to(dt)
    m_pts = torch.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2, dtype=torch.float64).to(dt)
    if mel_scale == "htk":
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    else:
        f_sp = 200.0 / 3
        min_log_mel = 1000.0 / f_sp
        logstep = math.log(6.4) / 27.0
        f_pts = torch.where(m_pts >= min_log_mel, 1000.0 * torch.exp(logstep * (m_pts - min_log_mel)), f_sp * m_pts)
    f_pts = f_pts.to(dt)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0)
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb.T.contiguous().to(torch.float32)
