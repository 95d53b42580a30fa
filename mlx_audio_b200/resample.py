"""Kaiser-sinc polyphase resampling (reference: resample.py:10-161).

The reference runs SciPy on the host; here torch CUDA tensors go through our polyphase kernel (csrc/dsp.cu:
resample_poly_kernel) with the SAME filter (SciPy ``firwin`` designs it once on the host, as the reference does) and the
same ``resample_poly(padtype="edge")`` indexing, float64 accumulation.  NumPy inputs keep the reference's host path.
Every output sample is an independent fixed-order sum, so chunked == whole-buffer holds bit-for-bit by construction
(the property the reference pins in tests/test_dsp.py:350-378)."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=None)
def _polyphase_filter(orig_sample_rate: int, sample_rate: int):
    """resample.py:10-26: (up, down, fir) with the kaiser_best-equivalent design."""
    from scipy import signal
    g = math.gcd(int(orig_sample_rate), int(sample_rate))
    up, down = sample_rate // g, orig_sample_rate // g
    mr = max(up, down)
    fir = signal.firwin(2 * 64 * mr + 1, 0.9475937167399596 / mr, window=("kaiser", 14.769656459379492))
    return up, down, fir


def _poly_plan(n_in: int, up: int, down: int, n_h: int):
    """The index bookkeeping of scipy.signal.resample_poly: (n_out, n_pre_pad, n_pre_remove)."""
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    half_len = (n_h - 1) // 2
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    return n_out, n_pre_pad, n_pre_remove


def resample_audio_array(audio, orig_sample_rate: int, sample_rate: int, axis: int = -1):
    """resample.py:29-47.  torch CUDA tensor -> GPU kernel; NumPy -> the reference's SciPy call."""
    if orig_sample_rate == sample_rate:
        return audio
    up, down, fir = _polyphase_filter(int(orig_sample_rate), int(sample_rate))
    try:
        import torch
        is_cuda = isinstance(audio, torch.Tensor) and audio.is_cuda
    except ImportError:                                        # pragma: no cover
        is_cuda = False
    if not is_cuda:
        from scipy import signal
        a = audio.detach().cpu().numpy() if not isinstance(audio, np.ndarray) and hasattr(audio, "detach") else np.asarray(audio)
        return signal.resample_poly(a, up, down, axis=axis, window=fir, padtype="edge").astype(np.float32, copy=False)
    from . import ops
    x = audio.to(torch.float32).movedim(axis, -1)
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).contiguous()
    n_out, n_pre_pad, n_pre_remove = _poly_plan(shp[-1], up, down, len(fir))
    h = torch.as_tensor(fir * up, dtype=torch.float64, device=x2.device)
    out = ops.resample_poly(x2, h, up, down, n_pre_pad, n_pre_remove, n_out)
    return out.reshape(*shp[:-1], n_out).movedim(-1, axis)


def resample_audio_chunks(chunks, orig_sample_rate: int, sample_rate: int, num_input_frames: int, chunk_duration_seconds: float = 1.0):
    """resample.py:50-161: time-first chunks -> the whole-buffer ``resample_poly`` result for the first ``num_input_frames``
    frames.  The reference overlaps chunks with a halo so that each retained region is bit-identical to the one-shot call; here
    every output sample is an independent fixed-order sum already, so the chunks are gathered (on the device when they are CUDA
    tensors) and resampled in one launch -- same samples, same edge rule, same errors."""
    if chunk_duration_seconds <= 0:
        raise ValueError("chunk_duration_seconds must be positive")
    try:
        import torch
    except ImportError:                                        # pragma: no cover
        torch = None
    parts = [c for c in chunks if c.shape[0] > 0]
    if not parts:
        return np.empty((0,), dtype=np.float32)
    cuda = torch is not None and isinstance(parts[0], torch.Tensor) and parts[0].is_cuda
    num_input_frames = max(0, int(num_input_frames))
    if num_input_frames == 0:
        shape = (0, *parts[0].shape[1:])
        return torch.empty(shape, dtype=torch.float32, device=parts[0].device) if cuda else np.empty(shape, dtype=np.float32)
    if cuda:
        whole = torch.cat([p.to(torch.float32) for p in parts], dim=0)[:num_input_frames]
    else:
        whole = np.concatenate([np.asarray(p, dtype=np.float32) for p in parts], axis=0)[:num_input_frames]
    if orig_sample_rate == sample_rate:
        return whole
    return resample_audio_array(whole, orig_sample_rate, sample_rate, axis=0)
