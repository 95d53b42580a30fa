"""Whisper log-mel frontend on B200 (reference: stt/models/whisper/audio.py).

Same constants and ``log_mel_spectrogram(audio, n_mels, padding)`` signature; the whole chain
(reflect pad, symmetric Hann, 400-point DFT, |.|^2, Slaney mel, log10, dynamic-range clamp, affine) is
two kernel launches (csrc/dsp.cu) and is batched over utterances, which the reference is not.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np
import torch

from .... import dsp, ops

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_FRAMES = N_SAMPLES // HOP_LENGTH
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def pad_or_trim(array: torch.Tensor, length: int = N_SAMPLES, *, axis: int = -1) -> torch.Tensor:
    """audio.py:24-38."""
    if array.shape[axis] > length:
        array = array.narrow(axis, 0, length)
    if array.shape[axis] < length:
        pad = [0, 0] * array.dim()
        pad[2 * (array.dim() - 1 - (axis % array.dim())) + 1] = length - array.shape[axis]
        array = torch.nn.functional.pad(array, pad)
    return array


@lru_cache(maxsize=None)
def _consts(n_mels: int, device: str):
    win = dsp.hanning(N_FFT).to(device)
    filt = dsp.mel_filters(SAMPLE_RATE, N_FFT, n_mels, norm="slaney", mel_scale=None).to(device)
    return win, filt


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, device="cuda") -> torch.Tensor:
    """audio [n] or [B, n] (16 kHz float) -> log-mel [n_frames, n_mels] or [B, n_frames, n_mels] (frames-major,
    as the reference returns despite its docstring, audio.py:62-63)."""
    if isinstance(audio, str):
        raise NotImplementedError("file decoding (audio_io) is outside the accelerated path; pass an array")
    if not isinstance(audio, torch.Tensor):
        audio = torch.as_tensor(np.asarray(audio))
    x = audio.to(device=device, dtype=torch.float32)
    squeeze = x.dim() == 1
    if squeeze:
        x = x[None]
    x = x.contiguous()
    n_total = x.shape[1] + padding
    frames = n_total // HOP_LENGTH          # 1 + n_total//hop STFT frames, last one dropped (audio.py:74)
    win, filt = _consts(n_mels, str(x.device))
    out = ops.whisper_logmel(x, padding, win, filt, frames)
    return out[0] if squeeze else out
