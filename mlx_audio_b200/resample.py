"""Kaiser-sinc polyphase resampling with the reference's filter design (resample.py:10-47).
Host-side SciPy in the reference as well; the GPU polyphase kernel is row next-2 of SURVEY.md section 8f."""
from __future__ import annotations

import math

import numpy as np


def _polyphase_filter(orig_sample_rate: int, sample_rate: int):
    from scipy import signal
    g = math.gcd(int(orig_sample_rate), int(sample_rate))
    up, down = sample_rate // g, orig_sample_rate // g
    mr = max(up, down)
    fir = signal.firwin(2 * 64 * mr + 1, 0.9475937167399596 / mr, window=("kaiser", 14.769656459379492))
    return up, down, fir


def resample_audio_array(audio: np.ndarray, orig_sample_rate: int, sample_rate: int, axis: int = -1) -> np.ndarray:
    from scipy import signal
    if orig_sample_rate == sample_rate:
        return audio
    up, down, fir = _polyphase_filter(orig_sample_rate, sample_rate)
    return signal.resample_poly(audio, up, down, axis=axis, window=fir, padtype="edge").astype(np.float32, copy=False)
