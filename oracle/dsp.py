"""Oracle for the DSP frontend: windows, STFT, iSTFT, mel filterbank, log-mel.

TEST INFRASTRUCTURE (see oracle/__init__.py).  float64 NumPy unless stated.
Restates ``dsp.py``, ``stt/models/whisper/audio.py``, ``tts/models/interpolate.py``
and ``tts/models/qwen3_tts/qwen3_tts.py:64-120`` of the reference.
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------- windows


def _cos_window(size: int, periodic: bool, coeffs) -> np.ndarray:
    """dsp.py:39-79 -- generalised cosine windows built from Python floats."""
    denom = size if periodic else size - 1
    n = np.arange(size, dtype=np.float64)
    out = np.zeros(size, dtype=np.float64)
    for order, a in enumerate(coeffs):
        out += a * np.cos(2.0 * math.pi * order * n / denom)
    return out


def hanning(size: int, periodic: bool = False) -> np.ndarray:
    """dsp.py:39-50: 0.5*(1-cos(2 pi n/denom)); symmetric unless ``periodic``."""
    return _cos_window(size, periodic, (0.5, -0.5))


def hamming(size: int, periodic: bool = False) -> np.ndarray:
    """dsp.py:53-64."""
    return _cos_window(size, periodic, (0.54, -0.46))


def blackman(size: int, periodic: bool = False) -> np.ndarray:
    """dsp.py:67-78."""
    return _cos_window(size, periodic, (0.42, -0.5, 0.08))


def bartlett(size: int, periodic: bool = False) -> np.ndarray:
    """dsp.py:81-85."""
    denom = size if periodic else size - 1
    n = np.arange(size, dtype=np.float64)
    return 1.0 - 2.0 * np.abs(n - denom / 2.0) / denom


WINDOWS = {"hann": hanning, "hanning": hanning, "hamming": hamming,
           "blackman": blackman, "bartlett": bartlett}

# ----------------------------------------------------------------------------- stft / istft


def stft(x, n_fft=800, hop_length=None, win_length=None, window="hann",
         center=True, pad_mode="reflect") -> np.ndarray:
    """dsp.py:385-433.  1-D input -> complex [num_frames, n_fft//2+1].

    Quirk kept: a string window resolves to the SYMMETRIC window of
    ``win_length`` points (dsp.py:399-403), zero-padded on the right to n_fft.
    """
    x = np.asarray(x, dtype=np.float64)
    hop_length = n_fft // 4 if hop_length is None else hop_length
    win_length = n_fft if win_length is None else win_length
    if isinstance(window, str):
        fn = WINDOWS.get(window.lower())
        if fn is None:
            raise ValueError(f"Unknown window function: {window}")
        w = fn(win_length)
    else:
        w = np.asarray(window, dtype=np.float64)
    if w.shape[0] < n_fft:
        w = np.concatenate([w, np.zeros(n_fft - w.shape[0])])
    if center:
        p = n_fft // 2
        if pad_mode == "constant":
            x = np.pad(x, (p, p))
        elif pad_mode == "reflect":
            x = np.concatenate([x[1:p + 1][::-1], x, x[-(p + 1):-1][::-1]])
        else:
            raise ValueError(f"Invalid pad_mode {pad_mode}")
    num_frames = 1 + (x.shape[0] - n_fft) // hop_length
    if num_frames <= 0:
        raise ValueError(
            f"Input is too short (length={x.shape[0]}) for n_fft={n_fft} with "
            f"hop_length={hop_length} and center={center}.")
    idx = np.arange(num_frames)[:, None] * hop_length + np.arange(n_fft)[None, :]
    return np.fft.rfft(x[idx] * w[None, :], axis=-1)


def istft(x, hop_length=None, win_length=None, window="hann", center=True,
          length=None, normalized=False) -> np.ndarray:
    """dsp.py:436-513.  complex [n_freq, num_frames] -> real signal.

    Quirk kept: a string window is the PERIODIC one here (``fn(win+1)[:-1]``,
    dsp.py:472) although ``stft`` uses the symmetric one.
    """
    x = np.asarray(x)
    # Quirk kept: the default reads x.shape[1] (the FRAME count), dsp.py:465-466.
    win_length = (x.shape[1] - 1) * 2 if win_length is None else win_length
    hop_length = win_length // 4 if hop_length is None else hop_length
    if isinstance(window, str):
        fn = WINDOWS.get(window.lower())
        if fn is None:
            raise ValueError(f"Unknown window function: {window}")
        w = fn(win_length + 1)[:-1]
    else:
        w = np.asarray(window, dtype=np.float64)
    if w.shape[0] < win_length:
        w = np.concatenate([w, np.zeros(win_length - w.shape[0])])
    num_frames = x.shape[1]
    t = (num_frames - 1) * hop_length + win_length
    frames = np.fft.irfft(x, axis=0).T                      # [frames, win]
    recon = np.zeros(t)
    wsum = np.zeros(t)
    wn = w * w if normalized else w
    for f in range(num_frames):                             # scatter-add OLA
        recon[f * hop_length:f * hop_length + win_length] += frames[f] * w
        wsum[f * hop_length:f * hop_length + win_length] += wn
    recon = np.where(wsum > 1e-10, recon / np.where(wsum > 1e-10, wsum, 1.0), recon)
    if center and length is None:
        recon = recon[win_length // 2: -win_length // 2]
    if length is not None:
        recon = recon[:length]
    return recon


def istft_cache(real_part, imag_part, n_fft, hop_length, win_length, window,
                center=True, audio_length=None, constrain_value_range=False):
    """dsp.py:663-738 (ISTFTCache.istft): batched [B, freq, T] -> [B, samples]."""
    real_part = np.asarray(real_part, dtype=np.float64)
    imag_part = np.asarray(imag_part, dtype=np.float64)
    w = np.asarray(window, dtype=np.float64)
    if w.shape[0] < n_fft:
        w = np.concatenate([w, np.zeros(n_fft - w.shape[0])])
    spec = real_part + 1j * imag_part
    frames = np.fft.irfft(np.transpose(spec, (0, 2, 1)), n=n_fft, axis=-1)
    if constrain_value_range:
        frames = np.clip(frames, -w, w)
    frames = frames * w
    b, nf, fl = frames.shape
    ola = (nf - 1) * hop_length + fl
    out = np.zeros((b, ola))
    norm = np.zeros(ola)
    for f in range(nf):
        out[:, f * hop_length:f * hop_length + fl] += frames[:, f]
        norm[f * hop_length:f * hop_length + fl] += w * w
    out = out / np.maximum(norm, 1e-10)[None, :]
    if center:
        out = out[:, n_fft // 2:]
    if audio_length is not None:
        out = out[:, :audio_length]
    return out

# ----------------------------------------------------------------------------- mel


def _hz_to_mel(freq: float, mel_scale) -> float:
    """dsp.py:540-552: anything other than "htk" is the Slaney scale."""
    if mel_scale == "htk":
        return 2595.0 * math.log10(1.0 + freq / 700.0)
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    if freq >= min_log_hz:
        mels = min_log_hz / f_sp + math.log(freq / min_log_hz) / (math.log(6.4) / 27.0)
    return mels


def _mel_to_hz(mels: np.ndarray, mel_scale) -> np.ndarray:
    """dsp.py:554-569."""
    if mel_scale == "htk":
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_mel = 1000.0 / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(mels >= min_log_mel, 1000.0 * np.exp(logstep * (mels - min_log_mel)), freqs)


def mel_filters(sample_rate, n_fft, n_mels, f_min=0.0, f_max=None, norm=None,
                mel_scale="htk", precise=False, dtype=None) -> np.ndarray:
    """dsp.py:519-609 -> [n_mels, n_fft//2+1].

    Quirks kept: FFT-bin axis ends at ``sample_rate // 2`` (integer division,
    dsp.py:577); the default build runs in float32 (``dtype=np.float32``),
    ``precise=True`` in float64 then casts (dsp.py:605-609).
    """
    if dtype is None:
        dtype = np.float64 if precise else np.float32
    f_max = f_max or sample_rate / 2
    n_freqs = n_fft // 2 + 1
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs).astype(dtype)
    m_pts = np.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale),
                        n_mels + 2).astype(dtype)
    f_pts = _mel_to_hz(m_pts, mel_scale).astype(dtype)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(np.zeros_like(down), np.minimum(down, up))
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return np.ascontiguousarray(fb.T).astype(np.float32)

# ----------------------------------------------------------------------------- whisper / qwen3 frontends

WHISPER_SR, WHISPER_NFFT, WHISPER_HOP = 16000, 400, 160


def whisper_log_mel(audio, n_mels=80, padding=0) -> np.ndarray:
    """stt/models/whisper/audio.py:41-82 -> [n_frames, n_mels] (frames-major).

    symmetric hanning(400) (audio.py:72), reflect-centred STFT, drop last frame,
    |.|^2 @ slaney-filters^T, log10 clamp 1e-10, clamp to global max-8, (x+4)/4.
    """
    audio = np.asarray(audio, dtype=np.float64)
    if padding > 0:
        audio = np.pad(audio, (0, padding))
    spec = stft(audio, window=hanning(WHISPER_NFFT), n_fft=WHISPER_NFFT, hop_length=WHISPER_HOP)
    mag = np.abs(spec[:-1]) ** 2
    filt = mel_filters(WHISPER_SR, WHISPER_NFFT, n_mels, norm="slaney", mel_scale=None).astype(np.float64)
    mel = mag @ filt.T
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


def pad_or_trim(a: np.ndarray, length: int, axis: int = -1) -> np.ndarray:
    """audio.py:24-38."""
    if a.shape[axis] > length:
        a = np.take(a, np.arange(length), axis=axis)
    if a.shape[axis] < length:
        pw = [(0, 0)] * a.ndim
        pw[axis] = (0, length - a.shape[axis])
        a = np.pad(a, pw)
    return a


def qwen3_mel_spectrogram(audio, n_fft=1024, num_mels=128, sample_rate=24000, hop_size=256,
                          win_size=1024, fmin=0.0, fmax=12000.0) -> np.ndarray:
    """tts/models/qwen3_tts/qwen3_tts.py:64-120 -> [B, frames, n_mels]."""
    audio = np.asarray(audio, dtype=np.float64)
    if audio.ndim == 1:
        audio = audio[None, :]
    basis = mel_filters(sample_rate, n_fft, num_mels, fmin, fmax, norm="slaney",
                        mel_scale="slaney").astype(np.float64)
    pad = (n_fft - hop_size) // 2
    out = []
    for s in audio:
        s = np.concatenate([s[1:pad + 1][::-1], s, s[-(pad + 1):-1][::-1]])
        spec = stft(s, n_fft=n_fft, hop_length=hop_size, win_length=win_size, window="hann",
                    center=False)
        mag = np.sqrt(np.abs(spec) ** 2 + 1e-9)
        out.append(np.log(np.clip(mag @ basis.T, 1e-5, None)))
    return np.stack(out, 0)

# ----------------------------------------------------------------------------- interpolate


def interpolate1d(x, size, mode="linear", align_corners=None) -> np.ndarray:
    """tts/models/interpolate.py:61-117 on [N, C, W]."""
    x = np.asarray(x, dtype=np.float64)
    n, c, w_in = x.shape
    size = max(size, 1)
    if mode == "nearest":
        if size == 1:
            idx = np.array([0])
        else:
            idx = np.clip(np.floor(np.arange(size) * (w_in / size)).astype(np.int64), 0, w_in - 1)
        return x[:, :, idx]
    if align_corners and size > 1:
        pos = np.arange(size) * ((w_in - 1) / (size - 1))
    elif size == 1:
        pos = np.array([0.0])
    else:
        pos = np.arange(size) * (w_in / size)
        if not align_corners:
            pos = np.maximum(pos + 0.5 * (w_in / size) - 0.5, 0.0)
    if w_in == 1:
        return np.broadcast_to(x, (n, c, size)).copy()
    lo = np.floor(pos).astype(np.int64)
    hi = np.minimum(lo + 1, w_in - 1)
    fr = pos - lo
    return x[:, :, lo] * (1 - fr)[None, None] + x[:, :, hi] * fr[None, None]


def interpolate(x, size=None, scale_factor=None, mode="nearest", align_corners=None):
    """tts/models/interpolate.py:7-58 (1-D only; size = ceil(W*scale))."""
    x = np.asarray(x)
    if x.ndim != 3:
        raise ValueError("Only 1D interpolation currently supported")
    if (size is None) == (scale_factor is None):
        raise ValueError("exactly one of size / scale_factor")
    if size is None:
        size = max(1, int(math.ceil(float(x.shape[2]) * float(scale_factor))))
    return interpolate1d(x, int(size), mode, align_corners)

# ----------------------------------------------------------------------------- resample


def resample(audio, orig_sr, target_sr, axis=-1) -> np.ndarray:
    """resample.py:10-47: Kaiser(beta 14.77, 64 zero crossings, rolloff .9476) sinc FIR designed
    with scipy.signal.firwin, applied by scipy.signal.resample_poly(padtype="edge").

    SciPy is the reference's own arithmetic here (resample.py:7,21-25,40-47), so this
    restatement calls the same two SciPy routines with the same arguments.
    """
    from scipy import signal
    if orig_sr == target_sr:
        return np.asarray(audio)
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = target_sr // g, orig_sr // g
    mr = max(up, down)
    fir = signal.firwin(2 * 64 * mr + 1, 0.9475937167399596 / mr, window=("kaiser", 14.769656459379492))
    return signal.resample_poly(audio, up, down, axis=axis, window=fir, padtype="edge").astype(np.float32)
