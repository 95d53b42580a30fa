"""Whisper frontend + encoder parity (BASELINE config 3 at a reduced batch) vs the float64 oracle.
Tolerances: log-mel 1e-4 abs; encoder output 1e-3 relative RMS."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mlx_audio_b200 import synth
from oracle import dsp as OD
from oracle import whisper as OW


def test_dsp_module_api_matches_reference_surface():
    from mlx_audio_b200 import dsp
    x = synth.whisper_audio(1, 12000, seed=1)[0]
    ref = OD.stft(x.numpy(), n_fft=800)
    y = dsp.stft(x)                                               # defaults: hop n_fft//4, symmetric hann, reflect
    assert y.shape == ref.shape and y.dtype == torch.complex64
    assert float(np.abs(y.cpu().numpy() - ref).max()) / np.abs(ref).max() < 2e-6
    spec = dsp.stft(x, n_fft=64, hop_length=16, window=dsp.hanning(64, periodic=True))
    rec = dsp.istft(spec.T.contiguous(), hop_length=16, win_length=64, window=dsp.hanning(64, periodic=True), normalized=True)
    assert float((rec.cpu() - x[: rec.shape[0]]).abs()[64:-64].max()) < 1e-5         # unity-gain round trip
    with pytest.raises(ValueError, match="Unknown window function"):
        dsp.stft(x, window="nope")
    with pytest.raises(ValueError, match="Input is too short"):
        dsp.stft(x[:10], n_fft=400, center=False)
    for kw in (dict(norm="slaney", mel_scale=None), dict(mel_scale="htk"), dict(norm="slaney", mel_scale="slaney", f_max=12000.0)):
        a = dsp.mel_filters(24000, 1024, 128, **kw).numpy()
        b = OD.mel_filters(24000, 1024, 128, **kw)
        assert np.abs(a - b).max() < 2e-5      # float32 pow in torch vs numpy


def test_whisper_encoder_parity_small_batch():
    from mlx_audio_b200.stt.models.whisper import Model, ModelDimensions
    dims = OW.WHISPER_SMALL
    P = synth.whisper_encoder_weights(dims)
    model = Model(ModelDimensions.from_dict(dims), device="cuda:0").load_weights(P)
    audio = synth.whisper_audio(2, 480000)
    mel_ref = np.stack([OD.whisper_log_mel(a.numpy(), 80, padding=480000)[:3000] for a in audio])
    from mlx_audio_b200.stt.models.whisper.audio import log_mel_spectrogram
    mel = log_mel_spectrogram(audio, 80, padding=480000, device="cuda:0")[:, :3000]
    assert mel.shape == (2, 3000, 80)
    assert float(np.abs(mel.cpu().numpy() - mel_ref).max()) < 1e-4
    ref = OW.encoder({k: v.double() for k, v in P.items()}, torch.as_tensor(mel_ref[:1]), dims)
    y = model.encode_audio(audio)
    assert y.shape == (2, 1500, 768)
    e = float(torch.sqrt(((y[:1].double().cpu() - ref) ** 2).mean()) / torch.sqrt((ref ** 2).mean()))
    assert e < 1e-3, e
