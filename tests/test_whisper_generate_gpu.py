"""Whisper long-form transcription on the GPU (Model.generate: 30-second windows, temperature fallback with the fused categorical
decode step, previous-text prompt, seek rule) against the float64 oracle's transcribe(), which tests/test_oracle_pins.py pins to the
reference's own Model.generate run.  Whisper-small widths and vocabulary, one encoder and two decoder layers (the oracle stays fast)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mlx_audio_b200 import synth
from mlx_audio_b200.configs import WHISPER_SMALL
from oracle import dsp as OD
from oracle import whisper as OW


class Stub:
    def decode(self, tokens):
        return " ".join(str(int(t)) for t in tokens)

    def encode(self, text):
        return []


def _setup():
    from mlx_audio_b200.stt.models.whisper import Model, ModelDimensions
    dims = dict(WHISPER_SMALL, n_audio_layer=1, n_text_layer=2)
    P = dict(synth.whisper_encoder_weights(dims))
    P.update(synth.whisper_decoder_weights(dims))
    model = Model(ModelDimensions.from_dict(dims), device="cuda:0").load_weights(P)
    return model, {k: v.double() for k, v in P.items()}, dims


def _audio(seconds, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(seconds * 16000) / 16000
    return (0.2 * np.sin(2 * np.pi * 330 * t) * (1 + np.sin(2 * np.pi * 0.4 * t)) + 0.05 * rng.standard_normal(t.shape)).astype(np.float32), rng


@pytest.mark.parametrize("case", ["fallback_and_prompt", "greedy_clips", "no_timestamps"])
def test_generate_matches_oracle_transcribe(case):
    from mlx_audio_b200.stt.models.whisper.whisper import TokenizerSpec
    model, P64, dims = _setup()
    audio, rng = _audio(47, seed=5)
    U = rng.random((32, 16)).astype(np.float32)
    kw = {"fallback_and_prompt": dict(temperature=(0.0, 0.5), logprob_threshold=-1.0, no_speech_threshold=0.6),     # every window falls back to t = 0.5
          "greedy_clips": dict(temperature=0.0, logprob_threshold=None, clip_timestamps="3,21,25,40", condition_on_previous_text=False),
          "no_timestamps": dict(temperature=(0.0,), logprob_threshold=None, return_timestamps=False, initial_prompt=[100, 101, 102])}[case]
    calls = [0]

    def uniforms(n_steps, batch):
        k = calls[0]
        calls[0] += 1
        return torch.as_tensor(U[k, :n_steps]).reshape(n_steps, 1).expand(n_steps, batch).contiguous()

    out = model.generate(audio, language="en", spec=TokenizerSpec(suppress=(11, 12)), tokenizer=Stub(), sample_len=12, uniforms=uniforms, **kw)
    okw = dict(kw)
    temps = okw.pop("temperature")
    okw["temperatures"] = (temps,) if isinstance(temps, float) else tuple(temps)
    if "clip_timestamps" in okw:
        okw["clip_timestamps"] = tuple(float(v) for v in okw["clip_timestamps"].split(","))
    okw["initial_prompt_tokens"] = tuple(okw.pop("initial_prompt", ()))
    okw.setdefault("compression_ratio_threshold", 2.4)
    mel = torch.as_tensor(OD.whisper_log_mel(audio, 80, padding=480000))
    text, segs = OW.transcribe(P64, mel, OW.TokenizerSpec(), dims, Stub(), suppress=(11, 12), sample_len=12, uniforms=lambda k: U[k].astype(np.float64), **okw)
    assert out.language == "en" and out.text == text and len(out.segments) == len(segs) > 0
    for a, b in zip(out.segments, segs):
        assert a["tokens"] == b["tokens"] and a["seek"] == b["seek"] and a["temperature"] == b["temperature"] and a["id"] == b["id"], (a, b)
        assert abs(a["start"] - b["start"]) < 1e-9 and abs(a["end"] - b["end"]) < 1e-9
        assert abs(a["avg_logprob"] - b["avg_logprob"]) < 2e-3 * max(1.0, abs(b["avg_logprob"]))
        assert abs(a["no_speech_prob"] - b["no_speech_prob"]) <= 2e-2 * b["no_speech_prob"] + 1e-30
        assert abs(a["compression_ratio"] - b["compression_ratio"]) < 1e-12
    if case == "fallback_and_prompt":
        assert calls[0] >= 2 and all(s["temperature"] == 0.5 for s in out.segments)


def test_generate_argument_errors_and_language_detection():
    from mlx_audio_b200.stt.models.whisper.whisper import TokenizerSpec
    model, P64, dims = _setup()
    audio, _ = _audio(4, seed=6)
    with pytest.raises(NotImplementedError):
        model.generate(audio, word_timestamps=True)
    with pytest.raises(ValueError, match="Unsupported language"):
        model.generate(audio, language="xx")
    # language=None on a multilingual vocabulary: one decoder pass on [sot], arg-max over the language tokens (decoding.py:20-77)
    mel = torch.as_tensor(OD.whisper_log_mel(audio, 80, padding=480000))[:3000]
    xa = OW.encoder(P64, mel[None], dims)
    lg, _ = OW.decoder_forward(P64, torch.tensor([[50258]]), xa, None, dims)
    want = int(lg[0, 0, 50259:50259 + 99].argmax())
    from mlx_audio_b200.stt.models.whisper.transcribe import LANGUAGE_CODES
    out = model.generate(audio, spec=TokenizerSpec(suppress=(11, 12)), tokenizer=Stub(), sample_len=4, temperature=0.0, logprob_threshold=None)
    assert out.language == LANGUAGE_CODES[want]
