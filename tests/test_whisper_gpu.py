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


def test_whisper_greedy_step_rules_match_reference_filters():
    """Fused decode-step kernel vs the oracle's SuppressBlank/SuppressTokens/ApplyTimestampRules + GreedyDecoder on crafted
    token histories that hit every branch (first step, text after a timestamp pair, single timestamp, eot rows, all-masked)."""
    from mlx_audio_b200 import ops
    spec = OW.TokenizerSpec()
    V, tb, sb = 51865, spec.timestamp_begin, 3
    g = torch.Generator().manual_seed(0)
    hist = [
        list(spec.sot_sequence),                                          # first sampled position
        list(spec.sot_sequence) + [tb + 5],                               # one timestamp -> must be text
        list(spec.sot_sequence) + [tb + 5, 100],                          # text
        list(spec.sot_sequence) + [tb + 5, 100, tb + 20],                 # text then timestamp -> timestamps/eot only
        list(spec.sot_sequence) + [tb + 5, 100, tb + 20, tb + 20],        # pair closed -> text
        list(spec.sot_sequence) + [tb + 5, 100, spec.eot],                # finished row stays eot
    ]
    dev = torch.device("cuda:0")
    for scale, ts_boost in ((3.0, 0.0), (3.0, 6.0), (0.01, 0.0)):
        for h in hist:
            rows = [h, h]
            logits = torch.randn(2, V, generator=g) * scale
            logits[:, tb:] += ts_boost
            ref_l = OW.apply_filters(logits.double(), rows, spec, sb, (7, 8, 9), 50)
            ref_tok, _, ref_lp = OW.greedy_update(rows, ref_l, torch.zeros(2, dtype=torch.float64), spec.eot)
            tokens = torch.zeros(2, 64, dtype=torch.int64)
            tokens[:, :len(h)] = torch.tensor(h)
            sup = torch.zeros(V); sup[[7, 8, 9]] = float("-inf")
            blank = torch.zeros(V); blank[list(spec.blank_ids) + [spec.eot]] = float("-inf")
            slp = torch.zeros(2, device=dev)
            nd = torch.zeros(1, dtype=torch.int32, device=dev)
            nxt = ops.whisper_greedy_step(logits.to(dev), tokens.to(dev), len(h), sb, suppress_mask=sup.to(dev), blank_mask=blank.to(dev),
                                          eot=spec.eot, no_timestamps=spec.no_timestamps, timestamp_begin=tb, max_initial_ts=50,
                                          without_timestamps=False, sum_logprobs=slp, not_done=nd)
            assert nxt.cpu().tolist() == [t[-1] for t in ref_tok], (h, scale, ts_boost)
            assert torch.allclose(slp.cpu().double(), ref_lp, atol=2e-4, equal_nan=True), (slp, ref_lp)
            assert int(nd.item()) == sum(t[-1] != spec.eot for t in ref_tok)


def test_whisper_decoder_and_greedy_decode_parity():
    """TextDecoder logits (1e-3 rel) and the full greedy decode loop (token ids bit-exact) vs the float64 oracle."""
    from mlx_audio_b200.stt.models.whisper import Model, ModelDimensions
    from mlx_audio_b200.stt.models.whisper.whisper import TokenizerSpec
    dims = dict(OW.WHISPER_SMALL)
    dims["n_text_layer"] = 4                                              # full widths / vocab, fewer layers: oracle stays fast
    P = synth.whisper_decoder_weights(dims)
    model = Model(ModelDimensions.from_dict(dims), device="cuda:0").load_weights(P)
    P64 = {k: v.double() for k, v in P.items()}
    xa = torch.randn(2, 1500, 768, generator=torch.Generator().manual_seed(0))
    spec = OW.TokenizerSpec()
    tok0 = torch.tensor([list(spec.sot_sequence)] * 2)
    ref_logits, _ = OW.decoder_forward(P64, tok0, xa.double(), None, dims)
    cache = model.decoder.new_cache(xa)
    logits = model.decoder(tok0.cuda(), cache)
    e = float((logits.double().cpu() - ref_logits[:, -1]).abs().max() / ref_logits[:, -1].abs().max())
    assert e < 1e-3, e
    ref_tokens, ref_lp, ref_ns = OW.greedy_decode(P64, xa.double(), spec, sample_len=10, suppress=(11, 12), dims=dims)
    tokens, lp, ns = model.greedy_decode(xa, TokenizerSpec(suppress=(11, 12)), sample_len=10)
    assert tokens == ref_tokens
    assert torch.allclose(lp.cpu().double(), ref_lp, rtol=1e-3, atol=1e-3)
    assert torch.allclose(ns.cpu().double(), ref_ns, rtol=1e-2, atol=1e-30)
    res = model.decode(xa, TokenizerSpec(suppress=(11, 12)), sample_len=10)          # public result objects (decoding.py:634-722)
    for r, row, s_lp, s_ns in zip(res, ref_tokens, ref_lp.tolist(), ref_ns.tolist()):
        cut = (row + [spec.eot])[3:]
        cut = cut[:cut.index(spec.eot)]
        assert r.tokens == cut and abs(r.avg_logprob - s_lp / (len(cut) + 1)) < 1e-3 and abs(r.no_speech_prob - s_ns) <= 1e-2 * s_ns + 1e-30
        assert r.temperature == 0.0 and r.language == "en" and r.text == ""
    # without timestamps the initial sequence carries <|notimestamps|> (decoding.py:463-465) and the timestamp rules are off
    ref_tokens, ref_lp, _ = OW.greedy_decode(P64, xa.double(), spec, sample_len=6, suppress=(11, 12), dims=dims, without_timestamps=True)
    tokens, lp, _ = model.greedy_decode(xa, TokenizerSpec(suppress=(11, 12)), sample_len=6, without_timestamps=True)
    assert tokens == ref_tokens and tokens[0][3] == spec.no_timestamps
    assert torch.allclose(lp.cpu().double(), ref_lp, rtol=1e-3, atol=1e-3)


def test_dsp_frontend_matches_vectors_produced_by_the_reference_code():
    """The CUDA front end against tests/golden/dsp_golden.npz (the reference's dsp.py / audio.py run with NumPy standing in for MLX,
    tests/golden/make_dsp_golden.py): STFT configurations, iSTFT variants, mel filterbanks, log-mel incl. BASELINE config 1."""
    import os
    import numpy as np
    from mlx_audio import dsp
    from mlx_audio.stt.models.whisper.audio import log_mel_spectrogram
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dsp_golden.npz"))
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4000).astype(np.float32)
    cases = {"whisper": dict(n_fft=400, hop_length=160, window="hann"), "kokoro": dict(n_fft=20, hop_length=5, window="hann"),
             "const": dict(n_fft=256, hop_length=64, window="hamming", pad_mode="constant"), "nocenter": dict(n_fft=128, hop_length=32, center=False),
             "shortwin": dict(n_fft=512, hop_length=128, win_length=400)}
    for tag, kw in cases.items():
        s = dsp.stft(x, **kw).cpu().numpy()
        want = g[f"stft_{tag}_re"] + 1j * g[f"stft_{tag}_im"]
        assert s.shape == want.shape and np.abs(s - want).max() < 2e-4 * max(1.0, float(np.abs(want).max())), tag
    s = dsp.stft(x, n_fft=256, hop_length=64)
    for tag, kw in (("default", dict(hop_length=64, win_length=256)), ("len", dict(hop_length=64, win_length=256, length=3900)),
                    ("norm", dict(hop_length=64, win_length=256, normalized=True))):
        y, want = dsp.istft(s.T, **kw).cpu().numpy(), g[f"istft_{tag}"]
        ok = np.isfinite(want)
        assert y.shape == want.shape and np.abs(y[ok] - want[ok]).max() < 5e-5, tag
    for tag, kw in {"whisper80": dict(sample_rate=16000, n_fft=400, n_mels=80, norm="slaney", mel_scale=None),
                    "qwen3": dict(sample_rate=24000, n_fft=1024, n_mels=128, f_min=0.0, f_max=12000.0, norm="slaney", mel_scale="slaney"),
                    "htk": dict(sample_rate=22050, n_fft=512, n_mels=40, norm=None, mel_scale="htk")}.items():
        assert np.abs(np.asarray(torch.as_tensor(dsp.mel_filters(**kw)).cpu()) - g[f"mel_{tag}"]).max() < 5e-6, tag   # float32 pow/exp rounding of the band edges (unnormalised HTK weights reach 1.0)
    a = (0.1 * rng.standard_normal(16000)).astype(np.float32)
    assert np.abs(log_mel_spectrogram(a, n_mels=80, padding=0).cpu().numpy() - g["logmel_noise"]).max() < 2e-4
    assert np.abs(log_mel_spectrogram(a[:4000], n_mels=80, padding=8000).cpu().numpy() - g["logmel_noise_padded"]).max() < 2e-4
    sine = np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000.0).astype(np.float32)
    d = np.abs(log_mel_spectrogram(sine, n_mels=80, padding=0).cpu().numpy() - g["logmel_sine440"])
    assert d.max() < 2e-3 and np.mean(d) < 1e-4
