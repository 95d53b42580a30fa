"""Pin the CPU oracle against every known-answer vector the reference's own tests hold
for the hot path.  Expected values are copied from the cited
reference tests (data, not code)."""
import math
import numpy as np
import torch
import pytest

from oracle import dsp as O


def _noise():
    np.random.seed(42)
    return np.random.randn(12000).astype(np.float32)


SEL = [0, 1, 2, 63, 126, 127]


class TestQwen3MelSnapshot:
    """mlx_audio/tts/tests/test_qwen3_tts.py:175-353 (rtol/atol 2e-3 there; we hold 2e-5)."""
    TOL = dict(rtol=2e-5, atol=2e-5)

    def test_shape(self):
        assert O.qwen3_mel_spectrogram(_noise()).shape == (1, 46, 128)

    def test_frame0(self):
        mel = O.qwen3_mel_spectrogram(_noise())[0]
        exp = [-0.21803714, 0.06630915, -0.31858957, -0.02480409, -0.4512914, -0.5911693]
        np.testing.assert_allclose(mel[0, SEL], exp, **self.TOL)

    def test_frame23(self):
        mel = O.qwen3_mel_spectrogram(_noise())[0]
        exp = [0.08127937, 0.4368576, 0.43200976, -0.7714137, -0.24601418, 0.04274124]
        np.testing.assert_allclose(mel[23, SEL], exp, **self.TOL)

    def test_last_frame_reflect(self):
        mel = O.qwen3_mel_spectrogram(_noise())[0]
        exp = [-0.16861804, 0.0474052, -0.3970174, -0.01738772, -0.28846806, -0.10941511]
        np.testing.assert_allclose(mel[-1, SEL], exp, **self.TOL)

    def test_sine(self):
        t = np.arange(12000, dtype=np.float32) / 24000.0
        mel = O.qwen3_mel_spectrogram(np.sin(2 * np.pi * 1000 * t).astype(np.float32))[0]
        exp = [-1.2959518, -1.2937515, -1.2902284, -1.2074544, -0.9268621, -2.3822036, -5.331841, -5.33782]
        np.testing.assert_allclose(mel[0, [0, 1, 2, 10, 20, 63, 126, 127]], exp, rtol=2e-4, atol=2e-4)

    def test_stats(self):
        mel = O.qwen3_mel_spectrogram(_noise())
        np.testing.assert_allclose(mel.mean(), -0.37329558, atol=2e-5)
        np.testing.assert_allclose(mel.std(), 0.37445435, atol=2e-5)


def test_convtranspose_scatter_rule_pin():
    """tts/tests/test_istftnet_fidelity.py:18-31: depthwise ConvTranspose(k3,s2,p0)[1:] with
    w=[1,2,3], x=[1,2,3,4] -> [2,5,4,9,6,13,8,12] (scatter rule, no kernel flip)."""
    from oracle import nn as N
    x = np.array([1.0, 2, 3, 4]).reshape(1, 4, 1)
    w = np.array([1.0, 2, 3]).reshape(1, 3, 1)          # (Cout, K, Cin/g) MLX layout
    y = N.conv_transpose1d(x, w, stride=2, padding=0, groups=1)[:, 1:, :]
    np.testing.assert_allclose(y.reshape(-1), [2, 5, 4, 9, 6, 13, 8, 12], rtol=1e-12)


def test_kokoro_stft_roundtrip_pin():
    """tts/tests/test_istftnet_fidelity.py:34-46: MLXSTFT(20,5,20) transform->inverse is unity gain."""
    from oracle import kokoro as K
    t = np.arange(2000, dtype=np.float32)
    x = (0.5 * np.sin(2 * np.pi * 220 * t / 24000)).astype(np.float32)
    mag, ph = K.mlxstft_transform(x[None, :], 20, 5, 20)
    rec = K.mlxstft_inverse(mag, ph, 20, 5, 20).reshape(-1)[: x.shape[0]]
    np.testing.assert_allclose(rec[20:-20], x[20:-20], atol=1e-5)   # reference tolerance 1e-3


def test_interpolate_pins():
    """tts/tests/test_interpolate.py:40-91."""
    x = np.array([[[1.0, 2.0, 3.0, 4.0]]])
    np.testing.assert_allclose(O.interpolate1d(x, 8, "nearest"), [[[1, 1, 2, 2, 3, 3, 4, 4]]])
    np.testing.assert_allclose(O.interpolate1d(x, 2, "nearest"), [[[1, 3]]])
    x = np.array([[[1.0, 3.0, 5.0, 7.0]]])
    np.testing.assert_allclose(O.interpolate1d(x, 7, "linear", True), [[[1, 2, 3, 4, 5, 6, 7]]], rtol=1e-6)
    np.testing.assert_allclose(O.interpolate1d(x, 7, "linear", False),
                               [[[1.0, 1.7142857, 2.8571429, 4.0, 5.1428576, 6.2857141, 7.0]]], rtol=1e-6)
    np.testing.assert_allclose(O.interpolate1d(np.array([[[5.0]]]), 4, "linear"), [[[5, 5, 5, 5]]])
    assert O.interpolate(np.zeros((2, 3, 4)), scale_factor=2).shape == (2, 3, 8)


def test_sinegen_shapes_pin():
    """tts/tests/test_sinegen_length_alignment.py:9-17."""
    from oracle import kokoro as K
    f0 = np.ones((1, 2, 1)) * 120
    sw, uv, noise = K.sinegen(f0, upsample_scale=300, harmonic_num=8,
                              rand_ini=np.zeros((1, 9)), noise=np.zeros((1, 2, 9)))
    assert sw.shape == (1, 2, 9) and uv.shape == (1, 2, 1) and noise.shape == (1, 2, 9)


def test_resample_pins():
    """mlx_audio/tests/test_dsp.py:299-324: alias rejection < 0.01 RMS; passband gain 0.70-0.72."""
    orig, target = 24000, 16000
    t = np.arange(2 * orig) / orig
    out = O.resample(np.sin(2 * np.pi * 8200.0 * t).astype(np.float32), orig, target)
    assert float(np.sqrt(np.mean(out[400:-400] ** 2))) < 0.01
    for f in (1000.0, 7000.0):
        out = O.resample(np.sin(2 * np.pi * f * t).astype(np.float32), orig, target)
        assert 0.70 < float(np.sqrt(np.mean(out[400:-400] ** 2))) < 0.72
    assert abs(len(O.resample(np.zeros(24000, np.float32), 24000, 16000)) - 16000) <= 1


def test_config1_whisper_logmel_shapes():
    """BASELINE config 1: 1 s 16 kHz 440 Hz sine -> (100, 80); with padding=480000 -> (3100, 80)."""
    x = np.sin(2 * np.pi * 440 * np.arange(16000) / 16000).astype(np.float32)
    m = O.whisper_log_mel(x)
    assert m.shape == (100, 80)
    assert O.whisper_log_mel(x, padding=480000).shape == (3100, 80)
    assert np.isfinite(m).all() and m.max() <= (np.log10(np.abs(m).max() * 0 + 1e10) + 4) / 4


def test_stft_errors_match_reference_messages():
    """dsp.py:402,426-428."""
    with pytest.raises(ValueError, match="Unknown window function"):
        O.stft(np.zeros(1000), window="nope")
    with pytest.raises(ValueError, match="Input is too short"):
        O.stft(np.zeros(10), n_fft=400, center=False)


# ------------------------------------------------------------------------------------------------ Qwen3-TTS oracle pins
def _small_tokenizer_cfg():
    from oracle import qwen3 as Q
    return dict(Q.TOKENIZER_DECODER, latent_dim=64, codebook_dim=32, codebook_size=64, decoder_dim=48, hidden_size=32, intermediate_size=64,
                head_dim=16, num_attention_heads=2, num_key_value_heads=2, num_hidden_layers=2)


def test_qwen3_interleaved_mrope_pattern():
    """talker.py:139-184: slot i rotates with the H position iff i%3==1 and i<60, with W iff i%3==2 and i<60, else T
    (sections [24,20,20] on 64 frequency slots) -- checked against the angle each slot actually receives."""
    import torch
    from oracle import qwen3 as Q
    pos3 = torch.tensor([[[1]], [[10]], [[100]]])
    cos, sin = Q.mrope_cos_sin(pos3, 128, 1e6, [24, 20, 20])
    inv = 1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.float64) / 128))
    ang = torch.atan2(sin[0, 0, :64], cos[0, 0, :64])
    for i in range(64):
        want = 10 if (i % 3 == 1 and i < 60) else (100 if (i % 3 == 2 and i < 60) else 1)
        a = want * float(inv[i])
        assert abs(math.remainder(float(ang[i]) - a, 2 * math.pi)) < 1e-9, i
    assert torch.equal(cos[..., :64], cos[..., 64:])                   # emb = concat(freqs, freqs) (talker.py:220)
    c2, s2 = Q.mrope_cos_sin(torch.tensor([[7]]), 128, 1e6, [24, 20, 20])
    c1, s1 = Q.rope_cos_sin(torch.tensor([[7]]), 128, 1e6)
    assert torch.allclose(c1, c2) and torch.allclose(s1, s2)            # equal positions on the three axes = plain RoPE


def test_qwen3_sampler_filters_hand_cases():
    """lm/sample_utils.py:131-239 on logits whose answer is computable by hand."""
    import torch
    from oracle import qwen3 as Q
    lg = torch.log(torch.tensor([0.1, 0.2, 0.3, 0.4], dtype=torch.float64))
    assert torch.isinf(Q.apply_top_k(lg, 2)).tolist() == [True, True, False, False]
    # ascending cumulative probs 0.1, 0.3, 0.6, 1.0: keep those > 1 - top_p
    assert torch.isinf(Q.apply_top_p(lg, 0.5)).tolist() == [True, True, False, False]
    assert torch.isinf(Q.apply_top_p(lg, 0.75)).tolist() == [True, False, False, False]
    assert torch.isinf(Q.apply_min_p(lg, 0.6)).tolist() == [True, True, False, False]     # p < 0.6 * 0.4 removed
    # inverse CDF in index order: u*1.0 against cumulative 0.1, 0.3, 0.6, 1.0
    picks = [Q.sample_token(lg, u, temperature=1.0, top_k=0, top_p=1.0, repetition_penalty=1.0) for u in (0.05, 0.25, 0.59, 0.61, 0.999)]
    assert picks == [0, 1, 2, 3, 3]
    # sign-aware repetition penalty on the SET of generated ids (qwen3_tts.py:830-842), greedy
    lg2 = torch.tensor([2.0, 1.9, -1.0, -3.0], dtype=torch.float64)
    assert Q.sample_token(lg2, 0.0, temperature=0.0, repetition_penalty=1.2, generated_tokens=[0, 0, 0]) == 1     # 2.0/1.2 < 1.9
    tok, f = Q.sample_token(lg2, 0.0, temperature=0.0, repetition_penalty=2.0, generated_tokens=[2, 7], return_filtered=True)
    assert tok == 0 and float(f[2]) == -2.0                                    # negative logits are multiplied; id 7 >= V ignored
    assert Q.sample_token(lg2, 0.0, temperature=0.0, suppress_tokens=[0]) == 1


def test_qwen3_vocoder_is_causal_and_1920_per_frame():
    """speech_tokenizer.py:843-880,932-954: 1920 samples per code frame; every layer is causal, so decoding a prefix gives the
    prefix of the full decode, and chunked_decode with unlimited left context reproduces the one-shot decode."""
    import torch
    from mlx_audio_b200 import synth
    from oracle import qwen3 as Q
    cfg = _small_tokenizer_cfg()
    P = {k: v.double() for k, v in synth.qwen3_tokenizer_weights(cfg, seed=3).items()}
    codes = synth.qwen3_codes(cfg, 11, batch=2, seed=4)
    full = Q.tokenizer_decode(P, codes, cfg)
    assert full.shape == (2, 1, 11 * 1920) and float(full.abs().max()) <= 1.0
    part = Q.tokenizer_decode(P, codes[..., :7], cfg)
    assert float((part - full[..., :7 * 1920]).abs().max()) < 1e-12
    chunked = Q.chunked_decode(P, codes, chunk_size=4, left_context_size=100, cfg=cfg)
    assert chunked.shape == full.shape and float((chunked - full).abs().max()) < 1e-12
    wav, lengths = Q.speech_tokenizer_decode(P, codes.transpose(1, 2), cfg)
    assert wav.shape == (2, 11 * 1920) and lengths.tolist() == [11 * 1920, 11 * 1920]
    with pytest.raises(ValueError, match="Expected 16 layers of codes"):
        Q.tokenizer_decode(P, codes[:, :8], cfg)


def test_qwen3_frame_loop_contract():
    """qwen3_tts.py:1323-1404 on a tiny talker: rows are [first code, 15 predictor codes]; first codes never come from the
    suppressed top-1024 range; greedy decoding is independent of the uniforms; EOS ends the loop without emitting its frame."""
    import torch
    from mlx_audio_b200 import synth
    from oracle import qwen3 as Q
    cfg = dict(Q.TALKER, num_hidden_layers=1, cp_num_hidden_layers=1)
    P = {k[len("talker."):]: v.double() for k, v in synth.qwen3_talker_weights(cfg, seed=5).items()}
    g = torch.Generator().manual_seed(0)
    x, tr, pad = (torch.randn(1, n, 1024, generator=g, dtype=torch.float64) for n in (6, 2, 1))
    u = torch.rand(5, 16, generator=g, dtype=torch.float64)
    codes = Q.generate_codes(P, x, tr, pad, u, 5, cfg=cfg)
    assert codes.shape == (5, 16) and int(codes[:, 0].max()) < 2048 and int(codes[:, 1:].max()) < 2048
    g1 = Q.generate_codes(P, x, tr, pad, u, 4, temperature=0.0, cfg=cfg)
    g2 = Q.generate_codes(P, x, tr, pad, torch.rand(4, 16, dtype=torch.float64), 4, temperature=0.0, cfg=cfg)
    assert torch.equal(g1, g2)
    stop = Q.generate_codes(P, x, tr, pad, u, 5, cfg=dict(cfg, codec_eos_token_id=int(codes[2, 0])))
    assert stop.shape[0] == 2 and torch.equal(stop, codes[:2])


def _golden_resample():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resample_golden.npz"))
    cases = [("24k_16k", 0, (4801,), 24000, 16000, -1), ("16k_24k", 1, (3000,), 16000, 24000, -1), ("44k1_16k", 2, (8820,), 44100, 16000, -1),
             ("48k_16k_2d", 3, (2, 4800), 48000, 16000, -1), ("22k05_24k_axis0", 4, (2205, 2), 22050, 24000, 0),
             ("8k_16k_short", 5, (37,), 8000, 16000, -1)]
    return g, cases


def test_resampler_matches_reference_golden_vectors():
    """tests/golden/resample_golden.npz was produced by the REFERENCE's resample.py itself (it needs only NumPy/SciPy), see
    tests/golden/make_resample_golden.py: the oracle, the host path and the chunked entry point reproduce it."""
    from mlx_audio_b200.resample import resample_audio_array, resample_audio_chunks
    g, cases = _golden_resample()
    for name, seed, shape, osr, tsr, axis in cases:
        x = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
        want = g[name]
        assert O.resample(x, osr, tsr, axis=axis).shape == want.shape
        assert float(np.abs(O.resample(x, osr, tsr, axis=axis) - want).max()) <= 1e-7
        assert np.array_equal(resample_audio_array(x, osr, tsr, axis=axis), want)
        if name + "_chunks" in g.files:
            got = resample_audio_chunks(iter(np.array_split(x, 3, axis=0)), osr, tsr, x.shape[0], chunk_duration_seconds=0.05)
            assert got.shape == g[name + "_chunks"].shape and float(np.abs(got - g[name + "_chunks"]).max()) <= 1e-7
            assert float(np.abs(g[name + "_chunks"] - want).max()) <= 1e-7        # the reference's own chunk-invariance


def _dsp_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "dsp_golden.npz"))


_STFT_CASES = {"whisper": dict(n_fft=400, hop_length=160, window="hann"), "kokoro": dict(n_fft=20, hop_length=5, window="hann"),
               "const": dict(n_fft=256, hop_length=64, window="hamming", pad_mode="constant"), "nocenter": dict(n_fft=128, hop_length=32, center=False),
               "shortwin": dict(n_fft=512, hop_length=128, win_length=400)}
_MEL_CASES = {"whisper80": dict(sample_rate=16000, n_fft=400, n_mels=80, norm="slaney", mel_scale=None),
              "whisper128": dict(sample_rate=16000, n_fft=400, n_mels=128, norm="slaney", mel_scale=None),
              "qwen3": dict(sample_rate=24000, n_fft=1024, n_mels=128, f_min=0.0, f_max=12000.0, norm="slaney", mel_scale="slaney"),
              "htk": dict(sample_rate=22050, n_fft=512, n_mels=40, norm=None, mel_scale="htk")}


def test_oracle_dsp_matches_vectors_produced_by_the_reference_code():
    """tests/golden/dsp_golden.npz = the reference's dsp.py / whisper audio.py RUN here with NumPy standing in for the MLX primitives
    (tests/golden/make_dsp_golden.py, numpy_mlx_shim.py): windows, five STFT configurations, three iSTFT variants, four mel
    filterbanks and three log-mel spectrograms (incl. BASELINE config 1's 440 Hz sine) pin the oracle's restatement."""
    g = _dsp_golden()
    for name in ("hanning", "hamming", "blackman", "bartlett"):
        for size in (20, 400):
            assert np.abs(getattr(O, name)(size) - g[f"win_{name}_{size}"]).max() < 1e-6
            assert np.abs(getattr(O, name)(size, periodic=True) - g[f"win_{name}_{size}_periodic"]).max() < 1e-6
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4000).astype(np.float32)
    for tag, kw in _STFT_CASES.items():
        s = O.stft(x, **kw)
        want = g[f"stft_{tag}_re"] + 1j * g[f"stft_{tag}_im"]
        assert s.shape == want.shape and np.abs(s - want).max() < 2e-4 * max(1.0, float(np.abs(want).max())), tag
    s = O.stft(x, n_fft=256, hop_length=64)
    for tag, kw in (("default", dict(hop_length=64, win_length=256)), ("len", dict(hop_length=64, win_length=256, length=3900)),
                    ("norm", dict(hop_length=64, win_length=256, normalized=True))):
        y, want = O.istft(s.T, **kw), g[f"istft_{tag}"]
        ok = np.isfinite(want)                        # the reference divides 0/0 at uncovered edge samples (dsp.py:503-505)
        assert y.shape == want.shape and np.abs(y[ok] - want[ok]).max() < 2e-5, tag
    for tag, kw in _MEL_CASES.items():
        assert np.abs(O.mel_filters(**kw) - g[f"mel_{tag}"]).max() < 2e-6, tag
    a = (0.1 * rng.standard_normal(16000)).astype(np.float32)
    assert np.abs(O.whisper_log_mel(a, 80, 0) - g["logmel_noise"]).max() < 2e-4
    assert np.abs(O.whisper_log_mel(a[:4000], 80, 8000) - g["logmel_noise_padded"]).max() < 2e-4
    sine = np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000.0).astype(np.float32)
    d = np.abs(O.whisper_log_mel(sine, 80, 0) - g["logmel_sine440"])
    assert d.max() < 2e-3 and np.mean(d) < 1e-4          # clamped low-energy bins of a pure tone amplify float32 FFT noise


def _golden(name):
    import os
    import sys
    here = os.path.join(os.path.dirname(__file__), "golden")
    if here not in sys.path:
        sys.path.insert(0, here)
    import synth_params
    g = np.load(os.path.join(here, name), allow_pickle=False)
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["params"]).items()} if "params" in g.files else None
    return g, P


def test_oracle_whisper_matches_the_reference_model_code():
    """tests/golden/whisper_golden.npz = the reference's AudioEncoder / TextDecoder classes (whisper.py:338-498) EXECUTED in float64 with
    NumPy standing in for MLX (tests/golden/make_whisper_golden.py): encoder output, prefill logits, three kv-cache steps."""
    from oracle import whisper as OW
    g, P = _golden("whisper_golden.npz")
    dims = dict(n_mels=80, n_audio_ctx=60, n_audio_state=64, n_audio_head=4, n_audio_layer=2, n_vocab=300, n_text_ctx=32, n_text_state=64,
                n_text_head=4, n_text_layer=2)
    assert np.abs(OW.sinusoids(60, 64).numpy() - g["sinusoids"]).max() < 1e-12
    xa = OW.encoder(P, torch.as_tensor(g["mel"]), dims)
    assert np.abs(xa.numpy() - g["xa"]).max() < 1e-11
    lg, cache = OW.decoder_forward(P, torch.as_tensor(g["tokens"]), xa, None, dims)
    assert np.abs(lg.numpy() - g["logits"]).max() < 1e-11
    for i in range(g["step_tokens"].shape[1]):
        lg, cache = OW.decoder_forward(P, torch.as_tensor(g["step_tokens"][:, i:i + 1]), xa, cache, dims)
        assert np.abs(lg.numpy()[:, 0] - g["step_logits"][:, i]).max() < 1e-11


def test_oracle_qwen3_matches_the_reference_model_code():
    """tests/golden/qwen3_golden.npz = the reference's own Qwen3-TTS code (talker.py, speech_tokenizer.py, qwen3_tts.py Model.generate ->
    generate_custom_voice -> _generate_with_instruct, lm/sample_utils.py, lm/models/cache.py) EXECUTED in float64 at a reduced configuration
    with NumPy standing in for MLX (tests/golden/make_qwen3_golden.py).  Pins: talker prefill + cached steps (interleaved MRoPE, GQA),
    code predictor with its per-step heads and small_to_mtp projection, the 12.5 Hz decoder (one-shot, chunked, public decode with lengths),
    prompt assembly (speaker / language / instruct), three whole generations -- EOS-terminated sampling with repetition penalty, top-p,
    and greedy -- and one left-padded batch of three through Model.batch_generate, whose code matrices must be IDENTICAL and whose waveforms agree to float32 storage precision."""
    import json
    from oracle import qwen3 as Q
    g, _ = _golden("qwen3_golden.npz")
    import synth_params
    cfg, tcfg = json.loads(str(g["cfg"])), json.loads(str(g["tok_cfg"]))
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["talker_params"]).items()}
    PT = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["tok_params"]).items()}
    cache = Q.make_cache(cfg["num_hidden_layers"])
    lg, h = Q.talker_forward(P, torch.as_tensor(g["t_x"]), cache, cfg=cfg)
    assert np.abs(lg.numpy() - g["t_logits"]).max() < 1e-11 and np.abs(h.numpy() - g["t_hidden"]).max() < 1e-11
    for i in range(g["t_step_x"].shape[1]):
        lg, _ = Q.talker_forward(P, torch.as_tensor(g["t_step_x"][:, i:i + 1]), cache, cfg=cfg)
        assert np.abs(lg.numpy()[:, 0] - g["t_step_logits"][:, i]).max() < 1e-11
    cc = Q.make_cache(cfg["cp_num_hidden_layers"])
    cl = Q.code_predictor_forward(P, torch.as_tensor(g["cp_x"]), cc, 0, cfg)
    assert np.abs(cl.numpy()[:, -1] - g["cp_logits"][:, 0]).max() < 1e-11
    for i in range(2):
        cl = Q.code_predictor_forward(P, torch.as_tensor(g["cp_x2"][:, i:i + 1]), cc, i + 1, cfg)
        assert np.abs(cl.numpy()[:, -1] - g["cp_logits"][:, i + 1]).max() < 1e-11
    wtol = 2e-7                                                         # waveforms are stored as float32 in the fixture
    assert np.abs(Q.tokenizer_decode(PT, torch.as_tensor(g["tok_codes"]), tcfg).numpy() - g["tok_wav"]).max() < wtol
    assert np.abs(Q.chunked_decode(PT, torch.as_tensor(g["tok_codes"]), 4, 2, tcfg).numpy() - g["tok_wav_chunked"]).max() < wtol
    w, ln = Q.speech_tokenizer_decode(PT, torch.as_tensor(g["tok_codes_bt"]), tcfg)
    assert np.abs(w.numpy() - g["tok_decode_wav"]).max() < wtol and ln.tolist() == g["tok_decode_lens"].tolist()
    # the incremental decoder (streaming_step, speech_tokenizer.py:889-930) fed 5 + 4 frames: exact everywhere except that the transposed-conv
    # bias is overlap-added twice after the boundary -- restated by the oracle, and clearly not the one-shot decode
    first = torch.as_tensor(g["tok_codes"][:1])
    assert np.abs(Q.tokenizer_decode(PT, first, tcfg, stream_boundaries=(5,)).numpy() - g["tok_stream_wav"]).max() < wtol
    assert np.abs(Q.tokenizer_decode(PT, first, tcfg).numpy() - g["tok_stream_wav"]).max() > 0.1
    P["codec_head.weight"] = P["codec_head.weight"].clone()
    P["codec_head.weight"][cfg["codec_eos_token_id"]] *= float(g["gen_eos_gain"])
    ids = dict(codec_nothink_id=1004, codec_think_id=1003, codec_think_bos_id=1005, codec_think_eos_id=1006, codec_pad_id=1001, codec_bos_id=1002)
    lang, spk = {"english": 1010, "german": 1011}, {"amy": 1020, "bob": 1021}
    for t in "abc":
        m = json.loads(str(g[f"gen_{t}_meta"]))
        ie, tr, pad = Q.prepare_generation_inputs_from_ids(P, m["text_ids"], (112, 113, 111), ids, lang.get(m["lang_code"]), spk[m["voice"].lower()],
                                                           m["instruct_ids"])
        assert np.abs(ie.numpy() - g[f"gen_{t}_input_embeds"]).max() < 1e-12 and np.abs(tr.numpy() - g[f"gen_{t}_trailing"]).max() < 1e-12
        codes = Q.generate_codes(P, ie, tr, pad, torch.as_tensor(g[f"gen_{t}_u"]), m["max_tokens"], temperature=m.get("temperature", 0.9),
                                 top_p=m.get("top_p", 1.0), cfg=cfg)
        assert np.array_equal(codes.numpy(), g[f"gen_{t}_codes"]), t
        wav, ln = Q.speech_tokenizer_decode(PT, codes[None], tcfg)
        assert int(ln[0]) == g[f"gen_{t}_audio"].shape[0] and np.abs(wav[0, :int(ln[0])].numpy() - g[f"gen_{t}_audio"]).max() < wtol
    assert json.loads(str(g["gen_a_meta"]))["draws_left"] > 0           # case a stopped on EOS, not on max_tokens
    # Model.batch_generate's own loop (qwen3_tts.py:1800-1935): three prompts of different length (left padding 0 / 18 / 10), one row
    # reaching EOS six frames before the others stop -> identical code matrices per row
    m = json.loads(str(g["batch_meta"]))
    calls, rows = list(m["tokenizer_calls"]), []
    for v, i in zip(m["voices"], m["instructs"]):
        tid = calls.pop(0)
        rows.append(Q.prepare_generation_inputs_from_ids(P, tid, (112, 113, 111), ids, lang[m["lang_code"]], spk[v], calls.pop(0) if i else None))
    assert len({r[0].shape[1] for r in rows}) == 3
    got = Q.generate_codes_batch(P, [r[0] for r in rows], [r[1] for r in rows], rows[0][2], torch.as_tensor(g["batch_u"]), m["max_tokens"], cfg=cfg)
    assert len({int(o.shape[0]) for o in got}) > 1
    for b, o in enumerate(got):
        assert np.array_equal(o.numpy(), g[f"batch_codes_{b}"]), b
        wav, _ = Q.speech_tokenizer_decode(PT, o[None], tcfg)
        assert np.abs(wav[0].numpy() - g[f"batch_audio_{b}"]).max() < wtol


def test_oracle_codecs_match_the_reference_model_code():
    """tests/golden/codec_golden.npz = the reference's SNAC.decode (snac/{snac,layers,vq}.py) and Mimi.decode (mimi/mimi.py + modules/) EXECUTED
    in float64 at reduced configurations with NumPy standing in for MLX (tests/golden/make_codec_golden.py).  SNAC: strides 8/3/4/2 (the
    output_padding<-groups argument quirk gives 3867 samples for 5 coarse frames), per-channel noise injected.  Mimi: split RVQ with the
    embedding_sum / cluster_usage codebooks, depthwise transposed-conv upsampling, rope transformer with a 6-step attention context over 18
    steps, SEANet decoder.  Waveforms are stored as float32 in the fixture, hence 2e-7.  The encode side of both codecs (SEANet / SNAC encoder, residual
    vector quantisation by nearest code) is pinned too: identical code streams."""
    import json
    from oracle import codec as OC
    g, _ = _golden("codec_golden.npz")
    import synth_params
    cfg = json.loads(str(g["snac_cfg"]))
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["snac_params"]).items()}
    codes = [torch.as_tensor(g[f"snac_codes_{i}"]).long() for i in range(3)]
    noises = [torch.as_tensor(g[f"snac_noise_{i}"]) for i in range(4)]
    y = OC.snac_decode(P, codes, cfg, noises)
    assert tuple(y.shape) == g["snac_audio"].shape == (2, 3867, 1) and np.abs(y.numpy() - g["snac_audio"]).max() < 2e-7
    cfg = json.loads(str(g["mimi_cfg"]))
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["mimi_params"]).items()}
    y = OC.mimi_decode(P, torch.as_tensor(g["mimi_codes"]).long(), cfg)
    assert tuple(y.shape) == g["mimi_pcm"].shape == (2, 1, 9 * 1920) and np.abs(y.numpy() - g["mimi_pcm"]).max() < 2e-7
    # SNAC.decode_stream (snac.py:106-162): a first call, then a call that prepends 2 / 4 / 8 context frames per level; the reference's
    # context trim slices the channel axis of the [B, T, 1] audio, so the second call returns context + new audio (3099 samples) -- kept
    scfg = json.loads(str(g["snac_cfg"]))
    PS = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["snac_params"]).items()}
    c1, c2 = ([torch.as_tensor(g[f"snac_stream_c{j}_{i}"]).long() for i in range(3)] for j in (1, 2))
    zs = [torch.as_tensor(g[f"snac_stream_noise_{i}"]) for i in range(8)]
    a1, ctx = OC.snac_decode_stream(PS, c1, None, 8, scfg, zs[:4])
    a2, ctx2 = OC.snac_decode_stream(PS, c2, ctx, 8, scfg, zs[4:])
    assert tuple(a1.shape) == (2, 2331, 1) and tuple(a2.shape) == (2, 3099, 1)
    assert np.abs(a1.numpy() - g["snac_stream_audio1"]).max() < 2e-7 and np.abs(a2.numpy() - g["snac_stream_audio2"]).max() < 2e-7
    assert all(np.array_equal(c.numpy(), g[f"snac_stream_ctx_{i}"]) for i, c in enumerate(ctx2))
    # encode side (integer results, identical): Mimi.encode on 12 frames + 700 samples, SNAC.encode on a length that needs right padding
    c = OC.mimi_encode(P, torch.as_tensor(g["mimi_enc_pcm"]), cfg)
    assert tuple(c.shape) == (2, 4, 13) and np.array_equal(c.numpy(), g["mimi_enc_codes"])
    cfg = json.loads(str(g["snac_cfg"]))
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["snac_params"]).items()}
    enc = OC.snac_encode(P, torch.as_tensor(g["snac_enc_audio"]), cfg)
    assert [tuple(e.shape) for e in enc] == [(2, 3), (2, 6), (2, 12)]
    for i, e in enumerate(enc):
        assert np.array_equal(e.numpy(), g[f"snac_enc_codes_{i}"]), i


def test_oracle_kokoro_matches_the_reference_model_code():
    """tests/golden/kokoro_golden.npz = the reference's Kokoro Model.__call__ (kokoro.py:111-177 over modules.py, istftnet.py, interpolate.py,
    dsp.py) EXECUTED in float64 on the public 82M configuration with NumPy standing in for MLX (tests/golden/make_kokoro_golden.py): ALBERT,
    duration encoder + LSTMs, duration rounding and alignment, F0/N predictor, text encoder, AdaIN decoder, harmonic source (voiced and
    unvoiced frames), noise convs, upsampling generator, iSTFT head.  The generator also asserts that the reference module tree and
    synth.kokoro_weights() name exactly the same 548 parameters.  The oracle runs with exact (unrounded) effective weights here, the one
    deliberate difference from its default (bf16 checkpoint dtype); the waveform is stored as float32, hence 2e-7."""
    import json
    from mlx_audio_b200 import synth
    from oracle import kokoro as OK
    g, _ = _golden("kokoro_golden.npz")
    m = json.loads(str(g["meta"]))
    cfg = OK.KOKORO_CONFIG
    P = {k: v.double() for k, v in synth.kokoro_weights(cfg, seed=0).items()}
    P["predictor.F0_proj.weight"] = P["predictor.F0_proj.weight"] * m["f0_gain"]
    rng = np.random.default_rng(71)
    rand_ini = rng.random((1, 9))
    assert np.array_equal(rand_ini, g["rand_ini"])
    noise = rng.standard_normal(tuple(int(v) for v in g["noise_shape"])).astype(np.float32).astype(np.float64)
    OK.EFFECTIVE_WEIGHTS_BF16, OK.TAP = False, {}
    try:
        audio, pd = OK.forward(P, torch.as_tensor(g["ids"])[None], torch.as_tensor(g["ref_s"]), cfg, speed=m["speed"], rand_ini=torch.as_tensor(rand_ini),
                               noise=torch.as_tensor(noise))
        voiced = float((OK.TAP["F0"] > 10).double().mean())
    finally:
        OK.EFFECTIVE_WEIGHTS_BF16, OK.TAP = True, None
    assert 0.05 < voiced < 0.95                                          # both branches of the harmonic source are exercised
    assert np.array_equal(pd.numpy(), g["pred_dur"])
    a, w = audio.numpy().reshape(-1), g["audio"].reshape(-1)
    assert a.shape == w.shape == (43200,) and np.abs(a - w).max() < 2e-7


def test_oracle_whisper_decoding_matches_the_reference_decoding_code():
    """tests/golden/whisper_golden.npz (dec_* / filt_* entries) = the reference's decoding.py EXECUTED through the NumPy MLX stand-in with a
    stub tokenizer on a 300-entry vocabulary laid out like Whisper's: SuppressBlank, SuppressTokens and ApplyTimestampRules applied to random
    logits under hand-built token histories (first step, timestamp pairs, text after a pair, EOT rows, the probability-mass rule), then
    GreedyDecoder.update; and DecodingTask._main_loop for 14 tokens with and without timestamps.  -inf patterns must be identical.
    This run is what showed that get_suppress_tokens() always adds the task / sot markers and no_speech (decoding.py:100-112)."""
    from oracle import whisper as OW
    g, P = _golden("whisper_golden.npz")
    dims = dict(n_mels=80, n_audio_ctx=60, n_audio_state=64, n_audio_head=4, n_audio_layer=2, n_vocab=300, n_text_ctx=32, n_text_state=64,
                n_text_head=4, n_text_layer=2)
    spec = OW.TokenizerSpec(eot=200, sot=201, no_timestamps=208, timestamp_begin=209, no_speech=207, blank_ids=(7,), language=202, task=203,
                            transcribe=203, translate=204, sot_lm=205, sot_prev=206)
    sup = g["dec_suppress"].tolist()
    assert OW.get_suppress_tokens(spec, sup) == (3, 4, 5, 201, 203, 204, 205, 206, 207, 250)
    for name in ("first", "mixed", "one", "two"):
        toks = g[f"filt_{name}_tokens"].tolist()
        y = OW.apply_filters(torch.as_tensor(g[f"filt_{name}_logits"]), toks, spec, 3, sup, max_initial_timestamp_index=2).numpy()
        want = g[f"filt_{name}_out"]
        fin = np.isfinite(want)
        assert np.array_equal(np.isinf(y), np.isinf(want)) and np.abs(y[fin] - want[fin]).max() < 1e-12, name
        nt, comp, slp = OW.greedy_update(toks, torch.as_tensor(y), torch.zeros(len(toks), dtype=torch.float64), spec.eot)
        assert np.array_equal(np.array(nt), g[f"filt_{name}_next"]) and comp == bool(g[f"filt_{name}_completed"])
        assert np.allclose(slp.numpy(), g[f"filt_{name}_sum_logprobs"], rtol=0, atol=1e-12, equal_nan=True)
    P["decoder.token_embedding.weight"] = P["decoder.token_embedding.weight"].clone()
    P["decoder.token_embedding.weight"][:200] *= float(g["dec_text_gain"])
    xa = torch.as_tensor(g["xa"])
    for tag, wt in (("ts", False), ("nots", True)):
        tok, slp, ns = OW.greedy_decode(P, xa, spec, sample_len=14, suppress=sup, dims=dims, max_initial_timestamp_index=2, without_timestamps=wt)
        assert np.array_equal(np.array(tok), g[f"dec_{tag}_tokens"]), tag
        assert np.abs(slp.numpy() - g[f"dec_{tag}_sum_logprobs"]).max() < 1e-11 and np.abs(ns.numpy() - g[f"dec_{tag}_no_speech"]).max() < 1e-12


def test_oracle_qwen3_default_batch_path_matches_the_reference_session():
    """qwen3_golden.npz session_* = the reference's DEFAULT batch path, Model.batch_generate(stream=False) -> Qwen3TTSBatchSession
    (continuous_batching.py) EXECUTED through the stand-in with per-row uniform streams: rows leave the batch when they finish (11, 15, 20
    frames), BatchKVCache rows are extracted / merged every step, and every row is decoded by _decode_generated_codes.  Each row must equal
    what the single-sequence loop generates from the same stream, and its audio the 15-frame / 5-context chunked decode."""
    import json
    from oracle import qwen3 as Q
    g, _ = _golden("qwen3_golden.npz")
    import synth_params
    cfg, tcfg = json.loads(str(g["cfg"])), json.loads(str(g["tok_cfg"]))
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["talker_params"]).items()}
    PT = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["tok_params"]).items()}
    P["codec_head.weight"] = P["codec_head.weight"].clone()
    P["codec_head.weight"][cfg["codec_eos_token_id"]] *= float(g["gen_eos_gain"])
    ids = dict(codec_nothink_id=1004, codec_think_id=1003, codec_think_bos_id=1005, codec_think_eos_id=1006, codec_pad_id=1001, codec_bos_id=1002)
    lang, spk = {"english": 1010, "german": 1011}, {"amy": 1020, "bob": 1021}
    m = json.loads(str(g["session_meta"]))
    calls, us = list(m["tokenizer_calls"]), torch.as_tensor(g["session_u"])
    lengths = []
    for b, (v, i) in enumerate(zip(m["voices"], m["instructs"])):
        tid = calls.pop(0)
        ie, tr, pad = Q.prepare_generation_inputs_from_ids(P, tid, (112, 113, 111), ids, lang[m["lang_code"]], spk[v], calls.pop(0) if i else None)
        codes = Q.generate_codes(P, ie, tr, pad, us[b], m["max_tokens"], cfg=cfg)
        assert np.array_equal(codes.numpy(), g[f"session_codes_{b}"]), b
        assert np.abs(Q.decode_generated_codes(PT, codes, tcfg).numpy() - g[f"session_audio_{b}"]).max() < 2e-7
        # the same audio is chunked_decode with (15, 5) in place of (300, 25) -- which is how the product computes it
        assert np.abs(Q.chunked_decode(PT, codes.T[None], 15, 5, tcfg)[0, 0].numpy() - g[f"session_audio_{b}"]).max() < 2e-7
        lengths.append(codes.shape[0])
    assert lengths == [11, 15, 20]
    one_shot = Q.tokenizer_decode(PT, torch.as_tensor(g["session_codes_2"]).long().T[None], tcfg)[0, 0]
    assert np.abs(one_shot.numpy() - g["session_audio_2"]).max() > 0.1          # ... and NOT a one-shot decode of the 20-frame row


def test_oracle_qwen3_voice_cloning_matches_the_reference_model_code():
    """qwen3_golden.npz, voice-cloning entries = the reference's own code EXECUTED (make_qwen3_golden.py): the ECAPA-TDNN speaker encoder on
    the 24 kHz mel front end (speaker_encoder.py, qwen3_tts.py:64-121,285-324), the speech-tokenizer ENCODER (speech_tokenizer.py:957-1058:
    Mimi SEANet encoder, full-causal half-split-RoPE transformer, replicate-padded stride-2 conv, split RVQ, first 16 of 20 code books) and
    two whole ``Model.generate(text, ref_audio, ref_text)`` runs on a base model (-> _generate_icl, qwen3_tts.py:2200-2510): in-context prompt,
    frame loop with repetition penalty 1.5, joint decode of [reference | generated] codes with the reference's share cut off -- one run
    reaching EOS, one whose first-code-book zero shortens the valid length (the ``codes > 0`` rule of speech_tokenizer.py:1113-1116)."""
    import json
    from oracle import dsp as D
    from oracle import qwen3 as Q
    g, _ = _golden("qwen3_golden.npz")
    import synth_params
    cfg, tcfg, ecfg, scfg = (json.loads(str(g[k])) for k in ("cfg", "tok_cfg", "tok_enc_cfg", "spk_cfg"))
    P, PT, PS = ({k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g[n]).items()} for n in ("talker_params", "tok_params", "spk_params"))
    audio = 0.4 * np.random.default_rng(133).standard_normal((2, 1, 5 * 1920 + 300))
    codes = Q.tokenizer_encode(PT, torch.as_tensor(audio), ecfg)
    assert tuple(codes.shape) == (2, 16, 6) and np.array_equal(codes.numpy(), g["tok_enc_codes"])
    spk_audio = 0.3 * np.random.default_rng(135).standard_normal((2, 9000))
    assert np.abs(np.asarray(D.qwen3_mel_spectrogram(spk_audio)) - g["spk_mel"]).max() < 5e-5     # the oracle's filterbank emulates float32
    assert np.abs(Q.speaker_encoder(PS, torch.as_tensor(g["spk_mel"]), scfg).numpy() - g["spk_embedding"]).max() < 1e-12
    P["codec_head.weight"] = P["codec_head.weight"].clone()
    P["codec_head.weight"][cfg["codec_eos_token_id"]] *= float(g["gen_eos_gain"])
    ids = dict(codec_nothink_id=1004, codec_think_id=1003, codec_think_bos_id=1005, codec_think_eos_id=1006, codec_pad_id=1001, codec_bos_id=1002)
    lengths = []
    for t in "ab":
        m = json.loads(str(g[f"icl_{t}_meta"]))
        rng = np.random.default_rng(m["seed"])
        ref_audio, us = 0.3 * rng.standard_normal(3 * 1920 + 500), rng.random((m["max_tokens"], 4))
        rc = Q.tokenizer_encode(PT, torch.as_tensor(ref_audio)[None, None], dict(ecfg, nq=m["enc_nq"]))
        assert np.array_equal(rc.numpy(), g[f"icl_{t}_ref_codes"])
        se = Q.speaker_encoder(PS, torch.as_tensor(np.asarray(D.qwen3_mel_spectrogram(ref_audio))).double(), scfg)
        assert np.abs(se.numpy() - g[f"icl_{t}_speaker_embed"]).max() < 1e-5
        ie, tr, pad = Q.prepare_icl_generation_inputs_from_ids(P, m["target_ids"], m["ref_ids"], rc, (112, 113, 111), ids, 1011,
                                                               torch.as_tensor(g[f"icl_{t}_speaker_embed"]), cfg)
        assert np.abs(ie.numpy() - g[f"icl_{t}_input_embeds"]).max() < 1e-12
        gen = Q.generate_codes(P, ie, tr, pad, torch.as_tensor(us), m["max_tokens"], repetition_penalty=m["repetition_penalty"], cfg=cfg)
        wav = Q.decode_icl_generated_codes(PT, gen, rc, tcfg).numpy()
        assert gen.shape[0] == m["token_count"] and wav.shape == g[f"icl_{t}_audio"].shape and np.abs(wav - g[f"icl_{t}_audio"]).max() < 2e-7
        lengths.append(wav.shape[0])
    assert lengths == [17829, 7680]


WHISPER_GEN_DIMS = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=32, n_audio_head=2, n_audio_layer=1, n_vocab=300, n_text_ctx=48, n_text_state=32,
                        n_text_head=2, n_text_layer=2)


def whisper_generate_audio():
    """The 75-second synthetic recording of tests/golden/make_whisper_golden.py:generate_case, rebuilt from its formula (seed 33), and the
    uniform table that follows it in the same random stream."""
    rng = np.random.default_rng(33)
    sr = 16000
    t = np.arange(75 * sr) / sr
    audio = (0.2 * np.sin(2 * np.pi * 220 * t) * (1 + np.sin(2 * np.pi * 0.3 * t)) + 0.05 * rng.standard_normal(t.shape)).astype(np.float32)
    audio[int(31 * sr):int(58 * sr)] *= 1e-3
    U = rng.random((64, WHISPER_GEN_DIMS["n_text_ctx"] // 2 + 4))
    return audio, U


class WhisperStubTokenizer:
    """decode / encode of the stub tokenizer the reference ran with (make_whisper_golden.py:StubTokenizer)."""

    def __init__(self, prompt=None):
        self.prompt = prompt

    def decode(self, tokens):
        return " ".join(str(int(t)) for t in tokens)

    def encode(self, text):
        return list(self.prompt) if self.prompt is not None else []


WHISPER_GEN_CASES = {
    "default": dict(temperatures=(0.0, 0.4, 0.8, 1.0), logprob_threshold=-4.6, compression_ratio_threshold=2.4, no_speech_threshold=0.6),
    "nocond": dict(temperatures=(0.0, 0.5), logprob_threshold=-4.3, condition_on_previous_text=False, no_speech_threshold=None, initial_prompt_tokens=(11, 12, 13)),
    "nots": dict(temperatures=(0.0,), return_timestamps=False, clip_timestamps=(5.0, 40.0)),
    "pairs": dict(temperatures=(0.0, 0.6), logprob_threshold=-5.2, no_speech_threshold=None, clip_timestamps=(0.0, 9.0)),
}


def _segments_match(got, want, tol=1e-6, score_tol=2e-4):
    """Token ids, seek positions and texts identical; times and temperatures to ``tol``; the scores to ``score_tol``: the stand-in runs the
    reference in the checkpoint's float32 and the oracle runs in float64, and a window's summed log-probability over logits of +-40..60
    carries ~4e-5 of float32 rounding (measured)."""
    assert len(got) == len(want), (len(got), len(want))
    for a, b in zip(got, want):
        assert a["tokens"] == b["tokens"] and a["seek"] == b["seek"] and a["text"] == b["text"] and a["id"] == b["id"], (a, b)
        for k in ("start", "end", "temperature", "compression_ratio"):
            assert abs(a[k] - b[k]) <= tol * max(1.0, abs(b[k])), (k, a[k], b[k])
        for k in ("avg_logprob", "no_speech_prob"):
            assert (a[k] != a[k] and b[k] != b[k]) or abs(a[k] - b[k]) <= score_tol * max(1.0, abs(b[k])), (k, a[k], b[k])


def test_oracle_whisper_generate_matches_the_reference_generate():
    """whisper_golden.npz gen_* = the reference's Model.generate (whisper.py:799-1318) EXECUTED through the NumPy stand-in on 75 s of
    synthetic audio with a stub tokenizer: log-mel front end, three to a dozen 30-second windows, decode_with_fallback over several
    temperatures (categorical draws injected per decode call), the no-speech skip, segment cutting at timestamp pairs, the seek rule, prompt
    conditioning and its reset after a hot window, clip_timestamps, return_timestamps=False.  The oracle's transcribe() must return the same
    segments: token ids, seek positions and temperatures identical, times and scores to 1e-6 (the reference builds its window and filterbank in float32)."""
    import json
    from oracle import dsp as OD
    from oracle import whisper as OW
    g, _ = _golden("whisper_golden.npz")
    import synth_params
    P0 = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["gen_params"]).items()}
    audio, U = whisper_generate_audio()
    assert np.array_equal(audio[:16], g["gen_audio_head"]) and np.array_equal(U, g["gen_U"])
    mel = torch.as_tensor(OD.whisper_log_mel(audio, 80, padding=480000))
    spec = OW.TokenizerSpec(eot=200, sot=201, no_timestamps=208, timestamp_begin=209, no_speech=207, blank_ids=(7,), language=202, task=203,
                            transcribe=203, translate=204, sot_lm=205, sot_prev=206)
    for tag, kw in WHISPER_GEN_CASES.items():
        want = json.loads(str(g[f"gen_{tag}"]))
        kw = dict(kw)
        P = dict(P0)
        P["decoder.token_embedding.weight"] = P0["decoder.token_embedding.weight"].clone()
        P["decoder.token_embedding.weight"][:200] *= float(g["gen_pairs_gain"] if tag == "pairs" else g["gen_text_gain"])
        prompt = kw.pop("initial_prompt_tokens", ())
        text, segs = OW.transcribe(P, mel, spec, WHISPER_GEN_DIMS, WhisperStubTokenizer(), initial_prompt_tokens=prompt,
                                   suppress=g["gen_suppress"].tolist(), sample_len=14, uniforms=lambda k: U[k], **kw)
        _segments_match(segs, want["segments"])
        assert text == want["text"], tag
