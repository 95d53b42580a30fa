"""Cross-checks against implementations that are independent of BOTH the reference's MLX code and this repository's reading of MLX:
PyTorch models shipped in the image's ``transformers`` package, run on the CPU in float64 with the synthetic hub-layout checkpoints of
tests/golden/checkpoint_layouts.py.  They validate the primitive semantics the oracle assumes (conv layouts, attention scaling, norms,
activations) and the checkpoint key mappings end to end."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
if HERE not in sys.path:
    sys.path.insert(0, HERE)
transformers = pytest.importorskip("transformers")


def test_whisper_oracle_and_sanitize_agree_with_the_transformers_implementation():
    """A HuggingFace-layout Whisper state dict -> (a) transformers' WhisperForConditionalGeneration, (b) the product's ``sanitize`` (pinned to the
    reference's in test_host_cpu.py) followed by the oracle.  Encoder output and decoder logits agree to 1e-12: the oracle's Whisper is
    Whisper, and the HF -> reference key / layout mapping loads the right tensors into the right places."""
    import checkpoint_layouts as L
    from oracle import whisper as OW
    from mlx_audio_b200.stt.models.whisper.whisper import Model as W
    d = L.WHISPER_DIMS
    cfg = transformers.WhisperConfig(
        vocab_size=d["n_vocab"], num_mel_bins=d["n_mels"], d_model=d["n_audio_state"], encoder_layers=d["n_audio_layer"], decoder_layers=d["n_text_layer"],
        encoder_attention_heads=d["n_audio_head"], decoder_attention_heads=d["n_text_head"], encoder_ffn_dim=4 * d["n_audio_state"],
        decoder_ffn_dim=4 * d["n_text_state"], max_source_positions=d["n_audio_ctx"], max_target_positions=d["n_text_ctx"], activation_function="gelu",
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1,
        suppress_tokens=None, begin_suppress_tokens=None)
    hf = transformers.WhisperForConditionalGeneration(cfg).double().eval()
    sd = {k: torch.as_tensor(v).double() for k, v in L.whisper_hf().items()}
    hf_sd = dict(sd)
    hf_sd["model.encoder.embed_positions.weight"] = OW.sinusoids(d["n_audio_ctx"], d["n_audio_state"]).double()   # the reference recomputes them
    hf_sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]                                             # tied output projection
    missing, unexpected = hf.load_state_dict(hf_sd, strict=False)
    assert not missing and not unexpected
    rng = np.random.default_rng(0)
    mel = torch.as_tensor(rng.standard_normal((2, 2 * d["n_audio_ctx"], d["n_mels"])))
    toks = torch.as_tensor(rng.integers(0, d["n_vocab"], size=(2, 6)))
    with torch.no_grad():
        enc = hf.model.encoder(mel.transpose(1, 2)).last_hidden_state
        logits = hf(input_features=mel.transpose(1, 2), decoder_input_ids=toks).logits
    P = W.sanitize(W.__new__(W), sd)
    xa = OW.encoder(P, mel, d)
    lg, _ = OW.decoder_forward(P, toks, xa, None, d)
    assert float((enc - xa).abs().max()) < 1e-12 and float((logits - lg).abs().max()) < 1e-12


def test_speech_tokenizer_encoder_oracle_agrees_with_the_transformers_mimi_encoder():
    """The Qwen3-TTS speech-tokenizer ENCODER is transformers' Mimi encoder re-hosted on the reference's Mimi modules (speech_tokenizer.py:957-1058,
    1253-1415).  A transformers MimiModel filled with the synthetic weights -> ``encode(audio)``; the same weights through the restated key mapping
    (equal to the reference's own ``sanitize``: sanitize_golden.json) -> the oracle's ``tokenizer_encode``.  The code streams must be IDENTICAL: SEANet
    encoder, half-split RoPE, replicate-padded stride-2 conv and the residual nearest-code search all mean what the oracle says they mean."""
    import json
    import zlib
    import checkpoint_layouts as L
    from oracle import qwen3 as Q
    W = L.qwen3_tokenizer_encoder_hf()
    mapped = L.map_hf_mimi_encoder(W)
    want = json.load(open(os.path.join(HERE, "sanitize_golden.json")))["qwen3_tokenizer_encoder"]
    assert {k: [list(v.shape), zlib.crc32(np.ascontiguousarray(v, dtype=np.float32).tobytes())] for k, v in mapped.items()} == want
    hf = transformers.MimiModel(transformers.MimiConfig(**L.HF_MIMI_SMALL)).double().eval()
    missing, unexpected = hf.load_state_dict({k[len("encoder."):]: torch.as_tensor(v).double() for k, v in W.items()}, strict=False)
    assert not unexpected and all(k.startswith(("decoder", "upsample")) for k in missing)
    P = {k: torch.as_tensor(np.asarray(v)).double() for k, v in mapped.items()}
    audio = torch.as_tensor(0.4 * np.random.default_rng(3).standard_normal((2, 1, 7 * 1920 + 500)))
    with torch.no_grad():
        codes = hf.encode(audio).audio_codes
    assert tuple(codes.shape) == (2, 6, 8) and torch.equal(codes, Q.tokenizer_encode(P, audio, L.ORACLE_MIMI_SMALL))


def test_reference_talker_outputs_agree_with_the_transformers_qwen3_decoder():
    """qwen3_golden.npz holds what the REFERENCE's talker and code predictor computed (make_qwen3_golden.py).  With the three MRoPE position axes
    equal -- the only case text-to-speech uses -- interleaved MRoPE is ordinary RoPE, and the talker is a Qwen3 decoder stack: transformers' Qwen3Model
    with the same weights (identical parameter names) must reproduce the reference's hidden states and logits.  5e-6: transformers builds its rotary
    table in float32."""
    import json
    import synth_params
    g = np.load(os.path.join(HERE, "qwen3_golden.npz"))
    cfg = json.loads(str(g["cfg"]))
    P = {k: torch.as_tensor(v) for k, v in synth_params.from_manifest(g["talker_params"]).items()}

    def stack(prefix, layers, hidden, inter, vocab):
        hc = transformers.Qwen3Config(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                                      num_attention_heads=cfg["num_attention_heads"], num_key_value_heads=cfg["num_key_value_heads"], head_dim=cfg["head_dim"],
                                      rms_norm_eps=cfg["rms_norm_eps"], rope_theta=cfg["rope_theta"], max_position_embeddings=256, attention_bias=False,
                                      tie_word_embeddings=False, use_sliding_window=False, attention_dropout=0.0)
        m = transformers.Qwen3Model(hc).double().eval()
        sd = {k: P[prefix + k] for k in m.state_dict() if prefix + k in P}
        sd["embed_tokens.weight"] = torch.zeros(vocab, hidden, dtype=torch.float64)                      # unused: inputs_embeds are given
        assert not m.load_state_dict(sd, strict=True).missing_keys
        return m
    talker = stack("model.", cfg["num_hidden_layers"], cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"])
    with torch.no_grad():
        h = talker(inputs_embeds=torch.as_tensor(g["t_x"])).last_hidden_state
    assert float((h - torch.as_tensor(g["t_hidden"])).abs().max()) < 5e-6
    assert float((h @ P["codec_head.weight"].T - torch.as_tensor(g["t_logits"])).abs().max()) < 5e-6
    cp = stack("code_predictor.model.", cfg["cp_num_hidden_layers"], cfg["cp_hidden_size"], cfg["cp_intermediate_size"], cfg["cp_vocab_size"])
    x = torch.as_tensor(g["cp_x"]) @ P["code_predictor.small_to_mtp_projection.weight"].T + P["code_predictor.small_to_mtp_projection.bias"]
    with torch.no_grad():
        hc_ = cp(inputs_embeds=x).last_hidden_state
    assert float((hc_[:, -1] @ P["code_predictor.lm_head.0.weight"].T - torch.as_tensor(g["cp_logits"][:, 0])).abs().max()) < 5e-6


def test_mimi_decode_oracle_agrees_with_the_transformers_mimi_decoder():
    """transformers' MimiModel.decode vs the oracle's ``mimi_decode`` on the same synthetic weights (keys mapped onto the reference's module tree,
    q / k rows re-interleaved because the reference rotates interleaved pairs where transformers rotates half-split pairs).  2e-7 with the tanh GELU
    the reference's MLP uses (``nn.gelu_approx``, transformer.py:141); with transformers' default exact GELU the two differ by ~1e-4 -- the size of the
    reference's own departure from the original model, noted here, not corrected."""
    import checkpoint_layouts as L
    import synth_params
    from oracle import codec as OC
    codes = torch.as_tensor(np.random.default_rng(1).integers(0, 64, size=(2, 6, 9)))
    err = {}
    for act in ("gelu_pytorch_tanh", "gelu"):
        hf = transformers.MimiModel(transformers.MimiConfig(**L.HF_MIMI_SMALL, hidden_act=act)).double().eval()
        sd = {k: torch.as_tensor(synth_params.value(k, tuple(v.shape))).double() for k, v in hf.state_dict().items()}
        hf.load_state_dict(sd)
        with torch.no_grad():
            a = hf.decode(codes).audio_values
        o = OC.mimi_decode(L.map_hf_mimi_decoder(sd, heads=4, head_dim=8), codes, dict(L.ORACLE_MIMI_SMALL))
        assert a.shape == o.shape == (2, 1, 9 * 1920)
        err[act] = float((a - o).abs().max())
    assert err["gelu_pytorch_tanh"] < 2e-7 and 1e-6 < err["gelu"] < 1e-2, err


def test_stft_oracle_agrees_with_torch_stft_where_the_reference_does():
    """torch.stft vs the oracle's ``stft`` (symmetric Hann, centred, reflect / constant padding): 1e-13 whenever the window fills the frame.
    With win_length < n_fft they differ BY DESIGN: the reference zero-pads the window on the right (dsp.py:399-403, pinned by dsp_golden.npz
    "shortwin"), torch and librosa centre it."""
    from oracle import dsp as O
    x = np.random.default_rng(2).standard_normal(4000)
    for n_fft, hop in ((400, 160), (20, 5), (256, 64), (1024, 256)):
        for pad_mode in ("reflect", "constant"):
            w = torch.as_tensor(O.hanning(n_fft, periodic=False))
            t = torch.stft(torch.as_tensor(x), n_fft, hop, n_fft, window=w, center=True, pad_mode=pad_mode, return_complex=True).T.numpy()
            o = O.stft(x, n_fft=n_fft, hop_length=hop, window="hann", center=True, pad_mode=pad_mode)
            assert t.shape == o.shape and np.abs(t - o).max() < 1e-12
    w = torch.as_tensor(O.hanning(400, periodic=False))
    t = torch.stft(torch.as_tensor(x), 512, 128, 400, window=w, center=True, pad_mode="reflect", return_complex=True).T.numpy()
    assert np.abs(t - O.stft(x, n_fft=512, hop_length=128, win_length=400)).max() > 1.0


def test_kokoro_albert_and_lstm_oracles_agree_with_transformers_and_torch():
    """Kokoro's text side against two independent implementations on the synthetic 82M weights: transformers' AlbertModel (same parameter names as
    the reference's CustomAlbert; exact GELU as the reference uses -- the original PL-BERT default "gelu_new" would differ by ~1e-3) and
    torch.nn.LSTM (the reference's hand-written bidirectional LSTM, modules.py:93-285, follows PyTorch's gate order and double bias)."""
    from mlx_audio_b200 import synth
    from oracle import kokoro as OK
    cfg, pb = OK.KOKORO_CONFIG, OK.KOKORO_CONFIG["plbert"]
    P = {k: v.double() for k, v in synth.kokoro_weights(cfg, seed=0).items()}
    ac = transformers.AlbertConfig(vocab_size=cfg["n_token"], embedding_size=128, hidden_size=pb["hidden_size"], num_hidden_layers=pb["num_hidden_layers"],
                                   num_attention_heads=pb["num_attention_heads"], intermediate_size=pb["intermediate_size"],
                                   max_position_embeddings=pb["max_position_embeddings"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                   hidden_act="gelu", num_hidden_groups=1, inner_group_num=1, type_vocab_size=2)
    hf = transformers.AlbertModel(ac).double().eval()
    assert not hf.load_state_dict({k: P["bert." + k] for k in hf.state_dict()}, strict=True).missing_keys
    ids = torch.as_tensor(np.random.default_rng(0).integers(1, cfg["n_token"], size=(1, 17)))
    with torch.no_grad():
        h = hf(input_ids=ids, attention_mask=torch.ones_like(ids)).last_hidden_state
    assert float((h - OK.albert(P, ids, torch.ones_like(ids), pb)).abs().max()) < 1e-12
    pre = "predictor.lstm"
    hid, inp = P[pre + ".Wh_forward"].shape[1], P[pre + ".Wx_forward"].shape[1]
    lstm = torch.nn.LSTM(inp, hid, batch_first=True, bidirectional=True).double()
    with torch.no_grad():
        for d, suffix in (("forward", ""), ("backward", "_reverse")):
            getattr(lstm, "weight_ih_l0" + suffix).copy_(P[f"{pre}.Wx_{d}"])
            getattr(lstm, "weight_hh_l0" + suffix).copy_(P[f"{pre}.Wh_{d}"])
            getattr(lstm, "bias_ih_l0" + suffix).copy_(P[f"{pre}.bias_ih_{d}"])
            getattr(lstm, "bias_hh_l0" + suffix).copy_(P[f"{pre}.bias_hh_{d}"])
        x = torch.as_tensor(np.random.default_rng(1).standard_normal((2, 23, inp)))
        want, _ = lstm(x)
    assert float((want - OK.lstm_bi(P, pre, x)).abs().max()) < 1e-12


def test_interpolate_and_istft_oracles_agree_with_torch():
    """torch.nn.functional.interpolate (nearest / linear, align_corners on and off, up- and down-scaling incl. Kokoro's x300 and /300) and torch.istft
    (window-squared normalisation = the reference's ``normalized=True``) against the oracle's restatements of tts/models/interpolate.py and
    dsp.py:436-513."""
    import torch.nn.functional as F
    from oracle import dsp as O
    rng = np.random.default_rng(0)
    for length, kw in ((50, dict(scale_factor=2.0)), (300, dict(scale_factor=1 / 3)), (7, dict(size=20)), (40, dict(size=13)), (10, dict(scale_factor=300.0)),
                       (3000, dict(scale_factor=1 / 300))):
        x = rng.standard_normal((2, 3, length))
        for mode, ac in (("nearest", None), ("linear", False), ("linear", True)):
            t = F.interpolate(torch.as_tensor(x), mode=mode, align_corners=ac, **kw).numpy()
            o = np.asarray(O.interpolate(x, mode=mode, align_corners=ac, **kw))
            assert t.shape == o.shape and np.abs(t - o).max() < 1e-12, (length, kw, mode, ac)
    y = rng.standard_normal(2048)
    for n_fft, hop in ((256, 64), (64, 16)):
        w = torch.as_tensor(O.hanning(n_fft, periodic=True))
        spec = torch.stft(torch.as_tensor(y), n_fft, hop, n_fft, window=w, center=True, pad_mode="reflect", return_complex=True)
        t = torch.istft(spec, n_fft, hop, n_fft, window=w, center=True).numpy()
        o = O.istft(spec.numpy(), hop_length=hop, win_length=n_fft, window=w.numpy(), center=True, normalized=True)
        n = min(t.shape[0], o.shape[0])
        assert abs(t.shape[0] - o.shape[0]) <= hop and np.abs(t[:n] - o[:n]).max() < 1e-10
