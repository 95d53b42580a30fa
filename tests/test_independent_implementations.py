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
