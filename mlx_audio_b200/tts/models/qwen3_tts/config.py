"""Qwen3-TTS configuration (reference: tts/models/qwen3_tts/config.py:19-251).

The reference declares eight dataclasses by hand; here each one is a field table fed to ``dataclasses.make_dataclass``, with the
same class names, field names, defaults and nesting rules, so ``ModelConfig.from_dict(config_json)`` yields an object that compares
equal, field by field, to the reference's (tests/test_host_cpu.py against tests/golden/config_golden.json, which the reference's own
dataclasses produced):

* unknown keys are dropped at every level (``filter_dict_for_dataclass``);
* a nested dict becomes the nested dataclass; a missing talker / code-predictor / speaker-encoder / tokenizer-decoder config becomes
  the default instance; a missing ``tokenizer_config`` and a missing tokenizer ``encoder_config`` stay ``None``;
* the code predictor's ``layer_types`` defaults to ``["full_attention"] * num_hidden_layers``.
"""
from __future__ import annotations

import copy
from dataclasses import field, fields, make_dataclass
from typing import Any

from ..base import BaseModelArgs


def filter_dict_for_dataclass(cls, data):
    valid = {f.name for f in fields(cls)}
    return {k: v for k, v in data.items() if k in valid}


def _table(name, spec, post_init=None, bases=()):
    cols = []
    for key, default in spec.items():
        if isinstance(default, (list, dict)):
            cols.append((key, Any, field(default_factory=lambda d=default: copy.deepcopy(d))))
        else:
            cols.append((key, Any, default))
    return make_dataclass(name, cols, bases=bases, namespace={"__post_init__": post_init} if post_init else {})


def _nest(obj, attr, cls, default_when_none=True):
    """``attr`` given as a dict -> ``cls`` built from its known keys; absent -> default instance (or left None)."""
    v = getattr(obj, attr)
    if isinstance(v, dict):
        setattr(obj, attr, cls(**filter_dict_for_dataclass(cls, v)))
    elif v is None and default_when_none:
        setattr(obj, attr, cls())


Qwen3TTSSpeakerEncoderConfig = _table("Qwen3TTSSpeakerEncoderConfig", {          # config.py:19-30 (ECAPA-TDNN)
    "mel_dim": 128, "enc_dim": 1024, "enc_channels": [512, 512, 512, 512, 1536], "enc_kernel_sizes": [5, 3, 3, 3, 1],
    "enc_dilations": [1, 2, 3, 4, 1], "enc_attention_channels": 128, "enc_res2net_scale": 8, "enc_se_channels": 128, "sample_rate": 24000})


def _cp_post(self):
    if self.layer_types is None:
        self.layer_types = ["full_attention"] * self.num_hidden_layers


Qwen3TTSTalkerCodePredictorConfig = _table("Qwen3TTSTalkerCodePredictorConfig", {  # config.py:33-55
    "vocab_size": 2048, "hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 5, "num_attention_heads": 16,
    "num_key_value_heads": 8, "head_dim": 128, "hidden_act": "silu", "max_position_embeddings": 65536, "rms_norm_eps": 1e-6,
    "rope_theta": 1000000.0, "rope_scaling": None, "attention_bias": False, "sliding_window": None, "layer_types": None,
    "attention_dropout": 0.0, "num_code_groups": 16}, _cp_post)

Qwen3TTSTalkerConfig = _table("Qwen3TTSTalkerConfig", {                          # config.py:58-102
    "code_predictor_config": None, "vocab_size": 3072, "hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 28,
    "num_attention_heads": 16, "num_key_value_heads": 8, "head_dim": 128, "hidden_act": "silu", "max_position_embeddings": 32768,
    "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
    "rope_scaling": {"interleaved": True, "mrope_section": [24, 20, 20], "rope_type": "default"},
    "attention_bias": False, "sliding_window": None, "attention_dropout": 0.0, "num_code_groups": 16, "text_hidden_size": 2048,
    "text_vocab_size": 151936, "codec_eos_token_id": 2150, "codec_think_id": 2154, "codec_nothink_id": 2155, "codec_think_bos_id": 2156,
    "codec_think_eos_id": 2157, "codec_pad_id": 2148, "codec_bos_id": 2149, "codec_language_id": None, "spk_id": None,
    "spk_is_dialect": None}, lambda self: _nest(self, "code_predictor_config", Qwen3TTSTalkerCodePredictorConfig))

Qwen3TTSTokenizerDecoderConfig = _table("Qwen3TTSTokenizerDecoderConfig", {      # config.py:105-133
    "attention_bias": False, "attention_dropout": 0.0, "latent_dim": 1024, "codebook_dim": 512, "codebook_size": 2048, "decoder_dim": 1536,
    "hidden_act": "silu", "hidden_size": 512, "intermediate_size": 1024, "layer_scale_initial_scale": 0.01, "max_position_embeddings": 8000,
    "head_dim": 64, "num_attention_heads": 16, "num_hidden_layers": 8, "num_key_value_heads": 16, "num_quantizers": 16,
    "num_semantic_quantizers": 1, "rms_norm_eps": 1e-5, "rope_theta": 10000.0, "semantic_codebook_size": 4096, "sliding_window": 72,
    "upsample_rates": [8, 5, 4, 3], "upsampling_ratios": [2, 2], "vector_quantization_hidden_dimension": 512})

Qwen3TTSTokenizerEncoderConfig = _table("Qwen3TTSTokenizerEncoderConfig", {      # config.py:136-173 (Mimi-style encoder, ICL voice cloning)
    "frame_rate": 12.5, "attention_bias": False, "attention_dropout": 0.0, "audio_channels": 1, "codebook_dim": 256, "codebook_size": 2048,
    "compress": 2, "dilation_growth_rate": 2, "head_dim": 64, "hidden_act": "gelu", "hidden_size": 512, "intermediate_size": 2048,
    "kernel_size": 7, "last_kernel_size": 3, "layer_scale_initial_scale": 0.01, "max_position_embeddings": 8000, "norm_eps": 1e-5,
    "num_attention_heads": 8, "num_filters": 64, "num_hidden_layers": 8, "num_key_value_heads": 8, "num_quantizers": 32,
    "num_residual_layers": 1, "num_semantic_quantizers": 1, "residual_kernel_size": 3, "rope_theta": 10000.0, "sampling_rate": 24000,
    "sliding_window": 250, "upsampling_ratios": [8, 6, 5, 4], "use_causal_conv": True, "use_conv_shortcut": False,
    "vector_quantization_hidden_dimension": 256})


def _tok_post(self):
    _nest(self, "encoder_config", Qwen3TTSTokenizerEncoderConfig, default_when_none=False)   # only needed for voice cloning
    _nest(self, "decoder_config", Qwen3TTSTokenizerDecoderConfig)


Qwen3TTSTokenizerConfig = _table("Qwen3TTSTokenizerConfig", {                    # config.py:176-201
    "encoder_config": None, "decoder_config": None, "encoder_valid_num_quantizers": 16, "input_sample_rate": 24000,
    "output_sample_rate": 24000, "decode_upsample_rate": 1920, "encode_downsample_rate": 1920}, _tok_post)


def _model_post(self):
    _nest(self, "talker_config", Qwen3TTSTalkerConfig)
    _nest(self, "speaker_encoder_config", Qwen3TTSSpeakerEncoderConfig)
    _nest(self, "tokenizer_config", Qwen3TTSTokenizerConfig, default_when_none=False)


ModelConfig = _table("ModelConfig", {                                            # config.py:204-251
    "model_type": "qwen3_tts", "talker_config": None, "speaker_encoder_config": None, "tokenizer_config": None,
    "tokenizer_type": "qwen3_tts_tokenizer_12hz", "tts_model_size": "0b6", "tts_model_type": "base", "im_start_token_id": 151644,
    "im_end_token_id": 151645, "tts_pad_token_id": 151671, "tts_bos_token_id": 151672, "tts_eos_token_id": 151673, "sample_rate": 24000},
    _model_post, bases=(BaseModelArgs,))
