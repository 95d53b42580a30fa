"""config.json dictionaries (hub style) fed to both the reference's and the product's ``from_dict`` (make_config_golden.py, tests/test_host_cpu.py)."""
WHISPER_CONFIGS = {
    "mlx_small": {"model_type": "whisper", "n_mels": 80, "n_audio_ctx": 1500, "n_audio_state": 768, "n_audio_head": 12, "n_audio_layer": 12,
                  "n_vocab": 51865, "n_text_ctx": 448, "n_text_state": 768, "n_text_head": 12, "n_text_layer": 12, "quantization": {"bits": 4},
                  "unknown_key": 1},
    "hf_small": {"model_type": "whisper", "architectures": ["WhisperForConditionalGeneration"], "d_model": 768, "encoder_layers": 12, "decoder_layers": 12,
                 "encoder_attention_heads": 12, "decoder_attention_heads": 12, "num_mel_bins": 80, "max_source_positions": 1500,
                 "max_target_positions": 448, "vocab_size": 51865, "encoder_ffn_dim": 3072},
    "hf_defaults": {"d_model": 1280},
}
_TALKER = {"vocab_size": 3072, "hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 28, "num_attention_heads": 16,
           "num_key_value_heads": 8, "head_dim": 128, "rope_scaling": {"interleaved": True, "mrope_section": [24, 20, 20], "rope_type": "default"},
           "num_code_groups": 16, "text_hidden_size": 2048, "text_vocab_size": 151936, "codec_eos_token_id": 2150, "codec_language_id": {"english": 2050, "chinese": 2055},
           "spk_id": {"vivian": 3065, "ryan": 3061}, "spk_is_dialect": {"vivian": False}, "unknown_talker_key": 5,
           "code_predictor_config": {"vocab_size": 2048, "hidden_size": 1024, "num_hidden_layers": 5, "num_code_groups": 16, "unknown_cp_key": 1}}
QWEN3_CONFIGS = {
    "custom_voice": {"model_type": "qwen3_tts", "tts_model_type": "custom_voice", "tts_model_size": "0b6", "talker_config": _TALKER,
                     "speaker_encoder_config": {"enc_dim": 1024, "sample_rate": 24000, "bogus": 2}, "tts_pad_token_id": 151671, "im_start_token_id": 151644,
                     "tokenizer_config": {"decoder_config": {"num_hidden_layers": 8, "sliding_window": 72, "junk": 0}, "encoder_valid_num_quantizers": 16},
                     "unknown_top_level": "x"},
    "minimal": {"model_type": "qwen3_tts"},
}
KOKORO_CONFIG_JSON = {
    "istftnet": {"upsample_kernel_sizes": [20, 12], "upsample_rates": [10, 6], "gen_istft_hop_size": 5, "gen_istft_n_fft": 20,
                 "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]], "resblock_kernel_sizes": [3, 7, 11], "upsample_initial_channel": 512},
    "dim_in": 64, "dropout": 0.2, "hidden_dim": 512, "max_conv_dim": 512, "max_dur": 50, "multispeaker": True, "n_layer": 3, "n_mels": 80, "n_token": 178,
    "style_dim": 128, "text_encoder_kernel_size": 5, "plbert": {"hidden_size": 768, "num_attention_heads": 12, "intermediate_size": 2048,
                                                               "max_position_embeddings": 512, "num_hidden_layers": 12, "dropout": 0.1},
    "vocab": {"a": 43, "b": 44}, "model_type": "kokoro", "extra": 1,
}
