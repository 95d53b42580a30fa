from mlx_audio_b200.resample import _polyphase_filter, resample_audio_array, resample_audio_chunks  # noqa: F401
