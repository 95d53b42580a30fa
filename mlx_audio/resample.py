from mlx_audio_b200.resample import _polyphase_filter, resample_audio_array  # noqa: F401
