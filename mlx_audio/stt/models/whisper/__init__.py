from mlx_audio_b200.stt.models.whisper import Model, ModelConfig, ModelDimensions  # noqa: F401
