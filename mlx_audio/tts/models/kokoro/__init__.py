from mlx_audio_b200.tts.models.kokoro import Model, ModelConfig  # noqa: F401
