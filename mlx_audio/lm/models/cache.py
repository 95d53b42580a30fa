from mlx_audio_b200.lm.models.cache import KVCache  # noqa: F401
