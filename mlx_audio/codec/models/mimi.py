from mlx_audio_b200.codec.models.mimi import *  # noqa: F401,F403
