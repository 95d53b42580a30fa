from mlx_audio_b200.codec.models.snac import *  # noqa: F401,F403
