from mlx_audio_b200.tts.models.base import *  # noqa: F401,F403
