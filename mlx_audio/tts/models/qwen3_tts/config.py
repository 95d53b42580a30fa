from mlx_audio_b200.tts.models.qwen3_tts.config import *  # noqa: F401,F403
