from mlx_audio_b200.tts.models.qwen3_tts.qwen3_tts import *  # noqa: F401,F403
