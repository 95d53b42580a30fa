from mlx_audio_b200.tts.models.qwen3_tts.speech_tokenizer import *  # noqa: F401,F403
