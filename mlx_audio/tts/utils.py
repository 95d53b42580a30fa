from mlx_audio_b200.tts.utils import load, load_model  # noqa: F401
