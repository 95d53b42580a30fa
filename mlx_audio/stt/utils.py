from mlx_audio_b200.stt.utils import load, load_model  # noqa: F401
