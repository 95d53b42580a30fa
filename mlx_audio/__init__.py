"""Drop-in import surface: the reference's ``mlx_audio.*`` paths for the accelerated hot path, backed by
``mlx_audio_b200`` (torch tensors on CUDA, sm_100a kernels).  Only the paths in SURVEY.md section 8(b) exist."""
from mlx_audio_b200 import __version__  # noqa: F401
