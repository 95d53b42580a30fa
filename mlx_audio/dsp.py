from mlx_audio_b200.dsp import *  # noqa: F401,F403
from mlx_audio_b200.dsp import __all__  # noqa: F401
