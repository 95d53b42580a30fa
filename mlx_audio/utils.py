from mlx_audio_b200.utils import *  # noqa: F401,F403
from mlx_audio_b200.utils import base_load_model, get_model_class, load_config, load_model, resample_audio  # noqa: F401
