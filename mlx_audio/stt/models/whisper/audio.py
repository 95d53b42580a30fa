from mlx_audio_b200.stt.models.whisper.audio import *  # noqa: F401,F403
from mlx_audio_b200.stt.models.whisper.audio import log_mel_spectrogram, pad_or_trim  # noqa: F401
