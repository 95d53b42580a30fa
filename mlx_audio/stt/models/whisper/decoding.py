"""Drop-in location of the reference's ``stt/models/whisper/decoding.py`` names that the accelerated path implements."""
from mlx_audio_b200.stt.models.whisper.whisper import DecodingResult, TokenizerSpec, get_suppress_tokens  # noqa: F401
