from .models.mimi import Mimi, MimiConfig, MimiStreamingDecoder, mimi_202407
from .models.snac import SNAC

__all__ = ["Mimi", "MimiConfig", "MimiStreamingDecoder", "mimi_202407", "SNAC"]
