from .models.mimi import Mimi, MimiConfig, mimi_202407
from .models.snac import SNAC

__all__ = ["Mimi", "MimiConfig", "mimi_202407", "SNAC"]
