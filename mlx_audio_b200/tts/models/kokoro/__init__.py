from .kokoro import Model, ModelConfig

__all__ = ["Model", "ModelConfig"]
