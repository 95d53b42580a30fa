from .whisper import Model, ModelConfig, ModelDimensions, TokenizerSpec, get_suppress_tokens

__all__ = ["Model", "ModelConfig", "ModelDimensions", "TokenizerSpec", "get_suppress_tokens"]
