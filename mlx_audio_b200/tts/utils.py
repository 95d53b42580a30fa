"""tts/utils.py:100-132 of the reference: category wrappers over the generic loader."""
from ..utils import base_load_model


def load_model(model_path, lazy: bool = False, strict: bool = True, **kwargs):
    return base_load_model(model_path, "tts", lazy, strict, **kwargs)


load = load_model
