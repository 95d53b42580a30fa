"""stt/utils.py:133-161 of the reference: category wrappers over the generic loader."""
from ..utils import base_load_model


def load_model(model_path, lazy: bool = False, strict: bool = False, **kwargs):
    return base_load_model(model_path, "stt", lazy, strict, **kwargs)


load = load_model
