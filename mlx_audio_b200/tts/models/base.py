"""Return contracts shared by the TTS models (reference: tts/models/base.py:8-99)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass

import torch


@dataclass
class BaseModelArgs:
    @classmethod
    def from_dict(cls, params):
        """Keep only the keys the dataclass declares (reference tts/models/base.py:10-18)."""
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})


def check_array_shape(arr) -> bool:
    """Layout heuristic of the reference's ``sanitize`` (tts/models/base.py:21-34): True when a 3-D
    conv weight already is (out, K, K')-shaped with out the largest and the last two equal."""
    shape = arr.shape
    if len(shape) != 3:
        return False
    out_channels, kh, kw = shape
    return (out_channels >= kh) and (out_channels >= kw) and (kh == kw)


@dataclass
class GenerationResult:
    audio: torch.Tensor
    samples: int
    sample_rate: int
    segment_idx: int
    token_count: int
    audio_duration: str
    real_time_factor: float
    prompt: dict
    audio_samples: dict
    processing_time_seconds: float
    peak_memory_usage: float
    is_streaming_chunk: bool = False
    is_final_chunk: bool = False


@dataclass
class BatchGenerationResult:
    audio: torch.Tensor
    sequence_idx: int
    samples: int
    sample_rate: int
    token_count: int
    audio_duration: str
    processing_time_seconds: float
    peak_memory_usage: float
    is_streaming_chunk: bool = False
    is_final_chunk: bool = False
