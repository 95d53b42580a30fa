"""Qwen3-TTS (reference: mlx_audio/tts/models/qwen3_tts/__init__.py)."""
from .config import (ModelConfig, Qwen3TTSTalkerCodePredictorConfig, Qwen3TTSTalkerConfig, Qwen3TTSTokenizerConfig,
                     Qwen3TTSTokenizerDecoderConfig)
from .qwen3_tts import Model
from .speech_tokenizer import Qwen3TTSSpeechTokenizer, Qwen3TTSSpeechTokenizerDecoder
from .talker import Qwen3TTSTalkerForConditionalGeneration

__all__ = ["Model", "ModelConfig", "Qwen3TTSTalkerConfig", "Qwen3TTSTalkerCodePredictorConfig", "Qwen3TTSTokenizerConfig",
           "Qwen3TTSTokenizerDecoderConfig", "Qwen3TTSSpeechTokenizer", "Qwen3TTSSpeechTokenizerDecoder",
           "Qwen3TTSTalkerForConditionalGeneration"]
