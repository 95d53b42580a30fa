"""Long-form transcription: ``Model.generate`` (reference stt/models/whisper/whisper.py:799-1318).

The audio is turned into one log-mel spectrogram on the device, then consumed in 30-second windows: each window is decoded (the
fused decode-step kernel, temperature fallback when a result looks degenerate), its tokens are cut into segments at consecutive
timestamp tokens, and ``seek`` advances to the last timestamp (or by a whole window) -- the rule of the reference's loop.  Everything
here is host logic around three device calls per window (encoder, decoder prefill, one decode step per token); word-level timestamps
(cross-attention DTW) are outside the accelerated path and raise NotImplementedError.
"""
from __future__ import annotations

from dataclasses import replace
from typing import List, Optional, Sequence, Union

import torch

from .audio import FRAMES_PER_SECOND, HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram, pad_or_trim

# language codes in token order: token id = sot + 1 + index (stt/models/whisper/tokenizer.py:3-104)
LANGUAGE_CODES = (
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue",
)


class IdTokenizer:
    """Stand-in when the checkpoint's tokenizer files are absent: token ids are rendered as decimal words, so that text-dependent rules
    (compression ratio, empty-segment clearing) stay well defined and the ids survive in ``segments[*]['tokens']``."""

    def decode(self, tokens) -> str:
        return " ".join(str(int(t)) for t in tokens)

    def encode(self, text: str):
        return [int(w) for w in text.split() if w.lstrip("-").isdigit()]


def spec_for(spec, language: Optional[str], task: str, num_languages: int):
    """Tokenizer constants for a language / task pair (HFTokenizerWrapper.sot_sequence, whisper.py:100-131)."""
    lang_tok = spec.language
    if language is not None:
        if language not in LANGUAGE_CODES[:num_languages]:
            raise ValueError(f"Unsupported language: {language}")
        lang_tok = spec.sot + 1 + LANGUAGE_CODES.index(language)
    return replace(spec, language=lang_tok, task=spec.translate if task == "translate" else spec.transcribe)


def _needs_fallback(r, compression_ratio_threshold, logprob_threshold, no_speech_threshold) -> bool:
    """whisper.py:976-992."""
    bad = False
    if compression_ratio_threshold is not None and r.compression_ratio > compression_ratio_threshold:
        bad = True                                        # too repetitive
    if logprob_threshold is not None and r.avg_logprob < logprob_threshold:
        bad = True                                        # average log probability is too low
    if no_speech_threshold is not None and r.no_speech_prob > no_speech_threshold:
        bad = False                                       # silence
    return bad


@torch.no_grad()
def transcribe(model, audio, *, verbose: Optional[bool] = None, language: Optional[str] = None, task: str = "transcribe",
               temperature: Union[float, Sequence[float]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0), compression_ratio_threshold: Optional[float] = 2.4,
               logprob_threshold: Optional[float] = -1.0, no_speech_threshold: Optional[float] = 0.6, condition_on_previous_text: bool = True,
               initial_prompt=None, return_timestamps: bool = True, word_timestamps: bool = False, clip_timestamps: Union[str, List[float]] = "0",
               hallucination_silence_threshold: Optional[float] = None, stream: bool = False, spec=None, tokenizer=None,
               sample_len: Optional[int] = None, max_initial_timestamp: Optional[float] = 1.0, uniforms=None, **_ignored):
    """-> STTOutput(text, segments, language).  ``audio``: float32 samples at 16 kHz (array / tensor).  ``spec``: TokenizerSpec (ids of the
    special tokens, suppress list); ``tokenizer``: ``decode`` / ``encode`` (IdTokenizer when None).  ``initial_prompt``: text (needs a real
    tokenizer) or a list of token ids.  ``uniforms``: callable ``(n_steps, batch) -> tensor`` supplying the sampler's uniforms at
    temperature > 0 (tests); default: drawn on the device."""
    from .whisper import STTOutput, TokenizerSpec
    if word_timestamps or hallucination_silence_threshold is not None:
        raise NotImplementedError("word-level timestamps (cross-attention DTW, whisper/timing.py) are outside the accelerated path")
    if stream:
        raise NotImplementedError("generate(stream=True) (AlignAtt streaming) is not implemented; use the windowed path")
    if isinstance(audio, str):
        raise NotImplementedError("pass samples (float32, 16 kHz); file decoding is audio_io's job, outside the hot path")
    dev = model.device
    tokenizer = tokenizer or IdTokenizer()
    spec = spec or TokenizerSpec()
    audio = torch.as_tensor(audio, dtype=torch.float32)
    mel = log_mel_spectrogram(audio, model.dims.n_mels, padding=N_SAMPLES, device=dev)          # [frames, n_mels]
    content_frames = mel.shape[-2] - N_FRAMES
    if language is None:
        if not model.is_multilingual:
            language = "en"
        else:
            _, probs = model.detect_language(pad_or_trim(mel, N_FRAMES, axis=-2), spec)
            language = max(probs, key=probs.get)
    spec = spec_for(spec, language, task, model.num_languages)

    if isinstance(clip_timestamps, str):
        clip_timestamps = [float(ts) for ts in (clip_timestamps.split(",") if clip_timestamps else [])]
    points = [round(ts * FRAMES_PER_SECOND) for ts in clip_timestamps] or [0]
    if len(points) % 2 == 1:
        points.append(content_frames)
    else:
        points[-1] = min(content_frames, points[-1])
    clips = list(zip(points[::2], points[1::2]))
    temperatures = [temperature] if isinstance(temperature, (int, float)) else list(temperature)
    input_stride = N_FRAMES // model.dims.n_audio_ctx                       # mel frames per output token: 2
    time_precision = input_stride * HOP_LENGTH / SAMPLE_RATE                # seconds per timestamp token: 0.02
    tb, eot = spec.timestamp_begin, spec.eot

    all_tokens: List[int] = []
    if initial_prompt is not None:
        prompt0 = tokenizer.encode(" " + initial_prompt.strip()) if isinstance(initial_prompt, str) else [int(t) for t in initial_prompt]
        all_tokens.extend(prompt0)
    else:
        prompt0 = []
    all_segments: List[dict] = []
    prompt_reset_since = 0

    def decode_window(segment):
        """decode_with_fallback (whisper.py:957-995): the encoder runs once per window, the temperatures share its features."""
        feats = model.encoder(segment[None])
        result = None
        for t in temperatures:
            u = None if (t == 0 or uniforms is None) else uniforms(sample_len or model.dims.n_text_ctx // 2, 1)
            result = model.decode(feats[0], spec, sample_len, without_timestamps=not return_timestamps, max_initial_timestamp=max_initial_timestamp,
                                  tokenizer=tokenizer, language=language, temperature=t, uniforms=u, prompt=all_tokens[prompt_reset_since:])
            if not _needs_fallback(result, compression_ratio_threshold, logprob_threshold, no_speech_threshold):
                break
        return result

    seek = clips[0][0]
    for _, clip_end in clips:
        while seek < clip_end:
            time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
            segment_size = min(N_FRAMES, content_frames - seek, clip_end - seek)
            segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE
            result = decode_window(pad_or_trim(mel[seek:seek + segment_size], N_FRAMES, axis=-2))
            tokens = list(result.tokens)
            if no_speech_threshold is not None:
                skip = result.no_speech_prob > no_speech_threshold
                if logprob_threshold is not None and result.avg_logprob > logprob_threshold:
                    skip = False                                            # confident text despite the no-speech probability
                if skip:
                    seek += segment_size
                    continue
            current: List[dict] = []

            def segment(start, end, toks):
                return {"seek": seek, "start": float(start), "end": float(end), "text": tokenizer.decode([t for t in toks if t < eot]),
                        "tokens": list(toks), "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                        "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob}

            is_ts = [t >= tb for t in tokens]
            single_ending = is_ts[-2:] == [False, True]
            cuts = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]       # positions after a closed timestamp pair
            if cuts:
                if single_ending:
                    cuts.append(len(tokens))
                last = 0
                for cut in cuts:
                    piece = tokens[last:cut]
                    current.append(segment(time_offset + (piece[0] - tb) * time_precision, time_offset + (piece[-1] - tb) * time_precision, piece))
                    last = cut
                if single_ending:
                    seek += segment_size                                    # nothing after the last timestamp
                else:
                    seek += (tokens[last - 1] - tb) * input_stride          # resume at the last closed timestamp
            else:
                duration = segment_duration
                stamps = [t for t in tokens if t >= tb]
                if stamps and stamps[-1] != tb:
                    duration = (stamps[-1] - tb) * time_precision
                current.append(segment(time_offset, time_offset + duration, tokens))
                seek += segment_size
            if verbose:
                for sg in current:
                    print(f"[{sg['start']:.2f} --> {sg['end']:.2f}] {sg['text']}")
            for sg in current:                                              # instantaneous or text-less segments are emptied
                if sg["start"] == sg["end"] or sg["text"].strip() == "":
                    sg["text"], sg["tokens"], sg["words"] = "", [], []
            all_segments.extend({"id": i, **sg} for i, sg in enumerate(current, start=len(all_segments)))
            all_tokens.extend(t for sg in current for t in sg["tokens"])
            if not condition_on_previous_text or result.temperature > 0.5:
                prompt_reset_since = len(all_tokens)                        # no prompt after a high-temperature window
    return STTOutput(text=tokenizer.decode(all_tokens[len(prompt0):]), segments=all_segments, language=language)
