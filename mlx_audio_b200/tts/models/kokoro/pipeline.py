"""Kokoro text / phoneme front end: voices, the <= 510-phoneme chunking rule and the per-chunk call of the model (reference
tts/models/kokoro/pipeline.py:94-528, voice.py:4-15).

Everything here is host-side string handling around ``Model.__call__`` (which replays the cached CUDA graphs).  Grapheme-to-phoneme
conversion itself is the optional ``misaki`` package exactly as in the reference; it is not shipped in this image, so a pipeline can
also be given any callable ``g2p(text) -> (phonemes, tokens)`` -- tokens being objects with ``text`` / ``phonemes`` / ``whitespace``
attributes (misaki's ``MToken``) -- or be fed phoneme strings directly (``generate_from_tokens``).
"""
from __future__ import annotations

import logging
import re
from dataclasses import dataclass
from numbers import Number
from pathlib import Path
from typing import Any, Callable, Generator, List, Optional, Sequence, Tuple, Union

import torch

MAX_PHONEMES = 510                                   # context 512 minus BOS / EOS (kokoro.py:122-125)

ALIASES = {"en": "a", "en-us": "a", "en-gb": "b", "es": "e", "fr-fr": "f", "fr": "f", "hi": "h", "it": "i", "pt-br": "p", "pt": "p", "ja": "j",
           "zh": "z"}
LANG_CODES = dict(a="American English", b="British English", e="es", f="fr-fr", h="hi", i="it", p="pt-br", j="Japanese", z="Mandarin Chinese")

MISAKI_INSTALL_MESSAGE = "Kokoro requires the optional 'misaki' package for text processing. Install it with: pip install misaki"


def load_voice_tensor(path: Union[str, Path]) -> torch.Tensor:
    """voice.py:4-15: a voice pack ``[510, 1, 256]`` from ``.safetensors`` (key ``voice``) -- or a torch ``.pt`` tensor, the hub's original format."""
    path = str(path)
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)["voice"]
    obj = torch.load(path, map_location="cpu", weights_only=True)
    return obj["voice"] if isinstance(obj, dict) else obj


def phonemes_of(tokens: Sequence[Any]) -> str:
    """pipeline.py:231-235."""
    return "".join(t.phonemes + (" " if t.whitespace else "") for t in tokens).strip()


def text_of(tokens: Sequence[Any]) -> str:
    return "".join(t.text + t.whitespace for t in tokens).strip()


def split_point(tokens: Sequence[Any], next_count: int, waterfall=("!.?…", ":;", ",—"), bumps=(")", "”")) -> int:
    """Where to cut a token list that is about to exceed 510 phonemes (pipeline.py:237-260): after the LAST sentence-final mark if the
    remainder then fits, else after the last colon / semicolon, else after the last comma / dash, else everything."""
    for marks in waterfall:
        at = next((i for i in range(len(tokens) - 1, -1, -1) if tokens[i].phonemes in set(marks)), None)
        if at is None:
            continue
        at += 1
        if at < len(tokens) and tokens[at].phonemes in bumps:
            at += 1
        if next_count - len(phonemes_of(tokens[:at])) <= MAX_PHONEMES:
            return at
    return len(tokens)


def chunk_tokens(tokens: Sequence[Any]) -> Generator[Tuple[str, str, List[Any]], None, None]:
    """English chunking (pipeline.py:266-293): walk the tokens, and whenever the next one would push the phoneme count over 510 emit the
    prefix chosen by ``split_point``.  Yields (graphemes, phonemes, tokens)."""
    held: List[Any] = []
    count = 0
    for t in tokens:
        t.phonemes = "" if t.phonemes is None else t.phonemes.replace("ɾ", "T")       # American English flap
        nxt = t.phonemes + (" " if t.whitespace else "")
        if count + len(nxt.rstrip()) > MAX_PHONEMES:
            z = split_point(held, count + len(nxt.rstrip()))
            yield text_of(held[:z]), phonemes_of(held[:z]), held[:z]
            held = held[z:]
            count = len(phonemes_of(held))
            if not held:
                nxt = nxt.lstrip()
        held.append(t)
        count += len(nxt)
    if held:
        yield text_of(held), phonemes_of(held), held


def chunk_text(graphemes: str, chunk_size: int = 400) -> List[str]:
    """Non-English chunking (pipeline.py:470-499): ~400-character chunks on sentence boundaries, else fixed slices."""
    parts = re.split(r"([.!?]+)", graphemes)
    chunks, cur = [], ""
    for i in range(0, len(parts), 2):
        sentence = parts[i] + (parts[i + 1] if i + 1 < len(parts) else "")
        if len(cur) + len(sentence) <= chunk_size:
            cur += sentence
        else:
            if cur:
                chunks.append(cur.strip())
            cur = sentence
    if cur:
        chunks.append(cur.strip())
    return chunks or [graphemes[i:i + chunk_size] for i in range(0, len(graphemes), chunk_size)]


def join_timestamps(tokens: Sequence[Any], pred_dur) -> None:
    """pipeline.py:363-399: start / end seconds per token from the predicted durations (half-frame counting, 80 half-frames per second)."""
    dur = [int(v) for v in (pred_dur.tolist() if hasattr(pred_dur, "tolist") else pred_dur)]
    if not tokens or len(dur) < 3:
        return
    left = right = 2 * max(0, dur[0] - 3)
    i = 1
    for t in tokens:
        if i >= len(dur) - 1:
            break
        if not t.phonemes:
            if t.whitespace:
                i += 1
                left = right + dur[i]
                right = left + dur[i]
                i += 1
            continue
        j = i + len(t.phonemes)
        if j >= len(dur):
            break
        t.start_ts = left / 80
        token_dur = sum(dur[i:j])
        space_dur = dur[j] if t.whitespace else 0
        left = right + 2 * token_dur + space_dur
        t.end_ts = left / 80
        right = left + space_dur
        i = j + (1 if t.whitespace else 0)


class KokoroPipeline:
    """pipeline.py:94-528 over the B200 model.  ``g2p``: callable (default: misaki, if installed); ``voices_dir``: where
    ``<voice>.safetensors`` / ``.pt`` packs live (the reference downloads them from the hub into a ``voices/`` folder)."""

    @dataclass
    class Result:
        graphemes: str
        phonemes: str
        tokens: Optional[list] = None
        output: Optional[Any] = None
        text_index: Optional[int] = None

        @property
        def audio(self):
            return None if self.output is None else self.output.audio

        @property
        def pred_dur(self):
            return None if self.output is None else self.output.pred_dur

        def __iter__(self):                           # (graphemes, phonemes, audio) unpacking, as the reference allows
            yield self.graphemes
            yield self.phonemes
            yield self.audio

        def __getitem__(self, i):
            return [self.graphemes, self.phonemes, self.audio][i]

        def __len__(self):
            return 3

    def __init__(self, lang_code: str, model=None, repo_id: Optional[str] = None, g2p: Optional[Callable] = None,
                 voices_dir: Optional[Union[str, Path]] = None):
        lang_code = ALIASES.get(lang_code.lower(), lang_code.lower())
        if lang_code not in LANG_CODES:
            raise AssertionError((lang_code, LANG_CODES))
        self.lang_code, self.model, self.repo_id = lang_code, model, repo_id
        self.voices_dir = None if voices_dir is None else Path(voices_dir)
        self.voices = {}
        self._g2p = g2p

    # ---- grapheme-to-phoneme (optional dependency, as in the reference)
    @property
    def g2p(self) -> Callable:
        if self._g2p is None:
            try:
                if self.lang_code in "ab":
                    from misaki import en
                    self._g2p = en.G2P(trf=False, british=self.lang_code == "b", fallback=None, unk="")
                else:
                    from misaki import espeak
                    self._g2p = espeak.EspeakG2P(language=LANG_CODES[self.lang_code])
            except ImportError as exc:
                raise ImportError(MISAKI_INSTALL_MESSAGE + "; or construct KokoroPipeline(..., g2p=callable) / pass phonemes directly") from exc
        return self._g2p

    # ---- voices
    def load_single_voice(self, voice: str) -> torch.Tensor:
        if voice in self.voices:
            return self.voices[voice]
        if voice.endswith((".safetensors", ".pt")):
            f = Path(voice)
        else:
            if self.voices_dir is None:
                raise FileNotFoundError(f"voice '{voice}': no voices_dir configured and hub download is outside this build (pass a file path or a tensor)")
            f = next((self.voices_dir / f"{voice}{ext}" for ext in (".safetensors", ".pt") if (self.voices_dir / f"{voice}{ext}").exists()), None)
            if f is None:
                raise FileNotFoundError(f"voice '{voice}' not found under {self.voices_dir}")
            if not voice.startswith(self.lang_code):
                logging.warning(f"Language mismatch, loading {voice} voice into {LANG_CODES.get(self.lang_code, self.lang_code)} pipeline.")
        self.voices[voice] = load_voice_tensor(f)
        return self.voices[voice]

    def load_voice(self, voice, delimiter: str = ",") -> torch.Tensor:
        """One voice or a comma-separated list, which is averaged (pipeline.py:219-229); a tensor is taken as a ready voice pack."""
        if isinstance(voice, torch.Tensor):
            return voice
        voice = str(voice)
        if voice in self.voices:
            return self.voices[voice]
        packs = [self.load_single_voice(v) for v in voice.split(delimiter)]
        if len(packs) > 1:
            self.voices[voice] = torch.stack([p.float() for p in packs]).mean(dim=0)
            return self.voices[voice]
        return packs[0]

    # ---- synthesis
    @staticmethod
    def infer(model, ps: str, pack: torch.Tensor, speed: Number = 1):
        return model(ps, pack[len(ps) - 1], speed, return_output=True)      # the style row is picked by the phoneme count (pipeline.py:296-303)

    def generate_from_tokens(self, tokens, voice, speed: Number = 1, model=None):
        """pipeline.py:305-361: a raw phoneme string (<= 510) or a list of G2P tokens, which is chunked."""
        model = model or self.model
        if model and voice is None:
            raise ValueError('Specify a voice: pipeline.generate_from_tokens(..., voice="af_heart")')
        pack = self.load_voice(voice) if model else None
        if isinstance(tokens, str):
            if len(tokens) > MAX_PHONEMES:
                raise ValueError(f"Phoneme string too long: {len(tokens)} > {MAX_PHONEMES}")
            yield self.Result(graphemes="", phonemes=tokens, output=self.infer(model, tokens, pack, speed) if model else None)
            return
        for gs, ps, tks in chunk_tokens(tokens):
            if not ps:
                continue
            if len(ps) > MAX_PHONEMES:
                logging.warning(f"Unexpected len(ps) == {len(ps)} > {MAX_PHONEMES}; truncating")
                ps = ps[:MAX_PHONEMES]
            out = self.infer(model, ps, pack, speed) if model else None
            if out is not None and out.pred_dur is not None:
                join_timestamps(tks, out.pred_dur)
            yield self.Result(graphemes=gs, phonemes=ps, tokens=tks, output=out)

    def __call__(self, text, voice=None, speed: Number = 1, split_pattern: Optional[str] = r"\n+"):
        """pipeline.py:428-528."""
        if voice is None:
            raise ValueError('Specify a voice: en_us_pipeline(text="Hello world!", voice="af_heart")')
        pack = self.load_voice(voice) if self.model else None
        if isinstance(text, str):
            text = re.split(split_pattern, text.strip()) if split_pattern else [text]
        for idx, graphemes in enumerate(text):
            if not graphemes.strip():
                continue
            if self.lang_code in "ab":
                _, tokens = self.g2p(graphemes)
                for gs, ps, tks in chunk_tokens(tokens):
                    if not ps:
                        continue
                    ps = ps[:MAX_PHONEMES]
                    out = self.infer(self.model, ps, pack, speed) if self.model else None
                    if out is not None and out.pred_dur is not None:
                        join_timestamps(tks, out.pred_dur)
                    yield self.Result(graphemes=gs, phonemes=ps, tokens=tks, output=out, text_index=idx)
            else:
                for chunk in chunk_text(graphemes):
                    if not chunk.strip():
                        continue
                    res = self.g2p(chunk)
                    ps = res[0] if isinstance(res, tuple) else res
                    if not ps:
                        continue
                    ps = ps[:MAX_PHONEMES]
                    yield self.Result(graphemes=chunk, phonemes=ps, output=self.infer(self.model, ps, pack, speed) if self.model else None, text_index=idx)
