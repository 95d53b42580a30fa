"""Qwen3-TTS generation loop on B200 (reference: tts/models/qwen3_tts/qwen3_tts.py).

Covers the base path of ``Model.generate`` (qwen3_tts.py:1122-1575): input assembly from token ids
(``_prepare_generation_inputs`` :326-484 minus the tokenizer / speaker encoder, which are host / "next" rows), the per-frame
loop (:1323-1404) with ``_sample_token`` (:805-860), and ``_decode_chunk`` (:1017-1048) through the speech tokenizer.

One frame = talker step + first-codebook sample + 15 code-predictor sub-steps (each with its sampler) + next-input
embedding sum: ~700 small launches.  They are captured ONCE into a CUDA graph; every scalar that changes between frames
(KV length, trailing-text index, uniforms, seen-token set, codes) lives in device memory, so the host only replays the
graph and reads back 16 integers per frame for the EOS test (the reference also syncs once per frame, :1398-1400).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch

from .... import ops
from ..base import GenerationResult
from .config import ModelConfig
from .speech_tokenizer import Qwen3TTSSpeechTokenizer
from .talker import Qwen3TTSTalkerForConditionalGeneration


def format_duration(seconds: float) -> str:
    """qwen3_tts.py:160-165."""
    hours = int(seconds // 3600)
    minutes = int((seconds % 3600) // 60)
    secs = int(seconds % 60)
    ms = int((seconds % 1) * 1000)
    return f"{hours:02d}:{minutes:02d}:{secs:02d}.{ms:03d}"


def mel_spectrogram(audio, n_fft: int = 1024, num_mels: int = 128, sample_rate: int = 24000, hop_size: int = 256,
                    win_size: int = 1024, fmin: float = 0.0, fmax: float = 12000.0, device="cuda") -> torch.Tensor:
    """qwen3_tts.py:64-120 (speaker-encoder front end): manual reflect pad of (n_fft-hop)/2, STFT (center=False, Hann),
    sqrt(|X|^2 + 1e-9) @ slaney-mel^T, log(clip(., 1e-5)).  [n] or [B, n] -> [B, frames, num_mels]; the batch is one STFT launch."""
    from .... import dsp
    a = torch.as_tensor(audio, dtype=torch.float32, device=device)
    if a.dim() == 1:
        a = a[None]
    pad = (n_fft - hop_size) // 2
    a = torch.cat([a[:, 1:pad + 1].flip(1), a, a[:, -(pad + 1):-1].flip(1)], dim=1)
    spec = dsp.stft(a, n_fft=n_fft, hop_length=hop_size, win_length=win_size, window="hann", center=False, pad_mode="reflect", device=device)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)
    basis = dsp.mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=num_mels, f_min=fmin, f_max=fmax, norm="slaney", mel_scale="slaney")
    basis = torch.as_tensor(basis, dtype=torch.float32, device=a.device)
    cw = ops.pack_linear(basis, None, a.device)
    mel = ops.linear(mag.contiguous(), cw)
    return torch.log(torch.clamp(mel, min=1e-5))


class Model:
    def __init__(self, config: ModelConfig, device="cuda"):
        self.config = config
        self.device = torch.device(device)
        self.talker = Qwen3TTSTalkerForConditionalGeneration(config.talker_config, device)
        self.speaker_encoder = None          # ECAPA-TDNN voice cloning: SURVEY.md section 8f "next"
        self.speech_tokenizer: Optional[Qwen3TTSSpeechTokenizer] = None
        self.tokenizer = None
        self.generate_config = None
        tc = config.talker_config
        self.supported_speakers = list(tc.spk_id.keys()) if tc.spk_id else []
        self.supported_languages = ["auto"] + [l for l in (tc.codec_language_id or {}) if "dialect" not in l]
        self._graph = None

    def get_supported_speakers(self):
        return self.supported_speakers

    def get_supported_languages(self):
        return self.supported_languages

    @property
    def sample_rate(self) -> int:
        return self.config.sample_rate

    @property
    def model_type(self) -> str:
        return self.config.model_type

    def load_weights(self, weights, strict: bool = True):
        """``weights`` (dict or list of pairs) with the checkpoint's ``talker.`` prefix (stripped here, talker.py:825-839)."""
        weights = dict(weights)
        self.talker.load_weights(self.talker.sanitize(weights))
        t = self.talker
        self._tabs_all = ops.EmbedTables([t.codec_embedding] + t.code_predictor.codec_embedding)
        self._tab0 = ops.EmbedTables([t.codec_embedding])
        self._tab_cp = [ops.EmbedTables([e]) for e in t.code_predictor.codec_embedding]
        return self

    def load_speech_tokenizer(self, speech_tokenizer: Qwen3TTSSpeechTokenizer):
        self.speech_tokenizer = speech_tokenizer

    def load_generate_config(self, generate_config: dict):
        self.generate_config = generate_config

    def eval(self):
        return self

    @staticmethod
    def sanitize(weights):
        """qwen3_tts.py:2914-2935: drop ``position_ids``; conv weights [out, in, K] -> [out, K, in] unless already MLX-layout."""
        from .speech_tokenizer import check_array_shape_qwen3
        out = {}
        for k, v in weights.items():
            if "position_ids" in k:
                continue
            if ("conv" in k or "speaker_encoder.fc" in k) and "weight" in k and v.dim() == 3:
                v = v if check_array_shape_qwen3(v) else v.permute(0, 2, 1)
            out[k] = v
        return out

    @classmethod
    def post_load_hook(cls, model: "Model", model_path):
        """qwen3_tts.py:2818-2911: HF tokenizer (when its files are present), ``speech_tokenizer/`` sub-model, generation config."""
        import json
        from pathlib import Path
        from .config import Qwen3TTSTokenizerConfig, Qwen3TTSTokenizerDecoderConfig, filter_dict_for_dataclass
        model_path = Path(model_path)
        try:
            from transformers import AutoTokenizer
            model.tokenizer = AutoTokenizer.from_pretrained(str(model_path))
        except Exception as e:                                           # same behaviour as the reference: warn and continue
            print(f"Warning: Could not load tokenizer: {e}")
        st_path = model_path / "speech_tokenizer"
        if st_path.exists():
            from safetensors.torch import load_file
            d = json.load(open(st_path / "config.json"))
            dec = Qwen3TTSTokenizerDecoderConfig(**filter_dict_for_dataclass(Qwen3TTSTokenizerDecoderConfig, d["decoder_config"])) \
                if "decoder_config" in d else None
            tc = Qwen3TTSTokenizerConfig(decoder_config=dec)
            for k, v in d.items():
                if k not in ("decoder_config", "encoder_config") and hasattr(tc, k):
                    setattr(tc, k, v)
            w = {}
            for wf in sorted(st_path.glob("*.safetensors")):
                w.update(load_file(str(wf)))
            if w:
                st = Qwen3TTSSpeechTokenizer(tc, model.device).load_weights(Qwen3TTSSpeechTokenizer.sanitize(w))
                model.load_speech_tokenizer(st)
        gen = model_path / "generation_config.json"
        if gen.exists():
            model.load_generate_config(json.load(open(gen)))
        return model

    # ------------------------------------------------------------------ input assembly
    @torch.no_grad()
    def prepare_generation_inputs_from_ids(self, input_ids, language_id: Optional[int] = None, speaker_id=None, speaker_embed=None,
                                           instruct_ids=None):
        """qwen3_tts.py:326-484 after tokenisation: ``input_ids`` = tokenizer.encode("<|im_start|>assistant\\n{text}<|im_end|>\\n
        <|im_start|>assistant\\n").  Returns (input_embeds [1,P,H], trailing_text_hidden [1,n,H], tts_pad_embed [1,1,H])."""
        t, cfg, dev = self.talker, self.config.talker_config, self.device
        ids = torch.as_tensor(input_ids, dtype=torch.int64, device=dev).reshape(-1)
        text_embed = t.text_projection(ops.gather_rows(t.text_embedding, ids)[None])                          # [1,L,H]
        tts_ids = torch.tensor([self.config.tts_bos_token_id, self.config.tts_eos_token_id, self.config.tts_pad_token_id], device=dev)
        tts = t.text_projection(ops.gather_rows(t.text_embedding, tts_ids)[None])
        tts_bos, tts_eos, tts_pad = tts[:, 0:1], tts[:, 1:2], tts[:, 2:3]
        if speaker_embed is None and speaker_id is not None:
            speaker_embed = ops.gather_rows(t.codec_embedding, torch.tensor([int(speaker_id)], device=dev))[None]
        if language_id is None:
            prefill = [cfg.codec_nothink_id, cfg.codec_think_bos_id, cfg.codec_think_eos_id]
        else:
            prefill = [cfg.codec_think_id, cfg.codec_think_bos_id, int(language_id), cfg.codec_think_eos_id]
        codec = ops.gather_rows(t.codec_embedding, torch.tensor(prefill, device=dev))[None]
        suffix = ops.gather_rows(t.codec_embedding, torch.tensor([cfg.codec_pad_id, cfg.codec_bos_id], device=dev))[None]
        parts = [codec] + ([speaker_embed.reshape(1, 1, -1).float()] if speaker_embed is not None else []) + [suffix]
        codec = torch.cat(parts, dim=1)
        role = text_embed[:, :3]
        combined = torch.cat([tts_pad.expand(1, codec.shape[1] - 2, -1), tts_bos], dim=1) + codec[:, :-1]
        first_text = text_embed[:, 3:4] + codec[:, -1:]
        parts = [role, combined, first_text]
        if instruct_ids is not None:                                      # "<|im_start|>user\n{instruct}<|im_end|>\n" (:452-458,473-476)
            iid = torch.as_tensor(instruct_ids, dtype=torch.int64, device=dev).reshape(-1)
            parts = [t.text_projection(ops.gather_rows(t.text_embedding, iid)[None])] + parts
        input_embeds = torch.cat(parts, dim=1).contiguous()
        trailing = torch.cat([text_embed[:, 4:-5], tts_eos], dim=1).contiguous()
        return input_embeds, trailing, tts_pad.contiguous()

    def _prepare_generation_inputs(self, text: str, language: str = "auto", speaker: Optional[str] = None, instruct: Optional[str] = None):
        """qwen3_tts.py:326-484: tokenise with the chat template, resolve speaker / language / dialect ids from the config."""
        if self.tokenizer is None:
            raise ValueError("Tokenizer not loaded. Call post_load_hook first.")
        cfg = self.config.talker_config
        ids = self.tokenizer.encode(f"<|im_start|>assistant\n{text}<|im_end|>\n<|im_start|>assistant\n")
        speaker_id = None
        if speaker and speaker.lower() in (cfg.spk_id or {}):
            sid = cfg.spk_id[speaker.lower()]
            speaker_id = sid[0] if isinstance(sid, (list, tuple)) else sid
        language_id = None
        if language.lower() != "auto" and cfg.codec_language_id and language.lower() in cfg.codec_language_id:
            language_id = cfg.codec_language_id[language.lower()]
        if language.lower() in ("chinese", "auto") and speaker and speaker.lower() in (cfg.spk_is_dialect or {}) \
                and cfg.spk_is_dialect[speaker.lower()]:
            dialect = cfg.spk_is_dialect[speaker.lower()]
            if dialect in (cfg.codec_language_id or {}):
                language_id = cfg.codec_language_id[dialect]
        instruct_ids = self.tokenizer.encode(f"<|im_start|>user\n{instruct}<|im_end|>\n") if instruct else None
        return self.prepare_generation_inputs_from_ids(ids, language_id, speaker_id, instruct_ids=instruct_ids)

    def _suppress_codec_tokens(self, eos_token_id: int):
        """qwen3_tts.py:927-933."""
        cfg = self.config.talker_config
        return [i for i in range(cfg.vocab_size - 1024, cfg.vocab_size) if i != eos_token_id]

    # ------------------------------------------------------------------ frame loop
    def _frame(self, x_in: torch.Tensor, sp) -> None:
        """One pass of the loop body qwen3_tts.py:1323-1398 for every batch row; results land in the state buffers."""
        t, cp = self.talker, self.talker.code_predictor
        B = x_in.shape[0]
        g = self.config.talker_config.num_code_groups
        logits, hidden = t(x_in, use_device_offset=True, kv_start=self._kv_start)
        ops.sample_token(logits[:, -1], temperature=sp["temperature"], top_k=sp["top_k"], top_p=sp["top_p"], u=self._u[0],
                         suppress_mask=self._suppress, seen=self._seen, repetition_penalty=sp["repetition_penalty"], mark_seen=True,
                         out=self._codes[:, 0], finished=self._finished, eos=sp["eos"])
        inp0 = self._cp_in0                                                                  # [B,2,H]: (hidden, embed(token 0))
        ops.copy2d(hidden[:, -1], inp0[:, 0])
        ops.embed_sum(self._codes[:, 0:1], self._tab0, out=inp0[:, 1], err=self._err)
        for ci in range(g - 1):
            if ci == 0:
                lg = cp(inp0, 0, 0)
            else:
                e = ops.embed_sum(self._codes[:, ci:ci + 1], self._tab_cp[ci - 1], out=self._cp_in, err=self._err)
                lg = cp(e[:, None], ci + 1, ci)
            ops.sample_token(lg[:, -1], temperature=sp["temperature"], top_k=sp["top_k"], top_p=sp["top_p"], u=self._u[ci + 1],
                             out=self._codes[:, ci + 1])
        if self._tidx is not None:      # batch rule: per-row trailing index, clamp-pad, advance unfinished rows (qwen3_tts.py:1903-1912)
            ops.embed_sum(self._codes, self._tabs_all, text=self._trailing, pad=self._pad, out=self._x_in[:, 0], err=self._err,
                          tidx=self._tidx, finished=self._finished)
        else:
            ops.embed_sum(self._codes, self._tabs_all, text=self._trailing, pad=self._pad, step_dev=t.offset_dev, step_sub=self._prefill_len,
                          out=self._x_in[:, 0], err=self._err)

    def _uniform_stream(self, seed):
        """Generator the sampler's uniforms are drawn from.  ``seed=None`` continues ONE stream owned by the model, so successive
        segments and calls see fresh draws the way the reference's global ``mx.random`` state advances (qwen3_tts.py:805-860);
        an integer starts a reproducible stream for this call only."""
        if seed is not None:
            return torch.Generator(device=self.device).manual_seed(int(seed))
        if getattr(self, "_rng", None) is None:
            self._rng = torch.Generator(device=self.device)
            self._rng.seed()
        return self._rng

    @torch.no_grad()
    def generate_codes(self, input_embeds, trailing_text_hidden, tts_pad_embed, *, max_tokens: int = 4096, temperature: float = 0.9,
                       top_k: int = 50, top_p: float = 1.0, repetition_penalty: float = 1.05, u=None, seed: int = 0,
                       use_graph: bool = True, stop_on_eos: bool = True, left_padding=None, batch_mode: bool = False,
                       trailing_rule: str = "clamp_pad"):
        """The generation loop of Model.generate for B prompts of equal prefill length: returns int64 codes [B, n_frames, 16]
        (B = 1: frames up to, not including, EOS; B > 1: until every row has hit EOS, rows padded with code 0 after their EOS,
        the convention of batch_generate / batch_decode).  ``u`` [max_tokens, 16, B] uniforms in [0,1) (drawn from ``seed`` when
        omitted; parity tests inject them).

        ``batch_mode`` = the loop of ``batch_generate`` (qwen3_tts.py:1861-1935): ``left_padding`` [B] rows of zero embeddings in front
        of shorter prompts (masked keys, positions from cumsum(mask) - 1), ``trailing_text_hidden`` [B, n, H] right-padded with the
        pad embedding, finished rows forced to EOS, per-row trailing indices with the clamp-pad rule; returns (codes [B, n, 16],
        lengths [B]) with rows zero-padded after their EOS.  ``trailing_rule="standard"`` is the rule of the default (non-streaming)
        batch path, where every row behaves as a single sequence (continuous_batching.py:261-278: text while its index is inside the
        trailing text, pad afterwards): one extra pad row is appended so that the kernel's clamp lands on it."""
        t, cfg, dev = self.talker, self.config.talker_config, self.device
        x = input_embeds.to(dev).float().contiguous()
        B, P, H = x.shape
        g, V = cfg.num_code_groups, cfg.vocab_size
        eos = cfg.codec_eos_token_id
        if u is None:
            u = torch.rand(max_tokens, g, B, device=dev, generator=self._uniform_stream(seed))
        u = u.to(dev).float().contiguous()
        sp = {"temperature": float(temperature), "top_k": int(top_k), "top_p": float(top_p), "repetition_penalty": float(repetition_penalty),
              "eos": int(eos)}
        batch_mode = batch_mode or left_padding is not None
        self._kv_start = None
        if left_padding is not None and any(int(v) for v in left_padding):
            self._kv_start = torch.tensor([int(v) for v in left_padding], dtype=torch.int32, device=dev)
        self._finished = torch.zeros(B, dtype=torch.uint8, device=dev) if batch_mode else None
        self._tidx = torch.zeros(B, dtype=torch.int32, device=dev) if batch_mode else None
        t.reset_cache(B, P + max_tokens + 1)
        self._prefill_len = P
        self._trailing = trailing_text_hidden.to(dev).float().expand(B, -1, -1).contiguous() if trailing_text_hidden.shape[0] != B \
            else trailing_text_hidden.to(dev).float().contiguous()
        self._pad = tts_pad_embed.to(dev).float().reshape(-1).contiguous()
        if batch_mode and trailing_rule == "standard":
            self._trailing = torch.cat([self._trailing, self._pad.reshape(1, 1, H).expand(B, 1, H)], dim=1).contiguous()
        elif trailing_rule not in ("clamp_pad", "standard"):
            raise ValueError(f"trailing_rule must be 'clamp_pad' or 'standard', got {trailing_rule!r}")
        self._suppress = torch.zeros(V, device=dev)
        self._suppress[torch.tensor(self._suppress_codec_tokens(eos), device=dev)] = float("-inf")
        self._seen = torch.zeros(B, V, dtype=torch.uint8, device=dev)
        self._codes = torch.zeros(B, g, dtype=torch.int64, device=dev)
        self._u = torch.zeros(g, B, device=dev)
        self._x_in = torch.zeros(B, 1, H, device=dev)
        self._cp_in0 = torch.zeros(B, 2, H, device=dev)
        self._cp_in = torch.zeros(B, H, device=dev)
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        out = torch.zeros(B, max_tokens, g, dtype=torch.int64, device=dev)
        lengths_dev = torch.zeros(B, dtype=torch.int64, device=dev)
        done = torch.zeros(B, dtype=torch.bool)
        n = 0
        graph = None
        for step in range(max_tokens):
            self._u.copy_(u[step])
            if step == 0:
                self._frame(x, sp)                                                   # prefill frame (S = P rows), eager
            elif use_graph:
                if graph is None:
                    # warm-up on a side stream is not needed: every kernel has already run once in the prefill frame except
                    # the S = 1 GEMV variants, which the capture below launches for the first time (lazy module load is done).
                    torch.cuda.synchronize(dev)
                    bufs = [t.offset_dev, self._seen, self._codes, self._x_in] + ([self._finished, self._tidx] if batch_mode else [])
                    state = [b_.clone() for b_ in bufs]
                    restore = lambda: [b_.copy_(s_) for b_, s_ in zip(bufs, state)]
                    l0 = ops.LAUNCHES[0]
                    self._frame(self._x_in, sp)                                      # eager run of the S = 1 path (loads kernels)
                    self._frame_launches = ops.LAUNCHES[0] - l0
                    restore()
                    t.offset = P + step - 1
                    torch.cuda.synchronize(dev)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        self._frame(self._x_in, sp)
                    restore()
                    t.offset = P + step - 1
                graph.replay()
                t.offset = P + step
                ops.LAUNCHES[0] += self._frame_launches
            else:
                l0 = ops.LAUNCHES[0]
                self._frame(self._x_in, sp)
                self._frame_launches = ops.LAUNCHES[0] - l0
            if batch_mode:
                # No per-frame host read: a finished row is masked on the device (its frame keeps the zeros `out` starts with, its length
                # stops growing) and the all-finished test is a sync every 8th frame only -- the frames replayed past the last EOS record
                # nothing.  (One .cpu() per frame held the loop at ~12 ms per frame; the graph itself replays in under 5.)
                fin = self._finished.bool()
                out[:, n] = torch.where(fin[:, None], out[:, n], self._codes)
                lengths_dev += (~fin).to(torch.int64)
                n += 1
                if (step & 7) == 7 and bool(fin.all()):
                    break
                continue
            if stop_on_eos:
                hit = self._codes[:, 0].cpu() == eos                                 # the per-frame sync (EOS test)
                done |= hit
                if bool(done.all()):
                    break
                live = ~done
                out[live.to(dev), n] = self._codes[live.to(dev)]
            else:
                out[:, n] = self._codes
            n += 1
        if int(self._err.item()) != 0:
            raise ValueError("generate_codes: a sampled code indexed outside its embedding table")
        self._graph = graph
        if batch_mode:
            lengths = lengths_dev.cpu()
            n = int(lengths.max()) if lengths.numel() else 0
            return out[:, :n], lengths
        return out[:, :n]

    # ------------------------------------------------------------------ batch generation
    @torch.no_grad()
    def prepare_batch_inputs_from_ids(self, ids_list, language_id=None, speaker_ids=None, instruct_ids=None):
        """_prepare_batch_inputs (qwen3_tts.py:486-604) after tokenisation: left-pad the prompts with zero rows, right-pad the trailing
        text with the pad embedding.  Returns (input_embeds [B,P,H], trailing [B,n,H], tts_pad [1,1,H], left_padding [B])."""
        per = [self.prepare_generation_inputs_from_ids(ids, language_id, None if speaker_ids is None else speaker_ids[i],
                                                       instruct_ids=None if instruct_ids is None else instruct_ids[i])
               for i, ids in enumerate(ids_list)]
        pad = per[0][2]
        pmax = max(e.shape[1] for e, _, _ in per)
        tmax = max(tr.shape[1] for _, tr, _ in per)
        H = pad.shape[-1]
        x = torch.zeros(len(per), pmax, H, device=self.device)
        trailing = pad.reshape(1, 1, H).expand(len(per), tmax, H).clone()
        left = []
        for i, (e, tr, _) in enumerate(per):
            left.append(pmax - e.shape[1])
            x[i, pmax - e.shape[1]:] = e[0]
            trailing[i, : tr.shape[1]] = tr[0]
        return x, trailing, pad, left

    def batch_generate_from_ids(self, ids_list, *, language_id=None, speaker_ids=None, temperature: float = 0.9, max_tokens: int = 4096,
                                top_k: int = 50, top_p: float = 1.0, repetition_penalty: float = 1.05, seed: int = 0, u=None,
                                stream: bool = False, **kwargs):
        """``Model.batch_generate`` (qwen3_tts.py:1651-2060) for already-tokenised texts, one BatchGenerationResult per sequence.

        ``stream=False`` (the reference's default) follows its batch session (continuous_batching.py): every row generates exactly what it
        would generate alone from its own uniform stream (standard trailing-text rule), and is decoded by ``_decode_generated_codes`` --
        15-frame chunks with 5 frames of left context (qwen3_tts.py:1050-1083).  ``stream=True`` follows the streaming branch
        (qwen3_tts.py:1861-2040) with one final chunk per row: finished rows forced to EOS, clamp-pad trailing rule, one decode of the whole
        row (chunks of 300 + 25 context)."""
        from ..base import BatchGenerationResult
        if self.speech_tokenizer is None:
            raise ValueError("Speech tokenizer not loaded")
        t0 = time.perf_counter()
        x, trailing, pad, left = self.prepare_batch_inputs_from_ids(ids_list, language_id, speaker_ids)
        codes, lengths = self.generate_codes(x, trailing, pad, max_tokens=max_tokens, temperature=temperature, top_k=top_k, top_p=top_p,
                                             repetition_penalty=repetition_penalty, seed=seed, u=u, left_padding=left, batch_mode=True,
                                             trailing_rule="clamp_pad" if stream else "standard")
        seqs = [codes[b, : int(lengths[b])] for b in range(codes.shape[0])]
        if stream:
            audios, _ = self.speech_tokenizer.batch_decode([s_ for s_ in seqs if s_.shape[0] > 0])
        else:
            # rows of equal length share their decode launches: the chunks of _decode_generated_codes do not interact, so chunk j of every
            # row goes through the vocoder as one batch (identical samples; 8 rows x 3 chunks: 24 decoder passes -> 2)
            live = [s_ for s_ in seqs if s_.shape[0] > 0]
            by_len: Dict[int, List[int]] = {}
            for i, s_ in enumerate(live):
                by_len.setdefault(int(s_.shape[0]), []).append(i)
            audios = [None] * len(live)
            for n_, idxs in by_len.items():
                wav = self.speech_tokenizer.decoder.chunked_decode(torch.stack([live[i] for i in idxs]).transpose(1, 2), chunk_size=15, left_context_size=5)
                for j, i in enumerate(idxs):
                    audios[i] = wav[j, 0]
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        it = iter(audios)
        for b, s_ in enumerate(seqs):
            if s_.shape[0] == 0:
                continue
            a = next(it)
            yield BatchGenerationResult(audio=a, sequence_idx=b, samples=int(a.shape[0]), sample_rate=self.sample_rate, token_count=int(s_.shape[0]),
                                        audio_duration=format_duration(a.shape[0] / self.sample_rate), processing_time_seconds=dt,
                                        peak_memory_usage=torch.cuda.max_memory_allocated(self.device) / 1e9, is_streaming_chunk=stream,
                                        is_final_chunk=stream)

    # ------------------------------------------------------------------ decode + public generate
    @torch.no_grad()
    def _decode_generated_codes(self, codes: torch.Tensor, *, decode_chunk: int = 15, decode_ctx: int = 5) -> torch.Tensor:
        """qwen3_tts.py:1050-1083: codes [n, G] of one sequence -> audio [1920 n], decoded in ``decode_chunk``-frame pieces with up to
        ``decode_ctx`` frames of left context whose samples are dropped."""
        if codes.shape[0] == 0:
            return torch.zeros(0, dtype=torch.float32, device=self.device)
        # the reference's loop is chunked_decode's loop with (15, 5) in place of (300, 25): same chunk boundaries, same "context only
        # when start > context" rule, context samples dropped (speech_tokenizer.py:932-954 vs qwen3_tts.py:1066-1078)
        return self.speech_tokenizer.decoder.chunked_decode(codes[None].transpose(1, 2), chunk_size=decode_chunk, left_context_size=decode_ctx)[0, 0]

    @torch.no_grad()
    def _decode_chunk(self, codes: torch.Tensor, chunk_tokens: int = 300) -> torch.Tensor:
        """qwen3_tts.py:1017-1048: codes [1, T, 16] -> audio [samples], trimmed to the frames whose first code is > 0."""
        chunks = list(self.speech_tokenizer.streaming_decode(codes, chunk_tokens=chunk_tokens))
        audio = torch.cat(chunks, dim=-1)[0]
        valid = int((codes[..., 0] > 0).sum().item()) * self.speech_tokenizer.decode_upsample_rate
        if 0 < valid < audio.shape[0]:
            audio = audio[:valid]
        return audio

    def generate_from_ids(self, input_ids, *, language_id=None, speaker_id=None, temperature: float = 0.9, max_tokens: int = 4096,
                          top_k: int = 50, top_p: float = 1.0, repetition_penalty: float = 1.05, seed: int = 0, u=None, **kwargs):
        """``Model.generate`` (qwen3_tts.py:1122-1575) for one already-tokenised segment; yields one GenerationResult."""
        if self.speech_tokenizer is None:
            raise ValueError("Speech tokenizer not loaded")
        t0 = time.perf_counter()
        x, trailing, pad = self.prepare_generation_inputs_from_ids(input_ids, language_id, speaker_id)
        codes = self.generate_codes(x, trailing, pad, max_tokens=max_tokens, temperature=temperature, top_k=top_k, top_p=top_p,
                                    repetition_penalty=repetition_penalty, seed=seed, u=u)
        if codes.shape[1] == 0:
            return
        audio = self._decode_chunk(codes[:1])
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        samples = int(audio.shape[0])
        dur = samples / self.sample_rate
        yield GenerationResult(audio=audio, samples=samples, sample_rate=self.sample_rate, segment_idx=0, token_count=int(codes.shape[1]),
                               audio_duration=format_duration(dur), real_time_factor=dur / dt if dt > 0 else 0.0,
                               prompt={"tokens": int(codes.shape[1]), "tokens-per-sec": round(codes.shape[1] / dt, 2) if dt > 0 else 0},
                               audio_samples={"samples": samples, "samples-per-sec": round(samples / dt, 2) if dt > 0 else 0},
                               processing_time_seconds=dt, peak_memory_usage=torch.cuda.max_memory_allocated(self.device) / 1e9)

    def _generate_segments(self, text, split_pattern, speaker, language, instruct, **gen):
        if self.speech_tokenizer is None:
            raise ValueError("Speech tokenizer not loaded")
        # base path: segments are split AND stripped (qwen3_tts.py:1268-1271); the instruct paths pass split_pattern=None and the text as is
        segments = [t.strip() for t in text.split(split_pattern) if t.strip()] if split_pattern else [text]
        for idx, seg in enumerate(segments):
            t0 = time.perf_counter()
            x, trailing, pad = self._prepare_generation_inputs(seg, language=language, speaker=speaker, instruct=instruct)
            seg_gen = dict(gen)
            if seg_gen.get("seed") is not None:                       # a fixed seed still gives every segment its own draws
                seg_gen["seed"] = int(seg_gen["seed"]) + idx
            codes = self.generate_codes(x, trailing, pad, **seg_gen)
            if codes.shape[1] == 0:
                continue
            audio = self._decode_chunk(codes[:1])
            torch.cuda.synchronize(self.device)
            dt = time.perf_counter() - t0
            samples = int(audio.shape[0])
            dur = samples / self.sample_rate
            yield GenerationResult(audio=audio, samples=samples, sample_rate=self.sample_rate, segment_idx=idx, token_count=int(codes.shape[1]),
                                   audio_duration=format_duration(dur), real_time_factor=dur / dt if dt > 0 else 0.0,
                                   prompt={"tokens": int(codes.shape[1]), "tokens-per-sec": round(codes.shape[1] / dt, 2) if dt > 0 else 0},
                                   audio_samples={"samples": samples, "samples-per-sec": round(samples / dt, 2) if dt > 0 else 0},
                                   processing_time_seconds=dt, peak_memory_usage=torch.cuda.max_memory_allocated(self.device) / 1e9)

    def generate(self, text: str, voice: Optional[str] = None, instruct: Optional[str] = None, temperature: float = 0.9, speed: float = 1.0,
                 lang_code: str = "auto", ref_audio=None, ref_text: Optional[str] = None, split_pattern: str = "\n", max_tokens: int = 4096,
                 verbose: bool = False, stream: bool = False, streaming_interval: float = 2.0, top_k: int = 50, top_p: float = 1.0,
                 repetition_penalty: float = 1.05, seed: Optional[int] = None, **kwargs):
        """Model.generate (qwen3_tts.py:1122-1575): routes on ``tts_model_type`` exactly as the reference (same errors)."""
        gen = dict(max_tokens=max_tokens, temperature=temperature, top_k=top_k, top_p=top_p, repetition_penalty=repetition_penalty, seed=seed)
        kind = getattr(self.config, "tts_model_type", "base")
        if kind == "voice_design":
            if not instruct:
                raise ValueError("VoiceDesign model requires 'instruct' to describe the voice "
                                 "(e.g., 'A cheerful young female voice with high pitch')")
            yield from self._generate_segments(text, None, None, lang_code, instruct, **gen)        # one utterance: _generate_with_instruct does not split
            return
        if kind == "custom_voice":
            if not voice:
                raise ValueError(f"CustomVoice model requires 'voice' (speaker name) (e.g., {self.supported_speakers})")
            if voice.lower() not in [s.lower() for s in self.supported_speakers]:
                raise ValueError(f"Speaker '{voice}' not supported. Available: {self.supported_speakers}")
            yield from self._generate_segments(text, None, voice, lang_code, instruct, **gen)
            return
        if self.speech_tokenizer is None:
            raise ValueError("Speech tokenizer not loaded")
        if ref_audio is not None:
            # with ref_text: in-context cloning (_generate_icl); alone: x-vector cloning through the speaker encoder (qwen3_tts.py:382-383).
            # Neither encoder has a CUDA path yet -- refuse instead of silently synthesising the default voice.
            raise NotImplementedError("voice cloning from ref_audio needs the speech-tokenizer encoder + speaker encoder "
                                      "(SURVEY.md section 8f 'next'); call without ref_audio for the default voice")
        if stream:
            raise NotImplementedError("stream=True (incremental audio chunks) is not implemented for generate(); use "
                                      "batch_generate(stream=True) or speech_tokenizer.streaming_decode on the returned codes")
        if voice is not None and voice.lower() not in [s.lower() for s in self.supported_speakers]:
            raise ValueError(f"Voice '{voice}' is not supported by this Base model. Base models have no built-in preset voices — "
                             "clone a voice by passing ref_audio and ref_text instead.")
        yield from self._generate_segments(text, split_pattern, voice, lang_code, None, **gen)
