"""Whisper on B200 -- audio encoder path (reference: stt/models/whisper/whisper.py:280-448,501-530).

``Model(ModelDimensions)`` with ``load_weights``, ``encoder(mel)``, ``embed_audio``; the text decoder
loop is row "next-1" of SURVEY.md section 8f.  Fusions: conv+GELU, conv+GELU+positional embedding,
q|k|v as one GEMM (the key bias is structurally zero), residual adds in the GEMM epilogues.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch

from .... import ops
from ....ops import ACT
from ....tts.models.base import BaseModelArgs
from .audio import log_mel_spectrogram, pad_or_trim  # noqa: F401  (re-exported like the reference)


@dataclass
class ModelDimensions(BaseModelArgs):
    """whisper.py:280-322."""
    n_mels: int = 80
    n_audio_ctx: int = 1500
    n_audio_state: int = 768
    n_audio_head: int = 12
    n_audio_layer: int = 12
    n_vocab: int = 51865
    n_text_ctx: int = 448
    n_text_state: int = 768
    n_text_head: int = 12
    n_text_layer: int = 12

    @classmethod
    def from_dict(cls, config: dict) -> "ModelDimensions":
        """whisper.py:293-322: MLX-format keys (n_mels, ...) are filtered to the known fields; a HuggingFace transformers config
        (d_model / encoder_layers / ...) is mapped, with the reference's large-v3 defaults for absent keys."""
        config = dict(config)
        if "d_model" in config or "encoder_layers" in config:
            return cls(n_mels=config.get("num_mel_bins", 128), n_audio_ctx=config.get("max_source_positions", 1500),
                       n_audio_state=config.get("d_model", 1280), n_audio_head=config.get("encoder_attention_heads", 20),
                       n_audio_layer=config.get("encoder_layers", 32), n_vocab=config.get("vocab_size", 51866),
                       n_text_ctx=config.get("max_target_positions", 448), n_text_state=config.get("d_model", 1280),
                       n_text_head=config.get("decoder_attention_heads", 20), n_text_layer=config.get("decoder_layers", 32))
        known = {f.name for f in cls.__dataclass_fields__.values()}
        return cls(**{k: v for k, v in config.items() if k in known})


ModelConfig = ModelDimensions


def sinusoids(length, channels, max_timescale=10000):
    """whisper.py:329-335."""
    assert channels % 2 == 0
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float64))
    st = torch.arange(length, dtype=torch.float64)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1).to(torch.float32)


class AudioEncoder:
    def __init__(self, dims: ModelDimensions, device):
        self.dims, self.device, self._w = dims, device, None

    def load(self, P):
        dev, d = self.device, self.dims.n_audio_state
        f = lambda t: t.float().to(dev).contiguous()
        W = {"conv1": ops.pack_conv(P["encoder.conv1.weight"].float(), P["encoder.conv1.bias"], 1, dev),
             "conv2": ops.pack_conv(P["encoder.conv2.weight"].float(), P["encoder.conv2.bias"], 1, dev),
             "pos": sinusoids(self.dims.n_audio_ctx, d).to(dev)[None].contiguous(), "blocks": []}
        # conv2 (k3, stride 2, pad 1) as a stride-1 2-tap conv over row PAIRS (space-to-depth): y[l] = W0 x[2l-1] + W1 x[2l] + W2 x[2l+1]
        # = [0 | W0] . x'[l-1] + [W1 | W2] . x'[l] with x' = x viewed as [B, L/2, 2C] -- a free view, and a shape the tcgen05 conv takes.
        w2 = P["encoder.conv2.weight"].float()                          # [Cout, 3, Cin]
        w2p = torch.stack([torch.cat([torch.zeros_like(w2[:, 0]), w2[:, 0]], dim=1), torch.cat([w2[:, 1], w2[:, 2]], dim=1)], dim=1)
        W["conv2_s2d"] = ops.pack_conv(w2p, P["encoder.conv2.bias"], 1, dev)
        for i in range(self.dims.n_audio_layer):
            L = f"encoder.blocks.{i}"
            wqkv = torch.cat([P[f"{L}.attn.{n}.weight"].float() for n in ("query", "key", "value")], 0)
            bqkv = torch.cat([P[f"{L}.attn.query.bias"].float(), torch.zeros(d), P[f"{L}.attn.value.bias"].float()], 0)
            W["blocks"].append({
                "attn_ln": (f(P[L + ".attn_ln.weight"]), f(P[L + ".attn_ln.bias"])), "mlp_ln": (f(P[L + ".mlp_ln.weight"]), f(P[L + ".mlp_ln.bias"])),
                "qkv": ops.pack_linear(wqkv, bqkv, dev), "out": ops.pack_linear(P[L + ".attn.out.weight"].float(), P[L + ".attn.out.bias"], dev),
                "mlp1": ops.pack_linear(P[L + ".mlp1.weight"].float(), P[L + ".mlp1.bias"], dev),
                "mlp2": ops.pack_linear(P[L + ".mlp2.weight"].float(), P[L + ".mlp2.bias"], dev)})
        W["ln_post"] = (f(P["encoder.ln_post.weight"]), f(P["encoder.ln_post.bias"]))
        self._w = W

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """whisper.py:438-448: mel [B, 3000, n_mels] -> [B, 1500, d]."""
        W, dims = self._w, self.dims
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        x = ops.conv1d(x, W["conv1"], pad_left=1, post_act=ACT["gelu"])
        if x.shape[1] % 2 == 0:
            x = ops.conv1d(x.view(x.shape[0], x.shape[1] // 2, -1), W["conv2_s2d"], pad_left=1, lout=x.shape[1] // 2, post_act=ACT["gelu"], res=W["pos"])
        else:
            x = ops.conv1d(x, W["conv2"], stride=2, pad_left=1, post_act=ACT["gelu"], res=W["pos"])
        assert x.shape[1:] == (dims.n_audio_ctx, dims.n_audio_state), "incorrect audio shape"
        d, nh = dims.n_audio_state, dims.n_audio_head
        for blk in W["blocks"]:
            h = ops.layernorm(x, *blk["attn_ln"])
            qkv = ops.linear(h, blk["qkv"])
            att = ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], n_heads=nh, scale=(d // nh) ** -0.5)   # (d^-.25)^2
            x = ops.linear(att, blk["out"], res=x)
            h = ops.layernorm(x, *blk["mlp_ln"])
            m = ops.linear(h, blk["mlp1"], post_act=ACT["gelu"])
            x = ops.linear(m, blk["mlp2"], res=x)
        return ops.layernorm(x, *W["ln_post"])


@dataclass
class TokenizerSpec:
    """The tokenizer constants the decode loop needs (decoding.py:349-442; HFTokenizerWrapper properties, whisper.py:46-175).
    Defaults: Whisper's multilingual vocabulary.  ``suppress`` = ids to suppress (the reference's "-1" expands to
    tokenizer.non_speech_tokens, which needs the tokenizer files; pass the list when they are available)."""
    eot: int = 50257
    sot: int = 50258
    no_timestamps: int = 50363
    timestamp_begin: int = 50364
    no_speech: int = 50362
    blank_ids: tuple = (220,)
    language: int = 50259
    task: int = 50359
    suppress: tuple = ()
    transcribe: int = 50359
    translate: int = 50358
    sot_lm: int = 50360
    sot_prev: int = 50361
    non_speech_tokens: Optional[tuple] = None

    @property
    def sot_sequence(self):
        return (self.sot, self.language, self.task)


def get_suppress_tokens(spec: "TokenizerSpec", suppress_tokens=None) -> tuple:
    """decoding.py:79-112: the ids the SuppressTokens filter masks.  ``-1`` expands to the tokenizer's non-speech tokens; the
    transcribe / translate / sot / sot_prev / sot_lm markers and no_speech are always added.  A falsy ``suppress_tokens`` means the
    filter is not installed at all (decoding.py:489-495) and yields ()."""
    suppress_tokens = spec.suppress if suppress_tokens is None else suppress_tokens
    if not suppress_tokens:
        return ()
    result = list(suppress_tokens)
    if -1 in result:
        if spec.non_speech_tokens is None:
            raise ValueError("suppress contains -1 but the spec carries no non_speech_tokens (they come from the tokenizer files)")
        result = [t for t in result if t >= 0] + list(spec.non_speech_tokens)
    result.extend([spec.transcribe, spec.translate, spec.sot, spec.sot_prev, spec.sot_lm])
    if spec.no_speech is not None:
        result.append(spec.no_speech)
    return tuple(sorted(set(result)))


@dataclass
class DecodingResult:
    """decoding.py:152-162, field for field."""
    audio_features: torch.Tensor
    language: str
    language_probs: Optional[dict] = None
    tokens: list = field(default_factory=list)
    text: str = ""
    avg_logprob: float = float("nan")
    no_speech_prob: float = float("nan")
    temperature: float = float("nan")
    compression_ratio: float = float("nan")


@dataclass
class STTOutput:
    """whisper.py:272-276."""
    text: str
    segments: Optional[list] = None
    language: Optional[str] = None


def compression_ratio(text: str) -> float:
    """decoding.py:19-21."""
    import zlib
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


def results_from_greedy(tokens, sum_logprobs, no_speech_probs, audio_features, spec: "TokenizerSpec", sample_begin: int, language: str = "en",
                        temperature: float = 0.0, tokenizer=None):
    """The tail of DecodingTask.run (decoding.py:664-722) for one candidate per audio: GreedyDecoder.finalize appends an EOT, each row is cut
    to ``[sample_begin, first EOT)``, avg_logprob = sum_logprob / (len(tokens) + 1); text needs a tokenizer (else "", whose compression ratio is 0.0 as the reference computes it)."""
    out = []
    for i, row in enumerate(tokens):
        row = list(row) + [spec.eot]
        cut = row[sample_begin:]
        cut = cut[: cut.index(spec.eot)]
        text = tokenizer.decode(cut).strip() if tokenizer is not None else ""
        out.append(DecodingResult(audio_features=audio_features[i], language=language, tokens=cut, text=text,
                                  avg_logprob=float(sum_logprobs[i]) / (len(cut) + 1), no_speech_prob=float(no_speech_probs[i]),
                                  temperature=temperature, compression_ratio=compression_ratio(text)))
    return out


class TextDecoder:
    """whisper.py:451-498 with an in-place KV cache: self-attention K|V rows are written by the projection GEMM straight
    into a preallocated [B, n_ctx, 2d] buffer (the reference concatenates every step, whisper.py:359-361); cross-attention
    K|V are computed once per audio window."""

    def __init__(self, dims: ModelDimensions, device):
        self.dims, self.device, self._w = dims, device, None

    def load(self, P):
        dev, d = self.device, self.dims.n_text_state
        f = lambda t: t.float().to(dev).contiguous()
        emb = P["decoder.token_embedding.weight"].float()
        W = {"emb": f(emb), "pos": f(P["decoder.positional_embedding"]), "logits": ops.pack_linear(emb, None, dev), "blocks": []}
        for i in range(self.dims.n_text_layer):
            L = f"decoder.blocks.{i}"
            kv = lambda a: ops.pack_linear(torch.cat([P[f"{L}.{a}.key.weight"].float(), P[f"{L}.{a}.value.weight"].float()], 0),
                                           torch.cat([torch.zeros(d), P[f"{L}.{a}.value.bias"].float()], 0), dev)
            lin = lambda n: ops.pack_linear(P[f"{L}.{n}.weight"].float(), P[f"{L}.{n}.bias"], dev)
            ln = lambda n: (f(P[f"{L}.{n}.weight"]), f(P[f"{L}.{n}.bias"]))
            W["blocks"].append({"attn_ln": ln("attn_ln"), "cross_ln": ln("cross_attn_ln"), "mlp_ln": ln("mlp_ln"),
                                "q": lin("attn.query"), "kv": kv("attn"), "o": lin("attn.out"),
                                "cq": lin("cross_attn.query"), "ckv": kv("cross_attn"), "co": lin("cross_attn.out"),
                                "mlp1": lin("mlp1"), "mlp2": lin("mlp2")})
        W["ln"] = (f(P["decoder.ln.weight"]), f(P["decoder.ln.bias"]))
        self._w = W

    def new_cache(self, xa: torch.Tensor):
        """Allocate the self-attention cache and project the cross-attention K|V of ``xa`` [B,1500,d] once."""
        W, d = self._w, self.dims.n_text_state
        B = xa.shape[0]
        xa = xa.to(device=self.device, dtype=torch.float32).contiguous()
        return {"offset": 0, "self": [torch.empty(B, self.dims.n_text_ctx, 2 * d, device=self.device) for _ in W["blocks"]],
                "cross": [ops.linear(xa, blk["ckv"]) for blk in W["blocks"]]}

    @torch.no_grad()
    def __call__(self, tokens: torch.Tensor, cache: dict, last_only: bool = True, also_first: bool = False, also_at: Optional[int] = None):
        """tokens int64 [B,n] -> logits [B,V] of the last position (and of position ``also_at`` -- 0 when ``also_first`` -- for the
        no-speech probability at the sot token, decoding.py:610-612)."""
        if also_first and also_at is None:
            also_at = 0
        also_first = also_at is not None
        W, dims = self._w, self.dims
        d, nh = dims.n_text_state, dims.n_text_head
        B, n = tokens.shape
        off = cache["offset"]
        x = ops.gather_rows(W["emb"], tokens.reshape(-1).contiguous(), add=W["pos"][off:off + n]).reshape(B, n, d)
        scale = (d // nh) ** -0.5                                           # q and k each carry d^-0.25 (whisper.py:373-375)
        for blk, kvbuf, ckv in zip(W["blocks"], cache["self"], cache["cross"]):
            h = ops.layernorm(x, *blk["attn_ln"])
            q = ops.linear(h, blk["q"])
            ops.linear(h, blk["kv"], out=kvbuf[:, off:off + n])               # K|V rows land in the cache
            att = ops.attention(q, kvbuf[:, :off + n, :d], kvbuf[:, :off + n, d:], n_heads=nh, scale=scale, causal=True, q_offset=off)
            x = ops.linear(att, blk["o"], res=x)
            h = ops.layernorm(x, *blk["cross_ln"])
            q = ops.linear(h, blk["cq"])
            att = ops.attention(q, ckv[:, :, :d], ckv[:, :, d:], n_heads=nh, scale=scale)
            x = ops.linear(att, blk["co"], res=x)
            h = ops.layernorm(x, *blk["mlp_ln"])
            m = ops.linear(h, blk["mlp1"], post_act=ACT["gelu"])
            x = ops.linear(m, blk["mlp2"], res=x)
        cache["offset"] = off + n
        if not last_only:                                                     # every position: Model.logits / __call__ (whisper.py:623-631)
            return ops.linear(ops.layernorm(x.contiguous(), *W["ln"]), W["logits"])
        rows = x[:, -1:] if not also_first else torch.cat([x[:, -1:], x[:, also_at:also_at + 1]], 1)
        hl = ops.layernorm(rows.contiguous(), *W["ln"])
        logits = ops.linear(hl, W["logits"])                                  # tied embedding (whisper.py:498)
        return (logits[:, 0], logits[:, 1]) if also_first else logits[:, 0]


class Model:
    """whisper.py:501-530."""

    def __init__(self, dims: ModelDimensions, dtype=torch.float16, device="cuda"):
        self.dims, self.dtype, self.device = dims, dtype, torch.device(device)
        self.encoder = AudioEncoder(dims, self.device)
        self.decoder = TextDecoder(dims, self.device)

    @property
    def sample_rate(self):
        return 16000

    def eval(self):
        return self

    _HF_KEY_MAP = (                                                      # whisper.py:562-585; specific patterns before generic ones
        ("encoder.embed_positions.weight", None),                        # recomputed (sinusoids)
        ("decoder.embed_positions.weight", "decoder.positional_embedding"),
        ("encoder.layer_norm.", "encoder.ln_post."), ("decoder.layer_norm.", "decoder.ln."),
        ("encoder.layers.", "encoder.blocks."), ("decoder.layers.", "decoder.blocks."),
        (".self_attn_layer_norm.", ".attn_ln."), (".final_layer_norm.", ".mlp_ln."), (".encoder_attn_layer_norm.", ".cross_attn_ln."),
        (".fc1.", ".mlp1."), (".fc2.", ".mlp2."),
        (".self_attn.q_proj.", ".attn.query."), (".self_attn.k_proj.", ".attn.key."), (".self_attn.v_proj.", ".attn.value."),
        (".self_attn.out_proj.", ".attn.out."),
        (".encoder_attn.q_proj.", ".cross_attn.query."), (".encoder_attn.k_proj.", ".cross_attn.key."),
        (".encoder_attn.v_proj.", ".cross_attn.value."), (".encoder_attn.out_proj.", ".cross_attn.out."),
        ("decoder.embed_tokens.", "decoder.token_embedding."),
    )

    def sanitize(self, weights):
        """whisper.py:551-618: a HuggingFace checkpoint (keys under ``model.``) is renamed to the reference's module tree and its conv
        weights go from (out, in, K) to (out, K, in); an MLX-format checkpoint passes through.  Like the reference (whisper.py:613-615)
        every floating-point tensor is rounded to the model dtype (``Model(dims, dtype=torch.float16)`` by default), so an fp32 checkpoint
        runs with fp16-rounded weights -- which are fp16-exact and therefore take the single-plane fp16 tcgen05 weight path."""
        dt = getattr(self, "dtype", None)
        is_hf = any(k.startswith("model.") for k in weights)
        out = {}
        for k, v in weights.items():
            if "_positional_embedding" in k:
                continue
            if is_hf:
                if k.startswith("model."):
                    k = k[6:]
                skip = False
                for old, new in self._HF_KEY_MAP:
                    if old in k:
                        if new is None:
                            skip = True
                            break
                        k = k.replace(old, new)
                if skip:
                    continue
                if ("conv1.weight" in k or "conv2.weight" in k) and v.dim() == 3:
                    v = v.permute(0, 2, 1).contiguous()
            if dt is not None and torch.is_tensor(v) and v.is_floating_point() and v.dtype != dt:
                v = v.to(dt)
            out[k] = v
        return out

    def load_weights(self, weights, strict=False):
        P = dict(weights)
        if any(k.startswith("encoder.") for k in P):
            self.encoder.load(P)
        if any(k.startswith("decoder.") for k in P):
            self.decoder.load(P)
        return self

    @torch.no_grad()
    def greedy_decode(self, audio_features: torch.Tensor, spec: Optional[TokenizerSpec] = None, sample_len: Optional[int] = None,
                      max_initial_timestamp_index: Optional[int] = 50, without_timestamps: bool = False, *, temperature: float = 0.0,
                      uniforms=None, prompt=None):
        """DecodingTask.run's sampling loop with GreedyDecoder (decoding.py:295-325, 588-632) on encoder features [B,1500,d]: the whole step
        -- logit filters, argmax or categorical draw, log-prob bookkeeping -- is one fused kernel on the device-resident token history; the
        only host read per step is the `completed` flag (the reference also syncs on it, decoding.py:625).

        ``temperature`` > 0 samples from softmax(filtered / temperature); ``uniforms`` [steps, B] in [0,1) drives the draws (inverse CDF in
        index order; parity tests inject it, otherwise drawn on the device).  ``prompt`` = previous-context token ids: the initial tokens
        become [sot_prev] + prompt[-(n_text_ctx // 2 - 1):] + sot sequence (decoding.py:538-549).
        Returns (tokens list per row incl. the initial tokens, sum_logprobs [B], no_speech_probs [B]); ``self.last_sample_begin`` holds the
        index of the first sampled position."""
        spec = spec or TokenizerSpec()
        dims, dev = self.dims, self.device
        B = audio_features.shape[0]
        sample_len = sample_len or dims.n_text_ctx // 2
        V = dims.n_vocab
        init = list(spec.sot_sequence) + ([spec.no_timestamps] if without_timestamps else [])   # decoding.py:463-465
        if prompt:
            init = [spec.sot_prev] + [int(t) for t in prompt][-(dims.n_text_ctx // 2 - 1):] + init
        sb = len(init)
        sot_index = init.index(spec.sot)
        self.last_sample_begin = sb
        tokens = torch.zeros(B, dims.n_text_ctx + 2, dtype=torch.int64, device=dev)
        tokens[:, :sb] = torch.tensor(init, device=dev)
        neg = float("-inf")
        sup = torch.zeros(V, device=dev)
        suppress = get_suppress_tokens(spec)
        if suppress:
            sup[list(suppress)] = neg
        blank = torch.zeros(V, device=dev)
        blank[list(spec.blank_ids) + [spec.eot]] = neg
        sum_lp = torch.zeros(B, device=dev)
        not_done = torch.zeros(1, dtype=torch.int32, device=dev)
        cache = self.decoder.new_cache(audio_features)
        cur = sb
        no_speech = None
        temperature = float(temperature)
        if temperature > 0:
            if uniforms is None:
                if getattr(self, "_rng", None) is None:
                    self._rng = torch.Generator(device=dev)
                    self._rng.seed()
                uniforms = torch.rand(sample_len, B, device=dev, generator=self._rng)
            uniforms = torch.as_tensor(uniforms, dtype=torch.float32).to(dev).reshape(-1, B).contiguous()
        for i in range(sample_len):
            if i == 0:
                logits, first = self.decoder(tokens[:, :cur], cache, also_at=sot_index)
                no_speech = torch.softmax(first, dim=-1)[:, spec.no_speech]            # one-off, not on the per-step path
            else:
                logits = self.decoder(tokens[:, cur - 1:cur], cache)
            not_done.zero_()
            nxt = ops.whisper_greedy_step(logits, tokens, cur, sb, suppress_mask=sup if suppress else None, blank_mask=blank,
                                          eot=spec.eot, no_timestamps=spec.no_timestamps, timestamp_begin=spec.timestamp_begin,
                                          max_initial_ts=-1 if max_initial_timestamp_index is None else max_initial_timestamp_index,
                                          without_timestamps=without_timestamps, sum_logprobs=sum_lp, not_done=not_done,
                                          temperature=temperature, u=uniforms[i] if temperature > 0 else None)
            tokens[:, cur] = nxt
            cur += 1
            if int(not_done.item()) == 0 or cur > dims.n_text_ctx:
                break
        return tokens[:, :cur].cpu().tolist(), sum_lp, no_speech

    def decode(self, mel: torch.Tensor, spec: Optional[TokenizerSpec] = None, sample_len: Optional[int] = None,
               without_timestamps: bool = False, max_initial_timestamp: Optional[float] = 1.0, tokenizer=None, language: str = "en", *,
               temperature: float = 0.0, uniforms=None, prompt=None):
        """decoding.decode / DecodingTask.run (decoding.py:634-765): ``mel`` [.., 3000, n_mels] log-mel windows, or encoder features
        [.., n_audio_ctx, n_audio_state] which skip the encoder (decoding.py:557-565) -> DecodingResult per window (one object for a single
        window).  Text needs ``tokenizer`` (anything with ``decode(list[int]) -> str``)."""
        from .audio import CHUNK_LENGTH
        spec = spec or TokenizerSpec()
        single = mel.dim() == 2
        mel = mel[None] if single else mel
        dims = self.dims
        feats = mel if tuple(mel.shape[-2:]) == (dims.n_audio_ctx, dims.n_audio_state) else self.encoder(mel)
        index = round(max_initial_timestamp / (CHUNK_LENGTH / dims.n_audio_ctx)) if max_initial_timestamp else None
        tokens, sum_lp, no_speech = self.greedy_decode(feats, spec, sample_len, index, without_timestamps, temperature=temperature,
                                                       uniforms=uniforms, prompt=prompt)
        res = results_from_greedy(tokens, sum_lp.cpu().tolist(), no_speech.cpu().tolist(), feats, spec, self.last_sample_begin, language,
                                  float(temperature), tokenizer)
        return res[0] if single else res

    def embed_audio(self, mel):
        return self.encoder(mel)

    def logits(self, tokens: torch.Tensor, audio_features: torch.Tensor) -> torch.Tensor:
        """whisper.py:623-624: logits [B, n, n_vocab] of every position of ``tokens`` [B, n] given encoder features."""
        tokens = tokens.to(device=self.device, dtype=torch.int64)
        return self.decoder(tokens, self.decoder.new_cache(audio_features.to(self.device)), last_only=False)

    def __call__(self, mel: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
        """whisper.py:630-631."""
        return self.logits(tokens, self.encoder(mel))

    @property
    def is_multilingual(self) -> bool:
        return self.dims.n_vocab >= 51865                                    # whisper.py:633-635

    @property
    def num_languages(self) -> int:
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)         # whisper.py:637-639

    def encode_audio(self, audio: torch.Tensor) -> torch.Tensor:
        """BASELINE config 3: audio [B, 480000] -> log-mel (reference's +30 s zero pad, first 3000 frames) -> encoder."""
        from .audio import N_FRAMES, N_SAMPLES
        mel = log_mel_spectrogram(audio, self.dims.n_mels, padding=N_SAMPLES, device=self.device)
        return self.encoder(mel[:, :N_FRAMES])

    def detect_language(self, mel: torch.Tensor, spec: Optional[TokenizerSpec] = None):
        """decoding.detect_language (decoding.py:20-77): one decoder pass on [sot], softmax over the language tokens only.
        Returns (language token ids [n], list of {code: probability})."""
        from .transcribe import LANGUAGE_CODES
        spec = spec or TokenizerSpec()
        single = mel.dim() == 2
        mel = mel[None] if single else mel
        feats = mel if tuple(mel.shape[-2:]) == (self.dims.n_audio_ctx, self.dims.n_audio_state) else self.encoder(mel)
        n = feats.shape[0]
        x = torch.full((n, 1), spec.sot, dtype=torch.int64, device=self.device)
        logits = self.logits(x, feats)[:, 0]
        ids = list(range(spec.sot + 1, spec.sot + 1 + self.num_languages))
        sub = logits[:, ids].double()
        probs = torch.softmax(sub, dim=-1).cpu()
        toks = torch.tensor(ids)[sub.argmax(dim=-1).cpu()]
        dicts = [{c: float(probs[i, j]) for j, c in enumerate(LANGUAGE_CODES[: len(ids)])} for i in range(n)]
        return (toks[0], dicts[0]) if single else (toks, dicts)

    def generate(self, audio, **kw):
        """Model.generate (whisper.py:799-1318): 30-second windows with the seek rule, temperature fallback, previous-text conditioning.
        See ``transcribe.transcribe`` for the arguments; ``tokenizer`` (``decode(list[int]) -> str``, ``encode``) is needed for text, token
        ids are always returned in the segments."""
        from .transcribe import transcribe
        return transcribe(self, audio, **kw)
