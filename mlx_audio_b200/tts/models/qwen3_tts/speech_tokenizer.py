"""Qwen3-TTS 12.5 Hz speech tokenizer, decode side, on B200 (reference: tts/models/qwen3_tts/speech_tokenizer.py).

``Qwen3TTSSpeechTokenizer(cfg).load_weights(...)``; ``decode(audio_codes[B,T,16]) -> (wav[B,samples], lengths)``
(speech_tokenizer.py:1099-1118), ``batch_decode`` (:1120-1179), ``streaming_decode`` (:1181-1217) and the decoder's
``__call__`` / ``chunked_decode`` (:843-880, 932-954).

B200 mapping: RVQ gather-sum in one kernel; every dense conv / linear runs on the tcgen05 conv kernel with the SnakeBeta /
LayerScale / gamma / residual / clip fused as prologue or epilogue; the 300-frame chunks of ``chunked_decode`` are
independent, so equal-length chunks are decoded as one batch instead of one after another.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .... import ops
from ....ops import ACT, Pre
from .config import Qwen3TTSTokenizerConfig, Qwen3TTSTokenizerDecoderConfig


def check_array_shape_qwen3(arr) -> bool:
    """True when a 3-D conv weight already is MLX-layout (out, K, in) (qwen3_tts.py:123-157)."""
    shape = tuple(arr.shape)
    if len(shape) != 3:
        return False
    _, dim2, dim3 = shape
    if dim2 == 1:
        return dim3 > 64
    if dim3 == 1:
        return not dim2 > 64
    return dim2 < dim3


class Qwen3TTSSpeechTokenizerDecoder:
    def __init__(self, config: Qwen3TTSTokenizerDecoderConfig, device="cuda"):
        self.config = config
        self.device = torch.device(device)
        self.total_upsample = int(np.prod(list(config.upsample_rates) + list(config.upsampling_ratios)))
        self._w = None

    # ------------------------------------------------------------------ weights
    def load_weights(self, weights, prefix="decoder."):
        """``weights``: MLX-side names (after ``Qwen3TTSSpeechTokenizer.sanitize``), conv weights [Cout, K, Cin/g]."""
        P, cfg, dev = {k[len(prefix):]: v for k, v in dict(weights).items() if k.startswith(prefix)}, self.config, self.device
        f = lambda t: t.float().to(dev).contiguous()
        conv = lambda pre, groups=1: ops.pack_conv(P[pre + ".weight"].float(), P.get(pre + ".bias"), groups, dev)
        lin = lambda pre: ops.pack_linear(P[pre + ".weight"].float(), P.get(pre + ".bias"), dev)

        def snake(pre):                                                 # SnakeBeta constants (speech_tokenizer.py:123-126)
            a, b = torch.exp(P[pre + ".alpha"].float()), torch.exp(P[pre + ".beta"].float())
            return Pre(act=ACT["snake"], a=f(a), b=f(1.0 / (b + 1e-9)))

        W = {}
        nsem, nq = cfg.num_semantic_quantizers, cfg.num_quantizers
        W["cb_first"] = f(torch.stack([P[f"quantizer.rvq_first.vq.layers.{i}.codebook.embed.weight"].float() for i in range(nsem)]))
        W["cb_rest"] = f(torch.stack([P[f"quantizer.rvq_rest.vq.layers.{i}.codebook.embed.weight"].float() for i in range(nq - nsem)]))
        W["proj_first"] = conv("quantizer.rvq_first.output_proj")
        W["proj_rest"] = conv("quantizer.rvq_rest.output_proj")
        W["pre_conv"] = conv("pre_conv.conv")
        T = "pre_transformer"
        W["in_proj"], W["out_proj"], W["norm"] = lin(T + ".input_proj"), lin(T + ".output_proj"), f(P[T + ".norm.weight"])
        W["layers"] = []
        for i in range(cfg.num_hidden_layers):
            L = f"{T}.layers.{i}"
            qkv = torch.cat([P[L + f".self_attn.{n}_proj.weight"].float() for n in "qkv"], dim=0)
            gu = torch.cat([P[L + ".mlp.gate_proj.weight"].float(), P[L + ".mlp.up_proj.weight"].float()], dim=0)
            W["layers"].append({
                "n1": f(P[L + ".input_layernorm.weight"]), "n2": f(P[L + ".post_attention_layernorm.weight"]),
                "qkv": ops.pack_linear(qkv, None, dev), "o": lin(L + ".self_attn.o_proj"),
                "gu": ops.pack_linear(gu, None, dev), "down": lin(L + ".mlp.down_proj"),
                "ls1": f(P[L + ".self_attn_layer_scale.scale"]), "ls2": f(P[L + ".mlp_layer_scale.scale"])})
        W["upsample"] = []
        for i, _ in enumerate(cfg.upsampling_ratios):
            U = f"upsample.{i}"
            W["upsample"].append({"up": conv(U + ".0.conv"), "dw": conv(U + ".1.dwconv.conv", cfg.latent_dim),
                                  "ln": (f(P[U + ".1.norm.weight"]), f(P[U + ".1.norm.bias"])),
                                  "pw1": lin(U + ".1.pwconv1"), "pw2": lin(U + ".1.pwconv2"), "gamma": f(P[U + ".1.gamma"])})
        W["init"] = conv("decoder.0.conv")
        W["blocks"] = []
        for bi, r in enumerate(cfg.upsample_rates):
            B_ = f"decoder.{bi + 1}.block"
            units = []
            for ui, d in enumerate((1, 3, 9)):
                U = f"{B_}.{ui + 2}"
                units.append({"d": d, "s1": snake(U + ".act1"), "c1": conv(U + ".conv1.conv"), "s2": snake(U + ".act2"), "c2": conv(U + ".conv2.conv")})
            W["blocks"].append({"r": r, "snake": snake(B_ + ".0"), "up": conv(B_ + ".1.conv"), "units": units})
        W["out_snake"] = snake(f"decoder.{len(cfg.upsample_rates) + 1}")
        W["out_conv"] = conv(f"decoder.{len(cfg.upsample_rates) + 2}.conv")
        self._w = W
        return self

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, codes: torch.Tensor, taps=None) -> torch.Tensor:
        """codes int [B, num_quantizers, T] -> audio [B, 1, 1920 T] clipped to [-1, 1] (speech_tokenizer.py:843-880)."""
        W, cfg, dev = self._w, self.config, self.device
        if codes.shape[1] != cfg.num_quantizers:
            raise ValueError(f"Expected {cfg.num_quantizers} layers of codes, got {codes.shape[1]}")
        codes = codes.to(device=dev, dtype=torch.int64).contiguous()
        B, nq, T = codes.shape
        nsem = cfg.num_semantic_quantizers
        x = ops.conv1d(ops.rvq_decode(codes[:, :nsem], W["cb_first"]), W["proj_first"])
        if nq > nsem:
            x = ops.conv1d(ops.rvq_decode(codes[:, nsem:], W["cb_rest"]), W["proj_rest"], res=x)
        h = ops.conv1d(x, W["pre_conv"], pad_left=2, lout=T)                                   # CausalConv1d k3
        if taps is not None:
            taps["pre_conv"] = h
        # ---- pre_transformer (speech_tokenizer.py:383-413): RMSNorm, RoPE (rotate_half), full causal, SwiGLU, LayerScale
        nh, hd, eps = cfg.num_attention_heads, cfg.head_dim, cfg.rms_norm_eps
        d = nh * hd
        x = ops.linear(h, W["in_proj"])
        for lw in W["layers"]:
            n = ops.layernorm(x, lw["n1"], None, eps=eps, rms=True)
            qkv = ops.linear(n, lw["qkv"])
            ops.rope_(qkv[:, :, :d], nh, offset=0, base=cfg.rope_theta, traditional=False)
            ops.rope_(qkv[:, :, d:2 * d], nh, offset=0, base=cfg.rope_theta, traditional=False)
            att = ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], n_heads=nh, scale=hd ** -0.5, causal=True)
            x = ops.linear(att, lw["o"], cscale=lw["ls1"], res=x)
            n = ops.layernorm(x, lw["n2"], None, eps=eps, rms=True)
            m = ops.swiglu(ops.linear(n, lw["gu"]))
            x = ops.linear(m, lw["down"], cscale=lw["ls2"], res=x)
        h = ops.linear(ops.layernorm(x, W["norm"], None, eps=eps, rms=True), W["out_proj"])
        if taps is not None:
            taps["transformer"] = h
        # ---- 2 x (CausalTransposeConv1d k=s=2, ConvNeXtBlock) (speech_tokenizer.py:86-159)
        for (uw, fct) in zip(W["upsample"], cfg.upsampling_ratios):
            h = ops.conv1d(h, uw["up"], stride=fct, pad_left=0, lout=h.shape[1] * fct, transpose=True)
            t = ops.conv1d(h, uw["dw"], pad_left=6, lout=h.shape[1])
            t = ops.layernorm(t, *uw["ln"], eps=1e-6)
            t = ops.linear(t, uw["pw1"], post_act=ACT["gelu"])
            h = ops.linear(t, uw["pw2"], cscale=uw["gamma"], res=h)
        if taps is not None:
            taps["upsample"] = h
        # ---- decoder: conv k7, 4 x (SnakeBeta, ConvT, 3 residual units), SnakeBeta, conv k7 -> 1, clip
        w = ops.conv1d(h, W["init"], pad_left=6, lout=h.shape[1])
        for bi, bw in enumerate(W["blocks"]):
            r = bw["r"]
            w = ops.conv1d(w, bw["up"], stride=r, pad_left=0, lout=w.shape[1] * r, pre=bw["snake"], transpose=True)
            for u in bw["units"]:
                t = ops.conv1d(w, u["c1"], dilation=u["d"], pad_left=6 * u["d"], lout=w.shape[1], pre=u["s1"])
                w = ops.conv1d(t, u["c2"], pre=u["s2"], res=w)
            if taps is not None:
                taps[f"block{bi}"] = w
        wav = ops.conv1d(w, W["out_conv"], pad_left=6, lout=w.shape[1], pre=W["out_snake"], post_act=ACT["clip1"])   # [B, L, 1]
        return wav.reshape(B, 1, -1)

    @torch.no_grad()
    def chunked_decode(self, codes: torch.Tensor, chunk_size: int = 300, left_context_size: int = 25) -> torch.Tensor:
        """speech_tokenizer.py:932-954, with chunks of equal length decoded as one batch (they do not interact)."""
        codes = codes.to(device=self.device, dtype=torch.int64)
        B, nq, T = codes.shape
        up = self.total_upsample
        spans, start = [], 0
        while start < T:
            end = min(start + chunk_size, T)
            ctx = left_context_size if start - left_context_size > 0 else start
            spans.append((start, end, ctx))
            start = end
        out = torch.empty(B, 1, T * up, device=self.device, dtype=torch.float32)
        groups: Dict[tuple, List[tuple]] = {}
        for sp in spans:
            groups.setdefault((sp[1] - sp[0] + sp[2], sp[2]), []).append(sp)
        for (length, ctx), members in groups.items():
            batch = torch.cat([codes[:, :, s - c: e] for (s, e, c) in members], dim=0)          # [n*B, nq, length]
            wav = self(batch)
            for i, (s, e, c) in enumerate(members):
                out[:, :, s * up: e * up] = wav[i * B: (i + 1) * B, :, c * up:]
        return out

    # streaming (speech_tokenizer.py:882-930) re-decodes with left context through streaming_decode below; the incremental
    # conv-buffer variant is SURVEY.md section 8f "next".
    def reset_streaming_state(self):
        pass


class Qwen3TTSSpeechTokenizer:
    """speech_tokenizer.py:1061-1217 (decode side)."""

    def __init__(self, config: Qwen3TTSTokenizerConfig, device="cuda"):
        self.config = config
        self.encoder_valid_num_quantizers = config.encoder_valid_num_quantizers
        self.input_sample_rate = config.input_sample_rate
        self.output_sample_rate = config.output_sample_rate
        self.decode_upsample_rate = config.decode_upsample_rate
        self.encode_downsample_rate = config.encode_downsample_rate
        self.decoder = Qwen3TTSSpeechTokenizerDecoder(config.decoder_config, device)
        self.encoder_model = None
        self.device = torch.device(device)

    @property
    def has_encoder(self) -> bool:
        return self.encoder_model is not None

    def load_weights(self, weights):
        self.decoder.load_weights(weights, prefix="decoder.")
        return self

    def encode(self, audio):
        raise ValueError("Encoder not available for this speech tokenizer")      # same error as speech_tokenizer.py:1092-1093

    def decode(self, audio_codes: torch.Tensor):
        """audio_codes [B, T, 16] -> (wav [B, samples], audio_lengths [B])."""
        audio_codes = audio_codes.to(self.device)
        wav = self.decoder.chunked_decode(audio_codes.transpose(1, 2)).squeeze(1)
        lengths = (audio_codes[..., 0] > 0).sum(dim=1) * self.decode_upsample_rate
        return wav, lengths

    def batch_decode(self, codes_list):
        """speech_tokenizer.py:1120-1179: pad to the longest with code 0, decode as one batch, trim per sequence."""
        if not codes_list:
            return [], []
        normed = [c[None] if c.dim() == 2 else c for c in codes_list]
        seq_lens = [c.shape[1] for c in normed]
        max_len, nq = max(seq_lens), normed[0].shape[2]
        batch = torch.zeros(len(normed), max_len, nq, dtype=torch.int64, device=self.device)
        for i, c in enumerate(normed):
            batch[i, : c.shape[1]] = c[0].to(self.device)
        wav = self.decoder.chunked_decode(batch.transpose(1, 2)).squeeze(1)
        lengths = [int(sl) * self.decode_upsample_rate for sl in seq_lens]
        audios = []
        for b, n in enumerate(lengths):
            a = wav[b]
            audios.append(a[:n] if 0 < n < a.shape[0] else a)
        return audios, lengths

    def streaming_decode(self, audio_codes: torch.Tensor, chunk_tokens: int = 100):
        """speech_tokenizer.py:1181-1217: yields [B, samples] per chunk of ``chunk_tokens`` frames (25 frames left context)."""
        codes = audio_codes.to(self.device).transpose(1, 2)
        total, start = codes.shape[-1], 0
        while start < total:
            end = min(start + chunk_tokens, total)
            ctx = 25 if start - 25 > 0 else start
            wav = self.decoder(codes[..., start - ctx: end])[..., ctx * self.decode_upsample_rate:]
            yield wav.squeeze(1)
            start = end

    @staticmethod
    def sanitize(weights):
        """Decoder half of speech_tokenizer.py:1220-1447: PyTorch conv layouts -> [out, K, in], transposed convs
        [in, out, K] -> [out, K, in], codebooks = embedding_sum / clip(cluster_usage, 1e-5)."""
        out, codebook = {}, {}
        for k, v in weights.items():
            if k.startswith("encoder."):
                continue                                                # encode side: SURVEY.md section 8f "next"
            if "_codebook.cluster_usage" in k or "_codebook.embedding_sum" in k:
                base = k.rsplit("._codebook.", 1)[0]
                codebook.setdefault(base, {})["cluster_usage" if "cluster_usage" in k else "embedding_sum"] = v
                continue
            is_tr = ("upsample" in k and ".0.conv.weight" in k) or ("decoder.decoder" in k and "block.1.conv.weight" in k)
            if is_tr and v.dim() == 3:
                v = v if check_array_shape_qwen3(v) else v.permute(1, 2, 0)
            elif ("conv.weight" in k or "_proj.weight" in k) and v.dim() == 3:
                v = v if check_array_shape_qwen3(v) else v.permute(0, 2, 1)
            out[k] = v
        for base, d in codebook.items():
            if "cluster_usage" in d and "embedding_sum" in d:
                out[f"{base}.codebook.embed.weight"] = d["embedding_sum"] / torch.clamp(d["cluster_usage"][:, None], min=1e-5)
        return out
