"""Kyutai Mimi codec, decode side, on B200 (reference: codec/models/mimi/mimi.py + modules/*).

``Mimi(mimi_202407(nq)).load_weights(...)``, ``decode(codes[B,nq,T]) -> [B,1,1920 T]`` (mimi.py:155-162).
RVQ gather-sum in one kernel, windowed causal attention (context 250) without materialising the
O(T^2) mask the reference builds (transformer.py:98-107), ELU / LayerScale / residual adds fused into
the neighbouring convs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import math

import torch

from ... import ops
from ...ops import ACT, Pre


@dataclass
class MimiConfig:
    """Flat restatement of MimiConfig / SeanetConfig / TransformerConfig (mimi.py:35-96)."""
    dimension: int = 512
    nfilters: int = 64
    ratios: list = field(default_factory=lambda: [8, 6, 5, 4])
    ksize: int = 7
    residual_ksize: int = 3
    last_ksize: int = 3
    compress: int = 2
    num_heads: int = 8
    num_layers: int = 8
    dim_feedforward: int = 2048
    context: int = 250
    max_period: float = 10000.0
    nq: int = 32
    bins: int = 2048
    qdim: int = 256
    upsample_stride: int = 2
    sample_rate: float = 24000.0
    frame_rate: float = 12.5


def mimi_202407(num_codebooks: int) -> MimiConfig:
    return MimiConfig(nq=num_codebooks)


class Mimi:
    def __init__(self, cfg: MimiConfig, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self._w = None
        self._stream_codes = None

    @property
    def frame_rate(self):
        return self.cfg.frame_rate

    @property
    def sample_rate(self):
        return self.cfg.sample_rate

    def reset_state(self):
        """mimi.py:138-144: forget the streaming state."""
        self._stream_codes = None

    @torch.no_grad()
    def decode_step(self, xs: torch.Tensor) -> torch.Tensor:
        """mimi.py:171-176: the next ``T_new`` code frames [B, nq, T_new] -> their 1920 T_new samples.  In the reference the incremental
        path (conv buffers, overlap-add with the bias handled, rotating KV cache) returns exactly the corresponding slice of a one-shot
        decode (7e-16 when its own code is run, tests/golden/make_codec_golden.py); here the codes seen so far are kept and re-decoded --
        the same samples, with the cost of a full decode per call (a 10 000-frame decode is 34 ms).  ``reset_state()`` / ``decode()`` start a
        new stream, as they do in the reference."""
        xs = xs.to(device=self.device, dtype=torch.int64)
        prev = getattr(self, "_stream_codes", None)
        codes = xs if prev is None else torch.cat([prev, xs], dim=-1)
        pcm = self.decode(codes)
        self._stream_codes = codes
        hop = pcm.shape[-1] // codes.shape[-1]
        return pcm[..., (codes.shape[-1] - xs.shape[-1]) * hop:]

    @property
    def span_halo(self) -> int:
        """Left context (code frames) that makes a span decode exact: the stack is causal; each of the transformer's layers looks back
        ``context`` positions (at ``upsample_stride`` positions per frame), so the receptive field of the stack is num_layers * context
        positions, plus a few frames for the causal convolutions either side of it."""
        c = self.cfg
        return -(-(c.num_layers * c.context) // c.upsample_stride) + 16

    @torch.no_grad()
    def decode_span(self, codes: torch.Tensor, start: int, end: int, halo=None) -> torch.Tensor:
        """Samples of code frames [start, end) -- equal to that slice of ``decode(codes)`` -- from the frames themselves plus ``halo``
        frames of left context (one stream sharded across GPUs).  Returns [B, 1, (end - start) * 1920]."""
        halo = self.span_halo if halo is None else halo
        T = codes.shape[-1]
        if not 0 <= start < end <= T:
            raise ValueError(f"decode_span: need 0 <= start < end <= {T}")
        rs = max(0, start - halo)
        pcm = self.decode(codes[:, :, rs:end])
        hop = pcm.shape[-1] // (end - rs)
        return pcm[..., (start - rs) * hop:]

    @staticmethod
    def sanitize_pytorch_weights(weights: dict) -> dict:
        """The key / layout mapping of ``Mimi.load_pytorch_weights`` (mimi.py:196-249): kyutai's PyTorch checkpoint names -> the reference's
        module tree (leading underscores dropped, SEANet ``model.N`` indices -> ``layers.i.{upsample,downsample,residuals.0}``, transformer
        ``in_proj_weight`` / ``linearN`` renames), conv weights (out, in, K) -> (out, K, in), transposed-conv weights (in, out/g, K) ->
        (out, K, in/g)."""
        out = {}
        for k, v in weights.items():
            k = ".".join(s.removeprefix("_") for s in k.split("."))
            if k.startswith("encoder.model."):
                k = k.replace("encoder.model.", "encoder.")
            if k.startswith("decoder.model."):
                k = k.replace("decoder.model.", "decoder.")
            if k.endswith(".in_proj_weight"):
                k = k.replace(".in_proj_weight", ".in_proj.weight")
            if k.endswith(".linear1.weight"):
                k = k.replace(".linear1.weight", ".gating.linear1.weight")
            if k.endswith(".linear2.weight"):
                k = k.replace(".linear2.weight", ".gating.linear2.weight")
            for li, di in enumerate((2, 5, 8, 11)):
                k = k.replace(f"decoder.{di}.", f"decoder.layers.{li}.upsample.")
                k = k.replace(f"decoder.{di + 1}.", f"decoder.layers.{li}.residuals.0.")
            for li, ei in enumerate((1, 4, 7, 10)):
                k = k.replace(f"encoder.{ei}.", f"encoder.layers.{li}.residuals.0.")
                k = k.replace(f"encoder.{ei + 2}.", f"encoder.layers.{li}.downsample.")
            k = k.replace("decoder.0.", "decoder.init_conv1d.").replace("decoder.14.", "decoder.final_conv1d.")
            k = k.replace("encoder.0.", "encoder.init_conv1d.").replace("encoder.14.", "encoder.final_conv1d.")
            k = k.replace(".block.1.", ".block.0.").replace(".block.3.", ".block.1.")
            if k.endswith((".conv.weight", ".output_proj.weight", ".input_proj.weight")):
                v = v.transpose(-1, -2)
            if k.endswith(".convtr.weight"):
                v = v.permute(0, 2, 1) if (v.dim() == 3 and v.shape[1] == 1) else v.permute(1, 2, 0)
            out[k] = v.contiguous()
        return out

    @classmethod
    def from_pretrained(cls, repo_id, filename: str = "tokenizer-e351c8d8-checkpoint125.safetensors", device="cuda"):
        """mimi.py:264-275: the 32-codebook 2024-07 configuration from kyutai's checkpoint; ``repo_id`` may be a local directory."""
        from pathlib import Path
        f = Path(repo_id) / filename
        if not f.exists():
            from huggingface_hub import hf_hub_download
            f = Path(hf_hub_download(repo_id, filename))
        return cls(mimi_202407(32), device=device).load_pytorch_weights(f, strict=True)

    def load_pytorch_weights(self, file, strict: bool = True):
        """mimi.py:192-262: ``file`` = path of kyutai's safetensors checkpoint (or an already loaded dict)."""
        if not isinstance(file, dict):
            from safetensors.torch import load_file
            file = load_file(str(file))
        return self.load_weights(self.sanitize_pytorch_weights(file), strict=strict)

    def load_weights(self, weights, strict=True):
        P, cfg, dev = dict(weights), self.cfg, self.device
        f = lambda t: t.float().to(dev).contiguous()
        bf = lambda t: t.float().to(torch.bfloat16).float()

        def emb(pre):                                                   # quantization.py:26-30
            usage = torch.clamp(P[pre + ".cluster_usage"].float(), min=1e-5)[:, None]
            return P[pre + ".embedding_sum"].float() / usage

        W = {}
        W["cb_first"] = f(torch.stack([emb("quantizer.rvq_first.vq.layers.0.codebook")]))
        W["cb_rest"] = f(torch.stack([emb(f"quantizer.rvq_rest.vq.layers.{i}.codebook") for i in range(cfg.nq - 1)])) if cfg.nq > 1 else None
        W["proj_first"] = ops.pack_conv(P["quantizer.rvq_first.output_proj.weight"].float(), None, 1, dev)
        W["proj_rest"] = ops.pack_conv(P["quantizer.rvq_rest.output_proj.weight"].float(), None, 1, dev) if cfg.nq > 1 else None
        W["upsample"] = ops.pack_conv(P["upsample.convtr.convtr.convtr.weight"].float(), None, cfg.dimension, dev)
        def tlayers(root):
            out = []
            for li in range(cfg.num_layers):
                L = f"{root}.transformer.layers.{li}"
                out.append({
                    "n1": (f(P[L + ".norm1.weight"]), f(P[L + ".norm1.bias"])), "n2": (f(P[L + ".norm2.weight"]), f(P[L + ".norm2.bias"])),
                    "in_proj": ops.pack_linear(P[L + ".self_attn.in_proj.weight"].float(), None, dev),
                    "out_proj": ops.pack_linear(P[L + ".self_attn.out_proj.weight"].float(), None, dev),
                    "l1": ops.pack_linear(P[L + ".gating.linear1.weight"].float(), None, dev),
                    "l2": ops.pack_linear(P[L + ".gating.linear2.weight"].float(), None, dev),
                    "ls1": f(P[L + ".layer_scale_1.scale"]), "ls2": f(P[L + ".layer_scale_2.scale"])})
            return out

        W["layers"] = []
        for li in range(cfg.num_layers):
            L = f"decoder_transformer.transformer.layers.{li}"
            W["layers"].append({
                "n1": (f(P[L + ".norm1.weight"]), f(P[L + ".norm1.bias"])), "n2": (f(P[L + ".norm2.weight"]), f(P[L + ".norm2.bias"])),
                "in_proj": ops.pack_linear(P[L + ".self_attn.in_proj.weight"].float(), None, dev),
                "out_proj": ops.pack_linear(P[L + ".self_attn.out_proj.weight"].float(), None, dev),
                "l1": ops.pack_linear(P[L + ".gating.linear1.weight"].float(), None, dev),
                "l2": ops.pack_linear(P[L + ".gating.linear2.weight"].float(), None, dev),
                "ls1": f(P[L + ".layer_scale_1.scale"]), "ls2": f(P[L + ".layer_scale_2.scale"])})
        cw = lambda pre: ops.pack_conv(P[pre + ".weight"].float(), P.get(pre + ".bias"), 1, dev)
        W["init"] = cw("decoder.init_conv1d.conv.conv")
        W["dec"] = []
        for li, r in enumerate(cfg.ratios):
            L = f"decoder.layers.{li}"
            W["dec"].append({"r": r, "up": cw(L + ".upsample.convtr.convtr"), "c0": cw(L + ".residuals.0.block.0.conv.conv"),
                             "c1": cw(L + ".residuals.0.block.1.conv.conv")})
        W["final"] = cw("decoder.final_conv1d.conv.conv")
        del bf
        self._w = W
        self._enc = None
        if "encoder.init_conv1d.conv.conv.weight" in P:                 # encode side (seanet.py:194-199, mimi.py:146-153, quantization.py:178-185)
            E = {"init": cw("encoder.init_conv1d.conv.conv"), "layers": [], "final": cw("encoder.final_conv1d.conv.conv"),
                 "tr": tlayers("encoder_transformer"), "down": cw("downsample.conv.conv.conv")}
            for li, r in enumerate(reversed(cfg.ratios)):
                L = f"encoder.layers.{li}"
                E["layers"].append({"r": r, "c0": cw(L + ".residuals.0.block.0.conv.conv"), "c1": cw(L + ".residuals.0.block.1.conv.conv"),
                                    "down": cw(L + ".downsample.conv.conv")})
            for name, cb in (("first", W["cb_first"]), ("rest", W["cb_rest"])):
                if cb is None:
                    continue
                E["in_" + name] = ops.pack_conv(P[f"quantizer.rvq_{name}.input_proj.weight"].float(), None, 1, dev)
                E["c2_" + name] = ((cb.double() ** 2).sum(-1) / 2).contiguous()             # |e|^2 / 2 of argmin(|e|^2 / 2 - x.e)
            self._enc = E
        return self

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes int64 [B, nq, T] -> pcm [B, 1, 1920 T]."""
        W, cfg, dev = self._w, self.cfg, self.device
        self._stream_codes = None                                       # decode() resets the streaming state (mimi.py:156-158)
        codes = codes.to(device=dev, dtype=torch.int64).contiguous()
        B, nq, T = codes.shape
        q = ops.rvq_decode(codes[:, :1], W["cb_first"])
        x = ops.conv1d(q, W["proj_first"])
        if nq > 1:
            q2 = ops.rvq_decode(codes[:, 1:], W["cb_rest"][: nq - 1])
            x = ops.conv1d(q2, W["proj_rest"], res=x)
        s = cfg.upsample_stride
        x = ops.conv1d(x, W["upsample"], stride=s, pad_left=0, lout=T * s, transpose=True)         # causal: trim k-s on the right
        x = self._transformer(x, W["layers"])
        elu = Pre(act=ACT["elu"])
        x = ops.conv1d(x, W["init"], pad_left=cfg.ksize - 1, lout=x.shape[1])
        for lw in W["dec"]:
            r = lw["r"]
            y = ops.conv1d(x, lw["up"], stride=r, pad_left=0, lout=x.shape[1] * r, pre=elu, transpose=True)
            t = ops.conv1d(y, lw["c0"], pad_left=cfg.residual_ksize - 1, lout=y.shape[1], pre=elu)
            x = ops.conv1d(t, lw["c1"], pre=elu, res=y)
        pcm = ops.conv1d(x, W["final"], pad_left=cfg.last_ksize - 1, lout=x.shape[1], pre=elu)      # [B, L, 1]
        return pcm.reshape(B, 1, -1)

    def _transformer(self, x: torch.Tensor, layers) -> torch.Tensor:
        """ProjectedTransformer (mimi/modules/transformer.py:63-261), fresh cache: pre-norm layers, traditional RoPE, causal attention
        inside a ``context``-position window, LayerScale on both residual branches."""
        cfg = self.cfg
        d, nh = cfg.dimension, cfg.num_heads
        for lw in layers:
            n1 = ops.layernorm(x, *lw["n1"], eps=1e-5)
            qkv = ops.linear(n1, lw["in_proj"])
            ops.rope_(qkv[:, :, :d], nh, offset=0, base=cfg.max_period, traditional=True)
            ops.rope_(qkv[:, :, d:2 * d], nh, offset=0, base=cfg.max_period, traditional=True)
            att = ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], n_heads=nh, scale=(d // nh) ** -0.5,
                                causal=True, window=cfg.context)
            x = ops.linear(att, lw["out_proj"], cscale=lw["ls1"], res=x)
            n2 = ops.layernorm(x, *lw["n2"], eps=1e-5)
            m = ops.linear(n2, lw["l1"], post_act=ACT["gelu_tanh"])
            x = ops.linear(m, lw["l2"], cscale=lw["ls2"], res=x)
        return x

    @staticmethod
    def _cconv(x, cw, ksize, stride=1, pre=None, pad_mode=0, res=None):
        """StreamableConv1d (mimi/modules/conv.py:224-243), causal: k - stride samples of left padding, the right edge padded up to a whole
        last frame (zeros, or the edge sample for ``pad_mode=1``)."""
        L = x.shape[1]
        pad_total = ksize - stride
        lout = int(math.ceil(max(L + pad_total - ksize, 0) / stride + 1.0))
        return ops.conv1d(x, cw, stride=stride, pad_left=pad_total, lout=lout, pad_mode=pad_mode, pre=pre, res=res)

    @torch.no_grad()
    def encode_latent(self, xs: torch.Tensor) -> torch.Tensor:
        """pcm [B, 1, n] -> the 12.5 Hz latent [B, ceil(n / 1920), 512] in front of the quantiser (mimi.py:146-152)."""
        if self._enc is None:
            raise ValueError("Mimi.encode: the loaded weights have no encoder (encoder.*, encoder_transformer.*, downsample.*)")
        E, cfg = self._enc, self.cfg
        x = xs.to(device=self.device, dtype=torch.float32)
        x = x.reshape(x.shape[0], -1, 1)                                 # [B, 1, n] -> [B, n, 1] (one channel: same memory)
        elu = Pre(act=ACT["elu"])
        x = self._cconv(x, E["init"], cfg.ksize)
        for lw in E["layers"]:
            t = self._cconv(x, lw["c0"], cfg.residual_ksize, pre=elu)
            y = self._cconv(t, lw["c1"], 1, pre=elu, res=x)               # block(x) + x  (seanet.py:61-66)
            x = self._cconv(y, lw["down"], 2 * lw["r"], stride=lw["r"], pre=elu)
        x = self._cconv(x, E["final"], cfg.last_ksize, pre=elu)
        x = self._transformer(x, E["tr"])
        s = cfg.upsample_stride
        return self._cconv(x, E["down"], 2 * s, stride=s, pad_mode=1)

    @torch.no_grad()
    def encode(self, xs: torch.Tensor) -> torch.Tensor:
        """mimi.py:146-153: pcm [B, 1, n] -> int64 codes [B, nq, ceil(n / 1920)]: SEANet encoder, encoder transformer, stride-2 replicate-padded
        down-sampling conv, then the split residual quantiser (quantization.py:178-185): the first codebook on its own projection, the other
        nq - 1 as a residual chain on theirs -- `rvq_encode_kernel` runs the chain (argmin |e|^2 / 2 - x.e, first index on ties)."""
        z = self.encode_latent(xs)
        E, W = self._enc, self._w
        B, T, _ = z.shape
        r1 = ops.conv1d(z, E["in_first"])
        codes = [ops.rvq_encode(r1.reshape(B * T, -1), W["cb_first"], E["c2_first"]).reshape(B, T, 1)]
        if W["cb_rest"] is not None:
            r2 = ops.conv1d(z, E["in_rest"])
            codes.append(ops.rvq_encode(r2.reshape(B * T, -1), W["cb_rest"], E["c2_rest"]).reshape(B, T, -1))
        return torch.cat(codes, dim=2).transpose(1, 2).contiguous()


class MimiStreamingDecoder:
    """mimi.py:278-320: keeps the codec's streaming state across calls.  The reference decodes frame by frame with ``decode_step``; since
    every step returns the matching slice of a one-shot decode, a block of frames is decoded here with a single ``decode_step``."""

    def __init__(self, mimi: Mimi) -> None:
        self._mimi = mimi
        self.reset()

    def reset(self) -> None:
        self._mimi.reset_state()

    def decode_frames(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens [B, C, T] or [C, T] -> waveform [B, 1, 1920 T] for these frames (continuing the stream)."""
        if tokens.dim() == 2:
            tokens = tokens[None]
        return self._mimi.decode_step(tokens)
