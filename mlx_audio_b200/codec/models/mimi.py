"""Kyutai Mimi codec, decode side, on B200 (reference: codec/models/mimi/mimi.py + modules/*).

``Mimi(mimi_202407(nq)).load_weights(...)``, ``decode(codes[B,nq,T]) -> [B,1,1920 T]`` (mimi.py:155-162).
RVQ gather-sum in one kernel, windowed causal attention (context 250) without materialising the
O(T^2) mask the reference builds (transformer.py:98-107), ELU / LayerScale / residual adds fused into
the neighbouring convs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

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

    @property
    def frame_rate(self):
        return self.cfg.frame_rate

    @property
    def sample_rate(self):
        return self.cfg.sample_rate

    def reset_state(self):
        """Full-sequence decode keeps no state; present for API parity (mimi.py:138-144)."""

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
        return self

    @torch.no_grad()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes int64 [B, nq, T] -> pcm [B, 1, 1920 T]."""
        W, cfg, dev = self._w, self.cfg, self.device
        codes = codes.to(device=dev, dtype=torch.int64).contiguous()
        B, nq, T = codes.shape
        q = ops.rvq_decode(codes[:, :1], W["cb_first"])
        x = ops.conv1d(q, W["proj_first"])
        if nq > 1:
            q2 = ops.rvq_decode(codes[:, 1:], W["cb_rest"][: nq - 1])
            x = ops.conv1d(q2, W["proj_rest"], res=x)
        s = cfg.upsample_stride
        x = ops.conv1d(x, W["upsample"], stride=s, pad_left=0, lout=T * s, transpose=True)         # causal: trim k-s on the right
        d, nh = cfg.dimension, cfg.num_heads
        for lw in W["layers"]:
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
        elu = Pre(act=ACT["elu"])
        x = ops.conv1d(x, W["init"], pad_left=cfg.ksize - 1, lout=x.shape[1])
        for lw in W["dec"]:
            r = lw["r"]
            y = ops.conv1d(x, lw["up"], stride=r, pad_left=0, lout=x.shape[1] * r, pre=elu, transpose=True)
            t = ops.conv1d(y, lw["c0"], pad_left=cfg.residual_ksize - 1, lout=y.shape[1], pre=elu)
            x = ops.conv1d(t, lw["c1"], pre=elu, res=y)
        pcm = ops.conv1d(x, W["final"], pad_left=cfg.last_ksize - 1, lout=x.shape[1], pre=elu)      # [B, L, 1]
        return pcm.reshape(B, 1, -1)

    def encode(self, xs):
        raise NotImplementedError("Mimi.encode is the 'next' row 2 of SURVEY.md section 8f (codec encode side)")
