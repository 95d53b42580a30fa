"""Whisper on B200 -- audio encoder path (reference: stt/models/whisper/whisper.py:280-448,501-530).

``Model(ModelDimensions)`` with ``load_weights``, ``encoder(mel)``, ``embed_audio``; the text decoder
loop is row "next-1" of SURVEY.md section 8f.  Fusions: conv+GELU, conv+GELU+positional embedding,
q|k|v as one GEMM (the key bias is structurally zero), residual adds in the GEMM epilogues.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

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


class Model:
    """whisper.py:501-530 (encoder side)."""

    def __init__(self, dims: ModelDimensions, dtype=torch.float16, device="cuda"):
        self.dims, self.dtype, self.device = dims, dtype, torch.device(device)
        self.encoder = AudioEncoder(dims, self.device)

    @property
    def sample_rate(self):
        return 16000

    def eval(self):
        return self

    def sanitize(self, weights):
        return {k: v for k, v in weights.items() if "_positional_embedding" not in k}

    def load_weights(self, weights, strict=False):
        self.encoder.load(dict(weights))
        return self

    def embed_audio(self, mel):
        return self.encoder(mel)

    def encode_audio(self, audio: torch.Tensor) -> torch.Tensor:
        """BASELINE config 3: audio [B, 480000] -> log-mel (reference's +30 s zero pad, first 3000 frames) -> encoder."""
        from .audio import N_FRAMES, N_SAMPLES
        mel = log_mel_spectrogram(audio, self.dims.n_mels, padding=N_SAMPLES, device=self.device)
        return self.encoder(mel[:, :N_FRAMES])

    def generate(self, audio, **kw):
        raise NotImplementedError("the Whisper decode loop is row next-1 of SURVEY.md section 8f; use embed_audio / encode_audio")
