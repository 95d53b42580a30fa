"""SNAC multi-scale RVQ codec, decode side, on B200 (reference: codec/models/snac/{snac,layers,vq}.py).

``SNAC(**config).load_weights(...)``, ``decode(codes) -> [B, T_out, 1]`` with the reference's
signature (snac.py:101-104).  Weight-norm is folded once at load; every Snake, bias, residual add
and NoiseBlock gain is a conv prologue/epilogue; the three codebook levels are gathered, projected,
repeat-interleaved and summed by one kernel (b2a_snac_from_codes).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from ... import ops
from ...ops import ACT, Pre


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fold_wn(v, g, except_dim=0):
    """snac/layers.py:9-14,57: g * v / ||v||, folded once at load in the checkpoint's dtype (SNAC checkpoints are
    float32 and the reference does not cast them, snac.py:184-199), so no rounding is applied."""
    v, g = v.double(), g.double()
    axes = tuple(i for i in range(v.dim()) if i != except_dim)
    return (g * v / torch.sqrt((v * v).sum(dim=axes, keepdim=True))).float()


class SNAC:
    def __init__(self, sampling_rate=44100, encoder_dim=64, encoder_rates=(3, 3, 7, 7), latent_dim=None, decoder_dim=1536,
                 decoder_rates=(7, 7, 3, 3), attn_window_size=32, codebook_size=4096, codebook_dim=8, vq_strides=(8, 4, 2, 1),
                 noise=True, depthwise=True, device="cuda"):
        self.sampling_rate = sampling_rate
        self.encoder_dim, self.encoder_rates = encoder_dim, list(encoder_rates)
        self.decoder_dim, self.decoder_rates = decoder_dim, list(decoder_rates)
        self.latent_dim = latent_dim or encoder_dim * (2 ** len(encoder_rates))
        self.hop_length = math.prod(encoder_rates)
        self.codebook_size, self.codebook_dim, self.vq_strides = codebook_size, codebook_dim, list(vq_strides)
        self.n_codebooks = len(vq_strides)
        self.attn_window_size, self.noise, self.depthwise = attn_window_size, noise, depthwise
        if attn_window_size is not None:
            raise NotImplementedError("SNAC LocalMHA (attn_window_size) is outside the accelerated configs (24 kHz model has none)")
        self.device = torch.device(device)
        self._w = None

    @classmethod
    def from_config(cls, config, device="cuda"):
        """snac.py:177-182: ``config`` = path of a config.json (as in the reference) or the already parsed dict."""
        if not isinstance(config, dict):
            import json
            with open(config, "r") as f:
                config = json.load(f)
        return cls(**config, device=device)

    @classmethod
    def from_pretrained(cls, repo_id, device="cuda", **kwargs):
        """snac.py:184-199 for a LOCAL snapshot directory (config.json + model.safetensors, already in the module-tree layout); a hub id
        is resolved through huggingface_hub only when that package can reach it."""
        from pathlib import Path
        path = Path(repo_id)
        if not path.exists():
            from huggingface_hub import snapshot_download
            path = Path(snapshot_download(repo_id=repo_id, allow_patterns=["*.safetensors", "*.json"]))
        from safetensors.torch import load_file
        model = cls.from_config(path / "config.json", device=device)
        return model.load_weights(list(load_file(str(path / "model.safetensors")).items()))

    @property
    def sample_rate(self):
        return self.sampling_rate

    def load_weights(self, weights, strict=True):
        P = dict(weights)
        dev = self.device
        f = lambda t: t.float().to(dev).contiguous()

        def wnconv(pre, groups=1):
            return ops.pack_conv(_fold_wn(P[pre + ".weight_v"], P[pre + ".weight_g"]), P.get(pre + ".bias"), groups, dev)

        def snake(name):
            a = P[name].float().reshape(-1)
            return f(a), f(1.0 / (a + 1e-9))                                   # x + sin(a x)^2 / (a + 1e-9), layers.py:124-130

        W = {"emb": [], "proj_w": [], "proj_b": []}
        for i in range(self.n_codebooks):
            q = f"quantizer.quantizers.{i}"
            W["emb"].append(f(P[q + ".codebook.weight"]))
            w = _fold_wn(P[q + ".out_proj.weight_v"], P[q + ".out_proj.weight_g"])      # [D,1,cd]
            W["proj_w"].append(f(w[:, 0, :].t()))                                          # [cd, D]
            W["proj_b"].append(f(P[q + ".out_proj.bias"]))
        pre = "decoder.model.layers"
        li = 0
        if self.depthwise:
            W["in_dw"] = wnconv(f"{pre}.{li}", groups=self.latent_dim); li += 1
            W["in_pw"] = wnconv(f"{pre}.{li}"); li += 1
        else:
            W["in_pw"] = wnconv(f"{pre}.{li}"); li += 1
        W["blocks"] = []
        for i, stride in enumerate(self.decoder_rates):
            cout = self.decoder_dim // (2 ** (i + 1))
            bp = f"{pre}.{li}.block.layers"; li += 1
            blk = {"stride": stride, "snake": snake(f"{bp}.0.alpha")}
            wt = _fold_wn(P[f"{bp}.1.weight_v"], P[f"{bp}.1.weight_g"], except_dim=0).permute(2, 1, 0)     # (in,K,out)->(out,K,in)
            blk["up"] = ops.pack_conv(wt, P.get(f"{bp}.1.bias"), 1, dev)
            bi = 2
            if self.noise:
                blk["noise"] = wnconv(f"{bp}.{bi}.linear"); bi += 1
            blk["res"] = []
            for d in (1, 3, 9):
                rp = f"{bp}.{bi}.block.layers"; bi += 1
                blk["res"].append({"d": d, "s1": snake(rp + ".0.alpha"), "c1": wnconv(rp + ".1", groups=cout if self.depthwise else 1),
                                   "s2": snake(rp + ".2.alpha"), "c2": wnconv(rp + ".3")})
            W["blocks"].append(blk)
        W["out_snake"] = snake(f"{pre}.{li}.alpha"); li += 1
        W["out_conv"] = wnconv(f"{pre}.{li}")
        self._w = W
        self._enc = None
        if "encoder.block.layers.0.weight_v" in P:                     # encode side (snac/layers.py:133-158, vq.py:22-109)
            E = {"in": wnconv("encoder.block.layers.0"), "blocks": [], "q": []}
            pre, li, d = "encoder.block.layers", 1, self.encoder_dim
            for stride in self.encoder_rates:
                bp = f"{pre}.{li}.block.layers"; li += 1
                blk = {"stride": stride, "res": [], "snake": snake(f"{bp}.3.alpha"), "down": wnconv(f"{bp}.4")}
                for bi, dil in enumerate((1, 3, 9)):
                    rp = f"{bp}.{bi}.block.layers"
                    blk["res"].append({"d": dil, "s1": snake(rp + ".0.alpha"), "c1": wnconv(rp + ".1", groups=d if self.depthwise else 1),
                                       "s2": snake(rp + ".2.alpha"), "c2": wnconv(rp + ".3")})
                E["blocks"].append(blk)
                d *= 2
            E["out"] = wnconv(f"{pre}.{li}", groups=d if self.depthwise else 1)
            for i in range(self.n_codebooks):
                q = f"quantizer.quantizers.{i}"
                cb = P[q + ".codebook.weight"].double()
                cn = (cb / cb.norm(dim=1, keepdim=True).clamp(min=1e-12)).float()            # the cosine search runs on the L2-normalised table (vq.py:62-66)
                E["q"].append({"in_proj": wnconv(q + ".in_proj"), "cn": f(cn)[None].contiguous(),
                               "c2": (cn.double() ** 2).sum(1)[None].to(dev).contiguous()})
            self._enc = E
        return self

    @torch.no_grad()
    def decode(self, codes: List[torch.Tensor], noises: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """snac.py:101-104.  codes[l] int64 [B, T/stride_l] -> audio [B, T_out, 1].
        ``noises[i]`` [B,1,C_i] injects the NoiseBlock draw (one per channel, layers.py:261-267); None draws it."""
        W, dev = self._w, self.device
        codes = [c.to(device=dev, dtype=torch.int64).contiguous() for c in codes]
        z = ops.snac_from_codes(codes, self.vq_strides, W["emb"], W["proj_w"], W["proj_b"], self.latent_dim)     # [B,T,latent]
        B = z.shape[0]
        x = ops.conv1d(z, W["in_dw"], pad_left=3) if self.depthwise else z
        x = ops.conv1d(x, W["in_pw"], pad_left=0 if self.depthwise else 3)
        for i, blk in enumerate(W["blocks"]):
            s = blk["stride"]
            L = x.shape[1]
            p = math.ceil(s / 2)
            lout = (L - 1) * s - 2 * p + (2 * s - 1) + 1 + 1          # output_padding = 1 (the reference's positional-arg quirk)
            a, ia = blk["snake"]
            y = ops.conv1d(x, blk["up"], stride=s, pad_left=p, lout=lout, pre=Pre(act=ACT["snake"], a=a, b=ia), transpose=True)
            if self.noise:
                nz = noises[i] if noises is not None else torch.randn(B, 1, y.shape[2], device=dev)
                nz = nz.to(device=dev, dtype=torch.float32).reshape(B, -1).contiguous()
                y = ops.conv1d(y, blk["noise"], cscale=nz, res=y)                   # x + noise * linear(x)
            for ru in blk["res"]:
                s1 = Pre(act=ACT["snake"], a=ru["s1"][0], b=ru["s1"][1])
                s2 = Pre(act=ACT["snake"], a=ru["s2"][0], b=ru["s2"][1])
                if ru["c1"].groups > 1 and ru["c2"].w_tc is not None and ops.emit_eligible(ru["c1"], y, y.shape[1], dilation=ru["d"]) \
                        and ru["c2"].cin_pad == ru["c1"].cout and ru["c2"].cin * ru["c2"].K >= ops.TC_MIN_K:
                    # depthwise conv writes Snake2(t) as the 1x1 conv's bf16 planes: no fp32 t, no prologue pass
                    t = ops.conv1d(y, ru["c1"], dilation=ru["d"], pad_left=3 * ru["d"], pre=s1, emit=s2)
                    y = ops.conv1d(t, ru["c2"], res=y)
                else:
                    t = ops.conv1d(y, ru["c1"], dilation=ru["d"], pad_left=3 * ru["d"], pre=s1)
                    y = ops.conv1d(t, ru["c2"], pre=s2, res=y)
            x = y
        a, ia = W["out_snake"]
        return ops.conv1d(x, W["out_conv"], pad_left=3, pre=Pre(act=ACT["snake"], a=a, b=ia), post_act=ACT["tanh"])

    # halo (in finest-level frames) that makes a span decode EXACT: first conv +-3, per block the transposed conv (+-1) and three residual
    # units of kernel 7 with dilations 1 / 3 / 9 (+-39 samples at that block's rate: 39/8 + 39/64 + ... < 6 frames), last conv +-3/512.
    SPAN_HALO = 16

    @torch.no_grad()
    def decode_span(self, codes: List[torch.Tensor], start: int, end: int, noises: Optional[List[torch.Tensor]] = None, halo: int = SPAN_HALO):
        """Audio of the finest-level frames [start, end) of a stream, decoded from those frames plus ``halo`` frames on each side: the decoder
        is convolutional, so this equals the corresponding slice of ``decode(codes)`` (one stream sharded across
        GPUs; the per-channel NoiseBlock draws ``noises`` must be the same on every shard).  ``start`` / ``end`` / ``halo`` are multiples of
        the coarsest code stride.  Returns [B, (end - start) * hop (+ the decoder's 75-sample tail when ``end`` is the stream's end), 1]."""
        top = max(self.vq_strides)
        T = codes[-1].shape[1] * self.vq_strides[-1]
        if start % top or (end % top and end != T) or halo % top or not 0 <= start < end <= T:
            raise ValueError(f"decode_span: start / end / halo must be multiples of {top} inside [0, {T}]")
        rs, re = max(0, start - halo), min(T, end + halo)
        part = [c[:, rs // st: -(-re // st)] for c, st in zip(codes, self.vq_strides)]
        y = self.decode(part, noises=noises)
        hop = math.prod(self.decoder_rates)
        lo = (start - rs) * hop
        hi = y.shape[1] if end == T else lo + (end - start) * hop
        return y[:, lo:hi]

    def decode_stream(self, codes: List[torch.Tensor], prev_codes: Optional[List[torch.Tensor]] = None, context_frames: int = 8, noises=None):
        """snac.py:106-162, literally: the first call decodes ``codes``; later calls prepend ``max(1, context_frames // stride_l)`` frames of
        context per level and decode the combination.  Kept quirk: the reference trims the context with ``full_audio[..., n:]`` on an audio
        tensor whose LAST axis is the channel ([B, T, 1]), so nothing is trimmed and the context's audio is returned again.
        -> (audio [B, T, 1], new context)."""
        new_context = [c[:, -context_frames:] if c.shape[1] > context_frames else c for c in codes]
        if prev_codes is None:
            return self.decode(codes, noises=noises), new_context
        combined = []
        for stride, prev, new in zip(self.vq_strides, prev_codes, codes):
            keep = max(1, context_frames // stride)
            prev = prev[:, -keep:] if prev.shape[1] > keep else prev
            combined.append(torch.cat([prev.to(new.device), new], dim=1))
        full_audio = self.decode(combined, noises=noises)
        context_samples = context_frames * self.hop_length
        audio = full_audio[..., context_samples:] if full_audio.shape[-1] > context_samples else full_audio
        return audio, new_context

    def preprocess(self, audio_data: torch.Tensor) -> torch.Tensor:
        """snac.py:67-84: right-pad [B, 1, n] to a multiple of hop x lcm(vq_strides)."""
        lcm = 1
        for v in self.vq_strides:
            lcm = lcm * v // math.gcd(lcm, v)
        pad_to = self.hop_length * lcm
        n = audio_data.shape[-1]
        return torch.nn.functional.pad(audio_data, (0, -n % pad_to))

    @torch.no_grad()
    def encode(self, audio_data: torch.Tensor) -> List[torch.Tensor]:
        """snac.py:95-99: audio [B, 1, n] -> codes, coarse to fine: [B, T/4], [B, T/2], [B, T] for the 24 kHz model.

        Encoder = the decoder's building blocks mirrored (k7 conv; per rate three Snake -> depthwise k7 (dilation 1 / 3 / 9) -> Snake -> 1x1
        residual units, Snake, strided conv k = 2s; depthwise k7 out conv), then per level of the residual quantiser: average-pool by the
        level's stride, 1x1 projection to 8 dims, cosine nearest-code search (`rvq_encode_kernel`, first index on ties), project the code
        back, repeat by the stride, subtract from the residual (vq.py:22-109).  Needs a checkpoint that carries the encoder."""
        residual = self.encode_latent(audio_data)                                                              # [B, T, latent]
        E = self._enc
        B, T, D = residual.shape
        codes = []
        for i, stride in enumerate(self.vq_strides):
            q = E["q"][i]
            xl = residual
            if stride > 1:
                t_ = (T - stride) // stride + 1
                xl = (residual[:, : t_ * stride].reshape(B, t_, stride, D).sum(dim=2) / stride).contiguous()
            ze = ops.conv1d(xl, q["in_proj"])                                                                  # [B, T', cd]
            idx = ops.rvq_encode(ze.reshape(-1, ze.shape[-1]), q["cn"], q["c2"], mode=1)[:, 0].reshape(B, -1).contiguous()
            codes.append(idx)
            if i + 1 < len(self.vq_strides):
                zq = ops.snac_from_codes([idx], [stride], [self._w["emb"][i]], [self._w["proj_w"][i]], [self._w["proj_b"][i]], D, check=False)
                residual = residual - zq[:, :T] if zq.shape[1] >= T else residual - torch.nn.functional.pad(zq, (0, 0, 0, T - zq.shape[1]))
        return codes

    @torch.no_grad()
    def encode_latent(self, audio_data: torch.Tensor) -> torch.Tensor:
        """The encoder alone (snac/layers.py:133-158): audio [B, 1, n] -> z [B, T, latent] in front of the quantiser."""
        if self._enc is None:
            raise ValueError("SNAC.encode: the loaded weights have no encoder (encoder.block.layers.*)")
        E, dev = self._enc, self.device
        x = self.preprocess(audio_data.to(device=dev, dtype=torch.float32))
        x = x.reshape(x.shape[0], -1, 1)                                                                       # [B, 1, n] -> [B, n, 1] (one channel: same memory)
        y = ops.conv1d(x, E["in"], pad_left=3)
        for blk in E["blocks"]:
            for ru in blk["res"]:
                s1 = Pre(act=ACT["snake"], a=ru["s1"][0], b=ru["s1"][1])
                s2 = Pre(act=ACT["snake"], a=ru["s2"][0], b=ru["s2"][1])
                t = ops.conv1d(y, ru["c1"], dilation=ru["d"], pad_left=3 * ru["d"], pre=s1)
                y = ops.conv1d(t, ru["c2"], pre=s2, res=y)
            s = blk["stride"]
            y = ops.conv1d(y, blk["down"], stride=s, pad_left=math.ceil(s / 2), pre=Pre(act=ACT["snake"], a=blk["snake"][0], b=blk["snake"][1]))
        return ops.conv1d(y, E["out"], pad_left=3)
