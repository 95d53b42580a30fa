"""Synthetic (random-init) checkpoints at the reference's real config shapes.

There is no network and no weights on disk, so parity and the bench
run on random weights whose NAMES and SHAPES are the reference's MLX parameter tree
(post-``sanitize``): the same dict feeds the CPU oracle and the CUDA product.
Recipe: N(0, 0.02) everywhere, norm gains 1 / biases 0,
weight-norm ``g = ||v||``, Snake alpha 1, values rounded to bf16 (a "bf16 checkpoint"),
returned as float32 tensors holding bf16-exact values.
"""
from __future__ import annotations

import math

import torch


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _Gen:
    def __init__(self, seed, std=0.02):
        self.g = torch.Generator().manual_seed(seed)
        self.std = std
        self.P = {}

    def normal(self, name, *shape, std=None):
        self.P[name] = _bf16(torch.randn(*shape, generator=self.g) * (self.std if std is None else std))
        return self.P[name]

    def const(self, name, value, *shape):
        self.P[name] = torch.full(shape, float(value))

    def linear(self, pre, out_f, in_f, bias=True, std=None):
        self.normal(pre + ".weight", out_f, in_f, std=std)
        if bias:
            self.normal(pre + ".bias", out_f, std=std)

    def layer_norm(self, pre, c):
        self.const(pre + ".weight", 1.0, c)
        self.const(pre + ".bias", 0.0, c)

    def conv_weighted(self, pre, out_c, k, in_c, bias_c=None, bias=True):
        """ConvWeighted params (kokoro/istftnet.py:96-126): weight_v [out,K,in], weight_g [out,1,1] = ||v||."""
        v = self.normal(pre + ".weight_v", out_c, k, in_c)
        self.P[pre + ".weight_g"] = _bf16(torch.sqrt((v * v).sum(dim=(1, 2), keepdim=True)))
        if bias:
            self.normal(pre + ".bias", out_c if bias_c is None else bias_c)

    def lstm(self, pre, in_f, hid):
        for d in ("forward", "backward"):
            self.normal(f"{pre}.Wx_{d}", 4 * hid, in_f)
            self.normal(f"{pre}.Wh_{d}", 4 * hid, hid)
            self.normal(f"{pre}.bias_ih_{d}", 4 * hid)
            self.normal(f"{pre}.bias_hh_{d}", 4 * hid)

    def adain(self, pre, style, c):
        self.linear(pre + ".fc", 2 * c, style)

    def adain_resblk1d(self, pre, din, dout, style, upsample=False):
        self.conv_weighted(pre + ".conv1", dout, 3, din)
        self.conv_weighted(pre + ".conv2", dout, 3, dout)
        self.adain(pre + ".norm1", style, din)
        self.adain(pre + ".norm2", style, dout)
        if din != dout:
            self.conv_weighted(pre + ".conv1x1", dout, 1, din, bias=False)
        if upsample:
            self.conv_weighted(pre + ".pool", din, 3, 1)

    def adain_resblock1(self, pre, c, k, style):
        for j in range(3):
            self.conv_weighted(f"{pre}.convs1.{j}", c, k, c)
            self.conv_weighted(f"{pre}.convs2.{j}", c, k, c)
            self.adain(f"{pre}.adain1.{j}", style, c)
            self.adain(f"{pre}.adain2.{j}", style, c)
            self.const(f"{pre}.alpha1.{j}", 1.0, 1, c, 1)
            self.const(f"{pre}.alpha2.{j}", 1.0, 1, c, 1)


def kokoro_weights(cfg, seed=0):
    """Parameter tree of ``tts/models/kokoro/kokoro.py:Model`` (reference) with random values."""
    g = _Gen(seed)
    pb, hd, st = cfg["plbert"], cfg["hidden_dim"], cfg["style_dim"]
    e = pb.get("embedding_size", 128)
    hs, it = pb["hidden_size"], pb["intermediate_size"]
    # ALBERT (modules.py:434-645)
    g.normal("bert.embeddings.word_embeddings.weight", cfg["n_token"], e)
    g.normal("bert.embeddings.position_embeddings.weight", pb["max_position_embeddings"], e)
    g.normal("bert.embeddings.token_type_embeddings.weight", 2, e)
    g.layer_norm("bert.embeddings.LayerNorm", e)
    g.linear("bert.encoder.embedding_hidden_mapping_in", hs, e)
    L = "bert.encoder.albert_layer_groups.0.albert_layers.0."
    for n in ("query", "key", "value", "dense"):
        g.linear(L + "attention." + n, hs, hs)
    g.layer_norm(L + "attention.LayerNorm", hs)
    g.layer_norm(L + "full_layer_layer_norm", hs)
    g.linear(L + "ffn", it, hs)
    g.linear(L + "ffn_output", hs, it)
    g.linear("bert.pooler", hs, hs)
    g.linear("bert_encoder", hd, hs)
    # prosody predictor (modules.py:288-411)
    for i in range(cfg["n_layer"]):
        g.lstm(f"predictor.text_encoder.lstms.{2 * i}", hd + st, hd // 2)
        g.linear(f"predictor.text_encoder.lstms.{2 * i + 1}.fc", 2 * hd, st)
    g.lstm("predictor.lstm", hd + st, hd // 2)
    g.linear("predictor.duration_proj.linear_layer", cfg["max_dur"], hd)
    g.lstm("predictor.shared", hd + st, hd // 2)
    for name in ("F0", "N"):
        g.adain_resblk1d(f"predictor.{name}.0", hd, hd, st)
        g.adain_resblk1d(f"predictor.{name}.1", hd, hd // 2, st, upsample=True)
        g.adain_resblk1d(f"predictor.{name}.2", hd // 2, hd // 2, st)
        g.normal(f"predictor.{name}_proj.weight", 1, 1, hd // 2)
        g.normal(f"predictor.{name}_proj.bias", 1)
    # text encoder (modules.py:21-68)
    g.normal("text_encoder.embedding.weight", cfg["n_token"], hd)
    for i in range(cfg["n_layer"]):
        g.conv_weighted(f"text_encoder.cnn.{i}.0", hd, cfg["text_encoder_kernel_size"], hd)
        g.layer_norm(f"text_encoder.cnn.{i}.1", hd)
    g.lstm("text_encoder.lstm", hd, hd // 2)
    # decoder (istftnet.py:936-997)
    g.adain_resblk1d("decoder.encode", hd + 2, 1024, st)
    for i in range(3):
        g.adain_resblk1d(f"decoder.decode.{i}", 1024 + 2 + 64, 1024, st)
    g.adain_resblk1d("decoder.decode.3", 1024 + 2 + 64, 512, st, upsample=True)
    g.conv_weighted("decoder.F0_conv", 1, 3, 1)
    g.conv_weighted("decoder.N_conv", 1, 3, 1)
    g.conv_weighted("decoder.asr_res.0", 64, 1, 512)
    # generator (istftnet.py:725-835)
    ist = cfg["istftnet"]
    G = "decoder.generator"
    g.linear(G + ".m_source.l_linear", 1, 9, std=0.3)
    c0 = ist["upsample_initial_channel"]
    rates, ks = ist["upsample_rates"], ist["upsample_kernel_sizes"]
    nfft = ist["gen_istft_n_fft"]
    nk = len(ist["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(rates, ks)):
        cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
        g.conv_weighted(f"{G}.ups.{i}", cin, k, cout, bias_c=cout)      # ConvWeighted(cout, cin, .., encode=True)
        for j, rk in enumerate(ist["resblock_kernel_sizes"]):
            g.adain_resblock1(f"{G}.resblocks.{i * nk + j}", cout, rk, st)
        if i + 1 < len(rates):
            sf0 = math.prod(rates[i + 1:])
            g.normal(f"{G}.noise_convs.{i}.weight", cout, sf0 * 2, nfft + 2)
            g.normal(f"{G}.noise_convs.{i}.bias", cout)
            g.adain_resblock1(f"{G}.noise_res.{i}", cout, 7, st)
        else:
            g.normal(f"{G}.noise_convs.{i}.weight", cout, 1, nfft + 2)
            g.normal(f"{G}.noise_convs.{i}.bias", cout)
            g.adain_resblock1(f"{G}.noise_res.{i}", cout, 11, st)
    g.conv_weighted(G + ".conv_post", nfft + 2, 7, c0 // (2 ** len(rates)))
    return g.P


def kokoro_inputs(n_phonemes=128, n_token=178, seed=1):
    """cfg2 inputs: token ids [1, T] with BOS/EOS 0, style ref_s [1,256]."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, n_token, (n_phonemes,), generator=g)
    ids = torch.cat([torch.zeros(1, dtype=torch.long), ids, torch.zeros(1, dtype=torch.long)])[None]
    ref_s = torch.randn(1, 256, generator=torch.Generator().manual_seed(seed + 1))
    return ids, ref_s


def kokoro_noise(n_samples, seed=3):
    """Injected randomness for SineGen: rand_ini [1,9] U[0,1), noise [1,n,9] N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, 9, generator=g), torch.randn(1, n_samples, 9, generator=g)
