"""Oracle for the Qwen3-TTS hot path: talker, code predictor, sampler, frame loop and the
12.5 Hz speech-tokenizer decoder, restated on torch-CPU tensors (float64 for checking).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs,
never by the product.  Each function cites the reference lines it follows.  Weights ``P`` use the reference's MLX-side
parameter paths (after ``sanitize``): conv weights ``[Cout, K, Cin/g]``, linear weights ``[out, in]``.

Deviation that is part of the contract: ``mx.random.categorical`` is not reproducible outside MLX, so the categorical
draw is the inverse CDF in index order driven by a caller-supplied uniform ``u`` -- the same definition the CUDA sampler
(b2a_sample_token) uses.  Everything before the draw (suppress, repetition penalty, temperature, top-k, top-p, min-p)
follows the reference literally.

parity: pinned against the reference's own shape/length contracts (tests/test_oracle_pins.py: 1920 samples per frame,
chunked == unchunked decode on the overlap-free region, interleaved MRoPE index pattern of talker.py:139-184, sampler
filters vs hand-computed cases of lm/sample_utils.py) AND against the reference's own talker / code predictor / sampler / Model.generate /
Model.batch_generate / speech-tokenizer code executed through a NumPy stand-in for MLX (tests/golden/make_qwen3_golden.py, qwen3_golden.npz:
logits 4e-15, code matrices identical).
"""
from __future__ import annotations

import math

import torch

from . import nn as N

TALKER = {   # qwen3_tts/config.py:57-102 (Qwen3TTSTalkerConfig) and :32-54 (code predictor)
    "vocab_size": 3072, "hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 28, "num_attention_heads": 16,
    "num_key_value_heads": 8, "head_dim": 128, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0, "mrope_section": [24, 20, 20],
    "num_code_groups": 16, "codec_eos_token_id": 2150, "text_hidden_size": 2048,
    "cp_vocab_size": 2048, "cp_hidden_size": 1024, "cp_intermediate_size": 3072, "cp_num_hidden_layers": 5,
    "cp_num_attention_heads": 16, "cp_num_key_value_heads": 8, "cp_head_dim": 128, "cp_rope_theta": 1000000.0,
}

TOKENIZER_DECODER = {   # qwen3_tts/config.py:105-133 (Qwen3TTSTokenizerDecoderConfig)
    "latent_dim": 1024, "codebook_dim": 512, "codebook_size": 2048, "decoder_dim": 1536, "hidden_size": 512,
    "intermediate_size": 1024, "layer_scale_initial_scale": 0.01, "head_dim": 64, "num_attention_heads": 16,
    "num_hidden_layers": 8, "num_key_value_heads": 16, "num_quantizers": 16, "num_semantic_quantizers": 1,
    "rms_norm_eps": 1e-5, "rope_theta": 10000.0, "upsample_rates": [8, 5, 4, 3], "upsampling_ratios": [2, 2],
}


# ------------------------------------------------------------------------------------------------ rotary embeddings
def rotate_half(x):
    """talker.py:14-18 / speech_tokenizer.py:214-217."""
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def rope_cos_sin(position_ids, dim, base, dtype=torch.float64):
    """RotaryEmbedding.__call__ (talker.py:87-113) / DecoderRotaryEmbedding (speech_tokenizer.py:202-211).
    position_ids [B,S] -> cos, sin [B,S,dim]."""
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=dtype) / dim))
    freqs = position_ids.to(dtype)[:, :, None] * inv[None, None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return torch.cos(emb), torch.sin(emb)


def mrope_cos_sin(position_ids, dim, base, section, dtype=torch.float64):
    """TalkerRotaryEmbedding.__call__ + apply_interleaved_mrope (talker.py:139-226).  position_ids [3,B,S] (or [B,S],
    broadcast to the three axes) -> cos, sin [B,S,dim]; frequency slot i takes the H position when i%3==1 and
    i < 3*section[1], the W position when i%3==2 and i < 3*section[2], else the T position."""
    if position_ids.dim() == 2:
        position_ids = position_ids[None].expand(3, -1, -1)
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=dtype) / dim))
    freqs = position_ids.to(dtype)[..., None] * inv                      # [3,B,S,dim/2]
    idx = torch.arange(dim // 2)
    h_mask = (idx % 3 == 1) & (idx < section[1] * 3)
    w_mask = (idx % 3 == 2) & (idx < section[2] * 3)
    comb = torch.where(h_mask, freqs[1], freqs[0])
    comb = torch.where(w_mask, freqs[2], comb)
    emb = torch.cat([comb, comb], dim=-1)
    return torch.cos(emb), torch.sin(emb)


def apply_rope(q, k, cos, sin):
    """apply_rotary_pos_emb / apply_multimodal_rotary_pos_emb (talker.py:22-65): q,k [B,H,S,D]; cos,sin [B,S,D]."""
    c, s = cos[:, None], sin[:, None]
    return q * c + rotate_half(q) * s, k * c + rotate_half(k) * s


def causal_mask(s_q, s_k, dtype, kv_start=None):
    """Additive mask for the last s_q of s_k positions (create_additive_causal_mask, rows = query positions)."""
    i = torch.arange(s_k - s_q, s_k)[:, None]
    j = torch.arange(s_k)[None, :]
    m = torch.where(j <= i, 0.0, -1e9).to(dtype)
    return m


# ------------------------------------------------------------------------------------------------ talker / code predictor
def _attn_block(P, L, x, cos, sin, cache, n_heads, n_kv, hd, eps, mask):
    """TalkerAttention / CodePredictorAttention.__call__ (talker.py:270-316, 542-581): q/k per-head RMSNorm, rotary,
    KV concat cache, GQA SDPA, o_proj."""
    b, s, _ = x.shape
    q = N.linear(x, P[L + ".q_proj.weight"]).reshape(b, s, n_heads, hd)
    k = N.linear(x, P[L + ".k_proj.weight"]).reshape(b, s, n_kv, hd)
    v = N.linear(x, P[L + ".v_proj.weight"]).reshape(b, s, n_kv, hd)
    q = N.rms_norm(q, P[L + ".q_norm.weight"], eps).transpose(1, 2)
    k = N.rms_norm(k, P[L + ".k_norm.weight"], eps).transpose(1, 2)
    v = v.transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    if cache is not None:
        if "k" in cache:
            k = torch.cat([cache["k"], k], dim=2)
            v = torch.cat([cache["v"], v], dim=2)
        cache["k"], cache["v"] = k, v
    o = N.sdpa(q, k, v, hd ** -0.5, mask)
    return N.linear(o.transpose(1, 2).reshape(b, s, n_heads * hd), P[L + ".o_proj.weight"])


def _mlp(P, L, x):
    """TalkerMLP / CodePredictorMLP (talker.py:319-336,584-600): down(silu(gate(x)) * up(x))."""
    g = N.linear(x, P[L + ".gate_proj.weight"])
    return N.linear(torch.nn.functional.silu(g) * N.linear(x, P[L + ".up_proj.weight"]), P[L + ".down_proj.weight"])


def _decoder_stack(P, pre, x, cos, sin, caches, n_layers, n_heads, n_kv, hd, eps, mask):
    """TalkerDecoderLayer / CodePredictorDecoderLayer loop + final norm (talker.py:381-400,478-495,615-632,688-698)."""
    for i in range(n_layers):
        L = f"{pre}.layers.{i}"
        c = None if caches is None else caches[i]
        h = N.rms_norm(x, P[L + ".input_layernorm.weight"], eps)
        x = x + _attn_block(P, L + ".self_attn", h, cos, sin, c, n_heads, n_kv, hd, eps, mask)
        h = N.rms_norm(x, P[L + ".post_attention_layernorm.weight"], eps)
        x = x + _mlp(P, L + ".mlp", h)
    return N.rms_norm(x, P[pre + ".norm.weight"], eps)


def make_cache(n_layers):
    return [dict() for _ in range(n_layers)]


def cache_offset(caches):
    return 0 if caches is None or "k" not in caches[0] else caches[0]["k"].shape[2]


def talker_forward(P, inputs_embeds, caches=None, position_ids=None, cfg=TALKER, attention_mask=None):
    """Qwen3TTSTalkerForConditionalGeneration.__call__ (talker.py:799-818) over Qwen3TTSTalkerModel.__call__
    (talker.py:435-496): inputs_embeds [B,S,1024] -> (logits [B,S,3072], hidden [B,S,1024]).  ``attention_mask`` [B, total_kv] (1 = real
    token) is the left-padded batch path (:449-476): positions cumsum(mask)-1 clamped at 0, additive -1e9 on padded keys."""
    b, s, _ = inputs_embeds.shape
    off = cache_offset(caches)
    dt = inputs_embeds.dtype
    if position_ids is None:
        if attention_mask is not None:
            pos = torch.clamp(torch.cumsum(attention_mask.to(torch.int64), dim=-1) - 1, min=0)[:, -s:]
            position_ids = torch.stack([pos, pos, pos], dim=0)
        else:
            position_ids = torch.arange(off, off + s)[None, :].expand(b, s)
    cos, sin = mrope_cos_sin(position_ids, cfg["head_dim"], cfg["rope_theta"], cfg["mrope_section"], dt)
    if attention_mask is not None:
        pad = (1 - attention_mask[:, None, None, :].to(dt)) * -1e9            # [B,1,1,total]
        mask = causal_mask(s, s, dt)[None, None] + pad if s > 1 else pad
    else:
        mask = causal_mask(s, s, dt) if s > 1 else None      # reference builds [S,S] (prefill starts at offset 0)
        if mask is not None and off > 0:
            mask = torch.cat([torch.zeros(s, off, dtype=mask.dtype), mask], dim=1)
    h = _decoder_stack(P, "model", inputs_embeds, cos, sin, caches, cfg["num_hidden_layers"], cfg["num_attention_heads"],
                       cfg["num_key_value_heads"], cfg["head_dim"], cfg["rms_norm_eps"], mask)
    return N.linear(h, P["codec_head.weight"]), h


def code_predictor_forward(P, inputs_embeds, caches, generation_step, cfg=TALKER):
    """Qwen3TTSTalkerCodePredictor.__call__ (talker.py:742-760) over CodePredictorModel.__call__ (:667-699)."""
    if "code_predictor.small_to_mtp_projection.weight" in P:
        inputs_embeds = N.linear(inputs_embeds, P["code_predictor.small_to_mtp_projection.weight"],
                                 P["code_predictor.small_to_mtp_projection.bias"])
    b, s, _ = inputs_embeds.shape
    off = cache_offset(caches)
    pos = torch.arange(off, off + s)[None, :].expand(b, s)
    cos, sin = rope_cos_sin(pos, cfg["cp_head_dim"], cfg["cp_rope_theta"], inputs_embeds.dtype)
    mask = causal_mask(s, s, inputs_embeds.dtype) if s > 1 else None
    h = _decoder_stack(P, "code_predictor.model", inputs_embeds, cos, sin, caches, cfg["cp_num_hidden_layers"],
                       cfg["cp_num_attention_heads"], cfg["cp_num_key_value_heads"], cfg["cp_head_dim"], cfg["rms_norm_eps"], mask)
    return N.linear(h, P[f"code_predictor.lm_head.{generation_step}.weight"])


# ------------------------------------------------------------------------------------------------ sampler
def apply_top_k(logprobs, top_k):
    """lm/sample_utils.py:131-151: everything outside the k largest -> -inf (ties: lower index wins, the CUDA sampler's rule)."""
    v = logprobs.shape[-1]
    order = sorted(range(v), key=lambda i: (-float(logprobs[i]), i))
    out = torch.full_like(logprobs, -math.inf)
    keep = torch.tensor(order[:top_k])
    out[keep] = logprobs[keep]
    return out


def apply_top_p(logprobs, top_p):
    """lm/sample_utils.py:206-239: ascending sort, cumulative probability, keep where cum > 1 - top_p."""
    probs = torch.exp(logprobs)
    order = torch.tensor(sorted(range(logprobs.shape[-1]), key=lambda i: (float(logprobs[i]), -i)))   # ascending; reverse of the top-k order
    cum = torch.cumsum(probs[order], dim=-1)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel())
    cum = cum[inv]
    return torch.where(cum > 1 - top_p, logprobs, torch.full_like(logprobs, -math.inf))


def apply_min_p(logprobs, min_p):
    """lm/sample_utils.py:156-203 (min_tokens_to_keep = 1)."""
    return torch.where(logprobs < logprobs.max() + math.log(min_p), torch.full_like(logprobs, -math.inf), logprobs)


def sample_token(logits, u, temperature=0.9, top_k=50, top_p=1.0, repetition_penalty=1.05, generated_tokens=None,
                 suppress_tokens=None, min_p=0.0, return_filtered=False):
    """Model._sample_token (qwen3_tts.py:805-860) for one row ``logits`` [V]; ``u`` in [0,1) drives the draw."""
    logits = logits.clone()
    if suppress_tokens:
        logits[torch.tensor(list(suppress_tokens))] = -math.inf
    if generated_tokens and repetition_penalty != 1.0:
        ids = [t for t in set(generated_tokens) if t < logits.shape[-1]]
        if ids:
            ids = torch.tensor(ids)
            sel = logits[ids]
            logits[ids] = torch.where(sel < 0, sel * repetition_penalty, sel / repetition_penalty)
    if temperature <= 0:
        tok = int(torch.argmax(logits))
        return (tok, logits) if return_filtered else tok
    if temperature != 1.0:
        logits = logits / temperature
    if 0 < top_k < logits.shape[-1]:
        logits = apply_top_k(logits, top_k)
    if 0.0 < top_p < 1.0 or min_p > 0.0:                                  # _apply_probability_filters (qwen3_tts.py:47-61)
        lp = torch.log_softmax(logits, dim=-1)
        if 0.0 < top_p < 1.0:
            lp = apply_top_p(lp, top_p)
        if min_p > 0.0:
            lp = apply_min_p(lp, min_p)
        logits = torch.where(torch.isinf(lp) & (lp < 0), torch.full_like(logits, -math.inf), logits)
    w = torch.exp((logits - logits.max()).double())
    cum = torch.cumsum(w, dim=-1)
    target = float(u) * float(cum[-1])
    live = torch.nonzero(w > 0).flatten()
    tok = int(live[-1])
    for i in live.tolist():
        if float(cum[i]) > target:
            tok = i
            break
    return (tok, logits) if return_filtered else tok


# ------------------------------------------------------------------------------------------------ frame loop
def generate_codes(P, input_embeds, trailing_text_hidden, tts_pad_embed, u, max_tokens, temperature=0.9, top_k=50, top_p=1.0,
                   repetition_penalty=1.05, cfg=TALKER, trace=None):
    """The per-frame loop of Model.generate (qwen3_tts.py:1300-1404), batch 1, after _prepare_generation_inputs:
    talker step -> first-codebook sample (suppress + repetition penalty) -> 15 code-predictor sub-steps with a fresh cache ->
    next input = trailing text (or pad) embedding + sum of the 16 code embeddings.  ``u`` [max_tokens, 16] uniforms.
    Returns int64 codes [n_frames, 16] (the EOS frame is not emitted)."""
    g = cfg["num_code_groups"]
    eos = cfg["codec_eos_token_id"]
    suppress = [i for i in range(cfg["vocab_size"] - 1024, cfg["vocab_size"]) if i != eos]
    cache = make_cache(cfg["num_hidden_layers"])
    generated, out = [], []
    trailing_idx = 0
    x = input_embeds
    emb0 = P["model.codec_embedding.weight"]
    for step in range(max_tokens):
        logits, hidden = talker_forward(P, x, cache, cfg=cfg)
        tok = sample_token(logits[0, -1], u[step, 0], temperature, top_k, top_p, repetition_penalty, generated or None, suppress)
        if trace is not None:
            trace.append({"logits": logits[0, -1].clone(), "hidden": hidden[0, -1].clone(), "token": tok})
        codes = [tok]
        code_cache = make_cache(cfg["cp_num_hidden_layers"])
        code_hidden = hidden[:, -1:, :]
        for ci in range(g - 1):
            if ci == 0:
                inp = torch.cat([code_hidden, emb0[torch.tensor([[tok]])]], dim=1)
            else:
                inp = P[f"code_predictor.model.codec_embedding.{ci - 1}.weight"][torch.tensor([[codes[-1]]])]
            cl = code_predictor_forward(P, inp, code_cache, ci, cfg)
            codes.append(sample_token(cl[0, -1], u[step, ci + 1], temperature, top_k, top_p, 1.05, None, None))
        if trailing_idx < trailing_text_hidden.shape[1]:
            text = trailing_text_hidden[:, trailing_idx: trailing_idx + 1]
            trailing_idx += 1
        else:
            text = tts_pad_embed
        ce = emb0[torch.tensor([[codes[0]]])]
        for i, c in enumerate(codes[1:]):
            ce = ce + P[f"code_predictor.model.codec_embedding.{i}.weight"][torch.tensor([[c]])]
        x = text + ce
        if tok == eos:
            break
        generated.append(tok)
        out.append(codes)
    return torch.tensor(out, dtype=torch.int64).reshape(-1, g)


def generate_codes_batch(P, embeds_list, trailing_list, tts_pad_embed, u, max_tokens, temperature=0.9, top_k=50, top_p=1.0,
                         repetition_penalty=1.05, cfg=TALKER):
    """Model.batch_generate's loop (qwen3_tts.py:1861-1935) over _prepare_batch_inputs' padding (:486-604): prompts left-padded with
    zero rows + attention mask, trailing text right-padded with the pad embedding, finished rows forced to EOS, trailing indices
    advanced for unfinished rows only, next input through _next_batch_input_embeds(pad_when_index_clamped=True) (:993-1015).
    ``u`` [max_tokens, 16, B].  Returns a list of int64 code matrices [n_b, 16]."""
    B = len(embeds_list)
    g, eos = cfg["num_code_groups"], cfg["codec_eos_token_id"]
    H = embeds_list[0].shape[-1]
    dt = embeds_list[0].dtype
    pmax = max(e.shape[1] for e in embeds_list)
    tmax = max(t.shape[1] for t in trailing_list)
    x = torch.zeros(B, pmax, H, dtype=dt)
    mask = torch.zeros(B, pmax, dtype=dt)
    trailing = tts_pad_embed.reshape(1, 1, H).expand(B, tmax, H).clone()
    for i, (e, t) in enumerate(zip(embeds_list, trailing_list)):
        x[i, pmax - e.shape[1]:] = e[0]
        mask[i, pmax - e.shape[1]:] = 1
        trailing[i, : t.shape[1]] = t[0]
    suppress = [i for i in range(cfg["vocab_size"] - 1024, cfg["vocab_size"]) if i != eos]
    cache = make_cache(cfg["num_hidden_layers"])
    gen_ids = [[] for _ in range(B)]
    out = [[] for _ in range(B)]
    finished = [False] * B
    tidx = [0] * B
    emb0 = P["model.codec_embedding.weight"]
    amask = mask if B > 1 else None                                       # bs = 1 drops the mask (:1826-1829)
    for step in range(max_tokens):
        logits, hidden = talker_forward(P, x, cache, cfg=cfg, attention_mask=amask)
        toks = []
        for b in range(B):
            tk = sample_token(logits[b, -1], u[step, 0, b], temperature, top_k, top_p, repetition_penalty, gen_ids[b] or None, suppress)
            toks.append(eos if finished[b] else tk)
        finished = [f or tk == eos for f, tk in zip(finished, toks)]
        codes = [[tk] for tk in toks]
        code_cache = make_cache(cfg["cp_num_hidden_layers"])
        first = torch.tensor(toks)[:, None]
        for ci in range(g - 1):
            if ci == 0:
                inp = torch.cat([hidden[:, -1:, :], emb0[first]], dim=1)
            else:
                inp = P[f"code_predictor.model.codec_embedding.{ci - 1}.weight"][torch.tensor([c[-1] for c in codes])[:, None]]
            cl = code_predictor_forward(P, inp, code_cache, ci, cfg)
            for b in range(B):
                codes[b].append(sample_token(cl[b, -1], u[step, ci + 1, b], temperature, top_k, top_p, 1.05, None, None))
        nxt = []
        for b in range(B):
            cl_i = min(tidx[b], tmax - 1)
            text = tts_pad_embed.reshape(1, H) if cl_i >= tmax - 1 else trailing[b, cl_i: cl_i + 1]
            ce = emb0[codes[b][0]][None]
            for i, c in enumerate(codes[b][1:]):
                ce = ce + P[f"code_predictor.model.codec_embedding.{i}.weight"][c][None]
            nxt.append(text + ce)
            if not finished[b]:
                tidx[b] += 1
        x = torch.stack(nxt, dim=0)
        if all(finished):
            break
        for b in range(B):
            if not finished[b]:
                gen_ids[b].append(toks[b])
                out[b].append(codes[b])
        if amask is not None:
            amask = torch.cat([amask, torch.ones(B, 1, dtype=dt)], dim=1)
    return [torch.tensor(o, dtype=torch.int64).reshape(-1, g) for o in out]


# ------------------------------------------------------------------------------------------------ speech-tokenizer decoder
def snake_beta(x, alpha, beta):
    """SnakeBeta (speech_tokenizer.py:110-126): x + 1/(exp(beta)+1e-9) * sin(x exp(alpha))^2, channels-last."""
    a, b = torch.exp(alpha.to(x.dtype)), torch.exp(beta.to(x.dtype))
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def causal_conv(P, pre, x, k, dilation=1, groups=1):
    """CausalConv1d.__call__ (speech_tokenizer.py:65-69), stride 1: left zero pad (k-1)*dilation; x [B,T,C]."""
    pad = (k - 1) * dilation
    xp = torch.nn.functional.pad(x, (0, 0, pad, 0))
    return N.conv1d(xp, P[pre + ".weight"].to(x.dtype), 1, 0, dilation, groups, P.get(pre + ".bias"))


def causal_convtr(P, pre, x, k, stride):
    """CausalTransposeConv1d / DecoderBlockUpsample.__call__ (speech_tokenizer.py:102-107,638-643): full scatter, drop k-stride on the right."""
    y = N.conv_transpose1d(x, P[pre + ".weight"].to(x.dtype), stride, 0, 1, 0, 1, P.get(pre + ".bias"))
    trim = k - stride
    return y[:, : y.shape[1] - trim] if trim > 0 else y


def quantizer_decode(P, codes, cfg=TOKENIZER_DECODER):
    """SplitResidualVectorQuantizer.decode (speech_tokenizer.py:577-582,532-541,483-490): codes [B,16,T] -> [B,T,512]."""
    nsem = cfg["num_semantic_quantizers"]
    out = None
    for name, qs in (("rvq_first", range(0, nsem)), ("rvq_rest", range(nsem, codes.shape[1]))):
        q = None
        for li, qi in enumerate(qs):
            e = P[f"decoder.quantizer.{name}.vq.layers.{li}.codebook.embed.weight"][codes[:, qi]]     # [B,T,256]
            q = e if q is None else q + e
        if q is None:
            continue
        y = N.conv1d(q, P[f"decoder.quantizer.{name}.output_proj.weight"].to(q.dtype))
        out = y if out is None else out + y
    return out


def decoder_transformer(P, x, cfg=TOKENIZER_DECODER):
    """DecoderTransformer.__call__ (speech_tokenizer.py:383-413), no cache: x [B,T,1024] -> [B,T,1024]."""
    pre = "decoder.pre_transformer"
    b, t, _ = x.shape
    nh, hd, eps = cfg["num_attention_heads"], cfg["head_dim"], cfg["rms_norm_eps"]
    x = N.linear(x, P[pre + ".input_proj.weight"], P[pre + ".input_proj.bias"])
    pos = torch.arange(t)[None, :].expand(b, t)
    cos, sin = rope_cos_sin(pos, hd, cfg["rope_theta"], x.dtype)
    mask = causal_mask(t, t, x.dtype) if t > 1 else None
    for i in range(cfg["num_hidden_layers"]):
        L = f"{pre}.layers.{i}"
        h = N.rms_norm(x, P[L + ".input_layernorm.weight"], eps)
        q = N.linear(h, P[L + ".self_attn.q_proj.weight"]).reshape(b, t, nh, hd).transpose(1, 2)
        k = N.linear(h, P[L + ".self_attn.k_proj.weight"]).reshape(b, t, nh, hd).transpose(1, 2)
        v = N.linear(h, P[L + ".self_attn.v_proj.weight"]).reshape(b, t, nh, hd).transpose(1, 2)
        q, k = apply_rope(q, k, cos, sin)
        a = N.sdpa(q, k, v, hd ** -0.5, mask).transpose(1, 2).reshape(b, t, nh * hd)
        x = x + N.linear(a, P[L + ".self_attn.o_proj.weight"]) * P[L + ".self_attn_layer_scale.scale"].to(x.dtype)
        h = N.rms_norm(x, P[L + ".post_attention_layernorm.weight"], eps)
        x = x + _mlp(P, L + ".mlp", h) * P[L + ".mlp_layer_scale.scale"].to(x.dtype)
    x = N.rms_norm(x, P[pre + ".norm.weight"], eps)
    return N.linear(x, P[pre + ".output_proj.weight"], P[pre + ".output_proj.bias"])


def convnext(P, pre, x):
    """ConvNeXtBlock.__call__ (speech_tokenizer.py:140-149)."""
    c = x.shape[-1]
    h = causal_conv(P, pre + ".dwconv.conv", x, 7, 1, c)
    h = N.layer_norm(h, P[pre + ".norm.weight"], P[pre + ".norm.bias"], 1e-6)
    h = N.gelu(N.linear(h, P[pre + ".pwconv1.weight"], P[pre + ".pwconv1.bias"]))
    h = N.linear(h, P[pre + ".pwconv2.weight"], P[pre + ".pwconv2.bias"])
    return x + P[pre + ".gamma"].to(x.dtype) * h


def tokenizer_decode(P, codes, cfg=TOKENIZER_DECODER, taps=None, stream_boundaries=()):
    """Qwen3TTSSpeechTokenizerDecoder.__call__ (speech_tokenizer.py:843-880): codes [B,16,T] -> wav [B,1,1920 T].

    ``stream_boundaries`` (frame indices where a new ``streaming_step`` call began) turns this into the concatenated output of the
    incremental decoder (speech_tokenizer.py:889-930).  Every buffered layer of that path is exact, with one exception that is kept:
    DecoderBlockUpsample.step (:645-656) overlap-adds the transposed-conv tail of the previous call INCLUDING its bias, so the bias is
    counted twice over the first ``kernel - stride`` (= stride) output samples after each boundary."""
    if codes.shape[1] != cfg["num_quantizers"]:
        raise ValueError(f"Expected {cfg['num_quantizers']} layers of codes, got {codes.shape[1]}")
    h = quantizer_decode(P, codes, cfg)
    h = causal_conv(P, "decoder.pre_conv.conv", h, 3)
    if taps is not None:
        taps["pre_conv"] = h
    h = decoder_transformer(P, h, cfg)
    if taps is not None:
        taps["transformer"] = h
    for i, f in enumerate(cfg["upsampling_ratios"]):
        h = causal_convtr(P, f"decoder.upsample.{i}.0.conv", h, f, f)
        h = convnext(P, f"decoder.upsample.{i}.1", h)
    if taps is not None:
        taps["upsample"] = h
    w = causal_conv(P, "decoder.decoder.0.conv", h, 7)
    for bi, r in enumerate(cfg["upsample_rates"]):
        B_ = f"decoder.decoder.{bi + 1}.block"
        w = snake_beta(w, P[B_ + ".0.alpha"], P[B_ + ".0.beta"])
        frames_in = w.shape[1]                                                        # NLC
        w = causal_convtr(P, B_ + ".1.conv", w, 2 * r, r)
        if stream_boundaries and (B_ + ".1.conv.bias") in P:
            per_frame = frames_in // codes.shape[-1]                                  # this block's input samples per code frame
            w = w.clone()
            for f in stream_boundaries:
                pos = f * per_frame * r
                w[:, pos: pos + r, :] += P[B_ + ".1.conv.bias"].to(w.dtype)[None, None, :]
        for ui, d in enumerate((1, 3, 9)):
            U = f"{B_}.{ui + 2}"
            y = snake_beta(w, P[U + ".act1.alpha"], P[U + ".act1.beta"])
            y = causal_conv(P, U + ".conv1.conv", y, 7, d)
            y = snake_beta(y, P[U + ".act2.alpha"], P[U + ".act2.beta"])
            w = causal_conv(P, U + ".conv2.conv", y, 1) + w
        if taps is not None:
            taps[f"block{bi}"] = w
    last = len(cfg["upsample_rates"]) + 1                                 # [0] initial conv, [1..n] blocks, [n+1] snake, [n+2] output conv
    w = snake_beta(w, P[f"decoder.decoder.{last}.alpha"], P[f"decoder.decoder.{last}.beta"])
    w = causal_conv(P, f"decoder.decoder.{last + 1}.conv", w, 7)
    return torch.clamp(w.transpose(1, 2), -1.0, 1.0)


def chunked_decode(P, codes, chunk_size=300, left_context_size=25, cfg=TOKENIZER_DECODER):
    """Qwen3TTSSpeechTokenizerDecoder.chunked_decode (speech_tokenizer.py:932-954)."""
    up = 1
    for r in list(cfg["upsample_rates"]) + list(cfg["upsampling_ratios"]):
        up *= r
    wavs, start = [], 0
    while start < codes.shape[-1]:
        end = min(start + chunk_size, codes.shape[-1])
        ctx = left_context_size if start - left_context_size > 0 else start
        w = tokenizer_decode(P, codes[..., start - ctx: end], cfg)
        wavs.append(w[..., ctx * up:])
        start = end
    return torch.cat(wavs, dim=-1)


def speech_tokenizer_decode(P, audio_codes, cfg=TOKENIZER_DECODER):
    """Qwen3TTSSpeechTokenizer.decode (speech_tokenizer.py:1099-1118): audio_codes [B,T,16] -> (wav [B,samples], lengths [B])."""
    wav = chunked_decode(P, audio_codes.transpose(1, 2), cfg=cfg).squeeze(1)
    lengths = (audio_codes[..., 0] > 0).sum(dim=1) * 1920
    return wav, lengths


def prepare_generation_inputs_from_ids(P, input_ids, tts_ids, cfg_ids, language_id=None, speaker_id=None, instruct_ids=None):
    """Model._prepare_generation_inputs (qwen3_tts.py:326-484) after tokenisation.  ``tts_ids`` = (bos, eos, pad) text-token ids
    (config.py:218-220); ``cfg_ids`` = dict with codec_nothink_id, codec_think_id, codec_think_bos_id, codec_think_eos_id,
    codec_pad_id, codec_bos_id (config.py:84-92)."""
    def text_projection(x):                                            # ResizeMLP (talker.py:339-364), silu
        h = torch.nn.functional.silu(N.linear(x, P["text_projection.linear_fc1.weight"], P["text_projection.linear_fc1.bias"]))
        return N.linear(h, P["text_projection.linear_fc2.weight"], P["text_projection.linear_fc2.bias"])
    te, ce = P["model.text_embedding.weight"], P["model.codec_embedding.weight"]
    ids = torch.as_tensor(input_ids, dtype=torch.int64).reshape(1, -1)
    text_embed = text_projection(te[ids])
    tts = text_projection(te[torch.tensor([list(tts_ids)])])
    tts_bos, tts_eos, tts_pad = tts[:, 0:1], tts[:, 1:2], tts[:, 2:3]
    if language_id is None:
        prefill = [cfg_ids["codec_nothink_id"], cfg_ids["codec_think_bos_id"], cfg_ids["codec_think_eos_id"]]
    else:
        prefill = [cfg_ids["codec_think_id"], cfg_ids["codec_think_bos_id"], language_id, cfg_ids["codec_think_eos_id"]]
    codec = ce[torch.tensor([prefill])]
    suffix = ce[torch.tensor([[cfg_ids["codec_pad_id"], cfg_ids["codec_bos_id"]]])]
    parts = [codec] + ([ce[torch.tensor([[speaker_id]])]] if speaker_id is not None else []) + [suffix]
    codec = torch.cat(parts, dim=1)
    role = text_embed[:, :3]
    combined = torch.cat([tts_pad.expand(1, codec.shape[1] - 2, -1), tts_bos], dim=1) + codec[:, :-1]
    first_text = text_embed[:, 3:4] + codec[:, -1:]
    parts = [role, combined, first_text]
    if instruct_ids is not None:                                       # instruct embedding is prepended (qwen3_tts.py:452-458,473-476)
        parts = [text_projection(te[torch.as_tensor(instruct_ids, dtype=torch.int64).reshape(1, -1)])] + parts
    input_embeds = torch.cat(parts, dim=1)
    trailing = torch.cat([text_embed[:, 4:-5], tts_eos], dim=1)
    return input_embeds, trailing, tts_pad


# ------------------------------------------------------------------------------------------------ speaker encoder (ECAPA-TDNN)
SPEAKER_ENCODER = {   # qwen3_tts/config.py:19-30
    "mel_dim": 128, "enc_dim": 1024, "enc_channels": [512, 512, 512, 512, 1536], "enc_kernel_sizes": [5, 3, 3, 3, 1],
    "enc_dilations": [1, 2, 3, 4, 1], "enc_attention_channels": 128, "enc_res2net_scale": 8, "enc_se_channels": 128,
}


def _tdnn(P, pre, x, k, dilation):
    """TimeDelayNetBlock (speaker_encoder.py:30-61) on NCL x: reflect "same" padding, conv, ReLU."""
    pad = (k - 1) * dilation // 2
    y = x.transpose(1, 2)
    if pad > 0:
        y = torch.cat([y[:, 1:pad + 1].flip(1), y, y[:, -(pad + 1):-1].flip(1)], dim=1)
    y = N.conv1d(y, P[pre + ".conv.weight"].to(x.dtype), 1, 0, dilation, 1, P[pre + ".conv.bias"])
    return torch.relu(y.transpose(1, 2))


def speaker_encoder(P, mel, cfg=SPEAKER_ENCODER, pre="speaker_encoder"):
    """Qwen3TTSSpeakerEncoder.__call__ (speaker_encoder.py:206-306): mel [B,T,128] -> embedding [B,enc_dim].  TDNN, three
    SE-Res2Net blocks (scale 8: chunk i is convolved after adding the previous chunk's output), concatenation of their outputs,
    TDNN, attentive statistics pooling (weighted mean | std), 1x1 projection."""
    ch, ks, dl, sc = cfg["enc_channels"], cfg["enc_kernel_sizes"], cfg["enc_dilations"], cfg["enc_res2net_scale"]
    x = _tdnn(P, f"{pre}.blocks.0", mel.transpose(1, 2), ks[0], dl[0])
    feats = []
    for i in range(1, len(ch) - 1):
        B_ = f"{pre}.blocks.{i}"
        r = x
        y = _tdnn(P, B_ + ".tdnn1", x, 1, 1)
        outs, prev = [], None
        for j, c in enumerate(torch.chunk(y, sc, dim=1)):                               # Res2NetBlock (:64-104)
            prev = c if j == 0 else _tdnn(P, f"{B_}.res2net_block.blocks.{j - 1}", c if j == 1 else c + prev, ks[i], dl[i])
            outs.append(prev)
        y = _tdnn(P, B_ + ".tdnn2", torch.cat(outs, dim=1), 1, 1)
        se = y.mean(dim=2, keepdim=True).transpose(1, 2)                                 # SqueezeExcitationBlock (:107-138)
        se = torch.relu(N.conv1d(se, P[B_ + ".se_block.conv1.weight"].to(y.dtype), bias=P[B_ + ".se_block.conv1.bias"]))
        se = torch.sigmoid(N.conv1d(se, P[B_ + ".se_block.conv2.weight"].to(y.dtype), bias=P[B_ + ".se_block.conv2.bias"])).transpose(1, 2)
        x = y * se + r
        feats.append(x)
    x = _tdnn(P, pre + ".mfa", torch.cat(feats, dim=1), ks[-1], dl[-1])
    eps = 1e-12                                                                          # AttentiveStatisticsPooling (:171-203)
    mean = x.mean(dim=2, keepdim=True)
    std = torch.sqrt(x.var(dim=2, unbiased=False, keepdim=True) + eps)
    a = torch.cat([x, mean.expand_as(x), std.expand_as(x)], dim=1)
    a = torch.tanh(_tdnn(P, pre + ".asp.tdnn", a, 1, 1))
    a = N.conv1d(a.transpose(1, 2), P[pre + ".asp.conv.weight"].to(x.dtype), bias=P[pre + ".asp.conv.bias"]).transpose(1, 2)
    a = torch.softmax(a, dim=2)
    mean = (a * x).sum(dim=2, keepdim=True)
    std = torch.sqrt(torch.clamp((a * (x - mean) ** 2).sum(dim=2, keepdim=True), min=eps))
    pooled = torch.cat([mean, std], dim=1).transpose(1, 2)
    return N.conv1d(pooled, P[pre + ".fc.weight"].to(x.dtype), bias=P[pre + ".fc.bias"])[:, 0]


# ------------------------------------------------------------------------------------------------ speech-tokenizer encoder (ICL)
TOKENIZER_ENCODER = {   # qwen3_tts/config.py:136-173 in oracle/codec.py's Mimi vocabulary
    "dimension": 512, "nfilters": 64, "ratios": [8, 6, 5, 4], "ksize": 7, "residual_ksize": 3, "last_ksize": 3, "compress": 2, "d_model": 512,
    "num_heads": 8, "num_layers": 8, "dim_feedforward": 2048, "context": 250, "max_period": 10000, "layer_scale": 0.01, "nq": 32, "bins": 2048,
    "qdim": 256, "upsample_stride": 2, "valid_num_quantizers": 16,
}


def tokenizer_encode(P, audio, cfg=TOKENIZER_ENCODER, root="encoder_model."):
    """Qwen3TTSSpeechTokenizerEncoder.encode (speech_tokenizer.py:1037-1058): audio [B,1,n] -> codes [B,16,ceil(n/1920)].  Mimi's SEANet
    encoder and split RVQ, with a FULL causal mask (no context window) and half-split RoPE in the transformer; only the first 16 of the
    32 code books are kept."""
    from . import codec as OC
    x = OC.mimi_seanet_encoder(P, audio, cfg, root)
    x = OC.mimi_transformer(P, root + "encoder_transformer", x, cfg, rope_traditional=False, full_causal=True)
    s = cfg["upsample_stride"]
    x = OC.mimi_causal_conv(P, root + "downsample.conv", x, 2 * s, stride=s, pad_mode="edge")
    return OC.mimi_quantizer_encode(P, x, cfg, root)[:, : cfg["valid_num_quantizers"]]


# ------------------------------------------------------------------------------------------------ ICL voice cloning
def prepare_icl_generation_inputs_from_ids(P, target_ids, ref_ids, ref_codes, tts_ids, cfg_ids, language_id=None, speaker_embed=None, cfg=TALKER):
    """Model._prepare_icl_generation_inputs (qwen3_tts.py:606-803) after tokenisation and reference encoding.  ``target_ids`` = ids of
    "<|im_start|>assistant\n{text}<|im_end|>\n<|im_start|>assistant\n", ``ref_ids`` = ids of "<|im_start|>assistant\n{ref_text}<|im_end|>\n",
    ``ref_codes`` [1, G, T_ref] from the speech-tokenizer encoder, ``speaker_embed`` [1, hidden] (x-vector) or None.
    -> (input_embeds, trailing_text_hidden = the pad embedding, tts_pad_embed)."""
    def text_projection(x):
        h = torch.nn.functional.silu(N.linear(x, P["text_projection.linear_fc1.weight"], P["text_projection.linear_fc1.bias"]))
        return N.linear(h, P["text_projection.linear_fc2.weight"], P["text_projection.linear_fc2.bias"])
    te, ce = P["model.text_embedding.weight"], P["model.codec_embedding.weight"]
    target = torch.as_tensor(target_ids, dtype=torch.int64).reshape(1, -1)
    ref = torch.as_tensor(ref_ids, dtype=torch.int64).reshape(1, -1)
    text_ids, ref_text_ids = target[:, 3:-5], ref[:, 3:-2]
    tts = text_projection(te[torch.tensor([list(tts_ids)])])
    tts_bos, tts_eos, tts_pad = tts[:, 0:1], tts[:, 1:2], tts[:, 2:3]
    text_embed = torch.cat([text_projection(te[torch.cat([ref_text_ids, text_ids], dim=1)]), tts_eos], dim=1)
    codes = torch.as_tensor(ref_codes, dtype=torch.int64)
    rc = ce[codes[:, 0]]
    for i in range(cfg["num_code_groups"] - 1):
        rc = rc + P[f"code_predictor.model.codec_embedding.{i}.weight"][codes[:, i + 1]]
    codec_icl = torch.cat([ce[torch.tensor([[cfg_ids["codec_bos_id"]]])], rc], dim=1)
    icl = torch.cat([text_embed + ce[torch.tensor([[cfg_ids["codec_pad_id"]]])], codec_icl + tts_pad], dim=1)   # all text, then all codec
    if language_id is None:
        prefill = [cfg_ids["codec_nothink_id"], cfg_ids["codec_think_bos_id"], cfg_ids["codec_think_eos_id"]]
    else:
        prefill = [cfg_ids["codec_think_id"], cfg_ids["codec_think_bos_id"], language_id, cfg_ids["codec_think_eos_id"]]
    parts = [ce[torch.tensor([prefill])]] + ([speaker_embed.reshape(1, 1, -1)] if speaker_embed is not None else []) \
        + [ce[torch.tensor([[cfg_ids["codec_pad_id"], cfg_ids["codec_bos_id"]]])]]
    prefix = torch.cat(parts, dim=1)
    combined = torch.cat([tts_pad.expand(1, prefix.shape[1] - 2, -1), tts_bos], dim=1) + prefix[:, :-1]
    return torch.cat([text_projection(te[target[:, :3]]), combined, icl], dim=1), tts_pad, tts_pad


def decode_icl_generated_codes(PT, gen_codes, ref_codes, cfg=TOKENIZER_DECODER):
    """Model._decode_icl_generated_codes (qwen3_tts.py:1085-1112): decode [reference codes | generated codes] together, keep the valid
    length, cut the reference's share ``int(ref_len / total_len * n_samples)`` off the front."""
    ref_t = torch.as_tensor(ref_codes, dtype=torch.int64).transpose(1, 2)
    full = torch.cat([ref_t, gen_codes[None]], dim=1)
    wav, lengths = speech_tokenizer_decode(PT, full, cfg)
    audio = wav[0]
    valid = int(lengths[0])
    if 0 < valid < audio.shape[0]:
        audio = audio[:valid]
    cut = int(ref_t.shape[1] / max(full.shape[1], 1) * audio.shape[0])
    return audio[cut:] if 0 < cut < audio.shape[0] else audio


def decode_generated_codes(PT, codes, cfg=TOKENIZER_DECODER, decode_chunk=15, decode_ctx=5):
    """Model._decode_generated_codes (qwen3_tts.py:1050-1083), the decode of the default (non-streaming) batch path: 15-frame chunks,
    each decoded with up to 5 frames of left context whose samples are dropped; no valid-length trimming.  codes [n, G] -> wav [1920 n]."""
    up = 1
    for r in list(cfg["upsample_rates"]) + list(cfg["upsampling_ratios"]):
        up *= r
    t = codes[None].transpose(1, 2)
    parts, start, n = [], 0, t.shape[-1]
    while start < n:
        end = min(start + decode_chunk, n)
        ctx = decode_ctx if start > decode_ctx else start
        wav = tokenizer_decode(PT, t[..., start - ctx: end], cfg)[0, 0]
        parts.append(wav[ctx * up:])
        start = end
    return torch.cat(parts)
