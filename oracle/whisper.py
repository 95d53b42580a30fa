"""Oracle for the Whisper audio encoder (stt/models/whisper/whisper.py:329-448) -- torch-CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The log-mel frontend is in oracle/dsp.py
(``whisper_log_mel``).  Parameter names are the reference's MLX parameter tree.
"""
from __future__ import annotations

import math

import torch

from . import nn as N

WHISPER_SMALL = {"n_mels": 80, "n_audio_ctx": 1500, "n_audio_state": 768, "n_audio_head": 12, "n_audio_layer": 12,
                 "n_vocab": 51865, "n_text_ctx": 448, "n_text_state": 768, "n_text_head": 12, "n_text_layer": 12}


def sinusoids(length, channels, max_timescale=10000, dtype=torch.float64):
    """whisper.py:329-335."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=dtype))
    st = torch.arange(length, dtype=dtype)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def mha(P, pre, x, n_head, xa=None, mask=None):
    """MultiHeadAttention (whisper.py:338-385): q and k each scaled by d^-0.25, softmax in the working dtype."""
    q = N.linear(x, P[pre + ".query.weight"], P[pre + ".query.bias"])
    src = x if xa is None else xa
    k = N.linear(src, P[pre + ".key.weight"])
    v = N.linear(src, P[pre + ".value.weight"], P[pre + ".value.bias"])
    b, t, d = q.shape
    scale = (d // n_head) ** -0.25
    qh = q.reshape(b, t, n_head, -1).permute(0, 2, 1, 3) * scale
    kh = k.reshape(b, k.shape[1], n_head, -1).permute(0, 2, 3, 1) * scale
    vh = v.reshape(b, v.shape[1], n_head, -1).permute(0, 2, 1, 3)
    qk = qh @ kh
    if mask is not None:
        qk = qk + mask[:t, :t]
    out = (torch.softmax(qk, dim=-1) @ vh).permute(0, 2, 1, 3).reshape(b, t, d)
    return N.linear(out, P[pre + ".out.weight"], P[pre + ".out.bias"])


def encoder(P, mel, dims=WHISPER_SMALL):
    """AudioEncoder.__call__ (whisper.py:438-448): mel [B, 3000, n_mels] (frames-major) -> [B, 1500, d]."""
    dt = mel.dtype
    x = N.gelu(N.conv1d(mel, P["encoder.conv1.weight"].to(dt), 1, 1, 1, 1, P["encoder.conv1.bias"]))
    x = N.gelu(N.conv1d(x, P["encoder.conv2.weight"].to(dt), 2, 1, 1, 1, P["encoder.conv2.bias"]))
    assert x.shape[1:] == (dims["n_audio_ctx"], dims["n_audio_state"]), "incorrect audio shape"
    x = x + sinusoids(dims["n_audio_ctx"], dims["n_audio_state"], dtype=dt)
    for i in range(dims["n_audio_layer"]):
        L = f"encoder.blocks.{i}"
        x = x + mha(P, L + ".attn", N.layer_norm(x, P[L + ".attn_ln.weight"], P[L + ".attn_ln.bias"]), dims["n_audio_head"])
        h = N.layer_norm(x, P[L + ".mlp_ln.weight"], P[L + ".mlp_ln.bias"])
        x = x + N.linear(N.gelu(N.linear(h, P[L + ".mlp1.weight"], P[L + ".mlp1.bias"])), P[L + ".mlp2.weight"], P[L + ".mlp2.bias"])
    return N.layer_norm(x, P["encoder.ln_post.weight"], P["encoder.ln_post.bias"])


# ============================================================================= text decoder + greedy decoding (row a8)

class TokenizerSpec:
    """The tokenizer constants the decode loop touches (decoding.py:349-442, whisper.py:46-175).  Defaults are the public
    ids of Whisper's multilingual vocabulary; no tokenizer files are needed for the numeric path."""
    def __init__(self, eot=50257, sot=50258, no_timestamps=50363, timestamp_begin=50364, no_speech=50362, blank_ids=(220,),
                 language=50259, task=50359, transcribe=50359, translate=50358, sot_lm=50360, sot_prev=50361, non_speech_tokens=None):
        self.eot, self.sot, self.no_timestamps, self.timestamp_begin, self.no_speech = eot, sot, no_timestamps, timestamp_begin, no_speech
        self.blank_ids, self.language, self.task = tuple(blank_ids), language, task
        self.transcribe, self.translate, self.sot_lm, self.sot_prev, self.non_speech_tokens = transcribe, translate, sot_lm, sot_prev, non_speech_tokens

    @property
    def sot_sequence(self):
        return (self.sot, self.language, self.task)


def get_suppress_tokens(spec, suppress_tokens):
    """decoding.py:79-112: the ids SuppressTokens masks.  -1 expands to the tokenizer's non-speech tokens; the five task / sot markers and
    no_speech are ALWAYS added.  (A falsy option means no SuppressTokens filter at all, decoding.py:489-495 -- the caller's check.)"""
    result = list(suppress_tokens) if suppress_tokens else []
    if -1 in result:
        if spec.non_speech_tokens is None:
            raise ValueError("suppress_tokens contains -1 but the spec carries no non_speech_tokens (they come from the tokenizer files)")
        result = [t for t in result if t >= 0] + list(spec.non_speech_tokens)
    result.extend([spec.transcribe, spec.translate, spec.sot, spec.sot_prev, spec.sot_lm])
    if spec.no_speech is not None:
        result.append(spec.no_speech)
    return tuple(sorted(set(result)))


def decoder_forward(P, tokens, xa, kv_cache=None, dims=WHISPER_SMALL):
    """TextDecoder.__call__ (whisper.py:476-498): tokens [B,n] int64, xa [B,1500,d] -> (logits [B,n,V], kv_cache)."""
    dt = xa.dtype
    nl, nh = dims["n_text_layer"], dims["n_text_head"]
    offset = kv_cache[0][0][0].shape[1] if kv_cache else 0
    x = P["decoder.token_embedding.weight"].to(dt)[tokens] + P["decoder.positional_embedding"].to(dt)[offset:offset + tokens.shape[-1]]
    n_ctx = dims["n_text_ctx"]
    mask = torch.triu(torch.full((n_ctx, n_ctx), float("-inf"), dtype=dt), diagonal=1)     # create_additive_causal_mask
    if kv_cache is None:
        kv_cache = [None] * nl
    new_cache = []
    for i in range(nl):
        L = f"decoder.blocks.{i}"
        kv, cross_kv = kv_cache[i] if kv_cache[i] else (None, None)
        # self attention with concatenated KV (whisper.py:354-361)
        h = N.layer_norm(x, P[L + ".attn_ln.weight"], P[L + ".attn_ln.bias"])
        q = N.linear(h, P[L + ".attn.query.weight"], P[L + ".attn.query.bias"])
        k = N.linear(h, P[L + ".attn.key.weight"])
        v = N.linear(h, P[L + ".attn.value.weight"], P[L + ".attn.value.bias"])
        if kv is not None:
            k, v = torch.cat([kv[0], k], 1), torch.cat([kv[1], v], 1)
        x = x + N.linear(_qkv_attention(q, k, v, nh, mask), P[L + ".attn.out.weight"], P[L + ".attn.out.bias"])
        # cross attention, K/V computed once (whisper.py:362-366)
        h = N.layer_norm(x, P[L + ".cross_attn_ln.weight"], P[L + ".cross_attn_ln.bias"])
        q = N.linear(h, P[L + ".cross_attn.query.weight"], P[L + ".cross_attn.query.bias"])
        if cross_kv is None:
            cross_kv = (N.linear(xa, P[L + ".cross_attn.key.weight"]), N.linear(xa, P[L + ".cross_attn.value.weight"], P[L + ".cross_attn.value.bias"]))
        x = x + N.linear(_qkv_attention(q, cross_kv[0], cross_kv[1], nh, None), P[L + ".cross_attn.out.weight"], P[L + ".cross_attn.out.bias"])
        h = N.layer_norm(x, P[L + ".mlp_ln.weight"], P[L + ".mlp_ln.bias"])
        x = x + N.linear(N.gelu(N.linear(h, P[L + ".mlp1.weight"], P[L + ".mlp1.bias"])), P[L + ".mlp2.weight"], P[L + ".mlp2.bias"])
        new_cache.append(((k, v), cross_kv))
    x = N.layer_norm(x, P["decoder.ln.weight"], P["decoder.ln.bias"])
    return x @ P["decoder.token_embedding.weight"].to(dt).T, new_cache


def _qkv_attention(q, k, v, n_head, mask):
    """whisper.py:371-385 (q and k each scaled by d^-0.25; mask[:n_ctx,:n_ctx] added)."""
    b, t, d = q.shape
    scale = (d // n_head) ** -0.25
    qh = q.reshape(b, t, n_head, -1).permute(0, 2, 1, 3) * scale
    kh = k.reshape(b, k.shape[1], n_head, -1).permute(0, 2, 3, 1) * scale
    vh = v.reshape(b, v.shape[1], n_head, -1).permute(0, 2, 1, 3)
    qk = qh @ kh
    if mask is not None:
        qk = qk + mask[:t, :t]
    return (torch.softmax(qk, dim=-1) @ vh).permute(0, 2, 1, 3).reshape(b, t, d)


def apply_filters(logits, tokens, spec, sample_begin, suppress, max_initial_timestamp_index=50, without_timestamps=False):
    """SuppressBlank -> SuppressTokens -> ApplyTimestampRules (decoding.py:349-442) on logits [B,V], tokens [B,n] (lists)."""
    logits = logits.clone()
    V = logits.shape[1]
    if len(tokens[0]) == sample_begin:                                               # SuppressBlank
        logits[:, list(spec.blank_ids) + [spec.eot]] = float("-inf")
    if suppress:
        logits[:, list(suppress)] = float("-inf")                                    # SuppressTokens
    if without_timestamps:
        return logits
    mask = torch.zeros_like(logits)
    mask[:, spec.no_timestamps] = float("-inf")
    tb = spec.timestamp_begin
    for k, row in enumerate(tokens):
        seq = row[sample_begin:]
        last_ts = len(seq) >= 1 and seq[-1] >= tb
        pen_ts = len(seq) < 2 or seq[-2] >= tb
        if last_ts:
            if pen_ts:
                mask[k, tb:] = float("-inf")
            else:
                mask[k, :spec.eot] = float("-inf")
        # Quirk kept (decoding.py:400-408): the reference collects the POSITIONS `i` of the timestamp tokens, not their
        # values, so `mask[k, timestamp_begin : last_timestamp]` slices [50364 : small index) -- an empty range.  The
        # "timestamps must not decrease" rule of the original Whisper is therefore inert here; restated literally.
        ts = [i for i, v in enumerate(seq) if v > tb]
        if ts:
            last = ts[-1]
            if not last or pen_ts:
                last += 1
            mask[k, tb:last] = float("-inf")
    if len(tokens[0]) == sample_begin:
        mask[:, :tb] = float("-inf")
        if max_initial_timestamp_index is not None:
            mask[:, tb + max_initial_timestamp_index + 1:] = float("-inf")
    logprobs = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    ts_lp = torch.logsumexp(logprobs[:, tb:], dim=-1, keepdim=True)
    max_text = logprobs[:, :tb].max(dim=-1, keepdim=True).values
    mask[:, :tb] = torch.where(ts_lp > max_text, torch.full_like(mask[:, :tb], float("-inf")), mask[:, :tb])
    return logits + mask


def greedy_update(tokens, logits, sum_logprobs, eot):
    """GreedyDecoder.update at temperature 0 (decoding.py:307-325)."""
    nxt = logits.argmax(dim=-1)
    logprobs = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    cur = logprobs[torch.arange(logits.shape[0]), nxt]
    last = torch.tensor([t[-1] for t in tokens])
    sum_logprobs = sum_logprobs + cur * (last != eot)
    nxt = torch.where(last == eot, torch.full_like(nxt, eot), nxt)
    tokens = [t + [int(n)] for t, n in zip(tokens, nxt)]
    return tokens, bool((nxt == eot).all()), sum_logprobs


def greedy_decode(P, xa, spec, sample_len=16, suppress=(), dims=WHISPER_SMALL, max_initial_timestamp_index=50, without_timestamps=False):
    """DecodingTask._main_loop with GreedyDecoder(temperature=0) (decoding.py:588-632) on encoder features xa [B,1500,d].
    Returns (tokens incl. the sot sequence, sum_logprobs, no_speech_probs)."""
    B = xa.shape[0]
    suppress = get_suppress_tokens(spec, suppress) if suppress else ()            # DecodingTask.__init__, decoding.py:489-495
    init = list(spec.sot_sequence) + ([spec.no_timestamps] if without_timestamps else [])     # sot_sequence_including_notimestamps, :463-465
    tokens = [list(init) for _ in range(B)]
    sample_begin = len(init)
    sum_lp = torch.zeros(B, dtype=xa.dtype)
    cache = None
    no_speech = None
    for i in range(sample_len):
        inp = torch.tensor(tokens if i == 0 else [[t[-1]] for t in tokens])
        pre, cache = decoder_forward(P, inp, xa, cache, dims)
        if i == 0:
            no_speech = torch.softmax(pre[:, 0], dim=-1)[:, spec.no_speech]          # sot_index = 0
        logits = apply_filters(pre[:, -1], tokens, spec, sample_begin, suppress, max_initial_timestamp_index, without_timestamps)
        tokens, completed, sum_lp = greedy_update(tokens, logits, sum_lp, spec.eot)
        if completed or len(tokens[0]) > dims["n_text_ctx"]:
            break
    return tokens, sum_lp, no_speech


# ------------------------------------------------------------------------------------------------ sampling + long-form loop
def sample_update(tokens, logits, sum_logprobs, eot, temperature, u):
    """GreedyDecoder.update at temperature > 0 (decoding.py:295-325): categorical draw from softmax(logits / temperature) -- defined, as
    everywhere in this repository, as the inverse CDF in index order driven by one uniform ``u[b]`` per row (MLX's PRNG is not
    reproducible) -- while the log-probability bookkeeping uses the un-tempered logits."""
    B = logits.shape[0]
    nxt = []
    u = torch.as_tensor(u, dtype=torch.float64).reshape(-1).expand(B) if torch.as_tensor(u).numel() == 1 else torch.as_tensor(u, dtype=torch.float64).reshape(-1)
    for b in range(B):
        row = logits[b].double() / float(temperature)
        w = torch.exp(row - row.max())
        cum = torch.cumsum(w, 0)
        live = torch.nonzero(w > 0).reshape(-1)
        hit = live[cum[live] > float(u[b]) * float(cum[-1])]
        nxt.append(int(hit[0]) if hit.numel() else int(live[-1]))
    nxt = torch.tensor(nxt)
    logprobs = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    cur = logprobs[torch.arange(B), nxt]
    last = torch.tensor([t[-1] for t in tokens])
    sum_logprobs = sum_logprobs + cur * (last != eot)
    nxt = torch.where(last == eot, torch.full_like(nxt, eot), nxt)
    tokens = [t + [int(n)] for t, n in zip(tokens, nxt)]
    return tokens, bool((nxt == eot).all()), sum_logprobs


def compression_ratio(text):
    """decoding.py:19-21."""
    import zlib
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


def decode_window(P, xa, spec, dims, *, temperature=0.0, uniforms=None, prompt=(), sample_len=None, suppress=(), max_initial_timestamp_index=50,
                  without_timestamps=False, tokenizer=None):
    """DecodingTask.run for one window (decoding.py:445-722): initial tokens incl. the previous-text prompt, the sampling loop, the
    result fields.  xa [1, n_audio_ctx, d] encoder features.  Returns a dict with the DecodingResult fields."""
    n_ctx = dims["n_text_ctx"]
    sample_len = sample_len or n_ctx // 2
    sup = get_suppress_tokens(spec, suppress) if suppress else ()
    init = list(spec.sot_sequence) + ([spec.no_timestamps] if without_timestamps else [])
    if prompt:
        init = [spec.sot_prev] + list(prompt)[-(n_ctx // 2 - 1):] + init
    sb, sot_index = len(init), init.index(spec.sot)
    tokens = [list(init)]
    sum_lp = torch.zeros(1, dtype=xa.dtype)
    cache, no_speech = None, None
    for i in range(sample_len):
        inp = torch.tensor(tokens if i == 0 else [[t[-1]] for t in tokens])
        pre, cache = decoder_forward(P, inp, xa, cache, dims)
        if i == 0:
            no_speech = torch.softmax(pre[:, sot_index], dim=-1)[:, spec.no_speech]
        logits = apply_filters(pre[:, -1], tokens, spec, sb, sup, max_initial_timestamp_index, without_timestamps)
        if temperature == 0:
            tokens, completed, sum_lp = greedy_update(tokens, logits, sum_lp, spec.eot)
        else:
            tokens, completed, sum_lp = sample_update(tokens, logits, sum_lp, spec.eot, temperature, uniforms[i])
        if completed or len(tokens[0]) > n_ctx:
            break
    row = tokens[0] + [spec.eot]                              # GreedyDecoder.finalize
    cut = row[sb:]
    cut = cut[: cut.index(spec.eot)]
    text = tokenizer.decode(cut).strip() if tokenizer is not None else ""
    return {"tokens": cut, "text": text, "avg_logprob": float(sum_lp[0]) / (len(cut) + 1), "no_speech_prob": float(no_speech[0]),
            "temperature": float(temperature), "compression_ratio": compression_ratio(text)}


def transcribe(P, mel, spec, dims, tokenizer, *, temperatures=(0.0, 0.2, 0.4, 0.6, 0.8, 1.0), compression_ratio_threshold=2.4,
               logprob_threshold=-1.0, no_speech_threshold=0.6, condition_on_previous_text=True, initial_prompt_tokens=(), return_timestamps=True,
               clip_timestamps=(0.0,), suppress=(), sample_len=None, uniforms=None, n_frames=3000, hop=160, sr=16000):
    """Model.generate's window loop (whisper.py:934-1318) on a log-mel spectrogram ``mel`` [frames + n_frames, n_mels] (already padded by
    one window of zeros, _prepare_audio :748-775).  ``uniforms(k)`` returns the per-step uniforms of the k-th decode call made at a
    temperature > 0.  Returns (text, segments)."""
    content_frames = mel.shape[0] - n_frames
    fps = sr // hop
    points = [round(ts * fps) for ts in clip_timestamps] or [0]
    if len(points) % 2 == 1:
        points.append(content_frames)
    else:
        points[-1] = min(content_frames, points[-1])
    clips = list(zip(points[::2], points[1::2]))
    input_stride = n_frames // dims["n_audio_ctx"]
    time_precision = input_stride * hop / sr
    tb, eot = spec.timestamp_begin, spec.eot
    precision = 30.0 / dims["n_audio_ctx"]
    mi = round(1.0 / precision)
    all_tokens = list(initial_prompt_tokens)
    all_segments, prompt_reset_since, hot_calls = [], 0, [0]

    def with_fallback(segment):
        xa = encoder(P, segment[None], dims)
        res = None
        for t in temperatures:
            u = None
            if t > 0:
                u = uniforms(hot_calls[0])
                hot_calls[0] += 1
            res = decode_window(P, xa, spec, dims, temperature=t, uniforms=u, prompt=all_tokens[prompt_reset_since:], sample_len=sample_len,
                                suppress=suppress, max_initial_timestamp_index=mi, without_timestamps=not return_timestamps, tokenizer=tokenizer)
            bad = False
            if compression_ratio_threshold is not None and res["compression_ratio"] > compression_ratio_threshold:
                bad = True
            if logprob_threshold is not None and res["avg_logprob"] < logprob_threshold:
                bad = True
            if no_speech_threshold is not None and res["no_speech_prob"] > no_speech_threshold:
                bad = False
            if not bad:
                break
        return res

    seek = clips[0][0]
    for _, clip_end in clips:
        while seek < clip_end:
            time_offset = float(seek * hop / sr)
            segment_size = min(n_frames, content_frames - seek, clip_end - seek)
            seg = mel[seek:seek + segment_size]
            if seg.shape[0] < n_frames:
                seg = torch.cat([seg, torch.zeros(n_frames - seg.shape[0], seg.shape[1], dtype=seg.dtype)], 0)
            res = with_fallback(seg)
            tokens = res["tokens"]
            if no_speech_threshold is not None:
                skip = res["no_speech_prob"] > no_speech_threshold
                if logprob_threshold is not None and res["avg_logprob"] > logprob_threshold:
                    skip = False
                if skip:
                    seek += segment_size
                    continue
            cur = []

            def mk(start, end, toks):
                return {"seek": seek, "start": float(start), "end": float(end), "text": tokenizer.decode([t for t in toks if t < eot]), "tokens": list(toks),
                        "temperature": res["temperature"], "avg_logprob": res["avg_logprob"], "compression_ratio": res["compression_ratio"],
                        "no_speech_prob": res["no_speech_prob"]}
            is_ts = [t >= tb for t in tokens]
            single = is_ts[-2:] == [False, True]
            cons = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
            if cons:
                if single:
                    cons.append(len(tokens))
                last = 0
                for c in cons:
                    sl = tokens[last:c]
                    cur.append(mk(time_offset + (sl[0] - tb) * time_precision, time_offset + (sl[-1] - tb) * time_precision, sl))
                    last = c
                if single:
                    seek += segment_size
                else:
                    seek += (tokens[last - 1] - tb) * input_stride
            else:
                duration = segment_size * hop / sr
                stamps = [t for t in tokens if t >= tb]
                if stamps and stamps[-1] != tb:
                    duration = (stamps[-1] - tb) * time_precision
                cur.append(mk(time_offset, time_offset + duration, tokens))
                seek += segment_size
            for sg in cur:
                if sg["start"] == sg["end"] or sg["text"].strip() == "":
                    sg["text"], sg["tokens"], sg["words"] = "", [], []
            all_segments.extend({"id": i, **sg} for i, sg in enumerate(cur, start=len(all_segments)))
            all_tokens.extend(t for sg in cur for t in sg["tokens"])
            if not condition_on_previous_text or res["temperature"] > 0.5:
                prompt_reset_since = len(all_tokens)
    return tokenizer.decode(all_tokens[len(initial_prompt_tokens):]), all_segments
