"""Synthetic checkpoints in the layouts the hub ships (HuggingFace / PyTorch names and axis orders), shared by
make_sanitize_golden.py (which feeds them to the REFERENCE's sanitize functions) and tests/test_host_cpu.py (which feeds the same
dicts to the product's).  Values come from synth_params.value(name, shape), float32."""
import numpy as np

import synth_params

WHISPER_DIMS = dict(n_mels=80, n_audio_ctx=60, n_audio_state=64, n_audio_head=4, n_audio_layer=2, n_vocab=300, n_text_ctx=32, n_text_state=64,
                    n_text_head=4, n_text_layer=2)


def _fill(entries):
    return {n: synth_params.value(n, sh).astype(np.float32) for n, sh in entries}


def whisper_hf(d=WHISPER_DIMS):
    """transformers' WhisperForConditionalGeneration state dict: ``model.`` prefix, (out, in, K) conv weights, no k_proj bias."""
    e = []
    a, t = d["n_audio_state"], d["n_text_state"]

    def attn(pre, n):
        for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
            e.append((f"{pre}.{p}.weight", (n, n)))
            if p != "k_proj":
                e.append((f"{pre}.{p}.bias", (n,)))

    def ln(pre, n):
        e.extend([(pre + ".weight", (n,)), (pre + ".bias", (n,))])
    e += [("model.encoder.conv1.weight", (a, d["n_mels"], 3)), ("model.encoder.conv1.bias", (a,)), ("model.encoder.conv2.weight", (a, a, 3)),
          ("model.encoder.conv2.bias", (a,)), ("model.encoder.embed_positions.weight", (d["n_audio_ctx"], a))]
    for i in range(d["n_audio_layer"]):
        L = f"model.encoder.layers.{i}"
        attn(L + ".self_attn", a)
        ln(L + ".self_attn_layer_norm", a)
        e += [(L + ".fc1.weight", (4 * a, a)), (L + ".fc1.bias", (4 * a,)), (L + ".fc2.weight", (a, 4 * a)), (L + ".fc2.bias", (a,))]
        ln(L + ".final_layer_norm", a)
    ln("model.encoder.layer_norm", a)
    e += [("model.decoder.embed_tokens.weight", (d["n_vocab"], t)), ("model.decoder.embed_positions.weight", (d["n_text_ctx"], t))]
    for i in range(d["n_text_layer"]):
        L = f"model.decoder.layers.{i}"
        attn(L + ".self_attn", t)
        ln(L + ".self_attn_layer_norm", t)
        attn(L + ".encoder_attn", t)
        ln(L + ".encoder_attn_layer_norm", t)
        e += [(L + ".fc1.weight", (4 * t, t)), (L + ".fc1.bias", (4 * t,)), (L + ".fc2.weight", (t, 4 * t)), (L + ".fc2.bias", (t,))]
        ln(L + ".final_layer_norm", t)
    ln("model.decoder.layer_norm", t)
    return _fill(e)


def qwen3_model_torch():
    """Keys of a Qwen3-TTS ``model.safetensors`` that Model.sanitize (qwen3_tts.py:2914-2935) has rules for: position_ids, conv weights in
    either axis order (the shape heuristic decides), the speaker encoder's 1x1 ``fc``, and plain matrices that must pass through."""
    return _fill([("talker.model.layers.0.self_attn.q_proj.weight", (64, 32)), ("talker.model.codec_embedding.weight", (50, 32)),
                  ("talker.rotary.position_ids", (1, 16)), ("speaker_encoder.blocks.0.conv.weight", (48, 40, 5)),
                  ("speaker_encoder.blocks.1.conv.weight", (48, 3, 96)), ("speaker_encoder.blocks.2.conv.weight", (48, 96, 1)),
                  ("speaker_encoder.blocks.3.conv.weight", (48, 1, 96)), ("speaker_encoder.blocks.4.conv.weight", (48, 1, 5)),
                  ("speaker_encoder.blocks.5.conv.weight", (48, 7, 1)), ("speaker_encoder.fc.weight", (32, 96, 1)), ("speaker_encoder.fc.bias", (32,)),
                  ("speaker_encoder.asp.conv.bias", (48,)), ("talker.code_predictor.lm_head.0.weight", (40, 32))])


def qwen3_tokenizer_torch():
    """Keys of ``speech_tokenizer/model.safetensors`` (decoder half + a few encoder keys, which the decoder-only product must ignore and the
    reference maps elsewhere): PyTorch conv (out, in, K), transposed conv (in, out, K), codebooks as embedding_sum + cluster_usage."""
    e = [("decoder.pre_conv.conv.weight", (32, 16, 3)), ("decoder.pre_conv.conv.bias", (32,)),
         ("decoder.pre_transformer.layers.0.self_attn.q_proj.weight", (32, 32)), ("decoder.pre_transformer.input_proj.weight", (32, 32)),
         ("decoder.pre_transformer.layers.0.self_attn_layer_scale.scale", (32,)),
         ("decoder.quantizer.rvq_first.output_proj.weight", (16, 8, 1)), ("decoder.quantizer.rvq_rest.output_proj.weight", (16, 8, 1)),
         ("decoder.quantizer.rvq_first.vq.layers.0._codebook.embedding_sum", (80, 8)), ("decoder.quantizer.rvq_first.vq.layers.0._codebook.cluster_usage", (80,)),
         ("decoder.quantizer.rvq_rest.vq.layers.0._codebook.embedding_sum", (80, 8)), ("decoder.quantizer.rvq_rest.vq.layers.0._codebook.cluster_usage", (80,)),
         ("decoder.quantizer.rvq_rest.vq.layers.1._codebook.embedding_sum", (80, 8)), ("decoder.quantizer.rvq_rest.vq.layers.1._codebook.cluster_usage", (80,)),
         ("decoder.upsample.0.0.conv.weight", (32, 32, 2)), ("decoder.upsample.0.0.conv.bias", (32,)), ("decoder.upsample.0.1.dwconv.conv.weight", (32, 1, 7)),
         ("decoder.upsample.0.1.pwconv1.weight", (128, 32)), ("decoder.upsample.0.1.gamma", (32,)),
         ("decoder.decoder.0.conv.weight", (48, 32, 7)), ("decoder.decoder.1.block.0.alpha", (48,)), ("decoder.decoder.1.block.1.conv.weight", (48, 24, 16)),
         ("decoder.decoder.1.block.2.conv1.conv.weight", (24, 24, 7)), ("decoder.decoder.1.block.2.conv2.conv.weight", (24, 24, 1)),
         ("decoder.decoder.2.block.1.conv.weight", (24, 12, 10)), ("decoder.decoder.6.conv.weight", (1, 3, 7)), ("decoder.decoder.6.conv.bias", (1,))]
    return _fill(e)


def kokoro_torch():
    """Key patterns of the PyTorch Kokoro-82M checkpoint that Model.sanitize (kokoro.py:179-276) and Decoder.sanitize (istftnet.py:999-1011)
    have rules for: ALBERT position_ids, LayerNorm gamma / beta, torch LSTM names, weight_v in either axis order, the 1x1 F0 / N projections
    and the generator's noise convs."""
    return _fill([("bert.embeddings.position_ids", (1, 512)), ("bert.embeddings.word_embeddings.weight", (178, 128)),
                  ("bert.encoder.albert_layer_groups.0.albert_layers.0.ffn.weight", (64, 32)), ("bert_encoder.weight", (32, 48)), ("bert_encoder.bias", (32,)),
                  ("text_encoder.embedding.weight", (178, 32)), ("text_encoder.cnn.0.0.weight_v", (32, 32, 5)), ("text_encoder.cnn.0.0.weight_g", (32, 1, 1)),
                  ("text_encoder.cnn.0.0.bias", (32,)), ("text_encoder.cnn.0.1.gamma", (32,)), ("text_encoder.cnn.0.1.beta", (32,)),
                  ("text_encoder.lstm.weight_ih_l0", (64, 32)), ("text_encoder.lstm.weight_hh_l0", (64, 16)), ("text_encoder.lstm.bias_ih_l0", (64,)),
                  ("text_encoder.lstm.bias_hh_l0", (64,)), ("text_encoder.lstm.weight_ih_l0_reverse", (64, 32)),
                  ("text_encoder.lstm.weight_hh_l0_reverse", (64, 16)), ("text_encoder.lstm.bias_ih_l0_reverse", (64,)),
                  ("text_encoder.lstm.bias_hh_l0_reverse", (64,)),
                  ("predictor.lstm.weight_ih_l0", (64, 40)), ("predictor.lstm.bias_hh_l0_reverse", (64,)), ("predictor.text_encoder.lstms.1.fc.weight", (64, 16)),
                  ("predictor.F0.0.conv1.weight_v", (32, 32, 3)), ("predictor.F0.1.pool.weight_v", (32, 1, 3)), ("predictor.F0.1.conv1x1.weight_v", (16, 32, 1)),
                  ("predictor.F0_proj.weight", (1, 16, 1)), ("predictor.F0_proj.bias", (1,)), ("predictor.N_proj.weight", (1, 16, 1)),
                  ("predictor.duration_proj.linear_layer.weight", (50, 32)),
                  ("decoder.encode.conv1.weight_v", (96, 34, 3)), ("decoder.encode.conv1.weight_g", (96, 1, 1)), ("decoder.asr_res.0.weight_v", (64, 512, 1)),
                  ("decoder.F0_conv.weight_v", (1, 1, 3)), ("decoder.generator.ups.0.weight_v", (64, 32, 20)), ("decoder.generator.ups.1.weight_v", (32, 12, 16)),
                  ("decoder.generator.noise_convs.0.weight", (32, 22, 12)), ("decoder.generator.noise_convs.0.bias", (32,)),
                  ("decoder.generator.noise_convs.1.weight", (16, 22, 1)), ("decoder.generator.resblocks.0.convs1.0.weight_v", (32, 32, 3)),
                  ("decoder.generator.resblocks.0.alpha1.0", (1, 32, 1)), ("decoder.generator.conv_post.weight_v", (22, 16, 7)),
                  ("decoder.generator.m_source.l_linear.weight", (1, 9))])


def mimi_torch():
    """Key patterns of kyutai's PyTorch Mimi checkpoint that Mimi.load_pytorch_weights (mimi.py:192-262) renames / re-lays-out."""
    e = [("encoder.model.0.conv.conv.weight", (4, 1, 7)), ("encoder.model.0.conv.conv.bias", (4,)), ("encoder.model.1.block.1.conv.conv.weight", (2, 4, 3)),
         ("encoder.model.1.block.3.conv.conv.weight", (4, 2, 1)), ("encoder.model.3.conv.conv.weight", (8, 4, 8)), ("encoder.model.4.block.1.conv.conv.bias", (4,)),
         ("encoder.model.6.conv.conv.weight", (16, 8, 10)), ("encoder.model.9.conv.conv.weight", (32, 16, 12)), ("encoder.model.10.block.3.conv.conv.weight", (32, 16, 1)),
         ("encoder.model.12.conv.conv.weight", (64, 32, 16)), ("encoder.model.14.conv.conv.weight", (32, 64, 3)),
         ("decoder.model.0.conv.conv.weight", (64, 32, 7)), ("decoder.model.2.convtr.convtr.weight", (64, 32, 16)), ("decoder.model.2.convtr.convtr.bias", (32,)),
         ("decoder.model.3.block.1.conv.conv.weight", (16, 32, 3)), ("decoder.model.3.block.3.conv.conv.weight", (32, 16, 1)),
         ("decoder.model.5.convtr.convtr.weight", (32, 16, 12)), ("decoder.model.6.block.1.conv.conv.bias", (8,)), ("decoder.model.8.convtr.convtr.weight", (16, 8, 10)),
         ("decoder.model.11.convtr.convtr.weight", (8, 4, 8)), ("decoder.model.12.block.3.conv.conv.weight", (4, 2, 1)), ("decoder.model.14.conv.conv.weight", (1, 4, 3)),
         ("encoder_transformer.transformer.layers.0.self_attn.in_proj_weight", (96, 32)), ("encoder_transformer.transformer.layers.0.self_attn.out_proj.weight", (32, 32)),
         ("decoder_transformer.transformer.layers.1.self_attn.in_proj_weight", (96, 32)), ("decoder_transformer.transformer.layers.1.linear1.weight", (64, 32)),
         ("decoder_transformer.transformer.layers.1.linear2.weight", (32, 64)), ("decoder_transformer.transformer.layers.1.norm1.weight", (32,)),
         ("decoder_transformer.transformer.layers.1.layer_scale_2.scale", (32,)),
         ("quantizer.rvq_first.input_proj.weight", (16, 32, 1)), ("quantizer.rvq_first.output_proj.weight", (32, 16, 1)),
         ("quantizer.rvq_first.vq.layers.0._codebook._initialized", (1,)), ("quantizer.rvq_first.vq.layers.0._codebook.embedding_sum", (64, 16)),
         ("quantizer.rvq_first.vq.layers.0._codebook.cluster_usage", (64,)), ("quantizer.rvq_rest.vq.layers.2._codebook.embedding_sum", (64, 16)),
         ("quantizer.rvq_rest.output_proj.weight", (32, 16, 1)), ("downsample.conv.conv.conv.weight", (32, 32, 4)), ("upsample.convtr.convtr.convtr.weight", (32, 1, 4))]
    return _fill(e)


HF_MIMI_SMALL = dict(hidden_size=32, num_filters=4, upsampling_ratios=[8, 6, 5, 4], intermediate_size=64, num_attention_heads=4, num_key_value_heads=4,
                     head_dim=8, num_hidden_layers=2, codebook_size=64, codebook_dim=16, vector_quantization_hidden_dimension=16, num_quantizers=6,
                     num_semantic_quantizers=1, sliding_window=250, upsample_groups=32)
ORACLE_MIMI_SMALL = {"dimension": 32, "nfilters": 4, "ratios": [8, 6, 5, 4], "ksize": 7, "residual_ksize": 3, "last_ksize": 3, "compress": 2, "d_model": 32,
                     "num_heads": 4, "num_layers": 2, "dim_feedforward": 64, "context": 250, "max_period": 10000, "layer_scale": 0.01, "nq": 6, "bins": 64,
                     "qdim": 16, "upsample_stride": 2, "valid_num_quantizers": 16}


def qwen3_tokenizer_encoder_hf():
    """The ENCODER half of ``speech_tokenizer/model.safetensors``: transformers' MimiModel state-dict names under an ``encoder.`` prefix
    (speech_tokenizer.py:1253-1388 maps them onto Mimi modules).  Names and shapes come from a small transformers MimiModel."""
    import transformers
    m = transformers.MimiModel(transformers.MimiConfig(**HF_MIMI_SMALL))
    keep = ("encoder.", "encoder_transformer.", "downsample.", "quantizer.")
    return _fill([("encoder." + k, tuple(v.shape)) for k, v in m.state_dict().items() if k.startswith(keep)])


def map_hf_mimi_encoder(weights):
    """The encoder branch of Qwen3TTSSpeechTokenizer.sanitize (speech_tokenizer.py:1253-1415), restated for tests: SEANet layer indices ->
    init / residual / downsample / final convs, q/k/v -> one in_proj, (out, in, K) -> (out, K, in), code books kept as sum + usage."""
    import re
    conv_map = {0: "encoder_model.encoder.init_conv1d", 3: "encoder_model.encoder.layers.0.downsample", 6: "encoder_model.encoder.layers.1.downsample",
                9: "encoder_model.encoder.layers.2.downsample", 12: "encoder_model.encoder.layers.3.downsample", 14: "encoder_model.encoder.final_conv1d"}
    res_map, blk_map = {1: 0, 4: 1, 7: 2, 10: 3}, {1: 0, 3: 1}
    tr = {"self_attn.o_proj.weight": "self_attn.out_proj.weight", "mlp.fc1.weight": "gating.linear1.weight", "mlp.fc2.weight": "gating.linear2.weight",
          "input_layernorm.weight": "norm1.weight", "input_layernorm.bias": "norm1.bias", "post_attention_layernorm.weight": "norm2.weight",
          "post_attention_layernorm.bias": "norm2.bias", "self_attn_layer_scale.scale": "layer_scale_1.scale", "mlp_layer_scale.scale": "layer_scale_2.scale"}
    out, qkv = {}, {}
    sw = lambda v: v.swapaxes(-1, -2) if v.ndim == 3 else v                                     # noqa: E731
    for k, v in weights.items():
        if not k.startswith("encoder."):
            continue
        parts = k.split(".")
        if k.startswith("encoder.encoder.layers."):
            n = int(parts[3])
            if "block" in k:
                if n not in res_map or int(parts[5]) not in blk_map:
                    continue
                base, suffix = f"encoder_model.encoder.layers.{res_map[n]}.residuals.0.block.{blk_map[int(parts[5])]}", ".".join(parts[6:])
            else:
                if n not in conv_map:
                    continue
                base, suffix = conv_map[n], ".".join(parts[4:])
            out[f"{base}.conv.{suffix}"] = sw(v) if "weight" in suffix else v
        elif k.startswith("encoder.encoder_transformer.layers."):
            li, rest = int(parts[3]), ".".join(parts[4:])
            pre = f"encoder_model.encoder_transformer.transformer.layers.{li}."
            m = re.match(r"self_attn\.([qkv])_proj\.weight", rest)
            if m:
                qkv.setdefault(li, {})[m.group(1)] = v
            elif rest in tr:
                out[pre + tr[rest]] = v
        elif k.startswith("encoder.downsample."):
            suffix = k.replace("encoder.downsample.", "")
            out[f"encoder_model.downsample.conv.conv.{suffix}"] = sw(v) if "weight" in suffix else v
        elif k.startswith("encoder.quantizer."):
            rest = k.replace("encoder.quantizer.", "")
            which = "rvq_first" if "semantic_residual_vector_quantizer" in rest else "rvq_rest"
            m = re.search(r"layers\.(\d+)\.codebook\.(cluster_usage|embed_sum)", rest)
            if m:
                out[f"encoder_model.quantizer.{which}.vq.layers.{m.group(1)}.codebook.{'embedding_sum' if m.group(2) == 'embed_sum' else 'cluster_usage'}"] = v
            elif "input_proj.weight" in rest or "output_proj.weight" in rest:
                out[f"encoder_model.quantizer.{which}.{'input_proj' if 'input_proj' in rest else 'output_proj'}.weight"] = sw(v)
    for li, d in qkv.items():
        if len(d) == 3:
            cat = __import__("numpy").concatenate if not hasattr(d["q"], "dim") else __import__("torch").cat
            out[f"encoder_model.encoder_transformer.transformer.layers.{li}.self_attn.in_proj.weight"] = cat([d["q"], d["k"], d["v"]], 0)
    return out


def map_hf_mimi_decoder(sd, heads, head_dim):
    """transformers MimiModel (decode side) -> the reference's Mimi module tree (codec/models/mimi), for the independent cross-check of the
    oracle's ``mimi_decode``.  The reference's Mimi rotates interleaved pairs (nn.RoPE(traditional=True), kyutai's convention) while
    transformers rotates half-split pairs of permuted projections: q / k rows are re-interleaved per head so that both compute the same
    attention.  torch tensors in, torch tensors out."""
    import re
    import torch

    def interleave(w):
        w = w.reshape(heads, head_dim, -1)
        out = torch.empty_like(w)
        out[:, 0::2], out[:, 1::2] = w[:, : head_dim // 2], w[:, head_dim // 2:]
        return out.reshape(heads * head_dim, -1)
    tr = {"self_attn.o_proj.weight": "self_attn.out_proj.weight", "mlp.fc1.weight": "gating.linear1.weight", "mlp.fc2.weight": "gating.linear2.weight",
          "input_layernorm.weight": "norm1.weight", "input_layernorm.bias": "norm1.bias", "post_attention_layernorm.weight": "norm2.weight",
          "post_attention_layernorm.bias": "norm2.bias", "self_attn_layer_scale.scale": "layer_scale_1.scale", "mlp_layer_scale.scale": "layer_scale_2.scale"}
    P, qkv = {}, {}
    sw = lambda v: v.transpose(-1, -2) if v.dim() == 3 else v                                  # noqa: E731
    for k, v in sd.items():
        p = k.split(".")
        if k.startswith("decoder.layers."):
            n = int(p[2])
            if n == 0:
                P["decoder.init_conv1d.conv.conv." + p[-1]] = sw(v)
            elif n == 14:
                P["decoder.final_conv1d.conv.conv." + p[-1]] = sw(v)
            elif n in (2, 5, 8, 11):
                P[f"decoder.layers.{(n - 2) // 3}.upsample.convtr.convtr." + p[-1]] = v.permute(1, 2, 0) if v.dim() == 3 else v
            else:
                P[f"decoder.layers.{(n - 3) // 3}.residuals.0.block.{ {1: 0, 3: 1}[int(p[4])] }.conv.conv." + p[-1]] = sw(v)
        elif k.startswith("decoder_transformer.layers."):
            li, rest = int(p[2]), ".".join(p[3:])
            m = re.match(r"self_attn\.([qkv])_proj\.weight", rest)
            if m:
                qkv.setdefault(li, {})[m.group(1)] = v
            elif rest in tr:
                P[f"decoder_transformer.transformer.layers.{li}." + tr[rest]] = v
        elif k == "upsample.conv.weight":
            P["upsample.convtr.convtr.convtr.weight"] = v.permute(0, 2, 1)
        elif k.startswith("quantizer."):
            which = "rvq_first" if "semantic" in k else "rvq_rest"
            m = re.search(r"layers\.(\d+)\.codebook\.(cluster_usage|embed_sum)", k)
            if m:
                P[f"quantizer.{which}.vq.layers.{m.group(1)}.codebook." + ("embedding_sum" if m.group(2) == "embed_sum" else "cluster_usage")] = v
            elif "output_proj.weight" in k:
                P[f"quantizer.{which}.output_proj.weight"] = sw(v)
    for li, d in qkv.items():
        P[f"decoder_transformer.transformer.layers.{li}.self_attn.in_proj.weight"] = torch.cat([interleave(d["q"]), interleave(d["k"]), d["v"]], 0)
    return P
