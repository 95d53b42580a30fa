#!/usr/bin/env python
"""BASELINE config 4 (Qwen3-TTS): frame rate of the talker + code-predictor loop (CUDA-graph replay per frame) against the
HBM floor of streaming the bf16 weights once per frame, and the 12.5 Hz vocoder's throughput on 300-frame chunks."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_audio_b200 import ops, synth
from mlx_audio_b200.tts.models.qwen3_tts import (Model, ModelConfig, Qwen3TTSSpeechTokenizer, Qwen3TTSTalkerConfig, Qwen3TTSTokenizerConfig)
from mlx_audio_b200.configs import QWEN3_TALKER as TALKER, QWEN3_TOKENIZER_DECODER as TOKENIZER_DECODER

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--batches", default="1,8")
ap.add_argument("--voc-frames", type=int, default=3000)
ap.add_argument("--which", default="talker,vocoder")
ap.add_argument("--eager", action="store_true", help="no CUDA graph (for ncu launch lists)")
args = ap.parse_args()
dev = torch.device("cuda:0")
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {}
hbm = float(peaks.get("hbm_gbs", peaks.get("hbm_gbs_burst", 6650.0))) if isinstance(peaks, dict) else 6650.0

if "talker" in args.which:
    P = synth.qwen3_talker_weights(dict(TALKER))
    model = Model(ModelConfig(talker_config=Qwen3TTSTalkerConfig(text_vocab_size=512), tts_pad_token_id=500, tts_bos_token_id=501,
                              tts_eos_token_id=502), dev).load_weights(P)
    frame_params = sum(v.numel() for k, v in P.items() if (".layers." in k or "codec_head" in k or "lm_head" in k) and "code_predictor" not in k)
    cp_params = sum(v.numel() for k, v in P.items() if "code_predictor.model.layers" in k) * 15 + sum(v.numel() for k, v in P.items() if "lm_head" in k)
    weight_bytes = 2 * (frame_params + cp_params)
    del P
    ids = torch.randint(0, 500, (40,), generator=torch.Generator().manual_seed(0)).tolist()
    x, trailing, pad = model.prepare_generation_inputs_from_ids(ids)
    model.config.talker_config.codec_eos_token_id = 3071        # suppressed id: never sampled, the loop runs to max_tokens
    for B in [int(b) for b in args.batches.split(",")]:
        xb = x.expand(B, -1, -1).contiguous()
        model.generate_codes(xb, trailing, pad, max_tokens=8, seed=1, use_graph=not args.eager)           # warm-up (loads kernels, captures once)
        torch.cuda.synchronize()
        ops.LAUNCHES[0] = 0
        t0 = time.perf_counter()
        codes = model.generate_codes(xb, trailing, pad, max_tokens=args.frames, seed=2, use_graph=not args.eager)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = codes.shape[1]
        ms = dt / n * 1e3
        print(json.dumps({"workload": f"qwen3-tts talker+code-predictor loop, batch {B}, {n} frames (prefill {x.shape[1]} rows, graph capture included)",
                          "ms_per_frame": ms, "frames_per_s": B * n / dt, "audio_s_per_s": B * n / dt / 12.5,
                          "weight_GB_per_frame": weight_bytes / 1e9, "hbm_floor_ms_per_frame": weight_bytes / (hbm * 1e9) * 1e3,
                          "frac_of_hbm_floor": weight_bytes / (hbm * 1e9) * 1e3 / ms, "launches_per_frame": ops.LAUNCHES[0] / n}))
    del model
    torch.cuda.empty_cache()

if "vocoder" in args.which:
    Pv = synth.qwen3_tokenizer_weights(dict(TOKENIZER_DECODER))
    st = Qwen3TTSSpeechTokenizer(Qwen3TTSTokenizerConfig(), dev).load_weights(Pv)
    codes = synth.qwen3_codes(TOKENIZER_DECODER, args.voc_frames).transpose(1, 2).to(dev)
    for _ in range(2):
        st.decode(codes)
    torch.cuda.synchronize()
    prof = {}
    ops.PROFILE = prof
    st.decode(codes)
    torch.cuda.synchronize()
    ops.PROFILE = None
    by_kind = {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in prof.items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        wav, _ = st.decode(codes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = 2.27e12 * args.voc_frames / 300.0           # 2.27 TFLOP per 300-frame chunk
    print(json.dumps({"workload": f"qwen3-tts vocoder chunked_decode, {args.voc_frames} frames ({args.voc_frames / 12.5:.0f} s of audio)", "ms": ms,
                      "audio_s_per_s": args.voc_frames / 12.5 / (ms / 1e3), "TFLOPs_algorithmic": flops / (ms / 1e3) / 1e12,
                      "tc_mode": ops.TC_MODE[0], "ms_by_kind": dict(sorted(by_kind.items(), key=lambda kv: -kv[1]))}))
