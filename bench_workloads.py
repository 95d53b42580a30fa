"""bench.py --workload whisper | codec | qwen3: the other BASELINE.json configurations under the same JSON contract as the Kokoro headline
(bench.py): device-resident `value`, host-buffer `e2e`, `roofline`, `cpu_baseline`, clocks, launches; `--impl reference` times the
CPU restatement of the reference (oracle/, rank 0 only, bounded sample).

  whisper  config 3: Whisper-small log-mel + encoder, 32 x 30 s windows per GPU (weak scaling: every rank encodes its own batch).
           Roofline: tensor pipe -- 345 GFLOP per window / time of the step.
  codec    config 5: ONE 10 000-frame SNAC-24k code stream (213 s of audio) decoded by all ranks: contiguous frame spans + the exact halo,
           one trailing NCCL all_gather of the waveform pieces (strong scaling; `parallel.decode_stream_sharded`).  Mimi (800 s) beside it.
           Roofline: HBM -- 27.65 GB per 10 000 SNAC frames (SURVEY.md section 8d, bf16 convention).
  qwen3    config 4: 64 utterances x 40 frames of Qwen3-TTS-0.6B batch-sharded over the ranks (strong scaling), frame loop in CUDA graphs +
           the 12.5 Hz vocoder.  Roofline: HBM -- 3.3 GB of weights streamed per frame and batch.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0}, "fallback"


class _Harness:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args, self.rank, self.world = torch, dist, args, rank, world
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.W, self.K = max(args.warmup, 3), args.steps

    def log(self, m):
        print(f"[bench r{self.rank} {time.strftime('%H:%M:%S')}] {m}", file=sys.stderr, flush=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def rank_max(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, flush=True):
        """K steps between barrier + synchronize brackets, CUDA events, max over ranks; the 256 MiB L2 flush between steps is timed
        separately and subtracted.  Returns (ms per step, last output)."""
        torch = self.torch
        for _ in range(self.W):
            out = fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(self.K):
            if flush:
                self.flush.zero_()
            out = fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if flush:
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(self.K):
                self.flush.zero_()
            f1.record()
            torch.cuda.synchronize(self.dev)
            ms -= f0.elapsed_time(f1)
        return self.rank_max(max(ms, 1e-6)) / self.K, out

    def wall(self, fn):
        for _ in range(self.W):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(self.K):
            fn()
        self.barrier()
        return self.rank_max(time.perf_counter() - t0) / self.K

    def finish(self, line):
        if self.rank == 0:
            print(json.dumps(line), flush=True)
        if self.world > 1:
            self.dist.destroy_process_group()


def _profile(ops, torch, fn, dev):
    prof = {}
    ops.PROFILE = prof
    fn()
    torch.cuda.synchronize(dev)
    ops.PROFILE = None
    return {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in prof.items()}, sum(len(v) for v in prof.values())


# ------------------------------------------------------------------------------------------------------------------------- whisper
WHISPER_BATCH = 32
WHISPER_GFLOP_PER_WINDOW = 345.0


def _whisper(args, rank, world, local_rank, top):
    H = _Harness(args, rank, world, local_rank)
    torch = H.torch
    from mlx_audio_b200 import ops, synth
    from mlx_audio_b200.configs import WHISPER_SMALL
    from mlx_audio_b200.stt.models.whisper import Model, ModelDimensions
    model = Model(ModelDimensions.from_dict(WHISPER_SMALL), device=H.dev).load_weights(synth.whisper_encoder_weights(WHISPER_SMALL))
    audio_h = synth.whisper_audio(WHISPER_BATCH, seed=4 + rank).pin_memory()
    audio_d = audio_h.to(H.dev)
    sampler = top.ClockSampler(local_rank)
    n0 = ops.LAUNCHES[0]
    sampler.start()
    ms, y = H.timed(lambda: model.encode_audio(audio_d))
    clocks = sampler.stop()
    launches = ops.LAUNCHES[0] - n0
    assert y.shape == (WHISPER_BATCH, 1500, 768) and bool(torch.isfinite(y).all())
    out_h = torch.empty(WHISPER_BATCH, 1500, 768).pin_memory()

    def e2e():
        out_h.copy_(model.encode_audio(audio_h.to(H.dev, non_blocking=True)), non_blocking=True)
        torch.cuda.current_stream(H.dev).synchronize()
    e2e_s = H.wall(e2e)
    by_kind, n_l = _profile(ops, torch, lambda: model.encode_audio(audio_d), H.dev)
    peaks, pk = _peaks()
    secs = 30.0 * WHISPER_BATCH
    tf = WHISPER_GFLOP_PER_WINDOW * 1e9 * WHISPER_BATCH / (ms / 1e3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    line = {"metric": "audio-sec/sec Whisper-small log-mel + encoder (batch 32 x 30 s)", "value": world * secs / (ms / 1e3), "unit": "audio-s/s",
            "n_gpus": world, "steps": H.K, "warmup": H.W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16x2" if ops.TC_MODE[0] == "x2" else "fp16", "data": "synthetic",
            "config": {"workload": "whisper-small cfg3: 32 x 30 s windows -> log-mel (+30 s zero pad) -> 12-layer encoder, per GPU",
                       "parallelism": f"window batches sharded x{world} (no data-path collective)", "l2": "256 MiB flush between timed steps (its cost subtracted)",
                       "weights": "synthetic fp16 checkpoint", "kernels_per_step": n_l},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": world * secs / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": int(audio_h.numel() * 4), "d2h_bytes_per_step": int(out_h.numel() * 4),
                    "call": "Model.encode_audio(pinned host audio) -> pinned host features"},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None, "peak_kind": pk,
                         "kernel": "encoder GEMMs + attention (tcgen05)", "ms_by_kind": by_kind,
                         "algorithmic_flops_per_step": WHISPER_GFLOP_PER_WINDOW * 1e9 * WHISPER_BATCH, "note": "one 16-bit product per MAC; the x2 mode issues two"}}
    if rank == 0 and world == 1 and args.cpu_utts > 0:
        line["cpu_baseline"] = _whisper_cpu(1, top.host_threads())
    H.finish(line)


def _whisper_cpu(n_windows, threads):
    import torch
    from mlx_audio_b200 import synth
    from mlx_audio_b200.configs import WHISPER_SMALL
    from oracle import dsp as OD
    from oracle import whisper as OW
    torch.set_num_threads(threads)
    P = {k: v.float() for k, v in synth.whisper_encoder_weights(WHISPER_SMALL).items()}
    audio = synth.whisper_audio(n_windows)
    t0 = time.perf_counter()
    for a in audio:
        mel = torch.as_tensor(OD.whisper_log_mel(a.numpy(), 80, padding=480000)[:3000]).float()
        OW.encoder(P, mel[None], WHISPER_SMALL)
    dt = time.perf_counter() - t0
    return {"value": 30.0 * n_windows / dt, "unit": "audio-s/s", "cores": threads, "kind": "port",
            "sample": f"{n_windows} window(s) of 30 s ({dt:.1f} s), torch-CPU fp32 restatement of the reference"}


# ------------------------------------------------------------------------------------------------------------------------- codec
CODEC_FRAMES = 10000
SNAC_BYTES_PER_10K = 27648.4e6


def _codec(args, rank, world, local_rank, top):
    H = _Harness(args, rank, world, local_rank)
    torch = H.torch
    from mlx_audio_b200 import ops, synth
    from mlx_audio_b200.codec import SNAC, Mimi, mimi_202407
    from mlx_audio_b200.configs import MIMI_202407, SNAC_24K
    from mlx_audio_b200.parallel import decode_stream_sharded
    T = CODEC_FRAMES // 4 * 4
    snac = SNAC.from_config(SNAC_24K, device=H.dev).load_weights(synth.snac_weights(SNAC_24K))
    codes_h = [c.pin_memory() for c in synth.snac_codes(SNAC_24K, T)]
    codes = [c.to(H.dev) for c in codes_h]
    noises = [n.to(H.dev) for n in synth.snac_noises(SNAC_24K)]
    sampler = top.ClockSampler(local_rank)
    n0 = ops.LAUNCHES[0]
    sampler.start()
    ms, y = H.timed(lambda: decode_stream_sharded(snac, codes, noises=noises))
    clocks = sampler.stop()
    launches = ops.LAUNCHES[0] - n0
    n_samples = T * 512 + 75
    if rank == 0:
        assert y.shape == (n_samples,) and bool(torch.isfinite(y).all())
    secs = n_samples / 24000.0
    out_h = torch.empty(n_samples).pin_memory() if rank == 0 else None

    def e2e():
        w = decode_stream_sharded(snac, [c.to(H.dev, non_blocking=True) for c in codes_h], noises=noises)
        if rank == 0:
            out_h.copy_(w, non_blocking=True)
        torch.cuda.current_stream(H.dev).synchronize()
    e2e_s = H.wall(e2e)
    # compute-only share of the step on this rank (span decode without the gather) and Mimi beside it
    from mlx_audio_b200.parallel import shard_span
    _, _, cs, ce = shard_span(T, rank, world, multiple=4)
    ms_local, _ = H.timed(lambda: snac.decode_span(codes, cs, ce, noises=noises))
    by_kind, n_l = _profile(ops, torch, lambda: snac.decode_span(codes, cs, ce, noises=noises), H.dev)
    mimi = Mimi(mimi_202407(32), device=H.dev).load_weights(synth.mimi_weights(MIMI_202407))
    mcodes = synth.mimi_codes(MIMI_202407, CODEC_FRAMES).to(H.dev)
    ms_mimi, ym = H.timed(lambda: decode_stream_sharded(mimi, mcodes))
    peaks, pk = _peaks()
    alg = SNAC_BYTES_PER_10K * T / 10000
    ach = alg / (ms / 1e3) / 1e9 / world                # per-GPU achieved bandwidth: the stream's algorithmic bytes are split over the ranks
    line = {"metric": "audio-sec/sec SNAC-24k decode of one 10k-frame stream (213 s)", "value": secs / (ms / 1e3), "unit": "audio-s/s", "n_gpus": world,
            "steps": H.K, "warmup": H.W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16x2(x3 weights)" if ops.TC_MODE[0] == "x2" else "bf16", "data": "synthetic",
            "config": {"workload": "codec cfg5: ONE SNAC-24k stream of 10 000 finest-level frames (5.12 M samples); Mimi 10 000 frames (19.2 M samples) beside it",
                       "parallelism": f"frame spans x{world} with a {snac.SPAN_HALO}-frame halo per side; one trailing all_gather of the pieces (NCCL)",
                       "l2": "256 MiB flush between timed steps (its cost subtracted)", "weights": "synthetic fp32 checkpoint", "kernels_per_step": n_l},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": secs / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": int(sum(c.numel() for c in codes_h) * 8), "d2h_bytes_per_step": int(n_samples * 4),
                    "call": "parallel.decode_stream_sharded(SNAC, pinned host codes) -> pinned host waveform on rank 0"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_kind": pk,
                         "kernel": "SNAC decoder conv stack (tcgen05 1x1 / transposed convs + staged depthwise convs)", "ms_by_kind": by_kind,
                         "ms_span_decode_this_rank": ms_local, "algorithmic_bytes_per_step": alg},
            "mimi": {"value": CODEC_FRAMES * 1920 / 24000.0 / (ms_mimi / 1e3), "unit": "audio-s/s", "ms_per_step": ms_mimi,
                     "halo_frames": mimi.span_halo, "note": "causal stack: left halo = num_layers x context transformer positions + convs"}}
    if rank == 0 and world == 1 and args.cpu_utts > 0:
        line["cpu_baseline"] = _codec_cpu(top.host_threads())
    H.finish(line)


def _codec_cpu(threads, frames=256):
    import torch
    from mlx_audio_b200 import synth
    from mlx_audio_b200.configs import SNAC_24K
    from oracle import codec as OC
    torch.set_num_threads(threads)
    P = {k: v.float() for k, v in synth.snac_weights(SNAC_24K).items()}
    codes = synth.snac_codes(SNAC_24K, frames)
    noises = [n.float() for n in synth.snac_noises(SNAC_24K)]
    t0 = time.perf_counter()
    y = OC.snac_decode(P, codes, noises=noises)
    dt = time.perf_counter() - t0
    return {"value": y.shape[1] / 24000.0 / dt, "unit": "audio-s/s", "cores": threads, "kind": "port",
            "sample": f"{frames} frames ({y.shape[1] / 24000.0:.1f} s of audio, {dt:.1f} s), torch-CPU fp32 restatement of the reference"}


# ------------------------------------------------------------------------------------------------------------------------- qwen3
QWEN3_UTTS, QWEN3_FRAMES = 64, 40
QWEN3_WEIGHT_GB_PER_FRAME = 3.3


def _qwen3(args, rank, world, local_rank, top):
    H = _Harness(args, rank, world, local_rank)
    torch = H.torch
    from mlx_audio_b200 import ops, synth
    from mlx_audio_b200.configs import QWEN3_TALKER, QWEN3_TOKENIZER_DECODER
    from mlx_audio_b200.parallel import shard_units
    from mlx_audio_b200.tts.models.qwen3_tts import Model, ModelConfig, Qwen3TTSSpeechTokenizer, Qwen3TTSTalkerConfig, Qwen3TTSTalkerCodePredictorConfig, Qwen3TTSTokenizerConfig, Qwen3TTSTokenizerDecoderConfig
    flat = dict(QWEN3_TALKER)
    P = synth.qwen3_talker_weights(flat, seed=11)
    cp = Qwen3TTSTalkerCodePredictorConfig(num_hidden_layers=flat["cp_num_hidden_layers"])
    tc = Qwen3TTSTalkerConfig(code_predictor_config=cp, num_hidden_layers=flat["num_hidden_layers"], text_vocab_size=512, codec_eos_token_id=flat["codec_eos_token_id"])
    model = Model(ModelConfig(talker_config=tc, tts_pad_token_id=500, tts_bos_token_id=501, tts_eos_token_id=502), H.dev).load_weights(P)
    st = Qwen3TTSSpeechTokenizer(Qwen3TTSTokenizerConfig(decoder_config=Qwen3TTSTokenizerDecoderConfig()), device=H.dev)
    st.load_weights(synth.qwen3_tokenizer_weights(dict(QWEN3_TOKENIZER_DECODER), seed=12))
    model.load_speech_tokenizer(st)
    g = torch.Generator().manual_seed(3)
    lens = [int(v) for v in torch.randint(10, 30, (QWEN3_UTTS,), generator=g)]
    prompts = [torch.randint(0, 500, (n,), generator=g).tolist() for n in lens]
    mine = shard_units(lens, rank, world)
    bs = 8
    groups = [mine[i:i + bs] for i in range(0, len(mine), bs)]

    def step():
        audio = []
        for grp in groups:
            res = list(model.batch_generate_from_ids([prompts[i] for i in grp], language_id=2050, max_tokens=QWEN3_FRAMES, seed=1 + grp[0], stop_on_eos=False))
            audio.extend(r.audio for r in res)
        return audio
    sampler = top.ClockSampler(local_rank)
    n0 = ops.LAUNCHES[0]
    sampler.start()
    ms, audio = H.timed(step, flush=False)
    clocks = sampler.stop()
    launches = ops.LAUNCHES[0] - n0
    frames_total = QWEN3_UTTS * QWEN3_FRAMES
    secs = frames_total / 12.5
    assert len(audio) == len(mine) and all(bool(torch.isfinite(a).all()) for a in audio)
    peaks, pk = _peaks()
    steps_per_rank = len(groups) * QWEN3_FRAMES
    ach = QWEN3_WEIGHT_GB_PER_FRAME * steps_per_rank / (ms / 1e3)
    line = {"metric": "audio-sec/sec Qwen3-TTS-0.6B, 64 utterances x 40 frames", "value": secs / (ms / 1e3), "unit": "audio-s/s", "n_gpus": world, "steps": H.K,
            "warmup": H.W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "qwen3-tts-0.6b cfg4: 64 prompts (10-30 text tokens), 40 code frames each (3.2 s), batch-sharded, batches of 8 per CUDA-graph frame loop, then the 12.5 Hz vocoder",
                       "parallelism": f"utterances sharded x{world} by length (no data-path collective)", "l2": "no flush: every frame streams 3.3 GB of weights (> L2)",
                       "weights": "synthetic bf16 checkpoint, 28 + 5 layers", "frames_per_s": frames_total / (ms / 1e3)},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": secs / (ms / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": int(sum(lens) * 8), "d2h_bytes_per_step": int(frames_total * 1920 * 4 // max(world, 1)),
                    "call": "Model.batch_generate_from_ids(host token ids) -> GenerationResult.audio (same timed region: inputs are host lists)"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_kind": pk,
                         "kernel": "decode GEMVs of the talker + code predictor (weights streamed once per frame and batch)",
                         "algorithmic_bytes_per_step": QWEN3_WEIGHT_GB_PER_FRAME * 1e9 * steps_per_rank}}
    H.finish(line)


def _reference(args, rank, world, top):
    if rank != 0:
        return
    cores = top.host_threads()
    if args.workload == "whisper":
        cb = _whisper_cpu(max(1, min(args.steps, 2)), cores)
        metric, wl = "audio-sec/sec Whisper-small log-mel + encoder (batch 32 x 30 s)", "whisper-small cfg3 (bounded sample: single windows)"
    elif args.workload == "codec":
        cb = _codec_cpu(cores)
        metric, wl = "audio-sec/sec SNAC-24k decode of one 10k-frame stream (213 s)", "codec cfg5 (bounded sample: 256 frames)"
    else:
        print(json.dumps({"impl": "reference", "unavailable": "the float64 Qwen3-TTS oracle needs minutes per frame batch; no bounded CPU sample is meaningful"}), flush=True)
        return
    line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": cb["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "parallelism": "cpu"}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main(args, rank, world, local_rank):
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200_bench_top", os.path.join(ROOT, "bench.py"))
    top = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(top)
    if args.impl == "reference":
        return _reference(args, rank, world, top)
    return {"whisper": _whisper, "codec": _codec, "qwen3": _qwen3}[args.workload](args, rank, world, local_rank, top)
