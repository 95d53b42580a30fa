#!/usr/bin/env python
"""bench.py -- Kokoro-82M TTS audio-seconds synthesised per wall-second on N B200s (BASELINE.json metric).

Workload (BASELINE.json configs[1]): one 128-phoneme utterance (T = 130 tokens
with BOS/EOS), durations pinned to 3 frames/token -> F = 390 frames -> 234 000 samples = 9.75 s of 24 kHz
audio, synthetic bf16 checkpoint at the real Kokoro-82M shapes (81.8 M parameters), SineGen noise drawn
on device every step.  A "step" = one utterance per GPU (weak scaling: every rank synthesises its own).

  value      device-resident inputs, CUDA-graph replay of the whole utterance, K steps between
             barrier+synchronize brackets, CUDA events, max over ranks.
  e2e        the public API path with HOST buffers: pinned ids/style -> device, replay, waveform -> pinned host.
  roofline   the dominant kernel family (dense conv1d / transposed conv of the decoder+generator stack):
             algorithmic bytes per utterance (137.2 MB per audio-second, bf16
             convention) / summed CUDA-event time of those launches, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle port (torch-CPU fp32 restatement of the reference; MLX is not installable) on
             the host cores, bounded sample.
`--impl reference` times that same CPU restatement as the reference arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AUDIO_S_PER_UTT = 9.75
N_PHONEMES, DUR = 128, 3
CONV_STACK_MB_PER_AUDIO_S = 137.2          # SURVEY.md section 8(d), Kokoro decoder+generator, bf16 convention
METRIC = "audio-sec/sec Kokoro-82M TTS (128-phoneme utterance, 9.75 s)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_run(n_utts: int, threads: int):
    """Time the oracle port (fp32 torch-CPU restatement of the reference) on `n_utts` cfg2 utterances."""
    import torch
    from mlx_audio_b200 import synth
    from oracle import kokoro as OK
    torch.set_num_threads(threads)
    print(f"[bench] cpu port: {n_utts} utterance(s) on {threads} threads", file=sys.stderr, flush=True)
    P = synth.kokoro_weights(OK.KOKORO_CONFIG)
    ids, ref = synth.kokoro_inputs(N_PHONEMES)
    T = ids.shape[1]
    _, nz = synth.kokoro_noise(T * DUR * 600)
    times = []
    for _ in range(n_utts):
        t0 = time.perf_counter()
        audio, _ = OK.forward(P, ids, ref, noise=nz, pred_dur_override=[DUR] * T)
        times.append(time.perf_counter() - t0)
    assert audio.shape[0] == int(AUDIO_S_PER_UTT * 24000)
    return times


def host_threads() -> int:
    """Threads the CPU arm may use: the cores this process is actually allowed to run on (cgroup / affinity aware), capped
    at 32 -- the restatement's small ops stop scaling long before that."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, p = f.read().split()
            if q != "max":
                n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    return max(1, min(n, 32))


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path, here its restatement (MLX cannot be installed)."""
    if rank != 0:
        return
    cores = host_threads()
    cpu_port_run(1, cores)           # warm-up (bounded: one utterance)
    times = cpu_port_run(args.steps, cores)
    total = sum(times)
    v = AUDIO_S_PER_UTT * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "kokoro-82m cfg2: 128 phonemes, F=390 frames, 9.75 s audio per step", "parallelism": "cpu"},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": f"{len(times)} full cfg2 utterances, torch-CPU fp32 restatement of the reference (oracle/kokoro.py)"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--cpu-utts", type=int, default=4, help="utterances in the cpu_baseline sample")
    ap.add_argument("--ncu", action="store_true", help="profiling aid: 2 eager warm-up steps, then ONE eager step between "
                    "cudaProfilerStart/Stop (run under `ncu --profile-from-start off`); prints no bench line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from mlx_audio_b200 import ops, synth
    from mlx_audio_b200.tts.models.kokoro import Model, ModelConfig
    from mlx_audio_b200.tts.models.kokoro.kokoro import CapturedUtterance
    from oracle.kokoro import KOKORO_CONFIG            # config dict only (data); the oracle itself runs in cpu_baseline

    W = max(args.warmup, 3)
    K = args.steps
    log = lambda m: print(f"[bench r{rank} {time.strftime('%H:%M:%S')}] {m}", file=sys.stderr, flush=True)
    log("building synthetic checkpoint")
    P = synth.kokoro_weights(KOKORO_CONFIG, seed=0)
    model = Model(ModelConfig.from_dict(KOKORO_CONFIG), device=dev).load_weights(list(P.items()))
    ids, ref_s = synth.kokoro_inputs(N_PHONEMES, seed=1 + rank)
    T = ids.shape[1]
    F = T * DUR
    n_samples = F * 600
    ids_d, ref_d = ids[0].to(dev), ref_s.to(dev)
    dur_d = torch.full((T,), DUR, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)            # > 126 MB L2

    log("model ready; capturing")
    cap = CapturedUtterance(model, T, F, seed=1234 + rank)
    cap.set_inputs(ids_d, ref_d, dur_d)
    noise_buf = torch.empty(1, n_samples, 9, device=dev)

    def step_eager():
        ops.randn_(noise_buf, 1234 + rank, 0)
        return model.forward_ids(ids_d, ref_d, noise=noise_buf, pred_dur=dur_d, n_frames=F)[0]

    if args.ncu:
        for _ in range(2):
            step_eager()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        step_eager()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return
    if args.no_graph:
        step = step_eager
        launches_per_step = None
    else:
        cap.capture()
        step = cap.replay
        launches_per_step = cap.launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    log("timing device-resident steps")
    # ---------------- device-resident timing
    for _ in range(W):
        step()
    n0 = ops.LAUNCHES[0]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        flush.zero_()                                                          # L2 flush between timed iterations
        audio = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = ops.LAUNCHES[0] - n0
    assert audio.shape[0] == n_samples and bool(torch.isfinite(audio).all())
    # L2 flush cost measured separately and subtracted (it is not part of the step)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    f0.record()
    for _ in range(K):
        flush.zero_()
    f1.record()
    torch.cuda.synchronize(dev)
    ms_flush = f0.elapsed_time(f1)
    ms_net = max(ms - ms_flush, 1e-6)
    t = torch.tensor([ms_net], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * AUDIO_S_PER_UTT / (ms_max / 1e3)

    log(f"value done: {ms_max / K:.3f} ms/step; timing e2e")
    # ---------------- end-to-end through the public call with host buffers
    ids_h = ids[0].clone().pin_memory()
    ref_h = ref_s.clone().pin_memory()
    out_h = torch.empty(n_samples, dtype=torch.float32).pin_memory()

    def step_e2e():
        if args.no_graph:
            a = model.forward_ids(ids_h.to(dev, non_blocking=True), ref_h.to(dev, non_blocking=True), noise=ops.randn_(noise_buf, 7, 0),
                                  pred_dur=dur_d, n_frames=F)[0]
        else:
            cap.set_inputs(ids_h, ref_h)
            a = cap.replay()
        out_h.copy_(a, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()                            # the caller owns the waveform when this returns
        return out_h

    for _ in range(W):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        flush.zero_()
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0 - ms_flush / 1e3
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K * AUDIO_S_PER_UTT / float(t.item())

    log("e2e done; instrumented pass")
    # ---------------- roofline of the dominant kernel family (instrumented eager pass, CUDA events per launch)
    prof = {"conv": [], "other": []}
    ops.PROFILE = prof
    nprof = 3
    for _ in range(nprof):
        step_eager()
    torch.cuda.synchronize(dev)
    ops.PROFILE = None
    by_kind = {k: sum(a.elapsed_time(b) for a, b in v) / nprof for k, v in prof.items()}
    conv_kinds = ("conv", "conv_tc", "prep")                   # the dense conv stack: CUDA-core convs, tcgen05 convs + their bf16 prologue
    conv_ms = sum(by_kind.get(k, 0.0) for k in conv_kinds)
    other_ms = sum(v for k, v in by_kind.items() if k not in conv_kinds)
    n_conv = sum(len(prof.get(k, [])) for k in conv_kinds) // nprof
    peaks, peak_kind = _peaks()
    alg_bytes = CONV_STACK_MB_PER_AUDIO_S * 1e6 * AUDIO_S_PER_UTT                # per utterance, all conv launches
    achieved = alg_bytes / (conv_ms / 1e3) / 1e9
    traffic = None                                                               # measured DRAM bytes of the same kernels (ncu capture, committed)
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01b_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = float(json.load(open(tpath))["conv_stack_bytes_per_step"])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_kind": peak_kind, "kernel": "dense conv stack: conv_tc_persist_kernel (tcgen05, persistent) + prep_bf16_kernel, conv1d_dense/convtr1d_dense (CUDA-core)",
                "launches_per_utterance": n_conv, "kernel_ms_per_utterance": conv_ms, "other_kernels_ms_per_utterance": other_ms,
                "ms_by_kind": {k: round(v, 3) for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1])},
                "algorithmic_bytes_per_utterance": alg_bytes}

    log("instrumented pass done")
    if rank == 0:
        cores = host_threads()
        cpu = None
        if world == 1 and args.cpu_utts > 0:
            times = cpu_port_run(args.cpu_utts, cores)
            cpu = {"value": AUDIO_S_PER_UTT * len(times) / sum(times), "unit": "audio-s/s", "cores": cores, "kind": "port",
                   "sample": f"{len(times)} full cfg2 utterances ({sum(times):.1f} s), torch-CPU fp32 restatement of the reference"}
        line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if ops.TC_MODE[0] == "off" else ("bf16x2" if ops.TC_MODE[0] == "x2" else "bf16"), "data": "synthetic",
                "config": {"workload": "kokoro-82m cfg2: 128 phonemes (T=130), durations pinned to 3 -> F=390 frames, 234000 samples = 9.75 s per step per GPU",
                           "parallelism": f"utterance-sharded x{world} (no data-path collective)", "l2": "256 MiB flush between timed steps (its cost subtracted)",
                           "launch": "eager" if args.no_graph else "cuda-graph replay", "weights": "synthetic bf16 checkpoint, 81.8 M params",
                           "activations": "fp32 in HBM; tensor-core mode " + ops.TC_MODE[0] + " (x2 = hi+lo bf16 planes, fp32-grade products)"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(ids_h.numel() * 8 + ref_h.numel() * 4),
                        "d2h_bytes_per_step": int(out_h.numel() * 4)},
                "roofline": roofline}
        if cpu:
            line["cpu_baseline"] = cpu
        if launches_per_step:
            line["config"]["kernels_per_step"] = launches_per_step
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
