"""Host-side wrappers: torch CUDA tensors -> C-ABI calls on torch's current stream.

torch is plumbing only (device memory, streams); every compute op here is one of our sm_100a
kernels.  All activations are float32 channels-last ``[B, L, C]``; tensors may be channel-slice
views of wider buffers (row stride = ``stride(1)``), which is how concatenations are avoided.
"""
from __future__ import annotations

import ctypes as C
import math
import contextlib
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import AttnParams, Conv1dParams, ConvFParams

ACT = {"none": 0, "lrelu": 1, "snake": 2, "elu": 3, "gelu": 4, "gelu_tanh": 5, "tanh": 6, "sigmoid": 7,
       "silu": 8, "clip1": 9}

LAUNCHES = [0]        # kernels launched through this module (bench.py reports it as gpu_launches)


PROFILE = None        # dict kind -> [(start_event, end_event)] when bench.py instruments a pass
PROFILE_TAGS = None   # dict kind -> [label] parallel to PROFILE (bench.py's per-launch table); TAG[0] is the label of the next call
TAG = [None]
PROFILE_EXTERNAL = False   # True while an instrumented CUDA graph is captured: the events become event-record NODES of the graph, so
                           # every replay re-stamps them and the per-launch times are those of the replayed graph (not of an eager pass)


def _call(kind, fn, n_launches, *args):
    """Invoke a C-ABI entry point, map its status to an exception, count its kernel launches and (when
    bench.py asks) bracket it with CUDA events on the launching stream."""
    if PROFILE is not None:
        a = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        b = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        a.record()
        rc = fn(*args)
        b.record()
        PROFILE.setdefault(kind, []).append((a, b))
        if PROFILE_TAGS is not None:
            PROFILE_TAGS.setdefault(kind, []).append(TAG[0])
        TAG[0] = None
    else:
        rc = fn(*args)
    _lib.check(rc)
    LAUNCHES[0] += n_launches


_SIDE_POOL = {}      # device -> [streams];  _SIDE_BUSY: streams handed out by fork() and not yet joined
_SIDE_BUSY = set()


def fork(device, n: int = 1):
    """Start ``n`` concurrent branches: returns side streams that wait on the current stream's present point.  Use
    ``with torch.cuda.stream(s): ...`` for each branch, then ``join(device, streams)``.  Inside a CUDA-graph capture
    the branches become parallel graph paths.  Tensors that cross branches must be allocated BEFORE the fork (on the
    joining stream); branch temporaries stay stream-local, and because every fork waits on the forking stream and every
    join waits on the branches, block reuse by the caching allocator stays ordered."""
    device = torch.device(device)
    cur = torch.cuda.current_stream(device)
    pool = _SIDE_POOL.setdefault(device, [])
    out = []
    for s in pool:
        if len(out) == n:
            break
        if s not in _SIDE_BUSY and s != cur:
            out.append(s)
    while len(out) < n:
        s = torch.cuda.Stream(device=device)
        pool.append(s)
        out.append(s)
    for s in out:
        _SIDE_BUSY.add(s)
        s.wait_stream(cur)
    return out


def join(device, streams) -> None:
    cur = torch.cuda.current_stream(torch.device(device))
    for s in streams:
        cur.wait_stream(s)
        _SIDE_BUSY.discard(s)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk3(x: torch.Tensor, name: str):
    if x.dtype != torch.float32 or not x.is_cuda or x.dim() != 3 or x.stride(2) != 1:
        raise ValueError(f"{name}: expected a CUDA float32 [B, L, C] tensor with unit channel stride, got "
                         f"{tuple(x.shape)} {x.dtype} {x.device} strides {x.stride()}")


@dataclass
class Pre:
    """Input transform fused into a conv: x*scale[b,c]+shift[b,c] then an activation."""
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    act: int = 0
    p0: float = 0.0
    a: Optional[torch.Tensor] = None
    b: Optional[torch.Tensor] = None


@dataclass
class ConvW:
    """Packed conv weights: ``w`` float32 [K, Cin/groups, Cout] (depthwise: [K, C]), optional bias [Cout];
    ``w_tc`` bf16 [K, Cout, cin_pad] for the tcgen05 path (dense layers only)."""
    w: torch.Tensor
    bias: Optional[torch.Tensor]
    K: int
    cin: int
    cout: int
    groups: int = 1
    w_tc: Optional[torch.Tensor] = None
    cin_pad: int = 0
    f16: bool = False          # w_tc is IEEE fp16 (fp16 checkpoints) instead of bf16
    w_tc_lo: Optional[torch.Tensor] = None   # fp32 checkpoints: w = w_tc + w_tc_lo (both bf16), kept to ~2^-17


# Tensor-core dispatch policy: "off" = CUDA-core fp32 everywhere; "x2" = tcgen05 with (hi, lo) bf16 activation planes
# (fp32-grade products); "x1" = tcgen05 with a single bf16 plane.
TC_MODE = [os.environ.get("B2A_TC", "x2")]
ATTN_MODE = [os.environ.get("B2A_ATTN", "tc")]      # "tc": tcgen05 flash attention for head_dim 64; "cuda": CUDA-core kernel
TC_MIN_K = 32                          # reduction length (Cin*K) below which the layer stays on the CUDA-core kernel


def _tc_eligible(cw: "ConvW", L: int, stride: int, transpose: bool, pad_mode: int, dilation: int = 1) -> bool:
    if TC_MODE[0] == "off" or cw.w_tc is None or pad_mode != 0 or cw.cout % 32 != 0 or L < 32:
        return False
    if transpose:                      # polyphase form: K = taps * stride, every phase is a `taps`-tap stride-1 conv
        return dilation == 1 and cw.K % stride == 0 and cw.cin * (cw.K // stride) >= TC_MIN_K
    return stride == 1 and cw.cin * cw.K >= TC_MIN_K


def _tc_transposed_weights(cw: "ConvW", stride: int):
    """([J, stride*Cout, cin_pad] hi, lo-or-None) with W[j, r*Cout + co, ci] = w[k = r + j*stride][ci][co] (cached per stride)."""
    cache = cw.__dict__.setdefault("_w_tc_tr", {})
    if stride not in cache:
        J = cw.K // stride
        w = cw.w.reshape(J, stride, cw.cin, cw.cout).permute(0, 1, 3, 2).reshape(J, stride * cw.cout, cw.cin)   # k = j*stride + r
        wt = torch.zeros(J, stride * cw.cout, cw.cin_pad, device=cw.w.device, dtype=cw.w_tc.dtype)
        wt[:, :, :cw.cin] = w.to(cw.w_tc.dtype)
        wl = None
        if cw.w_tc_lo is not None:
            wl = torch.zeros_like(wt)
            wl[:, :, :cw.cin] = (w - wt[:, :, :cw.cin].float()).to(wt.dtype)
        cache[stride] = (wt.contiguous(), None if wl is None else wl.contiguous())
    return cache[stride]


@dataclass
class Planes:
    """A tensor-core A operand already split by its producer: bf16 hi / lo planes [B, L, cin_pad] (what prep_bf16 would make)."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor]
    C: int

    @property
    def shape(self):
        return (self.hi.shape[0], self.hi.shape[1], self.C)


def pack_conv(w_mlx: torch.Tensor, bias=None, groups=1, device="cuda") -> ConvW:
    """MLX-layout conv weight [Cout, K, Cin/g] -> packed [K, Cin/g, Cout]."""
    cout, k, cin_g = w_mlx.shape
    if groups == 1:
        w = w_mlx.permute(1, 2, 0).contiguous()
        cin = cin_g
    else:
        if not (cin_g == 1 and groups == cout):
            raise NotImplementedError("only dense or depthwise convolutions are on the hot path")
        w = w_mlx[:, :, 0].t().contiguous()              # [K, C]
        cin = cout
    cwo = ConvW(w.to(device=device, dtype=torch.float32), None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous(),
                k, cin, cout, groups)
    if groups == 1 and cout % 32 == 0 and torch.device(device).type == "cuda":
        for dt in (torch.bfloat16, torch.float16):                     # 16-bit-exact weights only (bf16 / fp16 checkpoints)
            wb = w_mlx.float().to(dt)
            if torch.equal(wb.float(), w_mlx.float()):
                cpad = -(-cin // 64) * 64
                wt = torch.zeros(k, cout, cpad, dtype=dt)
                wt[:, :, :cin] = wb.permute(1, 0, 2)
                cwo.w_tc, cwo.cin_pad, cwo.f16 = wt.to(device).contiguous(), cpad, dt == torch.float16
                break
        else:                                                          # fp32 checkpoint (SNAC): split the weights as well
            cpad = -(-cin // 64) * 64
            w32 = w_mlx.float().permute(1, 0, 2)
            hi = w32.to(torch.bfloat16)
            lo = (w32 - hi.float()).to(torch.bfloat16)
            wt, wl = torch.zeros(k, cout, cpad, dtype=torch.bfloat16), torch.zeros(k, cout, cpad, dtype=torch.bfloat16)
            wt[:, :, :cin], wl[:, :, :cin] = hi, lo
            cwo.w_tc, cwo.w_tc_lo, cwo.cin_pad = wt.to(device).contiguous(), wl.to(device).contiguous(), cpad
    return cwo


def pack_linear(w: torch.Tensor, bias=None, device="cuda") -> ConvW:
    """nn.Linear weight [out, in] -> K=1 conv."""
    return pack_conv(w[:, None, :], bias, 1, device)


def conv1d(x: torch.Tensor, cw: ConvW, *, stride=1, dilation=1, pad_left=0, lout=None, pad_mode=0,
           pre: Optional[Pre] = None, post_act=0, post_p0=0.0, cscale=None, res=None, res_div=1,
           out_scale=1.0, out=None, accumulate=False, transpose=False, stats=False, emit: Optional[Pre] = None):
    """b2a_conv1d_cl / b2a_convtr1d_cl.  For ``transpose`` ``pad_left`` is the left crop of the scatter output.
    ``stats=True`` returns (y, partials): InstanceNorm partial sums of y from the tensor-core epilogue for ``adain_coeffs(partials=)``,
    or (y, None) when the layer does not run on that path."""
    if isinstance(x, Planes):                      # operand already split by its producer (conv1d(..., emit=...)): tensor-core path only
        B, L, cin = x.shape
        if cin != cw.cin or pre is not None or not _tc_eligible(cw, L, stride, transpose, pad_mode, dilation) or x.hi.shape[2] != cw.cin_pad:
            raise ValueError("conv1d: a Planes operand needs a tensor-core-eligible layer with matching channels and no prologue")
        if lout is None:
            lout = (L + 2 * pad_left - dilation * (cw.K - 1) - 1) // stride + 1
        return _conv1d_tc(x, cw, dilation, pad_left, lout, None, post_act, post_p0, cscale, res, res_div, out_scale, out, accumulate,
                          up_stride=stride if transpose else 0, stats=stats)
    _chk3(x, "conv1d x")
    B, L, cin = x.shape
    if cin != cw.cin:
        raise ValueError(f"conv1d: input has {cin} channels, weight expects {cw.cin}")
    if lout is None:
        if transpose:
            lout = (L - 1) * stride + cw.K - 2 * pad_left
        else:
            lout = (L + 2 * pad_left - dilation * (cw.K - 1) - 1) // stride + 1
    # Tiny GEMMs (a few rows: ALBERT at T = 130, decode-time prefills) are latency chains, not throughput problems: the fused kernel's
    # converter -> MMA -> split-K fix-up chain measured 30-40 us per launch there against ~14 us for the pre-split planes + TMA pipeline.
    small_gemm = SMALL_GEMM_SPLIT_PATH[0] and cw.K == 1 and B * L <= 256 and _tc_eligible(cw, L, stride, transpose, pad_mode, dilation)
    if FUSED_DISPATCH[0] and emit is None and not small_gemm and fused_eligible(x, cw, stride, dilation, transpose, pad_mode, out, res):
        y = conv_fused(FusedProblem(x, cw, stride=stride, dilation=dilation, pad_left=pad_left, lout=lout, pre=pre, post_act=post_act,
                                    post_p0=post_p0, cscale=cscale, res=res, res_div=res_div, out_scale=out_scale, out=out,
                                    accumulate=accumulate, transpose=transpose))[0]
        return (y, None) if stats else y
    if _tc_eligible(cw, L, stride, transpose, pad_mode, dilation):
        return _conv1d_tc(x, cw, dilation, pad_left, lout, pre, post_act, post_p0, cscale, res, res_div, out_scale, out, accumulate,
                          up_stride=stride if transpose else 0, stats=stats)
    planes = None
    if emit is not None:
        if not emit_eligible(cw, x, lout, stride, dilation, transpose) or res is not None or cscale is not None or accumulate or post_act or out is not None:
            raise ValueError("conv1d: emit= needs a stride-1 depthwise layer with Cout % 64 == 0 and no epilogue extras (see ops.emit_eligible)")
        planes = Planes(torch.empty(B, lout, cw.cout, device=x.device, dtype=torch.bfloat16),
                        torch.empty(B, lout, cw.cout, device=x.device, dtype=torch.bfloat16) if TC_MODE[0] == "x2" else None, cw.cout)
    elif out is None:
        out = torch.empty(B, lout, cw.cout, device=x.device, dtype=torch.float32)
    else:
        _chk3(out, "conv1d out")
        if out.shape != (B, lout, cw.cout):
            raise ValueError(f"conv1d: out has shape {tuple(out.shape)}, expected {(B, lout, cw.cout)}")
    p = Conv1dParams()
    p.x, p.x_bs, p.x_ld = x.data_ptr(), x.stride(0), x.stride(1)
    p.B, p.L, p.Cin = B, L, cin
    p.w, p.bias = cw.w.data_ptr(), _p(cw.bias)
    if planes is None:
        p.y, p.y_bs, p.y_ld = out.data_ptr(), out.stride(0), out.stride(1)
    else:
        p.emit_hi, p.emit_lo, p.emit_ld = planes.hi.data_ptr(), _p(planes.lo), cw.cout
        p.emit_act, p.emit_p0, p.emit_a, p.emit_b = emit.act, emit.p0, _p(emit.a), _p(emit.b)
    p.Lout, p.Cout = lout, cw.cout
    p.K, p.stride, p.dilation, p.pad_left, p.groups, p.pad_mode = cw.K, stride, dilation, pad_left, cw.groups, pad_mode
    if pre is not None:
        p.pre_scale, p.pre_shift = _p(pre.scale), _p(pre.shift)
        p.pre_act, p.pre_p0, p.pre_a, p.pre_b = pre.act, pre.p0, _p(pre.a), _p(pre.b)
    p.post_act, p.post_p0 = post_act, post_p0
    if cscale is not None:
        p.post_cscale = cscale.data_ptr()
        p.post_cscale_bs = cscale.stride(0) if cscale.dim() == 2 else 0
    if res is not None:
        _chk3(res, "conv1d res")
        p.res, p.res_bs, p.res_ld = res.data_ptr(), (res.stride(0) if res.shape[0] == B else 0), res.stride(1)
    p.res_div = res_div
    p.out_scale, p.accumulate = out_scale, int(accumulate)
    fn = _lib.lib().b2a_convtr1d_cl if transpose else _lib.lib().b2a_conv1d_cl
    if PROFILE_TAGS is not None:
        TAG[0] = f"cuda-core conv [{B}x{L}x{cw.cin}->{cw.cout} k{cw.K} s{stride} d{dilation} g{cw.groups}{' T' if transpose else ''}]"
    _call("conv" if cw.groups == 1 and cw.cin * cw.K >= 64 else "other", fn, 1, C.byref(p), _stream())
    if planes is not None:
        return planes
    return (out, None) if stats else out


def emit_eligible(cw: "ConvW", x: torch.Tensor, lout: int, stride: int = 1, dilation: int = 1, transpose: bool = False) -> bool:
    """True when conv1d(x, cw, emit=...) can write the next layer's bf16 planes directly (the vectorised depthwise kernel)."""
    return (TC_MODE[0] != "off" and not transpose and stride == 1 and cw.groups == cw.cin == cw.cout and cw.cout % 64 == 0 and cw.K <= 16
            and lout >= 128 and x.stride(1) % 4 == 0 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0
            and (128 + (cw.K - 1) * dilation) * 128 * 4 <= 160 * 1024)


def prep_bf16(x: torch.Tensor, pre: Optional[Pre], cpad: int, planes: int = 2, f16: bool = False):
    """Conv prologue -> (hi, lo) bf16 (or fp16) planes [B, L, cpad] (lo None when planes == 1)."""
    _chk3(x, "prep_bf16 x")
    B, L, Cc = x.shape
    dt = torch.float16 if f16 else torch.bfloat16
    hi = torch.empty(B, L, cpad, device=x.device, dtype=dt)
    lo = torch.empty(B, L, cpad, device=x.device, dtype=dt) if planes == 2 else None
    pre = pre or Pre()
    _call("prep", _lib.lib().b2a_prep_bf16, 1, x.data_ptr(), x.stride(0), x.stride(1), B, L, Cc, cpad, _p(pre.scale), _p(pre.shift),
          pre.act, pre.p0, _p(pre.a), _p(pre.b), hi.data_ptr(), _p(lo), int(f16), _stream())
    return hi, lo


# InstanceNorm partials from the conv epilogue: correct (tests/test_tc_gpu.py) but measured SLOWER on Kokoro (6.86 vs 6.74 ms per step:
# 51 fewer launches, yet the per-32-row float64 slots make the coefficient kernel strided and lengthen the epilogue that sits on the
# critical path) -> opt-in.
TC_STATS = [os.environ.get("B2A_TC_STATS", "0") != "0" and os.environ.get("B2A_TC_PERSIST", "1") != "0"]


def _conv1d_tc(x, cw, dilation, pad_left, lout, pre, post_act, post_p0, cscale, res, res_div, out_scale, out, accumulate,
               up_stride=0, stats=False):
    if isinstance(x, Planes):
        B, L, _ = x.shape
        hi, lo = x.hi, x.lo
    else:
        B, L, _ = x.shape
        hi, lo = prep_bf16(x, pre, cw.cin_pad, 2 if TC_MODE[0] == "x2" else 1, cw.f16)
    if out is None:
        out = torch.empty(B, lout, cw.cout, device=hi.device, dtype=torch.float32)
    else:
        _chk3(out, "conv1d out")
        if out.shape != (B, lout, cw.cout):
            raise ValueError(f"conv1d: out has shape {tuple(out.shape)}, expected {(B, lout, cw.cout)}")
    if up_stride:
        taps, n_total = cw.K // up_stride, up_stride * cw.cout
        w_tc, w_lo = _tc_transposed_weights(cw, up_stride)
        shifts = (C.c_int32 * taps)(*[-j for j in range(taps)])
    else:
        taps, n_total, w_tc, w_lo = cw.K, cw.cout, cw.w_tc, cw.w_tc_lo
        shifts = (C.c_int32 * taps)(*[k * dilation - pad_left for k in range(taps)])
    cs, cs_bs = (None, 0) if cscale is None else (cscale.data_ptr(), cscale.stride(0) if cscale.dim() == 2 else 0)
    r, r_bs, r_ld = (None, 0, 0)
    if res is not None:
        _chk3(res, "conv1d res")
        r, r_bs, r_ld = res.data_ptr(), (res.stride(0) if res.shape[0] == B else 0), res.stride(1)
    ws, slots = None, 0
    if stats and TC_STATS[0]:          # InstanceNorm partials of the output straight from the epilogue (persistent kernel only)
        mrows = (L + taps - 1) if up_stride else lout
        slots = -(-mrows // 128) * 4 * max(1, up_stride)
        ws = torch.empty(B, slots, cw.cout, 2, device=hi.device, dtype=torch.float64)
    _call("conv_tc", _lib.lib().b2a_conv1d_tc, 1, hi.data_ptr(), _p(lo), int(cw.f16), B, L, cw.cin_pad, w_tc.data_ptr(), _p(w_lo), taps, shifts, n_total, lout,
          _p(cw.bias), post_act, post_p0, cs, cs_bs, r, r_bs, r_ld, res_div, out_scale, int(accumulate), out.data_ptr(), out.stride(0),
          out.stride(1), up_stride, pad_left if up_stride else 0, _p(ws), slots, _stream())
    return (out, ws) if stats else out



# ---------------------------------------------------------------------------------------------------------------- fused tcgen05 conv
# One launch per layer (or per GROUP of independent layers): InstanceNorm / AdaIN coefficients from the producer's (sum, sumsq), the
# input activation, the 16-bit hi/lo split, the tap-summed GEMM, the epilogue and the output's (sum, sumsq) -- csrc/conv_fused.cu.
FUSED = [os.environ.get("B2A_FUSED", "1") != "0"]
# conv1d() / linear() route single layers to the fused kernel only on request: its A operand is converted by the CTA's own warps, which
# pays off where it removes whole passes (Kokoro's AdaIN statistics + prologue, called explicitly through conv_fused) but loses to the
# pre-split planes + TMA pipeline on plain wide GEMMs (measured, round 2: Whisper encoder GEMMs 23 -> 89 ms, Mimi 34 -> 89 ms, SNAC 22 -> 33 ms
# with the fused kernel as the default route).
FUSED_DISPATCH = [os.environ.get("B2A_FUSED_DISPATCH", "0") == "1"]


@contextlib.contextmanager
def fused_dispatch(on: bool = True):
    """Route eligible conv1d() / linear() calls inside the block through the fused kernel (Kokoro: thin layers at L <= 780 where the
    prologue pass of the split path is a separate launch on the critical path)."""
    old = FUSED_DISPATCH[0]
    FUSED_DISPATCH[0] = bool(on) and FUSED[0]
    try:
        yield
    finally:
        FUSED_DISPATCH[0] = old
SMALL_GEMM_SPLIT_PATH = [os.environ.get("B2A_SMALL_GEMM_SPLIT", "1") != "0"]
FUSED_WS_BYTES = 16 << 20
_FUSED_WS = {}


@dataclass
class PreStats:
    """Input transform whose scale/shift the kernel derives itself: InstanceNorm over L from ``stats`` [B, C, 2, 4] int64 (binned sum, sumsq:
    ``new_stats`` / ``stats_value``) as accumulated by the producing layer, folded with AdaIN's (1 + gamma) / beta rows ``gb`` [B, 2C] (None: plain InstanceNorm); then
    the activation."""
    stats: torch.Tensor
    gb: Optional[torch.Tensor] = None
    eps: float = 1e-5
    act: int = 0
    p0: float = 0.0
    a: Optional[torch.Tensor] = None
    b: Optional[torch.Tensor] = None


STAT_BINS, STAT_BIN0, STAT_BIN_BITS = 4, -100, 40          # csrc/common.cuh: B2A_NBIN, B2A_BIN0, B2A_BIN_BITS


def new_stats(B: int, Cc: int, device) -> torch.Tensor:
    """Zeroed (sum, sumsq) accumulator for ``FusedProblem(..., stats_out=)``: [B, C, 2, 4] int64 bins (multiples of 2^(-100 + 40 k)); integer
    atomics make the accumulation independent of the order in which CTAs arrive -> bit-reproducible statistics."""
    return torch.zeros(B, Cc, 2, STAT_BINS, device=device, dtype=torch.int64)


def stats_value(stats: torch.Tensor) -> torch.Tensor:
    """Binned accumulator [..., 4] int64 -> float64 values [...] (what repro_value computes on the device)."""
    w = torch.tensor([2.0 ** (STAT_BIN0 + STAT_BIN_BITS * k) for k in range(STAT_BINS)], dtype=torch.float64, device=stats.device)
    t = torch.zeros(stats.shape[:-1], dtype=torch.float64, device=stats.device)
    for k in range(STAT_BINS - 1, -1, -1):
        t = t + stats[..., k].double() * w[k]
    return t


def _al16(t: Optional[torch.Tensor]) -> bool:
    return t is None or (t.data_ptr() % 16 == 0 and t.stride(1) % 4 == 0 and t.stride(0) % 4 == 0)


def fused_eligible(x, cw: "ConvW", stride: int = 1, dilation: int = 1, transpose: bool = False, pad_mode: int = 0, out=None, res=None) -> bool:
    """Dense layers the fused kernel takes: tensor-core weights, stride 1 (or a polyphase transposed conv), taps spanning <= 64 rows,
    16-byte aligned fp32 rows."""
    if not FUSED[0] or TC_MODE[0] == "off" or cw.w_tc is None or pad_mode != 0 or cw.cout % 32 != 0 or cw.groups != 1 or isinstance(x, Planes):
        return False
    if x.dtype != torch.float32 or x.dim() != 3 or x.stride(2) != 1 or x.stride(1) % 4 or x.stride(0) % 4 or x.data_ptr() % 16:
        return False
    if x.stride(1) < -(-x.shape[2] // 4) * 4 or not _al16(out) or not _al16(res):
        return False
    if transpose:
        return dilation == 1 and cw.K % stride == 0 and cw.K // stride <= 32 and cw.cin * (cw.K // stride) >= TC_MIN_K
    return stride == 1 and cw.K <= 32 and (cw.K - 1) * dilation <= 64 and cw.cin * cw.K >= TC_MIN_K


class FusedProblem:
    """One problem of a fused launch: the filled C struct, its output tensor and the tensors it points into (kept alive)."""

    def __init__(self, x, cw: "ConvW", *, stride=1, dilation=1, pad_left=0, lout=None, pre=None, x_add=(), in_scale=1.0, post_act=0, post_p0=0.0,
                 cscale=None, res=None, res_div=1, out_scale=1.0, out=None, accumulate=False, transpose=False, stats_out=None):
        _chk3(x, "conv_fused x")
        B, L, cin = x.shape
        if cin != cw.cin:
            raise ValueError(f"conv_fused: input has {cin} channels, weight expects {cw.cin}")
        if lout is None:
            lout = (L - 1) * stride + cw.K - 2 * pad_left if transpose else (L + 2 * pad_left - dilation * (cw.K - 1) - 1) + 1
        if out is None:
            out = torch.empty(B, lout, cw.cout, device=x.device, dtype=torch.float32)
        else:
            _chk3(out, "conv_fused out")
            if out.shape != (B, lout, cw.cout):
                raise ValueError(f"conv_fused: out has shape {tuple(out.shape)}, expected {(B, lout, cw.cout)}")
        p = ConvFParams()
        p.x, p.x_bs, p.x_ld, p.in_scale = x.data_ptr(), x.stride(0), x.stride(1), float(in_scale)
        keep = [x, out, cw]
        for i, xa in enumerate(x_add):
            if xa.shape != x.shape or xa.stride() != x.stride() or xa.dtype != torch.float32:
                raise ValueError("conv_fused: x_add tensors must have x's shape and strides")
            setattr(p, "x1" if i == 0 else "x2", xa.data_ptr())
            keep.append(xa)
        p.B, p.L, p.Cin = B, L, cin
        if isinstance(pre, PreStats):
            if pre.stats.dtype != torch.int64 or tuple(pre.stats.shape) != (B, cin, 2, STAT_BINS) or not pre.stats.is_contiguous():
                raise ValueError("conv_fused: stats must be a contiguous int64 [B, Cin, 2, 4] tensor (ops.new_stats)")
            p.pre_mode, p.pre_stats, p.pre_eps = 2, pre.stats.data_ptr(), float(pre.eps)
            if pre.gb is not None:
                if pre.gb.shape[-1] != 2 * cin or pre.gb.stride(-1) != 1:
                    raise ValueError("conv_fused: gb must be [B, 2*Cin] rows (gamma | beta)")
                p.pre_gb, p.pre_gb_bs = pre.gb.data_ptr(), (pre.gb.stride(0) if pre.gb.dim() == 2 and pre.gb.shape[0] == B and B > 1 else 0)
            keep += [pre.stats, pre.gb]
        elif pre is not None and pre.scale is not None:
            p.pre_mode, p.pre_scale, p.pre_shift = 1, pre.scale.data_ptr(), pre.shift.data_ptr()
            keep += [pre.scale, pre.shift]
        if pre is not None:
            p.pre_act, p.pre_p0, p.pre_a, p.pre_b = pre.act, float(pre.p0), _p(pre.a), _p(pre.b)
            keep += [pre.a, pre.b]
        if transpose:
            taps, n_total = cw.K // stride, stride * cw.cout
            w_tc, w_lo = _tc_transposed_weights(cw, stride)
            shifts = [-j for j in range(taps)]
            p.up_stride, p.up_crop = stride, pad_left
        else:
            taps, n_total, w_tc, w_lo = cw.K, cw.cout, cw.w_tc, cw.w_tc_lo
            shifts = [k * dilation - pad_left for k in range(taps)]
        p.w_hi, p.w_lo, p.cin_pad, p.taps, p.N = w_tc.data_ptr(), _p(w_lo), cw.cin_pad, taps, n_total
        for i, sft in enumerate(shifts):
            p.shifts[i] = sft
        p.Lout, p.bias, p.post_act, p.post_p0 = lout, _p(cw.bias), post_act, float(post_p0)
        if cscale is not None:
            p.cscale, p.cscale_bs = cscale.data_ptr(), (cscale.stride(0) if cscale.dim() == 2 else 0)
            keep.append(cscale)
        if res is not None:
            _chk3(res, "conv_fused res")
            p.res, p.res_bs, p.res_ld = res.data_ptr(), (res.stride(0) if res.shape[0] == B else 0), res.stride(1)
            keep.append(res)
        p.res_div, p.out_scale, p.accumulate = res_div, float(out_scale), int(accumulate)
        p.y, p.y_bs, p.y_ld = out.data_ptr(), out.stride(0), out.stride(1)
        if stats_out is not None:
            if stats_out.dtype != torch.int64 or tuple(stats_out.shape) != (B, cw.cout, 2, STAT_BINS) or not stats_out.is_contiguous():
                raise ValueError("conv_fused: stats_out must be a contiguous int64 [B, Cout, 2, 4] tensor (ops.new_stats)")
            p.stats_out = stats_out.data_ptr()
            keep.append(stats_out)
        self.p, self.out, self.keep, self.f16 = p, out, keep, cw.f16


def conv_fused(problems) -> list:
    """Launch 1..4 independent fused conv problems on one persistent grid; returns their outputs."""
    if isinstance(problems, FusedProblem):
        problems = [problems]
    n = len(problems)
    if not 1 <= n <= 4:
        raise ValueError("conv_fused: 1..4 problems per launch")
    f16 = problems[0].f16
    if any(pr.f16 != f16 for pr in problems):
        raise ValueError("conv_fused: all problems of a launch must use the same 16-bit operand type")
    arr = (ConvFParams * n)(*[pr.p for pr in problems])
    dev = problems[0].out.device
    key = (dev, torch.cuda.current_stream().cuda_stream)
    ws = _FUSED_WS.get(key)
    if ws is None:
        ws = _FUSED_WS[key] = torch.zeros(FUSED_WS_BYTES, device=dev, dtype=torch.uint8)      # split-K partial tiles + (self-resetting) counters
    if PROFILE_TAGS is not None:
        TAG[0] = "fused " + " + ".join(f"[{q.B}x{q.L}x{q.Cin}->{q.N} k{q.taps}{' up' + str(q.up_stride) if q.up_stride else ''}]" for q in (pr.p for pr in problems))
    _call("conv_tc", _lib.lib().b2a_conv1d_fused, 1, arr, n, 2 if TC_MODE[0] == "x2" else 1, int(f16), ws.data_ptr(), ws.numel(), _stream())
    return [pr.out for pr in problems]


def linear(x: torch.Tensor, cw: ConvW, **kw) -> torch.Tensor:
    """nn.Linear on [..., in] via the K=1 conv; accepts [rows, in] or [B, L, in]."""
    if x.dim() == 2:
        o = kw.get("out")
        if o is not None and o.dim() == 2:
            kw["out"] = o[None]
        r = kw.get("res")
        if r is not None and r.dim() == 2:
            kw["res"] = r[None]
        return conv1d(x[None], cw, **kw)[0]
    return conv1d(x, cw, **kw)


def copy2d(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[r, c] = src[r, c] for 2-D (row-strided) float32 views."""
    rows, cols = src.shape
    assert dst.shape == src.shape and src.stride(1) == 1 and dst.stride(1) == 1
    _call("other", _lib.lib().b2a_copy2d, 1, src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, _stream())


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r, :] = src[idx[r], :] (+ add[r % add.shape[0], :]); src [N, C] float32, idx int64 [R]."""
    assert src.dim() == 2 and src.stride(1) == 1 and idx.dtype == torch.int64 and idx.is_contiguous()
    rows, cols = idx.shape[0], src.shape[1]
    if out is None:
        out = torch.empty(rows, cols, device=src.device, dtype=torch.float32)
    a, a_ld, a_per = (None, 0, 0) if add is None else (add.data_ptr(), add.stride(0), add.shape[0])
    _call("other", _lib.lib().b2a_gather_rows, 1, src.data_ptr(), src.stride(0), idx.data_ptr(), out.data_ptr(), out.stride(0),
          rows, cols, src.shape[0], a, a_ld, a_per, _stream())
    return out


def whisper_greedy_step(logits: torch.Tensor, tokens: torch.Tensor, cur_len: int, sample_begin: int, *, suppress_mask, blank_mask,
                        eot: int, no_timestamps: int, timestamp_begin: int, max_initial_ts: int, without_timestamps: bool,
                        sum_logprobs: torch.Tensor, not_done: torch.Tensor, temperature: float = 0.0, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused logit filters + greedy / categorical update (decoding.py:295-325,349-442).  logits [B,V] fp32, tokens [B, >=cur_len] int64;
    ``temperature`` > 0 draws from softmax(filtered / temperature) with one uniform per row ``u`` [B]."""
    B, V = logits.shape
    assert logits.stride(1) == 1 and tokens.dtype == torch.int64 and tokens.stride(1) == 1
    assert u is None or (u.dtype == torch.float32 and u.is_contiguous() and u.numel() == B)
    nxt = torch.empty(B, device=logits.device, dtype=torch.int64)
    _call("sampler", _lib.lib().b2a_whisper_greedy_step, 1, logits.data_ptr(), logits.stride(0), tokens.data_ptr(), tokens.stride(0), B, cur_len,
          sample_begin, V, _p(suppress_mask), _p(blank_mask), eot, no_timestamps, timestamp_begin, max_initial_ts, int(without_timestamps),
          nxt.data_ptr(), sum_logprobs.data_ptr(), not_done.data_ptr(), float(temperature), _p(u), _stream())
    return nxt


def durations_to_index(dur, max_frames: int, speed: float = 1.0):
    """Durations [T] (float32 pre-round sums, or int64 already-rounded) -> (pred_dur int64 [T], idx int64 [max_frames]
    whose first `total` entries are valid, total int64 [1] on device)."""
    assert dur.is_contiguous() and dur.dtype in (torch.float32, torch.int64)
    T = dur.shape[0]
    pred = torch.empty(T, device=dur.device, dtype=torch.int64)
    idx = torch.zeros(max_frames, device=dur.device, dtype=torch.int64)
    total = torch.zeros(1, device=dur.device, dtype=torch.int64)
    f, i = (dur.data_ptr(), None) if dur.dtype == torch.float32 else (None, dur.data_ptr())
    _call("other", _lib.lib().b2a_durations_to_index, 1, f, i, T, speed, pred.data_ptr(), idx.data_ptr(), max_frames, total.data_ptr(), _stream())
    return pred, idx, total


_WS = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)
        _WS[key] = ws
    return ws


def adain_coeffs(x: torch.Tensor, gb: Optional[torch.Tensor], eps=1e-5, partials: Optional[torch.Tensor] = None):
    """InstanceNorm stats of x [B,L,C] folded with AdaIN (gamma|beta) [B,2C] -> (scale, shift) [B,C].  ``partials`` [B,slots,C,2]
    float64 from ``conv1d(..., stats=True)`` skips the statistics pass over x."""
    _chk3(x, "adain_coeffs x")
    B, L, Cc = x.shape
    scale = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    shift = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    if partials is not None:
        assert partials.dtype == torch.float64 and partials.is_contiguous() and partials.shape[0] == B and partials.shape[2] == Cc
        _call("adain_stats", _lib.lib().b2a_adain_coeffs_from_partials, 1, partials.data_ptr(), partials.shape[1], B, L, Cc, _p(gb), eps,
              scale.data_ptr(), shift.data_ptr(), _stream())
        return scale, shift
    ws = _workspace(_lib.lib().b2a_adain_ws_bytes(B, L, Cc), x.device)
    _call("adain_stats", _lib.lib().b2a_adain_coeffs, 2, x.data_ptr(), x.stride(0), x.stride(1), B, L, Cc, _p(gb), eps,
                                           scale.data_ptr(), shift.data_ptr(), ws.data_ptr(), _stream())
    return scale, shift


def channel_stats(x: torch.Tensor, dsts) -> None:
    """Add (sum, sumsq) over L of x [B, L, C] to each binned accumulator in ``dsts`` (views [B, C, 2, 4] of possibly wider [B, C', 2, 4] buffers)."""
    _chk3(x, "channel_stats x")
    B, L, Cc = x.shape
    if isinstance(dsts, torch.Tensor):
        dsts = [dsts]
    n = len(dsts)
    for d in dsts:
        assert d.dtype == torch.int64 and tuple(d.shape) == (B, Cc, 2, STAT_BINS) and d.stride(3) == 1 and d.stride(2) == STAT_BINS and d.stride(1) == 2 * STAT_BINS
    ptrs = (C.c_void_p * n)(*[d.data_ptr() for d in dsts])
    bss = (C.c_int64 * n)(*[d.stride(0) for d in dsts])
    _call("adain_stats", _lib.lib().b2a_channel_stats, 1, x.data_ptr(), x.stride(0), x.stride(1), B, L, Cc, ptrs, bss, n, _stream())


def coeffs_from_stats(stats: torch.Tensor, L: int, gb: Optional[torch.Tensor], eps=1e-5):
    """Binned (sum, sumsq) [B, C, 2, 4] -> AdaIN (scale, shift) [B, C] float32 for consumers outside the fused conv (depthwise layers)."""
    B, Cc = stats.shape[:2]
    assert stats.dtype == torch.int64 and stats.is_contiguous() and tuple(stats.shape[2:]) == (2, STAT_BINS)
    scale = torch.empty(B, Cc, device=stats.device, dtype=torch.float32)
    shift = torch.empty(B, Cc, device=stats.device, dtype=torch.float32)
    _call("adain_stats", _lib.lib().b2a_coeffs_from_stats, 1, stats.data_ptr(), B, L, Cc, _p(gb), eps, scale.data_ptr(), shift.data_ptr(), _stream())
    return scale, shift


def layernorm(x: torch.Tensor, w=None, b=None, *, eps=1e-5, res=None, ada=None, rms=False, post_act=0, post_p0=0.0,
              out=None) -> torch.Tensor:
    """Row LayerNorm / RMSNorm over the last dim of a 2-D row-strided view."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]) if x.dim() != 2 else x
    assert x2.stride(1) == 1
    r2 = None
    if res is not None:
        r2 = res.reshape(-1, shp[-1]) if res.dim() != 2 else res
    if out is None:
        out = torch.empty(x2.shape, device=x.device, dtype=torch.float32)
    o2 = out.reshape(-1, shp[-1]) if out.dim() != 2 else out
    _call("layernorm", _lib.lib().b2a_layernorm, 1, x2.data_ptr(), x2.stride(0), _p(r2), 0 if r2 is None else r2.stride(0), o2.data_ptr(),
                                        o2.stride(0), x2.shape[0], shp[-1], _p(w), _p(b), _p(ada), eps, int(rms), post_act,
                                        post_p0, _stream())
    return out.reshape(shp) if out.dim() == 2 and len(shp) != 2 else out


def attention(q, k, v, *, n_heads, n_kv_heads=None, scale, causal=False, q_offset=0, window=0, k_len=None, out=None):
    """softmax(scale q k^T + mask) v.  q [B,Tq,H*D], k/v [B,Tk,Hkv*D] (row-strided views allowed)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk3(t, "attention " + n)
    B, Tq, hd = q.shape
    H = n_heads
    Hkv = n_kv_heads or H
    D = hd // H
    if out is None:
        out = torch.empty(B, Tq, hd, device=q.device, dtype=torch.float32)
    p = AttnParams()
    p.q, p.k, p.v, p.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    p.q_bs, p.q_ld, p.k_bs, p.k_ld = q.stride(0), q.stride(1), k.stride(0), k.stride(1)
    p.v_bs, p.v_ld, p.o_bs, p.o_ld = v.stride(0), v.stride(1), out.stride(0), out.stride(1)
    p.B, p.Tq, p.Tk, p.H, p.Hkv, p.D = B, Tq, k.shape[1], H, Hkv, D
    p.scale, p.causal, p.q_offset, p.window = scale, int(causal), q_offset, window
    p.k_len = _p(k_len)
    if ATTN_MODE[0] == "tc" and D == 64 and Hkv == H and k_len is None and k.shape[1] >= 64 and out.stride(1) % 4 == 0:
        ws = torch.empty(_lib.lib().b2a_attention_tc_ws_bytes(B, H, Tq, k.shape[1]), device=q.device, dtype=torch.uint8)
        _call("attention", _lib.lib().b2a_attention_tc, 4, C.byref(p), ws.data_ptr(), _stream())
        return out
    _call("attention", _lib.lib().b2a_attention, 1, C.byref(p), _stream())
    return out


def rope_(x: torch.Tensor, n_heads: int, *, offset=0, base=10000.0, traditional=True) -> torch.Tensor:
    """In-place rotary embedding on x [B,T,H*D]."""
    _chk3(x, "rope x")
    B, T, hd = x.shape
    _call("rope", _lib.lib().b2a_rope, 1, x.data_ptr(), x.stride(0), x.stride(1), B, T, n_heads, hd // n_heads, offset, base,
                                   int(traditional), _stream())
    return x


def lstm_bidir(xproj: torch.Tensor, wh: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """xproj [B,T,2*4H] (forward|backward input projections incl. biases), wh [2,4H,H] -> [B,T,2H]."""
    B, T, g8 = xproj.shape
    H = g8 // 8
    assert xproj.is_contiguous() and wh.is_contiguous() and wh.shape == (2, 4 * H, H)
    if out is None:
        out = torch.empty(B, T, 2 * H, device=xproj.device, dtype=torch.float32)
    assert out.stride(2) == 1 and (B == 1 or out.stride(0) == T * out.stride(1))
    _call("lstm", _lib.lib().b2a_lstm_bidir, 1, xproj.data_ptr(), wh.data_ptr(), out.data_ptr(), out.stride(1), B, T, H, _stream())
    return out


def stft(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int, pad_mode: int, frames: int):
    """x [B,n] -> (re, im) each [B, frames, n_fft//2+1]; pad_mode 0 none / 1 reflect / 2 constant."""
    B, n = x.shape
    assert x.stride(1) == 1 and window.shape[0] == n_fft
    nf = n_fft // 2 + 1
    re = torch.empty(B, frames, nf, device=x.device, dtype=torch.float32)
    im = torch.empty(B, frames, nf, device=x.device, dtype=torch.float32)
    _call("other", _lib.lib().b2a_stft, 1, x.data_ptr(), x.stride(0), B, n, window.data_ptr(), n_fft, hop, pad_mode, frames,
                                   re.data_ptr(), im.data_ptr(), _stream())
    return re, im


def whisper_logmel(x: torch.Tensor, padding: int, window: torch.Tensor, filters: torch.Tensor, frames: int) -> torch.Tensor:
    """x [B,n] float32 -> log-mel [B, frames, n_mels] (audio.py:41-82)."""
    B, n = x.shape
    assert x.stride(1) == 1 and filters.is_contiguous() and filters.shape[1] == 201
    out = torch.empty(B, frames, filters.shape[0], device=x.device, dtype=torch.float32)
    gmax = torch.empty(B, device=x.device, dtype=torch.float32)
    _call("logmel", _lib.lib().b2a_whisper_logmel, 2, x.data_ptr(), x.stride(0), B, n, padding, window.data_ptr(), filters.data_ptr(),
                                             filters.shape[0], frames, out.data_ptr(), gmax.data_ptr(), _stream())
    return out


def istft(re: torch.Tensor, im: torch.Tensor, n_fft: int, hop: int, window: torch.Tensor, *, norm_sq: bool, clamp_mode: int,
          trim: int, out_len: int) -> torch.Tensor:
    """re/im [B, n_freq, T] -> [B, out_len]."""
    B, nf, T = re.shape
    assert re.is_contiguous() and im.is_contiguous() and nf == n_fft // 2 + 1
    out = torch.empty(B, out_len, device=re.device, dtype=torch.float32)
    ws = torch.empty(B, T, n_fft, device=re.device, dtype=torch.float32)
    _call("other", _lib.lib().b2a_istft, 2, re.data_ptr(), im.data_ptr(), B, n_fft, T, hop, window.data_ptr(), int(norm_sq), clamp_mode,
                                    trim, out_len, out.data_ptr(), ws.data_ptr(), _stream())
    return out


def resample_poly(x: torch.Tensor, h: torch.Tensor, up: int, down: int, n_pre_pad: int, n_pre_remove: int, n_out: int) -> torch.Tensor:
    """x [B,n] float32, h float64 FIR (already * up) -> [B, n_out] (scipy.signal.resample_poly, padtype='edge')."""
    assert x.dim() == 2 and x.stride(1) == 1 and h.dtype == torch.float64 and h.is_contiguous()
    out = torch.empty(x.shape[0], n_out, device=x.device, dtype=torch.float32)
    _call("resample", _lib.lib().b2a_resample_poly, 1, x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], h.data_ptr(), h.shape[0], up, down,
          n_pre_pad, n_pre_remove, out.data_ptr(), n_out, _stream())
    return out


def kokoro_source(f0: torch.Tensor, noise: Optional[torch.Tensor], lin_w: torch.Tensor, lin_b: torch.Tensor) -> torch.Tensor:
    """F0 curve [B, nF] -> har [B, nF*60+1, 22] (magnitude | phase of the hn-NSF source STFT)."""
    B, nF = f0.shape
    assert f0.is_contiguous()
    har = torch.empty(B, nF * 60 + 1, 22, device=f0.device, dtype=torch.float32)
    src = torch.empty(B, nF * 300, device=f0.device, dtype=torch.float32)
    # length of the reference's down-sampled phase track: size = ceil(float(W) * float(scale)) with scale = 1/300
    # (tts/models/interpolate.py:43-50) -- nF or nF + 1 depending on floating-point rounding; evaluated the same way here.
    n_down = max(1, int(math.ceil(float(nF * 300) * float(1 / 300))))
    ph = torch.empty(B, n_down, 9, device=f0.device, dtype=torch.float64)
    if noise is not None:
        assert noise.is_contiguous() and noise.shape == (B, nF * 300, 9)
    _call("source", _lib.lib().b2a_kokoro_source, 3, f0.data_ptr(), B, nF, n_down, _p(noise), lin_w.data_ptr(), lin_b.data_ptr(), har.data_ptr(),
                                            src.data_ptr(), ph.data_ptr(), _stream())
    return har


def kokoro_istft_head(x: torch.Tensor) -> torch.Tensor:
    """conv_post output [B,T,22] -> waveform [B, (T-1)*5]."""
    _chk3(x, "kokoro_istft_head x")
    B, T, _ = x.shape
    audio = torch.empty(B, (T - 1) * 5, device=x.device, dtype=torch.float32)
    _call("istft_head", _lib.lib().b2a_kokoro_istft_head, 1, x.data_ptr(), x.stride(0), x.stride(1), B, T, audio.data_ptr(), _stream())
    return audio


def randn_(out: torch.Tensor, seed: int, offset: int = 0) -> torch.Tensor:
    """Fill a contiguous float32 tensor with N(0,1) draws from our Philox kernel."""
    assert out.is_contiguous() and out.dtype == torch.float32
    _call("other", _lib.lib().b2a_randn, 1, out.data_ptr(), out.numel(), seed, offset, _stream())
    return out


def randn_dev_(out: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """N(0,1) fill keyed by a device-resident Philox state (int64 [2] = seed, counter offset) that the call advances: graph-replay safe."""
    assert out.is_contiguous() and out.dtype == torch.float32 and state.dtype == torch.int64 and state.numel() == 2 and state.is_contiguous()
    _call("other", _lib.lib().b2a_randn_dev, 2, out.data_ptr(), out.numel(), state.data_ptr(), _stream())
    return out


def sample_token(logits: torch.Tensor, *, temperature: float, top_k: int = 0, top_p: float = 1.0, min_p: float = 0.0, u=None,
                 suppress_mask=None, seen=None, repetition_penalty: float = 1.0, return_filtered: bool = False, mark_seen: bool = False,
                 out: Optional[torch.Tensor] = None, finished=None, eos: int = -1):
    """Fused sampler on logits [B,V] (V <= 4096) -> int64 tokens [B] (and the filtered logits when asked).  ``out`` may be a
    strided int64 view (e.g. a column of the [B,16] code matrix)."""
    B, V = logits.shape
    assert logits.stride(1) == 1 and (seen is None or (seen.dtype == torch.uint8 and seen.stride(1) == 1))
    if out is None:
        out = torch.empty(B, device=logits.device, dtype=torch.int64)
    assert out.dtype == torch.int64 and out.dim() == 1 and out.shape[0] == B
    filt = torch.empty(B, V, device=logits.device, dtype=torch.float32) if return_filtered else None
    _call("sampler", _lib.lib().b2a_sample_token, 1, logits.data_ptr(), logits.stride(0), B, V, _p(suppress_mask), _p(seen),
          0 if seen is None else seen.stride(0), int(mark_seen), repetition_penalty, temperature, top_k, top_p, min_p, _p(u),
          out.data_ptr(), out.stride(0) if B > 1 else 1, _p(filt), _p(finished), eos, _stream())
    return (out, filt) if return_filtered else out


def gemv_eligible(cw: "ConvW") -> bool:
    """The decode GEMV streams ONE bf16 plane of weights: bf16-exact K=1 layers only (fp16 / fp32 checkpoints and layers whose
    Cout is not a multiple of 32 go through ``linear``)."""
    return cw.K == 1 and cw.w_tc is not None and not cw.f16 and cw.w_tc_lo is None


def gemv(x: torch.Tensor, cw: "ConvW", *, norm_w=None, norm_eps: float = 1e-6, swiglu: bool = False, res=None, out=None,
         prefetch: Optional["ConvW"] = None) -> torch.Tensor:
    """Decode-time nn.Linear on x [M, K] (any M; looped in groups of 8) with the bf16 weight rows of ``cw`` ([N, cin_pad]):
    optional fused RMSNorm prologue, SwiGLU (interleaved gate/up rows) and residual.  ``prefetch``: the next projection, whose
    weights are pulled into L2 while this one runs."""
    if not (x.dim() == 2 and x.stride(1) == 1 and gemv_eligible(cw)):
        raise NotImplementedError("gemv needs a bf16-exact K=1 weight with Cout % 32 == 0 (see ops.gemv_eligible); use ops.linear")
    M, K = x.shape
    N = cw.cout
    n_out = N // 2 if swiglu else N
    if out is None:
        out = torch.empty(M, n_out, device=x.device, dtype=torch.float32)
    assert K == cw.cin and out.shape == (M, n_out) and out.stride(1) == 1
    pf, pf_bytes = (None, 0) if prefetch is None or prefetch.w_tc is None else (prefetch.w_tc.data_ptr(), prefetch.w_tc.numel() * 2)
    for m0 in range(0, M, 8):
        m = min(8, M - m0)
        r = None if res is None else res[m0:m0 + m]
        _call("gemv", _lib.lib().b2a_gemv_bf16, 1, x[m0:m0 + m].data_ptr(), x.stride(0), m, K, cw.w_tc.data_ptr(), cw.cin_pad, N,
              _p(cw.bias), _p(norm_w), norm_eps, int(swiglu), _p(r), 0 if r is None else r.stride(0), out[m0:m0 + m].data_ptr(),
              out.stride(0), pf if m0 == 0 else None, pf_bytes, _stream())
    return out


def qknorm_rope_cache(qkv: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, k_cache: torch.Tensor, v_cache: torch.Tensor, *,
                      q_norm=None, k_norm=None, eps: float = 1e-6, pos3=None, base_dev=None, base: int = 0, mrope=(0, 0),
                      theta: float = 10000.0, q_out=None, pos_shift=None) -> torch.Tensor:
    """qkv [B,S,(Hq+2Hkv)D] -> q_out [B,S,Hq*D] (normed + rotated), k/v appended to caches [B,Smax,Hkv*D] at row base+s."""
    _chk3(qkv, "qkv")
    B, S, _ = qkv.shape
    if q_out is None:
        q_out = torch.empty(B, S, n_heads * head_dim, device=qkv.device, dtype=torch.float32)
    assert k_cache.shape == v_cache.shape and k_cache.stride() == v_cache.stride() and k_cache.stride(2) == 1
    assert pos3 is None or (pos3.dtype == torch.int32 and pos3.is_contiguous() and pos3.shape == (3, B, S))
    _call("rope", _lib.lib().b2a_qknorm_rope_cache, 1, qkv.data_ptr(), qkv.stride(0), qkv.stride(1), B, S, n_heads, n_kv, head_dim,
          _p(q_norm), _p(k_norm), eps, _p(pos3), _p(base_dev), base, mrope[0], mrope[1], theta, q_out.data_ptr(), q_out.stride(0),
          q_out.stride(1), k_cache.data_ptr(), v_cache.data_ptr(), k_cache.stride(0), k_cache.stride(1), k_cache.shape[1], _p(pos_shift), _stream())
    return q_out


def attn_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, *, scale: float,
                base_dev=None, base: int = 0, kv_start=None, max_k: Optional[int] = None, out=None) -> torch.Tensor:
    """Causal GQA attention of q [B,S,Hq*D] against cache rows [kv_start, base+s]; out [B,S,Hq*D]."""
    _chk3(q, "q")
    B, S, _ = q.shape
    if out is None:
        out = torch.empty(B, S, n_heads * head_dim, device=q.device, dtype=torch.float32)
    if max_k is None:
        max_k = k_cache.shape[1] if base_dev is not None else base + S
    _call("attention", _lib.lib().b2a_attn_decode, 1, q.data_ptr(), q.stride(0), q.stride(1), k_cache.data_ptr(), v_cache.data_ptr(),
          k_cache.stride(0), k_cache.stride(1), out.data_ptr(), out.stride(0), out.stride(1), B, S, n_heads, n_kv, head_dim, scale,
          _p(base_dev), base, _p(kv_start), max_k, _stream())
    return out


def attn_decode_fused(qkv: torch.Tensor, n_heads: int, n_kv: int, head_dim: int, k_cache: torch.Tensor, v_cache: torch.Tensor, *, scale: float,
                      q_norm=None, k_norm=None, eps: float = 1e-6, pos3=None, base_dev=None, base: int = 0, mrope=(0, 0),
                      theta: float = 10000.0, kv_start=None, out=None) -> torch.Tensor:
    """Single-token decode step: qkv [B,(Hq+2Hkv)D] -> attention output [B,Hq*D]; k/v appended to the caches at row ``base``."""
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and n_heads == 2 * n_kv and k_cache.stride() == v_cache.stride()
    B = qkv.shape[0]
    if out is None:
        out = torch.empty(B, n_heads * head_dim, device=qkv.device, dtype=torch.float32)
    assert pos3 is None or (pos3.dtype == torch.int32 and pos3.is_contiguous() and pos3.numel() == 3 * B)
    _call("attention", _lib.lib().b2a_attn_decode_fused, 1, qkv.data_ptr(), qkv.stride(0), B, n_heads, n_kv, head_dim, _p(q_norm), _p(k_norm),
          eps, _p(pos3), _p(base_dev), base, mrope[0], mrope[1], theta, k_cache.data_ptr(), v_cache.data_ptr(), k_cache.stride(0),
          k_cache.stride(1), k_cache.shape[1], scale, _p(kv_start), out.data_ptr(), out.stride(0), _stream())
    return out


def swiglu(x: torch.Tensor, out=None, interleaved: bool = False) -> torch.Tensor:
    """x [..., 2I] (gate | up halves, or interleaved pairs) -> silu(gate) * up [..., I]."""
    shp = x.shape
    I = shp[-1] // 2
    x2 = x.reshape(-1, shp[-1])
    assert x2.stride(1) == 1
    if out is None:
        out = torch.empty(*shp[:-1], I, device=x.device, dtype=torch.float32)
    o2 = out.reshape(-1, I)
    _call("other", _lib.lib().b2a_swiglu, 1, x2.data_ptr(), x2.stride(0), x2.shape[0], I, int(interleaved), o2.data_ptr(), o2.stride(0), _stream())
    return out


class EmbedTables:
    """Device arrays of table pointers / sizes for embed_sum (built once per model)."""

    def __init__(self, tables):
        self.tables = [t for t in tables]
        for t in self.tables:
            assert t.dim() == 2 and t.is_contiguous() and t.dtype == torch.float32 and t.shape[1] == self.tables[0].shape[1]
        dev = self.tables[0].device
        self.ptrs = torch.tensor([t.data_ptr() for t in self.tables], dtype=torch.int64, device=dev)
        self.bins = torch.tensor([t.shape[0] for t in self.tables], dtype=torch.int32, device=dev)
        self.dim = self.tables[0].shape[1]


def embed_sum(codes: torch.Tensor, tabs: EmbedTables, *, text=None, pad=None, step_dev=None, step_sub: int = 0, out=None, err=None,
              tidx=None, finished=None) -> torch.Tensor:
    """out[b] = text-or-pad(b) + sum_g tables[g][codes[b,g]]; codes int64 [B,G] (G <= len(tables))."""
    assert codes.dtype == torch.int64 and codes.dim() == 2 and codes.stride(1) == 1
    B, G = codes.shape
    assert G <= len(tabs.tables)
    if out is None:
        out = torch.empty(B, tabs.dim, device=codes.device, dtype=torch.float32)
    tb, ts, nt = (0, 0, 0) if text is None else (text.stride(0), text.stride(1), text.shape[1])
    _call("other", _lib.lib().b2a_embed_sum, 1, codes.data_ptr(), codes.stride(0), B, G, tabs.dim, tabs.ptrs.data_ptr(), tabs.bins.data_ptr(),
          _p(text), tb, ts, nt, _p(pad), _p(step_dev), step_sub, out.data_ptr(), out.stride(0), _p(err), _p(tidx), _p(finished), _stream())
    return out


def incr_(p: torch.Tensor, v: int = 1) -> None:
    assert p.dtype == torch.int32
    _call("other", _lib.lib().b2a_incr_i32, 1, p.data_ptr(), v, _stream())


def rvq_decode(codes: torch.Tensor, codebooks: torch.Tensor, out: Optional[torch.Tensor] = None, check=True) -> torch.Tensor:
    """codes int64 [B,nq,T], codebooks [nq,bins,dim] -> sum of gathers [B,T,dim]."""
    assert codes.dtype == torch.int64 and codes.stride(2) == 1 and codebooks.is_contiguous()
    B, nq, T = codes.shape
    _, bins, dim = codebooks.shape
    if out is None:
        out = torch.empty(B, T, dim, device=codes.device, dtype=torch.float32)
    err = torch.zeros(1, device=codes.device, dtype=torch.int32)
    _call("rvq", _lib.lib().b2a_rvq_decode, 1, codes.data_ptr(), codes.stride(0), codes.stride(1), B, nq, T, codebooks.data_ptr(), bins,
                                         dim, out.data_ptr(), out.stride(1), err.data_ptr(), _stream())
    if check and int(err.item()) != 0:
        raise ValueError(f"rvq_decode: code index out of range [0, {bins})")
    return out


def rvq_encode(x: torch.Tensor, codebooks: torch.Tensor, c2: torch.Tensor, *, mode: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Nearest-code search with the residual loop: x [R, D] fp32 rows, codebooks [nq, bins, D], c2 [nq, bins] float64 (|e|^2 / 2, or |en|^2 for
    ``mode=1`` = SNAC's single-level cosine search on an L2-normalised table) -> int64 codes [R, nq] (or ``out``, any strides)."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32 and codebooks.is_contiguous() and c2.dtype == torch.float64 and c2.is_contiguous()
    R, D = x.shape
    nq, bins, _ = codebooks.shape
    if out is None:
        out = torch.empty(R, nq, device=x.device, dtype=torch.int64)
    assert out.shape == (R, nq) and out.dtype == torch.int64
    _call("rvq", _lib.lib().b2a_rvq_encode, 1, x.data_ptr(), x.stride(0), R, D, codebooks.data_ptr(), c2.data_ptr(), bins, nq, mode, out.data_ptr(),
          out.stride(0), out.stride(1), _stream())
    return out


def snac_from_codes(codes, strides, embs, ws, biases, dim: int, check=True) -> torch.Tensor:
    """SNAC quantizer.from_codes: codes[l] int64 [B, T/stride_l]; embs[l] [bins, cd]; ws[l] [cd, dim]; -> [B,T,dim]."""
    n = len(codes)
    B = codes[0].shape[0]
    T = codes[-1].shape[1] * strides[-1]
    bins, cd = embs[0].shape
    out = torch.empty(B, T, dim, device=codes[0].device, dtype=torch.float32)
    err = torch.zeros(1, device=out.device, dtype=torch.int32)
    arr = lambda ts: (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
    for c in codes:
        assert c.dtype == torch.int64 and c.is_contiguous()
    _call("rvq", _lib.lib().b2a_snac_from_codes, 1, arr(codes), (C.c_int32 * n)(*strides), n, arr(embs), arr(ws), arr(biases), B, T, bins,
                                              cd, dim, out.data_ptr(), err.data_ptr(), _stream())
    if check and int(err.item()) != 0:
        raise ValueError(f"snac_from_codes: code index out of range [0, {bins})")
    return out
