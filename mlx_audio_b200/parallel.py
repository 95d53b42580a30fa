"""Multi-GPU partitioning of the hot path: one process per GPU, independent units (utterances, 30-s
windows, code-frame spans) sharded across ranks, NO collective inside a unit.
``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is used only for the trailing gather of
result lengths + waveforms and for bench barriers."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_units(lengths: Sequence[int], rank: int, world_size: int) -> List[int]:
    """Indices of the units this rank owns: longest-first round-robin (balances padded / sequential work)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return sorted(order[rank::world_size])


def shard_span(n_frames: int, rank: int, world_size: int, *, halo_left: int = 0, halo_right: int = 0, multiple: int = 1):
    """Contiguous span of a code-frame stream for this rank plus the halo it must decode to be exact.

    Returns (read_start, read_end, core_start, core_end): decode frames [read_start, read_end) and keep the
    outputs of [core_start, core_end).  ``multiple`` aligns core boundaries (e.g. SNAC's coarsest vq stride)."""
    per = -(-n_frames // world_size)
    per = -(-per // multiple) * multiple
    core_start = min(rank * per, n_frames)
    core_end = min(core_start + per, n_frames)
    hl = -(-halo_left // multiple) * multiple
    hr = -(-halo_right // multiple) * multiple
    return max(0, core_start - hl), min(n_frames, core_end + hr), core_start, core_end


def gather_waveforms(local: List[torch.Tensor], local_ids: List[int], n_total: int, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Trailing gather: every rank contributes its utterances' waveforms; rank ``dst`` gets them back in unit order.
    One all_gather of int64 lengths, then one padded all_gather of samples (KB-MB messages: latency-bound)."""
    rank, ws = world()
    if ws == 1:
        out = [None] * n_total
        for i, w in zip(local_ids, local):
            out[i] = w
        return out
    dev = local[0].device if local else torch.device("cpu")
    per = -(-n_total // ws)
    meta = torch.full((per, 2), -1, dtype=torch.int64, device=dev)
    for j, (i, w) in enumerate(zip(local_ids, local)):
        meta[j, 0], meta[j, 1] = i, w.numel()
    metas = [torch.empty_like(meta) for _ in range(ws)]
    dist.all_gather(metas, meta)
    max_len = int(max(int(m[:, 1].max()) for m in metas))
    buf = torch.zeros(per, max(max_len, 1), dtype=torch.float32, device=dev)
    for j, w in enumerate(local):
        buf[j, : w.numel()] = w.reshape(-1)
    bufs = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(bufs, buf)
    if rank != dst:
        return None
    out = [None] * n_total
    for m, b in zip(metas, bufs):
        for j in range(per):
            i, n = int(m[j, 0]), int(m[j, 1])
            if i >= 0:
                out[i] = b[j, :n].clone()
    return out


def decode_stream_sharded(model, codes, *, noises=None, dst: int = 0):
    """One long code stream decoded by all ranks: rank r decodes its contiguous span of frames (+ the halo that makes it exact,
    ``model.decode_span``) and rank ``dst`` receives the stitched waveform through one trailing all_gather (`gather_waveforms`); no
    collective inside the decode.  ``codes``: Mimi ``[B, nq, T]`` or SNAC's list of per-level code tensors (every rank holds the
    stream: it is a few hundred KB).  Returns the waveform on ``dst`` (None elsewhere); with one rank this is ``decode_span(0, T)``."""
    rank, ws = world()
    snac = isinstance(codes, (list, tuple))
    if snac:
        T = codes[-1].shape[1] * model.vq_strides[-1]
        mult = max(model.vq_strides)
    else:
        T, mult = codes.shape[-1], 1
    _, _, cs, ce = shard_span(T, rank, ws, multiple=mult)
    if ce > cs:
        y = model.decode_span(codes, cs, ce, noises=noises) if snac else model.decode_span(codes, cs, ce)
        piece = (y[0, :, 0] if snac else y[0, 0]).contiguous()
    else:
        piece = torch.zeros(0, device=model.device)
    parts = gather_waveforms([piece], [rank], ws, dst=dst)
    if parts is None:
        return None
    return torch.cat([p for p in parts if p is not None and p.numel()])
