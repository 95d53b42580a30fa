"""Host-visible KV cache with the reference's observable contract (lm/models/cache.py:104-176).

What callers can see -- and what tests/golden/cache_golden.npz (the reference class executed step by step) pins -- is:
``offset`` (rows written), the capacity ``keys.shape[2]`` (grows in 256-row blocks; a cache whose write pointer sits inside a
block is first cut back to the pointer, so capacities like 3 + 512 occur), prefix views from ``update_and_fetch`` / ``state``,
and ``trim`` as a pointer rewind.  The storage here is a pair of row slabs that own that policy; the Qwen3-TTS talker and the
Whisper decoder do not use this class on their hot path (their caches are pre-sized device buffers indexed by a device-side
length, tts/models/qwen3_tts/talker.py), it exists for callers that drive a cache themselves.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

BLOCK_ROWS = 256


def _capacity_after(rows_written: int, capacity: int, incoming: int, block: int) -> int:
    """Capacity rule of the reference: no change while the new rows fit; otherwise the slab is cut to the write pointer when
    that pointer is not on a block boundary, and ceil(incoming / block) fresh blocks are appended."""
    if capacity and rows_written + incoming <= capacity:
        return capacity
    kept = capacity if rows_written % block == 0 else rows_written
    return kept + -(-incoming // block) * block


class _RowSlab:
    """[B, heads, capacity, width] tensor addressed by a row pointer kept by the owner."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    @property
    def capacity(self) -> int:
        return 0 if self.buf is None else self.buf.shape[2]

    def resize(self, like: torch.Tensor, rows_written: int, capacity: int) -> None:
        B, H, _, W = like.shape
        fresh = torch.zeros(B, H, capacity, W, dtype=like.dtype, device=like.device)
        if self.buf is not None and rows_written:
            fresh[:, :, :rows_written] = self.buf[:, :, :rows_written]
        self.buf = fresh

    def write(self, at: int, rows: torch.Tensor) -> None:
        self.buf[:, :, at:at + rows.shape[2]] = rows

    def prefix(self, n: int) -> torch.Tensor:
        return self.buf[:, :, :n]


class KVCache:
    step = BLOCK_ROWS

    def __init__(self):
        self._k, self._v = _RowSlab(), _RowSlab()
        self.offset = 0

    # the reference exposes the raw buffers as attributes; some callers read ``cache.keys.shape[2]`` as the capacity
    @property
    def keys(self) -> Optional[torch.Tensor]:
        return self._k.buf

    @keys.setter
    def keys(self, t):
        self._k.buf = t

    @property
    def values(self) -> Optional[torch.Tensor]:
        return self._v.buf

    @values.setter
    def values(self, t):
        self._v.buf = t

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        incoming = keys.shape[2]
        want = _capacity_after(self.offset, self._k.capacity, incoming, self.step)
        if want != self._k.capacity:
            self._k.resize(keys, self.offset, want)
            self._v.resize(values, self.offset, want)
        self._k.write(self.offset, keys)
        self._v.write(self.offset, values)
        self.offset += incoming
        return self._k.prefix(self.offset), self._v.prefix(self.offset)

    def size(self) -> int:
        return self.offset

    @property
    def state(self):
        return self._k.prefix(self.offset), self._v.prefix(self.offset)

    @state.setter
    def state(self, kv):
        self._k.buf, self._v.buf = kv
        self.offset = self._k.capacity

    def is_trimmable(self) -> bool:
        return True

    def trim(self, n: int) -> int:
        cut = max(0, min(self.offset, n))
        self.offset -= cut
        return cut

    def empty(self) -> bool:
        return self._k.buf is None

    @property
    def nbytes(self) -> int:
        return sum(0 if s.buf is None else s.buf.numel() * s.buf.element_size() for s in (self._k, self._v))
