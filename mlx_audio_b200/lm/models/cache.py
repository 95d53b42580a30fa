"""KVCache (reference: lm/models/cache.py:104-176) on torch tensors.

Same contract: K/V ``[B, n_kv, S, hd]`` grown in blocks of ``step`` = 256 rows, ``update_and_fetch`` writes the new rows
at ``offset`` and returns views of the valid prefix, ``trim`` rewinds.  The Qwen3-TTS talker keeps its own device-resident
cache (tts/models/qwen3_tts/talker.py) so that the KV length never visits the host; this class is the host-visible
equivalent used by callers that manage caches themselves (and by the Whisper decoder's API surface).
"""
from __future__ import annotations

import torch


class KVCache:
    step = 256

    def __init__(self):
        self.keys = None
        self.values = None
        self.offset = 0

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor):
        prev = self.offset
        n_new = keys.shape[2]
        if self.keys is None or (prev + n_new) > self.keys.shape[2]:
            B, n_kv, _, kd = keys.shape
            vd = values.shape[3]
            n_steps = (self.step + n_new - 1) // self.step
            new_k = torch.zeros(B, n_kv, n_steps * self.step, kd, dtype=keys.dtype, device=keys.device)
            new_v = torch.zeros(B, n_kv, n_steps * self.step, vd, dtype=values.dtype, device=values.device)
            if self.keys is not None:
                if prev % self.step != 0:
                    self.keys = self.keys[..., :prev, :]
                    self.values = self.values[..., :prev, :]
                self.keys = torch.cat([self.keys, new_k], dim=2)
                self.values = torch.cat([self.values, new_v], dim=2)
            else:
                self.keys, self.values = new_k, new_v
        self.offset += n_new
        self.keys[..., prev:self.offset, :] = keys
        self.values[..., prev:self.offset, :] = values
        return self.keys[..., :self.offset, :], self.values[..., :self.offset, :]

    def size(self):
        return self.offset

    @property
    def state(self):
        if self.offset == self.keys.shape[2]:
            return self.keys, self.values
        return self.keys[..., :self.offset, :], self.values[..., :self.offset, :]

    @state.setter
    def state(self, v):
        self.keys, self.values = v
        self.offset = self.keys.shape[2]

    def is_trimmable(self):
        return True

    def trim(self, n):
        n = min(self.offset, n)
        self.offset -= n
        return n

    def empty(self):
        return self.keys is None

    @property
    def nbytes(self):
        if self.keys is None:
            return 0
        return self.keys.numel() * self.keys.element_size() + self.values.numel() * self.values.element_size()
