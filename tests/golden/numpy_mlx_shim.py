"""A NumPy stand-in for the handful of ``mlx.core`` primitives that the reference's ``dsp.py`` and ``stt/models/whisper/audio.py``
call, so that the REFERENCE'S OWN CODE can be executed in the build container (MLX has no Linux wheel here) to produce golden
vectors for the DSP front end (tests/golden/make_dsp_golden.py).  Only used to generate fixtures; never imported by tests at run
time, never by the product.

Semantics kept: float32 default dtype (python scalars are weak), int32 ``arange``, ``as_strided`` strides in ELEMENTS,
``x.at[idx].add(v)`` scatter-add, complex64 FFTs.  Everything else is NumPy's definition of the same-named function, which is also
what the MLX documentation specifies for these ops (pad, concatenate, where, tile, repeat, linspace, clip, ...)."""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np


class array(np.ndarray):
    """np.ndarray with the few MLX-only methods the reference uses."""

    def __new__(cls, obj=None, dtype=None):
        a = np.asarray(obj)
        if dtype is None and a.dtype == np.float64:
            dtype = np.float32
        if dtype is None and a.dtype == np.int64:
            dtype = np.int32
        if dtype is None and a.dtype == np.complex128:
            dtype = np.complex64
        return np.asarray(a, dtype=dtype).view(cls)

    class _At:
        def __init__(self, a, idx):
            self.a, self.idx = a, idx

        def add(self, v):
            out = np.array(self.a, copy=True)
            np.add.at(out, np.asarray(self.idx), np.asarray(v))
            return out.view(array)

    class _AtProxy:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            return array._At(self.a, idx)

    @property
    def at(self):
        return array._AtProxy(self)

    def abs(self):
        return np.abs(self).view(array)

    def square(self):
        return np.square(self).view(array)

    def log10(self):
        return np.log10(self).view(array)

    def log(self):
        return np.log(self).view(array)

    def moveaxis(self, a, b):
        return np.moveaxis(self, a, b).view(array)

    def exp(self):
        return np.exp(self).view(array)


def _w(x):
    return x.view(array) if isinstance(x, np.ndarray) else x


def _wrap(fn):
    def f(*a, **k):
        return _w(fn(*a, **k))
    return f


def build():
    core = types.ModuleType("mlx.core")
    core.array = array
    core.float32, core.float64, core.int32, core.complex64, core.float16 = np.float32, np.float64, np.int32, np.complex64, np.float16
    core.pi = np.pi
    core.cpu = "cpu"
    core.stream = lambda *_a, **_k: contextlib.nullcontext()
    core.zeros = lambda shape, dtype=np.float32: np.zeros(shape, dtype=dtype).view(array)
    core.ones = lambda shape, dtype=np.float32: np.ones(shape, dtype=dtype).view(array)
    core.zeros_like = lambda a: np.zeros_like(a).view(array)

    def arange(*args, dtype=None):
        a = np.arange(*args)
        if dtype is None:
            dtype = np.int32 if np.issubdtype(a.dtype, np.integer) else np.float32
        return a.astype(dtype).view(array)
    core.arange = arange
    core.linspace = lambda a, b, n=50, dtype=np.float32: np.linspace(a, b, n).astype(dtype).view(array)

    def pad(x, pad_width, mode="constant", constant_values=0):
        if isinstance(pad_width, tuple) and len(pad_width) == 2 and not isinstance(pad_width[0], (tuple, list)):
            pad_width = [(0, 0)] * (np.ndim(x) - 1) + [tuple(pad_width)]          # mx.pad(x, (lo, hi)) pads the last... MLX: all axes;
            if np.ndim(x) == 1:
                pad_width = [tuple(pad_width[-1])]
        return np.pad(np.asarray(x), pad_width, mode="constant", constant_values=constant_values).view(array)
    core.pad = pad
    for name in ("concatenate", "maximum", "minimum", "expand_dims", "cos", "sin", "where", "tile", "repeat", "log", "exp", "sum", "power",
                 "mean", "matmul", "clip", "abs", "sqrt", "stack", "transpose", "reshape", "square", "log10"):
        setattr(core, name, _wrap(getattr(np, name)))

    def as_strided(x, shape, strides, offset=0):
        a = np.ascontiguousarray(np.asarray(x)).reshape(-1)[offset:]
        item = a.dtype.itemsize
        v = np.lib.stride_tricks.as_strided(a, shape=tuple(shape), strides=tuple(int(s) * item for s in strides), writeable=False)
        return np.array(v, copy=True).view(array)
    core.as_strided = as_strided
    fft = types.ModuleType("mlx.core.fft")
    fft.rfft = lambda x, n=None, axis=-1: np.fft.rfft(np.asarray(x), n=n, axis=axis).astype(np.complex64).view(array)
    fft.irfft = lambda x, n=None, axis=-1: np.fft.irfft(np.asarray(x), n=n, axis=axis).astype(np.float32).view(array)
    core.fft = fft
    rnd = types.ModuleType("mlx.core.random")
    rnd.normal = lambda shape=(), **_k: np.zeros(shape, dtype=np.float32).view(array)        # dither is off in every golden case
    core.random = rnd
    core.eval = lambda *_a, **_k: None
    mlx = types.ModuleType("mlx")
    mlx.core = core
    return mlx, core


def install():
    mlx, core = build()
    sys.modules["mlx"] = mlx
    sys.modules["mlx.core"] = core
    sys.modules["mlx.core.fft"] = core.fft
    sys.modules["mlx.core.random"] = core.random
    return core
