"""NumPy stand-ins for ``mlx.core`` / ``mlx.nn`` / ``mlx.utils`` wide enough to EXECUTE THE REFERENCE'S OWN MODEL CODE (Whisper,
Qwen3-TTS talker / code predictor / speech-tokenizer decoder, SNAC, Mimi, Kokoro) in the build container, where MLX has no wheel.
Used only by the ``make_*_golden.py`` generators in this directory; never imported by the test-suite at run time, never by the
product.

What is restated here is MLX's documented semantics for each primitive, written independently of
``oracle/nn.py`` (plain NumPy loops over kernel taps rather than torch's conv routines) so that an error in one is unlikely to be
repeated in the other:

* arrays are channels-last; ``conv1d(x[N,L,Cin], w[Cout,K,Cin/g])`` is a cross-correlation with symmetric zero padding;
  ``conv_transpose1d`` scatters ``y[s*i + d*k - p] += x[i] . w[:, k, :]`` with no kernel flip and positional argument order
  ``(stride, padding, dilation, output_padding, groups)``;
* ``nn.Module`` keeps parameters as attributes; names beginning with ``_`` are not parameters; lists index as ``name.0.…``;
* ``x += y`` re-binds (MLX arrays are immutable) -- the array class below returns a new array from every in-place operator;
* ``precise=True`` maps every floating dtype to float64, which turns the run into a statement of the reference's ALGORITHM
  (composition, shapes, masks, cache handling, padding rules) with rounding out of the picture.
"""
from __future__ import annotations

import contextlib
import math
import sys
import types

import numpy as np

_FLOAT = np.float64


class array(np.ndarray):
    def __new__(cls, obj=None, dtype=None):
        a = np.asarray(obj)
        if dtype is None:
            if a.dtype.kind == "f":
                dtype = _FLOAT
            elif a.dtype == np.int64:
                dtype = np.int32
        return np.asarray(a, dtype=dtype).view(cls)

    # MLX arrays are immutable: in-place operators re-bind
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def astype(self, dtype, *a, **k):
        return np.ndarray.astype(self, _dt(dtype)).view(array)

    def flatten(self, start_axis=0, end_axis=-1):
        nd = self.ndim
        s, e = start_axis % nd, end_axis % nd
        return self.reshape(*self.shape[:s], -1, *self.shape[e + 1:])

    def square(self):
        return np.square(self)

    def sqrt(self):
        return np.sqrt(self)

    def rsqrt(self):
        return 1.0 / np.sqrt(self)

    def log10(self):
        return np.log10(self).view(array)

    def abs(self):
        return np.abs(self)

    def exp(self):
        return np.exp(self)

    def log(self):
        return np.log(self)

    def reciprocal(self):
        return 1.0 / self

    def logsumexp(self, axis=None, keepdims=False):
        m = np.max(np.asarray(self), axis=axis, keepdims=True)
        r = np.log(np.sum(np.exp(np.asarray(self) - m), axis=axis, keepdims=True)) + m
        return _w(r if keepdims else np.squeeze(r, axis=axis))

    def moveaxis(self, a, b):
        return np.moveaxis(self, a, b)

    def split(self, n, axis=0):
        return [_w(p) for p in np.split(np.asarray(self), n, axis=axis)]

    class _At:
        def __init__(self, a, idx):
            self.a, self.idx = a, idx

        def add(self, v):
            out = np.array(self.a, copy=True)
            np.add.at(out, self.idx, np.asarray(v))
            return out.view(array)

    class _AtProxy:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            return array._At(self.a, idx)

    @property
    def at(self):
        return array._AtProxy(self)


def _red(name):
    def f(self, axis=None, keepdims=False, **k):
        if isinstance(axis, list):
            axis = tuple(axis)
        return _w(getattr(np.ndarray, name)(self, axis=axis, keepdims=keepdims, **k))
    return f


for _n in ("sum", "mean", "var", "std", "max", "min", "prod"):     # MLX accepts a list of axes
    setattr(array, _n, _red(_n))


def _dt(d):
    if d in (np.float16, np.float32, np.float64, "float16", "float32", "bfloat16") or getattr(d, "__name__", "") == "bfloat16":
        return _FLOAT
    return d


def _w(x):
    if isinstance(x, np.ndarray):
        return x.view(array)
    if isinstance(x, (np.floating, np.integer, np.bool_)):
        return np.asarray(x).view(array)
    if isinstance(x, (list, tuple)) and x and isinstance(x[0], np.ndarray):
        return type(x)(_w(e) for e in x)
    return x


def _wrap(fn):
    def f(*a, **k):
        k.pop("stream", None)
        if isinstance(k.get("axis"), list):
            k["axis"] = tuple(k["axis"])
        if isinstance(k.get("axes"), list):
            k["axes"] = tuple(k["axes"])
        return _w(fn(*a, **k))
    return f


def conv1d(x, w, stride=1, padding=0, dilation=1, groups=1, **_k):
    x, w = np.asarray(x), np.asarray(w)
    n, l, cin = x.shape
    cout, k, cin_g = w.shape
    assert cin_g * groups == cin and cout % groups == 0, (x.shape, w.shape, groups)
    xp = np.zeros((n, l + 2 * padding, cin), dtype=x.dtype)
    xp[:, padding:padding + l] = x
    lout = (l + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    y = np.zeros((n, lout, cout), dtype=np.result_type(x, w))
    og = cout // groups
    for g in range(groups):
        xs = xp[:, :, g * cin_g:(g + 1) * cin_g]
        for t in range(k):
            seg = xs[:, t * dilation: t * dilation + (lout - 1) * stride + 1: stride]
            y[:, :, g * og:(g + 1) * og] += seg @ w[g * og:(g + 1) * og, t, :].T
    return y.view(array)


def conv_transpose1d(x, w, stride=1, padding=0, dilation=1, output_padding=0, groups=1, **_k):
    x, w = np.asarray(x), np.asarray(w)
    n, l, cin = x.shape
    cout, k, cin_g = w.shape
    assert cin_g * groups == cin and cout % groups == 0, (x.shape, w.shape, groups)
    full = (l - 1) * stride + dilation * (k - 1) + 1 + output_padding
    y = np.zeros((n, full, cout), dtype=np.result_type(x, w))
    og = cout // groups
    pos = np.arange(l) * stride
    for g in range(groups):
        xs = x[:, :, g * cin_g:(g + 1) * cin_g]
        for t in range(k):
            y[:, pos + t * dilation, g * og:(g + 1) * og] += xs @ w[g * og:(g + 1) * og, t, :].T
    return y[:, padding: full - padding].view(array)


def softmax(x, axis=-1, precise=False, **_k):
    x = np.asarray(x)
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return (e / np.sum(e, axis=axis, keepdims=True)).view(array)


def sdpa(q, k, v, *, scale, mask=None, **_k):
    """mx.fast.scaled_dot_product_attention: q [B,Hq,Tq,D], k/v [B,Hkv,Tk,D] (grouped heads share a kv head in blocks of Hq/Hkv);
    mask None | "causal" (aligned to the LAST key) | boolean keep-mask | additive array, broadcast to [B,Hq,Tq,Tk]."""
    q, k, v = np.asarray(q), np.asarray(k), np.asarray(v)
    hq, hkv = q.shape[1], k.shape[1]
    if hq != hkv:
        k, v = np.repeat(k, hq // hkv, axis=1), np.repeat(v, hq // hkv, axis=1)
    s = (q * scale) @ np.swapaxes(k, -1, -2)
    tq, tk = s.shape[-2:]
    if isinstance(mask, str):
        assert mask == "causal"
        keep = np.arange(tk)[None, :] <= (np.arange(tq)[:, None] + (tk - tq))
        s = np.where(keep, s, -np.inf)
    elif mask is not None:
        m = np.asarray(mask)
        s = np.where(m, s, -np.inf) if m.dtype == np.bool_ else s + m
    return (np.asarray(softmax(s, -1)) @ v).view(array)


def build(precise=True):
    global _FLOAT
    _FLOAT = np.float64 if precise else np.float32
    core = types.ModuleType("mlx.core")
    core.array = array
    core.Dtype = type
    core.float32 = core.float16 = core.float64 = _FLOAT
    core.bfloat16 = _FLOAT
    core.int32, core.int64, core.uint32, core.int16, core.uint8, core.bool_, core.complex64 = np.int32, np.int64, np.uint32, np.int16, np.uint8, np.bool_, np.complex128 if precise else np.complex64
    core.pi, core.inf, core.newaxis = np.pi, np.inf, None
    core.cpu, core.gpu = "cpu", "gpu"
    core.stream = lambda *_a, **_k: contextlib.nullcontext()
    core.eval = lambda *_a, **_k: None
    core.async_eval = lambda *_a, **_k: None
    core.nan = np.nan
    core.clear_cache = lambda *_a, **_k: None
    core.get_peak_memory = lambda: 0
    core.compile = lambda f=None, *a, **k: f if callable(f) else (lambda g: g)
    core.finfo = lambda d: np.finfo(np.float32)
    core.zeros = lambda shape, dtype=None: np.zeros(shape, dtype=_dt(dtype) if dtype is not None else _FLOAT).view(array)
    core.ones = lambda shape, dtype=None: np.ones(shape, dtype=_dt(dtype) if dtype is not None else _FLOAT).view(array)
    core.full = lambda shape, vals, dtype=None: np.full(shape, vals, dtype=_dt(dtype) if dtype is not None else (_FLOAT if isinstance(vals, float) else None)).view(array)
    core.zeros_like = lambda a: np.zeros_like(np.asarray(a)).view(array)
    core.ones_like = lambda a: np.ones_like(np.asarray(a)).view(array)

    def arange(*args, dtype=None):
        a = np.arange(*args)
        if dtype is None:
            dtype = np.int32 if np.issubdtype(a.dtype, np.integer) else _FLOAT
        return a.astype(_dt(dtype)).view(array)
    core.arange = arange
    core.linspace = lambda a, b, num=50, dtype=None: np.linspace(a, b, num).astype(_FLOAT).view(array)

    def pad(x, pad_width, mode="constant", constant_values=0, **_k):
        x = np.asarray(x)
        if isinstance(pad_width, int):
            pad_width = [(pad_width, pad_width)] * x.ndim
        elif len(pad_width) == 2 and not isinstance(pad_width[0], (tuple, list)):
            pad_width = [tuple(pad_width)] * x.ndim
        if mode == "edge":
            return np.pad(x, pad_width, mode="edge").view(array)
        return np.pad(x, pad_width, mode="constant", constant_values=constant_values).view(array)
    core.pad = pad
    for name in ("concatenate", "maximum", "minimum", "expand_dims", "cos", "sin", "where", "tile", "repeat", "log", "exp", "sum", "power",
                 "mean", "matmul", "clip", "abs", "sqrt", "stack", "transpose", "reshape", "square", "log10", "broadcast_to", "take_along_axis",
                 "swapaxes", "roll", "cumsum", "einsum", "reciprocal", "take", "triu", "tril", "tanh", "squeeze", "floor", "ceil", "var", "round",
                 "real", "imag", "nan_to_num", "arctan2", "argmax", "argmin", "max", "min", "sort", "argsort", "logical_and", "logical_or",
                 "logical_not", "isnan", "isinf", "moveaxis", "outer", "cumprod", "prod", "std", "sign", "negative", "floor_divide",
                 "remainder", "equal", "not_equal", "greater", "less", "any", "all", "flip", "diag", "eye", "identity", "log1p", "expm1", "array_equal"):
        setattr(core, name, _wrap(getattr(np, name)))
    core.concat = core.concatenate
    core.contiguous = lambda x, *a, **k: np.ascontiguousarray(x).view(array)
    core.stop_gradient = lambda x: x
    core.rsqrt = lambda x: (1.0 / np.sqrt(np.asarray(x))).view(array)
    core.sigmoid = lambda x: (1.0 / (1.0 + np.exp(-np.asarray(x)))).view(array)
    core.softmax = softmax
    core.addmm = lambda c, a, b, alpha=1.0, beta=1.0: (beta * np.asarray(c) + alpha * (np.asarray(a) @ np.asarray(b))).view(array)
    core.logsumexp = lambda x, axis=None, keepdims=False: _w(np.log(np.sum(np.exp(x - np.max(x, axis=axis, keepdims=True)), axis=axis, keepdims=True)) + np.max(x, axis=axis, keepdims=True)) if keepdims else _w(np.log(np.sum(np.exp(x - np.max(x, axis=axis, keepdims=True)), axis=axis)) + np.max(x, axis=axis))

    def split(x, indices_or_sections, axis=0, **_k):
        return [_w(p) for p in np.split(np.asarray(x), indices_or_sections, axis=axis)]
    core.split = split

    def put_along_axis(a, indices, values, axis=None, **_k):
        out = np.array(np.asarray(a), copy=True)
        idx = np.asarray(indices).astype(np.int64)
        np.put_along_axis(out, idx, np.broadcast_to(np.asarray(values, dtype=out.dtype), idx.shape) if np.ndim(values) else np.asarray(values, dtype=out.dtype), axis)
        return out.view(array)
    core.put_along_axis = put_along_axis
    core.argpartition = lambda a, kth, axis=-1, **_k: np.argpartition(np.asarray(a), kth, axis=axis).astype(np.int32).view(array)

    def erf(x):
        from scipy.special import erf as _erf
        return _erf(np.asarray(x)).view(array)
    core.erf = erf
    core.conv1d = conv1d
    core.conv_transpose1d = conv_transpose1d

    def as_strided(x, shape, strides, offset=0):
        a = np.ascontiguousarray(np.asarray(x)).reshape(-1)[offset:]
        item = a.dtype.itemsize
        v = np.lib.stride_tricks.as_strided(a, shape=tuple(shape), strides=tuple(int(s) * item for s in strides), writeable=False)
        return np.array(v, copy=True).view(array)
    core.as_strided = as_strided
    fft = types.ModuleType("mlx.core.fft")
    cdt = np.complex128 if precise else np.complex64
    fft.rfft = lambda x, n=None, axis=-1, **_k: np.fft.rfft(np.asarray(x), n=n, axis=axis).astype(cdt).view(array)
    fft.irfft = lambda x, n=None, axis=-1, **_k: np.fft.irfft(np.asarray(x), n=n, axis=axis).astype(_FLOAT).view(array)
    core.fft = fft

    # random draws come from a queue the generator fills (the oracle takes the same arrays as injected noise)
    rnd = types.ModuleType("mlx.core.random")
    rnd.queue = []
    rnd.state = []
    rnd.strict = False           # generators switch this on once the modules are built

    def _draw(kind, shape, **_k):
        if not rnd.queue:
            if rnd.strict:
                raise RuntimeError(f"mx.random.{kind}{tuple(shape)} called with an empty injection queue")
            return np.zeros(shape, dtype=_FLOAT).view(array)         # parameter initialisers: every value is overwritten afterwards
        want, a = rnd.queue.pop(0)
        if callable(a):
            a = a(tuple(shape))
        assert want == kind and tuple(a.shape) == tuple(shape), (want, kind, a.shape, shape)
        return np.asarray(a, dtype=_FLOAT).view(array)
    rnd.normal = lambda shape=(), dtype=None, loc=0.0, scale=1.0, key=None, **_k: _draw("normal", shape) * scale + loc
    rnd.uniform = lambda low=0.0, high=1.0, shape=(), dtype=None, key=None, **_k: _draw("uniform", shape) * (high - low) + low
    def categorical(logits, axis=-1, shape=None, num_samples=None, key=None, **_k):
        """Inverse CDF in index order driven by an injected uniform per row (the convention oracle/qwen3.py and the CUDA sampler use)."""
        lg = np.asarray(logits, dtype=np.float64)
        assert axis in (-1, lg.ndim - 1) and shape is None and num_samples is None
        us = np.asarray(_draw("categorical", lg.shape[:-1]), dtype=np.float64).reshape(-1)
        rows = lg.reshape(-1, lg.shape[-1])
        out = np.zeros(rows.shape[0], dtype=np.int32)
        for r in range(rows.shape[0]):
            w = np.exp(rows[r] - rows[r].max())
            cum = np.cumsum(w)
            live = np.nonzero(w > 0)[0]
            hit = live[cum[live] > us[r] * cum[-1]]
            out[r] = hit[0] if hit.size else live[-1]
        return out.reshape(lg.shape[:-1]).view(array)
    rnd.categorical = categorical
    rnd.seed = lambda *_a, **_k: None
    rnd.key = lambda *_a, **_k: None
    core.random = rnd

    fast = types.ModuleType("mlx.core.fast")
    fast.scaled_dot_product_attention = sdpa
    fast.rms_norm = lambda x, w, eps: _w(np.asarray(x) / np.sqrt(np.mean(np.square(np.asarray(x)), -1, keepdims=True) + eps) * (1.0 if w is None else np.asarray(w)))

    def layer_norm(x, w, b, eps):
        x = np.asarray(x)
        y = (x - x.mean(-1, keepdims=True)) / np.sqrt(x.var(-1, keepdims=True) + eps)
        if w is not None:
            y = y * np.asarray(w)
        if b is not None:
            y = y + np.asarray(b)
        return y.view(array)
    fast.layer_norm = layer_norm

    def rope(x, dims, *, traditional, base, scale, offset, freqs=None, **_k):
        x = np.asarray(x)
        t = x.shape[-2]
        half = dims // 2
        inv = (1.0 / np.asarray(freqs)) if freqs is not None else base ** (-np.arange(half, dtype=np.float64) * 2.0 / dims)
        ang = (np.arange(t, dtype=np.float64) + offset)[:, None] * scale * inv[None, :]
        c, s = np.cos(ang), np.sin(ang)
        out = np.array(x, copy=True)
        if traditional:
            a, b = x[..., 0:dims:2], x[..., 1:dims:2]
            out[..., 0:dims:2], out[..., 1:dims:2] = a * c - b * s, a * s + b * c
        else:
            a, b = x[..., :half], x[..., half:dims]
            out[..., :half], out[..., half:dims] = a * c - b * s, a * s + b * c
        return out.astype(x.dtype).view(array)
    fast.rope = rope
    core.fast = fast

    mlx = types.ModuleType("mlx")
    mlx.core = core
    nn = _build_nn(core)
    utils = _build_utils(core)
    mlx.nn, mlx.utils = nn, utils
    return mlx, core, nn, utils


# ----------------------------------------------------------------------------------------------------------------- nn
def _is_param_container(v):
    return isinstance(v, (np.ndarray, Module, list, tuple, dict))


class Module:
    def __init__(self):
        self.training = True

    def __call__(self, *a, **k):
        raise NotImplementedError

    # -- parameter tree (mlx.nn.Module.parameters(): arrays reachable through attributes not starting with "_") --
    def _items(self):
        return [(k, v) for k, v in self.__dict__.items() if not k.startswith("_") and k != "training" and _is_param_container(v)]

    def parameters(self):
        def walk(v):
            if isinstance(v, Module):
                out = {}
                for k, x in v._items():
                    w = walk(x)
                    if w is None or (isinstance(w, (dict, list)) and len(w) == 0):
                        continue
                    out[k] = w
                return out
            if isinstance(v, dict):
                return {k: walk(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [walk(x) for x in v]
            if isinstance(v, np.ndarray):
                return v
            return None
        return walk(self)

    trainable_parameters = parameters

    def children(self):
        return {k: v for k, v in self._items() if not isinstance(v, np.ndarray)}

    def named_modules(self):
        out = []

        def walk(prefix, v):
            if isinstance(v, Module):
                out.append((prefix, v))
                for k, x in v._items():
                    walk(f"{prefix}.{k}" if prefix else k, x)
            elif isinstance(v, dict):
                for k, x in v.items():
                    walk(f"{prefix}.{k}", x)
            elif isinstance(v, (list, tuple)):
                for i, x in enumerate(v):
                    walk(f"{prefix}.{i}", x)
        walk("", self)
        return out

    def modules(self):
        return [m for _, m in self.named_modules()]

    def apply_to_modules(self, fn):
        for n, m in self.named_modules():
            fn(n, m)
        return self

    def eval(self):
        for m in self.modules():
            m.training = False
        return self

    def train(self, mode=True):
        for m in self.modules():
            m.training = mode
        return self

    def freeze(self, *a, **k):
        return self

    def unfreeze(self, *a, **k):
        return self

    def set_dtype(self, *a, **k):
        return self

    def update(self, params):
        def walk(dst, src):
            if isinstance(src, dict):
                for k, v in src.items():
                    cur = dst[k] if isinstance(dst, dict) else getattr(dst, k)
                    if isinstance(v, np.ndarray):
                        if isinstance(dst, dict):
                            dst[k] = _w(v)
                        else:
                            setattr(dst, k, _w(v))
                    else:
                        walk(cur, v)
            elif isinstance(src, list):
                for i, v in enumerate(src):
                    if isinstance(v, np.ndarray):
                        dst[i] = _w(v)
                    else:
                        walk(dst[i], v)
        walk(self, params)
        return self

    def load_weights(self, weights, strict=True):
        if isinstance(weights, dict):
            weights = list(weights.items())
        have = dict(flat_parameters(self))
        for k, v in weights:
            if k not in have:
                if strict:
                    raise ValueError(f"Received parameters not in model: {k}")
                continue
            assert tuple(have[k].shape) == tuple(v.shape) or not strict, (k, have[k].shape, v.shape)
            set_parameter(self, k, v)
        if strict and set(have) - {k for k, _ in weights}:
            raise ValueError(f"Missing parameters: {sorted(set(have) - {k for k, _ in weights})[:5]}")
        return self


def flat_parameters(m):
    out = []

    def walk(prefix, v):
        if isinstance(v, dict):
            for k, x in v.items():
                walk(f"{prefix}.{k}" if prefix else k, x)
        elif isinstance(v, list):
            for i, x in enumerate(v):
                walk(f"{prefix}.{i}" if prefix else str(i), x)
        elif isinstance(v, np.ndarray):
            out.append((prefix, v))
    walk("", m.parameters())
    return out


def set_parameter(m, name, value):
    parts = name.split(".")
    cur = m
    for p in parts[:-1]:
        cur = cur[int(p)] if isinstance(cur, (list, tuple)) else (cur[p] if isinstance(cur, dict) else getattr(cur, p))
    v = np.asarray(value, dtype=_FLOAT if np.asarray(value).dtype.kind == "f" else None).view(array)
    if isinstance(cur, list):
        cur[int(parts[-1])] = v
    elif isinstance(cur, dict):
        cur[parts[-1]] = v
    else:
        setattr(cur, parts[-1], v)


def _build_nn(mx):
    nn = types.ModuleType("mlx.nn")
    nn.Module = Module
    Z = lambda *s: np.zeros(s, dtype=_FLOAT).view(array)          # noqa: E731  (values are overwritten by the generator)
    O = lambda *s: np.ones(s, dtype=_FLOAT).view(array)           # noqa: E731

    class Linear(Module):
        def __init__(self, input_dims, output_dims, bias=True):
            super().__init__()
            self.weight = Z(output_dims, input_dims)
            if bias:
                self.bias = Z(output_dims)

        def __call__(self, x):
            y = np.asarray(x) @ np.asarray(self.weight).T
            return _w(y + np.asarray(self.bias) if "bias" in self.__dict__ else y)

    class Embedding(Module):
        def __init__(self, num_embeddings, dims):
            super().__init__()
            self.weight = Z(num_embeddings, dims)

        def __call__(self, x):
            return _w(np.asarray(self.weight)[np.asarray(x)])

        def as_linear(self, x):
            return _w(np.asarray(x) @ np.asarray(self.weight).T)

    class LayerNorm(Module):
        def __init__(self, dims, eps=1e-5, affine=True, bias=True):
            super().__init__()
            self.eps, self.dims = eps, dims
            if affine:
                self.weight = O(dims)
                if bias:
                    self.bias = Z(dims)

        def __call__(self, x):
            return mx.fast.layer_norm(x, self.__dict__.get("weight"), self.__dict__.get("bias"), self.eps)

    class RMSNorm(Module):
        def __init__(self, dims, eps=1e-5):
            super().__init__()
            self.weight, self.eps = O(dims), eps

        def __call__(self, x):
            return mx.fast.rms_norm(x, self.weight, self.eps)

    class InstanceNorm(Module):
        def __init__(self, dims, eps=1e-5, affine=False):
            super().__init__()
            self.eps, self.dims = eps, dims
            if affine:
                self.weight, self.bias = O(dims), Z(dims)

        def __call__(self, x):
            x = np.asarray(x)
            ax = tuple(range(1, x.ndim - 1))
            y = (x - x.mean(ax, keepdims=True)) / np.sqrt(x.var(ax, keepdims=True) + self.eps)
            if "weight" in self.__dict__:
                y = y * np.asarray(self.weight) + np.asarray(self.bias)
            return _w(y)

    class Conv1d(Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
            super().__init__()
            self.weight = Z(out_channels, kernel_size, in_channels // groups)
            if bias:
                self.bias = Z(out_channels)
            self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups

        def __call__(self, x):
            y = conv1d(x, self.weight, self.stride, self.padding, self.dilation, self.groups)
            return _w(y + np.asarray(self.bias)) if "bias" in self.__dict__ else y

    class ConvTranspose1d(Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, output_padding=0, bias=True):
            super().__init__()
            self.weight = Z(out_channels, kernel_size, in_channels)
            if bias:
                self.bias = Z(out_channels)
            self.stride, self.padding, self.dilation, self.output_padding = stride, padding, dilation, output_padding

        def __call__(self, x):
            y = conv_transpose1d(x, self.weight, self.stride, self.padding, self.dilation, self.output_padding)
            return _w(y + np.asarray(self.bias)) if "bias" in self.__dict__ else y

    class Sequential(Module):
        def __init__(self, *modules):
            super().__init__()
            self.layers = list(modules)

        def __call__(self, x):
            for m in self.layers:
                x = m(x)
            return x

    class Dropout(Module):
        def __init__(self, p=0.5):
            super().__init__()
            self.p = p

        def __call__(self, x):
            if self.training and self.p > 0:
                raise RuntimeError("Dropout called in training mode: the golden run must be in eval mode")
            return x

    class Identity(Module):
        def __init__(self, *a, **k):
            super().__init__()

        def __call__(self, x, *a, **k):
            return x

    def gelu(x):
        x = np.asarray(x)
        return _w(x * (1.0 + np.asarray(mx.erf(x / math.sqrt(2.0)))) / 2.0)

    def gelu_approx(x):
        x = np.asarray(x)
        return _w(0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3))))

    def gelu_fast_approx(x):
        x = np.asarray(x)
        return _w(x / (1.0 + np.exp(-1.702 * x)))

    silu = lambda x: _w(np.asarray(x) / (1.0 + np.exp(-np.asarray(x))))                      # noqa: E731
    relu = lambda x: _w(np.maximum(np.asarray(x), 0))                                         # noqa: E731
    elu = lambda x, alpha=1.0: _w(np.where(np.asarray(x) > 0, np.asarray(x), alpha * (np.exp(np.minimum(np.asarray(x), 0)) - 1.0)))   # noqa: E731
    leaky_relu = lambda x, negative_slope=0.01: _w(np.where(np.asarray(x) > 0, np.asarray(x), negative_slope * np.asarray(x)))       # noqa: E731
    nn.gelu, nn.gelu_approx, nn.gelu_fast_approx, nn.silu, nn.relu, nn.elu, nn.leaky_relu = gelu, gelu_approx, gelu_fast_approx, silu, relu, elu, leaky_relu
    nn.tanh = lambda x: _w(np.tanh(np.asarray(x)))
    nn.sigmoid = mx.sigmoid
    nn.softmax = softmax
    nn.log_softmax = lambda x, axis=-1: _w(np.asarray(x) - np.max(np.asarray(x), axis=axis, keepdims=True) - np.log(np.sum(np.exp(np.asarray(x) - np.max(np.asarray(x), axis=axis, keepdims=True)), axis=axis, keepdims=True)))

    def _act_module(name, fn):
        def __call__(self, x):
            return fn(x)
        return type(name, (Module,), {"__call__": __call__})
    nn.GELU = type("GELU", (Module,), {"__init__": lambda self, approx="none": (Module.__init__(self), setattr(self, "_fn", {"none": gelu, "precise": gelu_approx, "tanh": gelu_approx, "fast": gelu_fast_approx}[approx]))[0],
                                       "__call__": lambda self, x: self._fn(x)})
    nn.SiLU, nn.ReLU, nn.Tanh, nn.Sigmoid, nn.ELU = (_act_module("SiLU", silu), _act_module("ReLU", relu), _act_module("Tanh", nn.tanh),
                                                      _act_module("Sigmoid", mx.sigmoid), _act_module("ELU", elu))

    class LeakyReLU(Module):
        def __init__(self, negative_slope=0.01):
            super().__init__()
            self._slope = negative_slope

        def __call__(self, x):
            return leaky_relu(x, self._slope)

    class Upsample(Module):
        def __init__(self, scale_factor, mode="nearest", align_corners=False):
            super().__init__()
            self.scale_factor, self.mode = scale_factor, mode
            assert mode == "nearest"

        def __call__(self, x):
            x = np.asarray(x)                                  # [N, L, C] -> nearest: out[i] = in[floor(i / scale)]
            s = self.scale_factor[0] if isinstance(self.scale_factor, (tuple, list)) else self.scale_factor
            lout = int(x.shape[1] * s)
            idx = np.floor(np.arange(lout) / s).astype(np.int64)
            return _w(x[:, idx])

    class RoPE(Module):
        def __init__(self, dims, traditional=False, base=10000, scale=1.0):
            super().__init__()
            self.dims, self.traditional, self.base, self.scale = dims, traditional, base, scale

        def __call__(self, x, offset=0):
            return mx.fast.rope(x, self.dims, traditional=self.traditional, base=self.base, scale=self.scale, offset=offset)

    class MultiHeadAttention(Module):
        @staticmethod
        def create_additive_causal_mask(N, dtype=None):
            i = np.arange(N)
            return _w((i[:, None] < i[None]).astype(_FLOAT) * float(np.finfo(np.float32).min))

    def quantize(*_a, **_k):
        raise RuntimeError("quantisation is not part of the golden runs")
    nn.quantize = quantize
    for c in (Linear, Embedding, LayerNorm, RMSNorm, InstanceNorm, Conv1d, ConvTranspose1d, Sequential, Dropout, Identity, LeakyReLU, Upsample, RoPE,
              MultiHeadAttention):
        setattr(nn, c.__name__, c)
    return nn


def _build_utils(mx):
    u = types.ModuleType("mlx.utils")

    def tree_flatten(tree, prefix="", is_leaf=None):
        out = []

        def walk(p, v):
            if isinstance(v, dict):
                for k, x in v.items():
                    walk(f"{p}.{k}" if p else k, x)
            elif isinstance(v, (list, tuple)):
                for i, x in enumerate(v):
                    walk(f"{p}.{i}" if p else str(i), x)
            else:
                out.append((p, v))
        walk(prefix, tree)
        return out

    def tree_unflatten(items):
        root = {}
        for k, v in items:
            cur = root
            parts = k.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            cur[parts[-1]] = v

        def fix(d):
            if isinstance(d, dict):
                d = {k: fix(v) for k, v in d.items()}
                if d and all(k.isdigit() for k in d):
                    return [d[str(i)] for i in range(len(d))]
            return d
        return fix(root)

    def tree_map(fn, tree, *rest, is_leaf=None):
        if isinstance(tree, dict):
            return {k: tree_map(fn, v, *(r[k] for r in rest)) for k, v in tree.items()}
        if isinstance(tree, (list, tuple)):
            return type(tree)(tree_map(fn, v, *(r[i] for r in rest)) for i, v in enumerate(tree))
        return fn(tree, *rest)
    u.tree_flatten, u.tree_unflatten, u.tree_map = tree_flatten, tree_unflatten, tree_map
    return u


def install(precise=True):
    mlx, core, nn, utils = build(precise)
    sys.modules["mlx"] = mlx
    sys.modules["mlx.core"] = core
    sys.modules["mlx.core.fft"] = core.fft
    sys.modules["mlx.core.random"] = core.random
    sys.modules["mlx.core.fast"] = core.fast
    sys.modules["mlx.nn"] = nn
    sys.modules["mlx.utils"] = utils
    return core, nn


def stub_package(name, path):
    """Register an empty package object whose ``__path__`` points at a reference directory, so that its submodules import by
    path without executing the reference's (dependency-heavy) ``__init__.py`` files."""
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m
