"""Golden trace from the REFERENCE'S OWN ``lm/models/cache.py:KVCache`` executed with NumPy standing in for MLX (numpy_mlx_nn.py): a
sequence of update_and_fetch / trim / state operations with the capacity, offset and fetched contents after each one.
python tests/golden/make_cache_golden.py -> tests/golden/cache_golden.npz"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy_mlx_nn as shim          # noqa: E402

REF = "/root/reference/mlx_audio"
mx, nn = shim.install(precise=True)
for name, path in (("mlx_audio", REF), ("mlx_audio.lm", f"{REF}/lm"), ("mlx_audio.lm.models", f"{REF}/lm/models")):
    shim.stub_package(name, path)
from mlx_audio.lm.models.cache import KVCache          # noqa: E402

OPS = [("update", 3), ("update", 300), ("update", 1), ("trim", 5), ("update", 2), ("update", 210), ("trim", 1000), ("update", 4), ("state", 0),
       ("update", 256), ("update", 1)]


def main():
    rng = np.random.default_rng(81)
    c = KVCache()
    out, trace = {}, []
    for i, (op, n) in enumerate(OPS):
        rec = {"op": op, "n": n}
        if op == "update":
            k, v = rng.standard_normal((1, 2, n, 4)), rng.standard_normal((1, 2, n, 3))
            fk, fv = c.update_and_fetch(mx.array(k), mx.array(v))
            out[f"k_{i}"], out[f"v_{i}"], out[f"fk_{i}"], out[f"fv_{i}"] = k, v, np.array(fk, copy=True), np.array(fv, copy=True)   # MLX slices are values, NumPy's are views
        elif op == "trim":
            rec["trimmed"] = int(c.trim(n))
        elif op == "state":
            sk, sv = c.state
            rec["state_len"] = int(sk.shape[2])
            c.state = (sk, sv)                                         # round trip through the setter
        rec.update(offset=int(c.offset), capacity=int(c.keys.shape[2]) if c.keys is not None else -1, size=int(c.size()), empty=bool(c.empty()),
                   trimmable=bool(c.is_trimmable()))
        trace.append(rec)
        print(rec)
    np.savez_compressed(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "cache_golden.npz"), trace=json.dumps(trace), **out)


if __name__ == "__main__":
    main()
