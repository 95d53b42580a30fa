#!/usr/bin/env python
"""Generate tests/golden/resample_golden.npz by running the REFERENCE's own resampler (mlx_audio/resample.py, NumPy/SciPy only,
so it imports without MLX) in the build container.  /root/reference does not exist on the GPU box: the vectors are committed.

    python tests/golden/make_resample_golden.py          # needs /root/reference

Inputs are regenerated from the seeds by the tests (np.random.default_rng(seed).standard_normal), only outputs are stored."""
import importlib.util
import os

import numpy as np

REF = "/root/reference/mlx_audio/resample.py"
CASES = [   # name, seed, shape, orig_sr, target_sr, axis
    ("24k_16k", 0, (4801,), 24000, 16000, -1),
    ("16k_24k", 1, (3000,), 16000, 24000, -1),
    ("44k1_16k", 2, (8820,), 44100, 16000, -1),
    ("48k_16k_2d", 3, (2, 4800), 48000, 16000, -1),
    ("22k05_24k_axis0", 4, (2205, 2), 22050, 24000, 0),
    ("8k_16k_short", 5, (37,), 8000, 16000, -1),
]


def main():
    spec = importlib.util.spec_from_file_location("ref_resample", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, seed, shape, osr, tsr, axis in CASES:
        x = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
        y = ref.resample_audio_array(x, osr, tsr, axis=axis)
        out[name] = np.asarray(y)
        if axis == 0 or x.ndim == 1:                                      # resample_audio_chunks is time-first (resample.py:50-161)
            chunks = np.array_split(x, 3, axis=0)
            out[name + "_chunks"] = np.asarray(ref.resample_audio_chunks(iter(chunks), osr, tsr, x.shape[0], chunk_duration_seconds=0.05))
    np.savez_compressed(os.path.join(os.environ.get("GOLDEN_OUT", os.path.dirname(os.path.abspath(__file__))), "resample_golden.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
