"""GPU polyphase resampler vs the reference's arithmetic (scipy.signal.resample_poly with the reference's Kaiser filter,
resample.py:10-47) and the reference's own pins (tests/test_dsp.py:299-378): alias rejection, pass-band gain, and
chunk-invariance (every output sample is an independent fixed-order sum)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dsp as OD


@pytest.mark.parametrize("orig,target,n", [(24000, 16000, 48137), (44100, 16000, 88337), (48000, 16000, 96137), (16000, 24000, 20011)])
def test_resample_matches_scipy(orig, target, n):
    from mlx_audio_b200.utils import resample_audio
    rng = np.random.default_rng(orig)
    x = rng.normal(0.0, 0.1, size=(2, n)).astype(np.float32)
    ref = OD.resample(x, orig, target, axis=-1)
    y = resample_audio(torch.as_tensor(x).cuda(), orig, target)
    assert isinstance(y, torch.Tensor) and y.is_cuda and y.shape == ref.shape
    assert float(np.abs(y.cpu().numpy() - ref).max()) <= 1.5e-7          # float32 rounding of a float64 sum
    # time-first layout through `axis`
    y2 = resample_audio(torch.as_tensor(x.T.copy()).cuda(), orig, target, axis=0)
    assert torch.equal(y2.T.contiguous(), y)


def test_resample_reference_pins_and_chunk_invariance():
    from mlx_audio_b200.utils import resample_audio
    orig, target = 24000, 16000
    t = np.arange(2 * orig) / orig
    out = resample_audio(torch.as_tensor(np.sin(2 * np.pi * 8200.0 * t).astype(np.float32)).cuda(), orig, target).cpu().numpy()
    assert float(np.sqrt(np.mean(out[400:-400] ** 2))) < 0.01
    for f in (1000.0, 7000.0):
        out = resample_audio(torch.as_tensor(np.sin(2 * np.pi * f * t).astype(np.float32)).cuda(), orig, target).cpu().numpy()
        assert 0.70 < float(np.sqrt(np.mean(out[400:-400] ** 2))) < 0.72
    x = torch.randn(48000, generator=torch.Generator().manual_seed(1)).cuda()
    whole = resample_audio(x, orig, target)
    # a chunk with enough halo reproduces the interior of the whole-buffer result bit-for-bit (tests/test_dsp.py:350-378 property)
    lo, hi, halo = 12000, 30000, 600
    part = resample_audio(x[lo - halo:hi + halo], orig, target)
    o_lo, o_hi = lo * 2 // 3, hi * 2 // 3
    off = (lo - halo) * 2 // 3
    assert torch.equal(part[o_lo - off:o_hi - off], whole[o_lo:o_hi])
    assert resample_audio(x, 16000, 16000) is x


def test_gpu_resampler_matches_reference_golden_vectors():
    """The CUDA polyphase kernel against vectors produced by the reference's own resample.py (tests/golden/)."""
    import os
    from mlx_audio_b200.resample import resample_audio_array, resample_audio_chunks
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resample_golden.npz"))
    for name, seed, shape, osr, tsr, axis in [("24k_16k", 0, (4801,), 24000, 16000, -1), ("16k_24k", 1, (3000,), 16000, 24000, -1),
                                              ("44k1_16k", 2, (8820,), 44100, 16000, -1), ("48k_16k_2d", 3, (2, 4800), 48000, 16000, -1),
                                              ("22k05_24k_axis0", 4, (2205, 2), 22050, 24000, 0), ("8k_16k_short", 5, (37,), 8000, 16000, -1)]:
        x = torch.as_tensor(np.random.default_rng(seed).standard_normal(shape).astype(np.float32)).cuda()
        y = resample_audio_array(x, osr, tsr, axis=axis)
        assert tuple(y.shape) == g[name].shape and float(np.abs(y.cpu().numpy() - g[name]).max()) <= 3e-7
        if name + "_chunks" in g.files:
            yc = resample_audio_chunks(iter(torch.tensor_split(x, 3, dim=0)), osr, tsr, x.shape[0])
            assert float(np.abs(yc.cpu().numpy() - g[name + "_chunks"]).max()) <= 3e-7
