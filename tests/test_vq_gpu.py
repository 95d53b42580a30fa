"""Nearest-code search of the codec ENCODE side (csrc/codec.cu: rvq_encode_kernel) against float64 torch on identical inputs: the code
indices are integer work and must be bit-exact (ties -> lowest index)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mimi_style_residual_search_is_bit_exact():
    """quantization.py:37-45, 90-101: argmin(|e|^2 / 2 - x.e) level after level on the residual; 7 levels x 2048 codes x 256 dims, 333 rows
    (not a multiple of the 4-row CTA), plus a crafted exact tie."""
    from mlx_audio_b200 import ops
    g = torch.Generator().manual_seed(0)
    nq, bins, D, R = 7, 2048, 256, 333
    cb = torch.randn(nq, bins, D, generator=g) * 0.3
    cb[0, 100] = cb[0, 7]                                            # duplicate code: the lower index must win
    x = torch.randn(R, D, generator=g)
    x[5] = cb[0, 100] * 1.0
    c2 = (cb.double() ** 2).sum(-1) / 2
    r = x.double().clone()
    want = []
    for q in range(nq):
        idx = (c2[q][None] - r @ cb[q].double().T).argmin(dim=-1)
        r = r - cb[q].double()[idx]
        want.append(idx)
    want = torch.stack(want, dim=1)
    got = ops.rvq_encode(x.cuda(), cb.cuda().contiguous(), c2.cuda().contiguous())
    assert got.shape == (R, nq) and torch.equal(got.cpu(), want) and int(got[5, 0]) == 7
    # strided output: codes laid out [B, nq, T] as the codecs return them
    out = torch.zeros(nq, R, dtype=torch.int64, device="cuda")
    ops.rvq_encode(x.cuda(), cb.cuda().contiguous(), c2.cuda().contiguous(), out=out.t())
    assert torch.equal(out.t().cpu(), want)


def test_snac_style_cosine_search_is_bit_exact():
    """snac/vq.py:56-73: L2-normalise encodings and code table, dist = |a|^2 - 2 a.b + |b|^2, arg-max of -dist; 4096 codes x 8 dims."""
    from mlx_audio_b200 import ops
    g = torch.Generator().manual_seed(1)
    bins, D, R = 4096, 8, 1001
    cb = torch.randn(bins, D, generator=g)
    x = torch.randn(R, D, generator=g) * 3
    cn = cb.double() / cb.double().norm(dim=1, keepdim=True).clamp(min=1e-12)
    en = x.double() / x.double().norm(dim=1, keepdim=True).clamp(min=1e-12)
    dist = (en ** 2).sum(1, keepdim=True) - 2 * en @ cn.T + (cn ** 2).sum(1, keepdim=True).T
    want = (-dist).argmax(1)
    cn32 = cn.float()                                              # the kernel takes the normalised table in fp32 (weights are prepared once)
    dist32 = (en ** 2).sum(1, keepdim=True) - 2 * en @ cn32.double().T + (cn32.double() ** 2).sum(1, keepdim=True).T
    want32 = (-dist32).argmax(1)
    got = ops.rvq_encode(x.cuda(), cn32[None].cuda().contiguous(), (cn32.double() ** 2).sum(1)[None].cuda().contiguous(), mode=1)[:, 0].cpu()
    assert torch.equal(got, want32) and float((got == want).float().mean()) > 0.999
