"""Codec decode parity (BASELINE config 5): SNAC-24k and Mimi through the CUDA path vs the float64 oracle.
Waveform tolerance 1e-3 relative RMS (north_star); output lengths are the reference's own pins
(codec/tests/test_snac.py:36 -> 120 907, codec/tests/test_mimi.py:18-21 -> 120 960)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from mlx_audio_b200 import synth
from oracle import codec as OC


def rel_rms(a, b):
    a, b = torch.as_tensor(a).double().cpu().reshape(-1), torch.as_tensor(b).double().cpu().reshape(-1)
    return float(torch.sqrt(((a - b) ** 2).mean()) / torch.sqrt((b ** 2).mean()))


@pytest.fixture(scope="module")
def snac():
    from mlx_audio_b200.codec import SNAC
    P = synth.snac_weights(OC.SNAC_24K)
    return SNAC.from_config(OC.SNAC_24K, device="cuda:0").load_weights(P), {k: v.double() for k, v in P.items()}


@pytest.fixture(scope="module")
def mimi():
    from mlx_audio_b200.codec import Mimi, mimi_202407
    P = synth.mimi_weights(OC.MIMI_202407)
    return Mimi(mimi_202407(32), device="cuda:0").load_weights(P), {k: v.double() for k, v in P.items()}


@pytest.mark.parametrize("t_fine,batch", [(64, 1), (100, 2)])
def test_snac_decode_parity(snac, t_fine, batch):
    model, P64 = snac
    codes = synth.snac_codes(OC.SNAC_24K, t_fine, batch)
    noises = synth.snac_noises(OC.SNAC_24K, batch)
    ref = OC.snac_decode(P64, codes, noises=[n.double() for n in noises])
    y = model.decode(codes, noises=noises)
    assert y.shape == ref.shape
    assert rel_rms(y, ref) < 1e-3


def test_snac_reference_length_pin_and_bad_codes(snac):
    model, _ = snac
    codes = synth.snac_codes(OC.SNAC_24K, 236)
    assert [c.shape[1] for c in codes] == [59, 118, 236]
    y = model.decode(codes, noises=synth.snac_noises(OC.SNAC_24K))
    assert y.shape == (1, 120907, 1) and bool(torch.isfinite(y).all()) and float(y.abs().max()) <= 1.0
    codes[1][0, 3] = 4096
    with pytest.raises(ValueError):
        model.decode(codes, noises=synth.snac_noises(OC.SNAC_24K))


@pytest.mark.parametrize("t,batch", [(20, 1), (140, 2)])
def test_mimi_decode_parity(mimi, t, batch):
    """T=140 -> 280 transformer positions: exercises the context-250 window."""
    model, P64 = mimi
    codes = synth.mimi_codes(OC.MIMI_202407, t, batch)
    ref = OC.mimi_decode(P64, codes)
    y = model.decode(codes)
    assert y.shape == ref.shape == (batch, 1, 1920 * t)
    assert rel_rms(y, ref) < 1e-3


def test_mimi_reference_length_pin_and_prefix_causality(mimi):
    """63 frames -> 120 960 samples; the stack is causal, so decoding a prefix gives a prefix (size-independent property)."""
    model, _ = mimi
    codes = synth.mimi_codes(OC.MIMI_202407, 63)
    y = model.decode(codes)
    assert y.shape == (1, 1, 120960)
    y2 = model.decode(codes[:, :, :40])
    assert torch.equal(y2, y[:, :, : 40 * 1920])        # default route (pre-split planes + TMA pipeline): no length-dependent split-K, exact
    from mlx_audio_b200 import ops
    ops.FUSED_DISPATCH[0] = True
    try:
        # the fused conv kernel splits K across CTAs when a layer has few output tiles, and how many ways depends on the length: the two
        # decodes add the same fp32 partial products in a different order (measured 4e-5 on a 5.9 full scale)
        assert torch.allclose(model.decode(codes[:, :, :40]), model.decode(codes)[:, :, : 40 * 1920], atol=3e-4, rtol=1e-4)
    finally:
        ops.FUSED_DISPATCH[0] = False


@pytest.mark.parametrize("parts", [2, 3])
def test_snac_span_decodes_stitch_to_the_one_shot_decode(snac, parts):
    """SURVEY.md section 8e (config 5): one stream sharded by contiguous frame spans with a 16-frame halo per side; the per-channel
    NoiseBlock draws are shared.  The stitched waveform equals the one-shot decode: bit for bit on the default route; routed through the
    fused kernel (length-dependent split-K factor) the same fp32 partial products are added in a different order, which four Snake stages
    amplify to 2-4e-4 of full scale (measured)."""
    from mlx_audio_b200 import ops
    from mlx_audio_b200.parallel import shard_span
    model, _ = snac
    T = 236
    codes = synth.snac_codes(OC.SNAC_24K, T)
    noises = synth.snac_noises(OC.SNAC_24K)

    def both():
        full = model.decode(codes, noises=noises)
        pieces = []
        for r in range(parts):
            _, _, cs, ce = shard_span(T, r, parts, multiple=max(model.vq_strides))
            if ce > cs:
                pieces.append(model.decode_span(codes, cs, ce, noises=noises))
        return torch.cat(pieces, dim=1), full

    got, full = both()
    assert got.shape == full.shape == (1, 120907, 1)
    assert torch.equal(got, full)                         # default route
    ops.FUSED_DISPATCH[0] = True
    try:
        got, full = both()
        assert float((got - full).abs().max()) <= 1.5e-3
    finally:
        ops.FUSED_DISPATCH[0] = False
    with pytest.raises(ValueError):
        model.decode_span(codes, 2, 40, noises=noises)


def test_mimi_span_decodes_stitch_to_the_one_shot_decode(mimi):
    """Mimi is causal: a span needs LEFT context only -- num_layers x context transformer positions plus the convolutions (span_halo).  With
    a too-short halo the result differs, which is what makes the exact halo worth stating."""
    model, _ = mimi
    T = 2400
    codes = synth.mimi_codes(OC.MIMI_202407, T)
    full = model.decode(codes)
    assert model.span_halo == 8 * 250 // 2 + 16
    a = model.decode_span(codes, 0, 1200)
    b = model.decode_span(codes, 1200, 2400)
    got = torch.cat([a, b], dim=-1)
    assert got.shape == full.shape and float((got - full).abs().max()) <= 1e-5 * max(1.0, float(full.abs().max()))
    short = model.decode_span(codes, 1200, 2400, halo=40)
    assert float((short - full[..., 1200 * 1920:]).abs().max()) > 1e-4 * float(full.abs().max())


def test_snac_encode_matches_the_oracle():
    """SURVEY.md section 8f row 2 (codec encode side), SNAC: the product's encoder + residual cosine quantiser against oracle.codec.snac_encode
    (itself pinned to the reference's SNAC.encode, tests/test_oracle_pins.py).  The latent in front of the quantiser to 2e-4; the code streams
    identical except where two codes are within the fp32 error of the latent of being equally near (a 4096-way arg-max over 8-dim cosines)."""
    from mlx_audio_b200.codec import SNAC
    P = synth.snac_weights(OC.SNAC_24K, encoder=True)
    model = SNAC.from_config(OC.SNAC_24K, device="cuda:0").load_weights(P)
    P64 = {k: v.double() for k, v in P.items()}
    g = torch.Generator().manual_seed(3)
    audio = torch.randn(2, 1, 9000, generator=g) * 0.3                  # 9000 -> right-padded to 10 240 = 20 finest frames
    z = model.encode_latent(audio)
    z_ref = OC.snac_encoder(P64, OC.snac_preprocess(audio.double(), OC.SNAC_24K).transpose(1, 2), OC.SNAC_24K)
    assert z.shape == z_ref.shape == (2, 20, 768)
    assert float((z.double().cpu() - z_ref).abs().max() / z_ref.abs().max()) < 2e-4
    codes = model.encode(audio)
    want = OC.snac_encode(P64, audio.double(), OC.SNAC_24K)
    assert [tuple(c.shape) for c in codes] == [tuple(w.shape) for w in want] == [(2, 5), (2, 10), (2, 20)]
    same = [float((c.cpu() == w).float().mean()) for c, w in zip(codes, want)]
    assert min(same) >= 0.9, same
    assert all(int(c.min()) >= 0 and int(c.max()) < 4096 and c.dtype == torch.int64 for c in codes)
    y = model.decode(codes)                                              # the code streams are valid decoder input
    assert y.shape[0] == 2 and bool(torch.isfinite(y).all())
    with pytest.raises(ValueError):
        SNAC.from_config(OC.SNAC_24K, device="cuda:0").load_weights(synth.snac_weights(OC.SNAC_24K)).encode(audio)


def test_mimi_encode_matches_the_oracle():
    """SURVEY.md section 8f row 2, Mimi: SEANet encoder + encoder transformer + replicate-padded stride-2 conv against the oracle's latent
    (2e-4), then the split residual quantiser against oracle.codec.mimi_encode (pinned to the reference's Mimi.encode).  32 codebooks deep
    the residual is small and fp32 differences of the latent move some arg-mins: the first codebooks must agree almost everywhere."""
    from mlx_audio_b200.codec import Mimi, mimi_202407
    cfg = OC.MIMI_202407
    P = synth.mimi_weights(cfg, encoder=True)
    model = Mimi(mimi_202407(32), device="cuda:0").load_weights(P)
    P64 = {k: v.double() for k, v in P.items()}
    g = torch.Generator().manual_seed(4)
    pcm = torch.randn(2, 1, 1920 * 6 + 700, generator=g) * 0.3          # 6 frames + 700 samples -> 7 frames
    z = model.encode_latent(pcm)
    x = OC.mimi_seanet_encoder(P64, pcm.double(), cfg)
    x = OC.mimi_transformer(P64, "encoder_transformer", x, cfg)
    z_ref = OC.mimi_causal_conv(P64, "downsample.conv", x, 4, stride=2, pad_mode="edge").transpose(1, 2)
    assert z.shape == z_ref.shape == (2, 7, 512)
    assert float((z.double().cpu() - z_ref).abs().max() / z_ref.abs().max()) < 2e-4
    codes = model.encode(pcm)
    want = OC.mimi_encode(P64, pcm.double(), cfg)
    assert codes.shape == want.shape == (2, 32, 7) and codes.dtype == torch.int64
    same = (codes.cpu() == want).float()
    assert float(same[:, :4].mean()) >= 0.9 and float(same.mean()) >= 0.6, (float(same[:, :4].mean()), float(same.mean()))
    y = model.decode(codes)
    assert y.shape == (2, 1, 7 * 1920) and bool(torch.isfinite(y).all())
