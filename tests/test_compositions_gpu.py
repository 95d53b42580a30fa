"""Host-side compositions of validated kernels (default batch path of Qwen3-TTS, Mimi.decode_step, SNAC.decode_stream, Whisper
Model.logits) against the oracle.  First hardware run: round 1 driver (all four passed there as xpass; the marker is gone, so a
regression now fails)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import qwen3 as Q


def test_default_batch_path_rows_equal_single_sequence_generation_and_chunked_decode():
    """Model.batch_generate(stream=False) of the reference = Qwen3TTSBatchSession (continuous_batching.py; pinned by
    tests/test_oracle_pins.py::test_oracle_qwen3_default_batch_path_matches_the_reference_session): every row generates what the
    single-sequence loop generates from its own uniform stream (standard trailing-text rule), and is decoded in 15-frame chunks with 5
    frames of left context.  Product: generate_codes(trailing_rule="standard") + _decode_generated_codes; codes bit-exact, waveform 1e-3
    relative RMS; rows run past 15 frames so that the chunking matters."""
    from test_qwen3_gpu import _talker, _tokenizer          # tests/ is on sys.path (rootdir-relative "prepend" import mode)
    model, Pt, flat = _talker({"num_hidden_layers": 2, "cp_num_hidden_layers": 1})
    st, P64, tflat = _tokenizer()
    model.load_speech_tokenizer(st)
    tc = model.config.talker_config
    cfg_ids = {k: getattr(tc, k) for k in ("codec_nothink_id", "codec_think_id", "codec_think_bos_id", "codec_think_eos_id", "codec_pad_id", "codec_bos_id")}
    g = torch.Generator().manual_seed(31)
    ids_list = [torch.randint(0, 500, (n,), generator=g).tolist() for n in (12, 30, 17)]          # trailing texts of 3, 21 and 8 rows
    n_frames = 18
    u = torch.rand(n_frames, 16, 3, generator=g)
    x, trailing, pad, left = model.prepare_batch_inputs_from_ids(ids_list, language_id=2050)
    codes, lengths = model.generate_codes(x, trailing, pad, max_tokens=n_frames, u=u, left_padding=left, batch_mode=True, trailing_rule="standard")
    for b, ids in enumerate(ids_list):
        ie, tr, pd = Q.prepare_generation_inputs_from_ids(Pt, ids, (501, 502, 500), cfg_ids, language_id=2050)
        want = Q.generate_codes(Pt, ie, tr, pd, u[:, :, b].double(), n_frames, cfg=flat)
        assert int(lengths[b]) == want.shape[0] and torch.equal(codes[b, : int(lengths[b])].cpu(), want), b
        audio = model._decode_generated_codes(codes[b, : int(lengths[b])])
        ref = Q.decode_generated_codes(P64, want, tflat)
        assert audio.shape[0] == want.shape[0] * 1920
        err = float(((audio.cpu().double() - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())
        assert err < 1e-3, (b, err)
    res = list(model.batch_generate_from_ids(ids_list, language_id=2050, max_tokens=n_frames, u=u))
    assert [r.sequence_idx for r in res] == [0, 1, 2] and all(r.samples == r.token_count * 1920 and not r.is_streaming_chunk for r in res)


def test_mimi_decode_step_returns_the_slices_of_a_one_shot_decode():
    """Mimi.decode_step (mimi.py:171-176): in the reference the incremental path equals the one-shot decode (checked on the reference's own
    code, tests/golden/make_codec_golden.py); the product re-decodes the codes seen so far, so the chunks are the slices."""
    from mlx_audio_b200 import synth
    from mlx_audio_b200.codec import Mimi, mimi_202407
    from oracle import codec as OC
    m = Mimi(mimi_202407(32), device="cuda:0").load_weights(synth.mimi_weights(OC.MIMI_202407))
    codes = synth.mimi_codes(OC.MIMI_202407, 12, batch=2)
    full = m.decode(codes)
    m.reset_state()
    parts = [m.decode_step(codes[:, :, :5]), m.decode_step(codes[:, :, 5:6]), m.decode_step(codes[:, :, 6:])]
    assert [p.shape[-1] for p in parts] == [5 * 1920, 1920, 6 * 1920]
    got = torch.cat(parts, dim=-1)
    assert float((got - full).abs().max()) <= 1e-4 * max(1.0, float(full.abs().max()))      # shorter decodes: same maths, other tile shapes
    m.decode(codes[:, :, :3])                                            # decode() starts a new stream
    assert m.decode_step(codes[:, :, :2]).shape[-1] == 2 * 1920
    from mlx_audio.codec import MimiStreamingDecoder
    sd = MimiStreamingDecoder(m)
    blocks = torch.cat([sd.decode_frames(codes[:, :, :7]), sd.decode_frames(codes[:, :, 7:])], dim=-1)
    assert float((blocks - full).abs().max()) <= 1e-4 * max(1.0, float(full.abs().max()))


def test_snac_decode_stream_follows_the_reference_function():
    """SNAC.decode_stream (snac.py:106-162; the oracle's restatement is pinned to the reference's own run incl. the untrimmed-context
    quirk): product vs oracle on the 24 kHz configuration, two calls."""
    from mlx_audio_b200 import synth
    from mlx_audio_b200.codec import SNAC
    from oracle import codec as OC
    P = synth.snac_weights(OC.SNAC_24K)
    model = SNAC.from_config(OC.SNAC_24K, device="cuda:0").load_weights(P)
    P64 = {k: v.double() for k, v in P.items()}
    c1, c2 = synth.snac_codes(OC.SNAC_24K, 24, 1, seed=6), synth.snac_codes(OC.SNAC_24K, 16, 1, seed=8)
    n1, n2 = synth.snac_noises(OC.SNAC_24K, 1, seed=7), synth.snac_noises(OC.SNAC_24K, 1, seed=9)
    a1, ctx = model.decode_stream(c1, noises=n1)
    a2, ctx2 = model.decode_stream(c2, prev_codes=ctx, context_frames=8, noises=n2)
    r1, rctx = OC.snac_decode_stream(P64, c1, None, 8, OC.SNAC_24K, [n.double() for n in n1])
    r2, rctx2 = OC.snac_decode_stream(P64, c2, rctx, 8, OC.SNAC_24K, [n.double() for n in n2])
    assert a1.shape == r1.shape and a2.shape == r2.shape and a2.shape[1] > a1.shape[1] * 16 // 24      # context audio is returned again
    for a, r in ((a1, r1), (a2, r2)):
        assert float(((a.cpu().double() - r) ** 2).mean().sqrt() / (r ** 2).mean().sqrt()) < 1e-3
    assert all(torch.equal(c.cpu(), rc) for c, rc in zip(ctx2, rctx2))


def test_whisper_logits_of_every_position():
    """Model.logits(tokens, audio_features) (whisper.py:623-624): all positions, vs the float64 oracle (1e-3 relative to the logit scale)."""
    from mlx_audio_b200 import synth
    from mlx_audio_b200.stt.models.whisper import Model, ModelDimensions
    from oracle import whisper as OW
    dims = dict(OW.WHISPER_SMALL)
    dims["n_text_layer"] = 2
    P = synth.whisper_decoder_weights(dims)
    model = Model(ModelDimensions.from_dict(dims), device="cuda:0").load_weights(P)
    xa = torch.randn(2, 1500, 768, generator=torch.Generator().manual_seed(0))
    tokens = torch.randint(0, dims["n_vocab"], (2, 9), generator=torch.Generator().manual_seed(1))
    want, _ = OW.decoder_forward({k: v.double() for k, v in P.items()}, tokens, xa.double(), None, dims)
    got = model.logits(tokens, xa)
    assert got.shape == want.shape == (2, 9, dims["n_vocab"])
    assert float((got.double().cpu() - want).abs().max() / want.abs().max()) < 1e-3
    assert model.is_multilingual and model.num_languages == 99
