"""Kokoro-82M parity: CUDA product vs the float64 oracle on the same synthetic bf16 checkpoint and inputs
(SURVEY.md section 8d cfg2).

Tolerances.  Durations (integer work): bit-exact.  Text/prosody side (ALBERT, 6 BiLSTMs, F0/N heads): 1e-4
relative RMS.  Waveform: 1e-3 relative RMS (north_star) ON IDENTICAL F0/N CURVES -- the hn-NSF source integrates
F0 over the whole utterance and multiplies the phase by 300 (istftnet.py:585-597), so a 1e-6 relative change of
F0 moves harmonic 9 by ~0.1 rad after 100 frames; no float32 implementation (the reference's included) can track a
float64 oracle through that, so decoder parity is asserted with the oracle's float32-rounded F0/N injected on both
sides, and the free-running waveform is only sanity-bounded (DESIGN.md "conditioning of the harmonic source")."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from mlx_audio_b200 import synth
from oracle import kokoro as OK

TOL_WAVE, TOL_TEXT = 1e-3, 1e-4


def rel_rms(a, b):
    a = torch.as_tensor(a).double().cpu().reshape(-1)
    b = torch.as_tensor(b).double().cpu().reshape(-1)
    return float(torch.sqrt(((a - b) ** 2).mean()) / (torch.sqrt((b ** 2).mean()) + 1e-30))


@pytest.fixture(scope="module")
def setup():
    from mlx_audio_b200.tts.models.kokoro import Model, ModelConfig
    cfg = OK.KOKORO_CONFIG
    P = synth.kokoro_weights(cfg, seed=0)
    model = Model(ModelConfig.from_dict(cfg), device="cuda:0")
    model.load_weights(list(P.items()))
    return model, {k: v.double() for k, v in P.items()}, cfg


def _oracle(P64, ids, ref_s, nz, pd, f0n=None):
    OK.TAP = {}
    audio, pred = OK.forward(P64, ids, ref_s.double(), noise=nz, pred_dur_override=pd, f0n_override=f0n)
    tap, OK.TAP = OK.TAP, None
    return audio, pred, tap


def _product(model, ids, ref_s, nz, pd, f0n=None):
    model.tap = {}
    audio, pred = model.forward_ids(ids[0], ref_s, 1.0, noise=nz, pred_dur=pd, f0n_override=f0n)
    tap, model.tap = model.tap, None
    return audio, pred, tap


def _stages(tap, tap_ref, keys):
    out = {}
    for k in keys:
        if k in tap and k in tap_ref:
            a, b = tap[k], tap_ref[k]
            if k == "d":
                pass
            out[k] = rel_rms(a.reshape(-1), b.reshape(-1)) if a.numel() == b.numel() else float("nan")
    return out


def _case(setup, n_ph, dur, seed):
    model, P64, cfg = setup
    ids, ref_s = synth.kokoro_inputs(n_ph, seed=seed)
    T = ids.shape[1]
    if dur is None:
        noise_fn = lambda n: synth.kokoro_noise(n, seed + 2)[1].double()
        ref_audio, ref_pd, tap_ref = _oracle(P64, ids, ref_s, noise_fn, None)
    else:
        ref_audio, ref_pd, tap_ref = _oracle(P64, ids, ref_s, synth.kokoro_noise(T * dur * 600, seed + 2)[1].double(), [dur] * T)
    F = int(ref_pd.sum())
    nz = synth.kokoro_noise(F * 600, seed + 2)[1]
    nz_d = nz.to("cuda:0").contiguous()
    pd = None if dur is None else [dur] * T
    audio, pred, tap = _product(model, ids, ref_s, nz_d, pd)
    assert pred.cpu().tolist() == ref_pd.tolist(), f"durations differ; pre-round sums {tap['dur'].cpu().tolist()}"
    text = _stages(tap, tap_ref, ["bert", "en", "F0", "N", "t_en", "dec_encode", "dec_out"])
    assert all(v < TOL_TEXT for v in text.values()), f"text/prosody stages: {text}"
    free = rel_rms(audio, ref_audio)
    # decoder + generator on identical (float32-rounded oracle) curves
    f0n = (tap_ref["F0"].float().reshape(-1), tap_ref["N"].float().reshape(-1))
    ref2, _, tap_ref2 = _oracle(P64, ids, ref_s, nz.double(), ref_pd.tolist(), f0n)
    audio2, _, tap2 = _product(model, ids, ref_s, nz_d, ref_pd.tolist(), f0n)
    gen = _stages(tap2, tap_ref2, ["har", "gen_stage0", "gen_stage1", "xpost"])
    e = rel_rms(audio2, ref2)
    print(f"\n[kokoro parity n_ph={n_ph}] text {text} | generator {gen} | waveform {e:.2e} | free-running waveform {free:.2e}")
    assert e < TOL_WAVE, f"waveform rel RMS {e:.3e}; generator stages {gen}"
    assert free < 0.5, f"free-running waveform diverged: {free:.3e}"
    return audio


def test_kokoro_small_pinned_durations(setup):
    audio = _case(setup, 16, 3, seed=1)
    assert audio.shape == (18 * 3 * 600,)


def test_kokoro_model_durations(setup):
    """The model's own duration head (round-half-even, clip) must agree bit-exactly, then the waveform."""
    _case(setup, 9, None, seed=5)


def test_kokoro_cfg2_shape_and_call_api(setup):
    """BASELINE config 2: 128 phonemes, durations pinned to 3 -> 234 000 samples (9.75 s); __call__ / generate API."""
    model, _, cfg = setup
    ids, ref_s = synth.kokoro_inputs(128, seed=1)
    audio, pred = model.forward_ids(ids[0], ref_s)
    assert audio.shape == (234000,) and bool(torch.isfinite(audio).all()) and pred.cpu().tolist() == [3] * 130
    model.vocab = {chr(97 + i): i + 1 for i in range(26)}
    out = model("hello world", ref_s, 1.0, return_output=True)
    assert out.audio.dim() == 2 and out.audio.shape[0] == 1 and out.pred_dur.shape[0] == 12   # 10 letters + BOS/EOS
    res = list(model.generate("ignored", phonemes=["hello", "world"], ref_s=ref_s))
    assert len(res) == 2 and res[0].sample_rate == 24000 and res[0].audio.dim() == 1
    with pytest.raises(AssertionError):
        model("a" * 600, ref_s)                                                                   # kokoro.py:122-125 context assert


def test_kokoro_graph_path_matches_eager_and_draws_fresh_noise(setup):
    """Model.synthesize_ids (what __call__ / generate / bench.py use): the two replayed graphs reproduce the eager launch sequence
    bit-for-bit on injected noise; without injection every call draws fresh SineGen noise from the device-resident Philox state
    (mx.random.normal semantics, istftnet.py:649) and `seed()` makes a run repeatable."""
    from mlx_audio_b200 import ops
    model, _, _ = setup
    ids, ref_s = synth.kokoro_inputs(20, seed=3)
    T, F = 22, 44
    dur = [2] * T
    noise = ops.randn_(torch.empty(1, F * 600, 9, device="cuda:0"), 99, 0)
    a, pa = model.synthesize_ids(ids[0], ref_s, pred_dur=dur, noise=noise)
    a = a.clone()
    b, pb = model.forward_ids(ids[0], ref_s, noise=noise, pred_dur=dur)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(pa, pb)
    assert sum(v["launches"] for v in model._graphs.values()) > 100
    assert abs(float(noise.mean())) < 0.01 and abs(float(noise.std()) - 1.0) < 0.01               # Philox N(0,1) sanity
    # the model's own duration head through the graph path: one host read of F, same result as eager
    c, pc = model.synthesize_ids(ids[0], ref_s, noise=None)
    Fm = int(pc.sum())
    nz = ops.randn_(torch.empty(1, Fm * 600, 9, device="cuda:0"), 5, 0)
    c = model.synthesize_ids(ids[0], ref_s, noise=nz)[0].clone()
    d = model.forward_ids(ids[0], ref_s, noise=nz)[0]
    assert c.shape == (Fm * 600,) and torch.equal(c, d)
    # fresh noise per call; repeatable under a seed
    model.seed(11)
    x1 = model.synthesize_ids(ids[0], ref_s)[0].clone()
    x2 = model.synthesize_ids(ids[0], ref_s)[0].clone()
    model.seed(11)
    y1 = model.synthesize_ids(ids[0], ref_s)[0].clone()
    y2 = model.synthesize_ids(ids[0], ref_s)[0].clone()
    assert not torch.equal(x1, x2) and torch.equal(x1, y1) and torch.equal(x2, y2)
    # __call__ returns a private copy (the graph's static buffer is overwritten by the next call)
    model.vocab = {chr(97 + i): i + 1 for i in range(26)}
    o1 = model("hello", ref_s)
    keep = o1.clone()
    model("world", ref_s)
    assert torch.equal(o1, keep)


def _bench_golden():
    import os
    import numpy as np
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "bench_shapes_golden.npz"))


@pytest.mark.parametrize("path", ["eager", "graph"])
def test_kokoro_cfg2_benched_shape_matches_cached_oracle_waveform(setup, path):
    """The shape bench.py times -- 128 phonemes, T = 130, model duration head -> F = 390, 234 000 samples -- against the float64 oracle's
    waveform for exactly this input (tests/golden/bench_shapes_golden.npz, made by make_bench_shape_golden.py), eager and through the
    replayed graphs, 1e-3 relative RMS on identical F0/N curves."""
    model, _, _ = setup
    g = _bench_golden()
    ids, ref_s = synth.kokoro_inputs(128, seed=1)
    nz = synth.kokoro_noise(130 * 3 * 600, 3)[1].to("cuda:0").contiguous()
    f0n = (torch.as_tensor(g["kokoro_f0"]), torch.as_tensor(g["kokoro_n"]))
    run = model.forward_ids if path == "eager" else model.synthesize_ids
    audio, pred = run(ids[0], ref_s, noise=nz, f0n_override=f0n)                 # durations from the model's own head
    assert pred.cpu().tolist() == [3] * 130 and audio.shape == (234000,)
    e = rel_rms(audio, g["kokoro_audio"])
    print(f"\n[kokoro cfg2 {path}] waveform rel RMS vs cached oracle {e:.2e}")
    assert e < TOL_WAVE, e


def test_kokoro_free_running_error_is_the_conditioning_of_the_harmonic_source(setup):
    """Free-running (no F0/N injection) the product's waveform differs from the float64 oracle's by far more than 1e-3 -- because the
    hn-NSF phase integrates F0 x 300 over the utterance, not because a kernel is off.  Evidence: perturb the ORACLE's own F0 curve by a
    random relative error of the size the product's F0 actually has (measured here, ~1e-6..1e-5) and the oracle's waveform moves by
    the same order as the product's free-running deviation."""
    model, P64, cfg = setup
    ids, ref_s = synth.kokoro_inputs(16, seed=1)
    T = ids.shape[1]
    nz = synth.kokoro_noise(T * 3 * 600, 3)[1]
    ref_audio, ref_pd, tap_ref = _oracle(P64, ids, ref_s, nz.double(), [3] * T)
    audio, _, tap = _product(model, ids, ref_s, nz.to("cuda:0").contiguous(), [3] * T)
    free = rel_rms(audio, ref_audio)
    f0_ref = tap_ref["F0"].double().reshape(-1)
    f0_err = (tap["F0"].double().cpu().reshape(-1) - f0_ref)
    delta = float(torch.sqrt((f0_err ** 2).mean()) / torch.sqrt((f0_ref ** 2).mean()))
    gperturb = torch.Generator().manual_seed(0)
    devs = []
    for _ in range(3):
        f0p = f0_ref + torch.randn(f0_ref.shape, generator=gperturb, dtype=torch.float64) * float(torch.sqrt((f0_err ** 2).mean()))
        pert, _, _ = _oracle(P64, ids, ref_s, nz.double(), [3] * T, (f0p, tap_ref["N"].double().reshape(-1)))
        devs.append(rel_rms(pert, ref_audio))
    print(f"\n[kokoro conditioning] product F0 rel err {delta:.2e}; free-running waveform dev {free:.2e}; oracle under an F0 perturbation "
          f"of that size moves by {[f'{d:.2e}' for d in devs]}")
    assert delta < TOL_TEXT
    assert free < 10 * max(devs) + 1e-3, (free, devs)


def test_kokoro_generate_text_path_with_a_g2p_callable_and_a_voice_pack(setup):
    """Model.generate(text, voice=...) (kokoro.py:293-370 over pipeline.py): G2P tokens -> <= 510-phoneme chunks -> one graph-replayed call
    per chunk with the style row voice[len(phonemes) - 1].  misaki is not in this image, so the G2P is a callable; the voice is a pack tensor."""
    model, _, _ = setup
    model.vocab = {chr(97 + i): i + 1 for i in range(26)}
    model.vocab[" "] = 27

    class Tok:
        def __init__(self, text, phonemes, whitespace):
            self.text, self.phonemes, self.whitespace = text, phonemes, whitespace

    def g2p(text):
        toks = [Tok(w, w, " ") for w in text.split()]
        return "", toks
    pack = torch.randn(510, 1, 256, generator=torch.Generator().manual_seed(4))
    model._pipelines = {}
    res = list(model.generate("hello brave new world\nsecond line here", voice=pack, g2p=g2p))
    assert len(res) == 2 and [r.token_count for r in res] == [len("hello brave new world"), len("second line here")]
    ps = "hello brave new world"
    want = model(ps, pack[len(ps) - 1], 1.0, noise=None)
    assert res[0].samples == want.shape[1] and res[0].sample_rate == 24000 and res[0].audio.dim() == 1
    with pytest.raises(ValueError, match="too long"):
        list(model.generate("x", phonemes="a" * 600, ref_s=pack[0]))
