"""Kokoro-82M end-to-end parity: CUDA product vs the float64 oracle on the same synthetic bf16
checkpoint and inputs.  Tolerance (north_star): waveform RMS error
<= 1e-3 of the reference RMS; durations (integer work) bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mlx_audio_b200 import synth
from oracle import kokoro as OK

TOL = 1e-3


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
    P64 = {k: v.double() for k, v in P.items()}
    return model, P64, cfg


def _run(setup, n_ph, dur, seed):
    model, P64, cfg = setup
    ids, ref_s = synth.kokoro_inputs(n_ph, seed=seed)
    T = ids.shape[1]
    pd = None if dur is None else [dur] * T
    OK.TAP = {}
    if pd is None:
        ref_audio, ref_pd = OK.forward(P64, ids, ref_s.double(), noise=lambda n: synth.kokoro_noise(n, seed + 2)[1].double())
    else:
        _, nz = synth.kokoro_noise(T * dur * 600, seed + 2)
        ref_audio, ref_pd = OK.forward(P64, ids, ref_s.double(), noise=nz.double(), pred_dur_override=pd)
    tap_ref, OK.TAP = OK.TAP, None
    F = int(ref_pd.sum())
    _, nz = synth.kokoro_noise(F * 600, seed + 2)
    model.tap = {}
    audio, pred = model.forward_ids(ids[0], ref_s, 1.0, noise=nz.to("cuda:0").contiguous(), pred_dur=pd)
    tap, model.tap = model.tap, None
    return audio, pred, ref_audio, ref_pd, tap, tap_ref


def _report(tap, tap_ref):
    rows = []
    for k in ("bert", "dec_encode", "dec_out", "har", "gen_stage0", "gen_stage1", "xpost"):
        if k in tap and k in tap_ref:
            a = tap[k][0] if tap[k].dim() == 3 else tap[k]
            b = tap_ref[k][0] if tap_ref[k].dim() == 3 else tap_ref[k]
            rows.append(f"{k}: {rel_rms(a, b):.2e}")
    return "; ".join(rows)


def test_kokoro_small_pinned_durations(setup):
    audio, pred, ref_audio, ref_pd, tap, tap_ref = _run(setup, 16, 3, seed=1)
    assert audio.shape == ref_audio.shape == (18 * 3 * 600,)
    assert pred.cpu().tolist() == ref_pd.tolist()
    e = rel_rms(audio, ref_audio)
    assert e < TOL, f"waveform rel RMS {e:.3e}; stages: {_report(tap, tap_ref)}"


def test_kokoro_model_durations(setup):
    """The model's own duration head (round-half-even, clip) must agree bit-exactly, then the waveform."""
    audio, pred, ref_audio, ref_pd, tap, tap_ref = _run(setup, 9, None, seed=5)
    dsum = tap["dur"].double().cpu()
    assert pred.cpu().tolist() == ref_pd.tolist(), f"durations differ (pre-round sums {dsum.tolist()})"
    e = rel_rms(audio, ref_audio)
    assert e < TOL, f"waveform rel RMS {e:.3e}; stages: {_report(tap, tap_ref)}"


def test_kokoro_cfg2_shape_and_call_api(setup):
    """BASELINE config 2: 128 phonemes, durations pinned to 3 -> 234 000 samples (9.75 s); __call__ API."""
    model, _, cfg = setup
    ids, ref_s = synth.kokoro_inputs(128, seed=1)
    audio, pred = model.forward_ids(ids[0], ref_s, pred_dur=[3] * 130)
    assert audio.shape == (234000,) and bool(torch.isfinite(audio).all())
    model.vocab = {chr(97 + i): i + 1 for i in range(26)}
    out = model("hello world", ref_s, 1.0, return_output=True)
    assert out.audio.dim() == 2 and out.audio.shape[0] == 1 and out.pred_dur.shape[0] == 12   # 10 letters + BOS/EOS
    res = list(model.generate("ignored", phonemes=["hello", "world"], ref_s=ref_s))
    assert len(res) == 2 and res[0].sample_rate == 24000 and res[0].audio.dim() == 1
    with pytest.raises(AssertionError):
        model("a" * 600, ref_s)                                                                   # kokoro.py:122-125 context assert
