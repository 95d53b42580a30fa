"""Golden vectors from the REFERENCE'S OWN Whisper model code (stt/models/whisper/whisper.py: AudioEncoder, TextDecoder with
its kv-cache protocol), executed in float64 with NumPy standing in for MLX (numpy_mlx_nn.py).  Run from the repo root in the
build container:  python tests/golden/make_whisper_golden.py   ->  tests/golden/whisper_golden.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy_mlx_nn as shim          # noqa: E402
import synth_params                  # noqa: E402

REF = "/root/reference/mlx_audio"
mx, nn = shim.install(precise=True)
for name, path in (("mlx_audio", REF), ("mlx_audio.stt", f"{REF}/stt"), ("mlx_audio.stt.models", f"{REF}/stt/models"),
                   ("mlx_audio.stt.models.whisper", f"{REF}/stt/models/whisper")):
    shim.stub_package(name, path)
import types                          # noqa: E402
for stub, names in (("mlx_audio.stt.utils", ("load_audio",)), ("huggingface_hub", ("snapshot_download",))):
    m = types.ModuleType(stub)
    for n in names:
        setattr(m, n, None)
    sys.modules[stub] = m
import mlx_audio.dsp as _dsp          # noqa: E402  (the reference's dsp.py, through the shim)
u = types.ModuleType("mlx_audio.utils")
u.hanning, u.mel_filters, u.stft = _dsp.hanning, _dsp.mel_filters, _dsp.stft
sys.modules["mlx_audio.utils"] = u
from mlx_audio.stt.models.whisper import whisper as W     # noqa: E402

DIMS = dict(n_mels=80, n_audio_ctx=60, n_audio_state=64, n_audio_head=4, n_audio_layer=2, n_vocab=300, n_text_ctx=32, n_text_state=64,
            n_text_head=4, n_text_layer=2)


def main():
    model = W.Model(W.ModelDimensions(**DIMS), dtype=mx.float32)
    names = [(n, v.shape) for n, v in shim.flat_parameters(model)]
    for n, sh in names:
        shim.set_parameter(model, n, synth_params.value(n, sh))
    rng = np.random.default_rng(21)
    mel = rng.standard_normal((2, 2 * DIMS["n_audio_ctx"], DIMS["n_mels"]))
    xa = model.encoder(mx.array(mel))
    tokens = rng.integers(0, DIMS["n_vocab"], size=(2, 7))
    logits, kv, cross_qk = model.decoder(mx.array(tokens), xa)
    step_tokens = rng.integers(0, DIMS["n_vocab"], size=(2, 3))
    step_logits = []
    for i in range(step_tokens.shape[1]):
        lg, kv, _ = model.decoder(mx.array(step_tokens[:, i:i + 1]), xa, kv_cache=kv)
        step_logits.append(np.asarray(lg))
    out = dict(params=synth_params.manifest(names), mel=mel, xa=np.asarray(xa), tokens=tokens, logits=np.asarray(logits),
               cross_qk_last=np.asarray(cross_qk[-1]), step_tokens=step_tokens, step_logits=np.stack(step_logits, 1)[:, :, 0],
               sinusoids=np.asarray(W.sinusoids(60, 64)))
    np.savez_compressed(os.path.join(HERE, "whisper_golden.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
