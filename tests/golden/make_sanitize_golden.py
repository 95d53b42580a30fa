"""Manifests (key -> shape, CRC-32 of the float32 bytes) of what the REFERENCE'S OWN ``sanitize`` functions return for the synthetic
hub-layout checkpoints of checkpoint_layouts.py, executed with NumPy standing in for MLX.
python tests/golden/make_sanitize_golden.py  ->  tests/golden/sanitize_golden.json"""
import json
import os
import sys
import types
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy_mlx_nn as shim          # noqa: E402
import checkpoint_layouts as L      # noqa: E402

ENC_HF = L.qwen3_tokenizer_encoder_hf()        # needs the real transformers / huggingface_hub: taken before the stand-ins are installed
REF = "/root/reference/mlx_audio"
mx, nn = shim.install(precise=True)
for name, path in (("mlx_audio", REF), ("mlx_audio.lm", f"{REF}/lm"), ("mlx_audio.lm.models", f"{REF}/lm/models"), ("mlx_audio.tts", f"{REF}/tts"),
                   ("mlx_audio.tts.models", f"{REF}/tts/models"), ("mlx_audio.tts.models.qwen3_tts", f"{REF}/tts/models/qwen3_tts"),
                   ("mlx_audio.tts.models.kokoro", f"{REF}/tts/models/kokoro"), ("mlx_audio.codec", f"{REF}/codec"),
                   ("mlx_audio.codec.models", f"{REF}/codec/models"), ("mlx_audio.codec.models.mimi", f"{REF}/codec/models/mimi"),
                   ("mlx_audio.stt", f"{REF}/stt"), ("mlx_audio.stt.models", f"{REF}/stt/models"), ("mlx_audio.stt.models.whisper", f"{REF}/stt/models/whisper")):
    shim.stub_package(name, path)
for stub, names in (("huggingface_hub", ("snapshot_download", "hf_hub_download")), ("mlx_audio.stt.utils", ("load_audio",))):
    m = types.ModuleType(stub)
    for n in names:
        setattr(m, n, None)
    sys.modules[stub] = m
import mlx_audio.dsp as _dsp          # noqa: E402
u = types.ModuleType("mlx_audio.utils")
u.load_audio = None
for n in ("hanning", "mel_filters", "stft", "istft"):
    setattr(u, n, getattr(_dsp, n))
sys.modules["mlx_audio.utils"] = u


def manifest(d):
    return {k: [list(np.asarray(v).shape), zlib.crc32(np.ascontiguousarray(np.asarray(v), dtype=np.float32).tobytes())] for k, v in d.items()}


def main():
    out = {}
    from mlx_audio.stt.models.whisper import whisper as W
    model = W.Model(W.ModelDimensions(**L.WHISPER_DIMS), dtype=mx.float32)
    s = model.sanitize({k: mx.array(v) for k, v in L.whisper_hf().items()})
    tree = {n for n, _ in shim.flat_parameters(model)}
    assert set(s) == tree, (sorted(set(s) - tree)[:5], sorted(tree - set(s))[:5])     # the synthetic HF names cover the whole module tree
    out["whisper_hf"] = manifest(s)
    from mlx_audio.tts.models.qwen3_tts import qwen3_tts as Q
    from mlx_audio.tts.models.qwen3_tts import speech_tokenizer as S
    out["qwen3_model"] = manifest(Q.Model.sanitize({k: mx.array(v) for k, v in L.qwen3_model_torch().items()}))
    full = S.Qwen3TTSSpeechTokenizer.sanitize({k: mx.array(v) for k, v in L.qwen3_tokenizer_torch().items()})
    out["qwen3_tokenizer_decoder"] = manifest({k: v for k, v in full.items() if not k.startswith("encoder_model.")})
    enc = S.Qwen3TTSSpeechTokenizer.sanitize({k: mx.array(v) for k, v in ENC_HF.items()})
    out["qwen3_tokenizer_encoder"] = manifest(enc)
    from mlx_audio.tts.models.kokoro import kokoro as K
    sys.path.insert(0, ROOT)
    from oracle.kokoro import KOKORO_CONFIG                          # constants only: the public configuration
    km = K.Model(K.ModelConfig(**json.loads(json.dumps(KOKORO_CONFIG)), vocab={}))
    out["kokoro_torch"] = manifest(km.sanitize({k: mx.array(v) for k, v in L.kokoro_torch().items()}))
    from mlx_audio.codec.models.mimi import mimi as M
    captured = {}

    class _Loaded:
        def filter_and_map(self, fn):
            return None

    class _Mimi(M.Mimi):                                             # only load_pytorch_weights is exercised: no module tree needed
        def __init__(self):
            pass

        def load_weights(self, weights, strict=True):
            captured.update(dict(weights))
            return _Loaded()
    mx.load = lambda f: {k: mx.array(v) for k, v in L.mimi_torch().items()}
    _Mimi().load_pytorch_weights("synthetic")
    out["mimi_torch"] = manifest(captured)
    for k, v in out.items():
        print(k, len(v))
    json.dump(out, open(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "sanitize_golden.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
