"""What the REFERENCE'S OWN config / result dataclasses parse and declare (imported with NumPy standing in for MLX): Whisper
ModelDimensions.from_dict on MLX- and HuggingFace-format configs, Qwen3-TTS ModelConfig.from_dict on a hub-style config.json, Kokoro
ModelConfig, and the field lists (name, default) of the result dataclasses of SURVEY.md row a22.
python tests/golden/make_config_golden.py -> tests/golden/config_golden.json"""
import dataclasses
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy_mlx_nn as shim          # noqa: E402
import config_cases as C            # noqa: E402

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


def fields(cls):
    out = []
    for f in dataclasses.fields(cls):
        d = None if f.default is dataclasses.MISSING else f.default
        out.append([f.name, f.default is not dataclasses.MISSING or f.default_factory is not dataclasses.MISSING, d if isinstance(d, (int, float, str, bool, type(None))) else repr(d)])
    return out


def main():
    out = {}
    from mlx_audio.stt.models.whisper import whisper as W
    out["whisper_dims"] = {k: dataclasses.asdict(W.ModelDimensions.from_dict(v)) for k, v in C.WHISPER_CONFIGS.items()}
    from mlx_audio.tts.models.qwen3_tts import config as QC
    out["qwen3_config"] = {k: dataclasses.asdict(QC.ModelConfig.from_dict(v)) for k, v in C.QWEN3_CONFIGS.items()}
    from mlx_audio.tts.models.kokoro import kokoro as K
    out["kokoro_config"] = dataclasses.asdict(K.ModelConfig.from_dict(C.KOKORO_CONFIG_JSON))
    from mlx_audio.tts.models import base as TB
    from mlx_audio.stt.models.whisper import decoding as D
    out["result_fields"] = {"GenerationResult": fields(TB.GenerationResult), "BatchGenerationResult": fields(TB.BatchGenerationResult),
                            "DecodingResult": fields(D.DecodingResult), "STTOutput": fields(W.STTOutput), "DecodingOptions": fields(D.DecodingOptions)}
    json.dump(out, open(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "config_golden.json"), "w"), indent=1, sort_keys=True)
    print({k: (list(v) if isinstance(v, dict) else type(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
