"""CPU-side tests: the C-ABI library loads and exports every symbol include/b200audio.h declares, host
logic (weight packing, sanitize, synthetic checkpoints, sharding) and a world_size-2 gloo run of the
multi-GPU plumbing.  No compute calls: there is no GPU here and no CPU fallback by design."""
import os
import re
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mlx_audio_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "b200audio.h")).read()
    declared = set(re.findall(r"\b(b2a_\w+)\s*\(", header))
    declared -= {"b2a_conv1d_t", "b2a_attn_t", "b2a_convf_t"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.PROTOTYPES), f"header vs ctypes prototypes differ: {declared ^ set(_lib.PROTOTYPES)}"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2a_version() == 100


def test_ctypes_struct_matches_c_layout(tmp_path):
    """sizeof/offsetof of the parameter structs as the C compiler sees them == the ctypes mirrors."""
    from mlx_audio_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text(textwrap.dedent('''
        #include <stdio.h>
        #include <stddef.h>
        #include "b200audio.h"
        int main(void) {
          printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b2a_conv1d_t), offsetof(b2a_conv1d_t, pre_scale), offsetof(b2a_conv1d_t, res),
                 offsetof(b2a_conv1d_t, accumulate), sizeof(b2a_attn_t), offsetof(b2a_attn_t, k_len),
                 sizeof(b2a_convf_t), offsetof(b2a_convf_t, shifts), offsetof(b2a_convf_t, res), offsetof(b2a_convf_t, stats_out));
          return 0; }'''))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    C, A, Fz = _lib.Conv1dParams, _lib.AttnParams, _lib.ConvFParams
    import ctypes
    assert got == [ctypes.sizeof(C), C.pre_scale.offset, C.res.offset, C.accumulate.offset, ctypes.sizeof(A), A.k_len.offset,
                   ctypes.sizeof(Fz), Fz.shifts.offset, Fz.res.offset, Fz.stats_out.offset]


def test_missing_library_fails_loudly(monkeypatch):
    from mlx_audio_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200audio.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_pack_conv_layouts():
    from mlx_audio_b200 import ops
    w = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)          # [Cout, K, Cin]
    cw = ops.pack_conv(w, torch.zeros(2), 1, "cpu")
    assert cw.w.shape == (3, 4, 2) and cw.K == 3 and cw.cin == 4 and cw.cout == 2
    assert float(cw.w[1, 2, 1]) == float(w[1, 1, 2])
    dw = ops.pack_conv(torch.arange(15.0).reshape(5, 3, 1), None, 5, "cpu")
    assert dw.w.shape == (3, 5) and float(dw.w[2, 4]) == 14.0
    with pytest.raises(NotImplementedError):
        ops.pack_conv(torch.zeros(4, 3, 2), None, 2, "cpu")
    lin = ops.pack_linear(torch.zeros(7, 5), None, "cpu")
    assert lin.K == 1 and lin.cin == 5 and lin.cout == 7


def test_kokoro_synthetic_checkpoint_and_sanitize():
    from mlx_audio_b200 import synth
    from mlx_audio_b200.tts.models.kokoro import Model, ModelConfig
    from oracle.kokoro import KOKORO_CONFIG
    P = synth.kokoro_weights(KOKORO_CONFIG)
    n = sum(v.numel() for v in P.values())
    assert 81.0e6 < n < 82.5e6                                                    # Kokoro-82M
    assert all(torch.equal(v, v.to(torch.bfloat16).float()) for v in P.values())  # a bf16 checkpoint
    cfg = ModelConfig.from_dict({**KOKORO_CONFIG, "unknown_key": 1})              # from_dict filters unknown keys (base.py:10-18)
    m = Model(cfg, device="cpu")
    # torch-layout checkpoint keys -> reference tree (kokoro.py:179-276)
    torch_ckpt = {"predictor.lstm.weight_ih_l0_reverse": torch.zeros(1024, 640), "text_encoder.cnn.0.1.gamma": torch.ones(512),
                  "decoder.generator.noise_convs.0.weight": torch.zeros(256, 22, 12), "bert.embeddings.position_ids": torch.zeros(1, 512),
                  "decoder.encode.conv1.weight_v": torch.zeros(1024, 514, 3), "predictor.F0_proj.weight": torch.zeros(1, 256, 1)}
    s = m.sanitize(torch_ckpt)
    assert "predictor.lstm.Wx_backward" in s and "text_encoder.cnn.0.1.weight" in s and "bert.embeddings.position_ids" not in s
    assert s["decoder.generator.noise_convs.0.weight"].shape == (256, 12, 22)
    assert s["decoder.encode.conv1.weight_v"].shape == (1024, 3, 514) and s["predictor.F0_proj.weight"].shape == (1, 1, 256)


def test_fold_weight_norm_matches_oracle_rule():
    from mlx_audio_b200.tts.models.kokoro.kokoro import fold_weight_norm
    from oracle.kokoro import weight_norm
    g = torch.Generator().manual_seed(0)
    v = (torch.randn(8, 3, 5, generator=g) * 0.02).to(torch.bfloat16).float()
    gg = torch.rand(8, 1, 1, generator=g).to(torch.bfloat16).float()
    assert torch.equal(fold_weight_norm(v, gg), weight_norm(v, gg))


def test_codec_oracle_reference_length_pins():
    """codec/tests/test_snac.py:30-36 (59/118/236 codes -> 120 907 samples) and test_mimi.py:18-21 (63 -> 120 960)."""
    from mlx_audio_b200 import synth
    from oracle import codec as OC
    P = synth.snac_weights(OC.SNAC_24K)
    y = OC.snac_decode(P, synth.snac_codes(OC.SNAC_24K, 236), noises=synth.snac_noises(OC.SNAC_24K))
    assert y.shape == (1, 120907, 1)
    y = OC.mimi_decode(synth.mimi_weights(OC.MIMI_202407), synth.mimi_codes(OC.MIMI_202407, 63))
    assert y.shape == (1, 1, 120960)


def test_shard_units_and_spans():
    from mlx_audio_b200.parallel import shard_span, shard_units
    lengths = [5, 50, 7, 30, 30, 1, 9, 12]
    parts = [shard_units(lengths, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(lengths)
    spans = [shard_span(10000, r, 8, halo_left=25) for r in range(8)]
    assert spans[0][:2] == (0, 1250) and spans[3] == (3750 - 25, 5000, 3750, 5000)
    assert [s[2] for s in spans] == [i * 1250 for i in range(8)] and spans[-1][3] == 10000
    s = shard_span(10000, 1, 3, halo_left=10, halo_right=10, multiple=4)
    assert s[2] % 4 == 0 and s[0] % 4 == 0 and s[0] <= s[2] - 10


GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from mlx_audio_b200.parallel import decode_stream_sharded, gather_waveforms, shard_units, world
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, ws = world()
lengths = [300, 100, 250, 50, 120]
mine = shard_units(lengths, rank, ws)
waves = [torch.full((lengths[i],), float(i)) for i in mine]          # stand-in for per-utterance synthesis
out = gather_waveforms(waves, mine, len(lengths), dst=0)
if rank == 0:
    assert [int(o.numel()) for o in out] == lengths and all(float(o[0]) == i for i, o in enumerate(out))
    print("GATHER_OK", mine)
else:
    assert out is None


class FakeSnac:                      # stands in for the GPU codec: 3 code levels, hop 4, sample value = global sample index
    vq_strides, device = [4, 2, 1], torch.device("cpu")
    def decode_span(self, codes, start, end, noises=None):
        assert start % 4 == 0
        n = (end - start) * 4 + (3 if end == codes[-1].shape[1] else 0)
        return (torch.arange(n, dtype=torch.float32) + start * 4).reshape(1, n, 1)


class FakeMimi:
    device = torch.device("cpu")
    def decode_span(self, codes, start, end):
        return (torch.arange((end - start) * 5, dtype=torch.float32) + start * 5).reshape(1, 1, -1)


T = 44
y = decode_stream_sharded(FakeSnac(), [torch.zeros(1, T // 4), torch.zeros(1, T // 2), torch.zeros(1, T)])
z = decode_stream_sharded(FakeMimi(), torch.zeros(1, 8, 31, dtype=torch.int64))
if rank == 0:
    assert torch.equal(y, torch.arange(T * 4 + 3, dtype=torch.float32)) and torch.equal(z, torch.arange(31 * 5, dtype=torch.float32))
    print("STREAM_OK")
else:
    assert y is None and z is None
dist.barrier(); dist.destroy_process_group()
'''


def test_gloo_world2_sharding_and_trailing_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0] and "STREAM_OK" in outs[0]


def test_drop_in_import_surface_and_loader(tmp_path):
    """The reference's import paths resolve and load_model follows utils.py:321-416 on a local dir."""
    import json
    from safetensors.torch import save_file
    import mlx_audio.dsp as d
    from mlx_audio.codec import SNAC, Mimi  # noqa: F401
    from mlx_audio.stt.models.whisper.audio import N_FRAMES, log_mel_spectrogram  # noqa: F401
    from mlx_audio.tts.utils import load_model
    from mlx_audio.utils import hanning, mel_filters, stft  # noqa: F401  (utils.py:31-40 re-exports)
    from mlx_audio_b200 import synth
    from oracle.kokoro import KOKORO_CONFIG
    assert N_FRAMES == 3000 and d.hanning(400).shape == (400,)
    with pytest.raises(FileNotFoundError):
        load_model(tmp_path / "missing")
    (tmp_path / "m").mkdir()
    with pytest.raises(FileNotFoundError, match="Config not found"):
        load_model(tmp_path / "m")
    json.dump({**KOKORO_CONFIG, "model_type": "kokoro", "vocab": {"a": 1}}, open(tmp_path / "m" / "config.json", "w"))
    P = synth.kokoro_weights(KOKORO_CONFIG)
    save_file({k: v.to(torch.bfloat16).contiguous() for k, v in synth.kokoro_to_torch_checkpoint(P).items()},
              str(tmp_path / "m" / "model.safetensors"))
    model = load_model(tmp_path / "m", device="cpu")
    assert model.sample_rate == 24000 and model._w is not None and model.vocab == {"a": 1}
    assert model._w["ups"][0].w.shape == (20, 512, 256)                    # ConvTranspose 512->256 k20 packed [K, Cin, Cout]
    assert all(torch.equal(model.parameters()[k].float(), P[k]) for k in P)  # sanitize inverted the torch layout exactly


def test_kvcache_contract():
    """lm/models/cache.py:104-176: 256-row growth, offset bookkeeping, prefix views, trim."""
    import torch
    from mlx_audio.lm.models.cache import KVCache
    c = KVCache()
    assert c.empty() and c.offset == 0 and c.nbytes == 0
    k1, v1 = torch.randn(1, 2, 3, 8), torch.randn(1, 2, 3, 4)
    k, v = c.update_and_fetch(k1, v1)
    assert k.shape == (1, 2, 3, 8) and v.shape == (1, 2, 3, 4) and c.keys.shape[2] == 256 and c.offset == 3
    big = torch.randn(1, 2, 300, 8)
    k, v = c.update_and_fetch(big, torch.randn(1, 2, 300, 4))
    assert c.offset == 303 and k.shape[2] == 303 and c.keys.shape[2] == 3 + 512          # prev%step != 0: trimmed to prev, then 2 new blocks
    assert torch.equal(k[:, :, :3], k1) and torch.equal(k[:, :, 3:], big)
    assert c.trim(1000) == 303 and c.offset == 0 and c.is_trimmable()
    c.state = (k1, v1)
    assert c.offset == 3 and c.size() == 3


def test_qwen3_config_and_routing_contract():
    """tts/tests/test_models.py:2152-2400 (TestQwen3TTSModel) restated for the parts that need no GPU: config parsing, speaker /
    language lists, shape heuristic of sanitize."""
    import torch
    from mlx_audio.tts.models.qwen3_tts.config import ModelConfig, Qwen3TTSTokenizerDecoderConfig
    from mlx_audio.tts.models.qwen3_tts.speech_tokenizer import Qwen3TTSSpeechTokenizer, check_array_shape_qwen3
    talker = {"vocab_size": 32, "hidden_size": 64, "intermediate_size": 128, "num_hidden_layers": 1, "num_attention_heads": 2,
              "num_key_value_heads": 1, "head_dim": 32, "num_code_groups": 4, "text_hidden_size": 64, "text_vocab_size": 100,
              "codec_eos_token_id": 30, "codec_pad_id": 28, "codec_bos_id": 29, "codec_language_id": {"english": 20, "chinese": 21},
              "spk_id": {"chelsie": 10, "ethan": 11}, "attention_dropout": 0.0,
              "code_predictor_config": {"vocab_size": 32, "hidden_size": 64, "intermediate_size": 128, "num_hidden_layers": 1,
                                        "num_attention_heads": 2, "num_key_value_heads": 1, "head_dim": 32, "num_code_groups": 4}}
    for kind in ("base", "custom_voice", "voice_design"):
        cfg = ModelConfig.from_dict({"model_type": "qwen3_tts", "tts_model_type": kind, "tts_model_size": "0b6", "talker_config": talker,
                                     "speaker_encoder_config": None, "tokenizer_config": None, "sample_rate": 24000})
        assert cfg.model_type == "qwen3_tts" and cfg.tts_model_type == kind and cfg.sample_rate == 24000
        assert cfg.talker_config.code_predictor_config.head_dim == 32 and cfg.talker_config.vocab_size == 32
    d = Qwen3TTSTokenizerDecoderConfig()
    assert (d.upsample_rates, d.upsampling_ratios, d.num_quantizers, d.codebook_size) == ([8, 5, 4, 3], [2, 2], 16, 2048)
    assert check_array_shape_qwen3(torch.zeros(8, 3, 16)) and not check_array_shape_qwen3(torch.zeros(8, 16, 3))
    assert check_array_shape_qwen3(torch.zeros(8, 1, 128)) and not check_array_shape_qwen3(torch.zeros(8, 128, 1))
    w = {"decoder.pre_conv.conv.weight": torch.zeros(8, 16, 3), "decoder.upsample.0.0.conv.weight": torch.zeros(16, 8, 2),
         "decoder.quantizer.rvq_first.vq.layers.0._codebook.cluster_usage": torch.tensor([2.0, 0.0]),
         "decoder.quantizer.rvq_first.vq.layers.0._codebook.embedding_sum": torch.tensor([[2.0, 4.0], [1.0, 1.0]]),
         "encoder.anything": torch.zeros(1)}
    s = Qwen3TTSSpeechTokenizer.sanitize(w)
    assert s["decoder.pre_conv.conv.weight"].shape == (8, 3, 16) and s["decoder.upsample.0.0.conv.weight"].shape == (8, 2, 16)
    emb = s["decoder.quantizer.rvq_first.vq.layers.0.codebook.embed.weight"]
    assert emb[0].tolist() == [1.0, 2.0] and abs(float(emb[1, 0]) - 1e5) < 1.0 and "encoder.anything" not in s


def test_whisper_suppress_token_expansion_follows_the_reference():
    """decoding.py:79-112 / 489-495: a falsy option installs no SuppressTokens filter; otherwise the transcribe / translate / sot /
    sot_prev / sot_lm markers and no_speech are always added, and -1 expands to the tokenizer's non-speech tokens."""
    from mlx_audio.stt.models.whisper.decoding import TokenizerSpec, get_suppress_tokens
    spec = TokenizerSpec()
    assert get_suppress_tokens(spec) == () and get_suppress_tokens(spec, []) == ()
    assert get_suppress_tokens(TokenizerSpec(suppress=(11, 12))) == (11, 12, 50258, 50358, 50359, 50360, 50361, 50362)
    with pytest.raises(ValueError):
        get_suppress_tokens(TokenizerSpec(suppress=(-1,)))
    assert get_suppress_tokens(TokenizerSpec(suppress=(-1, 7), non_speech_tokens=(1, 2))) == (1, 2, 7, 50258, 50358, 50359, 50360, 50361, 50362)


def test_kvcache_follows_the_trace_of_the_reference_class():
    """tests/golden/cache_golden.npz = the reference's lm/models/cache.py:KVCache EXECUTED (NumPy standing in for MLX,
    tests/golden/make_cache_golden.py) through updates that cross the 256-row growth rule in every way (first block, growth from a
    non-multiple offset, exact fill), trims, and a state round trip.  The product class must show the same capacity / offset / fetched
    contents after every operation."""
    import json
    import os
    import numpy as np
    import torch
    from mlx_audio.lm.models.cache import KVCache
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cache_golden.npz"))
    c = KVCache()
    for i, rec in enumerate(json.loads(str(g["trace"]))):
        if rec["op"] == "update":
            fk, fv = c.update_and_fetch(torch.as_tensor(g[f"k_{i}"]), torch.as_tensor(g[f"v_{i}"]))
            assert np.array_equal(fk.numpy(), g[f"fk_{i}"]) and np.array_equal(fv.numpy(), g[f"fv_{i}"]), i
        elif rec["op"] == "trim":
            assert c.trim(rec["n"]) == rec["trimmed"]
        else:
            sk, sv = c.state
            assert sk.shape[2] == rec["state_len"]
            c.state = (sk, sv)
        assert (c.offset, c.keys.shape[2], c.size(), c.empty(), c.is_trimmable()) == (rec["offset"], rec["capacity"], rec["size"], rec["empty"],
                                                                                      rec["trimmable"]), (i, rec)


def test_sanitize_matches_what_the_reference_sanitize_returns():
    """tests/golden/sanitize_golden.json = key -> (shape, CRC-32 of the float32 bytes) of the dictionaries the REFERENCE's own sanitize
    functions return for the synthetic hub-layout checkpoints of tests/golden/checkpoint_layouts.py (make_sanitize_golden.py, NumPy standing in
    for MLX): Whisper from a HuggingFace state dict (whisper.py:551-618), Qwen3-TTS Model.sanitize (qwen3_tts.py:2914-2935), the decoder
    half of Qwen3TTSSpeechTokenizer.sanitize (speech_tokenizer.py:1220-1447) Kokoro's Model.sanitize (kokoro.py:179-276) and the renaming / re-layout of
    Mimi.load_pytorch_weights (mimi.py:192-262).  The product's functions must return the same keys, shapes and
    values for the same inputs."""
    import json
    import os
    import sys
    import zlib
    import numpy as np
    import torch
    here = os.path.join(os.path.dirname(__file__), "golden")
    if here not in sys.path:
        sys.path.insert(0, here)
    import checkpoint_layouts as L
    want = json.load(open(os.path.join(here, "sanitize_golden.json")))

    def manifest(d):
        return {k: [list(v.shape), zlib.crc32(np.ascontiguousarray(v.detach().cpu().float().numpy()).tobytes())] for k, v in d.items()}

    def t(d):
        return {k: torch.as_tensor(v) for k, v in d.items()}
    from mlx_audio_b200.stt.models.whisper.whisper import Model as Whisper
    assert manifest(Whisper.sanitize(Whisper.__new__(Whisper), t(L.whisper_hf()))) == want["whisper_hf"]
    from mlx_audio.tts.models.qwen3_tts.qwen3_tts import Model as Qwen3
    from mlx_audio.tts.models.qwen3_tts.speech_tokenizer import Qwen3TTSSpeechTokenizer
    assert manifest(Qwen3.sanitize(t(L.qwen3_model_torch()))) == want["qwen3_model"]
    assert manifest(Qwen3TTSSpeechTokenizer.sanitize(t(L.qwen3_tokenizer_torch()))) == want["qwen3_tokenizer_decoder"]
    from mlx_audio_b200.tts.models.kokoro.kokoro import Model as Kokoro
    assert manifest(Kokoro.sanitize(Kokoro.__new__(Kokoro), t(L.kokoro_torch()))) == want["kokoro_torch"]
    from mlx_audio_b200.codec.models.mimi import Mimi
    assert manifest(Mimi.sanitize_pytorch_weights(t(L.mimi_torch()))) == want["mimi_torch"]


def test_codec_constructors_accept_what_the_reference_accepts(tmp_path):
    """SNAC.from_config takes the path of a config.json (snac.py:177-182) as well as a dict; Mimi exposes load_pytorch_weights /
    from_pretrained (mimi.py:192-275).  Construction only -- no kernels run."""
    import json
    from mlx_audio_b200.codec.models.mimi import Mimi
    from mlx_audio_b200.codec.models.snac import SNAC
    cfg = dict(sampling_rate=24000, encoder_dim=48, encoder_rates=[2, 4, 8, 8], decoder_dim=1024, decoder_rates=[8, 8, 4, 2], attn_window_size=None,
               codebook_size=4096, codebook_dim=8, vq_strides=[4, 2, 1], noise=True, depthwise=True)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    a, b = SNAC.from_config(str(tmp_path / "config.json"), device="cpu"), SNAC.from_config(cfg, device="cpu")
    assert a.sample_rate == b.sample_rate == 24000 and list(a.vq_strides) == [4, 2, 1]
    assert callable(Mimi.load_pytorch_weights) and callable(Mimi.from_pretrained) and callable(SNAC.from_pretrained)


def test_configs_and_result_contracts_match_the_reference_dataclasses():
    """tests/golden/config_golden.json = what the REFERENCE's own dataclasses parse / declare (make_config_golden.py): Whisper
    ModelDimensions.from_dict on MLX- and HuggingFace-format configs, Qwen3-TTS ModelConfig.from_dict (nested talker / code predictor /
    speaker encoder / tokenizer configs, unknown keys dropped, defaults), Kokoro ModelConfig, and the (name, has-default, default) lists of
    GenerationResult, BatchGenerationResult, DecodingResult."""
    import dataclasses
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(__file__), "golden")
    if here not in sys.path:
        sys.path.insert(0, here)
    import config_cases as C
    want = json.load(open(os.path.join(here, "config_golden.json")))
    from mlx_audio.stt.models.whisper import ModelDimensions
    for k, v in C.WHISPER_CONFIGS.items():
        assert dataclasses.asdict(ModelDimensions.from_dict(v)) == want["whisper_dims"][k], k
    from mlx_audio.tts.models.qwen3_tts.config import ModelConfig as Q
    for k, v in C.QWEN3_CONFIGS.items():
        got = json.loads(json.dumps(dataclasses.asdict(Q.from_dict(v))))
        assert got == want["qwen3_config"][k], (k, {f: (got.get(f), want["qwen3_config"][k].get(f)) for f in set(got) | set(want["qwen3_config"][k])
                                                   if got.get(f) != want["qwen3_config"][k].get(f)})
    from mlx_audio.tts.models.kokoro import ModelConfig as KC
    assert json.loads(json.dumps(dataclasses.asdict(KC.from_dict(C.KOKORO_CONFIG_JSON)))) == want["kokoro_config"]

    def fields(cls):
        out = []
        for f in dataclasses.fields(cls):
            d = None if f.default is dataclasses.MISSING else f.default
            out.append([f.name, f.default is not dataclasses.MISSING or f.default_factory is not dataclasses.MISSING,
                        d if isinstance(d, (int, float, str, bool, type(None))) else repr(d)])
        return out
    from mlx_audio.tts.models.base import BatchGenerationResult, GenerationResult
    assert fields(GenerationResult) == want["result_fields"]["GenerationResult"]
    assert fields(BatchGenerationResult) == want["result_fields"]["BatchGenerationResult"]
    from mlx_audio_b200.stt.models.whisper.whisper import DecodingResult, STTOutput
    assert [f[0] for f in fields(DecodingResult)] == [f[0] for f in want["result_fields"]["DecodingResult"]]
    assert [f[:2] for f in fields(DecodingResult)] == [f[:2] for f in want["result_fields"]["DecodingResult"]]
    assert fields(STTOutput) == want["result_fields"]["STTOutput"]


def test_whisper_decoding_results_are_assembled_like_the_reference_run():
    """tests/golden/whisper_golden.npz holds, next to the raw loop outputs of the reference's DecodingTask._main_loop, the DecodingResult
    objects its run() builds from them (decoding.py:664-722).  The product's results_from_greedy must build the same objects from the same
    loop outputs: token lists cut at the first EOT, avg_logprob = sum / (len + 1), text through the tokenizer, compression ratio."""
    import json
    import os
    import numpy as np
    import torch
    from mlx_audio.stt.models.whisper.decoding import DecodingResult, TokenizerSpec
    from mlx_audio_b200.stt.models.whisper.whisper import results_from_greedy
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "whisper_golden.npz"))
    spec = TokenizerSpec(eot=200, sot=201, no_timestamps=208, timestamp_begin=209, no_speech=207, blank_ids=(7,), language=202, task=203)

    class Tok:
        def decode(self, tokens):
            return " ".join(str(t) for t in tokens)
    for tag, sb in (("ts", 3), ("nots", 4)):
        want = json.loads(str(g[f"dec_{tag}_run"]))
        got = results_from_greedy(g[f"dec_{tag}_tokens"].tolist(), g[f"dec_{tag}_sum_logprobs"], g[f"dec_{tag}_no_speech"], torch.as_tensor(g["xa"]),
                                  spec, sb, tokenizer=Tok())
        assert len(got) == len(want) == 2 and all(isinstance(r, DecodingResult) for r in got)
        for r, w in zip(got, want):
            assert (r.language, r.tokens, r.text, r.temperature) == (w["language"], w["tokens"], w["text"], w["temperature"])
            assert abs(r.avg_logprob - w["avg_logprob"]) < 1e-12 and abs(r.no_speech_prob - w["no_speech_prob"]) < 1e-15
            assert abs(r.compression_ratio - w["compression_ratio"]) < 1e-12
    from mlx_audio_b200.stt.models.whisper.whisper import compression_ratio
    assert compression_ratio("") == 0.0


def test_product_never_touches_the_oracle_or_the_reference():
    """The oracle is test infrastructure: no module of the product (mlx_audio_b200/, the mlx_audio/ shim, the C sources) may import, open or
    mention-by-path ``oracle/`` or ``/root/reference``, and importing every product module must not pull ``oracle`` into sys.modules."""
    import importlib
    import os
    import pkgutil
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference|oracle/_ref|importlib[^\n]*oracle", re.M)
    offenders = []
    for base in ("mlx_audio_b200", "mlx_audio"):
        for dp, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    if pat.search(open(os.path.join(dp, f), errors="ignore").read()):
                        offenders.append(os.path.join(dp, f))
    assert not offenders, offenders
    code = ("import importlib, pkgutil, sys\n"
            "for root in ('mlx_audio_b200', 'mlx_audio'):\n"
            "    pkg = importlib.import_module(root)\n"
            "    for m in pkgutil.walk_packages(pkg.__path__, root + '.'):\n"
            "        if m.name.endswith(('.build', 'libb200audio')):\n"
            "            continue\n"
            "        importlib.import_module(m.name)\n"
            "assert not any(k == 'oracle' or k.startswith('oracle.') for k in sys.modules), 'oracle imported'\n"
            "print('clean')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-1500:]


def test_product_configs_say_what_the_oracle_configs_say():
    """mlx_audio_b200/configs.py (what bench.py and tools/ use) restates the public configurations; the oracle keeps its own copies next
    to the reference citations.  They must agree, and nothing outside tests/, bench.py's CPU arm and smoke() may import oracle/."""
    import os
    import re
    from mlx_audio_b200 import configs as C
    from oracle import codec as OC, kokoro as OK, qwen3 as OQ, whisper as OW
    assert C.KOKORO_82M == OK.KOKORO_CONFIG and C.SNAC_24K == OC.SNAC_24K and C.MIMI_202407 == OC.MIMI_202407
    assert C.WHISPER_SMALL == OW.WHISPER_SMALL and C.QWEN3_TALKER == OQ.TALKER and C.QWEN3_TOKENIZER_DECODER == OQ.TOKENIZER_DECODER
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle[\s.]|import\s+oracle)", re.M)
    for top in ("mlx_audio_b200", "mlx_audio", "tools"):
        for dp, _, fs in os.walk(os.path.join(root, top)):
            for f in fs:
                if f.endswith(".py"):
                    assert not pat.search(open(os.path.join(dp, f)).read()), os.path.join(dp, f)
    # bench.py: the only import sits inside cpu_port_run (the cpu_baseline / --impl reference leg)
    src = open(os.path.join(root, "bench.py")).read()
    hits = [m.start() for m in pat.finditer(src)]
    lo, hi = src.index("def cpu_port_run"), src.index("def host_threads")
    assert hits and all(lo < h < hi for h in hits)
    # bench_workloads.py (the other BASELINE configurations): oracle imports only inside the *_cpu functions (cpu_baseline / reference arm)
    src = open(os.path.join(root, "bench_workloads.py")).read()
    spans = [(m.start(), src.find("\ndef ", m.start() + 1)) for m in re.finditer(r"^def _\w+_cpu\(", src, re.M)]
    hits = [m.start() for m in pat.finditer(src)]
    assert hits and all(any(a < h < (b if b > 0 else len(src)) for a, b in spans) for h in hits)


def test_kokoro_pipeline_chunking_follows_the_reference_pipeline():
    """tests/golden/pipeline_golden.json = the reference's KokoroPipeline.en_tokenize / waterfall_last / join_timestamps and the sentence
    chunking of its non-English branch EXECUTED on synthetic token lists (make_pipeline_golden.py).  The product's chunk_tokens /
    join_timestamps / chunk_text must cut at the same places, and the model is called once per chunk with the voice row picked by the
    phoneme count (pipeline.py:296-303)."""
    import json
    import torch
    from mlx_audio_b200.tts.models.kokoro import pipeline as PL
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pipeline_golden.json")))

    class Tok:
        def __init__(self, text, phonemes, whitespace):
            self.text, self.phonemes, self.whitespace, self.start_ts, self.end_ts = text, phonemes, whitespace, None, None
    for case in g["chunk_cases"]:
        toks = [Tok(*s) for s in case["tokens"]]
        got = [[gs, ps, len(tks)] for gs, ps, tks in PL.chunk_tokens(toks)]
        assert got == case["chunks"] and all(len(c[1]) <= 510 for c in got)
    for case in g["timestamp_cases"]:
        toks = [Tok(*s) for s in case["tokens"]]
        PL.join_timestamps(toks, torch.tensor(case["pred_dur"]))
        assert [[t.start_ts, t.end_ts] for t in toks] == case["stamps"]
    for case in g["text_chunks"]:
        assert [c for c in PL.chunk_text(case["text"]) if c.strip()] == case["chunks"]
    # one model call per chunk; voice = comma-separated packs averaged; style row = pack[len(phonemes) - 1]
    calls = []

    class FakeModel:
        def __call__(self, ps, ref_s, speed, return_output=False):
            calls.append((ps, ref_s.clone(), speed))
            return type("O", (), {"audio": torch.zeros(1, 600 * len(ps)), "pred_dur": torch.ones(len(ps) + 2, dtype=torch.int64)})()
    va, vb = torch.arange(510 * 256, dtype=torch.float32).reshape(510, 1, 256), torch.ones(510, 1, 256)
    pipe = PL.KokoroPipeline("en-us", FakeModel(), g2p=lambda text: ("", [Tok(*s) for s in g["chunk_cases"][1]["tokens"]]))
    pipe.voices = {"af_a": va, "af_b": vb}
    res = list(pipe("ignored", voice="af_a,af_b", speed=1.25))
    assert [r.phonemes for r in res] == [c[1] for c in g["chunk_cases"][1]["chunks"]] and len(calls) == len(res)
    for (ps, ref_s, speed), r in zip(calls, res):
        assert torch.equal(ref_s, ((va + vb) / 2)[len(ps) - 1]) and speed == 1.25 and r.text_index == 0 and r.audio.shape[1] == 600 * len(ps)
    assert any(t.start_ts is not None for t in res[0].tokens)
    with pytest.raises(ValueError, match="Specify a voice"):
        list(pipe("x"))
    with pytest.raises(ValueError, match="too long"):
        list(pipe.generate_from_tokens("a" * 511, voice="af_a"))
    assert [r.phonemes for r in pipe.generate_from_tokens("abc", voice=va)] == ["abc"]
    with pytest.raises(ImportError, match="misaki"):
        PL.KokoroPipeline("a", FakeModel()).g2p


def test_torch_library_ops_are_registered_with_fake_implementations():
    """north_star / SURVEY.md section 8b: the C-ABI entry points are surfaced as torch custom ops (namespace b200audio) with
    ``register_fake``, so they trace without a GPU; on CPU tensors they refuse to run (no CPU fallback)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from mlx_audio_b200 import torch_ops
    for name in torch_ops.OPS:
        assert hasattr(torch.ops.b200audio, name), name
    with FakeTensorMode():
        x = torch.empty(2, 100, 64, device="cuda")
        w = torch.empty(128, 7, 64, device="cuda")
        assert torch.ops.b200audio.conv1d_cl(x, w, None, 1, 3, 9, 1, False, 2, 0.0, 0).shape == (2, 100, 128)
        wt = torch.empty(32, 16, 64, device="cuda")
        assert torch.ops.b200audio.conv1d_cl(x, wt, None, 8, 1, 4, 1, True, 1, 0.1, 0).shape == (2, 800, 32)
        assert torch.ops.b200audio.linear(x, torch.empty(256, 64, device="cuda"), None, 4).shape == (2, 100, 256)
        q = torch.empty(1, 130, 768, device="cuda")
        assert torch.ops.b200audio.attention(q, q, q, 12, 0.125, False, 0).shape == q.shape
        assert torch.ops.b200audio.lstm_bidir(torch.empty(1, 130, 2048, device="cuda"), torch.empty(2, 1024, 256, device="cuda")).shape == (1, 130, 512)
        assert torch.ops.b200audio.layernorm(q, None, None, 1e-5, False).shape == q.shape
        assert torch.ops.b200audio.whisper_logmel(torch.empty(32, 480000, device="cuda"), 80, 480000).shape == (32, 6000, 80)
        codes = torch.empty(1, 32, 63, dtype=torch.int64, device="cuda")
        assert torch.ops.b200audio.rvq_decode(codes, torch.empty(32, 2048, 256, device="cuda")).shape == (1, 63, 256)
        enc = torch.ops.b200audio.rvq_encode(torch.empty(63, 256, device="cuda"), torch.empty(32, 2048, 256, device="cuda"),
                                             torch.empty(32, 2048, dtype=torch.float64, device="cuda"), 0)
        assert enc.shape == (63, 32) and enc.dtype == torch.int64
        re, im = torch.ops.b200audio.stft(torch.empty(1, 16000, device="cuda"), torch.empty(400, device="cuda"), 400, 160, 1)
        assert re.shape == im.shape == (1, 101, 201)
        assert torch.ops.b200audio.kokoro_istft_head(torch.empty(1, 46801, 22, device="cuda")).shape == (1, 234000)
        tok = torch.ops.b200audio.sample_token(torch.empty(8, 3072, device="cuda"), torch.empty(8, device="cuda"), 0.9, 50, 1.0, 0.0)
        assert tok.shape == (8,) and tok.dtype == torch.int64
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.b200audio.layernorm(torch.zeros(2, 8), None, None, 1e-5, False)
