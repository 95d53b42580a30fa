"""Golden vectors from the REFERENCE'S OWN Kokoro pipeline (tts/models/kokoro/pipeline.py: tokens_to_ps / waterfall_last / en_tokenize /
join_timestamps and the non-English sentence chunking of __call__), executed here with NumPy standing in for MLX and a stub for misaki's
token class.  Run from the repo root in the build container:  python tests/golden/make_pipeline_golden.py -> tests/golden/pipeline_golden.json"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy_mlx_nn as shim          # noqa: E402

REF = "/root/reference/mlx_audio"
mx, nn = shim.install(precise=True)
for name, path in (("mlx_audio", REF), ("mlx_audio.tts", f"{REF}/tts"), ("mlx_audio.tts.models", f"{REF}/tts/models"),
                   ("mlx_audio.tts.models.kokoro", f"{REF}/tts/models/kokoro")):
    shim.stub_package(name, path)
hub = types.ModuleType("huggingface_hub")
hub.snapshot_download = hub.hf_hub_download = None
sys.modules["huggingface_hub"] = hub
from mlx_audio.tts.models.kokoro import pipeline as RP          # noqa: E402


class Tok:
    """The attributes of misaki.en.MToken the pipeline touches."""

    def __init__(self, text, phonemes, whitespace):
        self.text, self.phonemes, self.whitespace = text, phonemes, whitespace
        self.start_ts = self.end_ts = None


PUNCT = [".", ",", ";", "!", "?", ":", "—", "…"]


def make_tokens(rng, n_words, punct_every, flap=0.05, none=0.02, bumps=0.1):
    """(text, phonemes, whitespace) triples: words of 1-9 phonemes, punctuation now and then, a few closing brackets / quotes after a mark,
    some flaps (rewritten to T by the pipeline) and some tokens without phonemes."""
    toks = []
    for i in range(n_words):
        n = int(rng.integers(1, 10))
        ph = "".join(rng.choice(list("abdefhijklmnoprstuvwzæɑɪʊθðŋʃʒ"), size=n))
        if rng.random() < flap:
            ph = ph[:-1] + "ɾ"
        if rng.random() < none:
            ph = None
        toks.append((f"w{i}", ph, " "))
        if punct_every and rng.random() < 1.0 / punct_every:
            toks[-1] = (toks[-1][0], toks[-1][1], "")
            toks.append((str(rng.choice(PUNCT)), None, " "))
            toks[-1] = (toks[-1][0], toks[-1][0], " ")
            if rng.random() < bumps:
                toks[-1] = (toks[-1][0], toks[-1][1], "")
                b = str(rng.choice([")", "”"]))
                toks.append((b, b, " "))
    return toks


def main():
    rng = np.random.default_rng(77)
    out = {"chunk_cases": [], "timestamp_cases": [], "text_chunks": []}
    pipe = RP.KokoroPipeline.__new__(RP.KokoroPipeline)
    for n_words, every in ((40, 6), (300, 7), (300, 40), (400, 0), (250, 3), (600, 12)):
        spec = make_tokens(rng, n_words, every)
        toks = [Tok(*s) for s in spec]
        chunks = [(gs, ps, len(tks)) for gs, ps, tks in pipe.en_tokenize(toks)]
        out["chunk_cases"].append({"tokens": spec, "chunks": chunks})
        print("tokens", len(spec), "chunks", [(len(ps), n) for _, ps, n in chunks])
    for n_words in (5, 12):
        spec = make_tokens(rng, n_words, 4, flap=0.0, none=0.1)
        toks = [Tok(*s) for s in spec]
        for t in toks:
            t.phonemes = "" if t.phonemes is None else t.phonemes
        n_ph = len(RP.KokoroPipeline.tokens_to_ps(toks))
        dur = rng.integers(1, 9, size=n_ph + 2)
        RP.KokoroPipeline.join_timestamps(toks, mx.array(dur))
        out["timestamp_cases"].append({"tokens": [[t.text, t.phonemes, t.whitespace] for t in toks], "pred_dur": dur.tolist(),
                                       "stamps": [[t.start_ts, t.end_ts] for t in toks]})
    # the sentence chunking of the non-English branch is inline in __call__: run it through a pipeline object with a recording g2p
    for text in ("Uno. Dos! Tres? " * 40, "x" * 950, "Hola mundo. " + "palabra " * 80 + ". Fin."):
        seen = []
        p2 = RP.KokoroPipeline.__new__(RP.KokoroPipeline)
        p2.lang_code, p2.model, p2.voices = "e", None, {}
        p2.g2p = lambda chunk, _s=seen: (_s.append(chunk) or ("p" * min(len(chunk), 30), None))
        list(p2(text, voice="ef_x", split_pattern=None))
        out["text_chunks"].append({"text": text, "chunks": seen})
    with open(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "pipeline_golden.json"), "w") as f:
        json.dump(out, f, ensure_ascii=False)
    print("wrote pipeline_golden.json")


if __name__ == "__main__":
    main()
