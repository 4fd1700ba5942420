"""Compares the JSON written on the GPU box by tools/gpu_edgecheck.py with the oracle's answers for the same seeded cases.
Run (CPU): python tools/verify_edgecheck.py gpurun_out/edgecheck.json"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_edgecheck as ge  # noqa: E402
from helpers import oracle_recognizer  # noqa: E402
from oracle import fbank_ref, search_ref as sr  # noqa: E402
from sherpa_vietnamese_asr_b200 import synth, weights  # noqa: E402


def want_for(orec, audio, beam):
    feats = fbank_ref.fbank(audio, np.float64) if len(audio) else np.zeros((0, 80))
    if feats.shape[0] < 9:
        return {"tokens": [], "frames": [], "num_frames": None, "lps": []}
    orec["dec_cache"].clear()
    toks, frames, lps, T, _ = sr.modified_beam_search(orec, feats, beam)
    return {"tokens": list(toks), "frames": list(frames), "num_frames": int(T), "lps": [float(x) for x in lps]}


def same(got, want, tag):
    ok = got["tokens"] == want["tokens"] and got["frames"] == want["frames"]
    if want["num_frames"] is not None:
        ok = ok and got["num_frames"] == want["num_frames"]
    if ok and want["lps"]:
        ok = float(np.max(np.abs(np.array(got["lps"]) - np.array(want["lps"])))) <= 5e-3
    if not ok:
        print("MISMATCH", tag, got["tokens"][:12], want["tokens"][:12], got["num_frames"], want["num_frames"])
    return ok


def main(path):
    got = json.load(open(path))
    bad = 0
    with tempfile.TemporaryDirectory() as d:
        orecs = {}
        for name, (model, seed, beam, sizes) in ge.cases().items():
            if (model, seed) not in orecs:
                paths = weights.write_model_dir(os.path.join(d, model), weights.CONFIGS[model](), seed)
                orecs[(model, seed)] = oracle_recognizer(paths, beam=beam)[0]
            orec = orecs[(model, seed)]
            n_tok = 0
            for i, n in enumerate(sizes):
                w = want_for(orec, ge.audio_for(name, i, n), beam)
                n_tok += len(w["tokens"])
                bad += not same(got[name][i], w, f"{name}[{i}] n={n}")
            print(name, "compared", len(sizes), "streams,", n_tok, "tokens")
        orec = orecs[("zipformer-tiny", 3)]
        a = synth.speech_like(16000 * 5, 7300)
        bad += not same(got["regrow"]["first"], want_for(orec, a[:32000], 4), "regrow first")
        bad += not same(got["regrow"]["second"], want_for(orec, a, 4), "regrow second")
        bad += not same(got["regrow"]["third"], want_for(orec, a, 4), "regrow third")
    print("EDGECHECK", "OK" if bad == 0 else f"FAILED ({bad})")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1]) else 0)
