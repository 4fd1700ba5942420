"""Torch-free GPU check of the chunk-planner additions (a fresh box pays about a minute for the first `import torch`):
1. csrc/energy.cu flags against NumPy's float32 flags (the body of tests/test_gpu_parity.py::test_energy_scan_flags_equal_numpy);
2. chunking.transcribe_long on the seeded tiny model; the per-chunk word lists are written to gpurun_out/ so that they can be
   compared with the oracle's decode of the same chunks off the box (the comparison itself is
   tests/test_gpu_parity.py::test_transcribe_long_matches_oracle_chunks).
Run: python tools/gpu_quickcheck.py [out.json]"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import chunk_cases as cc  # noqa: E402  (seeded inputs only)
from sherpa_vietnamese_asr_b200 import _capi, asr_engine, chunking, synth, weights  # noqa: E402


def say(*a):
    print(*a, flush=True)


def energy():
    rng = np.random.default_rng(12)
    n_frames = 200_003
    scale = np.exp(rng.uniform(np.log(0.003), np.log(0.03), n_frames)).astype(np.float32)
    scale[::7] = np.float32(0.01) * (1 + rng.uniform(-3e-7, 3e-7, len(scale[::7]))).astype(np.float32)
    x = rng.normal(0, 1, (n_frames, 160)).astype(np.float32)
    x /= np.sqrt(np.mean(x.astype(np.float64) ** 2, axis=1, keepdims=True)).astype(np.float32)
    x = np.ascontiguousarray((x * scale[:, None]).reshape(-1))
    want = np.sqrt(np.mean(x.reshape(n_frames, 160) ** 2, axis=1)) < 0.01
    quiet = np.full(n_frames, 9, dtype=np.uint8)
    t = time.time()
    rc = _capi.lib().B200AsrSilentFrames(_capi.fptr(x), len(x), 16000, float(np.float32(0.01)),
                                         quiet.ctypes.data_as(C.POINTER(C.c_uint8)), 0)
    say("energy rc", rc, _capi.last_error() if rc < 0 else "", "first call s", round(time.time() - t, 3))
    bad = int((quiet.astype(bool) != want).sum())
    say("ENERGY flags mismatches:", bad, "of", n_frames, "quiet share", float(want.mean()))
    ok = rc == n_frames and bad == 0
    for seed, sec in [(3, 7.3), (5, 125.7), (6, 400.0)]:
        audio = cc.silence_audio(seed, sec)
        t = time.time()
        g = chunking.find_silent_regions_gpu(audio)
        tg = time.time() - t
        t = time.time()
        h = chunking.find_silent_regions(audio)
        say("regions", sec, "s audio:", len(g), "gpu", round(tg * 1e3, 2), "ms host", round((time.time() - t) * 1e3, 2), "ms equal", g == h)
        ok = ok and g == h
    say("ENERGY", "OK" if ok else "FAILED")
    return ok


def long_recording(out_path):
    cfg = weights.CONFIGS["zipformer-tiny"]()
    with tempfile.TemporaryDirectory() as d:
        weights.write_model_dir(d, cfg, 3)
        rec = asr_engine.create_recognizer(d, max_active_paths=4)
        audio = synth.speech_like(16000 * 41 + 321, 4100)
        for a, b in [(9.6, 10.3), (19.0, 19.5), (31.2, 32.0)]:
            audio[int(a * 16000):int(b * 16000)] *= 0.01
        vad = [(8000, 16000 * 25), (16000 * 26, len(audio) - 4000)]
        res = chunking.transcribe_long(rec, audio, vad, segment_samples=16000 * 10)
        keep = ("text", "start", "end", "local_start", "local_end", "prob")
        out = {"chunk_plan": res["chunk_plan"], "text": res["text"],
               "chunks": [[{k: w[k] for k in keep} for w in c["words"]] for c in res["chunk_results"]],
               "words": [{k: w[k] for k in keep} for w in res["words"]]}
        with open(out_path, "w", encoding="utf-8") as f:
            json.dump(out, f, ensure_ascii=False)
        say("LONG plan", res["chunk_plan"], "words per chunk", [len(c) for c in out["chunks"]], "stitched", len(out["words"]))


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "quickcheck_long.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    ok = energy()
    long_recording(out)
    sys.exit(0 if ok else 1)
