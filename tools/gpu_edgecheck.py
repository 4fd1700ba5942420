"""Torch-free GPU run of edge cases whose expected answers are computed off the box with the oracle
(tools/verify_edgecheck.py): long utterances (35 s / 61 s, past every size the parity tests use), empty and sub-frame
streams, re-decoding a stream that has grown, and a batch of 300 short streams. Writes gpurun_out/<name>.json.
Run: python tools/gpu_edgecheck.py [out.json]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sherpa_vietnamese_asr_b200 import synth, weights  # noqa: E402
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer  # noqa: E402


def cases():
    """name -> (model, seed, beam, [sample counts])"""
    rng = np.random.default_rng(5)
    return {
        "tiny_long": ("zipformer-tiny", 3, 4, [16000 * 35, 16000 * 61 + 77, 0, 79, 1, 16000 * 2]),
        "m30_35s": ("zipformer-30m", 30, 4, [16000 * 35 + 123]),
        "tiny_many": ("zipformer-tiny", 3, 4, [int(x) for x in rng.integers(1600, 24000, 300)]),
    }


def audio_for(name, i, n):
    base = {"tiny_long": 7000, "m30_35s": 7100, "tiny_many": 7200}[name]
    return synth.speech_like(n, base + i) if n else np.zeros(0, np.float32)


def res(s):
    r = s.result
    return {"tokens": list(r.token_ids), "frames": list(r.frames), "num_frames": int(r.num_frames),
            "lps": [float(x) for x in r.ys_log_probs]}


def main(out_path):
    out = {}
    with tempfile.TemporaryDirectory() as d:
        recs = {}
        for name, (model, seed, beam, sizes) in cases().items():
            t = time.time()
            if (model, seed) not in recs:
                paths = weights.write_model_dir(os.path.join(d, model), weights.CONFIGS[model](), seed)
                recs[(model, seed)] = OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"],
                                                                        joiner=paths["joiner"], tokens=paths["tokens"],
                                                                        decoding_method="modified_beam_search", max_active_paths=beam)
            rec = recs[(model, seed)]
            streams = []
            for i, n in enumerate(sizes):
                s = rec.create_stream()
                if n:
                    s.accept_waveform(16000, audio_for(name, i, n))
                streams.append(s)
            rec.decode_streams(streams)
            out[name] = [res(s) for s in streams]
            print(name, "ok", [len(r["tokens"]) for r in out[name]][:8], round(time.time() - t, 2), "s", flush=True)
        # a stream decoded, grown, decoded again == the whole audio decoded once
        rec = recs[("zipformer-tiny", 3)]
        a = synth.speech_like(16000 * 5, 7300)
        s = rec.create_stream()
        s.accept_waveform(16000, a[:32000])
        rec.decode_stream(s)
        first = res(s)
        s.accept_waveform(16000, a[32000:])
        rec.decode_stream(s)
        again = res(s)
        rec.decode_stream(s)                       # unchanged audio: idempotent
        out["regrow"] = {"first": first, "second": again, "third": res(s)}
        print("regrow ok", len(first["tokens"]), len(again["tokens"]), flush=True)
    with open(out_path, "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    p = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "edgecheck.json")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    main(p)
