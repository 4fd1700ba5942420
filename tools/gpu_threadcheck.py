"""Torch-free GPU check of the threading contract (a recognizer is shared, streams are independent,
calls from several Python threads are safe and serialise): worker threads accept and decode their own streams on shared
recognizers - one model shared by all threads, and two models used at the same time - and every result must equal the
single-threaded decode of the same audio. Run: python tools/gpu_threadcheck.py"""
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sherpa_vietnamese_asr_b200 import synth, weights  # noqa: E402
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer  # noqa: E402


def decode(rec, audios, split=False):
    ss = []
    for a in audios:
        s = rec.create_stream()
        if split:
            s.accept_waveform(16000, a[: len(a) // 3])
            s.accept_waveform(16000, a[len(a) // 3:])
        else:
            s.accept_waveform(16000, a)
        ss.append(s)
    rec.decode_streams(ss)
    return [(list(s.result.token_ids), list(s.result.frames)) for s in ss]


def main():
    rng = np.random.default_rng(1)
    with tempfile.TemporaryDirectory() as d:
        recs = []
        for model, seed in (("zipformer-tiny", 3), ("zipformer-30m", 30)):
            p = weights.write_model_dir(os.path.join(d, model), weights.CONFIGS[model](), seed)
            recs.append(OfflineRecognizer.from_transducer(encoder=p["encoder"], decoder=p["decoder"], joiner=p["joiner"], tokens=p["tokens"],
                                                          decoding_method="modified_beam_search", max_active_paths=4))
        n_threads, n_iter, n_utt = 4, 5, 6
        audio = {(t, i, u): synth.speech_like(int(rng.integers(8000, 90000)), 8000 + 100 * t + 10 * i + u)
                 for t in range(n_threads) for i in range(n_iter) for u in range(n_utt)}
        want = {(t, i): decode(recs[t % 2], [audio[(t, i, u)] for u in range(n_utt)]) for t in range(n_threads) for i in range(n_iter)}
        got, errors = {}, []

        def worker(t):
            try:
                for i in range(n_iter):
                    got[(t, i)] = decode(recs[t % 2], [audio[(t, i, u)] for u in range(n_utt)], split=bool(i % 2))
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        t0 = time.time()
        threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        bad = sum(got.get(k) != v for k, v in want.items())
        print("threads", n_threads, "batches", len(want), "mismatching batches", bad, "errors", errors, "tokens",
              sum(len(tk) for v in want.values() for tk, _ in v), "wall", round(time.time() - t0, 2), "s", flush=True)
        print("THREADCHECK", "OK" if bad == 0 and not errors else "FAILED")
        return bad or len(errors)


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
