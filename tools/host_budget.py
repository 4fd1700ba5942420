"""Host-side cost of the steps around the recognizer for one 15-minute recording (CPU only): what is left on the host once
the GPU decodes 900 s of audio in ~25 ms. Run: python tools/host_budget.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import chunk_cases as cc  # noqa: E402  (seeded synthetic audio only)
from sherpa_vietnamese_asr_b200 import chunking, postprocess, staging, vad  # noqa: E402


def standin_net(rows):
    return (1.0 - np.exp(-12.0 * np.sqrt(np.mean(rows.astype(np.float32) ** 2, axis=1)))).astype(np.float32)


def fake_decode(_rec, chunks, offsets):
    out = []
    for c, o in zip(chunks, offsets):
        n = int(len(c) / 16000 * 3)        # three words per second
        out.append([{"text": "xin", "start": o + i / 3, "end": o + i / 3 + 0.25, "local_start": i / 3, "local_end": i / 3 + 0.25,
                     "prob": 0.9, "tsallis_max": 0.01, "margin_min": 0.9} for i in range(n)])
    return out


def main():
    audio = np.tile(cc.silence_audio(1, 90.0), 10)
    rows = []

    def timed(name, fn):
        t = time.perf_counter()
        r = fn()
        rows.append((name, (time.perf_counter() - t) * 1e3))
        return r

    segs, probs = timed("VAD host logic (window matrix, stand-in network, segments)", lambda: vad.get_vad_segments(audio, standin_net))
    staged = timed("preprocess_audio (RMS normalise + peak limit)", lambda: staging.preprocess_audio(audio, segs, enable_rms_normalize=True))
    merged = vad.merge_close_segments(segs, vad.MAX_VAD_GAP, True)
    speech, _ = timed("concat_vad_speech", lambda: chunking.concat_vad_speech(staged, merged))
    regions = timed("find_silent_regions (NumPy)", lambda: chunking.find_silent_regions(speech))
    timed("plan_chunks", lambda: chunking.plan_chunks(len(speech), regions))
    res = timed("transcribe_long host side (plan, time map, stitch; decode faked)",
                lambda: chunking.transcribe_long(None, staged, merged, decode_chunks=fake_decode))
    timed("finish_transcript (suspect flags, fillers)", lambda: postprocess.finish_transcript(res["words"], staged, False, probs))
    print(f"{len(audio) / 16000:.0f} s recording, {len(res['chunk_plan'])} chunks, {len(res['words'])} words")
    for name, ms in rows:
        print(f"  {name:70s} {ms:8.1f} ms")
    print(f"  {'total':70s} {sum(ms for _, ms in rows):8.1f} ms   (GPU decode of the same audio at 36 k audio-s/s: {len(audio) / 16000 / 36.0:.0f} ms)")


if __name__ == "__main__":
    main()
