"""CPU check of the one-call transcription phase (pipeline.transcribe_recording): it is exactly the composition of the
parity-tested steps, in the reference's order (core/asr_engine.py:2076-2161, :2326-2496, :2556-2580)."""
import copy

import numpy as np

from oracle import chunk_cases as cc
from sherpa_vietnamese_asr_b200 import chunking, pipeline, postprocess, staging, vad


def _fake_decode(rec, chunks, offsets):
    out = []
    for c, off in zip(chunks, offsets):
        n = len(c) // 8000
        out.append([{"text": ["ờ", "xin", "chào", "bạn"][(int(round(off * 2)) + i) % 4], "start": off + 0.5 * i, "end": off + 0.5 * i + 0.3,
                     "local_start": 0.5 * i, "local_end": 0.5 * i + 0.3, "prob": 0.8, "tsallis_max": 0.01 * (i % 9), "margin_min": 0.5}
                    for i in range(n)])
    return out


def _prob(rows):
    return (1.0 - np.exp(-12.0 * np.sqrt(np.mean(rows.astype(np.float64) ** 2, axis=1)))).astype(np.float32)


def test_transcribe_recording_is_the_composition_of_its_steps():
    audio = (cc.silence_audio(21, 95.0) * np.float32(1.7)).astype(np.float32)
    audio[16000 * 40:16000 * 52] = 0                                   # a pause longer than the 5 s merge
    res = pipeline.transcribe_recording(None, audio, vad_prob_fn=_prob, rms_normalize=True, decode_chunks=_fake_decode)

    segs, probs = vad.get_vad_segments(audio, _prob)
    staged = staging.preprocess_audio(audio, segs, enable_rms_normalize=True)
    merged = vad.merge_close_segments(segs, vad.MAX_VAD_GAP, True)
    long = chunking.transcribe_long(None, staged, merged, decode_chunks=_fake_decode)
    words, text = postprocess.finish_transcript(copy.deepcopy(long["words"]), staged, False, probs)
    assert res["vad_segments"] == merged and len(merged) >= 2
    assert res["chunk_plan"] == long["chunk_plan"] and len(res["chunk_plan"]) >= 3
    assert res["words"] == words and res["text"] == text
    assert np.max(np.abs(staged)) <= 0.95 + 1e-6 and np.max(np.abs(audio)) > 0.95
    assert all(w["text"] != "ờ" for w in res["words"]) and any(w.get("_suspect_level") for w in res["words"])
    assert res["text"][0].isupper()


def test_transcribe_recording_without_vad_uses_the_whole_recording():
    audio = cc.silence_audio(22, 33.0)
    res = pipeline.transcribe_recording(None, audio, decode_chunks=_fake_decode)
    assert res["vad_segments"] is None
    assert res["chunk_plan"] == chunking.plan_chunks(len(audio), chunking.find_silent_regions(audio))
    given = pipeline.transcribe_recording(None, audio, vad_segments=[(0, 16000 * 10), (16000 * 12, len(audio))], skip_preprocessing=True,
                                          decode_chunks=_fake_decode)
    assert given["vad_segments"] == [(0, len(audio))]                  # 2 s gap: merged by the 5 s rule


def test_vad_failure_falls_back_to_silence_chunking():
    """core/asr_engine.py:2171-2204: any error in the VAD phase -> the whole recording, chunked at silences."""
    audio = cc.silence_audio(23, 70.0)

    def broken(rows):
        raise RuntimeError("VAD model missing")

    res = pipeline.transcribe_recording(None, audio, vad_prob_fn=broken, decode_chunks=_fake_decode)
    base = pipeline.transcribe_recording(None, audio, decode_chunks=_fake_decode)
    assert "VAD model missing" in res["vad_error"] and base["vad_error"] is None
    assert res["vad_segments"] is None and res["chunk_plan"] == base["chunk_plan"] and res["text"] == base["text"]
    assert set(res["timing"]) == {"vad_preprocess", "transcription", "postprocess"} and all(v >= 0 for v in res["timing"].values())


def test_rover_mode_merges_per_chunk_in_recording_time_and_flags_disagreement():
    """ROVER (core/asr_engine.py:2333-2369, :2469-2486, :2556-2566): both models over the same chunks, word times mapped to
    the recording before the per-chunk merge, disagreeing words end up flagged as suspects."""
    from sherpa_vietnamese_asr_b200 import asr_engine
    audio = cc.silence_audio(24, 50.0)
    vad_segments = [(16000, 16000 * 20), (16000 * 28, 16000 * 49)]
    seen = []

    def decode(rec, chunks, offsets):
        seen.append(rec)
        out = []
        for c, off in zip(chunks, offsets):
            ws = []
            for i in range(len(c) // 8000):
                text = ["xin", "chào", "bạn"][i % 3]
                if rec == "B" and i % 7 == 3:
                    text = "khác"                                   # model B hears another word here
                ws.append({"text": text, "start": off + 0.5 * i, "end": off + 0.5 * i + 0.3, "local_start": 0.5 * i,
                           "local_end": 0.5 * i + 0.3, "prob": 0.9 if rec == "A" else 0.6, "tsallis_max": 0.0, "margin_min": 1.0})
            out.append(ws)
        return out

    res = pipeline.transcribe_recording("A", audio, vad_segments=vad_segments, rover_recognizer="B", skip_preprocessing=True,
                                        decode_chunks=decode)
    assert seen == ["A", "B"]
    single = pipeline.transcribe_recording("A", audio, vad_segments=vad_segments, skip_preprocessing=True, decode_chunks=decode)
    assert res["chunk_plan"] == single["chunk_plan"]
    assert any(w.get("_suspect_level") for w in res["words"]) and not any(w.get("_suspect_level") for w in single["words"])
    assert all("_disagree" not in w for w in res["words"])
    # every chunk is the product's rover_merge_words (itself pinned to the reference's) of the two decodes
    a, b = decode("A", *_chunks(res, audio, vad_segments)), decode("B", *_chunks(res, audio, vad_segments))
    n_dis = 0
    for wa, wb, got in zip(a, b, res["chunk_results"]):
        merged, dis = asr_engine.rover_merge_words(wa, wb)
        assert [w["text"] for w in merged] == [w["text"] for w in got["words"]]
        n_dis += len(dis)
    assert n_dis > 0 and "khác" in res["text"].lower()


def _chunks(res, audio, vad_segments):
    speech, _ = chunking.concat_vad_speech(audio, vad.merge_close_segments(vad_segments, vad.MAX_VAD_GAP, True))
    return [speech[s:e] for s, e, _ in res["chunk_plan"]], [s / 16000.0 for s, _, _ in res["chunk_plan"]]


def test_transcribe_corpus_pools_chunks_across_recordings_and_ranks():
    """C5 shape in miniature: several recordings, chunks pooled across them into length-sorted batches; every recording's
    result equals its own transcribe_recording; two ranks (recordings dealt by duration) give the same answers."""
    recs = [cc.silence_audio(40 + i, sec) for i, sec in enumerate([75.0, 12.0, 140.0, 33.0, 0.5])]
    batches = []

    def decode(rec, chunks, offsets):
        batches.append([len(c) for c in chunks])
        return _fake_decode(rec, chunks, offsets)

    got = pipeline.transcribe_corpus(None, recs, max_batch_seconds=120.0, decode_chunks=decode)
    assert len(got) == len(recs)
    for audio, g in zip(recs, got):
        want = pipeline.transcribe_recording(None, audio, decode_chunks=_fake_decode)
        assert g["chunk_plan"] == want["chunk_plan"] and g["text"] == want["text"]
        assert [(w["text"], w["start"]) for w in g["words"]] == [(w["text"], w["start"]) for w in want["words"]]
    n_chunks = sum(len(g["chunk_plan"]) for g in got)
    assert sum(len(b) for b in batches) == n_chunks and len(batches) < n_chunks          # pooled, not one pass per chunk
    assert all(sum(b) / 16000.0 <= 120.0 + 35.0 for b in batches)
    flat = [x for b in batches for x in b]
    assert flat == sorted(flat, reverse=True)                                          # longest chunks first, across recordings
    # two ranks, gathered in-process
    parts = [None, None]
    outs = []
    for rank in (1, 0):
        def gather(obj, rank=rank):
            parts[rank] = obj
            return parts if rank == 0 else None
        outs.append(pipeline.transcribe_corpus(None, recs, rank=rank, world_size=2, max_batch_seconds=120.0, decode_chunks=_fake_decode,
                                               gather=gather))
    assert outs[0] is None
    assert [g["text"] for g in outs[1]] == [g["text"] for g in got]
    assert all(p for p in parts) and set(parts[0]) | set(parts[1]) == set(range(len(recs))) and not set(parts[0]) & set(parts[1])


def test_corpus_threads_do_not_change_results_and_surface_errors(monkeypatch):
    """transcribe_corpus runs planner threads ahead of the decoder and a finisher thread behind it (the GPU decodes the next
    round during the host-only tail of the last one): the thread counts change nothing in the output, and an exception in
    either side reaches the caller instead of hanging the queues."""
    import pytest
    recs = [cc.silence_audio(60 + i, sec) for i, sec in enumerate([40.0, 75.0, 9.0, 33.0, 61.0, 20.0])]
    base = pipeline.transcribe_corpus(None, recs, max_batch_seconds=90.0, decode_chunks=_fake_decode, planners=1, prefetch=2)
    for planners, prefetch in ((3, 2), (2, 8), (4, 1)):
        got = pipeline.transcribe_corpus(None, recs, max_batch_seconds=90.0, decode_chunks=_fake_decode, planners=planners, prefetch=prefetch)
        assert [g["text"] for g in got] == [g["text"] for g in base]
        assert [g["chunk_plan"] for g in got] == [g["chunk_plan"] for g in base]
        assert [[(w["text"], w["start"], w["end"]) for w in g["words"]] for g in got] == \
               [[(w["text"], w["start"], w["end"]) for w in g["words"]] for g in base]
    stats = {}
    pipeline.transcribe_corpus(None, recs, max_batch_seconds=90.0, decode_chunks=_fake_decode, stats=stats)
    assert stats["recordings"] == len(recs) and stats["chunks"] == sum(len(g["chunk_plan"]) for g in base)

    real_finish = postprocess.finish_transcript
    calls = []

    def failing_finish(*a, **k):
        calls.append(1)
        if len(calls) == 3:
            raise RuntimeError("finisher failed")
        return real_finish(*a, **k)
    monkeypatch.setattr(postprocess, "finish_transcript", failing_finish)
    with pytest.raises(RuntimeError, match="finisher failed"):
        pipeline.transcribe_corpus(None, recs, max_batch_seconds=90.0, decode_chunks=_fake_decode, prefetch=2)
    monkeypatch.setattr(postprocess, "finish_transcript", real_finish)

    def failing_plan(*a, **k):
        raise ValueError("planner failed")
    monkeypatch.setattr(chunking, "plan_recording", failing_plan)
    with pytest.raises(ValueError, match="planner failed"):
        pipeline.transcribe_corpus(None, recs, max_batch_seconds=90.0, decode_chunks=_fake_decode)


def test_corpus_of_one_equals_transcribe_recording_with_vad():
    """Both entry points share prepare_recording (VAD -> preprocess_audio -> 5 s merge, core/asr_engine.py:2076-2128): with the
    same VAD input - a probability function, or given segments - they build the same speech concatenation, chunk plan, words and
    suspect flags (the gap rule needs the VAD probabilities, so those travel too)."""
    audio = (cc.silence_audio(31, 95.0) * np.float32(1.7)).astype(np.float32)
    audio[16000 * 40:16000 * 52] = 0
    want = pipeline.transcribe_recording(None, audio, vad_prob_fn=_prob, rms_normalize=True, decode_chunks=_fake_decode)
    got = pipeline.transcribe_corpus(None, [audio], vad_prob_fn=_prob, rms_normalize=True, decode_chunks=_fake_decode)[0]
    assert got["vad_segments"] == want["vad_segments"] and len(want["vad_segments"]) >= 2
    assert got["chunk_plan"] == want["chunk_plan"] and got["words"] == want["words"] and got["text"] == want["text"]
    assert np.max(np.abs(audio)) > 0.95                                 # the peak limiter had something to do
    segs = [(16000 * 2, 16000 * 30), (16000 * 33, 16000 * 39), (16000 * 55, len(audio) - 8000)]
    want = pipeline.transcribe_recording(None, audio, vad_segments=segs, decode_chunks=_fake_decode)
    stats = {}
    got = pipeline.transcribe_corpus(None, [audio], vad_segments=[segs], decode_chunks=_fake_decode, stats=stats)[0]
    assert got["chunk_plan"] == want["chunk_plan"] and got["words"] == want["words"] and got["text"] == want["text"]
    assert stats["recordings"] == 1 and stats["batches"] >= 1 and stats["chunks"] == len(want["chunk_plan"])


def test_corpus_with_per_recording_vad_functions():
    """vad_prob_fns: one probability function per recording (bench.py's C5 feeds ground-truth intervals this way)."""
    recs = [cc.silence_audio(50 + i, sec) for i, sec in enumerate([48.0, 70.0])]

    def gt(iv):
        def fn(rows):
            p = np.full(rows.shape[0], 0.02, dtype=np.float32)
            for s, e in iv:
                p[s // 512:(e + 511) // 512] = 0.95
            return p
        return fn
    ivs = [[(16000 * 2, 16000 * 20), (16000 * 30, 16000 * 44)], [(16000 * 5, 16000 * 60)]]
    got = pipeline.transcribe_corpus(None, recs, vad_prob_fns=[gt(iv) for iv in ivs], decode_chunks=_fake_decode)
    for audio, iv, g in zip(recs, ivs, got):
        want = pipeline.transcribe_recording(None, audio, vad_prob_fn=gt(iv), decode_chunks=_fake_decode)
        assert g["vad_segments"] == want["vad_segments"] and g["words"] == want["words"]
    assert got[0]["vad_segments"] != got[1]["vad_segments"]


def test_real_engine_branch_wiring(monkeypatch):
    """The branch taken with a real recognizer (no injected decoder): the energy scan goes to the recognizer's GPU and the
    chunks to asr_engine.decode_chunks. Both are replaced here by host stand-ins, so only the wiring is under test."""
    from sherpa_vietnamese_asr_b200 import asr_engine
    calls = {"scan": [], "decode": 0}

    def scan(audio, sample_rate=16000, threshold=0.01, min_silence_duration=0.3, device_id=0):
        calls["scan"].append(device_id)
        return chunking.find_silent_regions(audio, sample_rate, threshold, min_silence_duration)

    def decode(rec, chunks, offsets=None):
        calls["decode"] += 1
        return _fake_decode(rec, chunks, offsets)

    monkeypatch.setattr(chunking, "find_silent_regions_gpu", scan)
    monkeypatch.setattr(asr_engine, "decode_chunks", decode)

    class Engine:
        device_id = 3

    class Rec(dict):
        engine = Engine()

    audio = cc.silence_audio(25, 70.0)
    want = pipeline.transcribe_recording(None, audio, decode_chunks=_fake_decode)
    got = pipeline.transcribe_recording(Rec(), audio)
    assert calls["scan"] == [3] and calls["decode"] == 1
    assert got["text"] == want["text"] and got["chunk_plan"] == want["chunk_plan"]
    calls["scan"].clear()
    corpus = pipeline.transcribe_corpus(Rec(), [audio, audio[: 16000 * 20]])
    assert calls["scan"] == [3, 3] and corpus[0]["text"] == want["text"]
    long = chunking.transcribe_long(Rec(), audio, rover_recognizer=Rec())
    assert long["chunk_plan"] == want["chunk_plan"]
