"""CPU suite for the chunk planner and overlap stitcher: our host code against
(a) golden vectors written by the reference's own functions (oracle/make_golden.py -> tests/golden/chunking.json) and
(b) those functions imported live from /root/reference when it is present."""
import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import chunk_cases as cc
from sherpa_vietnamese_asr_b200 import chunking as ck

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"
have_ref = os.path.isdir(os.path.join(REF, "core"))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "chunking.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def ae():
    if not have_ref:
        pytest.skip("/root/reference not present (GPU box)")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with redirect_stdout(io.StringIO()):
        import core.asr_engine as m
    return m


# ----------------------------------------------------------------------------- golden
def test_alignment_golden(gold):
    cases = cc.alignment_cases()
    assert len(cases) == len(gold["alignment"])
    for (tail, head), want in zip(cases, gold["alignment"]):
        assert list(ck.find_overlap_alignment(tail, head)) == want
    assert {w[1] for w in gold["alignment"]} == {"none", "cut_head", "drop_head", "drop_tail"}


def test_merge_golden(gold):
    for chunks, want in zip(cc.merge_cases(), gold["merge"]):
        words, text = ck.merge_chunks_with_overlap(copy.deepcopy(chunks))
        assert text == want["text"]
        assert [w["start"] for w in words] == want["starts"]


def test_silent_regions_golden(gold):
    for (seed, sec), want in zip(cc.silence_cases(), gold["silence"]):
        assert [list(r) for r in ck.find_silent_regions(cc.silence_audio(seed, sec))] == want
    assert sum(len(w) for w in gold["silence"]) > 30


def test_chunk_plan_golden(gold):
    for (total, regions), want, split in zip(cc.plan_cases(), gold["plan"], gold["split"]):
        assert [list(c) for c in ck.plan_chunks(total, regions)] == want
        assert ck.find_best_split_point(total // 2, total, regions) == split


def test_long_segment_and_time_map_golden(gold):
    for (s, e), want in zip(cc.segment_cases(), gold["segment"]):
        assert [list(c) for c in ck.chunk_long_segment(s, e)] == want
    for (segs, total, times), want in zip(cc.offset_map_cases(), gold["offset_map"]):
        audio = np.arange(total, dtype=np.float32)
        concat, omap = ck.concat_vad_speech(audio, segs)
        assert len(concat) == want["len"] and [list(m) for m in omap] == want["map"]
        if segs:
            assert np.array_equal(concat, np.concatenate([audio[a:b] for a, b in segs]))
        tm = ck.ConcatTimeMap(omap)
        assert [tm(t) for t in times] == want["times"]
        assert [ck.map_concat_time_to_original(t, omap) for t in times] == want["times"]


# ----------------------------------------------------------------------------- live against the reference
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_alignment_live(ae, seed):
    with redirect_stdout(io.StringIO()):
        for tail, head in cc.alignment_cases(seed, 80):
            assert ck.find_overlap_alignment(tail, head) == ae.find_overlap_alignment(tail, head)


def test_words_match_live(ae):
    rng = np.random.default_rng(0)
    vocab = [ck.normalize_word_for_overlap(w) for w in cc.SYLLABLES] + ["", "a", "ab", "abc", "abcd", "nghieng", "nghiêng"]
    for w in cc.SYLLABLES + ["  Hà-Nội!  ", "VIỆT", "é"]:
        assert ck.normalize_word_for_overlap(w) == ae.normalize_word_for_overlap(w)
    for _ in range(3000):
        a, b = vocab[int(rng.integers(len(vocab)))], vocab[int(rng.integers(len(vocab)))]
        assert ck.words_match(a, b) == ae.words_match(a, b)


@pytest.mark.parametrize("seed", [201, 202])
def test_merge_live(ae, seed):
    with redirect_stdout(io.StringIO()):
        for chunks in cc.merge_cases(seed, 15):
            got = ck.merge_chunks_with_overlap(copy.deepcopy(chunks))
            want = ae.merge_chunks_with_overlap(copy.deepcopy(chunks))
            assert got == (want[0], want[1])


def test_silence_and_split_live(ae):
    for seed, sec in [(31, 47.0), (32, 200.0), (33, 0.31)]:
        audio = cc.silence_audio(seed, sec)
        regions = ae.find_silent_regions(audio)
        assert ck.find_silent_regions(audio) == regions
        for thr, mind in [(0.02, 0.1), (0.005, 0.5)]:
            assert ck.find_silent_regions(audio, 16000, thr, mind) == ae.find_silent_regions(audio, 16000, thr, mind)
        rng = np.random.default_rng(seed)
        for _ in range(200):
            t = int(rng.integers(0, len(audio) + 1))
            assert ck.find_best_split_point(t, len(audio), regions) == ae.find_best_split_point(t, len(audio), regions)


def test_chunk_plan_live(ae):
    sys.path.insert(0, os.path.dirname(GOLD))
    from oracle.make_golden import reference_plan
    for total, regions in cc.plan_cases(77, 60):
        assert [list(c) for c in ck.plan_chunks(total, regions)] == reference_plan(ae, total, regions)


def test_segments_and_time_map_live(ae):
    rng = np.random.default_rng(4)
    for _ in range(200):
        s = int(rng.integers(0, 10 ** 6))
        e = s + int(rng.uniform(0.1, 400) * 16000)
        assert ck.chunk_long_segment(s, e) == ae.chunk_long_segment(s, e)
    for segs, total, times in cc.offset_map_cases(21, 12):
        audio = np.zeros(total, np.float32)
        _, omap = ae.concat_vad_speech(audio, segs)
        assert ck.concat_vad_speech(audio, segs)[1] == omap
        tm = ck.ConcatTimeMap(omap)
        for t in times:
            assert tm(t) == ae.map_concat_time_to_original(t, omap)


# ----------------------------------------------------------------------------- properties and the batch driver
def test_plan_covers_audio_without_gaps():
    for total, regions in cc.plan_cases(3, 50):
        plan = ck.plan_chunks(total, regions)
        assert plan[0][0] == 0 and plan[-1][1] == total and plan[0][2] == 0
        for (s0, e0, _), (s1, e1, o1) in zip(plan, plan[1:]):
            assert s1 + o1 == e0                               # logical boundaries abut, the overlap reaches back from them
            assert o1 == min(ck.OVERLAP_SAMPLES, e0)
        for s, e, o in plan[:-1]:
            assert 20 * 16000 < e - s - o <= 33 * 16000        # 30 s +- (2 s window + half a silent run of <= 2 s), never under 20 s


def test_merge_of_exact_overlaps_reconstructs_the_word_stream():
    """Two decodes of the overlap that agree word for word: the stitched stream is the original stream."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(40, 200))
        stream = [{"text": f"w{i}", "start": 0.4 * i, "end": 0.4 * i + 0.3, "prob": 0.9} for i in range(n)]
        total = 0.4 * n
        chunks, t = [], 0.0
        while t < total:
            end = min(total, t + float(rng.uniform(10, 20)))
            start = max(0.0, t - 3.0)
            ws = [dict(w, local_start=w["start"] - start, local_end=w["end"] - start) for w in stream if start <= w["start"] < end]
            chunks.append({"words": ws, "audio_start_abs": start, "audio_end_abs": end})
            t = end
        words, text = ck.merge_chunks_with_overlap(chunks)
        assert [w["text"] for w in words] == [w["text"] for w in stream]


class FakeRecognizer:
    pass


def test_transcribe_long_batches_every_chunk_once():
    """One decode_chunks call for the whole recording; chunk audio and offsets follow the plan over the speech-only
    concatenation; word times come back in original-recording time."""
    audio = cc.silence_audio(8, 100.0)
    vad = [(16000 * 2, 16000 * 40), (16000 * 45, 16000 * 99)]
    calls = []

    def fake_decode(rec, chunks, offsets):
        calls.append((chunks, offsets))
        out = []
        for c, off in zip(chunks, offsets):
            n = len(c) // 8000
            out.append([{"text": f"t{int(round((off + 0.5 * i) * 2))}", "start": off + 0.5 * i, "end": off + 0.5 * i + 0.4,
                         "local_start": 0.5 * i, "local_end": 0.5 * i + 0.4, "prob": 0.8} for i in range(n)])
        return out

    res = ck.transcribe_long(FakeRecognizer(), audio, vad, decode_chunks=fake_decode)
    assert len(calls) == 1
    chunks, offsets = calls[0]
    speech, omap = ck.concat_vad_speech(audio, vad)
    plan = ck.plan_chunks(len(speech), ck.find_silent_regions(speech))
    assert res["chunk_plan"] == plan and len(plan) == 4
    for c, off, (s, e, o) in zip(chunks, offsets, plan):
        assert np.array_equal(c, speech[s:e]) and off == s / 16000.0
    # the fake decoder names a word after its concat-time slot: stitched, every slot appears exactly once and in order
    slots = [int(w["text"][1:]) for w in res["words"]]
    assert slots == sorted(set(slots)) and slots[0] == 0 and slots[-1] == len(speech) // 8000 - 1
    assert res["text"] == " ".join(w["text"] for w in res["words"])
    # times are in the original recording: inside a VAD segment and monotone
    starts = [w["start"] for w in res["words"]]
    assert starts == sorted(starts)
    assert all(any(a / 16000.0 <= t <= b / 16000.0 for a, b in vad) for t in starts)
    assert starts[0] == 2.0


def test_transcribe_long_single_chunk_and_no_vad():
    audio = cc.silence_audio(9, 12.0)
    res = ck.transcribe_long(FakeRecognizer(), audio, decode_chunks=lambda r, c, o: [[{"text": "x", "start": 1.0, "end": 1.2,
                                                                                      "local_start": 1.0, "local_end": 1.2}]])
    assert res["chunk_plan"] == [(0, len(audio), 0)] and res["text"] == "x" and res["words"][0]["start"] == 1.0


def test_transcribe_long_with_oracle_decoder(model_dirs):
    """The whole host pipeline on CPU with real word lists: the oracle stands in for the GPU decode (tests only)."""
    from helpers import oracle_recognizer
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d = model_dirs("zipformer-tiny", 3)
    orec = oracle_recognizer(paths, beam=4)[0]

    def oracle_decode(rec, chunks, offsets):
        out = []
        for c, off in zip(chunks, offsets):
            rec["dec_cache"].clear()
            out.append(sr.decode_chunk(rec, c, off, precomputed_features=fbank_ref.fbank(c, np.float64)))
        return out

    audio = synth.speech_like(16000 * 26 + 321, 4100)
    audio[int(9.6 * 16000):int(10.3 * 16000)] *= 0.01
    vad = [(8000, 16000 * 14), (16000 * 15, len(audio) - 4000)]
    res = ck.transcribe_long(orec, audio, vad, decode_chunks=oracle_decode, segment_samples=16000 * 10)
    assert len(res["chunk_plan"]) == 3 and len(res["words"]) > 0
    starts = [w["start"] for w in res["words"]]
    assert all(vad[0][0] / 16000.0 <= t <= vad[1][1] / 16000.0 for t in starts)
    words, text = ck.merge_chunks_with_overlap(copy.deepcopy(res["chunk_results"]))
    assert text == res["text"]
