"""TEST INFRASTRUCTURE ONLY. Seeded inputs for the chunk planner / overlap stitcher parity tests
(tests/test_chunking.py) and for the golden vectors oracle/make_golden.py writes from the reference's own functions
(/root/reference core/asr_engine.py:44-237, :521-573, :583-676, :2141-2161)."""
import numpy as np

SYLLABLES = ("xin chào các bạn hôm nay chúng ta sẽ nói về thành phố hồ chí minh và những điều thú vị ở đây một hai ba bốn "
             "năm sáu bảy tám chín mười trăm nghìn triệu người dân đang sinh sống làm việc học tập tại trường đại học quốc gia "
             "Việt Nam, Hà-Nội! kinh tế phát triển nhanh chóng trong thời gian qua nghiên nghiêng nghiêm a ở u").split()


def _words(rng, n, t0, dur, vocab=SYLLABLES):
    """n words spread over [t0, t0 + dur) in local time, with probabilities."""
    starts = np.sort(rng.uniform(0, dur, n))
    out = []
    for i, s in enumerate(starts):
        e = float(starts[i + 1]) if i + 1 < n else dur
        out.append({"text": vocab[int(rng.integers(len(vocab)))], "local_start": round(float(s), 4),
                    "local_end": round(e, 4), "start": round(t0 + float(s), 4), "end": round(t0 + e, 4),
                    "prob": round(float(rng.uniform(0.2, 1.0)), 4)})
    return out


def _perturb(rng, words, p_drop, p_sub, p_fuzz):
    """What a second decode of the same audio looks like: some words missing, replaced or spelled slightly differently."""
    out = []
    for w in words:
        r = rng.uniform()
        if r < p_drop:
            continue
        w = dict(w)
        w["prob"] = round(float(rng.uniform(0.2, 1.0)), 4)
        if r < p_drop + p_sub:
            w["text"] = SYLLABLES[int(rng.integers(len(SYLLABLES)))]
        elif r < p_drop + p_sub + p_fuzz and len(w["text"]) > 2:
            k = int(rng.integers(len(w["text"])))
            w["text"] = w["text"][:k] + w["text"][k + 1:]
        out.append(w)
    return out


def alignment_cases(seed=7, n=60):
    """(tail_words, head_words) pairs: shared stretch + independent edges, at several perturbation strengths."""
    rng = np.random.default_rng(seed)
    cases = [([], []), ([], _words(rng, 3, 0, 3)), (_words(rng, 3, 0, 3), [])]
    for c in range(n):
        n_shared = int(rng.integers(0, 14))
        shared = _words(rng, n_shared, 0, 3.0) if n_shared else []
        level = c % 4
        p = [(0, 0, 0), (0.1, 0.05, 0.1), (0.25, 0.2, 0.2), (0.5, 0.4, 0.0)][level]
        tail = _words(rng, int(rng.integers(0, 4)), 0, 1.0) + shared + _perturb(rng, _words(rng, int(rng.integers(0, 3)), 0, 1.0), 0, 0, 0)
        head = _perturb(rng, shared, *p) + _words(rng, int(rng.integers(0, 6)), 3.0, 2.0)
        cases.append((tail, head))
    # longer than MAX_OVERLAP_WORDS on both sides
    big = _words(rng, 130, 0, 60.0)
    cases.append((big, _perturb(rng, big[-110:], 0.05, 0.05, 0.05) + _words(rng, 20, 3, 5)))
    # identical sequences, and a pure repetition ("một một một")
    rep = [dict(w, text="một") for w in _words(rng, 6, 0, 3)]
    cases.append((rep, [dict(w) for w in rep[2:]]))
    same = _words(rng, 8, 0, 3)
    cases.append((same, [dict(w) for w in same]))
    return cases


def merge_cases(seed=11, n=12):
    """Chunk-result lists as the transcription phase builds them: consecutive chunks whose last/first 3 s hold two
    decodes of the same words."""
    rng = np.random.default_rng(seed)
    cases = [[]]
    for c in range(n):
        n_chunks = int(rng.integers(1, 6))
        level = [(0, 0, 0), (0.1, 0.05, 0.1), (0.3, 0.3, 0.1)][c % 3]
        chunks, t = [], 0.0
        carry = []
        for k in range(n_chunks):
            dur = float(rng.uniform(8.0, 33.0))
            start = t if k == 0 else t - 3.0
            ov = 0.0 if k == 0 else 3.0
            body = _words(rng, int(rng.integers(0, 60)), start + ov, dur - ov)
            for w in body:                                   # local times relative to the chunk start
                w["local_start"] = round(w["local_start"] + ov, 4)
                w["local_end"] = round(w["local_end"] + ov, 4)
            head = []
            for w in _perturb(rng, carry, *level):
                w = dict(w)
                w["local_start"] = round(w["start"] - start, 4)
                w["local_end"] = round(w["end"] - start, 4)
                if 0 <= w["local_start"]:
                    head.append(w)
            words = head + body
            end = start + dur
            chunks.append({"words": words, "audio_start_abs": round(start, 4), "audio_end_abs": round(end, 4),
                           "overlap_sec": ov, "text": " ".join(w["text"] for w in words), "vad_group": 0})
            carry = [w for w in words if w["local_start"] >= dur - 3.0]
            t = end
        cases.append(chunks)
    return cases


def silence_audio(seed, seconds, sr=16000):
    """Speech-like bursts (noise at 0.05-0.3 RMS) separated by quiet stretches (noise at 0.001-0.02) of 0.05-1.5 s."""
    rng = np.random.default_rng(seed)
    out, n = [], 0
    total = int(seconds * sr)
    while n < total:
        burst = int(rng.uniform(0.3, 9.0) * sr)
        out.append(rng.normal(0, rng.uniform(0.05, 0.3), burst))
        gap = int(rng.uniform(0.05, 1.5) * sr)
        out.append(rng.normal(0, rng.uniform(0.001, 0.02), gap))
        n += burst + gap
    return np.concatenate(out)[:total].astype(np.float32)


def silence_cases():
    return [(1, 0.005), (2, 0.2), (3, 7.3), (4, 61.0), (5, 125.7), (6, 400.0)]


def plan_cases(seed=5, n=40):
    """(total_samples, silent_regions) for the chunk plan, without audio: regions placed at random."""
    rng = np.random.default_rng(seed)
    cases = [(0, []), (16000 * 30, []), (16000 * 30 + 1, []), (16000 * 95, [])]
    for _ in range(n):
        total = int(rng.uniform(1, 900) * 16000)
        regions, p = [], 0
        while True:
            p += int(rng.exponential(rng.uniform(1, 25)) * 16000) + 1
            ln = int(rng.uniform(0.3, 2.0) * 16000)
            if p + ln >= total:
                break
            regions.append((p, p + ln))
            p += ln
        cases.append((total, regions))
    return cases


def segment_cases():
    return [(0, 16000 * 5), (100, 100 + 16000 * 30), (0, 16000 * 30 + 1), (777, 777 + 16000 * 61), (0, int(16000 * 89.99)),
            (123456, 123456 + int(16000 * 333.3)), (0, 16000 * 90), (5, 5 + int(16000 * 30.0001))]


def offset_map_cases(seed=9, n=10):
    rng = np.random.default_rng(seed)
    cases = []
    for c in range(n):
        segs, p = [], int(rng.integers(0, 16000))
        for _ in range(int(rng.integers(1, 30))):
            ln = int(rng.integers(0, 8 * 16000)) if c % 3 else int(rng.integers(1, 3) * 16000)
            if c == 4 and rng.uniform() < 0.3:
                ln = 0                                        # empty VAD segments
            segs.append((p, p + ln))
            p += ln + int(rng.integers(0, 5 * 16000))
        total = sum(e - s for s, e in segs)
        times = [-0.5, 0.0, total / 16000.0, total / 16000.0 + 1.0] + [float(x) for x in rng.uniform(0, max(total, 1) / 16000.0, 50)]
        times += [s / 16000.0 for s in np.cumsum([e - s for s, e in segs]).tolist()]
        cases.append((segs, p + 16000, times))
    cases.append(([], 16000, [0.0, 0.5, 2.0]))
    return cases
