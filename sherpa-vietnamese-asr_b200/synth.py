"""Synthetic workloads named by BASELINE.json `configs`.

No datasets are reachable offline, so the engine is exercised on seeded speech-like 16 kHz signals:
3-5 harmonics of a wandering f0 (90-220 Hz), formant-ish spectral shaping, 3-5 Hz syllabic AM,
150-400 ms pauses and a -35 dB white floor, peak 0.6.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000


def speech_like(n_samples: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(n_samples)
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    # wandering f0: smooth random walk between 90 and 220 Hz
    n_ctrl = max(4, n // 1600 + 2)
    ctrl = np.clip(150 + np.cumsum(rng.normal(0, 12, n_ctrl)), 90, 220)
    f0 = np.interp(np.linspace(0, n_ctrl - 1, n), np.arange(n_ctrl), ctrl)
    phase = 2 * np.pi * np.cumsum(f0) / SAMPLE_RATE
    nh = int(rng.integers(3, 6))
    formants = rng.uniform([300, 900, 2200], [800, 2200, 3400])
    sig = np.zeros(n)
    for h in range(1, nh * 6 + 1):
        fh = f0 * h
        gain = sum(np.exp(-0.5 * ((fh - fc) / 180.0) ** 2) for fc in formants) + 0.05 / h
        sig += gain * np.sin(h * phase + rng.uniform(0, 2 * np.pi))
    # syllabic AM 3-5 Hz
    am_f = rng.uniform(3, 5)
    am = 0.55 + 0.45 * np.sin(2 * np.pi * am_f * t + rng.uniform(0, 2 * np.pi))
    sig *= am
    # pauses 150-400 ms every 1.5-4 s
    gate = np.ones(n)
    pos = int(rng.uniform(0.5, 2.0) * SAMPLE_RATE)
    while pos < n:
        ln = int(rng.uniform(0.15, 0.4) * SAMPLE_RATE)
        gate[pos:pos + ln] = 0.0
        pos += ln + int(rng.uniform(1.5, 4.0) * SAMPLE_RATE)
    k = 160
    if n >= k:
        gate = np.convolve(gate, np.ones(k) / k, mode="same")
    sig *= gate
    peak = np.max(np.abs(sig)) + 1e-9
    sig = sig / peak * 0.6
    sig += rng.normal(0, 0.6 * 10 ** (-35 / 20), n)
    return np.clip(sig, -1.0, 1.0).astype(np.float32)


def c1_clip() -> np.ndarray:
    """Config C1: one 10.0 s clip, seed 1234."""
    return speech_like(160000, 1234)


def c2_durations(n_segments: int = 256, seed: int = 256) -> np.ndarray:
    """Config C2: durations clip(lognormal(ln 9 s, 0.7), 1, 30) s."""
    rng = np.random.default_rng(seed)
    return np.clip(rng.lognormal(np.log(9.0), 0.7, n_segments), 1.0, 30.0)


def c2_segments(n_segments: int = 256, seed: int = 256):
    durs = c2_durations(n_segments, seed)
    return [speech_like(int(round(d * SAMPLE_RATE)), seed * 100003 + i) for i, d in enumerate(durs)]


def planner_chunks(n_chunks: int, seed: int):
    """Pipeline-faithful variant: 20-35 s chunks as produced by the reference's chunk planner
    (core/asr_engine.py:2141-2161: 30 s +/- 2 s silence-snapped windows + 3 s overlap)."""
    rng = np.random.default_rng(seed)
    durs = rng.uniform(28.0, 32.0, n_chunks) + 3.0
    durs[0] -= 3.0
    return [speech_like(int(round(d * SAMPLE_RATE)), seed * 7919 + i) for i, d in enumerate(durs)]


def random_hotwords(n_phrases: int, vocab_size: int, seed: int, planted=None):
    """Config C3: token-id phrases of length 2-8 over ids [3, V); 10 % boosted scores; 20 % share a
    prefix with another phrase; `planted` (list of token lists) supplies phrases cut from real
    decodes so boosts fire. Returns (token_sequences, scores)."""
    rng = np.random.default_rng(seed)
    seqs, scores = [], []
    planted = [p for p in (planted or []) if len(p) >= 2]
    for i in range(n_phrases):
        r = rng.random()
        if planted and r < 0.25:
            src = planted[int(rng.integers(0, len(planted)))]
            ln = int(rng.integers(2, min(8, len(src)) + 1))
            st = int(rng.integers(0, len(src) - ln + 1))
            seq = [int(x) for x in src[st:st + ln]]
        elif seqs and r < 0.45:
            base = seqs[int(rng.integers(0, len(seqs)))]
            pl = int(rng.integers(1, min(3, len(base)) + 1))
            ln = int(rng.integers(max(2, pl + 1), 9))
            seq = list(base[:pl]) + [int(x) for x in rng.integers(3, vocab_size, ln - pl)]
        else:
            ln = int(np.clip(round(rng.normal(4, 1.5)), 2, 8))
            seq = [int(x) for x in rng.integers(3, vocab_size, ln)]
        seqs.append(seq)
        scores.append(float(rng.choice([2.0, 2.5])) if rng.random() < 0.10 else 1.5)
    return seqs, scores


def corpus_recordings(n_files: int = 40, minutes: float = 15.0, seed: int = 36000, bank_seconds: float = 420.0):
    """Config C5: `n_files` recordings of `minutes` each with a speech / silence layout and its ground-truth speech intervals
    (the Silero model is not available offline, so the VAD stage is fed these intervals). A bank of speech-like utterances
    (2-14 s, seeded) is synthesised once; every recording strings randomly chosen, randomly scaled bank utterances together
    with 0.3-4 s pauses of -50 dB noise, so a 10 h corpus costs seconds to build instead of minutes.
    Returns (recordings, intervals): float32 arrays and, per recording, a list of (start_sample, end_sample)."""
    rng = np.random.default_rng(seed)
    bank, tot = [], 0.0
    while tot < bank_seconds:
        d = float(np.clip(rng.lognormal(np.log(6.0), 0.5), 2.0, 14.0))
        bank.append(speech_like(int(round(d * SAMPLE_RATE)), seed * 131 + len(bank)))
        tot += d
    n_total = int(round(minutes * 60 * SAMPLE_RATE))
    recs, ivals = [], []
    for f in range(n_files):
        r = np.random.default_rng(seed + 1 + f)
        out = (r.standard_normal(n_total) * 0.003).astype(np.float32)      # pause floor
        iv, pos = [], int(r.uniform(0.2, 2.0) * SAMPLE_RATE)
        while True:
            u = bank[int(r.integers(0, len(bank)))]
            if pos + len(u) >= n_total:
                break
            out[pos:pos + len(u)] += u * np.float32(r.uniform(0.5, 1.0))
            iv.append((pos, pos + len(u)))
            pos += len(u) + int(np.clip(r.lognormal(np.log(0.8), 0.8), 0.3, 4.0) * SAMPLE_RATE)
        recs.append(out)
        ivals.append(iv)
    return recs, ivals
