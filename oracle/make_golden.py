"""Regenerates tests/golden/* from the reference's own code. Needs /root/reference (this container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

What is pinned, and by what:
  search_cases.npz / .json   `_ort_beam_search` from /root/reference core/asr_engine.py:1023-1153 run UNMODIFIED
                             on duck-typed sessions (oracle decoder/joiner of the seeded zipformer-tiny model)
                             over seeded random encoder outputs, beams 1/4/8, with and without a ContextGraph
                             built by the reference's core/hotword_context.py.
  context_graph.json         ContextGraph.forward_one_step / finalize traces (reference class).
  entropy.json               `_compute_token_entropy` (:1159-1181) on seeded logits rows.
  words.json                 `decode_chunk` (:1209-1326) word lists (with precomputed_features + fake encoder).
  rover.json                 `rover_merge_words` (:1446-1577) on seeded word lists.
  chunking.json              `find_overlap_alignment` / `merge_chunks_with_overlap` (:70-237), `find_silent_regions` /
                             `find_best_split_point` (:521-573), the inline chunk plan (:2141-2161, executed from the
                             reference file), `chunk_long_segment` / `concat_vad_speech` / `map_concat_time_to_original`
                             (:583-676) on the seeded inputs of oracle/chunk_cases.py.
  postprocess.json           `suspect_detect` (with the VAD-probability cache of core/vad_utils.py set to seeded values),
                             `remove_filler_words`, `count_energy_peaks`, `_compute_gap_features`,
                             `compute_disagree_indices` (:1587-1865) and core/asr_json.py `serialize_segments` /
                             `deserialize_segments` on the seeded inputs of oracle/post_cases.py.
  staging.json               core/audio_preprocessing.py `per_segment_rms_normalize` / `preprocess_audio` digests on the
                             seeded cases of tests/test_staging.py.
  fbank_5000.npz             torchaudio.compliance.kaldi.fbank (independent Kaldi restatement) on a seeded clip;
                             kaldi-native-fbank itself is not installable offline.
"""
from __future__ import annotations

import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with redirect_stdout(io.StringIO()):
        import core.asr_engine as ae
        import core.hotword_context as hc
    return ae, hc


def tiny_model():
    from oracle import zipformer_ref
    from sherpa_vietnamese_asr_b200 import weights
    cfg = weights.zipformer_tiny()
    W = weights.init_weights(cfg, 3)
    id2token = {i: t for i, t in enumerate(weights.make_tokens(cfg))}
    return cfg, W, id2token, zipformer_ref


class FixedEncoder:
    """enc_sess stand-in returning a preset encoder_out (the search does not care where it came from)."""

    def __init__(self, enc_out):
        self.enc_out = enc_out

    def run(self, _, feeds):
        return [self.enc_out[None].astype(np.float32), np.array([self.enc_out.shape[0]], dtype=np.int64)]


def search_cases(ae, hc):
    cfg, W, id2token, zr = tiny_model()
    from sherpa_vietnamese_asr_b200 import synth
    rng = np.random.default_rng(2024)
    cases, arrays = [], {}
    for ci, (T, beam, with_graph) in enumerate([(40, 1, False), (60, 4, False), (60, 8, False), (90, 4, True), (75, 8, True),
                                                 (3, 4, False), (120, 4, True)]):
        # smooth-ish random encoder output with enough swing to emit tokens
        base = rng.standard_normal((T, cfg.joiner_dim)).astype(np.float32)
        enc = (0.6 * base + 0.4 * np.roll(base, 1, axis=0)) * 1.6
        rec = zr.make_recognizer(W, cfg, id2token=id2token, max_active_paths=beam)
        rec["enc_sess"] = FixedEncoder(enc)
        graph_spec = None
        if with_graph:
            plain = ae._ort_beam_search(rec, np.zeros((T * 4 + 9, 80), np.float32), beam)[0]
            seqs, scores = synth.random_hotwords(60, cfg.vocab_size, 77 + ci, planted=[plain])
            g = hc.ContextGraph()
            g.build(seqs, scores)
            rec["context_graph"] = g
            rec["dec_cache"] = {}
            graph_spec = {"seqs": seqs, "scores": scores}
        toks, frames, lps, Tout, emit = ae._ort_beam_search(rec, np.zeros((T * 4 + 9, 80), np.float32), beam)
        arrays[f"enc_{ci}"] = enc
        cases.append({"id": ci, "T": T, "beam": beam, "graph": graph_spec, "tokens": [int(t) for t in toks],
                      "frames": [int(f) for f in frames], "tok_lp": [float(x) for x in lps], "T_out": int(Tout),
                      "entropy": [ae._compute_token_entropy(e, cfg.vocab_size) for e in emit]})
    np.savez_compressed(os.path.join(GOLD, "search_cases.npz"), **arrays)
    with open(os.path.join(GOLD, "search_cases.json"), "w") as f:
        json.dump({"model": "zipformer-tiny", "seed": 3, "cases": cases}, f)
    return cases


def context_graph_traces(hc):
    rng = np.random.default_rng(5)
    out = []
    specs = [([[1, 2, 3], [2, 3, 4], [1, 2]], [1.5, 2.0, 1.0]),
             ([[5, 6, 7], [6, 7, 8], [5, 6], [9], [5, 6, 7, 8, 9], [7, 8]], [1.5, 2.0, 1.0, 1.5, 2.5, 1.5]),
             ([[3, 3, 3], [3, 3], [3], [4, 3, 3, 5]], [2.0, 1.5, 1.0, 2.5]),
             ([[7, 8, 9, 10], [8, 9], [9, 10, 11], [], [7, 8]], [1.5, 1.5, 2.0, 1.5, 2.5])]
    for seqs, scores in specs:
        g = hc.ContextGraph()
        g.build(seqs, scores)
        st = g.root
        toks = [int(t) for t in rng.integers(1, 13, 300)]
        deltas, fins = [], []
        for t in toks:
            d, st = g.forward_one_step(st, t)
            deltas.append(float(d))
            fins.append(float(g.finalize(st)))
        out.append({"seqs": seqs, "scores": scores, "tokens": toks, "deltas": deltas, "finalize": fins,
                    "n_phrases": g.n_phrases})
    with open(os.path.join(GOLD, "context_graph.json"), "w") as f:
        json.dump(out, f)


def entropy_cases(ae):
    rng = np.random.default_rng(9)
    out = []
    for V, scale in [(2000, 4.0), (2000, 0.5), (500, 8.0), (50, 2.0)]:
        lg = (rng.standard_normal(V) * scale).astype(np.float32)
        out.append({"V": V, "logits": [float(x) for x in lg], "stats": ae._compute_token_entropy(lg, V)})
    with open(os.path.join(GOLD, "entropy.json"), "w") as f:
        json.dump(out, f)


def word_cases(ae):
    cfg, W, id2token, zr = tiny_model()
    rng = np.random.default_rng(31)
    out, arrays = [], {}
    for ci, (T, n_samples, off) in enumerate([(80, 16000 * 13, 0.0), (50, 16000 * 8 + 11, 12.5)]):
        base = rng.standard_normal((T, cfg.joiner_dim)).astype(np.float32)
        enc = (0.6 * base + 0.4 * np.roll(base, 1, axis=0)) * 1.6
        rec = zr.make_recognizer(W, cfg, id2token=id2token, max_active_paths=4)
        rec["enc_sess"] = FixedEncoder(enc)
        audio = np.zeros(n_samples, dtype=np.float32)
        with redirect_stdout(io.StringIO()):
            words = ae.decode_chunk(rec, audio, off, precomputed_features=np.zeros((T * 4 + 9, 80), np.float32))
        arrays[f"enc_{ci}"] = enc
        out.append({"id": ci, "n_samples": n_samples, "time_offset": off, "words": words})
    np.savez_compressed(os.path.join(GOLD, "words_enc.npz"), **arrays)
    with open(os.path.join(GOLD, "words.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    return out


def rover_cases(ae):
    rng = np.random.default_rng(12)
    vocab = ["xin", "chào", "các", "bạn", "hôm", "nay", "trời", "đẹp", "quá", "ban", "tổ", "chức", "ký", "kết", "hợp", "đồng"]

    def words(n, t0=0.0):
        ws, t = [], t0
        for _ in range(n):
            d = float(rng.uniform(0.1, 0.4))
            ws.append({"text": str(rng.choice(vocab)), "start": t, "end": t + d, "prob": float(rng.uniform(0.3, 1.0)),
                       "margin_min": round(float(rng.uniform(0.0, 1.0)), 4), "tsallis_max": round(float(rng.uniform(0.0, 0.8)), 4)})
            t += d + float(rng.uniform(0.0, 0.2))
        return ws
    out = []
    ae._hotword_phrases_cache = ["ban tổ chức", "hợp đồng"]
    for ci in range(6):
        a = words(int(rng.integers(5, 25)))
        b = copy.deepcopy(a)
        # perturb b: substitutions, deletions, insertions
        for w in b:
            if rng.random() < 0.25:
                w["text"] = str(rng.choice(vocab))
                w["margin_min"] = round(float(rng.uniform(0.0, 1.0)), 4)
        b = [w for w in b if rng.random() > 0.12]
        for _ in range(int(rng.integers(0, 3))):
            k = int(rng.integers(0, len(b) + 1))
            ins = words(1, t0=b[k - 1]["end"] if k > 0 else 0.0)[0]
            b.insert(k, ins)
        ia, ib = copy.deepcopy(a), copy.deepcopy(b)
        with redirect_stdout(io.StringIO()):
            merged, dis = ae.rover_merge_words(a, b)
        out.append({"a": ia, "b": ib, "merged": merged, "disagree": sorted(int(i) for i in dis),
                    "hotword_phrases": ["ban tổ chức", "hợp đồng"]})
    with redirect_stdout(io.StringIO()):
        out.append({"a": [], "b": words(3), "merged": None, "disagree": [], "hotword_phrases": []})
        out[-1]["merged"] = ae.rover_merge_words([], copy.deepcopy(out[-1]["b"]))[0]
    ae._hotword_phrases_cache = None
    with open(os.path.join(GOLD, "rover.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)


def fbank_case():
    import torch
    import torchaudio

    from sherpa_vietnamese_asr_b200 import synth
    a = synth.speech_like(5000, 4242)
    ta = torchaudio.compliance.kaldi.fbank(torch.from_numpy(a)[None].double(), num_mel_bins=80, dither=0.0, snip_edges=False,
                                           window_type="povey", low_freq=20, high_freq=7600, sample_frequency=16000,
                                           energy_floor=1.0)
    np.savez_compressed(os.path.join(GOLD, "fbank_5000.npz"), audio=a, feats=ta.numpy().astype(np.float64))


def reference_plan(ae, total, silent_regions):
    """The chunk plan is inline in the reference's transcription method (core/asr_engine.py:2141-2161); run those very
    lines, read from the file where it lies, on our inputs."""
    import textwrap
    with open(os.path.join(REF, "core", "asr_engine.py"), encoding="utf-8") as f:
        lines = f.read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.strip() == "segment_samples = 16000 * 30")
    b = next(i for i in range(a, len(lines)) if lines[i].strip().startswith('print(f"[Chunk]'))
    ns = {"concat_total": total, "concat_silent_regions": silent_regions, "find_best_split_point": ae.find_best_split_point,
          "OVERLAP_SAMPLES": ae.OVERLAP_SAMPLES}
    exec(textwrap.dedent("\n".join(lines[a:b])), ns)
    return [list(map(int, c)) for c in ns["chunk_plan"]]


def chunking_cases(ae):
    from oracle import chunk_cases as cc
    out = {}
    with redirect_stdout(io.StringIO()):
        out["alignment"] = [list(ae.find_overlap_alignment(t, h)) for t, h in cc.alignment_cases()]
        merges = []
        for chunks in cc.merge_cases():
            words, text = ae.merge_chunks_with_overlap(copy.deepcopy(chunks))
            merges.append({"text": text, "starts": [w["start"] for w in words]})
        out["merge"] = merges
        out["silence"] = [[list(map(int, r)) for r in ae.find_silent_regions(cc.silence_audio(seed, sec))]
                          for seed, sec in cc.silence_cases()]
        out["plan"] = [reference_plan(ae, total, regions) for total, regions in cc.plan_cases()]
        out["split"] = [int(ae.find_best_split_point(total // 2, total, regions)) for total, regions in cc.plan_cases()]
        out["segment"] = [[list(map(int, c)) for c in ae.chunk_long_segment(s, e)] for s, e in cc.segment_cases()]
        maps = []
        for segs, total, times in cc.offset_map_cases():
            audio = np.zeros(total, np.float32)
            concat, omap = ae.concat_vad_speech(audio, segs)
            maps.append({"len": int(len(concat)), "map": [list(map(int, m)) for m in omap],
                         "times": [ae.map_concat_time_to_original(t, omap) for t in times]})
        out["offset_map"] = maps
    with open(os.path.join(GOLD, "chunking.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    return out


def _suspect_summary(words):
    return [[w.get("_suspect_level"), w.get("gap_after_ms"), w.get("gap_before_ms")] for w in words]


def postprocess_cases(ae):
    import core.asr_json as aj
    import core.vad_utils as vu
    from oracle import post_cases as pc
    out = {}
    with redirect_stdout(io.StringIO()):
        sus = []
        for words, audio, disagree, vad in pc.suspect_cases():
            vu._last_vad_probs = vad          # the cache suspect_detect reads (core/vad_utils.py:51-55)
            got = ae.suspect_detect(copy.deepcopy(words), audio, disagree_indices=disagree)
            kept = ae.remove_filler_words(got)
            sus.append({"flags": _suspect_summary(got), "kept": [w["start"] for w in kept]})
        vu._last_vad_probs = None
        out["suspect"] = sus
        out["gap"] = [{"peaks": ae.count_energy_peaks(seg), "features": list(ae._compute_gap_features(seg))} for seg in pc.gap_segments()]
        out["disagree"] = [sorted(ae.compute_disagree_indices(m, o)) for m, o in pc.disagree_cases()]
        ser = []
        for c in pc.segment_cases():
            data = aj.serialize_segments(copy.deepcopy(c["segments"]), c["speaker_name_mapping"], c["speaker_colors"], c["model_name"],
                                         c["model_type"], c["duration_sec"], c["timing"], c["overlap_segments"])
            data.pop("created_at")
            back = aj.deserialize_segments(json.loads(json.dumps(data)))
            ser.append({"json": data, "back": list(back)})
        out["serialize"] = ser
    with open(os.path.join(GOLD, "postprocess.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    return out


def staging_cases_golden():
    """core/audio_preprocessing.py per_segment_rms_normalize / preprocess_audio on the seeded cases of
    tests/test_staging.py; full-length outputs are summarised as (length, sum, sum |x|, max |x|) in float64."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_staging import _digest, staging_cases
    with redirect_stdout(io.StringIO()):
        import core.audio_preprocessing as ap
    out = []
    for audio, segs in staging_cases():
        out.append({"rms": _digest(ap.per_segment_rms_normalize(audio.copy(), segs)), "pre": _digest(ap.preprocess_audio(audio, segs)),
                    "limit": _digest(ap.preprocess_audio(audio, segs, enable_rms_normalize=False))})
    with open(os.path.join(GOLD, "staging.json"), "w", encoding="utf-8") as f:
        json.dump(out, f)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    ae, hc = load_reference()
    cases = search_cases(ae, hc)
    print("search cases:", [(c["T"], c["beam"], len(c["tokens"])) for c in cases])
    context_graph_traces(hc)
    entropy_cases(ae)
    w = word_cases(ae)
    print("word cases:", [len(x["words"]) for x in w])
    rover_cases(ae)
    fbank_case()
    c = chunking_cases(ae)
    print("chunking cases:", {k: len(v) for k, v in c.items()})
    c = postprocess_cases(ae)
    print("postprocess cases:", {k: len(v) for k, v in c.items()})
    print("staging cases:", len(staging_cases_golden()))
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
