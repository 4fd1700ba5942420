"""`core.asr_engine`-shaped entry points over the CUDA engine (Seam 2 of SURVEY.md §8b).

What `TranscriberPipeline` calls in the reference and what stands for it here:

  create_recognizer(model_path, ...)          core/asr_engine.py:903-1020   -> Recognizer (dict-compatible)
  compute_fbank_ort(audio)                    :698-721                      -> compute_fbank_ort (CUDA fbank)
  decode_chunk(recognizer, chunk, t0, feats)  :1209-1326                    -> decode_chunk / decode_chunks (batched)
  rover_merge_words(words_a, words_b)         :1446-1577                    -> rover_merge_words
  find_silent_regions / find_best_split_point / chunk plan / concat_vad_speech / map_concat_time_to_original /
  merge_chunks_with_overlap                   :44-237, :521-676, :2141-2161 -> chunking.py (re-exported here), and
                                                                               chunking.transcribe_long = plan -> one
                                                                               ragged GPU batch -> stitch
  suspect_detect / remove_filler_words / compute_disagree_indices / count_energy_peaks   :1587-1865 -> postprocess.py
  core/asr_json.py serialize_segments / deserialize_segments                  -> postprocess.py
  enc_sess / dec_sess / joi_sess `.run`       :1047,1055,1085,1092          -> session adapters on the raw CUDA stages

The token -> word merge, ROVER and time mapping are host glue in the reference too; the per-token entropy
statistics come out of the CUDA search (no logits rows travel to the host).
"""
from __future__ import annotations

import difflib
import math
import os
import re
import unicodedata
from typing import Dict, List, Optional, Sequence

import numpy as np

from .chunking import (chunk_long_segment, concat_vad_speech, find_best_split_point, find_overlap_alignment,  # noqa: F401
                       find_silent_regions, map_concat_time_to_original, merge_chunks_with_overlap, plan_chunks,
                       transcribe_long, words_match)
from .postprocess import (compute_disagree_indices, count_energy_peaks, finish_transcript, remove_filler_words,  # noqa: F401
                          suspect_detect)
from .recognizer import OfflineRecognizer

ROVER_MODEL_IDS = ["zipformer-30m-rnnt-6000h", "sherpa-onnx-zipformer-vi-2025-04-20"]
HOTWORD_ROVER_BONUS = 0.5
_CONTEXT_WORDS = 3


# ----------------------------------------------------------------------------- session adapters
class _Shape:
    def __init__(self, shape):
        self.shape = shape


class EncoderSession:
    """`enc_sess.run(None, {"x": [N,T,80], "x_lens": [N]})` -> [encoder_out [N,T'max,512], lens] on the GPU."""

    def __init__(self, rec: OfflineRecognizer):
        self._rec = rec

    def run(self, _names, feeds):
        x, lens = np.asarray(feeds["x"], dtype=np.float32), np.asarray(feeds["x_lens"]).astype(np.int64)
        outs = self._rec.encoder([x[n, : int(lens[n])] for n in range(x.shape[0])])
        tmax = max((o.shape[0] for o in outs), default=0)
        out = np.zeros((len(outs), tmax, self._rec.joiner_dim), dtype=np.float32)
        for n, o in enumerate(outs):
            out[n, : o.shape[0]] = o
        return [out, np.array([o.shape[0] for o in outs], dtype=np.int64)]


class DecoderSession:
    def __init__(self, rec: OfflineRecognizer):
        self._rec = rec

    def run(self, _names, feeds):
        return [self._rec.decoder(np.asarray(feeds["y"], dtype=np.int64))]


class JoinerSession:
    def __init__(self, rec: OfflineRecognizer):
        self._rec = rec

    def get_outputs(self):
        return [_Shape(["N", self._rec.vocab_size])]

    def run(self, _names, feeds):
        return [self._rec.joiner(feeds["encoder_out"], feeds["decoder_out"])]


class Recognizer(dict):
    """The dict `create_recognizer` returns (core/asr_engine.py:1005-1012) plus the engine handle."""

    @property
    def engine(self) -> OfflineRecognizer:
        return self["_engine"]


def _find(model_path: str, prefix: str) -> Optional[str]:
    files = sorted(f for f in os.listdir(model_path) if f.startswith(prefix) and f.endswith(".b200w"))
    plain = [f for f in files if "int8" not in f]
    pick = plain or files
    return os.path.join(model_path, pick[0]) if pick else None


def create_recognizer(model_path: str, cpu_threads: int = 4, max_active_paths: int = 8, execution_provider: str = "cuda",
                      hotwords_file: str = "", hotwords_score: float = 1.5, device_id: int = 0,
                      precision: str = "fp32") -> Recognizer:
    """Same discovery rules as the reference (encoder-*/decoder-*/joiner-* + tokens.txt, non-int8 preferred);
    `cpu_threads` is accepted and ignored; any provider other than cuda/cpu-default is an error."""
    enc, dec, joi = (_find(model_path, p) for p in ("encoder-", "decoder-", "joiner-"))
    tokens = os.path.join(model_path, "tokens.txt")
    if not all([enc, dec, joi]) or not os.path.exists(tokens):
        raise FileNotFoundError(f"model files missing in: {model_path}")
    bpe_model = os.path.join(model_path, "bpe.model")
    eng = OfflineRecognizer.from_transducer(
        encoder=enc, decoder=dec, joiner=joi, tokens=tokens, decoding_method="modified_beam_search",
        max_active_paths=max_active_paths, hotwords_file=hotwords_file, hotwords_score=hotwords_score,
        modeling_unit="bpe", provider=execution_provider, device_id=device_id, precision=precision,
        bpe_model=bpe_model if os.path.exists(bpe_model) else "")
    info = {"actual_provider": "B200CudaExecutionProvider"}
    return Recognizer({
        "enc_sess": EncoderSession(eng), "dec_sess": DecoderSession(eng), "joi_sess": JoinerSession(eng),
        "id2token": dict(eng.id2token), "vocab_size": eng.vocab_size, "max_active_paths": max_active_paths,
        "model_path": model_path, "dec_cache": {}, "context_graph": None,
        "provider_info": {"encoder": info, "decoder": info, "joiner": info}, "_engine": eng,
    })


def compute_fbank_ort(recognizer: Recognizer, audio, sr: int = 16000) -> np.ndarray:
    if sr != 16000:
        raise ValueError("16 kHz only")
    return recognizer.engine.fbank(audio)


# ----------------------------------------------------------------------------- tokens -> words
def _mean_like_numpy(xs: List[float]) -> float:
    """np.mean of a short list of Python floats, bit for bit: NumPy adds fewer than 8 values left to right (its pairwise
    scheme starts at 8), which is what a plain loop does; longer lists go to NumPy itself."""
    if len(xs) < 8:
        acc = 0.0
        for x in xs:
            acc += x
        return acc / len(xs)
    return float(np.mean(xs))


def words_from_result(res, id2token: Dict[int, str], n_samples: int, time_offset: float = 0.0) -> List[dict]:
    """Host half of decode_chunk (core/asr_engine.py:1228-1326) from an engine result: pieces lower-cased,
    time = frame / T' * duration, U+2581-merge into words with prob / entropy aggregates and the word-end estimate."""
    ids = list(res.token_ids)
    if not ids or res.num_frames <= 0:
        return []
    pieces = [id2token.get(t, "") for t in ids]
    dur = n_samples / 16000.0
    ts = [f / res.num_frames * dur for f in res.frames]
    step = (ts[-1] - ts[0]) / (len(ts) - 1) if len(ts) >= 2 else 0.08
    # per-token statistics rounded to 4 decimals as _compute_token_entropy returns them (:1176-1181); tuples, not dicts:
    # this loop runs once per token of every chunk
    ents = [(round(float(a), 4), round(float(b), 4), round(float(c), 4))
            for a, b, c in zip(res.tsallis, res.margin, res.entropy)]
    words: List[dict] = []
    for j, piece in enumerate(pieces):
        text = piece.lower()
        start = ts[j]
        end = ts[j + 1] if j + 1 < len(ts) else start + step
        prob = math.exp(res.ys_log_probs[j])
        opens = text.startswith(" ") or text.startswith("▁")
        if opens or not words:
            words.append({"text": text.lstrip(" ").lstrip("▁") if opens else text, "start": start + time_offset,
                          "end": end + time_offset, "local_start": start, "local_end": end,
                          "_last": start + time_offset, "_probs": [prob], "_ents": [ents[j]]})
        else:
            w = words[-1]
            w["text"] += text
            w["end"], w["local_end"], w["_last"] = end + time_offset, end, start + time_offset
            w["_probs"].append(prob)
            w["_ents"].append(ents[j])
    for i, w in enumerate(words):
        probs, es = w.pop("_probs"), w.pop("_ents")
        w["prob"] = sum(probs) / len(probs)
        w["tsallis_max"] = round(max(e[0] for e in es), 4)
        w["margin_min"] = round(min(e[1] for e in es), 4)
        w["entropy_norm"] = round(_mean_like_numpy([e[2] for e in es]), 4)
        w["_conf"] = round(sum(e[1] * (1.0 - e[0]) for e in es) / len(es), 4)
    words[0]["_chunk_bpe_tokens"] = list(pieces)
    words[0]["_chunk_bpe_timestamps_local"] = list(ts)
    for i, w in enumerate(words):
        est = w.pop("_last") + step
        if i + 1 < len(words):
            est = min(est, words[i + 1]["start"])
        w["end"] = est
        w["local_end"] = est - time_offset
    return words


def decode_chunks(recognizer: Recognizer, chunks: Sequence[np.ndarray], time_offsets: Optional[Sequence[float]] = None,
                  precomputed_features: Optional[Sequence[np.ndarray]] = None) -> List[List[dict]]:
    """Batched decode_chunk: all chunks go through one ragged-batch `decode_streams` call. With `precomputed_features`
    (one [T, 80] array per chunk, as compute_fbank_ort returns) the streams carry the features instead of the samples and the
    pass skips its fbank stage - how ROVER runs two models on one fbank (core/asr_engine.py:2346-2350)."""
    eng = recognizer.engine
    offs = list(time_offsets) if time_offsets is not None else [0.0] * len(chunks)
    streams = [eng.create_stream() for _ in chunks]
    if precomputed_features is not None:
        if len(precomputed_features) != len(chunks):
            raise ValueError("precomputed_features and chunks differ in length")
        for s, c, f in zip(streams, chunks, precomputed_features):
            s.accept_features(f, len(c))
    else:
        eng.accept_waveforms(streams, chunks)
    eng.decode_streams(streams)
    return [words_from_result(s.result, recognizer["id2token"], len(c), o) for s, c, o in zip(streams, chunks, offs)]


def decode_chunk(recognizer: Recognizer, audio_chunk, time_offset: float = 0.0, precomputed_features=None) -> List[dict]:
    """Signature of core/asr_engine.py:1209: `precomputed_features` ([T, 80], ROVER's shared fbank) replaces the fbank stage
    when given (:1212-1216)."""
    chunk = np.asarray(audio_chunk, dtype=np.float32)
    feats = None if precomputed_features is None else [np.asarray(precomputed_features, dtype=np.float32)]
    return decode_chunks(recognizer, [chunk], [time_offset], feats)[0]


# ----------------------------------------------------------------------------- ROVER v3
def normalize_word_for_overlap(word: str) -> str:
    return re.sub(r"[^\w]", "", unicodedata.normalize("NFC", word.lower().strip()), flags=re.UNICODE)


def _word_confidence(w: dict) -> float:
    if w.get("margin_min") is not None and w.get("tsallis_max") is not None:
        return w["margin_min"] * (1.0 - w["tsallis_max"])
    return w.get("prob", 0.5)


def _block_confidence(ws: Sequence[dict]) -> float:
    return sum(map(_word_confidence, ws)) / len(ws) if ws else 0.0


def _hotword_share(block, phrases, before, after) -> float:
    if not block or not phrases:
        return 0.0
    seq = list(before or []) + list(block) + list(after or [])
    norms = [normalize_word_for_overlap(w["text"]) for w in seq]
    text = " ".join(norms)
    covered = np.zeros(len(text) + 1, dtype=bool)
    for ph in phrases:
        at = text.find(ph)
        while at >= 0:
            covered[at:at + len(ph)] = True
            at = text.find(ph, at + 1)
    if not covered.any():
        return 0.0
    lo, hi, pos, hits = len(before or []), len(before or []) + len(block), 0, 0
    for i, nw in enumerate(norms):
        a = text.find(nw, pos)
        if a < 0:
            continue
        if lo <= i < hi and covered[a:a + len(nw)].any():
            hits += 1
        pos = a + len(nw)
    return hits / len(block)


def rover_merge_words(words_a: List[dict], words_b: List[dict], hotword_phrases: Sequence[str] = ()):
    """ROVER v3 of the reference: equal/delete keep A, replace picks the block with the higher mean
    margin*(1-tsallis) (+0.5 * hotword share when only one side matches a hotword), B-only inserts kept when
    confidence > 0.20, then time sort and de-duplication of supplements. Returns (merged, disagree indices)."""
    if not words_a:
        return list(words_b or []), set()
    if not words_b:
        return list(words_a), set()
    ops = difflib.SequenceMatcher(None, [normalize_word_for_overlap(w["text"]) for w in words_a],
                                  [normalize_word_for_overlap(w["text"]) for w in words_b], autojunk=False).get_opcodes()
    merged: List[dict] = []
    supplements = 0
    for k, (tag, a0, a1, b0, b1) in enumerate(ops):
        if tag in ("equal", "delete"):
            merged += words_a[a0:a1]
            continue
        if tag == "insert":
            for w in words_b[b0:b1]:
                if _word_confidence(w) > 0.20:
                    w["_source"], w["_disagree"] = "B_supplement", True
                    merged.append(w)
                    supplements += 1
            continue
        blk_a, blk_b = words_a[a0:a1], words_b[b0:b1]
        prev_eq = ops[k - 1] if k > 0 and ops[k - 1][0] == "equal" else None
        next_eq = ops[k + 1] if k + 1 < len(ops) and ops[k + 1][0] == "equal" else None
        ctx = lambda ws, s, e: ws[s:e]
        pa = ctx(words_a, max(prev_eq[1], prev_eq[2] - _CONTEXT_WORDS), prev_eq[2]) if prev_eq else None
        pb = ctx(words_b, max(prev_eq[3], prev_eq[4] - _CONTEXT_WORDS), prev_eq[4]) if prev_eq else None
        na = ctx(words_a, next_eq[1], min(next_eq[2], next_eq[1] + _CONTEXT_WORDS)) if next_eq else None
        nb = ctx(words_b, next_eq[3], min(next_eq[4], next_eq[3] + _CONTEXT_WORDS)) if next_eq else None
        ca, cb = _block_confidence(blk_a), _block_confidence(blk_b)
        ha, hb = _hotword_share(blk_a, hotword_phrases, pa, na), _hotword_share(blk_b, hotword_phrases, pb, nb)
        if ha > 0 and hb == 0:
            ca += ha * HOTWORD_ROVER_BONUS
        elif hb > 0 and ha == 0:
            cb += hb * HOTWORD_ROVER_BONUS
        winner = blk_b if cb > ca else blk_a
        for w in winner:
            w["_disagree"] = True
        merged += winner
    merged.sort(key=lambda w: w["start"])
    if supplements:
        kept: List[dict] = []
        for w in merged:
            if w.get("_source") == "B_supplement":
                nw = normalize_word_for_overlap(w["text"])
                if any(e.get("_source") != "B_supplement" and abs(e["start"] - w["start"]) < 0.15 and
                       normalize_word_for_overlap(e["text"]) == nw for e in kept):
                    continue
            kept.append(w)
        merged = kept
    disagree = {i for i, w in enumerate(merged) if w.get("_disagree")}
    for w in merged:
        w.pop("_source", None)
    return merged, disagree


def rover_decode_chunks(rec_a: Recognizer, rec_b: Recognizer, chunks, time_offsets=None, hotword_phrases=()):
    """ROVER mode (BASELINE config C4): both models over the same chunks, hypotheses combined per chunk
    (core/asr_engine.py:2333-2369,2469-2486)."""
    chunks = [np.asarray(c, dtype=np.float32) for c in chunks]
    feats = rec_a.engine.fbank_batch(chunks) if chunks else []      # one fbank for both models (core/asr_engine.py:2346-2350)
    wa = decode_chunks(rec_a, chunks, time_offsets, feats)
    wb = decode_chunks(rec_b, chunks, time_offsets, feats)
    return [rover_merge_words(a, b, hotword_phrases) for a, b in zip(wa, wb)]
