"""sherpa-onnx compatible recognizer surface over libb200asr.so.

Mirrors what the reference calls on `sherpa_onnx.OfflineRecognizer`
(/root/reference streaming_asr.py:224-243 `from_transducer`, :308 `create_stream`, :285,312,355
`accept_waveform`, :358,408 `decode_stream`, :359 `stream.result.text`;
web_service/audio_quality.py:268-292 reads `result.tokens / timestamps / ys_log_probs`): same names, same
argument meaning, errors raised as exceptions. Model arguments take `.b200w` containers
(weights.write_model_dir) where the reference passes `.onnx` files.
"""
from __future__ import annotations

import ctypes as C
import os
import unicodedata
from typing import List, Optional, Sequence

import numpy as np

from . import _capi


class OfflineRecognitionResult:
    """`stream.result` of sherpa-onnx plus the per-token statistics the reference derives from the joiner
    logits (core/asr_engine.py:1159-1181)."""

    def __init__(self, text="", tokens=None, token_ids=None, timestamps=None, ys_log_probs=None, frames=None,
                 tsallis=None, margin=None, entropy=None, top1=None, num_frames=0, duration=0.0, json=""):
        self.text = text
        self.tokens = tokens or []
        self.token_ids = token_ids or []
        self.timestamps = timestamps or []
        self.ys_log_probs = ys_log_probs or []
        self.frames = frames or []
        self.tsallis = tsallis or []
        self.margin = margin or []
        self.entropy = entropy or []
        self.top1 = top1 or []
        self.num_frames = num_frames
        self.duration = duration
        self.lang = self.emotion = self.event = ""
        self._json = json

    def __str__(self):
        return self._json or self.text


def result_from_struct(r: "_capi.Result") -> OfflineRecognitionResult:
    """Library-owned B200AsrOfflineRecognizerResult -> Python lists. Element-wise ctypes reads: at the 40-100 tokens of a
    chunk they beat a NumPy view per array (12 us against 18 us for 40 tokens)."""
    n = int(r.count)

    def take(ptr):
        return [ptr[i] for i in range(n)]

    return OfflineRecognitionResult(
        text=(r.text or b"").decode("utf-8"), tokens=[(r.tokens[i] or b"").decode("utf-8") for i in range(n)],
        token_ids=take(r.token_ids), timestamps=take(r.timestamps), ys_log_probs=take(r.ys_log_probs),
        frames=take(r.frames), tsallis=take(r.tsallis), margin=take(r.margin), entropy=take(r.entropy),
        top1=take(r.top1), num_frames=r.num_frames, duration=r.duration, json=(r.json or b"").decode("utf-8"))


class OfflineStream:
    def __init__(self, recognizer: "OfflineRecognizer", hotword_ids: Optional[str] = None):
        self._rec = recognizer
        if hotword_ids:
            self._h = _capi.lib().B200AsrCreateOfflineStreamWithHotwords(recognizer._h, hotword_ids.encode())
        else:
            self._h = _capi.lib().B200AsrCreateOfflineStream(recognizer._h)
        if not self._h:
            raise RuntimeError(_capi.last_error())
        self._result: Optional[OfflineRecognitionResult] = None

    def accept_waveform(self, sample_rate: int, waveform) -> None:
        """Copies float samples in [-1, 1]; repeated calls append (streaming_asr.py:285)."""
        w = np.ascontiguousarray(waveform, dtype=np.float32).reshape(-1)
        if int(sample_rate) != 16000:
            raise ValueError("only 16000 Hz input is supported (the reference resamples upstream, core/asr_engine.py:467-518)")
        if _capi.lib().B200AsrAcceptWaveformOffline(self._h, int(sample_rate), _capi.fptr(w), int(w.shape[0])) != 0:
            raise RuntimeError("accept_waveform failed: " + _capi.last_error())
        self._result = None

    def accept_features(self, features, num_samples: int) -> None:
        """Precomputed [T, 80] fbank of `num_samples` samples in place of the samples (decode_chunk's `precomputed_features`)."""
        f = np.ascontiguousarray(features, dtype=np.float32)
        if f.ndim != 2:
            raise ValueError("features must be [T, 80]")
        if _capi.lib().B200AsrAcceptFeaturesOffline(self._h, _capi.fptr(f), int(f.shape[0]), int(f.shape[1]), int(num_samples)) != 0:
            raise RuntimeError("accept_features failed: " + _capi.last_error())
        self._result = None

    @property
    def result(self) -> OfflineRecognitionResult:
        if self._result is None:
            p = _capi.lib().B200AsrGetOfflineStreamResult(self._h)
            if not p:
                return OfflineRecognitionResult()
            self._result = result_from_struct(p.contents)
        return self._result

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _capi.lib().B200AsrDestroyOfflineStream(self._h)
                self._h = None
        except Exception:
            pass


def parse_hotwords_text(text: str, default_score: float = 1.5):
    """`PHRASE :score` lines, `#` comments, NFC + upper-case (core/hotword_context.py:191-222)."""
    out = []
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        score = default_score
        if ":" in line:
            head, tail = line.rsplit(":", 1)
            try:
                score = float(tail.strip())
                line = head.strip()
            except ValueError:
                pass
        phrase = unicodedata.normalize("NFC", line.strip().upper())
        if phrase:
            out.append((phrase, score))
    return out


class _TableTokenizer:
    """Greedy longest-match over tokens.txt pieces, used when no SentencePiece model is supplied."""

    def __init__(self, tokens: Sequence[str]):
        self.tok2id = {t: i for i, t in enumerate(tokens) if t and i > 2}
        self.maxlen = max((len(t) for t in self.tok2id), default=1)

    def encode(self, phrase: str) -> List[int]:
        ids = []
        for word in phrase.split():
            s = "▁" + word
            i = 0
            while i < len(s):
                for ln in range(min(self.maxlen, len(s) - i), 0, -1):
                    t = self.tok2id.get(s[i:i + ln])
                    if t is not None:
                        ids.append(t)
                        i += ln
                        break
                else:
                    return []          # un-encodable phrase is skipped, like an empty sp.encode
        return ids


class OfflineRecognizer:
    def __init__(self):
        raise TypeError("use OfflineRecognizer.from_transducer(...)")

    @classmethod
    def from_transducer(cls, encoder: str, decoder: str, joiner: str, tokens: str, num_threads: int = 1,
                        sample_rate: int = 16000, feature_dim: int = 80, dither: float = 0.0,
                        decoding_method: str = "greedy_search", max_active_paths: int = 4,
                        hotwords_file: str = "", hotwords_score: float = 1.5, blank_penalty: float = 0.0,
                        modeling_unit: str = "cjkchar", bpe_vocab: str = "", debug: bool = False,
                        provider: str = "cuda", model_type: str = "transducer", device_id: int = 0,
                        precision: str = "fp32", bpe_model: str = "", **_ignored) -> "OfflineRecognizer":
        """Keyword-compatible with sherpa_onnx.OfflineRecognizer.from_transducer as called at
        streaming_asr.py:224-243 / core/config.py:405-412. `provider="cpu"` (sherpa's default) is accepted
        as "use the default provider", which here is CUDA; any other provider string is an error."""
        if dither != 0.0:
            raise ValueError("dither must be 0 (core/asr_engine.py:704)")
        self = object.__new__(cls)
        prov = "" if provider in ("cpu", "cuda", "", None) else provider
        cfg = _capi.RecognizerConfig()
        cfg.feat_config.sample_rate = sample_rate
        cfg.feat_config.feature_dim = feature_dim
        cfg.model_config.transducer.encoder = os.fsencode(encoder)
        cfg.model_config.transducer.decoder = os.fsencode(decoder)
        cfg.model_config.transducer.joiner = os.fsencode(joiner)
        cfg.model_config.tokens = os.fsencode(tokens)
        cfg.model_config.num_threads = num_threads
        cfg.model_config.debug = int(bool(debug))
        cfg.model_config.provider = prov.encode()
        cfg.model_config.model_type = model_type.encode()
        cfg.model_config.modeling_unit = modeling_unit.encode()
        cfg.model_config.bpe_vocab = os.fsencode(bpe_vocab) if bpe_vocab else b""
        cfg.decoding_method = decoding_method.encode()
        cfg.max_active_paths = max_active_paths
        cfg.hotwords_file = os.fsencode(hotwords_file) if (hotwords_file and modeling_unit == "token_id") else b""
        cfg.hotwords_score = hotwords_score
        cfg.blank_penalty = blank_penalty
        cfg.device_id = device_id
        modes = {"fp32": 0, "tf32": 1, "fp32_simt": 2, "bf16": 3}
        if precision not in modes:
            raise ValueError(f"precision must be one of {sorted(modes)}")
        cfg.precision = modes[precision]
        self._cfg = cfg
        self._h = _capi.lib().B200AsrCreateOfflineRecognizer(C.byref(cfg))
        if not self._h:
            raise RuntimeError("B200AsrCreateOfflineRecognizer failed: " + _capi.last_error())
        self.vocab_size = _capi.lib().B200AsrVocabSize(self._h)
        self.joiner_dim = _capi.lib().B200AsrEncoderOutDim(self._h)
        self.device_id = device_id
        self.decoding_method = decoding_method
        self.max_active_paths = max_active_paths
        self.hotwords_score = hotwords_score
        self.id2token = {}
        with open(tokens, "r", encoding="utf-8") as f:
            for line in f:
                parts = line.strip().split()
                if len(parts) >= 2:
                    self.id2token[int(parts[-1])] = parts[0]
        self._sp = None
        if not bpe_model and bpe_vocab:        # get_hotwords_config hands over bpe.vocab; the model it was written from sits beside it
            beside = os.path.join(os.path.dirname(os.path.abspath(bpe_vocab)), "bpe.model")
            bpe_model = beside if os.path.exists(beside) else ""
        if bpe_model and os.path.exists(bpe_model):
            import sentencepiece as spm
            self._sp = spm.SentencePieceProcessor()
            self._sp.load(bpe_model)
        if hotwords_file and modeling_unit != "token_id":
            if not os.path.exists(hotwords_file):
                raise FileNotFoundError(hotwords_file)
            with open(hotwords_file, "r", encoding="utf-8") as f:
                self.set_hotwords_text(f.read(), hotwords_score)
        return self

    # ---- hotwords
    def encode_phrase(self, phrase: str) -> List[int]:
        if self._sp is not None:
            return list(self._sp.encode(phrase, out_type=int))
        if not hasattr(self, "_tt"):
            toks = [self.id2token.get(i, "") for i in range(self.vocab_size)]
            self._tt = _TableTokenizer(toks)
        return self._tt.encode(phrase)

    def set_hotwords_text(self, text: str, default_score: float = 1.5) -> int:
        """build_context_graph (core/hotword_context.py:222-259): parse, tokenise, build. Returns phrases kept."""
        seqs, scores = [], []
        for phrase, score in parse_hotwords_text(text, default_score):
            ids = self.encode_phrase(phrase)
            if ids:
                seqs.append(ids)
                scores.append(score)
        self.set_hotwords_token_ids(seqs, scores)
        return len(seqs)

    def set_hotwords_token_ids(self, token_sequences, scores) -> None:
        flat = np.array([t for s in token_sequences for t in s], dtype=np.int32)
        offs = np.zeros(len(token_sequences) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(s) for s in token_sequences])
        sc = np.array(scores, dtype=np.float32)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.int32)
        if sc.size == 0:
            sc = np.zeros(1, dtype=np.float32)
        rc = _capi.lib().B200AsrSetHotwordsTokenIds(self._h, _capi.i32ptr(flat), _capi.i32ptr(offs), _capi.fptr(sc),
                                                   len(token_sequences))
        if rc != 0:
            raise RuntimeError(_capi.last_error())

    def context_forward_one_step(self, state: int, token: int):
        nxt = C.c_int32(0)
        d = _capi.lib().B200AsrContextForwardOneStep(self._h, state, token, C.byref(nxt))
        return d, nxt.value

    def context_finalize(self, state: int) -> float:
        return _capi.lib().B200AsrContextFinalize(self._h, state)

    # ---- sherpa surface
    def create_stream(self, hotwords: Optional[str] = None) -> OfflineStream:
        """`hotwords` (sherpa-onnx: phrases separated by '/', optional ' :score' each) gives the stream its own automaton in
        place of the recognizer's; phrases are parsed and tokenised like a hotwords file (core/hotword_context.py:191-259)."""
        if not hotwords:
            return OfflineStream(self)
        parts = []
        for phrase, score in parse_hotwords_text(hotwords.replace("/", "\n"), self.hotwords_score):
            ids = self.encode_phrase(phrase)
            if ids:
                parts.append(" ".join(str(i) for i in ids) + f" :{score}")
        return OfflineStream(self, "/".join(parts))

    def accept_waveforms(self, streams: Sequence[OfflineStream], waveforms, sample_rate: int = 16000) -> None:
        """accept_waveform for a whole batch in one library call (stream i takes waveforms[i])."""
        n = len(streams)
        if n != len(waveforms):
            raise ValueError("streams and waveforms differ in length")
        if n == 0:
            return
        if int(sample_rate) != 16000:
            raise ValueError("only 16000 Hz input is supported (the reference resamples upstream, core/asr_engine.py:467-518)")
        ws = [np.ascontiguousarray(w, dtype=np.float32).reshape(-1) for w in waveforms]
        hs = (C.c_void_p * n)(*[s._h for s in streams])
        ps = (C.c_void_p * n)(*[w.ctypes.data for w in ws])
        ns = np.array([w.shape[0] for w in ws], dtype=np.int32)
        if _capi.lib().B200AsrAcceptWaveformsOffline(hs, int(sample_rate), ps, _capi.i32ptr(ns), n) != 0:
            raise RuntimeError("accept_waveforms failed: " + _capi.last_error())
        for s in streams:
            s._result = None

    def decode_stream(self, s: OfflineStream) -> None:
        self.decode_streams([s])

    def decode_streams(self, ss: Sequence[OfflineStream]) -> None:
        n = len(ss)
        if n == 0:
            return
        arr = (C.c_void_p * n)(*[s._h for s in ss])
        rc = _capi.lib().B200AsrDecodeMultipleOfflineStreams(self._h, arr, n)
        if rc != 0:
            raise RuntimeError("decode failed: " + _capi.last_error())
        for s in ss:
            s._result = None

    def set_config(self, decoding_method: Optional[str] = None, max_active_paths: Optional[int] = None,
                   hotwords_score: Optional[float] = None, blank_penalty: Optional[float] = None) -> None:
        """Arguments left at None keep their current value (a blank_penalty given to from_transducer survives a
        decoding-method switch)."""
        cfg = _capi.RecognizerConfig()
        cfg.decoding_method = (decoding_method or "").encode()
        cfg.max_active_paths = max_active_paths or 0
        cfg.hotwords_score = hotwords_score if hotwords_score is not None else 0.0
        cfg.blank_penalty = blank_penalty if blank_penalty is not None else float("nan")
        if _capi.lib().B200AsrOfflineRecognizerSetConfig(self._h, C.byref(cfg)) != 0:
            raise RuntimeError(_capi.last_error())
        if hotwords_score:
            self.hotwords_score = hotwords_score
        if decoding_method:
            self.decoding_method = decoding_method
        if max_active_paths:
            self.max_active_paths = max_active_paths

    # ---- raw stages (parity tests, ncu, asr_engine-style sessions)
    def fbank(self, samples) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        T = (x.shape[0] + 80) // 160
        out = np.empty((T, 80), dtype=np.float32)
        if T == 0:
            return out
        rc = _capi.lib().B200AsrFbank(self._h, _capi.fptr(x), x.shape[0], _capi.fptr(out))
        if rc < 0:
            raise RuntimeError(_capi.last_error())
        return out

    def fbank_batch(self, utterances):
        offs = np.zeros(len(utterances) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(u) for u in utterances])
        x = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.float32) for u in utterances]) if utterances else np.zeros(0, np.float32))
        foffs = np.zeros(len(utterances) + 1, dtype=np.int64)
        tot = _capi.lib().B200AsrFbankBatch(self._h, _capi.fptr(x), _capi.i64ptr(offs), len(utterances), None, _capi.i64ptr(foffs))
        out = np.empty((max(tot, 0), 80), dtype=np.float32)
        rc = _capi.lib().B200AsrFbankBatch(self._h, _capi.fptr(x), _capi.i64ptr(offs), len(utterances), _capi.fptr(out), _capi.i64ptr(foffs))
        if rc < 0:
            raise RuntimeError(_capi.last_error())
        return [out[foffs[i]:foffs[i + 1]] for i in range(len(utterances))]

    def encoder(self, feats_list):
        """List of [T,80] arrays -> list of [T',joiner_dim] arrays (ragged batch, unpadded semantics)."""
        lens = np.array([f.shape[0] for f in feats_list], dtype=np.int32)
        x = np.ascontiguousarray(np.concatenate(feats_list, axis=0), dtype=np.float32)
        out_lens = np.zeros(len(feats_list), dtype=np.int32)
        tot = _capi.lib().B200AsrEncoder(self._h, _capi.fptr(x), _capi.i32ptr(lens), len(feats_list), None, _capi.i32ptr(out_lens))
        out = np.empty((max(tot, 0), self.joiner_dim), dtype=np.float32)
        rc = _capi.lib().B200AsrEncoder(self._h, _capi.fptr(x), _capi.i32ptr(lens), len(feats_list), _capi.fptr(out), _capi.i32ptr(out_lens))
        if rc < 0:
            raise RuntimeError(_capi.last_error())
        offs = np.concatenate([[0], np.cumsum(out_lens)])
        return [out[offs[i]:offs[i + 1]] for i in range(len(feats_list))]

    def encoder_tap(self, name: str) -> np.ndarray:
        dim = C.c_int32(0)
        rows = _capi.lib().B200AsrEncoderTap(self._h, name.encode(), None, C.byref(dim))
        if rows < 0:
            raise RuntimeError(_capi.last_error())
        out = np.empty((rows, dim.value), dtype=np.float32)
        _capi.lib().B200AsrEncoderTap(self._h, name.encode(), _capi.fptr(out), C.byref(dim))
        return out

    def decoder(self, y) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.int64).reshape(-1, 2)
        out = np.empty((y.shape[0], self.joiner_dim), dtype=np.float32)
        if _capi.lib().B200AsrDecoder(self._h, _capi.i64ptr(y), y.shape[0], _capi.fptr(out)) != 0:
            raise RuntimeError(_capi.last_error())
        return out

    def joiner(self, enc, dec) -> np.ndarray:
        enc = np.ascontiguousarray(enc, dtype=np.float32)
        dec = np.ascontiguousarray(dec, dtype=np.float32)
        out = np.empty((enc.shape[0], self.vocab_size), dtype=np.float32)
        if _capi.lib().B200AsrJoiner(self._h, _capi.fptr(enc), _capi.fptr(dec), enc.shape[0], _capi.fptr(out)) != 0:
            raise RuntimeError(_capi.last_error())
        return out

    def decoder_joiner_input(self, y, enc=None):
        """The search's decoder kernel on rows y[m,2]: (decoder_out[m,512], tanh(enc + decoder_out)[m,512])."""
        y = np.ascontiguousarray(y, dtype=np.int64).reshape(-1, 2)
        m = y.shape[0]
        dec = np.empty((m, self.joiner_dim), dtype=np.float32)
        x = np.empty((m, self.joiner_dim), dtype=np.float32)
        e = np.ascontiguousarray(enc, dtype=np.float32) if enc is not None else None
        rc = _capi.lib().B200AsrDecoderJoinerInput(self._h, _capi.i64ptr(y), _capi.fptr(e) if e is not None else None, m,
                                                   _capi.fptr(dec), _capi.fptr(x))
        if rc != 0:
            raise RuntimeError(_capi.last_error())
        return dec, x

    def joiner_records(self, x, kb: int = 4) -> np.ndarray:
        """The search's joiner GEMM on rows x[m,512]: records [m, ceil(V/32), 4 + 2*kb] (see include/b200asr.h)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        m = x.shape[0]
        per_row = _capi.lib().B200AsrJoinerRecords(self._h, None, 0, kb, None)
        if per_row < 0:
            raise RuntimeError(_capi.last_error())
        out = np.empty((m, per_row // (4 + 2 * kb), 4 + 2 * kb), dtype=np.float32)
        if _capi.lib().B200AsrJoinerRecords(self._h, _capi.fptr(x), m, kb, _capi.fptr(out)) < 0:
            raise RuntimeError(_capi.last_error())
        return out

    def gemm(self, A, W, bias=None, R=None, act: int = 0, impl: str = "fp32", reps: int = 1):
        """One Linear: act(A W^T + bias) (+ R) through the selected kernel. Returns (C, ms_per_launch)."""
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        M, K = A.shape
        N = W.shape[0]
        out = np.empty((M, N), dtype=np.float32)
        b = np.ascontiguousarray(bias, dtype=np.float32) if bias is not None else None
        Rr = np.ascontiguousarray(R, dtype=np.float32) if R is not None else None
        ms = C.c_float(0)
        rc = _capi.lib().B200AsrGemm(self._h, _capi.fptr(A), _capi.fptr(W), _capi.fptr(b) if b is not None else None,
                                    _capi.fptr(Rr) if Rr is not None else None, _capi.fptr(out), M, N, K, act,
                                    {"fp32": 0, "tc": 1, "tc3": 2, "f16x3": 3, "bf16": 4}[impl], reps, C.byref(ms))
        if rc != 0:
            raise RuntimeError(_capi.last_error())
        return out, ms.value

    def beam_search(self, enc_list, method: str = "modified_beam_search", beam: int = 4):
        """Search from encoder outputs. Returns per utterance (tokens, frames, tok_logprobs, stats[U,4])."""
        lens = np.array([e.shape[0] for e in enc_list], dtype=np.int32)
        x = np.ascontiguousarray(np.concatenate(enc_list, axis=0), dtype=np.float32)
        n = len(enc_list)
        cap = max(1, int(lens.max()) if n else 1)
        toks = np.zeros((n, cap), dtype=np.int32)
        frames = np.zeros((n, cap), dtype=np.int32)
        lps = np.zeros((n, cap), dtype=np.float32)
        stats = np.zeros((n, cap, 4), dtype=np.float32)
        ntok = np.zeros(n, dtype=np.int32)
        rc = _capi.lib().B200AsrBeamSearch(self._h, _capi.fptr(x), _capi.i32ptr(lens), n,
                                          0 if method == "greedy_search" else 1, beam, cap, _capi.i32ptr(toks),
                                          _capi.i32ptr(frames), _capi.fptr(lps), _capi.fptr(stats), _capi.i32ptr(ntok))
        if rc != 0:
            raise RuntimeError(_capi.last_error())
        return [(toks[i, :ntok[i]].tolist(), frames[i, :ntok[i]].tolist(), lps[i, :ntok[i]].tolist(), stats[i, :ntok[i]].copy())
                for i in range(n)]

    # ---- device-resident benchmarking hooks
    def stage_batch(self, utterances) -> int:
        offs = np.zeros(len(utterances) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(u) for u in utterances])
        x = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.float32) for u in utterances]))
        h = _capi.lib().B200AsrStageBatch(self._h, _capi.fptr(x), _capi.i64ptr(offs), len(utterances))
        if h < 0:
            raise RuntimeError(_capi.last_error())
        self._staged_n = getattr(self, "_staged_n", {})
        self._staged_n[h] = len(utterances)
        return h

    def run_staged(self, handle: int) -> np.ndarray:
        ntok = np.zeros(self._staged_n[handle], dtype=np.int32)
        if _capi.lib().B200AsrRunStagedBatch(self._h, handle, _capi.i32ptr(ntok)) != 0:
            raise RuntimeError(_capi.last_error())
        return ntok

    def last_pass_tokens(self, u: int, cap: int = 4096):
        """(token ids, frames) of utterance u of the last staged run / decode pass."""
        toks = np.zeros(cap, dtype=np.int32)
        frames = np.zeros(cap, dtype=np.int32)
        n = _capi.lib().B200AsrLastPassTokens(self._h, int(u), _capi.i32ptr(toks), _capi.i32ptr(frames), cap)
        if n < 0:
            raise RuntimeError(_capi.last_error())
        return toks[:n].tolist(), frames[:n].tolist()

    def last_pipeline_stats(self) -> dict:
        ng, busy, d2h = C.c_int32(0), C.c_float(0), C.c_int64(0)
        lanes = (C.c_float * 8)()
        _capi.lib().B200AsrLastPipelineStats(self._h, C.byref(ng), C.byref(busy), lanes, C.byref(d2h))
        return {"groups": ng.value, "search_busy_ms": busy.value, "lane_ms": [lanes[i] for i in range(ng.value)],
                "d2h_bytes": d2h.value}

    def last_pipeline_timeline(self):
        """Per group of the last pass: (encoder begin, encoder end, search begin, search end) in device ms from the pass start."""
        t = (C.c_float * 32)()
        n = _capi.lib().B200AsrLastPipelineTimeline(self._h, t, 8)
        return [tuple(round(t[g * 4 + k], 3) for k in range(4)) for g in range(max(n, 0))]

    def run_staged_chained(self, handle: int, reps: int):
        """`reps` (<= 8) passes over a staged batch back to back, each pass's search beside the next pass's encoder.
        Returns (token counts of the last pass, device ms for all passes)."""
        ntok = np.zeros(self._staged_n[handle], dtype=np.int32)
        ms = C.c_float(0)
        if _capi.lib().B200AsrRunStagedBatchChained(self._h, handle, int(reps), _capi.i32ptr(ntok), C.byref(ms)) != 0:
            raise RuntimeError(_capi.last_error())
        return ntok, ms.value

    def release_batch(self, handle: int) -> None:
        _capi.lib().B200AsrReleaseBatch(self._h, handle)

    def last_timings(self) -> dict:
        t = (C.c_float * 6)()
        nl = C.c_int64(0)
        _capi.lib().B200AsrLastTimings(self._h, t, C.byref(nl))
        return {"fbank_ms": t[0], "encoder_ms": t[1], "search_ms": t[2], "total_ms": t[3], "h2d_ms": t[4],
                "d2h_ms": t[5], "launches": nl.value}

    def last_gemm_stats(self) -> dict:
        ms, fl, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        _capi.lib().B200AsrLastGemmStats(self._h, C.byref(ms), C.byref(fl), C.byref(n))
        return {"ms": ms.value, "flops": fl.value, "launches": n.value, "bytes": _capi.lib().B200AsrLastGemmBytes(self._h)}

    def set_profiling(self, on: bool) -> None:
        _capi.lib().B200AsrSetProfiling(self._h, int(on))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _capi.lib().B200AsrDestroyOfflineRecognizer(self._h)
                self._h = None
        except Exception:
            pass
