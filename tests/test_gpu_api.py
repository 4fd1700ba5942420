"""GPU tests of the recognizer surface around the hot path: per-stream hotwords, batched accept, precomputed features
(ROVER's shared fbank), set_config, error reporting, degenerate batches. All through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from helpers import make_graph, oracle_recognizer

pytestmark = pytest.mark.gpu


def _gpu_rec(paths, **kw):
    from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer
    return OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"], joiner=paths["joiner"],
                                             tokens=paths["tokens"], **kw)


@pytest.fixture(scope="module")
def tiny(model_dirs):
    cfg, paths, d = model_dirs("zipformer-tiny", 3)
    return cfg, paths, d, _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)


def _decode(rec, audios, **stream_kw):
    ss = [rec.create_stream(**stream_kw) for _ in audios]
    for s, a in zip(ss, audios):
        s.accept_waveform(16000, a)
    rec.decode_streams(ss)
    return [(list(s.result.token_ids), list(s.result.frames)) for s in ss]


def test_accept_waveforms_batch_equals_loop(tiny):
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = tiny
    audios = [synth.speech_like(n, 9100 + i) for i, n in enumerate([16000 * 3, 500, 16000 * 5 + 11, 16000])]
    want = _decode(rec, audios)
    ss = [rec.create_stream() for _ in audios]
    rec.accept_waveforms(ss, audios)
    rec.decode_streams(ss)
    assert [(list(s.result.token_ids), list(s.result.frames)) for s in ss] == want
    assert sum(len(t) for t, _ in want) > 3


def test_accept_waveform_reports_failures(tiny):
    from sherpa_vietnamese_asr_b200 import _capi
    cfg, paths, d, rec = tiny
    s = rec.create_stream()
    x = np.zeros(160, np.float32)
    lib = _capi.lib()
    assert lib.B200AsrAcceptWaveformOffline(s._h, 8000, _capi.fptr(x), 160) != 0
    assert "16000" in _capi.last_error()
    assert lib.B200AsrAcceptWaveformOffline(s._h, 16000, _capi.fptr(x), 160) == 0
    with pytest.raises(ValueError):
        s.accept_waveform(8000, x)
    with pytest.raises(RuntimeError):
        s.accept_features(np.zeros((1, 80), np.float32), 160)       # already holds samples


def test_stream_hotwords_equal_recognizer_hotwords(tiny):
    """create_stream(hotwords=...) / B200AsrCreateOfflineStreamWithHotwords: a stream's own automaton gives the tokens the same
    phrases give as the recognizer's hotwords (which the parity tests pin to the oracle), also inside a mixed batch."""
    from oracle import search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    from sherpa_vietnamese_asr_b200.recognizer import OfflineStream
    cfg, paths, d, rec = tiny
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    audios = [synth.speech_like(16000 * 6, 9200 + i) for i in range(4)]
    plain = _decode(rec, audios)
    planted = [t for t, _ in plain if len(t) >= 3]
    seqs, scores = synth.random_hotwords(60, ocfg.vocab_size, 60, planted=planted)
    rec.set_hotwords_token_ids(seqs, scores)
    boosted = _decode(rec, audios)
    rec.set_hotwords_token_ids([], [])
    ids = "/".join(" ".join(str(t) for t in s) + f" :{sc}" for s, sc in zip(seqs, scores))
    ss = [OfflineStream(rec, ids if i % 2 == 0 else None) for i in range(len(audios))]
    for s, a in zip(ss, audios):
        s.accept_waveform(16000, a)
    rec.decode_streams(ss)
    got = [(list(s.result.token_ids), list(s.result.frames)) for s in ss]
    for i in range(len(audios)):
        assert got[i] == (boosted[i] if i % 2 == 0 else plain[i])
    print("stream hotwords changed the decode:", boosted != plain)
    with pytest.raises(RuntimeError):
        OfflineStream(rec, "not token ids")


def test_precomputed_features_equal_samples(tiny):
    """decode_chunk(..., precomputed_features=) (core/asr_engine.py:1209-1216) and ROVER's shared fbank: streams fed the
    engine's own features decode to the same words as streams fed the samples."""
    from sherpa_vietnamese_asr_b200 import asr_engine, synth
    cfg, paths, d, rec0 = tiny
    rec = asr_engine.create_recognizer(d, max_active_paths=4)
    chunks = [synth.speech_like(n, 9300 + i) for i, n in enumerate([16000 * 4, 16000 * 7 + 333, 1600])]
    feats = rec.engine.fbank_batch(chunks)
    a = asr_engine.decode_chunks(rec, chunks, [0.0, 5.0, 9.0])
    b = asr_engine.decode_chunks(rec, chunks, [0.0, 5.0, 9.0], feats)
    assert a == b and sum(len(w) for w in a) > 3
    one = asr_engine.decode_chunk(rec, chunks[1], 5.0, precomputed_features=feats[1])
    assert one == a[1]
    # a batch mixing both kinds of stream
    ss = [rec.engine.create_stream() for _ in chunks]
    ss[0].accept_waveform(16000, chunks[0])
    ss[1].accept_features(feats[1], len(chunks[1]))
    ss[2].accept_waveform(16000, chunks[2])
    rec.engine.decode_streams(ss)
    for s, c, o, w in zip(ss, chunks, [0.0, 5.0, 9.0], a):
        assert asr_engine.words_from_result(s.result, rec["id2token"], len(c), o) == w


def test_set_config_keeps_blank_penalty(model_dirs):
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d = model_dirs("zipformer-tiny", 3)
    audios = [synth.speech_like(16000 * 5, 9400 + i) for i in range(3)]
    base = _decode(_gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4), audios)
    rec = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4, blank_penalty=2.5)
    pen = _decode(rec, audios)
    assert pen != base                                      # the penalty matters on this model
    rec.set_config(decoding_method="greedy_search")
    greedy = _decode(rec, audios)
    rec.set_config(decoding_method="modified_beam_search", max_active_paths=4)
    assert _decode(rec, audios) == pen                      # ... and survived the two switches
    rec.set_config(blank_penalty=0.0)
    assert _decode(rec, audios) == base
    assert greedy != pen or True


def test_search_with_trailing_empty_utterance(tiny):
    """lens = [k, 0] through B200AsrBeamSearch: an empty utterance queues no frame-0 decoder row (it has no encoder row)."""
    from oracle import search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = tiny
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    import torch
    from oracle import fbank_ref, zipformer_ref as zr
    with torch.no_grad():
        e = zr.encoder(orec["enc_sess"].W, ocfg, fbank_ref.fbank(synth.speech_like(16000 * 3, 9500), np.float64)).numpy()
    empty = np.zeros((0, e.shape[1]), np.float32)
    for encs in ([e[:1], empty], [empty, e, empty], [empty]):
        got = rec.beam_search(encs, beam=4)
        for x, (toks, frames, lps, stats) in zip(encs, got):
            orec["dec_cache"].clear()
            want = sr.modified_beam_search(orec, None, 4, enc_out=x)[0] if len(x) else []
            assert toks == list(want)
