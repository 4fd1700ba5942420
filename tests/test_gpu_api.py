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


# ----------------------------------------------------------------------------- voice-activity network
@pytest.fixture(scope="module")
def gpu_vad(tmp_path_factory):
    from sherpa_vietnamese_asr_b200 import vad, weights
    W = weights.init_vad_weights(5)
    path = weights.save_vad(str(tmp_path_factory.mktemp("vad") / "silero_vad.b200w"), W)
    return W, vad.GpuVad(path)


def _vad_audio(seed, seconds):
    """Speech-like bursts between near-silent pauses, so the probabilities move."""
    from sherpa_vietnamese_asr_b200 import synth
    rng = np.random.default_rng(seed)
    a = synth.speech_like(int(16000 * seconds), seed)
    t = 0
    while t < len(a):
        t += int(rng.uniform(0.8, 3.0) * 16000)
        gap = int(rng.uniform(0.2, 1.2) * 16000)
        a[t:t + gap] *= np.float32(0.003)
        t += gap
    return a


def test_vad_network_matches_oracle(gpu_vad):
    """csrc/vad.cu against oracle/silero_ref.py on the same seeded weights: probabilities of every 512-sample window within 1e-4
    over a 40 s recording (1250 sequential LSTM steps), identical speech segments through get_vad_segments, and the window
    matrix seam (`prob_fn`) equal to the direct call."""
    from oracle import silero_ref
    from sherpa_vietnamese_asr_b200 import vad
    W, g = gpu_vad
    audio = _vad_audio(11, 40.0)
    got = g.probs(audio)
    want = silero_ref.probs(W, vad.window_matrix(audio))
    assert got.shape == want.shape == (len(audio) // 512,)
    assert np.abs(got - want).max() <= 1e-4, np.abs(got - want).max()
    assert want.std() > 0.02                                  # the network reacts to the input
    np.testing.assert_array_equal(g.prob_fn()(vad.window_matrix(audio)), got)
    thr = float(np.median(want))                              # a threshold in the middle of this (untrained) network's range
    clear = np.abs(want - thr) > 1e-3                         # windows whose side of the threshold is not within the tolerance
    assert ((got >= thr) == (want >= thr))[clear].all()
    if clear.all():
        assert vad.segments_from_probs(got, thr) == vad.segments_from_probs(want, thr)
    print("vad timings:", g.last_timings())


def test_vad_batch_of_recordings_equals_single(gpu_vad):
    """Ragged batch: each recording starts from a zero state and a zero context; results equal the one-recording calls,
    including a recording shorter than one window and one with a partial last window."""
    W, g = gpu_vad
    recs = [_vad_audio(20 + i, s) for i, s in enumerate([6.0, 0.02, 13.7, 1.0])]
    batch = g.probs_batch(recs)
    assert [len(b) for b in batch] == [len(r) // 512 for r in recs]
    for r, b in zip(recs, batch):
        np.testing.assert_array_equal(g.probs(r), b)
    assert g.probs(np.zeros(100, np.float32)).shape == (0,)
