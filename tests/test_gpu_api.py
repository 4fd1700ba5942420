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


# ----------------------------------------------------------------------------- audio staging
def _staging_case(seed, seconds, scale):
    from oracle import chunk_cases as cc
    rng = np.random.default_rng(seed)
    audio = (cc.silence_audio(seed, seconds) * np.float32(scale)).astype(np.float32)
    n = len(audio)
    segs, t = [], int(rng.uniform(0, 0.5) * 16000)
    while t < n - 2000:
        ln = int(rng.choice([600, 2500, 9000, 40000, 120000]))       # incl. segments under the 100 ms minimum
        e = min(n, t + ln)
        segs.append((t, e))
        audio[t:e] *= np.float32(rng.choice([0.05, 0.3, 1.0, 2.0]))   # loudness differs from segment to segment
        t = e + int(rng.choice([0, 1, 50, 4000, 30000]))              # incl. touching segments (fade onto the neighbour's gain)
    return audio, segs


@pytest.mark.parametrize("seed,seconds,scale", [(1, 30.0, 1.0), (2, 95.0, 1.7), (3, 12.0, 0.2), (4, 0.5, 1.0)])
def test_device_staging_matches_host_preprocess(seed, seconds, scale):
    """csrc/staging.cu against staging.preprocess_audio (itself bit-equal to core/audio_preprocessing.py:46-292 in
    tests/test_staging.py): peak limiter only -> bit-equal; with per-segment RMS normalisation -> 1e-6 of the peak (segment RMS is
    accumulated in float64 on the device); with the load step's low-volume boost in front as well."""
    from sherpa_vietnamese_asr_b200 import staging
    audio, segs = _staging_case(seed, seconds, scale)
    want = staging.preprocess_audio(audio, segs, enable_rms_normalize=False)
    got = staging.preprocess_audio_gpu(audio, segs, enable_rms_normalize=False)
    np.testing.assert_array_equal(got, want)
    want = staging.preprocess_audio(audio, segs, enable_rms_normalize=True)
    got = staging.preprocess_audio_gpu(audio, segs, enable_rms_normalize=True)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-6 * max(1.0, float(np.abs(want).max()))
    assert np.abs(want - audio).max() > 1e-3 or seconds < 1.0          # the normalisation did something
    boosted = staging.boost_low_volume(audio)
    want = staging.preprocess_audio(boosted, segs, enable_rms_normalize=True)
    got = staging.preprocess_audio_gpu(audio, segs, enable_rms_normalize=True, boost_low=True)
    assert np.abs(got - want).max() <= 1e-6 * max(1.0, float(np.abs(want).max()))
    np.testing.assert_array_equal(staging.preprocess_audio_gpu(audio, [], enable_rms_normalize=True),
                                  staging.preprocess_audio(audio, [], enable_rms_normalize=True))


# ----------------------------------------------------------------------------- the whole phase on real device statistics
def test_pipeline_flags_from_device_statistics_equal_oracle_chain(model_dirs, gpu_vad):
    """transcribe_recording on the real engine - GPU VAD probabilities, GPU staging, GPU energy scan, one ragged GPU decode whose
    per-token tsallis / margin / entropy statistics come out of the joiner epilogue - against the same chain fed by the oracle:
    oracle VAD probabilities and oracle decode_chunk (core/asr_engine.py:1209-1326) per chunk. Word texts, times, the suspect
    flags of suspect_detect (:1711-1865) and the final text must agree; `_conf` / entropy features within the rounding of the
    statistics (4 decimals)."""
    from oracle import fbank_ref, search_ref as sr, silero_ref
    from sherpa_vietnamese_asr_b200 import asr_engine, pipeline, synth
    cfg, paths, d = model_dirs("zipformer-tiny", 3)
    W, g = gpu_vad
    rec = asr_engine.create_recognizer(d, max_active_paths=4)
    orec = oracle_recognizer(paths, beam=4)[0]
    audio = _vad_audio(77, 75.0)
    thr = None

    def oracle_decode(_rec, chunks, offsets, **kw):
        out = []
        for c, o in zip(chunks, offsets):
            orec["dec_cache"].clear()
            out.append(sr.decode_chunk(orec, c, o, precomputed_features=fbank_ref.fbank(c, np.float64)))
        return out

    # the (untrained) network's probabilities sit around 0.5: pick the segments once, from the oracle, so both chains get the
    # same VAD segments and differ only in where the numbers after that come from
    from sherpa_vietnamese_asr_b200 import vad as vadmod
    segs, probs_o = vadmod.get_vad_segments(audio, silero_ref.prob_fn(W), threshold=0.5)
    _, probs_g = vadmod.get_vad_segments(audio, g.prob_fn(), threshold=0.5)
    assert np.abs(probs_o - probs_g).max() <= 1e-4
    got = pipeline.transcribe_recording(rec, audio, vad_segments=segs, rms_normalize=True)
    want = pipeline.transcribe_recording(None, audio, vad_segments=segs, rms_normalize=True, decode_chunks=oracle_decode)
    assert got["chunk_plan"] == want["chunk_plan"] and len(want["chunk_plan"]) >= 2
    assert got["text"] == want["text"] and len(want["words"]) > 20
    n_flag = 0
    for a, b in zip(got["words"], want["words"]):
        assert a["text"] == b["text"]
        assert abs(a["start"] - b["start"]) <= 1e-4 and abs(a["end"] - b["end"]) <= 1e-4
        assert a.get("_suspect_level") == b.get("_suspect_level"), (a, b)
        assert a.get("_suspect_reasons") == b.get("_suspect_reasons")
        for k in ("tsallis_max", "margin_min", "entropy_norm", "_conf"):
            if k in b:
                assert abs(a[k] - b[k]) <= 2e-4, (k, a[k], b[k])
        n_flag += bool(b.get("_suspect_level"))
    print("suspect words:", n_flag, "of", len(want["words"]))


def test_one_decode_call_spanning_several_batches(tiny):
    """A decode_streams call with more audio than one encoder pass takes (70 min) runs as chained batches - the search of batch k
    beside the encoder of batch k + 1: results equal the stream-by-stream decode."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = tiny
    clips = [synth.speech_like(16000 * 30 + 37 * i, 9700 + i) for i in range(6)]
    audios = [clips[i % 6][: len(clips[i % 6]) - 1000 * (i // 6)] for i in range(170)]       # 170 x ~30 s = 85 min, ragged
    ss = [rec.create_stream() for _ in audios]
    rec.accept_waveforms(ss, audios)
    rec.decode_streams(ss)
    st = rec.last_pipeline_stats()
    print("chained batches:", st)
    assert st["groups"] >= 2
    for i in (0, 5, 77, 139, 140, 169):
        s1 = rec.create_stream(); s1.accept_waveform(16000, audios[i]); rec.decode_stream(s1)
        assert s1.result.token_ids == ss[i].result.token_ids and s1.result.frames == ss[i].result.frames
    ntok = sum(len(s.result.token_ids) for s in ss)
    assert ntok > 1000
