"""GPU parity: every stage of the CUDA path, called through the C-ABI, against the CPU oracle.
Tolerances (BASELINE.json north_star): fbank 1e-4 abs, encoder/joiner logits 1e-3 relative, token ids exact."""
import numpy as np
import pytest

from helpers import make_graph, oracle_recognizer, rel_err, rel_l2, row_err

pytestmark = pytest.mark.gpu


def _gpu_rec(paths, **kw):
    from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer
    return OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"], joiner=paths["joiner"],
                                             tokens=paths["tokens"], **kw)


@pytest.fixture(scope="module")
def tiny(model_dirs):
    cfg, paths, d = model_dirs("zipformer-tiny", 3)
    return cfg, paths, _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)


@pytest.fixture(scope="module")
def m30(model_dirs):
    cfg, paths, d = model_dirs("zipformer-30m", 30)
    return cfg, paths, _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)


@pytest.mark.parametrize("n", [160000, 48037, 5000, 401, 160, 1])
def test_fbank_matches_oracle(tiny, n):
    from oracle import fbank_ref
    from sherpa_vietnamese_asr_b200 import synth
    _, _, rec = tiny
    a = synth.speech_like(n, 11 + n)
    got = rec.fbank(a)
    want64 = fbank_ref.fbank(a, np.float64)
    assert got.shape == want64.shape
    if got.size == 0:
        return
    # 1e-4 absolute where the bin is not sitting on the log floor / in cancellation noise
    loud = want64 > -9.0
    assert np.abs(got - want64)[loud].max() <= 1e-4 if loud.any() else True
    assert np.abs(got - want64).max() <= 5e-3


def test_fbank_ragged_batch_equals_single(tiny):
    from sherpa_vietnamese_asr_b200 import synth
    _, _, rec = tiny
    utts = [synth.speech_like(n, 70 + i) for i, n in enumerate([16000, 401, 52345, 80, 31999])]
    batch = rec.fbank_batch(utts)
    for u, b in zip(utts, batch):
        np.testing.assert_array_equal(rec.fbank(u), b)


def _encoder_taps_check(cfg, paths, rec, audio, tol):
    import torch
    from oracle import fbank_ref, zipformer_ref as zr
    orec, ocfg, tensors = oracle_recognizer(paths)
    feats = fbank_ref.fbank(audio, np.float64)
    W = zr.Weights(tensors)
    with torch.no_grad():
        want, inter = zr.encoder(W, ocfg, feats, return_intermediate=True)
    got = rec.encoder([feats])[0]
    errs, l2s, rows = {}, {}, {}
    for name, ref in list(inter.items()) + [("enc_out", want)]:
        tap = got if name == "enc_out" else rec.encoder_tap(name)
        assert tap.shape == tuple(ref.shape), name
        errs[name] = rel_err(tap, ref.numpy())
        l2s[name] = rel_l2(tap, ref.numpy())
        rows[name] = row_err(tap, ref.numpy())
    print("encoder relative errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    print("encoder rel_l2:", {k: f"{v:.2e}" for k, v in l2s.items()}, "worst row:", {k: f"{v:.2e}" for k, v in rows.items()})
    for k in errs:
        # max-error / max-value, the reference gate's rel_l2 (core/calibration.py:1057-1090), and the worst single row
        assert errs[k] <= tol and l2s[k] <= tol and rows[k] <= 10 * tol, (k, errs[k], l2s[k], rows[k])
    return feats, got, want.numpy()


def test_encoder_tiny_matches_oracle(tiny):
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = tiny
    _encoder_taps_check(cfg, paths, rec, synth.speech_like(16000 * 4 + 123, 5), 1e-4)


@pytest.mark.parametrize("n", [16000 * 10, 1600 * 7 + 5, 1600])
def test_encoder_30m_matches_oracle(m30, n):
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    _encoder_taps_check(cfg, paths, rec, synth.speech_like(n, 1234), 1e-4)


@pytest.fixture(scope="module")
def m68(model_dirs):
    """The headline model (BASELINE configs 2-5): Zipformer-68M, seed 68 as bench.py uses."""
    cfg, paths, d = model_dirs("zipformer-68m", 68)
    return cfg, paths, d, _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)


@pytest.mark.parametrize("n", [16000 * 30, 16000 * 10 + 77, 1600 * 12 + 5])
def test_encoder_68m_matches_oracle(m68, n):
    """68M-only shapes (D 384/512, 8 heads at D 512, 3-4 layers per stack, FFN 1920, the three-piece full-dim concat) against
    the oracle: taps of the embed, the six stacks and encoder_out, incl. a 30 s chunk (T = 3000, T' = 748)."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = m68
    feats, got, want = _encoder_taps_check(cfg, paths, rec, synth.speech_like(n, 6800 + n % 97), 1e-4)
    if n == 16000 * 30:
        assert feats.shape[0] == 3000 and got.shape == (748, 512)


def test_softmax_single_pass_two_pass_and_retry(model_dirs, monkeypatch):
    """Attention weights are written in one pass (shift = the row's diagonal score, unnormalised, consumers divide by the row
    sum); the exact two-pass kernel stays behind B200ASR_SOFTMAX_2PASS=1 and is what a pass is repeated with when a row sum
    leaves the safe range. All three routes against the oracle, the retry route bit-equal to the two-pass one."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d = model_dirs("zipformer-30m", 30)
    audio = synth.speech_like(16000 * 7 + 31, 4242)
    rec1 = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)
    feats, got1, want = _encoder_taps_check(cfg, paths, rec1, audio, 1e-4)
    monkeypatch.setenv("B200ASR_SOFTMAX_2PASS", "1")
    rec2 = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)
    monkeypatch.delenv("B200ASR_SOFTMAX_2PASS")
    _, got2, _ = _encoder_taps_check(cfg, paths, rec2, audio, 1e-4)
    assert rel_err(got1, got2) <= 2e-5
    rec3 = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)
    monkeypatch.setenv("B200ASR_DBG_FORCE_SOFTMAX_RETRY", "1")
    got3 = rec3.encoder([feats])[0]
    monkeypatch.delenv("B200ASR_DBG_FORCE_SOFTMAX_RETRY")
    np.testing.assert_array_equal(got3, got2)
    np.testing.assert_array_equal(rec3.encoder([feats])[0], got2)        # the recognizer stays on the exact kernel
    # a ragged batch through the decode path (the retry there repeats the whole pass)
    rec4 = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)
    utts = [synth.speech_like(n, 900 + i) for i, n in enumerate([16000 * 3, 16000 * 5 + 7, 9000])]

    def decode(rec):
        ss = []
        for a in utts:
            s = rec.create_stream()
            s.accept_waveform(16000, a)
            ss.append(s)
        rec.decode_streams(ss)
        return [s.result.token_ids for s in ss]
    want_tokens = decode(rec1)
    monkeypatch.setenv("B200ASR_DBG_FORCE_SOFTMAX_RETRY", "1")
    assert decode(rec4) == want_tokens
    monkeypatch.delenv("B200ASR_DBG_FORCE_SOFTMAX_RETRY")
    assert decode(rec2) == want_tokens


def test_c2_slice_68m_streams_token_exact(m68):
    """BASELINE config C2 on the model it names: the first 20 segments of bench.py's workload (ragged, 1-30 s) through
    create_stream / accept_waveform / decode_streams, token- and frame-exact against the oracle run segment by segment."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = m68
    orec = oracle_recognizer(paths, beam=4)[0]
    durs = synth.c2_durations(256, 256)[:20]
    audios = [synth.speech_like(int(round(x * 16000)), 256 * 100003 + i) for i, x in enumerate(durs)]
    assert _stream_case(rec, orec, audios) > 100


def test_c3_500_hotwords_68m(m68):
    """BASELINE config C3 on Zipformer-68M: 500-phrase ContextGraph with planted phrases, beam 4."""
    from oracle import search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = m68
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    audios = [synth.speech_like(n, 6900 + i) for i, n in enumerate([16000 * 9, 16000 * 4 + 321, 16000 * 14, 16000 * 2])]
    encs = _oracle_enc(orec, ocfg, audios)
    planted = []
    for e in encs:
        orec["dec_cache"].clear()
        planted.append(sr.modified_beam_search(orec, None, 4, enc_out=e)[0])
    seqs, scores = synth.random_hotwords(500, ocfg.vocab_size, 500, planted=[p for p in planted if len(p) >= 2])
    assert len(seqs) == 500
    assert _search_case(rec, orec, encs, 4, "modified_beam_search", graph_args=(seqs, scores)) > 20
    rec.set_hotwords_token_ids([], [])


def test_c4_rover_30m_68m_matches_oracle(model_dirs, m68):
    """BASELINE config C4 on the real pair (core/asr_engine.py:899-900,2041-2058): Zipformer-30M + Zipformer-68M over the same
    chunks, ROVER-combined; product path against the oracle's decode_chunk + rover_merge_words."""
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import asr_engine, synth
    chunks = [synth.speech_like(n, 1300 + i) for i, n in enumerate([16000 * 7, 16000 * 12 + 99, 16000 * 3])]
    offs = [0.0, 10.0, 30.0]
    recs, orecs = [], []
    for name, seed in (("zipformer-30m", 30), ("zipformer-68m", 68)):
        cfg, paths, d = model_dirs(name, seed)
        recs.append(asr_engine.create_recognizer(d, max_active_paths=4))
        orecs.append(oracle_recognizer(paths, beam=4)[0])
    got = asr_engine.rover_decode_chunks(recs[0], recs[1], chunks, time_offsets=offs)
    n_words = 0
    for c, off, (merged, disagree) in zip(chunks, offs, got):
        feats = fbank_ref.fbank(c, np.float64)
        want_words = []
        for o in orecs:
            o["dec_cache"].clear()
            want_words.append(sr.decode_chunk(o, c, off, precomputed_features=feats))
        want, want_dis = sr.rover_merge_words(want_words[0], want_words[1])
        assert [w["text"] for w in merged] == [w["text"] for w in want]
        assert disagree == want_dis
        for a, b in zip(merged, want):
            assert abs(a["start"] - b["start"]) <= 1e-6 and abs(a["end"] - b["end"]) <= 1e-6
            assert abs(a["prob"] - b["prob"]) <= 2e-3
        n_words += len(want)
    assert n_words > 10


def test_pipelined_groups_equal_single_pass(m68, monkeypatch):
    """A batch large enough for the pipelined decode (length-sorted groups, searches on their own streams beside the next
    group's encoder): token ids, frames and log-probs equal the utterance-by-utterance decode of the same recognizer, and the
    pass really ran as several groups."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = m68
    durs = synth.c2_durations(256, 256)[:96]
    audios = [synth.speech_like(int(round(x * 16000)), 256 * 100003 + i) for i, x in enumerate(durs)]
    ss = []
    for a in audios:
        s = rec.create_stream(); s.accept_waveform(16000, a); ss.append(s)
    rec.decode_streams(ss)
    st = rec.last_pipeline_stats()
    print("pipeline:", st)
    import os
    if os.environ.get("B200ASR_PIPELINE", "0") not in ("", "0"):      # the split is opt-in (see engine.cu plan_groups)
        assert st["groups"] >= 2
    for i in list(range(0, 96, 7)) + [95]:
        s1 = rec.create_stream(); s1.accept_waveform(16000, audios[i]); rec.decode_stream(s1)
        assert s1.result.token_ids == ss[i].result.token_ids and s1.result.frames == ss[i].result.frames
        np.testing.assert_allclose(s1.result.ys_log_probs, ss[i].result.ys_log_probs, atol=1e-5)


def test_encoder_ragged_batch_equals_single(m30):
    from oracle import fbank_ref
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    feats = [fbank_ref.fbank(synth.speech_like(n, 40 + i), np.float64) for i, n in enumerate([32000, 16000 * 5 + 77, 9000, 1000, 48123])]
    batch = rec.encoder(feats)
    for f, b in zip(feats, batch):
        single = rec.encoder([f])[0]
        assert single.shape == b.shape
        # same kernels, same per-utterance arithmetic: ragged batching must not leak across utterances
        assert rel_err(b, single) <= 2e-6


def test_decoder_joiner_rows(m30):
    import torch
    from oracle import zipformer_ref as zr
    cfg, paths, rec = m30
    orec, ocfg, tensors = oracle_recognizer(paths)
    W = zr.Weights(tensors)
    rng = np.random.default_rng(0)
    y = rng.integers(0, ocfg.vocab_size, (37, 2))
    y[0] = (0, 0)
    with torch.no_grad():
        want_dec = zr.decoder(W, ocfg, y).numpy()
        enc = rng.standard_normal((37, ocfg.joiner_dim)).astype(np.float32)
        want_lg = zr.joiner(W, torch.from_numpy(enc), torch.from_numpy(want_dec)).numpy()
    got_dec = rec.decoder(y)
    assert rel_err(got_dec, want_dec) <= 1e-5
    got_lg = rec.joiner(enc, want_dec)
    assert rel_err(got_lg, want_lg) <= 1e-5


@pytest.mark.parametrize("kb", [4, 8, 16])
def test_product_decoder_and_joiner_record_kernels(m30, kb, monkeypatch):
    """The kernels a frame step of the device search launches, pinned directly (not through downstream tokens):
    the decoder rows a step reads (the V^2 context table built by the Linear-layer GEMM; decoder_joinin_kernel with
    B200ASR_DEC_TABLE=0, test_on_demand_decoder_equals_decoder_table) -> decoder_out and X = tanh(enc + dec) against the
    oracle decoder; the tcgen05 joiner GEMM with
    the record epilogue -> log-sum-exp, top-k and the entropy sums rebuilt from the records against the oracle's logits."""
    import torch
    from oracle import zipformer_ref as zr
    cfg, paths, rec = m30
    orec, ocfg, tensors = oracle_recognizer(paths)
    W = zr.Weights(tensors)
    V = ocfg.vocab_size
    rng = np.random.default_rng(kb)
    m = 150                                              # two 128-row tiles, the second ragged
    y = rng.integers(0, V, (m, 2))
    y[0] = (0, 0)
    enc = rng.standard_normal((m, ocfg.joiner_dim)).astype(np.float32)
    with torch.no_grad():
        want_dec = zr.decoder(W, ocfg, y).numpy()
        want_lg = zr.joiner(W, torch.from_numpy(enc), torch.from_numpy(want_dec)).numpy().astype(np.float64)
    got_dec, got_x = rec.decoder_joiner_input(y, enc)
    assert row_err(got_dec, want_dec) <= 2e-5
    want_x = np.tanh(enc.astype(np.float64) + want_dec.astype(np.float64))
    assert np.abs(got_x - np.tanh(enc.astype(np.float64) + got_dec.astype(np.float64))).max() <= 1e-6    # the activation itself
    assert np.abs(got_x - want_x).max() <= 3e-5                                                          # decoder error carried through
    recs = rec.joiner_records(want_x.astype(np.float32), kb)
    P = (V + 31) // 32
    assert recs.shape == (m, P, 4 + 2 * kb)
    # the frame step's joiner kernel takes X already split into fp16 hi / lo planes (both operands straight from TMA); the
    # kernel that converts fp32 X on the fly issues the same products in the same order: identical records
    monkeypatch.setenv("B200ASR_JOINER_SS", "0")
    recs_conv = rec.joiner_records(want_x.astype(np.float32), kb)
    monkeypatch.delenv("B200ASR_JOINER_SS")
    np.testing.assert_array_equal(recs.view(np.int32), recs_conv.view(np.int32))
    pm, ps, pu, pt = (recs[:, :, i].astype(np.float64) for i in range(4))
    vals = recs[:, :, 4:4 + kb].astype(np.float64)
    cols = recs[:, :, 4 + kb:].copy().view(np.int32)
    M = pm.max(axis=1)
    want_M = want_lg.max(axis=1)
    tol = 1e-5 * np.abs(want_lg).max()           # fp32-grade product of a K = 512 contraction, relative to the logits' scale
    assert np.abs(M - want_M).max() <= tol
    lse = M + np.log((ps * np.exp(pm - M[:, None])).sum(axis=1))
    want_lse = want_M + np.log(np.exp(want_lg - want_M[:, None]).sum(axis=1))
    assert np.abs(lse - want_lse).max() <= tol
    # sum p log p from the records (search.cu select_partials) against the direct sum
    S = (ps * np.exp(pm - M[:, None])).sum(axis=1)
    dm = pm - M[:, None]
    su = (np.exp(dm) * (pu + dm * ps)).sum(axis=1)
    plogp = su / S - np.log(S)
    p_ref = np.exp(want_lg - want_lse[:, None])
    want_plogp = (p_ref * np.log(p_ref + 1e-300)).sum(axis=1)
    assert np.abs(plogp - want_plogp).max() <= 2e-4
    st = (np.exp(dm / 3.0) * pt).sum(axis=1) * S ** (-1.0 / 3.0)
    assert np.abs(st - (p_ref ** (1.0 / 3.0)).sum(axis=1)).max() <= 2e-3 * (p_ref ** (1.0 / 3.0)).sum(axis=1).max()
    # per part: the kb best (value desc, column asc) of its 32 columns
    for r in range(0, m, 7):
        for q in range(P):
            lo, hi = 32 * q, min(32 * q + 32, V)
            part = want_lg[r, lo:hi]
            order = np.argsort(-part, kind="stable")[:kb]
            n_have = len(order)
            got_c = cols[r, q, :n_have]
            np.testing.assert_allclose(vals[r, q, :n_have], part[order], atol=tol)
            if not np.array_equal(got_c, order + lo):     # a swap is legitimate only inside fp32 noise
                for a, b in zip(got_c, order + lo):
                    assert abs(want_lg[r, a] - want_lg[r, b]) <= 2 * tol
            assert (cols[r, q, n_have:] == -1).all()


def _search_case(rec, orec, enc_list, beam, method, graph_args=None):
    from oracle import search_ref as sr
    V = orec["vocab_size"]
    if graph_args is not None:
        rec.set_hotwords_token_ids(*graph_args)
        orec["context_graph"] = make_graph(*graph_args)
    else:
        rec.set_hotwords_token_ids([], [])
        orec["context_graph"] = None
    got = rec.beam_search(enc_list, method=method, beam=beam)
    n_tok = 0
    for e, (toks, frames, lps, stats) in zip(enc_list, got):
        orec["dec_cache"].clear()
        if method == "greedy_search":
            w_toks, w_frames, w_lps, _, w_emit = sr.greedy_search(orec, None, enc_out=e)
        else:
            w_toks, w_frames, w_lps, _, w_emit = sr.modified_beam_search(orec, None, beam, enc_out=e)
        assert toks == w_toks
        assert frames == w_frames
        np.testing.assert_allclose(lps, w_lps, rtol=0, atol=2e-4)
        for j, lg in enumerate(w_emit):
            st = sr.token_entropy(lg, V, rounded=False)
            np.testing.assert_allclose(stats[j], [st["tsallis_norm"], st["margin"], st["entropy_norm"], st["top1_prob"]], atol=2e-4)
        n_tok += len(toks)
    return n_tok


def _oracle_enc(orec, cfg, audios):
    import torch
    from oracle import fbank_ref, zipformer_ref as zr
    out = []
    W = orec["enc_sess"].W
    with torch.no_grad():
        for a in audios:
            out.append(zr.encoder(W, cfg, fbank_ref.fbank(a, np.float64)).numpy())
    return out


@pytest.mark.parametrize("method,beam", [("greedy_search", 1), ("modified_beam_search", 4), ("modified_beam_search", 8),
                                         ("modified_beam_search", 16)])
def test_search_from_oracle_encoder_out(m30, method, beam):
    """Search alone: same encoder_out (the oracle's) on both sides, so token ids must be identical."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    orec, ocfg, _ = oracle_recognizer(paths, beam=beam)
    audios = [synth.speech_like(n, 300 + i) for i, n in enumerate([16000 * 6, 16000 * 3 + 11, 16000 * 9, 2000])]
    encs = _oracle_enc(orec, ocfg, audios)
    n_tok = _search_case(rec, orec, encs, beam, method)
    assert n_tok > 10


def _window_cases(orec, ocfg, n_cases):
    """Many ragged encoder_out sequences cut from a few oracle encoder runs (each window is a valid input)."""
    from sherpa_vietnamese_asr_b200 import synth
    audios = [synth.speech_like(16000 * 7, 700 + i) for i in range(3)]
    base = _oracle_enc(orec, ocfg, audios)
    rng = np.random.default_rng(77)
    out = []
    for i in range(n_cases):
        e = base[i % len(base)]
        a = int(rng.integers(0, e.shape[0] - 12))
        b = int(rng.integers(a + 1, min(e.shape[0], a + 60) + 1))
        out.append(np.ascontiguousarray(e[a:b]))
    return out


def test_search_batch_spanning_several_row_tiles(m30):
    """70 utterances x beam 4 = 280 joiner rows (three 128-row tiles) with ragged lengths: the active prefix shrinks
    through tile boundaries while the per-row partial records, the recompute list and the row copies stay aligned."""
    cfg, paths, rec = m30
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    encs = _window_cases(orec, ocfg, 70)
    assert _search_case(rec, orec, encs, 4, "modified_beam_search") > 50


def test_on_demand_decoder_equals_decoder_table(m30, monkeypatch):
    """Frame steps read the decoder output of a context from the V^2 table the engine builds on its first search (the
    reference's dec_cache filled ahead of time, core/asr_engine.py:1072-1088); B200ASR_DEC_TABLE=0 keeps the on-demand path
    (decoder_joinin_kernel: decoder_proj only for changed contexts). Both are held to the oracle: identical tokens / frames,
    decoder rows within 2e-5, and the two product paths agree with each other at fp32 rounding level."""
    import torch
    from oracle import zipformer_ref as zr
    cfg, paths, rec_table = m30
    monkeypatch.setenv("B200ASR_DEC_TABLE", "0")
    rec = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4)
    orec, ocfg, tensors = oracle_recognizer(paths, beam=4)
    encs = _window_cases(orec, ocfg, 40)
    assert _search_case(rec, orec, encs, 4, "modified_beam_search") > 30      # the search issues decoder_joinin here
    assert _search_case(rec, orec, encs[:9], 1, "greedy_search") > 5
    rng = np.random.default_rng(5)
    y = rng.integers(0, ocfg.vocab_size, (150, 2))
    y[0] = (0, 0)
    enc = rng.standard_normal((150, ocfg.joiner_dim)).astype(np.float32)
    with torch.no_grad():
        want_dec = zr.decoder(zr.Weights(tensors), ocfg, y).numpy()
    dec_a, x_a = rec.decoder_joiner_input(y, enc)
    monkeypatch.delenv("B200ASR_DEC_TABLE")
    assert _search_case(rec_table, orec, encs, 4, "modified_beam_search") > 30
    dec_b, x_b = rec_table.decoder_joiner_input(y, enc)
    assert row_err(dec_a, want_dec) <= 2e-5 and row_err(dec_b, want_dec) <= 2e-5
    assert row_err(dec_a, dec_b) <= 1e-5 and np.abs(x_a - x_b).max() <= 1e-5


def test_search_cuda_core_mode_uses_full_logits_selection(model_dirs):
    """precision='fp32_simt': CUDA-core joiner GEMM + selection from the full logits rows (no partial records)."""
    cfg, paths, d = model_dirs("zipformer-30m", 30)
    rec = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4, precision="fp32_simt")
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    encs = _window_cases(orec, ocfg, 9)
    assert _search_case(rec, orec, encs, 4, "modified_beam_search") > 5
    assert _search_case(rec, orec, encs, 1, "greedy_search") > 5


def test_search_with_hotwords(m30):
    from oracle import search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    audios = [synth.speech_like(16000 * 8, 500 + i) for i in range(3)]
    encs = _oracle_enc(orec, ocfg, audios)
    planted = []
    for e in encs:
        orec["dec_cache"].clear()
        planted.append(sr.modified_beam_search(orec, None, 4, enc_out=e)[0])
    seqs, scores = synth.random_hotwords(200, ocfg.vocab_size, 500, planted=planted)
    base = [sr.modified_beam_search(orec, None, 4, enc_out=e)[0] for e in encs]
    _search_case(rec, orec, encs, 4, "modified_beam_search", graph_args=(seqs, scores))
    boosted = [sr.modified_beam_search(orec, None, 4, enc_out=e)[0] for e in encs]
    print("hotwords changed the decode:", base != boosted)
    rec.set_hotwords_token_ids([], [])


def test_search_with_500_hotwords(m30):
    """BASELINE config C3: a 500-entry ContextGraph (planted phrases so boosts fire), beam 4, ragged batch."""
    from oracle import search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    encs = _window_cases(orec, ocfg, 12)
    planted = []
    for e in encs[:6]:
        orec["dec_cache"].clear()
        planted.append(sr.modified_beam_search(orec, None, 4, enc_out=e)[0])
    seqs, scores = synth.random_hotwords(500, ocfg.vocab_size, 500, planted=[p for p in planted if len(p) >= 2])
    assert len(seqs) == 500
    assert _search_case(rec, orec, encs, 4, "modified_beam_search", graph_args=(seqs, scores)) > 10
    rec.set_hotwords_token_ids([], [])


def test_rover_two_models_matches_oracle(model_dirs):
    """BASELINE config C4: two models over the same chunks, word lists combined with ROVER; the product path
    (GPU decode -> words_from_result -> rover_merge_words) against the oracle's decode_chunk + rover_merge_words."""
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import asr_engine, synth
    chunks = [synth.speech_like(n, 1200 + i) for i, n in enumerate([16000 * 4, 16000 * 6 + 99, 16000 * 3])]
    recs, orecs = [], []
    for name, seed in (("zipformer-30m", 30), ("zipformer-tiny", 3)):
        cfg, paths, d = model_dirs(name, seed)
        recs.append(asr_engine.create_recognizer(d, max_active_paths=4))
        orecs.append(oracle_recognizer(paths, beam=4)[0])
    got = asr_engine.rover_decode_chunks(recs[0], recs[1], chunks, time_offsets=[0.0, 10.0, 20.0])
    for c, off, (merged, disagree) in zip(chunks, [0.0, 10.0, 20.0], got):
        feats = fbank_ref.fbank(c, np.float64)
        want_words = []
        for o in orecs:
            o["dec_cache"].clear()
            want_words.append(sr.decode_chunk(o, c, off, precomputed_features=feats))
        want, want_dis = sr.rover_merge_words(want_words[0], want_words[1])
        assert [w["text"] for w in merged] == [w["text"] for w in want]
        assert disagree == want_dis
        for a, b in zip(merged, want):
            assert abs(a["start"] - b["start"]) <= 1e-6 and abs(a["end"] - b["end"]) <= 1e-6
            assert abs(a["prob"] - b["prob"]) <= 2e-3


def test_context_graph_matches_oracle(m30):
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    seqs = [[5, 6, 7], [6, 7, 8], [5, 6], [9], [5, 6, 7, 8, 9], [7, 8]]
    scores = [1.5, 2.0, 1.0, 1.5, 2.5, 1.5]
    rec.set_hotwords_token_ids(seqs, scores)
    g = make_graph(seqs, scores)
    rng = np.random.default_rng(1)
    st_o, st_d = g.root, 0
    for _ in range(500):
        tok = int(rng.integers(4, 11))
        d_o, st_o = g.forward_one_step(st_o, tok)
        d_d, st_d = rec.context_forward_one_step(st_d, tok)
        assert d_o == d_d
        assert g.finalize(st_o) == rec.context_finalize(st_d)
    rec.set_hotwords_token_ids([], [])


def test_end_to_end_streams_match_oracle(m30):
    """create_stream / accept_waveform / decode_streams on a ragged batch vs the oracle run utterance by
    utterance (fbank -> encoder -> modified_beam_search, beam 4)."""
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = m30
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    audios = [synth.speech_like(n, 900 + i) for i, n in enumerate([16000 * 5, 16000 * 2 + 7, 16000 * 8 + 1234, 3000, 100])]
    streams = []
    for a in audios:
        s = rec.create_stream()
        s.accept_waveform(16000, a[: len(a) // 2])
        s.accept_waveform(16000, a[len(a) // 2:])
        streams.append(s)
    rec.decode_streams(streams)
    exact = 0
    for a, s in zip(audios, streams):
        feats = fbank_ref.fbank(a, np.float64)
        if feats.shape[0] < 9:
            assert s.result.token_ids == []
            exact += 1
            continue
        orec["dec_cache"].clear()
        toks, frames, lps, T, _ = sr.modified_beam_search(orec, feats, 4)
        assert s.result.num_frames == T
        assert s.result.token_ids == toks
        assert s.result.frames == frames
        np.testing.assert_allclose(s.result.ys_log_probs, lps, atol=5e-3)
        exact += 1
    assert exact == len(audios)


def test_c1_greedy_10s_clip_end_to_end(model_dirs):
    """BASELINE config C1: Zipformer-30M, greedy_search, one synthetic 10 s 16 kHz clip (seed 1234) through
    create_stream / accept_waveform / decode_stream, against the oracle's fbank -> encoder -> greedy search."""
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d = model_dirs("zipformer-30m", 30)
    rec = _gpu_rec(paths, decoding_method="greedy_search")
    orec, ocfg, _ = oracle_recognizer(paths, beam=1)
    a = synth.speech_like(160000, 1234)
    s = rec.create_stream()
    s.accept_waveform(16000, a)
    rec.decode_stream(s)
    feats = fbank_ref.fbank(a, np.float64)
    assert feats.shape == (1000, 80)
    toks, frames, lps, T, _ = sr.greedy_search(orec, feats)
    assert T == 248 and s.result.num_frames == 248
    assert s.result.token_ids == toks and len(toks) > 5
    assert s.result.frames == frames
    np.testing.assert_allclose(s.result.ys_log_probs, lps, atol=5e-3)
    assert abs(s.result.timestamps[0] - frames[0] / 248 * 10.0) < 1e-4


# ----------------------------------------------------------------------------- GEMM kernels
GEMM_SHAPES = [(1000, 272, 192), (517, 48, 192), (4096, 640, 192), (130, 2000, 512), (777, 192, 2432), (64, 16, 48),
               (2500, 384, 128), (3000, 512, 512), (129, 130, 36)]


@pytest.mark.parametrize("impl", ["fp32", "tc", "tc3", "f16x3", "bf16"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_kernels_match_numpy(tiny, impl, M, N, K):
    """Every GEMM kernel against a float64 product, relative to the output scale. FP32 CUDA-core kernel: fp32 rounding only;
    tcgen05 TF32 (10-bit mantissa) 1.5e-3; the two fp32-grade operand splits - 3xTF32 and fp16 hi + lo (f16x3) - 2e-5;
    BF16 operands (8-bit mantissa) 1e-2. All accumulate in FP32."""
    _, _, rec = tiny
    rng = np.random.default_rng(M * 7 + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    R = rng.standard_normal((M, N)).astype(np.float32)
    for act in (0, 1, 2):
        got, _ = rec.gemm(A, W, b, R, act=act, impl=impl)
        z = A.astype(np.float64) @ W.astype(np.float64).T + b
        if act == 1:
            z = np.logaddexp(0, z - 4.0) - 0.08 * z - 0.035
        elif act == 2:
            z = np.logaddexp(0, z - 1.0) - 0.08 * z - 0.313261687
        want = z + R
        tol = {"fp32": 5e-6, "tc": 1.5e-3, "tc3": 2e-5, "f16x3": 2e-5, "bf16": 1e-2}[impl]
        assert rel_err(got, want) <= tol, (impl, act, rel_err(got, want))
    got, _ = rec.gemm(A, W, None, None, act=0, impl=impl)
    assert rel_err(got, A.astype(np.float64) @ W.astype(np.float64).T) <= {"fp32": 5e-6, "tc": 1.5e-3, "tc3": 2e-5, "f16x3": 2e-5, "bf16": 1e-2}[impl]


def test_f16x3_gemm_keeps_small_and_large_operands(tiny):
    """The fp16 operand split scales activations by 64 and weights by 1024 so the low parts stay normal fp16 numbers: rows of
    very small and of large activations (the range a BiasNorm'd residual stream and a Swoosh hidden layer span) keep fp32-grade
    accuracy relative to their own scale."""
    _, _, rec = tiny
    rng = np.random.default_rng(3)
    M, N, K = 512, 256, 384
    A = rng.standard_normal((M, K)).astype(np.float32)
    A *= np.exp(rng.uniform(np.log(1e-4), np.log(60.0), (M, 1))).astype(np.float32)       # per-row scale 1e-4 .. 60
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    W[::7] *= np.float32(1e-3)
    got, _ = rec.gemm(A, W, None, None, act=0, impl="f16x3")
    want = A.astype(np.float64) @ W.astype(np.float64).T
    assert row_err(got, want) <= 2e-5


def test_tensor_core_mode_end_to_end(model_dirs):
    """precision='tf32' = single-pass TF32 operands on tcgen05 (not the BF16 mode, see test_bf16_mode_*): encoder within
    2e-2 relative of the oracle, token edit distance vs the FP32 oracle decode small."""
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d = model_dirs("zipformer-30m", 30)
    rec = _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4, precision="tf32")
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    audios = [synth.speech_like(n, 900 + i) for i, n in enumerate([16000 * 5, 16000 * 8 + 1234, 16000 * 3])]
    import torch
    from oracle import zipformer_ref as zr
    feats = [fbank_ref.fbank(a, np.float64) for a in audios]
    encs = rec.encoder(feats)
    tot, dist = 0, 0
    for f, e in zip(feats, encs):
        with torch.no_grad():
            want = zr.encoder(orec["enc_sess"].W, ocfg, f).numpy()
        err = rel_err(e, want)
        print("tensor-core encoder rel err", err)
        assert err <= 2e-2
    ss = []
    for a in audios:
        s = rec.create_stream(); s.accept_waveform(16000, a); ss.append(s)
    rec.decode_streams(ss)
    import difflib
    for f, s in zip(feats, ss):
        orec["dec_cache"].clear()
        toks = sr.modified_beam_search(orec, f, 4)[0]
        sm = difflib.SequenceMatcher(None, toks, s.result.token_ids, autojunk=False)
        dist += sum(max(i2 - i1, j2 - j1) for tag, i1, i2, j1, j2 in sm.get_opcodes() if tag != "equal")
        tot += len(toks)
    print("tensor-core mode token edit distance:", dist, "/", tot)
    assert dist <= max(2, 0.05 * tot)


def _edit_distance(a, b):
    import difflib
    sm = difflib.SequenceMatcher(None, a, b, autojunk=False)
    return sum(max(i2 - i1, j2 - j1) for tag, i1, i2, j1, j2 in sm.get_opcodes() if tag != "equal")


def test_bf16_mode_68m(m68):
    """The BF16 mode (bf16 weights, activations rounded to bf16 at every Linear, one kind::f16 MMA per K step, FP32 accumulate,
    fp32 residual stream) on the headline model. What can be held to a bar with random-init weights is the arithmetic: the
    encoder output stays within bf16's error budget of the oracle (8-bit mantissa through ~100 chained Linears). The token
    edit distance against the FP32-mode decode over > 2000 tokens is measured and printed, not asserted: an untrained network
    has no decision margins - the single-pass TF32 mode (10-bit mantissa) measured beside it flips tokens too - so the north
    star's 0.5 % budget, which presumes a trained checkpoint, cannot be checked offline."""
    import torch
    from oracle import fbank_ref, zipformer_ref as zr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, d, rec = m68
    orec, ocfg, _ = oracle_recognizer(paths, beam=4)
    modes = {"bf16": _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4, precision="bf16"),
             "tf32": _gpu_rec(paths, decoding_method="modified_beam_search", max_active_paths=4, precision="tf32")}
    feats = fbank_ref.fbank(synth.speech_like(16000 * 12, 6868), np.float64)
    with torch.no_grad():
        want = zr.encoder(orec["enc_sess"].W, ocfg, feats).numpy()
    errs = {k: rel_l2(r.encoder([feats])[0], want) for k, r in modes.items()}
    errs["fp32"] = rel_l2(rec.encoder([feats])[0], want)
    print("encoder rel_l2 by mode:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["fp32"] <= 1e-4 and errs["tf32"] <= 2e-2 and errs["bf16"] <= 1e-1
    assert errs["fp32"] < errs["tf32"] < errs["bf16"]
    durs = synth.c2_durations(256, 256)[:40]
    audios = [synth.speech_like(int(round(x * 16000)), 256 * 100003 + i) for i, x in enumerate(durs)]
    outs = {}
    for name, r in [("fp32", rec)] + list(modes.items()):
        ss = [r.create_stream() for _ in audios]
        r.accept_waveforms(ss, audios)
        r.decode_streams(ss)
        outs[name] = [list(s.result.token_ids) for s in ss]
    tot = sum(len(t) for t in outs["fp32"])
    assert tot >= 2000
    for name in ("tf32", "bf16"):
        dist = sum(_edit_distance(a, b) for a, b in zip(outs["fp32"], outs[name]))
        print(f"{name} mode token edit distance vs FP32 mode: {dist} / {tot} = {100.0 * dist / tot:.2f} % (random-init weights)")


def test_transcribe_long_matches_oracle_chunks(model_dirs):
    """a long recording through the chunk planner -> ONE ragged GPU batch -> overlap stitcher.
    Per-chunk word lists equal the oracle's decode of the same chunks; the stitched result is the stitcher applied to them."""
    import copy
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import asr_engine, chunking, synth
    cfg, paths, d = model_dirs("zipformer-tiny", 3)
    rec = asr_engine.create_recognizer(d, max_active_paths=4)
    orec = oracle_recognizer(paths, beam=4)[0]
    audio = synth.speech_like(16000 * 41 + 321, 4100)
    for a, b in [(9.6, 10.3), (19.0, 19.5), (31.2, 32.0)]:           # silences for the planner to snap to
        audio[int(a * 16000):int(b * 16000)] *= 0.01
    vad = [(8000, 16000 * 25), (16000 * 26, len(audio) - 4000)]
    res = chunking.transcribe_long(rec, audio, vad, segment_samples=16000 * 10)
    speech, omap = chunking.concat_vad_speech(audio, vad)
    assert len(res["chunk_plan"]) >= 4 and res["chunk_plan"][-1][1] == len(speech)
    tm = chunking.ConcatTimeMap(omap)
    for (s, e, o), got in zip(res["chunk_plan"], res["chunk_results"]):
        orec["dec_cache"].clear()
        chunk = speech[s:e]
        want = sr.decode_chunk(orec, chunk, s / 16000.0, precomputed_features=fbank_ref.fbank(chunk, np.float64))
        assert [w["text"] for w in got["words"]] == [w["text"] for w in want]
        for a, b in zip(got["words"], want):
            assert abs(a["start"] - tm(b["start"])) <= 1e-4 and abs(a["local_start"] - b["local_start"]) <= 1e-6
            assert abs(a["prob"] - b["prob"]) <= 2e-3
    words, text = chunking.merge_chunks_with_overlap(copy.deepcopy(res["chunk_results"]))
    assert text == res["text"] and len(words) == len(res["words"]) > 0


def test_energy_scan_flags_equal_numpy():
    """csrc/energy.cu through B200AsrSilentFrames: per-frame quiet flags equal NumPy's float32 flags, including frames whose
    RMS sits on the threshold; regions equal the host find_silent_regions (core/asr_engine.py:521-553)."""
    import ctypes as C
    from oracle import chunk_cases as cc
    from sherpa_vietnamese_asr_b200 import _capi, chunking
    rng = np.random.default_rng(12)
    # RMS log-uniform around the 0.01 threshold, many frames within a few ulp-scale steps of it
    n_frames = 200_003
    scale = np.exp(rng.uniform(np.log(0.003), np.log(0.03), n_frames)).astype(np.float32)
    scale[::7] = np.float32(0.01) * (1 + rng.uniform(-3e-7, 3e-7, len(scale[::7]))).astype(np.float32)
    x = (rng.normal(0, 1, (n_frames, 160)).astype(np.float32))
    x /= np.sqrt(np.mean(x.astype(np.float64) ** 2, axis=1, keepdims=True)).astype(np.float32)
    x = np.ascontiguousarray((x * scale[:, None]).reshape(-1))
    x = np.concatenate([x, rng.normal(0, 0.1, 77).astype(np.float32)])        # a partial last frame is ignored
    want = np.sqrt(np.mean(x[:n_frames * 160].reshape(n_frames, 160) ** 2, axis=1)) < 0.01
    assert 0.2 < want.mean() < 0.8
    quiet = np.full(n_frames, 9, dtype=np.uint8)
    lib = _capi.lib()
    assert lib.B200AsrSilentFrames(_capi.fptr(x), len(x), 16000, 0.01, None, 0) == n_frames
    rc = lib.B200AsrSilentFrames(_capi.fptr(x), len(x), 16000, float(np.float32(0.01)), quiet.ctypes.data_as(C.POINTER(C.c_uint8)), 0)
    assert rc == n_frames, _capi.last_error()
    assert np.array_equal(quiet.astype(bool), want)
    assert lib.B200AsrSilentFrames(_capi.fptr(x), len(x), 8000, 0.01, quiet.ctypes.data_as(C.POINTER(C.c_uint8)), 0) < 0
    for seed, sec in [(1, 0.005), (3, 7.3), (5, 125.7), (6, 400.0)]:
        audio = cc.silence_audio(seed, sec)
        assert chunking.find_silent_regions_gpu(audio) == chunking.find_silent_regions(audio)
        assert chunking.find_silent_regions_gpu(audio, 16000, 0.02, 0.1) == chunking.find_silent_regions(audio, 16000, 0.02, 0.1)


def _stream_case(rec, orec, audios, beam=4):
    from oracle import fbank_ref, search_ref as sr
    streams = []
    for a in audios:
        s = rec.create_stream()
        if len(a):
            s.accept_waveform(16000, a)
        streams.append(s)
    rec.decode_streams(streams)
    n_tok = 0
    for a, s in zip(audios, streams):
        feats = fbank_ref.fbank(a, np.float64) if len(a) else np.zeros((0, 80))
        if feats.shape[0] < 9:
            assert s.result.token_ids == []
            continue
        orec["dec_cache"].clear()
        toks, frames, lps, T, _ = sr.modified_beam_search(orec, feats, beam)
        assert s.result.num_frames == T
        assert s.result.token_ids == toks and s.result.frames == frames
        np.testing.assert_allclose(s.result.ys_log_probs, lps, atol=5e-3)
        n_tok += len(toks)
    return n_tok


def test_long_and_degenerate_utterances(tiny, m30):
    """Maximum sizes and empties: 35 s (the longest chunk the planner makes) and 61 s utterances, no samples at all, less
    than one frame, one sample - in one ragged batch; and a 35 s chunk through the 30M model."""
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = tiny
    orec = oracle_recognizer(paths, beam=4)[0]
    sizes = [16000 * 35, 16000 * 61 + 77, 0, 79, 1, 16000 * 2]
    audios = [synth.speech_like(n, 7000 + i) if n else np.zeros(0, np.float32) for i, n in enumerate(sizes)]
    assert _stream_case(rec, orec, audios) > 50
    cfg, paths, rec = m30
    orec = oracle_recognizer(paths, beam=4)[0]
    assert _stream_case(rec, orec, [synth.speech_like(16000 * 35 + 123, 7100)]) > 5


def test_batch_of_300_short_streams(tiny):
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = tiny
    orec = oracle_recognizer(paths, beam=4)[0]
    rng = np.random.default_rng(5)
    audios = [synth.speech_like(int(n), 7200 + i) for i, n in enumerate(rng.integers(1600, 24000, 300))]
    assert _stream_case(rec, orec, audios) > 300


def test_regrown_stream_is_decoded_again(tiny):
    """decode_stream is re-callable on a stream that received more audio (streaming_asr.py:408) and idempotent on an
    unchanged one."""
    from oracle import fbank_ref, search_ref as sr
    from sherpa_vietnamese_asr_b200 import synth
    cfg, paths, rec = tiny
    orec = oracle_recognizer(paths, beam=4)[0]
    a = synth.speech_like(16000 * 5, 7300)
    s = rec.create_stream()
    s.accept_waveform(16000, a[:32000])
    rec.decode_stream(s)
    first = list(s.result.token_ids)
    s.accept_waveform(16000, a[32000:])
    rec.decode_stream(s)
    second = list(s.result.token_ids)
    rec.decode_stream(s)
    for part, got in ((a[:32000], first), (a, second), (a, list(s.result.token_ids))):
        orec["dec_cache"].clear()
        assert got == sr.modified_beam_search(orec, fbank_ref.fbank(part, np.float64), 4)[0]
    assert len(second) > len(first) > 0


def test_threads_share_recognizers(tiny, m30):
    """Threading contract: recognizers shared by several Python threads, each with its own streams;
    results equal the single-threaded decode (which the tests above pin to the oracle)."""
    import threading
    from sherpa_vietnamese_asr_b200 import synth
    recs = [tiny[2], m30[2]]
    rng = np.random.default_rng(1)
    n_threads, n_iter, n_utt = 4, 4, 5

    def decode(rec, audios, split):
        ss = []
        for a in audios:
            s = rec.create_stream()
            if split:
                s.accept_waveform(16000, a[: len(a) // 3])
                s.accept_waveform(16000, a[len(a) // 3:])
            else:
                s.accept_waveform(16000, a)
            ss.append(s)
        rec.decode_streams(ss)
        return [(list(s.result.token_ids), list(s.result.frames)) for s in ss]

    audio = {(t, i): [synth.speech_like(int(rng.integers(8000, 90000)), 8000 + 100 * t + 10 * i + u) for u in range(n_utt)]
             for t in range(n_threads) for i in range(n_iter)}
    want = {k: decode(recs[k[0] % 2], v, False) for k, v in audio.items()}
    got, errors = {}, []

    def worker(t):
        try:
            for i in range(n_iter):
                got[(t, i)] = decode(recs[t % 2], audio[(t, i)], bool(i % 2))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not errors and got == want
    assert sum(len(tk) for v in want.values() for tk, _ in v) > 100
