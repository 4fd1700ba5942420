"""CPU restatements of the index/merge arithmetic the CUDA kernels rely on (no GPU needed).

* the per-32-column partial records of the joiner epilogue (csrc/gemm_tc.cu, ACT_JOINER) merged the way
  select_partials_kernel (csrc/search.cu) merges them, against the direct `_compute_token_entropy` restatement
  (oracle/search_ref.py::token_entropy, pinned to /root/reference core/asr_engine.py:1159-1181) and the float32
  log-softmax of the beam search (:1096-1098);
* the 256-point FFT decomposition of csrc/fbank.cu (radix 8 x 8 x 4 over register index, exchange buffer, lane quad).
"""
import numpy as np

from oracle import search_ref as sr


def _records(logits_row, kb=4):
    """What one epilogue thread emits per 32-column part: max, sum e^(x-m), sum e^(x-m)(x-m), sum e^((x-m)/3), top-kb."""
    V = logits_row.shape[0]
    recs = []
    for p0 in range(0, V, 32):
        x = logits_row[p0:p0 + 32].astype(np.float32)
        m = x.max()
        dx = x - m
        e = np.exp(dx)
        order = np.lexsort((np.arange(len(x)), -x))[:kb]          # value desc, column asc
        recs.append((m, e.sum(dtype=np.float32), (e * dx).sum(dtype=np.float32), np.exp(dx / 3).sum(dtype=np.float32),
                     x[order], order + p0))
    return recs


def _merge(recs, V):
    """select_partials_kernel phase (a): row max, log-sum-exp and the token statistics from the parts."""
    ms = np.array([r[0] for r in recs], np.float32)
    M = ms.max()
    sc = np.exp(ms - M)
    S = float(np.sum(np.array([r[1] for r in recs]) * sc))
    U = float(np.sum(sc * (np.array([r[2] for r in recs]) + (ms - M) * np.array([r[1] for r in recs]))))
    T = float(np.sum(np.exp((ms - M) / 3) * np.array([r[3] for r in recs])))
    lse = np.log(S)
    vals = np.concatenate([r[4] for r in recs])
    top = np.sort(vals)[::-1]
    p1, p2 = np.exp(top[0] - M) / S, np.exp(top[1] - M) / S
    a = 1.0 / 3.0
    ts_max = (1.0 / (a - 1.0)) * (1.0 - V ** (1.0 - a))
    tsallis = (1.0 / (a - 1.0)) * (1.0 - T * np.exp(-lse / 3))
    return {"M": M, "lse": lse, "tsallis_norm": tsallis / ts_max, "margin": p1 - p2,
            "entropy_norm": -(U / S - lse) / np.log(V), "top1_prob": p1}


def test_partial_records_reproduce_logsoftmax_topk_and_token_statistics():
    rng = np.random.default_rng(5)
    V = 2000
    for trial in range(20):
        logits = (rng.standard_normal(V) * rng.uniform(0.5, 6.0)).astype(np.float32)
        if trial % 3 == 0:
            logits[0] += 8.0                                       # a dominant blank
        recs = _records(logits)
        got = _merge(recs, V)
        # log-softmax as the reference computes it (float32)
        lp = sr.log_softmax_f32(logits[None])[0]
        mine = (logits - got["M"]) - np.float32(got["lse"])
        np.testing.assert_allclose(mine, lp, atol=2e-6)
        # exact top-4 (value desc, index asc) is contained in the union of the parts' top-4
        want = np.lexsort((np.arange(V), -logits))[:4]
        cand = np.concatenate([r[5] for r in recs])
        cv = logits[cand]
        pick = cand[np.lexsort((cand, -cv))[:4]]
        assert list(pick) == list(want)
        st = sr.token_entropy(logits, V, rounded=False)
        for k in ("tsallis_norm", "margin", "entropy_norm", "top1_prob"):
            assert abs(got[k] - st[k]) <= 1e-5, (k, got[k], st[k])


def test_fbank_fft_decomposition_matches_numpy():
    """256 = 8 x 8 x 4 exactly as the kernel walks it: lane = n2, registers n1; exchange [k1][36]; lane' = (k1, b);
    radix-4 across the quad with two xor butterflies; output k = k1 + 8 c + 64 d with d = {0,2,1,3}[b]."""
    rng = np.random.default_rng(0)
    z = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    W = lambda N, m: np.exp(-2j * np.pi * m / N)

    def dft4(p):
        u0, u1, v0, v1 = p[0] + p[2], p[0] - p[2], p[1] + p[3], (p[1] - p[3]) * (-1j)
        return [u0 + v0, u1 + v1, u0 - v0, u1 - v1]

    def dft8(x):
        a = [x[j] + x[j + 4] for j in range(4)]
        b = [x[j] - x[j + 4] for j in range(4)]
        b[1] *= W(8, 1); b[2] *= W(8, 2); b[3] *= W(8, 3)
        e, o = dft4(a), dft4(b)
        y = [0] * 8
        for i in range(4):
            y[2 * i], y[2 * i + 1] = e[i], o[i]
        return y

    sm = np.zeros(8 * 36, complex)
    for lane in range(32):
        y = dft8([z[32 * n1 + lane] for n1 in range(8)])
        for k1 in range(8):
            sm[k1 * 36 + lane] = y[k1] * W(256, lane * k1)
    regs = np.zeros((32, 8), complex)
    for lp in range(32):
        k1, b = lp >> 2, lp & 3
        y = dft8([sm[k1 * 36 + 4 * a + b] for a in range(8)])
        for c in range(8):
            regs[lp, c] = y[c] * W(32, b * c)
    X = np.zeros(256, complex)
    lanes = np.arange(32)
    for c in range(8):
        v = regs[:, c]
        p = v[lanes ^ 2]
        u = np.where((lanes & 2) == 0, v + p, p - v)
        u = np.where((lanes & 3) == 3, u * (-1j), u)
        q = u[lanes ^ 1]
        o = np.where((lanes & 1) == 0, u + q, q - u)
        for lp in range(32):
            k1, b = lp >> 2, lp & 3
            d = ((b & 1) << 1) | (b >> 1)
            X[k1 + 8 * c + 64 * d] = o[lp]
    np.testing.assert_allclose(X, np.fft.fft(z), atol=1e-10)
    # bank check of the exchange buffer: stores [k1*36 + lane] and loads [k1*36 + 4a + b] hit 32 distinct banks
    for a in range(8):
        assert len({((lp >> 2) * 36 + 4 * a + (lp & 3)) % 32 for lp in range(32)}) == 32
    # natural-order spectrum with 8 floats of padding per 64: the stores of one register index are conflict-free
    for c in range(8):
        addr = []
        for lp in range(32):
            k1, b = lp >> 2, lp & 3
            k = k1 + 8 * c + 64 * (((b & 1) << 1) | (b >> 1))
            addr.append((k + 8 * (k >> 6)) % 32)
        assert len(set(addr)) == 32


# ----------------------------------------------------------------------------- energy scan (csrc/energy.cu)
def _energy_kernel_order(frames):
    """The kernel's arithmetic, lane for lane, in NumPy float32: two halves of 80 squares, 8 interleaved accumulators
    per half added in index order, xor-tree fold (1, 2, 4), half 0 + half 1, / 160, sqrt."""
    sq = (frames * frames).astype(np.float32)
    halves = []
    for h in range(2):
        a = sq[:, h * 80:(h + 1) * 80].reshape(-1, 10, 8)
        r = a[:, 0, :].copy()
        for i in range(1, 10):
            r = (r + a[:, i, :]).astype(np.float32)
        for step in (1, 2, 4):                       # lane j receives r[j] + r[j ^ step]
            r = (r + r[:, np.arange(8) ^ step]).astype(np.float32)
        halves.append(r[:, 0])
    tot = (halves[0] + halves[1]).astype(np.float32)
    return np.sqrt((tot / np.float32(160)).astype(np.float32)).astype(np.float32)


def test_energy_scan_order_reproduces_numpy_float32_bits():
    """The flags decide where chunks are cut, so the kernel restates NumPy's pairwise float32 summation exactly; this
    pins that statement against NumPy itself (the reference computes np.sqrt(np.mean(frames ** 2, axis=1)),
    core/asr_engine.py:532-536)."""
    rng = np.random.default_rng(0)
    for scale in (1e-20, 1e-3, 0.01, 0.3, 1.0):
        x = rng.normal(0, scale, (50000, 160)).astype(np.float32)
        got = _energy_kernel_order(x)
        want = np.sqrt(np.mean(x ** 2, axis=1))
        assert want.dtype == np.float32
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        per_frame = np.array([np.sqrt(np.mean(f ** 2)) for f in x[:500]])
        assert np.array_equal(per_frame.view(np.uint32), got[:500].view(np.uint32))
        # and the comparison the reference makes (weak Python float against float32)
        assert np.array_equal(got < np.float32(0.01), want < 0.01)


# ----------------------------------------------------------------------------- fp16-hi / bf16-lo operand split (csrc/gemm_tc_f16.cu)
def _bf16_rn(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def _split16(x):
    """The converter's arithmetic (split_pair / split_w16_kernel): hi = fp16(clamp(x, +-65504)), lo = bf16(x - hi)."""
    hi = np.clip(x, -65504.0, 65504.0).astype(np.float16).astype(np.float32)
    return hi, _bf16_rn((x - hi).astype(np.float32))


def test_f16_split_product_is_fp32_grade():
    """A*W ~ A_lo*W_hi + A_hi*W_lo + A_hi*W_hi with fp32 accumulation: relative error of the split itself (products in
    float64 here, so only the operand rounding shows) stays at the 3xTF32 level at the network's operand scales; a whole
    operand far under the fp16 normal range (6.1e-5) or over 65504 degrades smoothly towards bf16 precision - never to
    zeros or infinities."""
    rng = np.random.default_rng(0)
    for scale_a, scale_w, tol in ((1.0, 0.05, 2e-6), (0.01, 1e-3, 2e-6), (6.9, 1.0, 2e-6), (1e-6, 1e-3, 1e-4), (3e5, 1.0, 2e-2)):
        a = (rng.normal(0, scale_a, (64, 256))).astype(np.float32)
        w = (rng.normal(0, scale_w, (48, 256))).astype(np.float32)
        ah, al = _split16(a)
        wh, wl = _split16(w)
        f = np.float64
        got = al.astype(f) @ wh.astype(f).T + ah.astype(f) @ wl.astype(f).T + ah.astype(f) @ wh.astype(f).T
        want = a.astype(f) @ w.astype(f).T
        assert np.isfinite(got).all()
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < tol, (scale_a, err)
    # one operand: |x - (hi + lo)| <= max(2^-19 |x|, 2^-33) over 12 decades, overflow side excluded
    x = (rng.normal(0, 1, 200000) * 10.0 ** rng.uniform(-9, 3, 200000)).astype(np.float32)
    hi, lo = _split16(x)
    assert np.all(np.abs(x - (hi + lo)) <= np.maximum(2.0 ** -19 * np.abs(x), 2.0 ** -33))
