"""CPU study for a next-round GEMM mode: can the FP32 (token-exact) mode run on two-term FP16 operands?

Today's FP32 mode is error-compensated 3xTF32: x = hi + lo (hi = 13 low mantissa bits cleared), three TF32 MMAs per K step.
The candidate: hi = fp16(x), lo' = fp16((x - hi) * 2^11); A*W ~ A_hi*W_hi + 2^-11 (A_lo'*W_hi + A_hi*W_lo') with FP32
accumulation - also three MMAs, but kind::f16 runs at twice the TF32 rate and the operands are half the bytes (the GEMMs are
bound by operand traffic). FP16's range is the risk: hi overflows above 65504 and loses bits below 6e-5.

This script runs the oracle encoder on synthetic audio with the bench model's random-init weights, records every Linear's
operands, and compares the schemes against a float64 product. No GPU needed.

    python tools/fp16_split_study.py [zipformer-30m|zipformer-68m] [seconds]
    python tools/fp16_split_study.py --tokens [model]      # decoded token ids under each scheme vs fp32
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import fbank_ref, zipformer_ref as zr  # noqa: E402
from sherpa_vietnamese_asr_b200 import synth, weights  # noqa: E402


def tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def gemm_3xtf32(a, w):
    ah, wh = tf32_trunc(a), tf32_trunc(w)
    al, wl = tf32_trunc(a - ah), tf32_trunc(w - wh)          # the tensor core truncates lo as well
    f = np.float64
    return (al.astype(f) @ wh.astype(f).T + ah.astype(f) @ wl.astype(f).T + ah.astype(f) @ wh.astype(f).T).astype(np.float32)


def gemm_2xfp16(a, w):
    ah, wh = a.astype(np.float16), w.astype(np.float16)
    al = ((a - ah.astype(np.float32)) * 2048.0).astype(np.float16)
    wl = ((w - wh.astype(np.float32)) * 2048.0).astype(np.float16)
    f = np.float64
    main = ah.astype(f) @ wh.astype(f).T
    cross = al.astype(f) @ wh.astype(f).T + ah.astype(f) @ wl.astype(f).T
    return (main + cross / 2048.0).astype(np.float32), bool(np.isinf(ah).any() or np.isinf(wh).any())


def bf16_round(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000          # round to nearest even
    return u.astype(np.uint32).view(np.float32)


def gemm_f16_bf16lo(a, w):
    """hi in fp16, lo (unscaled) in bf16: one FP32 accumulator, kind::f16 MMAs with mixed a/b formats."""
    ah, wh = a.astype(np.float16).astype(np.float32), w.astype(np.float16).astype(np.float32)
    al, wl = bf16_round(a - ah), bf16_round(w - wh)
    f = np.float64
    return (al.astype(f) @ wh.astype(f).T + ah.astype(f) @ wl.astype(f).T + ah.astype(f) @ wh.astype(f).T).astype(np.float32)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "zipformer-30m"
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    cfg = weights.CONFIGS[name]()
    d = tempfile.mkdtemp()
    paths = weights.write_model_dir(d, cfg, 68 if "68" in name else 30)
    tensors = {}
    for part in ("encoder", "decoder", "joiner"):
        tensors.update(weights.load_container(paths[part])[1])
    W = zr.Weights(tensors)
    feats = fbank_ref.fbank(synth.speech_like(int(16000 * secs), 7), np.float64)
    calls = []
    orig = F.linear

    def rec(x, w, b=None):
        if x.dim() == 2 and x.shape[0] >= 8:
            calls.append((x.detach().numpy().astype(np.float32), w.detach().numpy().astype(np.float32)))
        return orig(x, w, b)

    F.linear = rec
    try:
        with torch.no_grad():
            zr.encoder(W, cfg, feats)
    finally:
        F.linear = orig
    worst = {"fp32": 0.0, "3xtf32": 0.0, "2xfp16": 0.0, "f16+bf16lo": 0.0}
    amax, amin_nz, overflow = 0.0, 1e30, False
    rng = np.random.default_rng(0)
    for a, w in calls:
        if a.shape[0] > 256:
            a = a[rng.choice(a.shape[0], 256, replace=False)]
        ref = a.astype(np.float64) @ w.astype(np.float64).T
        scale = np.abs(ref).max() + 1e-30
        worst["fp32"] = max(worst["fp32"], float(np.abs((a @ w.T) - ref).max() / scale))
        worst["3xtf32"] = max(worst["3xtf32"], float(np.abs(gemm_3xtf32(a, w) - ref).max() / scale))
        g, ovf = gemm_2xfp16(a, w)
        overflow |= ovf
        worst["2xfp16"] = max(worst["2xfp16"], float(np.abs(g - ref).max() / scale))
        worst["f16+bf16lo"] = max(worst["f16+bf16lo"], float(np.abs(gemm_f16_bf16lo(np.ascontiguousarray(a), np.ascontiguousarray(w)) - ref).max() / scale))
        amax = max(amax, float(np.abs(a).max()), float(np.abs(w).max()))
        nz = np.abs(a[a != 0])
        if nz.size:
            amin_nz = min(amin_nz, float(np.percentile(nz, 1)))
    print(f"{name}: {len(calls)} Linear calls on {secs:.0f} s of audio")
    print(f"  largest |operand| = {amax:.3g} (fp16 max 65504; overflow seen: {overflow}); 1st percentile of |activation| = {amin_nz:.3g}")
    for k, v in worst.items():
        print(f"  worst max-error / output scale, {k:10s}: {v:.3e}")




def token_study(name="zipformer-30m", secs=(5.0, 3.0, 7.0), beam=4):
    """End to end on the CPU oracle: every Linear (encoder, decoder_proj, joiner) computed with an emulated operand scheme,
    modified_beam_search on top, token ids compared with the plain fp32 run."""
    from oracle import search_ref as sr
    cfg = weights.CONFIGS[name]()
    d = tempfile.mkdtemp()
    paths = weights.write_model_dir(d, cfg, 68 if "68" in name else 30)
    tensors = {}
    for part in ("encoder", "decoder", "joiner"):
        tensors.update(weights.load_container(paths[part])[1])
    audios = [synth.speech_like(int(16000 * s), 40 + i) for i, s in enumerate(secs)]
    feats = [fbank_ref.fbank(a, np.float64) for a in audios]
    orig = F.linear

    def tf32_1(a, w):
        return (tf32_trunc(a).astype(np.float64) @ tf32_trunc(w).astype(np.float64).T).astype(np.float32)

    schemes = {"fp32": None, "tf32 x1": tf32_1, "3xtf32": gemm_3xtf32, "f16+bf16lo": gemm_f16_bf16lo,
               "2xfp16": lambda a, w: gemm_2xfp16(a, w)[0]}
    results = {}
    for sname, fn in schemes.items():
        def lin(x, w, b=None, fn=fn):
            if fn is None or x.dim() != 2:
                return orig(x, w, b)
            y = torch.from_numpy(fn(np.ascontiguousarray(x.detach().numpy(), dtype=np.float32),
                                    np.ascontiguousarray(w.detach().numpy(), dtype=np.float32)))
            return y + b if b is not None else y
        F.linear = lin
        try:
            rec = zr.make_recognizer(tensors, cfg, max_active_paths=beam)
            toks = []
            with torch.no_grad():
                for f in feats:
                    rec["dec_cache"].clear()
                    toks.append(sr.modified_beam_search(rec, f, beam)[0])
        finally:
            F.linear = orig
        results[sname] = toks
    base = results["fp32"]
    print(f"token study, {name}, beam {beam}, {sum(len(t) for t in base)} tokens in the fp32 decode of {len(base)} utterances")
    for sname, toks in results.items():
        diff = sum(1 for a, b in zip(toks, base) if a != b)
        print(f"  {sname:10s}: {diff} of {len(base)} utterances differ from fp32")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--tokens":
        token_study(*(sys.argv[2:3] or ["zipformer-30m"]))
    else:
        main()
