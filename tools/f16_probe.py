"""Hardware bring-up of the fp16-hi / bf16-lo operand-split GEMM (csrc/gemm_tc_f16.cu) through the C-ABI (B200AsrGemm, impl
tc3 with B200ASR_GEMM_F16SPLIT=1). For each bring-up variant (B200ASR_F16_VARIANT, read per launch) prints the error against a
float64 product on a few shapes, so one GPU call tells which assumptions about tcgen05 kind::f16 hold:
  0   the design (A_lo(bf16) W_hi(f16) + A_hi(f16) W_lo(bf16) + A_hi W_hi)         expect ~1e-6
  3   hi*hi only                                                                    expect ~5e-4 (fp16 rounding)
  11  hi*hi only, half-words of a packed A column swapped                           expect ~5e-4 iff the order is the other one
  16  lo parts as fp16 too (no mixed a/b formats)                                   expect ~1e-6 for O(1) data
  5 / 6  a single cross term (lo*hi dropped + hi*hi dropped -> only hi*lo, etc.)    tells which mixed descriptor misbehaves
"""
import os
import sys
import tempfile

os.environ["B200ASR_GEMM_F16SPLIT"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sherpa_vietnamese_asr_b200 import weights  # noqa: E402
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer  # noqa: E402

d = tempfile.mkdtemp()
p = weights.write_model_dir(d, weights.zipformer_tiny(), 3)
rec = OfflineRecognizer.from_transducer(encoder=p["encoder"], decoder=p["decoder"], joiner=p["joiner"], tokens=p["tokens"])
rng = np.random.default_rng(0)
shapes = [(300, 272, 192), (1000, 128, 64), (517, 48, 192), (777, 192, 2432), (260, 130, 144), (3000, 512, 512)]
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16]   # one process per variant: a trap is sticky


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


for (M, N, K) in shapes:
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    want = A.astype(np.float64) @ W.astype(np.float64).T
    # the exact split the kernel makes, for the single-term variants
    Ah = A.astype(np.float16).astype(np.float64)
    Wh = W.astype(np.float16).astype(np.float64)
    Al, Wl = A.astype(np.float64) - Ah, W.astype(np.float64) - Wh
    A6 = (A.astype(np.float64) * 64.0).astype(np.float16).astype(np.float64)      # the scaled all-fp16 variant's hi parts
    W10 = (W.astype(np.float64) * 1024.0).astype(np.float16).astype(np.float64)
    terms = {0: want, 16: want, 3: Ah @ Wh.T, 11: Ah @ Wh.T, 5: Ah @ Wl.T, 6: Al @ Wh.T, 19: A6 @ W10.T / 65536.0,
             27: A6 @ W10.T / 65536.0}
    line = []
    for v in variants:
        os.environ["B200ASR_F16_VARIANT"] = str(v)
        try:
            got, _ = rec.gemm(A, W, None, None, act=0, impl="tc3")
            ref = terms.get(v, want)
            line.append(f"v{v}: {rel(got, ref):.2e} (vs full {rel(got, want):.2e})")
        except Exception as e:  # noqa: BLE001
            line.append(f"v{v}: ERROR {e}")
    print(f"M={M} N={N} K={K} | " + " | ".join(line), flush=True)
os.environ["B200ASR_F16_VARIANT"] = "0"
