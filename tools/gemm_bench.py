"""Micro-benchmark of the GEMM kernels through the C-ABI (B200AsrGemm): device time per launch, TFLOP/s, GB/s."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sherpa_vietnamese_asr_b200 import weights
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer

shapes = [(143000, 384, 192), (143000, 192, 640), (143000, 272, 192), (400000, 384, 128), (400000, 128, 384), (71500, 768, 256),
          (18000, 1920, 512), (143000, 192, 2432), (1024, 2000, 512)]
impls = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "tc", "tc3"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if len(sys.argv) > 3:
    shapes = shapes[: int(sys.argv[3])]
d = tempfile.mkdtemp(); p = weights.write_model_dir(d, weights.zipformer_tiny(), 3)
rec = OfflineRecognizer.from_transducer(encoder=p["encoder"], decoder=p["decoder"], joiner=p["joiner"], tokens=p["tokens"])
rng = np.random.default_rng(0)
for (M, N, K) in shapes:
    A = rng.standard_normal((M, K)).astype(np.float32); W = rng.standard_normal((N, K)).astype(np.float32)
    R = rng.standard_normal((M, N)).astype(np.float32); b = rng.standard_normal(N).astype(np.float32)
    out = []
    for impl in impls:
        _, ms = rec.gemm(A, W, b, R, act=1, impl=impl, reps=reps)
        out.append(f"{impl}: {ms*1000:8.1f} us {2*M*N*K/ms/1e9:7.1f} TFLOP/s {4*(M*K+2*M*N+N*K)/ms/1e6:7.0f} GB/s")
    print(f"M={M} N={N} K={K} | " + " | ".join(out), flush=True)
