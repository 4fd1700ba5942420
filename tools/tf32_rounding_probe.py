import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sherpa_vietnamese_asr_b200 import weights
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer
d = tempfile.mkdtemp(); p = weights.write_model_dir(d, weights.zipformer_tiny(), 3)
rec = OfflineRecognizer.from_transducer(encoder=p["encoder"], decoder=p["decoder"], joiner=p["joiner"], tokens=p["tokens"])
A = np.zeros((128, 32), np.float32); W = np.zeros((16, 32), np.float32)
A[:, 0] = np.float32(1 + 3 * 2.0 ** -12); W[:, 0] = 1.0
A[:, 1] = np.float32(-(1 + 3 * 2.0 ** -12)); W[1, :] = 0; W[1, 1] = 1.0
W[2, :] = 0; W[2, 2] = np.float32(1 + 3 * 2.0 ** -12); A[:, 2] = 1.0
C, _ = rec.gemm(A, W, impl="tc")
print("A operand +:", repr(C[0, 0]), " A operand -:", repr(C[0, 1]), " W operand:", repr(C[0, 2]))
print("truncation" if C[0, 0] == 1.0 else "rounding", "| expected RN value", 1 + 2.0 ** -10)
