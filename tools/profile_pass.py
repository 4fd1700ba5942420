"""Device-resident passes of the bench workload (C2: 256 segments, Zipformer-68M, beam 4) and nothing else:
the command profiled under ncu for profiles/ (launch list, --set full captures). Prints the stage timings and the
pipeline shape of every pass (environment switches: B200ASR_PIPELINE, B200ASR_GROUPS, B200ASR_SM_RESERVE, ...)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer


class A:
    segments = int(os.environ.get("SEGMENTS", "256"))
    model = os.environ.get("MODEL", "zipformer-68m")


cfg, paths = bench.model_dir(A.model, 68 if "68" in A.model else 30)
rec = OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"], joiner=paths["joiner"], tokens=paths["tokens"],
                                        decoding_method="modified_beam_search", max_active_paths=4,
                                        precision=os.environ.get("PRECISION", "fp32"))
h = rec.stage_batch(bench.workload(A, 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tot = []
for i in range(n):
    ntok = rec.run_staged(h)
    tm = rec.last_timings()
    tot.append(tm["total_ms"])
    if i >= n - 2:
        print({k: round(v, 2) if isinstance(v, float) else v for k, v in tm.items()}, rec.last_pipeline_stats(), "tokens", int(ntok.sum()), flush=True)
print("TOTAL_MS", " ".join(f"{t:.2f}" for t in tot), flush=True)
