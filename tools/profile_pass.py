"""Two device-resident passes of the bench workload (C2: 256 segments, Zipformer-68M, beam 4) and nothing else:
the command profiled under ncu for profiles/ (launch list, --set full captures). Prints the stage timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer


class A:
    segments = int(os.environ.get("SEGMENTS", "256"))
    model = "zipformer-68m"


cfg, paths = bench.model_dir(A.model, 68)
rec = OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"], joiner=paths["joiner"], tokens=paths["tokens"],
                                        decoding_method="modified_beam_search", max_active_paths=4)
h = rec.stage_batch(bench.workload(A, 0))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    rec.run_staged(h)
    print(rec.last_timings(), flush=True)
