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
chain = int(os.environ.get("CHAIN", "0"))
if chain > 1:     # `chain` passes issued back to back: search of pass k beside the encoder of pass k + 1
    for i in range(n):
        ntok, ms = rec.run_staged_chained(h, chain)
        print(f"chained x{chain}: {ms:.2f} ms total, {ms / chain:.2f} ms per pass", rec.last_pipeline_stats(), "tokens", int(ntok.sum()), flush=True)
        if i == n - 1:      # text timeline: one row per batch, '=' encoder, '#' search, 1 column = 2 ms
            tl = rec.last_pipeline_timeline()
            print("timeline (device ms from the start of the call; batch: encoder begin-end | search begin-end)")
            for g, (a, b, c, d) in enumerate(tl):
                row = [" "] * (int(ms / 2) + 2)
                for x in range(int(a / 2), int(b / 2) + 1):
                    row[x] = "="
                for x in range(int(c / 2), int(d / 2) + 1):
                    row[x] = "#" if row[x] == " " else "%"
                print(f"  batch {g}: {a:7.2f}-{b:7.2f} | {c:7.2f}-{d:7.2f}  " + "".join(row))
    sys.exit(0)
tot = []
prof = os.environ.get("PROFILE_LAST") is not None      # ncu --profile-from-start off: only the last pass is captured
if prof:
    from cuda import cuda as _cu
for i in range(n):
    if prof and i == n - 1:
        _cu.cuProfilerStart()
    ntok = rec.run_staged(h)
    if prof and i == n - 1:
        _cu.cuProfilerStop()
    tm = rec.last_timings()
    tot.append(tm["total_ms"])
    if i >= n - 2:
        print({k: round(v, 2) if isinstance(v, float) else v for k, v in tm.items()}, rec.last_pipeline_stats(), "tokens", int(ntok.sum()), flush=True)
print("TOTAL_MS", " ".join(f"{t:.2f}" for t in tot), flush=True)
