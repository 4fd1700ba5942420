"""End-to-end probe on a B200: create_stream / accept_waveform / decode_streams over the C2 batch, with the host phases of
decode (B200ASR_HOST_PROF) and the stage timings of the pass. Usage: python tools/e2e_probe.py"""
import sys, os, time
os.environ.setdefault("B200ASR_HOST_PROF", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer
class A: pass
args = A(); args.segments=256; args.model="zipformer-68m"
cfg, paths = bench.model_dir("zipformer-68m", 68)
rec = OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"], joiner=paths["joiner"], tokens=paths["tokens"], decoding_method="modified_beam_search", max_active_paths=4)
audios = bench.workload(args, 0)
for it in range(3):
    t0=time.perf_counter()
    ss=[rec.create_stream() for _ in audios]
    t1=time.perf_counter()
    for s,a in zip(ss,audios): s.accept_waveform(16000,a)
    t2=time.perf_counter()
    rec.decode_streams(ss)
    t3=time.perf_counter()
    print("create %.1f ms accept %.1f ms decode %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), rec.last_timings())
    del ss
    t4=time.perf_counter(); print("del %.1f ms"%((t4-t3)*1e3))
