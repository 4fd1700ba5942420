"""Separates the accept-time PCM upload from the decode pass: accept, device-wide sync, decode, with two generations of
streams alive as in bench.py. Usage: python tools/upload_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200ASR_HOST_PROF", "1")
import numpy as np, torch
import bench
from sherpa_vietnamese_asr_b200.recognizer import OfflineRecognizer
class A: pass
args = A(); args.segments=256; args.model="zipformer-68m"
cfg, paths = bench.model_dir("zipformer-68m", 68)
rec = OfflineRecognizer.from_transducer(encoder=paths["encoder"], decoder=paths["decoder"], joiner=paths["joiner"], tokens=paths["tokens"], decoding_method="modified_beam_search", max_active_paths=4)
audios = bench.workload(args, 0)
for it in range(4):
    t0=time.perf_counter()
    ss=[rec.create_stream() for _ in audios]
    for s,a in zip(ss,audios): s.accept_waveform(16000,a)
    t1=time.perf_counter()
    torch.cuda.synchronize()
    t2=time.perf_counter()
    rec.decode_streams(ss)
    t3=time.perf_counter()
    print("accept %.1f ms  sync-after-accept %.1f ms  decode %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), flush=True)
    if it == 1: keep = ss   # second generation alive, like bench.py
