#!/bin/bash
# GPU call 7 of round 2: search step time against the number of live rows (trace by eighth), feed-forward row chunks in L2.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 5 2>&1 | grep -v "^\[b200asr search prof" | tail -5; }
{
run B200ASR_SEARCH_PROF=1
run B200ASR_SEARCH_PROF=1 SEGMENTS=64
run B200ASR_FFN_CHUNK_MB=16
run B200ASR_FFN_CHUNK_MB=32
run B200ASR_FFN_CHUNK_MB=48
run B200ASR_FFN_CHUNK_MB=0
} > gpurun_out/r4b_sweep.log 2>&1
cat gpurun_out/r4b_sweep.log
