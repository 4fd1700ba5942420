#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
B200ASR_SEARCH_PROF=1 timeout 300 python tools/profile_pass.py 2 > gpurun_out/r4o_prof.log 2>&1
grep "b200asr search prof" gpurun_out/r4o_prof.log | tail -1 | cut -c1-700
