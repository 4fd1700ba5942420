#!/bin/bash
# GPU call 8 of round 2: search step trace / phase counters.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
B200ASR_SEARCH_PROF=1 timeout 300 python tools/profile_pass.py 3 > gpurun_out/r4c_prof256.log 2>&1
B200ASR_SEARCH_PROF=1 SEGMENTS=64 timeout 300 python tools/profile_pass.py 3 > gpurun_out/r4c_prof64.log 2>&1
tail -4 gpurun_out/r4c_prof256.log | cut -c1-900
