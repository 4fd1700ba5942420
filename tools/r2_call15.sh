#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
B200ASR_SEARCH_PROF=1 timeout 300 python tools/profile_pass.py 2 > gpurun_out/r4j_prof.log 2>&1
grep "search prof" gpurun_out/r4j_prof.log | tail -1 | cut -c1-600
echo "== c5 N=1"; timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r4j_c5.log 2> gpurun_out/r4j_c5.err; echo "rc=$?"; grep -o '"per_rank.*"note' gpurun_out/r4j_c5.log | cut -c1-400; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4j_c5.log | head -1; tail -2 gpurun_out/r4j_c5.err
