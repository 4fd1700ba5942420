#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== search tests"; timeout 1200 python -m pytest tests -m gpu -q -x -k "search or decoder or c2_slice or c3_500 or end_to_end or trailing_empty or hotwords or rover or c1_greedy or spanning or threads or regrown or degenerate or 300" > gpurun_out/r4n_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r4n_tests.log
B200ASR_SEARCH_PROF=1 timeout 300 python tools/profile_pass.py 3 > gpurun_out/r4n_prof.log 2>&1
grep "b200asr search" gpurun_out/r4n_prof.log | tail -3 | cut -c1-700
tail -2 gpurun_out/r4n_prof.log | cut -c1-200
