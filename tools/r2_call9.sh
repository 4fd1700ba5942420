#!/bin/bash
# GPU call 9 of round 2: pre-split joiner GEMM (both operands from TMA) in the decoder-table search.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== search tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "search or decoder or c2_slice or c3_500 or end_to_end or trailing_empty or hotwords" > gpurun_out/r4d_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r4d_tests.log
B200ASR_SEARCH_PROF=1 timeout 300 python tools/profile_pass.py 3 > gpurun_out/r4d_prof256.log 2>&1
B200ASR_SEARCH_PROF=1 B200ASR_JOINER_SS=0 timeout 300 python tools/profile_pass.py 3 > gpurun_out/r4d_prof256_conv.log 2>&1
grep "search trace" gpurun_out/r4d_prof256.log | tail -2 | cut -c1-700
tail -2 gpurun_out/r4d_prof256.log | cut -c1-400
tail -2 gpurun_out/r4d_prof256_conv.log | cut -c1-400
