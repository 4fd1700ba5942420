#!/bin/bash
# GPU call 6 of round 2: decoder-table mode of the search (no decoder kernel in a frame step) against the on-demand path.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== search tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "search or decoder or c2_slice or c3_500 or end_to_end or trailing_empty or hotwords" > gpurun_out/r4a_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r4a_tests.log
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 5 2>&1 | tail -4; }
{
run B200ASR_DEC_TABLE=0
run B200ASR_DEC_TABLE=1
run B200ASR_DEC_TABLE=1 B200ASR_SEARCH_PROF=1
run B200ASR_DEC_TABLE=1 CHAIN=4
} > gpurun_out/r4a_sweep.log 2>&1
cat gpurun_out/r4a_sweep.log
