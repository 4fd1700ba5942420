#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/profile_pass.py 4 2>&1 | tail -2 | cut -c1-200
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "encoder or gemm or c2_slice or c1_greedy or end_to_end or rover or c3_500 or softmax or transcribe_long" > gpurun_out/r4l_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r4l_tests.log
