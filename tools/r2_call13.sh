#!/bin/bash
# GPU call 13 of round 2: single-pass attention weights (diagonal shift, consumers normalise).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q -x -s -k "encoder or softmax or c2_slice or rover or c1_greedy or end_to_end" > gpurun_out/r4h_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r4h_tests.log; grep "encoder rel" gpurun_out/r4h_tests.log | tail -4 | cut -c1-300
timeout 300 python tools/profile_pass.py 4 2>&1 | tail -3
B200ASR_SOFTMAX_2PASS=1 timeout 300 python tools/profile_pass.py 4 2>&1 | tail -2
