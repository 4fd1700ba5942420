#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "encoder or softmax or c2_slice or c1_greedy or end_to_end or rover or degenerate or bf16 or tensor_core" > gpurun_out/r4q_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r4q_tests.log
timeout 300 python tools/profile_pass.py 4 2>&1 | tail -3 | cut -c1-200
B200ASR_ATTN_SIMT=1 SEGMENTS=16 timeout 300 python tools/profile_pass.py 2 2>&1 | tail -1 | cut -c1-100
