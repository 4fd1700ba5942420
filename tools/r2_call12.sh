#!/bin/bash
# GPU call 12 of round 2: converter with F2FP.SATFINITE (one third fewer instructions per element).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/gemm_bench.py f16x3 6 8 2>&1 | tail -8 > gpurun_out/r4g_gemm.log; cat gpurun_out/r4g_gemm.log
timeout 300 python tools/profile_pass.py 4 2>&1 | tail -3
echo "== gemm / encoder / search tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "gemm or encoder or c2_slice or decoder or f16x3" > gpurun_out/r4g_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r4g_tests.log
