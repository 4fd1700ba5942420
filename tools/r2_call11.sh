#!/bin/bash
# GPU call 11 of round 2: how the fp16-split GEMM's time depends on the operand bytes per stage (timing experiment, results wrong).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for d in 0 1 2 3; do echo "-- B200ASR_DBG_GEMM=$d"; B200ASR_DBG_GEMM=$d timeout 300 python tools/gemm_bench.py f16x3 6 8 2>&1 | tail -8; done > gpurun_out/r4f_gemm_dbg.log 2>&1
cat gpurun_out/r4f_gemm_dbg.log
