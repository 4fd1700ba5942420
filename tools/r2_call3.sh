#!/bin/bash
# GPU call 3 of round 2: dynamic tile scheduler validation, pipeline sweep on top of it, fp16-split GEMM as the FP32 mode.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r3c_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r3c_tests.log
echo "== pipeline sweep"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 6 2>&1 | tail -3; }
{
run B200ASR_PIPELINE=0
run B200ASR_PIPELINE=0 B200ASR_STATIC_TILES=1
run B200ASR_SM_RESERVE=16
run B200ASR_SM_RESERVE=16 B200ASR_STATIC_TILES=1
run B200ASR_SM_RESERVE=8
run B200ASR_SM_RESERVE=24
run B200ASR_SM_RESERVE=32
run B200ASR_SM_RESERVE=0
run B200ASR_SM_RESERVE=16 B200ASR_PIPE_MAX_GROUPS=3
run B200ASR_SM_RESERVE=16 B200ASR_PIPE_KAPPA=30
} > gpurun_out/r3c_sweep.log 2>&1
cat gpurun_out/r3c_sweep.log
echo "== f16split gemm tests"; B200ASR_GEMM_F16SPLIT=1 timeout 300 python -m pytest tests -m gpu -q -k "gemm_kernels and tc3" > gpurun_out/r3c_f16tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r3c_f16tests.log
echo "== f16split gemm bench"; B200ASR_GEMM_F16SPLIT=1 timeout 200 python tools/gemm_bench.py tc3 6 > gpurun_out/r3c_f16gemm.log 2>&1; tail -10 gpurun_out/r3c_f16gemm.log
timeout 200 python tools/gemm_bench.py tc3 6 > gpurun_out/r3c_tc3gemm.log 2>&1; tail -10 gpurun_out/r3c_tc3gemm.log
echo "== f16split pass"; 
{
run B200ASR_GEMM_F16SPLIT=1 B200ASR_PIPELINE=0
run B200ASR_GEMM_F16SPLIT=1
} 2>&1
echo "== f16split full suite"; B200ASR_GEMM_F16SPLIT=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r3c_f16suite.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r3c_f16suite.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3c_bench.log 2> gpurun_out/r3c_bench.err; echo "rc=$?"; cat gpurun_out/r3c_bench.log; tail -3 gpurun_out/r3c_bench.err
