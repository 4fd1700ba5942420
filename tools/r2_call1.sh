#!/bin/bash
# GPU call 1 of round 2: full GPU suite, f16-split bring-up probe, pipeline sweep, bench.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3a_tests.log
echo "== f16 probe"; timeout 240 python tools/f16_probe.py > gpurun_out/r3a_f16probe.log 2>&1; echo "rc=$?"; cat gpurun_out/r3a_f16probe.log | tail -12
echo "== pipeline sweep"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 5 2>&1 | tail -3; }
{
run B200ASR_PIPELINE=0
run B200ASR_SM_RESERVE=16
run B200ASR_SM_RESERVE=8
run B200ASR_SM_RESERVE=24
run B200ASR_SM_RESERVE=0
run B200ASR_SM_RESERVE=16 B200ASR_GROUPS=32,32,64
run B200ASR_SM_RESERVE=16 B200ASR_PIPE_MAX_GROUPS=4
run B200ASR_SM_RESERVE=16 B200ASR_PIPE_KAPPA=90
run B200ASR_SM_RESERVE=12 B200ASR_PIPE_MIN_AUDIO=100
} > gpurun_out/r3a_sweep.log 2>&1
cat gpurun_out/r3a_sweep.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3a_bench.log 2> gpurun_out/r3a_bench.err; echo "rc=$?"; cat gpurun_out/r3a_bench.log; tail -3 gpurun_out/r3a_bench.err
echo "== f16split gemm tests"; B200ASR_GEMM_F16SPLIT=1 timeout 300 python -m pytest tests -m gpu -q -k "gemm_kernels and tc3" > gpurun_out/r3a_f16tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r3a_f16tests.log
echo "== f16split gemm bench"; B200ASR_GEMM_F16SPLIT=1 timeout 200 python tools/gemm_bench.py tc3 6 > gpurun_out/r3a_f16gemm.log 2>&1; tail -10 gpurun_out/r3a_f16gemm.log
timeout 200 python tools/gemm_bench.py tc3 6 > gpurun_out/r3a_tc3gemm.log 2>&1; tail -10 gpurun_out/r3a_tc3gemm.log
echo "== f16split pass"; B200ASR_GEMM_F16SPLIT=1 timeout 300 python tools/profile_pass.py 5 2>&1 | tail -3
