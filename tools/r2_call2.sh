#!/bin/bash
# GPU call 2 of round 2: f16 probe (one process per variant), full suite, pipeline sweep with the SM partition, bench.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== f16 probe"
for v in 19 16 3 0 27; do echo "-- variant $v"; timeout 120 python tools/f16_probe.py $v 2>&1 | tail -7; done > gpurun_out/r3b_f16probe.log 2>&1
cat gpurun_out/r3b_f16probe.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r3b_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r3b_tests.log
echo "== pipeline sweep"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 6 2>&1 | tail -3; }
{
run B200ASR_PIPELINE=0
run B200ASR_DEBUG=1
run B200ASR_SM_RESERVE=24
run B200ASR_SM_RESERVE=32
run B200ASR_SM_RESERVE=8
run B200ASR_PIPE_NOOVERLAP=1
run B200ASR_NO_GREEN_CTX=1
run B200ASR_PIPE_KAPPA=30
run B200ASR_PIPE_KAPPA=90
run B200ASR_PIPE_MAX_GROUPS=3
run B200ASR_PIPE_MAX_GROUPS=2
} > gpurun_out/r3b_sweep.log 2>&1
cat gpurun_out/r3b_sweep.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3b_bench.log 2> gpurun_out/r3b_bench.err; echo "rc=$?"; cat gpurun_out/r3b_bench.log; tail -3 gpurun_out/r3b_bench.err
echo "== f16split gemm tests"; B200ASR_GEMM_F16SPLIT=1 timeout 300 python -m pytest tests -m gpu -q -k "gemm_kernels and tc3" > gpurun_out/r3b_f16tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r3b_f16tests.log
echo "== f16split gemm bench"; B200ASR_GEMM_F16SPLIT=1 timeout 200 python tools/gemm_bench.py tc3 6 > gpurun_out/r3b_f16gemm.log 2>&1; tail -10 gpurun_out/r3b_f16gemm.log
echo "== f16split pass"; B200ASR_GEMM_F16SPLIT=1 timeout 300 python tools/profile_pass.py 5 2>&1 | tail -3
B200ASR_GEMM_F16SPLIT=1 B200ASR_PIPELINE=0 timeout 300 python tools/profile_pass.py 5 2>&1 | tail -3
