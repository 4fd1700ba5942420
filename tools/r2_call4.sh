#!/bin/bash
# GPU call 4 of round 2: fp16-split as the FP32 mode (engine-owned weight copies), BF16 mode, staging / chain tests, C5 at N=1.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu (FP32 mode = fp16 operand split)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3d_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r3d_tests.log
grep -h "BF16 mode token\|suspect words\|vad timings\|pipeline:" gpurun_out/r3d_tests.log
echo "== pytest -m gpu with 3xTF32"; B200ASR_GEMM_3XTF32=1 timeout 1500 python -m pytest tests -m gpu -q -x -k "not bf16" > gpurun_out/r3d_tests3x.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3d_tests3x.log
echo "== passes"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 6 2>&1 | tail -3; }
{
run B200ASR_PIPELINE=0
run B200ASR_PIPELINE=0 B200ASR_GEMM_3XTF32=1
run B200ASR_PIPELINE=0 PRECISION=bf16
run B200ASR_PIPELINE=0 PRECISION=tf32
} > gpurun_out/r3d_sweep.log 2>&1
cat gpurun_out/r3d_sweep.log
echo "== gemm bench"; timeout 300 python tools/gemm_bench.py tc3,f16x3,bf16 6 > gpurun_out/r3d_gemm.log 2>&1; cat gpurun_out/r3d_gemm.log
echo "== bench"; B200ASR_PIPELINE=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3d_bench.log 2> gpurun_out/r3d_bench.err; echo "rc=$?"; cat gpurun_out/r3d_bench.log; tail -3 gpurun_out/r3d_bench.err
echo "== c5 N=1"; B200ASR_PIPELINE=0 timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r3d_c5.log 2> gpurun_out/r3d_c5.err; echo "rc=$?"; cat gpurun_out/r3d_c5.log; tail -5 gpurun_out/r3d_c5.err
echo "== pipelined path still exact"; B200ASR_PIPELINE=1 timeout 600 python -m pytest tests -m gpu -q -k "pipelined_groups" 2>&1 | tail -3
