#!/bin/bash
# GPU call 5 of round 2: new tests, chained passes, bench (c2 / c3 / c4), ncu launch list + full captures of the steady pass.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== new / changed tests"; timeout 900 python -m pytest tests -m gpu -q -s -k "bf16_mode or several_batches or staging or pipeline_flags or f16x3 or gemm_kernels" > gpurun_out/r3e_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r3e_tests.log
grep -h "mode token edit\|encoder rel_l2 by mode\|suspect words\|chained batches" gpurun_out/r3e_tests.log
echo "== chained passes"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 4 2>&1 | tail -3; }
{
run CHAIN=0
run CHAIN=4 B200ASR_SM_RESERVE=0
run CHAIN=4 B200ASR_SM_RESERVE=8
run CHAIN=4 B200ASR_SM_RESERVE=16
run CHAIN=4 B200ASR_SM_RESERVE=32
run CHAIN=8 B200ASR_SM_RESERVE=8
run CHAIN=2 B200ASR_SM_RESERVE=8
} > gpurun_out/r3e_chain.log 2>&1
cat gpurun_out/r3e_chain.log
echo "== bench c2"; timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/r3e_bench.log 2> gpurun_out/r3e_bench.err; echo "rc=$?"; cat gpurun_out/r3e_bench.log; tail -3 gpurun_out/r3e_bench.err
echo "== bench c3"; timeout 600 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench_c3.log 2> gpurun_out/r3e_bench_c3.err; echo "rc=$?"; cat gpurun_out/r3e_bench_c3.log; tail -3 gpurun_out/r3e_bench_c3.err
echo "== bench c4"; timeout 600 python bench.py --workload c4 --steps 4 --warmup 2 > gpurun_out/r3e_bench_c4.log 2> gpurun_out/r3e_bench_c4.err; echo "rc=$?"; cat gpurun_out/r3e_bench_c4.log; tail -3 gpurun_out/r3e_bench_c4.err
echo "== ncu launch list"
timeout 300 python tools/profile_pass.py 3 > gpurun_out/r3e_plain.log 2>&1 && \
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3e_launches.csv python tools/profile_pass.py 3 > gpurun_out/r3e_ncu1.log 2>&1
echo "ncu1 rc=$?"; tail -2 gpurun_out/r3e_ncu1.log
echo "== ncu full: gemm / search step / attention"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_f16_tcgen05 -s 60 -c 3 -o gpurun_out/r3e_gemm_f16 python tools/profile_pass.py 3 > gpurun_out/r3e_ncu2.log 2>&1; echo "ncu2 rc=$?"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:decoder_joinin|select_partials" -s 600 -c 2 -o gpurun_out/r3e_search_step python tools/profile_pass.py 3 > gpurun_out/r3e_ncu3.log 2>&1; echo "ncu3 rc=$?"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:attn_weights_tcgen05|attn_apply_tcgen05|glu_dwconv|fbank_kernel" -c 5 -o gpurun_out/r3e_encoder_kernels python tools/profile_pass.py 3 > gpurun_out/r3e_ncu4.log 2>&1; echo "ncu4 rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -5
