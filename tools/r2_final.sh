#!/bin/bash
# Final GPU call of round 2: full suite, bench lines (c2 with CPU baseline, c3, c4, bf16, reference arm), ncu launch list and full captures.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T=${TAG:-r5}
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${T}_tests.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${T}_smoke.log | cut -c1-200
echo "== bench c2"; timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/${T}_bench.log; tail -2 gpurun_out/${T}_bench.err
echo "== bench c3"; timeout 600 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c3.log 2> gpurun_out/${T}_bench_c3.err; echo "rc=$?"
echo "== bench c4"; timeout 600 python bench.py --workload c4 --steps 4 --warmup 2 > gpurun_out/${T}_bench_c4.log 2> gpurun_out/${T}_bench_c4.err; echo "rc=$?"
echo "== bench bf16"; timeout 600 python bench.py --precision bf16 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_bf16.log 2> gpurun_out/${T}_bench_bf16.err; echo "rc=$?"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.log 2> gpurun_out/${T}_bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/${T}_bench_ref.log
echo "== ncu launch list"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_pass.py 3 > gpurun_out/${T}_ncu1.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_f16_tcgen05 -s 60 -c 3 -o gpurun_out/${T}_gemm_f16 python tools/profile_pass.py 3 > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu2 rc=$?"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:joiner_f16ss|select_partials" -s 600 -c 2 -o gpurun_out/${T}_search_step python tools/profile_pass.py 3 > gpurun_out/${T}_ncu3.log 2>&1; echo "ncu3 rc=$?"
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:attn_weights_tcgen05|attn_apply_tcgen05|glu_dwconv" -c 4 -o gpurun_out/${T}_encoder_kernels python tools/profile_pass.py 3 > gpurun_out/${T}_ncu4.log 2>&1; echo "ncu4 rc=$?"
ls -la gpurun_out/${T}_*.ncu-rep
