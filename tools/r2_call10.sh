#!/bin/bash
# GPU call 10 of round 2: six-stage attention application, pipelined end-to-end measurement.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/profile_pass.py 4 2>&1 | tail -3
echo "== attention / encoder tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "encoder or c2_slice" > gpurun_out/r4e_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r4e_tests.log
echo "== bench c2"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4e_bench.log 2> gpurun_out/r4e_bench.err; echo "rc=$?"; cat gpurun_out/r4e_bench.log; tail -3 gpurun_out/r4e_bench.err
