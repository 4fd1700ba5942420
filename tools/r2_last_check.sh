#!/bin/bash
# last check of the round on the library as shipped: smoke, full GPU suite, default bench invocation
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r10_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r10_tests.log
timeout 900 python bench.py > gpurun_out/r10_bench.log 2> gpurun_out/r10_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r10_bench.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','steps','warmup')}, l['e2e']['value'], l['e2e']['ms_per_step'], l['config']['parity_checked'], l['config']['parity_mismatches'], l['cpu_baseline']['value'], l['roofline']['frac'], l['roofline']['traffic'])
PY
