#!/bin/bash
# full GPU suite + bench line
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r4p_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r4p_tests.log
echo "== bench c2"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4p_bench.log 2> gpurun_out/r4p_bench.err; echo "rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r4p_bench.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, l['e2e'], l['config']['chained_passes'], l['config']['stage_ms'], l['roofline']['frac'], l['roofline']['gemm_ms_per_step'], l['config'].get('parity_checked'), l['config'].get('parity_mismatches'))
PY
tail -2 gpurun_out/r4p_bench.err
