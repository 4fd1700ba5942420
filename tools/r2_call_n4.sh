#!/bin/bash
# 4-GPU weak-scaling C2 bench line
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi -L | wc -l; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-4} --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus ${NG:-4} --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_n${NG:-4}.log 2> gpurun_out/r7_bench_n${NG:-4}.err; echo "rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r7_bench_n${NG:-4}.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}, {k:v for k,v in l['e2e'].items() if k!='mode'}, l['clocks'])
PY
tail -2 gpurun_out/r7_bench_n${NG:-4}.err
