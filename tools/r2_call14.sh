#!/bin/bash
# GPU call 14 of round 2 (2 GPUs): weak-scaling C2 bench line and the C5 corpus through the shared pull queue at N = 2.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi -L | head -4
echo "== bench c2 N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r4i_bench_n2.log 2> gpurun_out/r4i_bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/r4i_bench_n2.log | cut -c1-400; grep -o '"e2e": {[^}]*}' gpurun_out/r4i_bench_n2.log; tail -2 gpurun_out/r4i_bench_n2.err
echo "== c5 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --steps 1 --warmup 1 > gpurun_out/r4i_c5_n2.log 2> gpurun_out/r4i_c5_n2.err; echo "rc=$?"; tail -1 gpurun_out/r4i_c5_n2.log | cut -c1-1500; tail -2 gpurun_out/r4i_c5_n2.err
