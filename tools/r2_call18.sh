#!/bin/bash
# ncu launch list of the steady pass
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/profile_pass.py 5 > gpurun_out/r4m_plain.log 2>&1; tail -3 gpurun_out/r4m_plain.log | cut -c1-220
PROFILE_LAST=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4m_launches.csv python tools/profile_pass.py 3 > gpurun_out/r4m_ncu1.log 2>&1
echo "ncu1 rc=$?"
