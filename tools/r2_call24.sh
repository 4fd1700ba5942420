#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
B200ASR_HOST_PROF=1 CHAIN=4 timeout 300 python tools/profile_pass.py 2 2>&1 | grep "issue" | tail -8 | cut -c1-160
timeout 300 python tools/profile_pass.py 4 2>&1 | tail -2 | cut -c1-200
CHAIN=4 timeout 300 python tools/profile_pass.py 3 2>&1 | tail -6
