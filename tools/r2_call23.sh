#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for d in 0 16 19 27; do echo "-- B200ASR_DBG_ATTN=$d"; B200ASR_DBG_ATTN=$d timeout 300 python tools/profile_pass.py 4 2>&1 | tail -2 | cut -c1-120; done
