#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() { echo "-- $*"; env "$@" timeout 300 python tools/profile_pass.py 4 2>&1 | tail -2 | cut -c1-200; }
{
run B200ASR_EMBED_CHUNK_MB=0
run B200ASR_EMBED_CHUNK_MB=16
run B200ASR_EMBED_CHUNK_MB=32
run B200ASR_EMBED_CHUNK_MB=64
} > gpurun_out/r4k_embed.log 2>&1
cat gpurun_out/r4k_embed.log
