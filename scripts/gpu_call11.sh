#!/bin/bash
# GPU call 11: k_sweep3 640-thread tiles (no spills) A/B
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
for T in 3 4; do VDN_MG_TILE=$T $B > gpurun_out/b11_f4_tile$T.json 2>&1; done
$B > gpurun_out/b11_default.json 2>&1
ls -la gpurun_out
