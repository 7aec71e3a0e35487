#!/bin/bash
# compute-sanitizer pass over the smoke step (memcheck: out-of-bounds / misaligned accesses; initcheck: reads of uninitialised global memory --
# the umac ghost faces deliberately hold the reference's 1.d20 poison, advance_timestep.f90:77, which IS initialised).  Run on the GPU box:
#   bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck initcheck; do
  echo "=== compute-sanitizer --tool $tool: __graft_entry__.smoke() ==="
  timeout 900 $CS --tool $tool --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25
  echo "exit code: ${PIPESTATUS[0]}"
done
