#!/bin/bash
# round 2, GPU call 21 (1 GPU, the last seconds): smoke() and the golden-fixture GPU tests on the final library (after the Helmholtz template change)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 14 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c21_smoke.log 2>&1; tail -2 gpurun_out/r2c21_smoke.log ) &
timeout 14 python -m pytest tests/test_golden.py -m gpu -q -x > gpurun_out/r2c21_pytest_golden.log 2>&1; tail -3 gpurun_out/r2c21_pytest_golden.log
wait
