#!/bin/bash
# GPU call 10: full parity suite on the default configuration (k_sweep3), default bench (e2e + cpu baseline), reference arm,
# ncu launch list + full capture of the top kernels, 512^3 and ratio-1000 runs
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu10.log
tail -4 gpurun_out/pytest_gpu10.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke10.log 2>&1; tail -2 gpurun_out/smoke10.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/b10_default.json 2> gpurun_out/b10_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b10_reference.json 2> gpurun_out/b10_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches10.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:'k_sweep3|k_mf_|k_vp_|k_update|k_mkvelforce' -s 20 -c 40 -o /tmp/prof_top10 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_top10.log 2>&1
ncu -i /tmp/prof_top10.ncu-rep --page raw --csv > gpurun_out/prof_top10_raw.csv 2>/dev/null
timeout 300 python bench.py --ratio 1000 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/b10_ratio1000.json 2>&1
timeout 300 python bench.py --n 512 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b10_n512.json 2>&1
ls -la gpurun_out
