#!/bin/bash
# round 2, GPU call 16 (1 GPU): the full GPU test tier, smoke(), the driver's default bench line, the 256^3 line, config 5 (512^3, 1000:1),
# fuse-min A/B, ncu launch list of one 256^3 step and ncu --set full of the fused smoother / marching kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
make -C varden_b200/csrc -j16 > gpurun_out/r2c16_build.log 2>&1 || { tail -20 gpurun_out/r2c16_build.log; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c16_pytest.log
tail -6 gpurun_out/r2c16_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c16_smoke.log 2>&1; tail -2 gpurun_out/r2c16_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2c16_default_n1.json 2> gpurun_out/r2c16_default_n1.err ) 2> gpurun_out/r2c16_default_n1.time
timeout 300 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r2c16_256.json 2> gpurun_out/r2c16_256.err
timeout 300 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e --fuse-min 64 > gpurun_out/r2c16_256_fusemin64.json 2> gpurun_out/r2c16_256_fusemin64.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c16_512_ratio1000.json 2> gpurun_out/r2c16_512_ratio1000.err
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 2700 --csv --log-file gpurun_out/r2c16_launches.csv \
    python bench.py --config 2 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c16_ncu_launches.log 2>&1
timeout 600 $NCU --set full --import-source on --clock-control none -k regex:"k_sweep3|march" --launch-skip 150 -c 14 -o gpurun_out/r2c16_top \
    python bench.py --config 2 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c16_ncu_full.log 2>&1
$NCU -i gpurun_out/r2c16_top.ncu-rep --page raw --csv > gpurun_out/r2c16_top_raw.csv 2>/dev/null
for f in gpurun_out/r2c16*.err; do echo "== $f"; tail -n 3 "$f"; done
cat gpurun_out/r2c16_default_n1.time
