#!/bin/bash
# round 2, GPU call 1: parity of the plane-marching Godunov kernels + tile / components-per-launch variants + ncu capture
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# the library is rebuilt on the box so that it always matches the snapshot's sources
make -C varden_b200/csrc -j16 > gpurun_out/r2c1_build.log 2>&1 || { tail -20 gpurun_out/r2c1_build.log; exit 1; }
( cd varden_b200/csrc && for v in "16 16 1" "12 12 1" "8 8 1" "16 8 3" "12 8 2"; do set -- $v
    ( nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -fmad=false -DMARCH_TYT=$1 -DMARCH_TYT_VP=$2 -DMARCH_NCG=$3 -c vdn_godunov.cu -o /tmp/god_$1_$2_$3.o &&
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libvdn_t$1_v$2_n$3.so vdn_ctx.o /tmp/god_$1_$2_$3.o vdn_stream.o vdn_mg.o vdn_comm.o -lcudart -lnccl ) &
  done; wait )
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
B="timeout 200 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu --no-e2e"
$B > gpurun_out/r2c1_default.json 2> gpurun_out/r2c1_default.err
for v in t16_v16_n1 t12_v12_n1 t8_v8_n1 t16_v8_n3 t12_v8_n2; do
  VDN_LIB=$PWD/varden_b200/libvdn_$v.so $B > gpurun_out/r2c1_$v.json 2>&1
done
timeout 600 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c1_n512.json 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'march' -s 5 -c 6 -o /tmp/prof_march python bench.py --config 2 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c1_ncu.log 2>&1
cp /tmp/prof_march.ncu-rep gpurun_out/r2c1_march.ncu-rep
ncu -i /tmp/prof_march.ncu-rep --page raw --csv > gpurun_out/r2c1_march_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
