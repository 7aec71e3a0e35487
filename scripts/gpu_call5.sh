#!/bin/bash
# GPU call 5: k_sweep parity + A/B benches over tile / sweeps-per-launch, ncu full capture of k_sweep, PCIe rate
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
tail -5 gpurun_out/pytest_gpu5.log
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$B > gpurun_out/b5_default.json 2>&1
for T in 0 1 2; do
  VDN_MG_TILE=$T $B > gpurun_out/b5_nsw1_tile$T.json 2>&1
  VDN_MG_NSW=2 VDN_MG_TILE=$T $B > gpurun_out/b5_nsw2_tile$T.json 2>&1
done
VDN_MG_FUSE=0 $B > gpurun_out/b5_plain.json 2>&1
VDN_MG_FUSE_MIN=64 $B > gpurun_out/b5_min64.json 2>&1
VDN_MG_FUSE_MIN=256 $B > gpurun_out/b5_min256.json 2>&1
timeout 500 ncu --set full --clock-control none -k regex:'k_sweep' -s 8 -c 12 -o /tmp/prof_sweep5 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_sweep5.log 2>&1
ncu -i /tmp/prof_sweep5.ncu-rep --page raw --csv > gpurun_out/prof_sweep5_raw.csv 2>/dev/null
python - <<'P' > gpurun_out/pcie.txt 2>&1
import torch, time
a = torch.empty(1 << 27, dtype=torch.float64).pin_memory(); d = torch.empty_like(a, device="cuda")
for name, f in (("h2d", lambda: d.copy_(a, non_blocking=True)), ("d2h", lambda: a.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): f()
    torch.cuda.synchronize(); print(name, 3 * a.numel() * 8 / (time.perf_counter() - t) / 1e9, "GB/s")
P
cat gpurun_out/pcie.txt
ls -la gpurun_out
