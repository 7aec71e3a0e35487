"""CPU-side checks of the measurement contract: the reference arm of bench.py prints one well-formed JSON line, the ncu summariser maps
kernel names to the bench's kernel families, and the committed round evidence is self-consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-n", "32"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Gcell-updates/s") and d["unit"] == "Gcell-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--cpu-n", "32"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ncu_summariser_maps_kernels_to_families():
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import summarize_ncu as S
    assert S.family("void k_sweep3<0, 2, 32, 16>(WaveArgs)", "") == "mg_wave_down_l0"
    assert S.family("void k_sweep3<0, 3, 32, 24>(WaveArgs)", "") == "mg_wave_up_l0"
    assert S.family("void k_sweep3<1, 0, 32, 24>(WaveArgs)", "") == "mg_wave_pro_l0"
    assert S.family("void k_sweep3<0, 0, 32, 24>(WaveArgs)", "") == "mg_wave_smooth_l0"
    assert S.family("void k_sweep<1, 0, 2, 64, 32, 1>(WaveArgs)", "") == "mg_wave_down_l0"
    assert S.family("k_mf_trans6(MfArgs)", "") == "mf_trans6"
    assert S.family("void <unnamed>::k_setval(SetArgs)", "") is None


def test_committed_bench_line_and_traffic_agree():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_256_v6_final.json")))
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    roof = d["roofline"]
    assert roof["bound"] == "hbm" and 0 < roof["frac"] < 1 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    fam = d["kernels"]
    nl = sum(fam[n]["launches"] for n in roof["families"])
    want = sum(t[n]["dram_bytes_per_launch"] * fam[n]["launches"] for n in roof["families"]) / nl
    assert abs(roof["traffic"] - want) <= 1e-6 * want
    # fused kernels: real DRAM traffic is below the algorithmic (per-colour) accounting, never above it
    assert roof["traffic"] < roof["alg_bytes_per_launch"]
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["clocks"]["reasons"] == []
