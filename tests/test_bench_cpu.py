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
    assert S.family("void march::k_velpred_march<16, 0, 1>(march::VpmArgs)", "") == "velpred"
    assert S.family("void march::k_mkflux_march<1, 1, 16, 0, 1>(march::MfmArgs)", "") == "mkflux_scal"
    assert S.family("void march::k_mkflux_march<1, 0, 16, 0, 1>(march::MfmArgs)", "") == "mkflux_vel"
    assert S.family("void k_sweep3<0, 2, 32, 16, 0>(WaveArgs)", "") == "mg_wave_down_l0"


def _check_line(line_file, traffic_file, exact):
    txt = open(os.path.join(ROOT, "profiles", line_file)).read().strip().splitlines()[-1]
    d = json.loads(txt)
    t = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
    roof = d["roofline"]
    assert roof["bound"] == "hbm" and 0 < roof["frac"] < 1 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    fam = d["kernels"]
    nl = sum(fam[n]["launches"] for n in roof["families"])
    want = sum(t[n]["dram_bytes_per_launch"] * fam[n]["launches"] for n in roof["families"]) / nl
    if exact:
        assert abs(roof["traffic"] - want) <= 1e-6 * want
    # fused kernels: real DRAM traffic is below the algorithmic (per-colour) accounting, never above it
    assert roof["traffic"] < roof["alg_bytes_per_launch"] and want < roof["alg_bytes_per_launch"]
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["clocks"]["reasons"] == []
    return d


def test_committed_bench_line_and_traffic_agree():
    """round 1: the line's `traffic` is the launch-weighted DRAM bytes of the ncu capture of that round (profiles/ncu_traffic_r01.json).
    round 2: the lines of GPU call 16 were printed while the round-1 table was still in place (0.85 GB per launch); the round-2 capture
    (profiles/ncu_traffic.json, taken in the same call: 0.99 GB per launch -- the inverse-diagonal array is 8 B/cell more) is what later runs
    report; both stay below the algorithmic bytes."""
    _check_line("r01_bench_256_v6_final.json", "ncu_traffic_r01.json", True)
    d = _check_line("r02_bench_c16_256.json", "ncu_traffic.json", False)
    assert d["config"]["workload"].startswith("BASELINE configs[1]") and d["cpu_baseline"]["cores"] >= 1
    d = _check_line("r02_bench_c16_default_n1_512.json", "ncu_traffic.json", False)
    assert d["config"]["workload"].startswith("BASELINE configs[2]") and d["scaling"] == "strong"


def test_committed_scaling_lines_are_the_north_star_config():
    """the strong-scaling lines under profiles/ are BASELINE configs[2] (512^3) at N = 1, 2, 4, 8 and their efficiencies are what DESIGN.md states"""
    t = {}
    for n, f in ((1, "r02_bench_c16_default_n1_512.json"), (2, "r02_bench_c10_strong_n2_push.json"), (4, "r02_bench_c14_strong_n4_push.json"),
                 (8, "r02_bench_c18_strong_n8_push.json")):
        d = json.loads(open(os.path.join(ROOT, "profiles", f)).read().strip().splitlines()[-1])
        assert d["n_gpus"] == n and d["scaling"] == "strong" and "512x512x512" in d["config"]["workload"]
        t[n] = d["ms_per_step"]
    eff = {n: t[1] / (n * t[n]) for n in t}
    assert eff[2] > 0.85 and eff[4] > 0.75 and eff[8] > 0.65
