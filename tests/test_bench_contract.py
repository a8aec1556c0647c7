"""bench.py's driver contract, checked on the CPU arm (`--impl reference` = the C++ restatement of the reference's CPU path on
all host cores): exactly ONE JSON line on stdout, carrying the metric / unit / config of the GPU arm plus the cpu_baseline and e2e
objects of the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"].startswith("leapfrog-steps/sec") and line["unit"] == "leapfrog-steps/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert "configs[1]" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_bench_lines_follow_the_contract():
    """The evidence under profiles/ is what bench.py printed on the GPU boxes: one object per file with every key of the driver's
    contract (metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype / data /
    config.workload / e2e / gpu_launches / clocks / roofline / cpu_baseline) and the sub-records DESIGN.md section 5 describes."""
    prof = os.path.join(ROOT, "profiles")
    values = {}
    for n, name in [(1, "r2_bench_line.json"), (2, "r2_bench_line_2gpu.json"), (4, "r2_bench_line_4gpu.json"), (8, "r2_bench_line_8gpu.json")]:
        line = json.load(open(os.path.join(prof, name)))
        assert line["metric"].startswith("leapfrog-steps/sec") and line["unit"] == "leapfrog-steps/s" and line["n_gpus"] == n
        assert line["steps"] >= 1 and line["warmup"] >= 3 and line["higher_is_better"] is True and line["scaling"] == "weak"
        assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic"
        assert "configs[1]" in line["config"]["workload"] and line["gpu_launches"] >= line["steps"]
        assert abs(line["value"] - line["config"]["leapfrogs_per_step_all_gpus"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
        e2e = line["e2e"]
        assert 0 < e2e["value"] < line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] > 4e8
        assert e2e["host_sink_ceiling"]["GBps_per_rank"] > 0
        ck = line["clocks"]
        assert ck["sm_mhz"] > 0.9 * ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        rf = line["roofline"]
        assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["traffic"] > 0
        assert set(rf["binding_ceilings"]) >= {"checkpoint_hbm", "fp64_pipe", "algorithmic_48d"}
        cb = line["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        cfg = line["config"]
        assert 0 < cfg["tuning_phase"]["ratio_to_sampling_rate"] < 1.2
        assert set(cfg["other_configs"]) == {"c3", "c4", "c5"} and cfg["other_configs"]["c5"]["chains_all_ranks"] == 65536
        assert cfg["other_configs"]["c5"]["chains_rank0"] == 65536 // n  # config 5 is a strong split of its chains over the ranks
        assert cfg["whole_run"]["wall_s"] > 0 and cfg["draw_gather"]["own_shard_intact_rank0"] is True
        values[n] = line["value"]
    # weak scaling of the sharded path: no collective in the timed region
    for n in (2, 4, 8):
        assert values[n] / (n * values[1]) > 0.95
    # the last line of the round (final build) adds the Tier-2 plane-leapfrog sub-record: TMA-staged rows vs the register path
    last = json.load(open(os.path.join(prof, "r5_bench_line.json")))
    assert last["n_gpus"] == 1 and abs(last["value"] / values[1] - 1) < 0.1
    pl = last["config"]["plane_leapfrog_rank0"]
    assert pl["algorithmic_bytes_per_launch"] == 80 * pl["dim"] * pl["chains"]
    assert pl["tma"]["frac_of_hbm_peak"] > 0.9 > pl["register_path"]["frac_of_hbm_peak"] > 0.5
    assert pl["tma"]["diverged"] == 0 and pl["register_path"]["diverged"] == 0
    ref = json.load(open(os.path.join(prof, "r2_bench_reference_line.json")))
    assert ref["impl"] == "reference" and ref["unit"] == "leapfrog-steps/s" and 0 < ref["value"] < values[1]
