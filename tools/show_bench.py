"""One-screen summary of a bench.py JSON line.   python tools/show_bench.py gpurun_out/bench_8gpu_x.json"""
import json, sys
d = json.load(open(sys.argv[1]))
c = d["config"]
print(f"n_gpus {d['n_gpus']}  value {d['value']:.4g}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.4g}  d2h {d['e2e'].get('d2h_GBps_per_rank', 0):.1f} GB/s/rank")
hs = d["e2e"].get("host_sink_ceiling")
if hs and hs.get("GBps_per_rank"): print(f"host sink ceiling {hs['GBps_per_rank']:.1f} GB/s/rank = {hs['leapfrogs_per_s_all_gpus']:.4g} lf/s")
t = c["tuning_phase"]; print(f"tuning {t['leapfrogs_per_s']:.4g} ({t['kernel_ms']:.1f} ms) ratio {t['ratio_to_sampling_rate']:.3f}")
g = c.get("draw_gather") or {}
if "leapfrogs_per_s" in g: print(f"gather pipeline {g['leapfrogs_per_s']:.4g} ratio {g['ratio_to_value']:.3f} {g['gather_GBps_per_rank']:.0f} GB/s/rank last {g['last_gather_device_ms_rank0']:.2f} ms")
else: print("gather", g)
for k, v in c.get("other_configs", {}).items(): print(f"{k} {v['value']:.4g} tuning {v['tuning_phase_leapfrogs_per_s']:.4g} chains {v['chains_rank0']}/{v['chains_all_ranks']} {v['scaling'][:6]}")
w = c.get("whole_run") or {}
if w: print(f"whole run {w['wall_s']:.3f} s  {w['leapfrogs_per_s']:.4g} lf/s")
print("clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), "roofline frac", round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline", {}).get("value"))
