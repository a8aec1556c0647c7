"""Quick device-arm throughput of one BASELINE config (or any shape) with the engine chosen by NUTS_B200_ENGINE.

  python tools/quick.py c2|c3|c4|c5|diag:<dim>:<chains> [tune] [draws_per_launch] [launches]
Prints the tuning-phase and sampling-phase leapfrog rates and a checksum of the draws (to compare engine variants)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from nuts_rs_b200 import lib

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
if ":" in name:  # ad-hoc shape "diag:<dim>:<chains>": the config-2 target at another size
    _, _d, _n = name.split(":")
    cfg = dict(bench.CONFIGS["c2"], dim=int(_d), chains=int(_n))
else:
    cfg = bench.CONFIGS[name]
tune = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["num_tune"]
dpl = int(sys.argv[3]) if len(sys.argv) > 3 else 10
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 5
N = int(os.environ.get("BENCH_N", cfg["chains"]))
d = cfg["dim"]
s = bench.config_settings(cfg, tune)
m = lib.CudaMath(N, d, cfg["kind"], **cfg["model"](d))
S = lib.Sampler(m, s, seed=bench.SEED)
st = S.set_position(bench.initial_positions(N, 0, d))
S.draw_device(tune)
tune_ms, _ = S.last_timing()
lf_t, _ = S.counters()
tot_ms = 0.0
for _ in range(launches):
    S.draw_device(dpl)
    tot_ms += S.last_timing()[0]
lf1, _ = S.counters()
draws, stats = S.draw(4)
print(f"{name} N={N} d={d} engine={os.environ.get('NUTS_B200_ENGINE')} bad_init={(st != 0).sum()} | tune {tune} draws: {tune_ms:.1f} ms "
      f"{lf_t / max(tune_ms, 1e-9) * 1e3:.4g} lf/s | sample: {tot_ms / launches:.3f} ms per {dpl} draws, {(lf1 - lf_t) / tot_ms * 1e3:.4g} lf/s | "
      f"depth {stats['depth'].mean():.2f} div {stats['diverging'].mean():.4f} checksum {np.nansum(draws):.12e}")
S.close()
m.close()
