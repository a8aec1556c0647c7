"""Short driver for ncu.  python tools/prof_run.py [config] [mode]
  mode sample (default): set_position, tuning, then PROF_LAUNCHES sampling launches of 10 draws (profile with --launch-skip 3)
  mode tune:             set_position, then ONE launch of PROF_TUNE_DRAWS tuning draws (profile with --launch-skip 1 -c 1)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nuts_rs_b200 import lib

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "sample"
cfg = bench.CONFIGS[name]
N, d = int(os.environ.get("PROF_N", cfg["chains"])), cfg["dim"]
math = lib.CudaMath(N, d, cfg["kind"], **cfg["model"](d))
s = lib.Sampler(math, bench.config_settings(cfg), seed=bench.SEED)
assert (s.set_position(bench.initial_positions(N, 0, d)) == 0).all()
if mode == "tune":
    s.draw_device(int(os.environ.get("PROF_TUNE_DRAWS", 100)))
    print(s.last_timing(), s.counters())
else:
    s.draw_device(cfg["num_tune"])
    for _ in range(int(os.environ.get("PROF_LAUNCHES", 3))):
        s.draw_device(10)
        print(s.last_timing(), s.counters())
s.close()
math.close()
