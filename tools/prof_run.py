"""Short driver for ncu: set_position, tuning, then a few draw launches of the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from nuts_rs_b200 import _abi, lib
N = int(os.environ.get("PROF_N", bench.CHAINS_PER_GPU)); d = bench.DIM
math = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=bench.model_sigma())
s = lib.Sampler(math, bench.settings(), seed=bench.SEED)
assert (s.set_position(bench.initial_positions(N, 0)) == 0).all()
s.draw_device(bench.NUM_TUNE)
for _ in range(int(os.environ.get("PROF_LAUNCHES", 3))):
    s.draw_device(10)
    print(s.last_timing(), s.counters())
s.close(); math.close()
