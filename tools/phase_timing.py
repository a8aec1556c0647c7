import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from nuts_rs_b200 import _abi, lib
N = int(os.environ.get('PROF_N', bench.CHAINS_PER_GPU)); d = bench.DIM
math = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=bench.model_sigma())
s = lib.Sampler(math, bench.settings(), seed=bench.SEED)
s.set_position(bench.initial_positions(N, 0)); s.draw_device(bench.NUM_TUNE)
out = (C.c_ulonglong * 8)(); L = lib.load(); L.nuts_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
L.nuts_debug_phase_clocks(s.h, out)
lf0, _ = s.counters(); s.draw_device(10); ms, _ = s.last_timing(); lf1, _ = s.counters()
L.nuts_debug_phase_clocks(s.h, out)
names = ["init_traj", "leapfrog", "leaf+store", "merges", "doubling pro/epilogue", "materialise", "adapt", "whole draw"]
tot = out[7]
print("engine", os.environ.get("NUTS_B200_ENGINE"), "%.2f ms, %d leapfrogs, %.0f leapfrogs/chain-draw" % (ms, lf1 - lf0, (lf1 - lf0) / N / 10))
for n, v in zip(names, out): print("  %-24s %6.1f%%   %8.0f cycles per chain-draw" % (n, 100.0 * v / tot, v / N / 10))
