import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from nuts_rs_b200 import _abi, lib
N = bench.CHAINS_PER_GPU; d = bench.DIM
math = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=bench.model_sigma())
s = lib.Sampler(math, bench.settings(), seed=bench.SEED)
s.set_position(bench.initial_positions(N, 0))
out = (C.c_ulonglong * 8)(); L = lib.load(); L.nuts_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
L.nuts_debug_phase_clocks(s.h, out)
names = ["init_traj", "leapfrog", "leaf+store", "merges", "doubling pro/epilogue", "materialise", "adapt", "whole draw"]
for lo, n in ((0, 100), (100, 150), (250, 150), (400, 50)):
    lf0, _ = s.counters(); s.draw_device(n); ms, _ = s.last_timing(); lf1, _ = s.counters()
    L.nuts_debug_phase_clocks(s.h, out)
    tot = out[7]
    print("draws %d..%d: %.1f ms, %.1f leapfrogs/chain-draw, %.3g lf/s" % (lo, lo + n, ms, (lf1 - lf0) / N / n, (lf1 - lf0) / ms * 1e3))
    print("   " + "  ".join("%s %.0fk" % (nm, v / N / n / 1e3) for nm, v in zip(names, out)))
