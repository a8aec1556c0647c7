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
names = ["load+pre", "estimators fg", "estimators bg", "mass matrix", "rest", "-", "-", "cold_adapt total"]
for lo, n in ((0, 100), (100, 200), (300, 100), (400, 50)):
    s.draw_device(n); ms, _ = s.last_timing()
    L.nuts_debug_phase_clocks(s.h, out)
    print("draws %d..%d: %.1f ms  " % (lo, lo + n, ms) + "  ".join("%s %.1fk" % (nm, v / N / n / 1e3) for nm, v in zip(names, out) if nm != "-"))
