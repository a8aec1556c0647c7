import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from nuts_rs_b200 import _abi, lib
N = bench.CHAINS_PER_GPU; d = bench.DIM
math = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=bench.model_sigma())
s = lib.Sampler(math, bench.settings(), seed=bench.SEED)
s.set_position(bench.initial_positions(N, 0)); s.draw_device(bench.NUM_TUNE)
out = (C.c_ulonglong * 8)(); L = lib.load(); L.nuts_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
L.nuts_debug_phase_clocks(s.h, out)
lf0, _ = s.counters(); s.draw_device(10); ms, _ = s.last_timing(); lf1, _ = s.counters()
L.nuts_debug_phase_clocks(s.h, out)
busy, total, polls, prog, events = out[0], out[1], out[2], out[3], out[4]
print("kernel %.2f ms, %d leapfrogs; leader busy %.1f%% of its time; %.0f cycles per productive poll, %.2f lanes per productive poll, %d events (%.2f per leapfrog), polls %d"
      % (ms, lf1 - lf0, 100.0 * busy / max(total, 1), busy / max(prog, 1), events / max(prog, 1), events, events / (lf1 - lf0), polls))
