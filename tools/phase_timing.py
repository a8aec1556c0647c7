"""Per-phase clock totals of an instrumented build (make OUT=../libnuts_b200_phase.so BUILD=build_phase
EXTRA="-DNB_PHASE_TIMING -DNB_PHASE_TIMING_COLD" ...; run with NUTS_B200_LIB=.../libnuts_b200_phase.so): the tuning phase and a
sampling launch of a BASELINE config, cycles per chain-draw of thread 0 of every team."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nuts_rs_b200 import lib

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = bench.CONFIGS[name]
N, d, tune = int(os.environ.get("PROF_N", cfg["chains"])), cfg["dim"], cfg["num_tune"]
math = lib.CudaMath(N, d, cfg["kind"], **cfg["model"](d))
s = lib.Sampler(math, bench.config_settings(cfg), seed=bench.SEED)
s.set_position(bench.initial_positions(N, 0, d))
out = (C.c_ulonglong * 16)()
L = lib.load()
L.nuts_debug_phase_clocks.argtypes = [C.c_void_p, C.c_void_p]
L.nuts_debug_phase_clocks(s.h, out)
names = ["init_traj", "leapfrog", "leaf+store", "merges", "doubling pro/epilogue", "materialise", "adapt (cold call)", "whole draw",
         "cold: load + schedule", "cold: vector pass", "cold: dual avg / step / stats / store", "  pass: load wait", "  pass: estimators + stores", "  pass: mass matrix + stores", "", "cold total"]


def report(label, draws):
    lf0, _ = s.counters()
    s.draw_device(draws)
    ms, _ = s.last_timing()
    lf1, _ = s.counters()
    L.nuts_debug_phase_clocks(s.h, out)
    print(f"{label}: engine {os.environ.get('NUTS_B200_ENGINE')} {ms:.2f} ms, {lf1 - lf0} leapfrogs, {(lf1 - lf0) / N / draws:.1f} per chain-draw, "
          f"{(lf1 - lf0) / ms * 1e3:.4g} lf/s")
    tot = out[7] or 1
    for n, v in zip(names, out):
        if n:
            print("  %-40s %6.1f%%   %9.0f cycles per chain-draw" % (n, 100.0 * v / tot, v / N / draws))


report("tuning", tune)
report("sampling", 10)
