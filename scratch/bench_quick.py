"""Quick device-arm throughput for the bench workload with the engine chosen by NUTS_B200_ENGINE."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nuts_rs_b200 import _abi, lib
N = int(os.environ.get('BENCH_N', bench.CHAINS_PER_GPU)); d = bench.DIM
math = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=bench.model_sigma())
s = lib.Sampler(math, bench.settings(), seed=bench.SEED)
assert (s.set_position(bench.initial_positions(N, 0)) == 0).all()
s.draw_device(bench.NUM_TUNE); tune_ms, _ = s.last_timing(); lf_t, _ = s.counters()
tot_ms = 0; lf0, _ = s.counters()
for _ in range(5):
    s.draw_device(10); ms, _ = s.last_timing(); tot_ms += ms
lf1, _ = s.counters()
print("N", N, "engine", os.environ.get("NUTS_B200_ENGINE"), "tune: %.1f ms %.3g lf/s | sample: %.2f ms/10 draws, %.4g leapfrogs/s" % (tune_ms, lf_t / tune_ms * 1e3, tot_ms / 5, (lf1 - lf0) / tot_ms * 1e3))
s.close(); math.close()
