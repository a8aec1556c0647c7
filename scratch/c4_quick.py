import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nuts_rs_b200 import _abi, lib
N = int(os.environ.get("BENCH_N", 256)); d = int(os.environ.get('BENCH_D', 10000)); tune = int(os.environ.get("BENCH_TUNE", 300))
s = lib.DiagNutsSettings(num_tune=tune, seed=42)
x0 = np.random.default_rng(42).normal(size=(N, d))
m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.0, sigma=10 ** np.linspace(-3, 3, d)); S = lib.Sampler(m, s, seed=42)
st = S.set_position(x0)
S.draw_device(tune); tune_ms, _ = S.last_timing(); lf_t, _ = S.counters()
draws, stats = S.draw(20)
ms, _ = S.last_timing(); lf, _ = S.counters()
z = draws / 10 ** np.linspace(-3, 3, d)
print(f"d={d} N={N} engine {os.environ.get('NUTS_B200_ENGINE')} tune {tune_ms:.0f} ms ({lf_t/tune_ms*1e3:.3g} lf/s) sample {ms:.1f} ms ({(lf-lf_t)/ms*1e3:.3g} lf/s) depth {stats['depth'].mean():.2f} "
      f"div {stats['diverging'].mean():.4f} z std {z.std():.4f} checksum {draws.sum():.10e}")
