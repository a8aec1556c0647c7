"""Decoupled engine vs the register-resident 64x16 engine: same seeds => bit-identical draws and statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nuts_rs_b200 import _abi, lib

def run(engine, N, d, num_tune, n, maxdepth=8, **kw):
    if engine: os.environ["NUTS_B200_ENGINE"] = engine
    else: os.environ.pop("NUTS_B200_ENGINE", None)
    m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
    st = lib.DiagNutsSettings(num_tune=num_tune, maxdepth=maxdepth, **kw)
    s = lib.Sampler(m, st, seed=7)
    x0 = np.random.default_rng(1).normal(size=(N, d))
    status = s.set_position(x0)
    draws, stats = s.draw(n)
    lf, _ = s.counters()
    s.close(); m.close()
    return status, draws, stats, lf

for (N, d, tune, n, md) in [(8, 1000, 0, 3, 3), (64, 1000, 30, 40, 8), (300, 700, 60, 30, 10), (1100, 1000, 20, 10, 6)]:
    a = run("64,16,4", N, d, tune, n, md)
    b = run("64,16,107", N, d, tune, n, md)
    ok = (a[0] == b[0]).all() and np.array_equal(a[1], b[1], equal_nan=True) and a[3] == b[3]
    bad = [k for k in a[2] if not np.array_equal(a[2][k], b[2][k], equal_nan=True)]
    print(f"N={N} d={d} tune={tune} draws={n} maxdepth={md}: identical={ok} stats_diff={bad} leapfrogs={a[3]},{b[3]} depth max {a[2]['depth'].max()}")
    if not ok or bad:
        dd = np.argwhere(~np.isclose(a[1], b[1], rtol=0, atol=0, equal_nan=True))
        print("  first diffs", dd[:5], a[2]['depth'][:3, :4], b[2]['depth'][:3, :4])
        sys.exit(1)
print("v2 parity ok")
