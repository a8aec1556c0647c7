import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nuts_rs_b200 import _abi, lib

def run(name, kind, N, d, num_tune, n_draws, **mk):
    s = lib.DiagNutsSettings(num_tune=num_tune, seed=42)
    x0 = np.random.default_rng(42).normal(size=(N, d))
    t0 = time.time()
    m = lib.CudaMath(N, d, kind, **mk); S = lib.Sampler(m, s, seed=42)
    st = S.set_position(x0)
    S.draw_device(num_tune); tune_ms, _ = S.last_timing(); lf_t, _ = S.counters()
    draws, stats = S.draw(n_draws)
    ms, _ = S.last_timing(); lf, _ = S.counters()
    x = draws[:, st == 0]
    print(f"{name}: N={N} d={d} bad_init={(st!=0).sum()} tune {tune_ms:.0f} ms ({lf_t/tune_ms*1e3:.3g} lf/s) sample {ms:.0f} ms ({(lf-lf_t)/ms*1e3:.3g} lf/s) wall {time.time()-t0:.1f}s"
          f" depth mean {stats['depth'].mean():.2f} max {stats['depth'].max()} div {stats['diverging'].mean():.4f} eps med {np.median(stats['step_size']):.3g} acc {stats['mean_tree_accept'].mean():.3f}")
    S.close(); m.close()
    return x, stats

x, st = run("C2", _abi.NUTS_LOGP_GAUSS_DIAG, 1024, 1000, 400, 50, mu=0.5, sigma=np.exp(np.linspace(-1, 1, 1000)))
z = (x - 0.5) / np.exp(np.linspace(-1, 1, 1000)); print("  z mean %.4f std %.4f" % (z.mean(), z.std()))
x, st = run("C3", _abi.NUTS_LOGP_FUNNEL, 8192, 10, 400, 100, funnel_scale=3.0)
print("  v mean %.3f std %.3f (prior sd 3)" % (x[..., 0].mean(), x[..., 0].std()))
x, st = run("C5/8", _abi.NUTS_LOGP_GAUSS_RANK1, 8192, 100, 400, 50, mu=0.0, rank1_scale=0.5)
print("  mean %.4f var %.4f (1.5) cov offdiag %.3f (0.5)" % (x.mean(), x.var(), np.mean((x[..., 0] * x[..., 1]))))
x, st = run("C4", _abi.NUTS_LOGP_GAUSS_DIAG, 256, 10000, 1000, 20, mu=0.0, sigma=10 ** np.linspace(-3, 3, 10000))
z = x / 10 ** np.linspace(-3, 3, 10000); print("  z mean %.4f std %.4f" % (z.mean(), z.std()))
