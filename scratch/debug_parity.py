import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nuts_rs_b200 import _abi, lib as L
from oracle import oracle as O

def run(kind, N, d, n_draws, seed=42, x0=None, thresh=1e-9, **kw):
    print("=== kind", kind, "N", N, "d", d, kw)
    s = L.DiagNutsSettings(**kw)
    mk = {0: dict(mu=3.0), 1: dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, d))), 2: dict(mu=0.0, rank1_scale=0.5), 3: dict(funnel_scale=3.0)}[kind]
    if x0 is None: x0 = np.random.default_rng(seed).normal(size=(N, d))
    m = L.CudaMath(N, d, kind, **mk); S = L.Sampler(m, s, seed=seed); S.set_position(x0)
    om = O.Model(kind, d, **mk); OS = O.Sampler(om, s, seed=seed, nchains=N); OS.set_position(x0)
    dr, st = S.draw(n_draws); odr, ost = OS.draw(n_draws)
    err = np.abs(dr - odr).max(axis=2) / np.maximum(1, np.abs(odr).max(axis=2))
    last = 1e-16
    for t in range(n_draws):
        flags = []
        for name in ("depth", "n_steps", "index_in_trajectory", "diverging"):
            bad = np.nonzero(st[name][t] != ost[name][t])[0]
            if len(bad): flags.append((name, bad.tolist(), st[name][t][bad].tolist(), ost[name][t][bad].tolist()))
        c = int(np.argmax(err[t]))
        if err[t].max() > 10 * last or flags:
            last = max(err[t].max(), 1e-16)
            print("t=%d maxerr %.2e chain %d depth %d/%d nsteps %d idx %d eps %.3g acc %.3f |x|max %.3g  flags=%s" % (t, err[t].max(), c, st["depth"][t][c], ost["depth"][t][c], st["n_steps"][t][c], st["index_in_trajectory"][t][c], st["step_size"][t][c], st["mean_tree_accept"][t][c], np.abs(odr[t][c]).max(), flags))
            if flags: break
    S.close(); m.close()

run(3, 16, 10, 60, seed=5, num_tune=100, maxdepth=8)
run(3, 16, 10, 60, seed=6, num_tune=0, maxdepth=8)
run(1, 2, 1025, 40, seed=1025, num_tune=20, maxdepth=6)
run(1, 2, 10000, 20, seed=10000, num_tune=10, maxdepth=6)
run(1, 2, 2048, 30, seed=2048, num_tune=15, maxdepth=6)
