"""Low-rank vs diagonal adaptation on BASELINE config 5's target (rank-1 correlated Gaussian): leapfrogs per draw, step size and
fisher distance |z + grad_z|^2 after the warm-up.   python tools/lowrank_demo.py [dim] [chains] [num_tune]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import numpy as np
from nuts_rs_b200 import lib, _abi, lowrank

d = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
s = lib.DiagNutsSettings(num_tune=T, maxdepth=8)
m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_RANK1, mu=0.0, rank1_scale=0.5)
x0 = np.random.default_rng(0).normal(size=(N, d))
for name in ("lowrank", "diag"):
    smp = lowrank.LowRankSampler(m, s, seed=3, rank_max=16) if name == "lowrank" else lib.Sampler(m, s, seed=3)
    smp.set_position(x0)
    t0 = time.perf_counter()
    _, st = smp.draw(T + 100)
    dt = time.perf_counter() - t0
    post = slice(T, None)
    extra = f" updates {smp.updates} ranks {smp.last_ranks.tolist()}" if name == "lowrank" else ""
    print(f"{name:8s} d={d} N={N}: n_steps {st['n_steps'][post].mean():.2f} step {np.median(st['step_size'][post]):.3f} "
          f"fisher median {np.median(st['fisher_distance'][post]):.3e} max {st['fisher_distance'][post].max():.3e} wall {dt:.2f} s{extra}")
    smp.close()
m.close()
