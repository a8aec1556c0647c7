"""Small workloads for compute-sanitizer (racecheck / memcheck): they exercise the hand-rolled synchronisation of the engines -
draw migration between teams (more chains than resident teams: release / acquire on EngineParams::done), the CTA-team
reductions (double-buffered shared scratch, one barrier), the decoupled engine's flag protocol.

  compute-sanitizer --tool racecheck python tools/sanitize_run.py <case>
cases: c1 (warp teams, 4 chains x d=10), migrate (d=1000, 64x16 CTA teams, more chains than one wave of teams via NUTS_B200_GRID),
       large (large-dim engine, d=5000), funnel, rank1, cluster (rank-1 at d=4200: one chain on the 4 CTAs of a cluster, DSMEM reductions)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nuts_rs_b200 import _abi, lib

case = sys.argv[1]
shapes = {
    "c1": (_abi.NUTS_LOGP_GAUSS_ISO, 4, 10, dict(mu=3.0), 12, 8),
    "migrate": (_abi.NUTS_LOGP_GAUSS_DIAG, 40, 1000, dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, 1000))), 6, 4),
    "large": (_abi.NUTS_LOGP_GAUSS_DIAG, 3, 5000, dict(mu=0.0, sigma=np.exp(np.linspace(-1, 1, 5000))), 6, 3),
    "funnel": (_abi.NUTS_LOGP_FUNNEL, 64, 10, dict(funnel_scale=3.0), 10, 6),
    "rank1": (_abi.NUTS_LOGP_GAUSS_RANK1, 24, 100, dict(mu=0.0, rank1_scale=0.5), 8, 5),
    "cluster": (_abi.NUTS_LOGP_GAUSS_RANK1, 3, 4200, dict(mu=0.0, rank1_scale=0.5), 4, 3),
}
kind, N, d, mk, tune, maxdepth = shapes[case]
s = lib.DiagNutsSettings(num_tune=tune, maxdepth=maxdepth)
m = lib.CudaMath(N, d, kind, **mk)
S = lib.Sampler(m, s, seed=3)
st = S.set_position(np.random.default_rng(3).normal(size=(N, d)))
draws, stats = S.draw(tune + 4)
print(case, "ok: leapfrogs", int(stats["n_steps"].sum()), "finite", bool(np.isfinite(draws[:, st == 0]).all()))
S.close()
m.close()
