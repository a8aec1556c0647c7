"""Small workloads for compute-sanitizer (racecheck / memcheck): they exercise the hand-rolled synchronisation of the engines -
draw migration between teams (more chains than resident teams: release / acquire on EngineParams::done), the CTA-team
reductions (double-buffered shared scratch, one barrier), the decoupled engine's flag protocol.

  compute-sanitizer --tool racecheck python tools/sanitize_run.py <case>
cases: c1 (warp teams, 4 chains x d=10), migrate (d=1000, 64x16 CTA teams, more chains than one wave of teams via NUTS_B200_GRID),
       large (large-dim engine, d=5000), funnel, rank1, cluster (rank-1 at d=4200: one chain on the 4 CTAs of a cluster, DSMEM reductions),
       tunebuild (640 chains x d=1000 on the full grid: aligned warm-up build, then the plain one), lowrank (SM_LOWRANK engine),
       plane (Tier-2 plane kernels: TMA-staged leapfrog, ExactNormal / Microcanonical leapfrogs, ESH update, flows)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nuts_rs_b200 import _abi, lib

case = sys.argv[1]
if case == "plane":
    # Tier 1 / Tier 2 plane kernels added last: the TMA-staged leapfrog (cp.async.bulk + mbarrier; one / two / four chunks, i.e. with
    # stage refills and an odd tail) and the ExactNormal / Microcanonical kernels (ESH: three team reductions per update)
    total = 0
    for d in (3, 513, 1537):
        N = 5
        rng = np.random.default_rng(d)
        m = lib.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
        m.set_transform(np.exp(0.3 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d)))
        p, st = m.init_state(rng.normal(size=(N, d)))
        for kind in (_abi.NUTS_KINETIC_EUCLIDEAN, _abi.NUTS_KINETIC_EXACT_NORMAL, _abi.NUTS_KINETIC_MICROCANONICAL):
            m.initialize_trajectory(p, True, 9, 0, 3, kind=kind)
            cur = p
            for step in range(3):
                cur, status, ee = m.leapfrog(cur, 0.01, direction=np.array([1, -1, 1, -1, 1], dtype=np.int8), kind=kind,
                                             active=np.array([1, 1, 0, 1, 1], dtype=np.uint8) if step == 1 else None)
                total += int((status == 0).sum())
        a, b, c = m.from_host(rng.normal(size=(N, d))), m.from_host(rng.normal(size=(N, d))), m.new_array()
        m.std_norm_flow(a, c, b, 0.3)
        m.std_norm_grad_flow(a, b, c, c, 0.1)
        m.array_normalize(b)
        dke = m.esh_momentum_update(a, b, 0.05)
        assert np.isfinite(dke).all() and np.isfinite(b.box_array()).all()
        m.close()
    print("plane ok: leapfrogs", total, "finite True")
    sys.exit(0)
shapes = {
    "c1": (_abi.NUTS_LOGP_GAUSS_ISO, 4, 10, dict(mu=3.0), 12, 8),
    "migrate": (_abi.NUTS_LOGP_GAUSS_DIAG, 40, 1000, dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, 1000))), 6, 4),
    "large": (_abi.NUTS_LOGP_GAUSS_DIAG, 3, 5000, dict(mu=0.0, sigma=np.exp(np.linspace(-1, 1, 5000))), 6, 3),
    "funnel": (_abi.NUTS_LOGP_FUNNEL, 64, 10, dict(funnel_scale=3.0), 10, 6),
    "rank1": (_abi.NUTS_LOGP_GAUSS_RANK1, 24, 100, dict(mu=0.0, rank1_scale=0.5), 8, 5),
    "cluster": (_abi.NUTS_LOGP_GAUSS_RANK1, 3, 4200, dict(mu=0.0, rank1_scale=0.5), 4, 3),
    # full grid (no NUTS_B200_GRID): more chains than the 592 resident teams, so the warm-up launch runs the aligned four-team build
    # (named barriers per team + the CTA alignment barrier) and hands over to the plain build afterwards
    "tunebuild": (_abi.NUTS_LOGP_GAUSS_DIAG, 640, 1000, dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, 1000))), 3, 3),
    "lowrank": (_abi.NUTS_LOGP_GAUSS_RANK1, 12, 100, dict(mu=0.0, rank1_scale=0.5), 4, 4),
}
kind, N, d, mk, tune, maxdepth = shapes[case]
if case == "tunebuild":
    os.environ.pop("NUTS_B200_GRID", None)
s = lib.DiagNutsSettings(num_tune=tune, maxdepth=maxdepth)
m = lib.CudaMath(N, d, kind, **mk)
S = lib.Sampler(m, s, seed=3, lowrank_rank_max=4 if case == "lowrank" else 0)
st = S.set_position(np.random.default_rng(3).normal(size=(N, d)))
if case == "lowrank":  # a rank-2 transformation on the SM_LOWRANK engine (eigenvector reductions inside the leapfrog, kernel mode 2)
    rng = np.random.default_rng(4)
    vecs = np.stack([np.ascontiguousarray(np.linalg.qr(rng.normal(size=(d, 2)))[0].T) for _ in range(N)])
    S.set_lowrank_transform(np.ones((N, d)), np.zeros((N, d)), np.exp(rng.normal(size=(N, 2))), vecs, np.zeros((N, d)))
if case == "tunebuild":
    d1, s1 = S.draw(tune)       # the warm-up launch: aligned build
    d2, s2 = S.draw(2)          # plain build
    draws, stats = np.concatenate([d1, d2]), {k: np.concatenate([s1[k], s2[k]]) for k in s1}
else:
    draws, stats = S.draw(tune + 4)
print(case, "ok: leapfrogs", int(stats["n_steps"].sum()), "finite", bool(np.isfinite(draws[:, st == 0]).all()))
S.close()
m.close()
