"""BASELINE.json configs 2-5 at full size on one B200, checked through size-independent properties (the oracle cannot run
1024 x 1000-dim chains for 1400 draws inside a test): posterior moments, adapted scales, tree invariants, reproducibility,
stat consistency; plus, for every config, an ORACLE SLICE of the same full-size run: the first 4 chains followed one draw
ahead (teacher forced, tests/test_gpu_teacher_forced.py) through the first draws of the warm-up and again late in the schedule,
every draw and the whole adaptation state within 1e-9 with identical tree shapes."""
import numpy as np
import pytest

from nuts_rs_b200 import _abi
from test_gpu_teacher_forced import teacher_forced

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _run(L, kind, N, d, num_tune, n_draws, seed=42, **mk):
    s = L.DiagNutsSettings(num_tune=num_tune, seed=seed)
    x0 = np.random.default_rng(seed).normal(size=(N, d))
    m = L.CudaMath(N, d, kind, **mk)
    S = L.Sampler(m, s, seed=seed)
    st = S.set_position(x0)
    assert (st == 0).all()
    S.draw(num_tune, want_draws=False, stats=False)
    draws, stats = S.draw(n_draws)
    state = S.state()
    total, done = S.counters()
    S.close()
    m.close()
    return draws, stats, state, s, x0


def _tree_invariants(stats, maxdepth=10):
    depth, n_steps = stats["depth"].astype(np.int64), stats["n_steps"].astype(np.int64)
    assert depth.max() <= maxdepth
    assert (n_steps >= 2 ** depth - 1).all() and (n_steps <= 2 ** (depth + 1) - 1).all()
    md = stats["maxdepth_reached"].astype(bool)
    assert (n_steps[md] == 2 ** maxdepth - 1).all()
    assert (np.abs(stats["index_in_trajectory"]) <= n_steps).all()
    ok = ~stats["diverging"].astype(bool)
    assert ((stats["mean_tree_accept"] >= 0) & (stats["mean_tree_accept"] <= 1)).all()
    # the draw's energy error is finite and below the divergence threshold
    assert np.isfinite(stats["energy_error"][ok]).all() and (stats["energy_error"][ok] <= 1000).all()
    assert not stats["tuning"].any()


def test_config2_1000dim_diag_gaussian_1024_chains(L, orc):
    d, N = 1000, 1024
    sigma = np.exp(np.linspace(-1, 1, d))
    draws, stats, state, settings, x0 = _run(L, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, 400, 40, mu=0.5, sigma=sigma)
    _tree_invariants(stats)
    z = (draws - 0.5) / sigma
    assert abs(z.mean()) < 2e-3 and abs(z.std() - 1) < 2e-3
    # per-coordinate moments over 40 x 1024 draws
    assert np.abs(z.mean(axis=(0, 1))).max() < 0.05 and np.abs(z.std(axis=(0, 1)) - 1).max() < 0.05
    assert stats["diverging"].sum() == 0
    assert abs(stats["mean_tree_accept"].mean() - 0.8) < 0.03  # dual averaging hit target_accept
    # adapted mass matrix: stds^2 = sqrt(var_x / var_grad) = sigma^2 (reference src/math/cpu_math.rs:695)
    ratio = state["stds"] ** 2 / sigma[None, :] ** 2
    assert 0.9 < np.median(ratio) < 1.1 and (ratio > 0.4).all() and (ratio < 2.5).all()
    # oracle on a slice of the SAME run: chains 0..3 (their streams do not depend on the other chains)
    om = orc.Model(_abi.NUTS_LOGP_GAUSS_DIAG, d, mu=0.5, sigma=sigma)
    osamp = orc.Sampler(om, settings, seed=42, nchains=4, nthreads=4)
    osamp.set_position(x0[:4])
    odraws, ostats = osamp.draw(8)
    # re-run the GPU for the first draws of the tuning phase (fresh sampler, same seed)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=sigma)
    S = L.Sampler(m, settings, seed=42)
    S.set_position(x0)
    gdraws, gstats = S.draw(8)
    S.close()
    m.close()
    np.testing.assert_array_equal(gstats["n_steps"][:, :4], ostats["n_steps"])
    np.testing.assert_allclose(gdraws[:1, :4], odraws[:1], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gdraws[:, :4], odraws, rtol=1e-5, atol=1e-5)


def test_config3_funnel_8192_chains(L):
    draws, stats, state, _, _ = _run(L, _abi.NUTS_LOGP_FUNNEL, 8192, 10, 400, 100, funnel_scale=3.0)
    _tree_invariants(stats)
    div = stats["diverging"].astype(bool)
    assert 0.001 < div.mean() < 0.2  # the funnel neck produces divergences
    assert stats["depth"].max() - stats["depth"].min() >= 6  # depth spread: the stress this config is about
    assert stats["maxdepth_reached"].sum() >= 0
    v = draws[..., 0]
    assert np.isfinite(draws).all()
    assert -1.0 < v.mean() < 3.0 and 1.5 < v.std() < 3.5  # N(0, 3^2) truncated by the well-known NUTS funnel bias
    # x_i | v ~ N(0, e^v): standardised x has unit scale where the sampler mixes (v > 0)
    w = draws[..., 1:] * np.exp(-0.5 * v[..., None])
    sel = v > 0
    assert abs(w[sel].std() - 1.0) < 0.1


def test_config4_10000dim_ill_conditioned_256_chains(L):
    d, N = 10000, 256
    sigma = 10.0 ** np.linspace(-3, 3, d)
    draws, stats, state, _, _ = _run(L, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, 1000, 10, mu=0.0, sigma=sigma)
    _tree_invariants(stats)
    z = draws / sigma
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert stats["diverging"].sum() == 0
    # diag mass-matrix tuning recovered 6 orders of magnitude of scale
    ratio = state["stds"] ** 2 / sigma[None, :] ** 2
    assert 0.9 < np.median(ratio) < 1.1 and (ratio > 0.3).all() and (ratio < 3.0).all()
    assert abs(stats["mean_tree_accept"].mean() - 0.8) < 0.05


def test_config5_rank1_correlated_8192_chains_per_gpu(L):
    d, N = 100, 8192
    draws, stats, state, _, _ = _run(L, _abi.NUTS_LOGP_GAUSS_RANK1, N, d, 400, 20, mu=0.0, rank1_scale=0.5)
    _tree_invariants(stats)
    x = draws.reshape(-1, d)
    assert abs(x.mean()) < 5e-3
    cov = np.cov(x[:, :6].T)
    np.testing.assert_allclose(np.diag(cov), 1.5, atol=0.05)  # Sigma = I + 0.5 * 11^T
    np.testing.assert_allclose(cov[np.triu_indices(6, 1)], 0.5, atol=0.05)
    assert stats["diverging"].sum() == 0


def test_reproducible_and_seed_sensitive(L):
    a = _run(L, _abi.NUTS_LOGP_GAUSS_ISO, 64, 10, 50, 20, seed=1, mu=3.0)[0]
    b = _run(L, _abi.NUTS_LOGP_GAUSS_ISO, 64, 10, 50, 20, seed=1, mu=3.0)[0]
    c = _run(L, _abi.NUTS_LOGP_GAUSS_ISO, 64, 10, 50, 20, seed=2, mu=3.0)[0]
    np.testing.assert_array_equal(a, b)  # bitwise reproducible regardless of which SM ran which chain
    assert not np.array_equal(a, c)


def test_leapfrog_counter_matches_n_steps(L):
    s = L.DiagNutsSettings(num_tune=0, seed=3)
    m = L.CudaMath(32, 50, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    S = L.Sampler(m, s, seed=3)
    S.set_position(np.random.default_rng(3).normal(size=(32, 50)))
    before, _ = S.counters()  # leapfrogs of the initial step-size search
    _, stats = S.draw(25)
    after, done = S.counters()
    assert done == 25 and after - before == int(stats["n_steps"].sum())
    S.close()
    m.close()


# ---------------------------------------------------------------------------------------------------------------------------
# Oracle slices at full size: the GPU runs the whole BASELINE config (all chains resident / scheduled as in the bench), the
# oracle follows chains 0..3 one draw ahead.  Two windows per config: the first draws of the warm-up (initial mass matrix, early
# window switches, the step-size re-search) and a window late in the schedule that crosses the end of the tuning phase.
# ---------------------------------------------------------------------------------------------------------------------------
SLICES = {
    "config2": dict(kind=_abi.NUTS_LOGP_GAUSS_DIAG, N=1024, d=1000, tune=400, early=24, late=(388, 24),
                    mk=lambda d: dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))),
    "config3": dict(kind=_abi.NUTS_LOGP_FUNNEL, N=8192, d=10, tune=400, early=60, late=(380, 60), rtol=1e-8, knife=4,
                    mk=lambda d: dict(funnel_scale=3.0)),
    "config4": dict(kind=_abi.NUTS_LOGP_GAUSS_DIAG, N=256, d=10000, tune=1000, early=16, late=(990, 16),
                    mk=lambda d: dict(mu=0.0, sigma=10.0 ** np.linspace(-3, 3, d))),
    "config5": dict(kind=_abi.NUTS_LOGP_GAUSS_RANK1, N=8192, d=100, tune=400, early=24, late=(388, 24),
                    mk=lambda d: dict(mu=0.0, rank1_scale=0.5)),
}


@pytest.mark.parametrize("name", sorted(SLICES))
def test_oracle_slice_of_full_size_config(L, orc, name):
    c = SLICES[name]
    s = L.DiagNutsSettings(num_tune=c["tune"], seed=42)
    kw = dict(rtol=c.get("rtol", 1e-9), max_knife_edge=c.get("knife", 1), oracle_chains=4)
    mk = c["mk"](c["d"])
    teacher_forced(L, orc, c["kind"], c["N"], c["d"], s, c["early"], 42, mk, **kw)
    skip, n = c["late"]
    stats = teacher_forced(L, orc, c["kind"], c["N"], c["d"], s, n, 42, mk, skip_draws=skip, **kw)
    k = c["tune"] - skip
    assert stats["tuning"][:k].all() and not stats["tuning"][k:].any()
