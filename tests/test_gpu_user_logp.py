"""NUTS_LOGP_USER: the user-supplied device log density (include/nuts_user_logp.cuh; the device-side form of
CpuLogpFunc::logp, reference src/math/cpu_math.rs:885-970) and its LogpError channel (src/math/math.rs:9-13).

The library is built with the example density nuts_rs_b200/csrc/user_models/diag_gaussian.cuh: the diagonal Gaussian written
against the user interface (params = [mu | 1/sigma^2 | limit]).  It must sample exactly like the built-in diagonal Gaussian and
like the oracle; beyond `limit` it raises a recoverable error (-> divergence), far beyond a fatal one (-> the chain stops)."""
import numpy as np
import pytest

from nuts_rs_b200 import _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _params(mu, sigma, limit=0.0):
    d = len(sigma)
    return np.concatenate([np.full(d, mu), 1.0 / np.asarray(sigma) ** 2, [limit]])


def _run(L, kind, N, d, settings, n_draws, seed, x0, **mk):
    m = L.CudaMath(N, d, kind, **mk)
    s = L.Sampler(m, settings, seed=seed)
    st = s.set_position(x0)
    draws, stats = s.draw(n_draws)
    state = s.chain_state()
    s.close()
    m.close()
    return st, draws, stats, state


@pytest.mark.parametrize("d,N", [(1000, 4), (100, 6), (10, 8), (2048, 2)])
def test_user_density_samples_like_the_builtin_and_the_oracle(L, orc, d, N):
    sigma = np.exp(np.linspace(-1, 1, d))
    s = L.DiagNutsSettings(num_tune=30, maxdepth=6)
    x0 = np.random.default_rng(d).normal(size=(N, d))
    st_u, du, su, _ = _run(L, _abi.NUTS_LOGP_USER, N, d, s, 50, 7, x0, user_params=_params(0.5, sigma))
    st_b, db, sb, _ = _run(L, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, 50, 7, x0, mu=0.5, sigma=sigma)
    assert (st_u == 0).all() and (st_b == 0).all()
    om = orc.Model(_abi.NUTS_LOGP_GAUSS_DIAG, d, mu=0.5, sigma=sigma)
    osamp = orc.Sampler(om, s, seed=7, nchains=N, nthreads=4)
    osamp.set_position(x0)
    do, so = osamp.draw(50)
    # the first draws (before rounding noise is amplified by the adaptation): identical trees, 1e-9 on the draws
    k = 6
    for name in ("depth", "n_steps", "diverging", "index_in_trajectory"):
        np.testing.assert_array_equal(su[name][:k], so[name][:k], err_msg=name)
        np.testing.assert_array_equal(su[name][:k], sb[name][:k], err_msg=name)
    np.testing.assert_allclose(du[:k], do[:k], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(du[:k], db[:k], rtol=1e-9, atol=1e-9)


def test_recoverable_error_is_a_divergence_and_fatal_error_stops_the_chain(L):
    d, N = 20, 8
    sigma = np.ones(d)
    s = L.DiagNutsSettings(num_tune=0, maxdepth=6)
    rng = np.random.default_rng(3)
    # |x0_i| ~ 1: the initial mass matrix (1 / sqrt|grad|, no adaptation in this test) is then close to the identity
    x0 = rng.choice([-1.0, 1.0], size=(N, d)) * rng.uniform(0.7, 1.3, size=(N, d))
    # limit 2.5 sigma: N(0,1) trajectories in 20 dims cross it regularly -> recoverable errors -> divergences, never fatal (25 sigma)
    st, draws, stats, state = _run(L, _abi.NUTS_LOGP_USER, N, d, s, 200, 5, x0, user_params=_params(0.0, sigma, limit=2.5))
    assert (st == 0).all()
    assert stats["diverging"].sum() > 0
    assert np.isfinite(draws).all() and (np.abs(draws) <= 2.5 + 1e-12).all()  # a draw never comes from a leaf that raised the error
    assert state["alive"].all()
    # the same target without the limit has no divergences at all
    _, _, stats0, _ = _run(L, _abi.NUTS_LOGP_USER, N, d, s, 200, 5, x0, user_params=_params(0.0, sigma))
    assert stats0["diverging"].sum() == 0
    # fatal: coordinate 0 beyond 10 x limit.  limit 0.05 => |x_0| > 0.5 is fatal, |x_i| > 0.05 recoverable: chains die quickly
    x1 = np.zeros((N, d)) + 0.01
    x1[:, 1] = 0.02
    st, draws, stats, state = _run(L, _abi.NUTS_LOGP_USER, N, d, s, 50, 5, x1, user_params=_params(0.0, sigma * 10.0, limit=0.05))
    assert (st == 0).all()
    dead = state["alive"] == 0
    assert dead.any()
    for c in np.nonzero(dead)[0]:
        first_nan = int(np.argmax(np.isnan(draws[:, c, 0])))
        assert np.isnan(draws[first_nan:, c]).all() and np.isfinite(draws[:first_nan, c]).all()
        assert (stats["n_steps"][first_nan:, c] == 0).all()
    # an error at the initial point is a bad initial point
    x2 = x0.copy()
    x2[3, 5] = 100.0
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_USER, user_params=_params(0.0, sigma, limit=2.5))
    smp = L.Sampler(m, s, seed=1)
    st = smp.set_position(x2)
    assert st[3] == 3 and (np.delete(st, 3) == 0).all()
    smp.close()
    m.close()


def test_user_density_through_the_math_ops(L):
    """Math::logp_array (Tier 1) with the user density."""
    d, N = 37, 5
    sigma = np.exp(np.linspace(-1, 1, d))
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_USER, user_params=_params(0.5, sigma, limit=50.0))
    x = np.random.default_rng(0).normal(size=(N, d))
    x[2, 4] = 60.0  # beyond the limit: recoverable error -> status 2, NaN logp
    pos, grad = m.from_host(x), m.new_array()
    logp, status = m.logp_array(pos, grad)
    ref = -0.5 * (((x - 0.5) / sigma) ** 2).sum(axis=1)
    ok = np.arange(N) != 2
    np.testing.assert_allclose(logp[ok], ref[ok], rtol=1e-12)
    np.testing.assert_allclose(grad.box_array()[ok], (-(x - 0.5) / sigma**2)[ok], rtol=1e-12)
    assert status[2] == 2 and (status[ok] == 0).all() and np.isnan(logp[2])
    m.close()
