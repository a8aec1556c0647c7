"""GPU: the low-rank mass matrix at Tier 1 (Math::apply_lowrank_transform, reference src/math/cpu_math.rs:332-425) and Tier 2
(LowRankMassMatrix, reference src/transform/low_rank.rs:97-404 inside Hamiltonian::init_state / initialize_trajectory / leapfrog)
through the C ABI - the reference's known-answer tests on the device (src/transform/low_rank.rs:437-533, src/transform/mod.rs:383-674,
1e-12) and parity with the oracle on random transformations (1e-11 per leapfrog like the diagonal case)."""
import math

import numpy as np
import pytest

from nuts_rs_b200 import _abi
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _orthonormal(rng, d, r):
    q, _ = np.linalg.qr(rng.normal(size=(d, r)))
    return np.ascontiguousarray(q.T)  # [r, d]


@pytest.mark.parametrize("d,r", [(3, 1), (17, 3), (100, 8), (1000, 11), (4097, 20), (300, 64)])
def test_apply_lowrank_transform(L, orc, d, r):
    """dest = rhs + U (diag(vals) - I) U^T rhs per chain; different eigenvectors, eigenvalues and ranks per chain; in place; rank 0."""
    N = 5
    rng = np.random.default_rng(d + r)
    vecs = np.stack([_orthonormal(rng, d, r) for _ in range(N)])
    vals = np.exp(rng.normal(size=(N, r)))
    rank = np.array([r, max(r - 1, 0), 0, r, min(2, r)], dtype=np.int32)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    eigs = m.new_eigs(vecs, vals, rank)
    rhs_h = rng.normal(size=(N, d))
    rhs, dest = m.new_array().read_from_slice(rhs_h), m.new_array()
    m.apply_lowrank_transform(eigs, rhs, dest)
    got = dest.box_array()
    for c in range(N):
        k = rank[c]
        want = orc.apply_lowrank_transform(vecs[c, :k], vals[c, :k], rhs_h[c])
        scale = np.maximum(1.0, np.abs(want))
        assert np.max(np.abs(got[c] - want) / scale) < 1e-13 * max(1, d) ** 0.5, c
        if k == 0:
            np.testing.assert_array_equal(got[c], rhs_h[c])
    np.testing.assert_array_equal(rhs.box_array(), rhs_h)  # rhs untouched
    m.apply_lowrank_transform_inplace(eigs, rhs)
    np.testing.assert_array_equal(rhs.box_array(), got)  # the in-place form is the same computation
    m.close()


def _point_from_x(m, x):
    p, status = m.init_state(x)
    assert (status == 0).all()
    return p


def test_reference_known_answers_on_the_device(L):
    """src/transform/mod.rs:391-456, 511-674 and src/transform/low_rank.rs:437-533, one chain each in ONE batched call."""
    N, d = 4, 3
    # chain 0: N(0, diag(1,4,9)), empty low-rank part; chain 1: N(0, diag(4,1,1)) preconditioned exactly by lambda = 4 along e_1;
    # chain 2: non-zero mean; chain 3: rank-1 correction with mean and mu (round trip only)
    sigma2 = np.array([[1.0, 4.0, 9.0], [4.0, 1.0, 1.0], [4.0, 1.0, 9.0], [1.0, 1.0, 1.0]])
    target_mu = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [2.0, -1.0, 3.0], [0.0, 0.0, 0.0]])
    # (the device model has one mu / sigma for all chains: evaluate chain by chain with four contexts)
    stds = np.array([np.sqrt(sigma2[0]), np.ones(3), np.sqrt(sigma2[2]), np.ones(3)])
    mean = np.array([np.zeros(3), np.zeros(3), target_mu[2], [1.0, -0.5, 0.0]])
    vals = np.array([[1.0], [4.0], [1.0], [4.0]])
    vecs = np.tile(np.array([[[1.0, 0.0, 0.0]]]), (N, 1, 1))
    rank = np.array([0, 1, 0, 1], dtype=np.int32)
    mu_lr = np.array([np.zeros(3), np.zeros(3), np.zeros(3), [0.2, -0.1, 0.0]])
    xs = np.array([[1.0, 2.0, 3.0], [2.0, 1.0, 1.0], target_mu[2] + np.sqrt(sigma2[2]), [2.0, 0.5, -1.3]])
    for c in range(N):
        m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=target_mu[c], sigma=np.sqrt(sigma2[c]))
        ok = m.set_lowrank_transform(stds, mean, vals, vecs, mu_lr, rank)
        assert ok.all()
        t = m.transform()
        assert (t["id"] == 0).all()  # -1 + one update (low_rank.rs:139, 189)
        p = _point_from_x(m, np.tile(xs[c], (N, 1)))
        z, gz, s = p.vec(p.Z)[c], p.vec(p.GZ)[c], p.scalars()
        if c < 3:
            np.testing.assert_allclose(z, [1.0, 1.0, 1.0], atol=1e-12, rtol=0)
        if c < 2:
            np.testing.assert_allclose(gz, [-1.0, -1.0, -1.0], atol=1e-12, rtol=0)
            expected_logdet = sum(-0.5 * math.log(v) for v in sigma2[c])  # chain 1: -1/2 ln 4 comes from lambda instead of sigma
            assert abs(s["logdet"][c] - expected_logdet) < 1e-12
            norm = -0.5 * (d * math.log(math.tau) - sum(math.log(1.0 / v) for v in sigma2[c]))
            assert abs((s["logp"][c] + norm) - s["logdet"][c] - (-0.5 * (d * math.log(math.tau) + float(np.sum(z * z))))) < 1e-12
        # round trip z -> x through one leapfrog of step size 0 is not available; use the forward map of a zero-velocity step:
        # a leapfrog with eps = 0 maps (z, v) to itself and recomputes x = F(z)
        m.initialize_trajectory(p, True, seed=1, chain_offset=0, counter=0)
        nxt, st, _ = m.leapfrog(p, 0.0)
        np.testing.assert_allclose(nxt.vec(nxt.X)[c], xs[c], atol=1e-12, rtol=0)
        np.testing.assert_allclose(nxt.vec(nxt.GZ)[c], gz, atol=1e-12, rtol=0)
        m.close()


@pytest.mark.parametrize("kind,d,r", [(_abi.NUTS_LOGP_GAUSS_DIAG, 50, 4), (_abi.NUTS_LOGP_GAUSS_RANK1, 100, 2), (_abi.NUTS_LOGP_FUNNEL, 12, 3),
                                      (_abi.NUTS_LOGP_GAUSS_DIAG, 1000, 9)])
def test_leapfrog_parity_with_lowrank_transformation(L, orc, kind, d, r):
    """init_state, initialize_trajectory (re-whitening after the transformation changed) and 4 + 4 leapfrogs against the oracle, every
    chain with its own (sigma, mean, U, lambda, mu_lr); one chain of the batch stays diagonal (rank 0)."""
    N = 4
    rng = np.random.default_rng(d * 31 + r)
    kw = {_abi.NUTS_LOGP_GAUSS_DIAG: dict(mu=0.3, sigma=np.exp(0.5 * rng.normal(size=d))), _abi.NUTS_LOGP_GAUSS_RANK1: dict(mu=0.0, rank1_scale=0.5),
          _abi.NUTS_LOGP_FUNNEL: dict(funnel_scale=3.0)}[kind]
    tol = 1e-9 if kind == _abi.NUTS_LOGP_FUNNEL else 1e-11
    m = L.CudaMath(N, d, kind, **kw)
    om = orc.Model(kind, d, **kw)
    stds, mean = np.exp(0.3 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d))
    vecs = np.stack([_orthonormal(rng, d, r) for _ in range(N)])
    vals = np.exp(0.7 * rng.normal(size=(N, r)))
    mu_lr = 0.05 * rng.normal(size=(N, d))
    rank = np.array([r, r - 1, 0, r], dtype=np.int32)
    x0 = rng.normal(size=(N, d))
    # first a diagonal transformation, then the low-rank update: the points are re-whitened by initialize_trajectory
    m.set_transform(stds, mean)
    p, status = m.init_state(x0)
    assert (status == 0).all()
    assert m.set_lowrank_transform(stds * 1.1, mean, vals, vecs, mu_lr, rank).all()
    m.initialize_trajectory(p, True, seed=9, chain_offset=0, counter=3)
    eps = 0.05 + 0.02 * rng.random(N)
    for c in range(N):
        h = orc.Hamiltonian(om)
        h.set_transform(stds[c], mean[c])
        op, ost = h.init_state(x0[c])
        assert ost == 0
        k = rank[c]
        assert h.set_lowrank_transform(stds[c] * 1.1, mean[c], vals[c, :k], vecs[c, :k], mu_lr[c])
        h.initialize_trajectory(op, True, 9, c + 1, 3)
        t, ot = m.transform(), h.transform()
        assert t["id"][c] == ot["id"] == 1
        assert abs(t["logdet"][c] - ot["logdet"]) < 1e-12 * max(1.0, abs(ot["logdet"]))
        assert rel_err(p.vec(p.Z)[c], op.vec(op.Z)) < 1e-12
        assert rel_err(p.vec(p.GZ)[c], op.vec(op.GZ)) < 1e-12
        sc, osc = p.scalars(), op.scalars()
        assert abs(sc["initial_energy"][c] - osc["initial_energy"]) <= 1e-12 * max(1.0, abs(osc["initial_energy"]))
        for direction in (1, -1):
            cur, ocur = p, op
            for step in range(4):
                nxt, st, ee = m.leapfrog(cur, eps, direction=direction)
                onxt, ost2, oee = h.leapfrog(ocur, eps[c], direction)
                assert st[c] == ost2
                for which in range(5):
                    assert rel_err(nxt.vec(which)[c], onxt.vec(which)) < tol, (which, step)
                s1, s2 = nxt.scalars(), onxt.scalars()
                for key in ("logp", "kinetic_energy", "logdet", "initial_energy"):
                    assert abs(s1[key][c] - s2[key]) <= tol * max(1.0, abs(s2[key])), key
                assert abs(ee[c] - oee) <= 1e-9 * max(1.0, abs(oee), abs(s2["initial_energy"]))
                cur, ocur = nxt, onxt
    m.close()


def test_non_finite_update_keeps_the_old_transformation_and_diag_update_drops_the_correction(L):
    """low_rank.rs:168-173 per chain, and update_from_grad / set_transform -> inner = None (low_rank.rs:143-156)."""
    N, d, r = 3, 6, 2
    rng = np.random.default_rng(5)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    stds, mean = np.exp(0.2 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d))
    vecs = np.stack([_orthonormal(rng, d, r) for _ in range(N)])
    vals = np.exp(rng.normal(size=(N, r)))
    mu = np.zeros((N, d))
    assert m.set_lowrank_transform(stds, mean, vals, vecs, mu).all()
    before = m.transform()
    x = rng.normal(size=(N, d))
    z_before = _point_from_x(m, x).vec(2)
    bad_vals = vals.copy()
    bad_vals[1, 0] = np.inf
    ok = m.set_lowrank_transform(stds * 2.0, mean + 1.0, bad_vals * 1.5, vecs, mu + 0.1)
    np.testing.assert_array_equal(ok, [True, False, True])
    after = m.transform()
    np.testing.assert_array_equal(after["id"], before["id"] + np.array([1, 0, 1]))
    assert after["logdet"][1] == before["logdet"][1] and after["logdet"][0] != before["logdet"][0]
    np.testing.assert_array_equal(after["stds"][1], before["stds"][1])
    z_after = _point_from_x(m, x).vec(2)
    np.testing.assert_array_equal(z_after[1], z_before[1])  # chain 1: same transformation as before
    assert not np.array_equal(z_after[0], z_before[0])
    m.set_transform(stds, mean)  # diagonal update: the correction is gone
    z_diag = _point_from_x(m, x).vec(2)
    np.testing.assert_array_equal(z_diag, (x - mean) * (1.0 / stds))
    m.close()


# ------------------------------------------------------------------------------------------------ tier 3: whole draws
def _lr_settings(L, **kw):
    s = L.DiagNutsSettings(**kw)
    return s


def _freeze_mass_matrix(s):
    """device-side adaptation limited to the step size: no window ever switches, no update is ever due"""
    big = 1 << 40
    s.adapt_options.mass_matrix_update_freq = big
    s.adapt_options.early_mass_matrix_switch_freq = big
    s.adapt_options.mass_matrix_switch_freq = big
    return s


@pytest.mark.parametrize("kind,d,r,N", [(_abi.NUTS_LOGP_GAUSS_DIAG, 50, 4, 5), (_abi.NUTS_LOGP_GAUSS_RANK1, 100, 2, 4), (_abi.NUTS_LOGP_FUNNEL, 12, 3, 6),
                                        (_abi.NUTS_LOGP_GAUSS_DIAG, 300, 9, 3), (_abi.NUTS_LOGP_GAUSS_DIAG, 1000, 6, 3)])
@pytest.mark.parametrize("num_tune", [0, 30])
def test_whole_draws_with_a_lowrank_transformation(L, orc, kind, d, r, N, num_tune):
    """Chain::draw x n on an SM_LOWRANK engine against the oracle with the same per-chain transformation (installed after
    set_position: the first update re-runs the step size search on both sides).  num_tune = 0: step size and transformation fixed,
    every draw and statistic to 1e-9 with identical trees; num_tune = 30: dual averaging runs on top (exact-agreement prefix)."""
    rng = np.random.default_rng(d * 7 + r + num_tune)
    kw = {_abi.NUTS_LOGP_GAUSS_DIAG: dict(mu=0.3, sigma=np.exp(0.5 * rng.normal(size=d))), _abi.NUTS_LOGP_GAUSS_RANK1: dict(mu=0.0, rank1_scale=0.5),
          _abi.NUTS_LOGP_FUNNEL: dict(funnel_scale=3.0)}[kind]
    s = _freeze_mass_matrix(L.DiagNutsSettings(num_tune=num_tune, maxdepth=5))
    m = L.CudaMath(N, d, kind, **kw)
    smp = L.Sampler(m, s, seed=5, lowrank_rank_max=16)
    om = orc.Model(kind, d, **kw)
    osmp = orc.Sampler(om, s, seed=5, nchains=N, nthreads=4)
    x0 = rng.normal(size=(N, d))
    if kind == _abi.NUTS_LOGP_FUNNEL:
        x0[:, 0] = 0.1
    st, ost = smp.set_position(x0), osmp.set_position(x0)
    np.testing.assert_array_equal(st, ost)
    stds, mean = np.exp(0.2 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d))
    vecs = np.stack([_orthonormal(rng, d, r) for _ in range(N)])
    vals = np.exp(0.6 * rng.normal(size=(N, r)))
    mu_lr = 0.05 * rng.normal(size=(N, d))
    rank = np.array([r, r - 1, 0] + [r] * (N - 3), dtype=np.int32)
    ok = smp.set_lowrank_transform(stds, mean, vals, vecs, mu_lr, rank)
    ook = osmp.set_lowrank_transform(stds, mean, vals, vecs, mu_lr, rank)
    assert ok.all() and ook.all()
    g0, o0 = smp.state(), osmp.state()
    np.testing.assert_allclose(g0["step_size"], o0["step_size"], rtol=1e-9)  # the re-run step size search agrees
    np.testing.assert_array_equal(g0["rng_counter"], o0["rng_counter"])
    n = 14
    draws, stats = smp.draw(n)
    odraws, ostats = osmp.draw(n)
    strict = n if num_tune == 0 and kind != _abi.NUTS_LOGP_FUNNEL else 3
    for c in range(N):
        same = np.ones(n, dtype=bool)
        for name in ("depth", "n_steps", "diverging", "index_in_trajectory"):
            same &= stats[name][:, c] == ostats[name][:, c]
        common = n if same.all() else int(np.argmin(same))
        assert common >= strict, f"chain {c}: tree shapes differ from draw {common}"
        k = min(strict, common)
        scale = np.maximum(1.0, np.abs(odraws[:k, c]))
        assert np.max(np.abs(draws[:k, c] - odraws[:k, c]) / scale) < 1e-9, c
        for name in ("logp", "energy", "step_size", "mean_tree_accept"):
            a, b = stats[name][:k, c], ostats[name][:k, c]
            assert np.all(np.abs(a - b) <= 1e-9 * np.maximum(1.0, np.abs(b))), (c, name)
    smp.close()
    m.close()


# ------------------------------------------------------------------------------------------------ adaptation (host estimator + GPU)
def test_low_rank_exact_gaussian(L):
    """The reference's end-to-end known-answer test tests/sample_normal.rs:320-356: 10-dim N(0, I + 0.5 11^T), LowRankNutsSettings with
    num_tune = 500, eigval_cutoff = 1.00001, start at x = 1: after the warm-up the transformation whitens the target exactly, so on
    EVERY post-warm-up draw fisher_distance = |z + grad_z|^2 < 1e-10.  Here with 6 chains (own windows, own transformations)."""
    from nuts_rs_b200 import lowrank

    N, d = 6, 10
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_RANK1, mu=0.0, rank1_scale=0.5)
    s = L.DiagNutsSettings(num_tune=500, maxdepth=6)  # LowRankNutsSettings::default(): maxdepth 6 (sampler.rs:636-642)
    smp = lowrank.LowRankSampler(m, s, seed=42, rank_max=10, eigval_cutoff=1.00001)
    status = smp.set_position(np.ones((N, d)))
    assert (status == 0).all()
    draws, stats = smp.draw(600)
    post = stats["tuning"] == 0
    assert post[500:].all() and not post[:500].any()
    fd = stats["fisher_distance"][500:]
    assert np.isfinite(fd).all() and fd.max() < 1e-10, fd.max()
    assert smp.updates >= N * 10 and (smp.last_ranks >= 1).all()
    # and the draws are draws of the target: variance 1.5, covariance 0.5
    x = draws[500:].reshape(-1, d)
    cov = np.cov(x.T)
    assert abs(np.mean(np.diag(cov)) - 1.5) < 0.25 and abs((cov.sum() - np.trace(cov)) / (d * d - d) - 0.5) < 0.25
    assert stats["diverging"][500:].sum() == 0
    # an exactly whitened Gaussian: every post-warm-up tree has the same shape
    assert stats["depth"][500:].max() <= 3
    smp.close()
    m.close()


def test_lowrank_adaptation_beats_diagonal_on_a_correlated_target(L):
    """Config 5's target (rank-1 correlated Gaussian, here d = 100): diagonal adaptation cannot whiten it (fisher distance |z + grad_z|^2
    of order dim), the low-rank transformation does (near zero) at no more leapfrogs per draw."""
    from nuts_rs_b200 import lowrank

    N, d = 8, 100
    kw = dict(mu=0.0, rank1_scale=0.5)
    s = L.DiagNutsSettings(num_tune=300, maxdepth=8)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_RANK1, **kw)
    lr = lowrank.LowRankSampler(m, s, seed=3, rank_max=16)
    x0 = np.random.default_rng(0).normal(size=(N, d))
    assert (lr.set_position(x0) == 0).all()
    _, st_lr = lr.draw(400)
    lr_ranks = lr.last_ranks.copy()
    lr.close()
    diag = L.Sampler(m, s, seed=3)
    assert (diag.set_position(x0) == 0).all()
    _, st_d = diag.draw(400)
    diag.close()
    m.close()
    n_lr, n_d = st_lr["n_steps"][300:].mean(), st_d["n_steps"][300:].mean()
    assert n_lr <= 1.05 * n_d, (n_lr, n_d)
    # the estimator finds the one eigenvalue that is far from 1 (51 = 1 + 0.5 * dim in the target, rescaled by the diagonal part)
    assert (lr_ranks == 1).all(), lr_ranks
    f_lr, f_d = st_lr["fisher_distance"][300:], st_d["fisher_distance"][300:]
    assert np.median(f_lr) < 0.5 * np.median(f_d) and f_lr.max() < 0.2 * f_d.max(), (np.median(f_lr), np.median(f_d), f_lr.max(), f_d.max())


def test_rank_can_grow_between_updates(L, orc):
    """A later update with more eigenvectors than any before (the window grew): the device buffers are re-allocated and the new
    transformation is what the oracle computes; a smaller rank afterwards works too."""
    N, d = 3, 40
    rng = np.random.default_rng(8)
    kw = dict(mu=0.1, sigma=np.exp(0.3 * rng.normal(size=d)))
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, **kw)
    om = orc.Model(_abi.NUTS_LOGP_GAUSS_DIAG, d, **kw)
    x = rng.normal(size=(N, d))
    for r in (2, 7, 3):
        stds, mean = np.exp(0.2 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d))
        vecs = np.stack([_orthonormal(rng, d, r) for _ in range(N)])
        vals, mu = np.exp(0.5 * rng.normal(size=(N, r))), 0.05 * rng.normal(size=(N, d))
        assert m.set_lowrank_transform(stds, mean, vals, vecs, mu).all()
        p, status = m.init_state(x)
        assert (status == 0).all()
        for c in range(N):
            h = orc.Hamiltonian(om)
            assert h.set_lowrank_transform(stds[c], mean[c], vals[c], vecs[c], mu[c])
            op, ost = h.init_state(x[c])
            assert rel_err(p.vec(p.Z)[c], op.vec(op.Z)) < 1e-12 and rel_err(p.vec(p.GZ)[c], op.vec(op.GZ)) < 1e-12
            assert abs(m.transform()["logdet"][c] - h.transform()["logdet"]) < 1e-12 * max(1.0, abs(h.transform()["logdet"]))
    m.close()
