"""Pins the oracle's low-rank mass matrix (oracle/nuts_oracle.hpp: apply_lowrank_transform, LowRankInner, DiagMassMatrix::inner)
against the reference's known-answer tests: src/transform/low_rank.rs:437-533 (round trips) and src/transform/mod.rs:383-674
(position / gradient / logdet / adapted density, rank-1 correction, non-zero mean), all to 1e-12; and the restated estimator
(oracle/lowrank_estimator.hpp) against src/transform/adapt/low_rank.rs:354-407 and the product's numpy estimator."""
import math

import numpy as np

from nuts_rs_b200 import _abi


def _gauss(orc, sigma2, mu=0.0):
    return orc.Model(_abi.NUTS_LOGP_GAUSS_DIAG, len(sigma2), mu=mu, sigma=np.sqrt(np.asarray(sigma2, dtype=np.float64)))


def _mass(orc, model, stds, mean, vals, vecs, mu_lr):
    ham = orc.Hamiltonian(model)
    assert ham.set_lowrank_transform(stds, mean, vals, np.asarray(vecs, dtype=np.float64).reshape(len(vals), model.dim), mu_lr)
    return ham


def _round_trip_x(ham, x):
    p = ham.new_point()
    p.set_vec(p.X, x)
    ham.init_from_untransformed(p)
    q = ham.new_point()
    q.set_vec(q.Z, p.vec(p.Z))
    ham.init_from_transformed(q)
    return p, q


def _round_trip_z(ham, z):
    q = ham.new_point()
    q.set_vec(q.Z, z)
    ham.init_from_transformed(q)
    p = ham.new_point()
    p.set_vec(p.X, q.vec(q.X))
    ham.init_from_untransformed(p)
    return q, p


def test_apply_lowrank_transform(orc):
    # src/math/cpu_math.rs:332-377: dest = rhs + U (diag(vals) - I) U^T rhs; zero columns = copy
    rng = np.random.default_rng(0)
    d, r = 17, 3
    q, _ = np.linalg.qr(rng.normal(size=(d, r)))
    vals = np.array([4.0, 0.25, 1.0])
    rhs = rng.normal(size=d)
    expect = rhs + q @ ((vals - 1.0) * (q.T @ rhs))
    np.testing.assert_allclose(orc.apply_lowrank_transform(q.T, vals, rhs), expect, rtol=1e-14, atol=1e-14)
    np.testing.assert_array_equal(orc.apply_lowrank_transform(np.zeros((0, d)), [], rhs), rhs)
    # inverse map: vals -> 1 / vals
    back = orc.apply_lowrank_transform(q.T, 1.0 / vals, orc.apply_lowrank_transform(q.T, vals, rhs))
    np.testing.assert_allclose(back, rhs, rtol=1e-13, atol=1e-13)


def test_diagonal_round_trips(orc):
    # low_rank.rs:437-482: empty vals / vecs = the pure diagonal transformation
    model = _gauss(orc, [1.0, 1.0, 1.0])
    ham = _mass(orc, model, [1.0, 2.0, 3.0], [0.5, -1.0, 2.0], [], np.zeros((0, 3)), np.zeros(3))
    p, q = _round_trip_x(ham, [1.5, -0.3, 4.2])
    np.testing.assert_allclose(q.vec(q.X), [1.5, -0.3, 4.2], atol=1e-12, rtol=0)
    q, p = _round_trip_z(ham, [0.7, -1.1, 0.3])
    np.testing.assert_allclose(p.vec(p.Z), [0.7, -1.1, 0.3], atol=1e-12, rtol=0)


def test_lowrank_round_trips(orc):
    # low_rank.rs:484-533: rank-1 correction along e_1 with eigenvalue 4, non-zero mean and mu
    model = _gauss(orc, [1.0, 1.0, 1.0])
    ham = _mass(orc, model, np.ones(3), [1.0, -0.5, 0.0], [4.0], [[1.0, 0.0, 0.0]], [0.2, -0.1, 0.0])
    p, q = _round_trip_x(ham, [2.0, 0.5, -1.3])
    np.testing.assert_allclose(q.vec(q.X), [2.0, 0.5, -1.3], atol=1e-12, rtol=0)
    q, p = _round_trip_z(ham, [1.0, -0.3, 0.8])
    np.testing.assert_allclose(p.vec(p.Z), [1.0, -0.3, 0.8], atol=1e-12, rtol=0)


def test_lowrank_transform_position_and_gradient(orc):
    # transform/mod.rs:391-456: empty low-rank part on N(0, diag(1, 4, 9))
    sigma2 = [1.0, 4.0, 9.0]
    ham = _mass(orc, _gauss(orc, sigma2), np.sqrt(sigma2), np.zeros(3), [], np.zeros((0, 3)), np.zeros(3))
    p = ham.new_point()
    p.set_vec(p.X, [1.0, 2.0, 3.0])
    ham.init_from_untransformed(p)
    s = p.scalars()
    z = p.vec(p.Z)
    np.testing.assert_allclose(z, [1.0, 1.0, 1.0], atol=1e-12, rtol=0)
    np.testing.assert_allclose(p.vec(p.GZ), [-1.0, -1.0, -1.0], atol=1e-12, rtol=0)
    assert abs(s["logdet"] - sum(-(0.5 * math.log(v)) for v in sigma2)) < 1e-12
    norm = -0.5 * (3 * math.log(math.tau) - sum(math.log(1.0 / v) for v in sigma2))  # MvNormal's constant (mod.rs:112-117)
    assert abs((s["logp"] + norm) - s["logdet"] - (-0.5 * (3 * math.log(math.tau) + float(np.sum(z * z))))) < 1e-12


def test_lowrank_round_trip_of_the_hamiltonian(orc):
    # transform/mod.rs:458-509
    sigma2 = [2.0, 0.5, 3.0]
    ham = _mass(orc, _gauss(orc, sigma2), np.sqrt(sigma2), np.zeros(3), [], np.zeros((0, 3)), np.zeros(3))
    p, q = _round_trip_x(ham, [0.7, -1.2, 3.3])
    np.testing.assert_allclose(q.vec(q.X), [0.7, -1.2, 3.3], atol=1e-12, rtol=0)
    assert abs(p.scalars()["logp"] - q.scalars()["logp"]) < 1e-12
    assert abs(p.scalars()["logdet"] - q.scalars()["logdet"]) < 1e-12


def test_lowrank_with_rank1_correction(orc):
    # transform/mod.rs:511-617: sigma = 1, lambda = [4], u = e_1 is the exact preconditioner of N(0, diag(4, 1, 1))
    sigma2 = [4.0, 1.0, 1.0]
    ham = _mass(orc, _gauss(orc, sigma2), np.ones(3), np.zeros(3), [4.0], [[1.0, 0.0, 0.0]], np.zeros(3))
    p = ham.new_point()
    p.set_vec(p.X, [2.0, 1.0, 1.0])
    ham.init_from_untransformed(p)
    s = p.scalars()
    z = p.vec(p.Z)
    np.testing.assert_allclose(z, [1.0, 1.0, 1.0], atol=1e-12, rtol=0)
    np.testing.assert_allclose(p.vec(p.GZ), [-1.0, -1.0, -1.0], atol=1e-12, rtol=0)
    assert abs(s["logdet"] - (-0.5 * math.log(4.0))) < 1e-12
    norm = -0.5 * (3 * math.log(math.tau) - sum(math.log(1.0 / v) for v in sigma2))
    assert abs((s["logp"] + norm) - s["logdet"] - (-0.5 * (3 * math.log(math.tau) + 3.0))) < 1e-12
    q = ham.new_point()
    q.set_vec(q.Z, z)
    ham.init_from_transformed(q)
    np.testing.assert_allclose(q.vec(q.X), [2.0, 1.0, 1.0], atol=1e-12, rtol=0)
    assert ham.transform()["id"] == 0  # low_rank.rs:139 id = -1, :189 += 1


def test_lowrank_nonzero_mean(orc):
    # transform/mod.rs:619-674
    sigma2, mu = [4.0, 1.0, 9.0], [2.0, -1.0, 3.0]
    ham = _mass(orc, _gauss(orc, sigma2, mu=np.asarray(mu)), np.sqrt(sigma2), mu, [], np.zeros((0, 3)), np.zeros(3))
    x = np.asarray(mu) + np.sqrt(sigma2)
    p, q = _round_trip_x(ham, x)
    np.testing.assert_allclose(p.vec(p.Z), [1.0, 1.0, 1.0], atol=1e-12, rtol=0)
    np.testing.assert_allclose(q.vec(q.X), x, atol=1e-12, rtol=0)


def test_non_finite_update_is_ignored_and_diag_update_drops_the_correction(orc):
    # low_rank.rs:168-173 (non-finite input: return without touching anything); :143-156 (update_from_grad: inner = None)
    model = _gauss(orc, [1.0, 1.0, 1.0])
    ham = _mass(orc, model, np.ones(3), np.zeros(3), [4.0], [[1.0, 0.0, 0.0]], np.zeros(3))
    before = ham.transform()
    assert not ham.set_lowrank_transform([1.0, np.nan, 1.0], np.zeros(3), [4.0], [[1.0, 0.0, 0.0]], np.zeros(3))
    assert not ham.set_lowrank_transform(np.ones(3), np.zeros(3), [np.inf], [[1.0, 0.0, 0.0]], np.zeros(3))
    after = ham.transform()
    assert after["id"] == before["id"] and after["logdet"] == before["logdet"]
    ham.update_diag_grad(np.ones(3), -np.ones(3))
    p = ham.new_point()
    p.set_vec(p.X, [2.0, 1.0, 1.0])
    ham.init_from_untransformed(p)
    t = ham.transform()
    np.testing.assert_allclose(p.vec(p.Z), (np.array([2.0, 1.0, 1.0]) - t["mean"]) * t["inv_stds"], atol=1e-15, rtol=0)


# ------------------------------------------------------------------------------------------------ the estimator
def test_estimator_spd_mean_known_answer(orc):
    # adapt/low_rank.rs:354-381
    out = orc.lowrank_spd_mean(np.diag([1.0, 4.0, 8.0]), np.diag([1.0, 1.0, 0.5]))
    np.testing.assert_allclose(out, np.diag([1.0, 2.0, 4.0]), rtol=1e-10, atol=1e-10)


def test_estimator_estimate_mass_matrix_known_answer(orc):
    # adapt/low_rank.rs:383-407: grads = -draws => every eigenvalue is 1 to 1e-5
    rng = np.random.default_rng(1)
    draws = rng.normal(size=(20, 3))
    vals, vecs = orc.lowrank_estimate_mass_matrix(draws, -draws, 0.0001)
    assert (vals > 0).all() and np.isfinite(vecs).all()
    np.testing.assert_allclose(vals, np.ones(20), rtol=1e-5, atol=1e-5)


def _lowrank_operator(vals, vecs, d):
    return np.eye(d) + vecs.T @ np.diag(vals - 1.0) @ vecs


def test_product_estimator_agrees_with_the_restatement(orc):
    """nuts_rs_b200/lowrank.py (numpy / scipy factorisations) against the C++ restatement (own Jacobi factorisations): eigenvector
    signs, order and the basis of degenerate eigenspaces are conventions, the transformation is not - compare sigma, mean, mu_lr,
    the sorted eigenvalues and I + U (diag(vals) - I) U^T."""
    from nuts_rs_b200 import lowrank

    rng = np.random.default_rng(11)
    for d, n, cutoff in [(10, 40, 1.00001), (30, 12, 2.0), (6, 6, 1.5), (50, 25, 2.0)]:
        cov = np.eye(d) + 0.5 * np.ones((d, d)) + np.diag(rng.uniform(0.0, 3.0, size=d))
        m = rng.normal(size=d)
        x = m + rng.multivariate_normal(np.zeros(d), cov, size=n)
        g = -(x - m) @ np.linalg.inv(cov) + (0.01 * rng.normal(size=(n, d)) if n < d else 0.0)
        a = lowrank.compute_update(x, g, gamma=1e-5, eigval_cutoff=cutoff)
        b = orc.lowrank_compute_update(x, g, gamma=1e-5, eigval_cutoff=cutoff)
        assert a is not None and b is not None
        np.testing.assert_allclose(a[0], b[0], rtol=1e-10)  # stds
        np.testing.assert_allclose(a[1], b[1], rtol=1e-10, atol=1e-12)  # mean
        assert len(a[2]) == len(b[2]), (d, n, a[2], b[2])
        # A window that spans the space (n > d) determines everything: agreement to rounding.  With n <= d the centred window has
        # rank n - 1 and the thin SVD returns one singular vector of singular value ~0 - an arbitrary direction of the null space
        # (LAPACK, faer and the restatement each pick their own; the restatement drops it) - that enters the joint subspace, and the
        # 1 / gamma = 1e5 regularisation gives the geometric mean a condition number of ~1e14: the estimate itself is only defined to
        # ~1e-2 there.
        tol = 1e-9 if n > d else 1e-2
        np.testing.assert_allclose(np.sort(a[2]), np.sort(b[2]), rtol=tol)
        np.testing.assert_allclose(_lowrank_operator(a[2], a[3], d), _lowrank_operator(b[2], b[3], d), rtol=tol, atol=tol)
        np.testing.assert_allclose(a[4], b[4], rtol=tol, atol=tol)  # mean_low_rank


def test_restated_estimator_whitens_an_exact_gaussian(orc):
    # the property behind tests/sample_normal.rs:320-356 (see tests/test_lowrank_estimator.py for the product side)
    rng = np.random.default_rng(2)
    d, n = 10, 40
    cov = np.eye(d) + 0.5 * np.ones((d, d))
    prec = np.linalg.inv(cov)
    x = rng.multivariate_normal(np.zeros(d), cov, size=n)
    stds, mean, vals, vecs, mu = orc.lowrank_compute_update(x, -x @ prec, gamma=1e-5, eigval_cutoff=1.00001)
    a_fwd = np.eye(d) + vecs.T @ np.diag(np.sqrt(vals) - 1.0) @ vecs
    a_inv = np.eye(d) + vecs.T @ np.diag(1.0 / np.sqrt(vals) - 1.0) @ vecs
    for _ in range(5):
        xt = rng.multivariate_normal(np.zeros(d), cov)
        z = a_inv @ ((xt - mean) / stds - mu)
        gz = a_fwd @ ((-prec @ xt) * stds)
        assert np.sum((z + gz) ** 2) < 1e-10
