"""Pins the oracle's low-rank mass matrix (oracle/nuts_oracle.hpp: apply_lowrank_transform, LowRankInner, DiagMassMatrix::inner)
against the reference's known-answer tests: src/transform/low_rank.rs:437-533 (round trips) and src/transform/mod.rs:383-674
(position / gradient / logdet / adapted density, rank-1 correction, non-zero mean), all to 1e-12."""
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
