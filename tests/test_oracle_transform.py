"""Pins the oracle's DiagMassMatrix against the reference's known-answer tests
(reference src/transform/mod.rs:175-377, all to 1e-12) and its update rules (src/transform/diagonal.rs:85-162)."""
import math

import numpy as np

from nuts_rs_b200 import _abi


def _gauss(orc, sigma2, mu=0.0):
    # make_math(precision): N(mu, diag(sigma2)) log density up to a constant
    return orc.Model(_abi.NUTS_LOGP_GAUSS_DIAG, len(sigma2), mu=mu, sigma=np.sqrt(np.asarray(sigma2, dtype=np.float64)))


def _exact_mass(orc, sigma2, draw_mean=None):
    model = _gauss(orc, sigma2)
    ham = orc.Hamiltonian(model)
    d = len(sigma2)
    sigma2 = np.asarray(sigma2, dtype=np.float64)
    ham.update_diag_draw_grad(np.zeros(d) if draw_mean is None else draw_mean, np.zeros(d), sigma2, 1.0 / sigma2, None, (1e-20, 1e20))
    return ham


def test_diag_transform_position_and_gradient(orc):
    # transform/mod.rs:175-250
    sigma2 = [1.0, 4.0, 9.0]
    ham = _exact_mass(orc, sigma2)
    p = ham.new_point()
    p.set_vec(p.X, [1.0, 2.0, 3.0])
    ham.init_from_untransformed(p)
    s = p.scalars()
    np.testing.assert_allclose(p.vec(p.Z), [1.0, 1.0, 1.0], atol=1e-12, rtol=0)
    np.testing.assert_allclose(p.vec(p.GZ), [-1.0, -1.0, -1.0], atol=1e-12, rtol=0)
    expected_logdet = sum(-(0.5 * math.log(v)) for v in sigma2)
    assert abs(s["logdet"] - expected_logdet) < 1e-12
    z = p.vec(p.Z)
    # transform/mod.rs:240-249: logp_adapted = logp - logdet equals the FULLY NORMALISED standard-normal density at z.  The
    # reference's test target (MvNormal, transform/mod.rs:104-117) carries its normalising constant
    # norm = -0.5 * (d * ln(2 pi) - ln det P); our device / oracle targets are un-normalised (test_logps.rs:49-58), so add it here.
    d = len(sigma2)
    log_det_p = sum(math.log(1.0 / v) for v in sigma2)
    norm = -0.5 * (d * math.log(math.tau) - log_det_p)
    standard_normal_logp = -0.5 * (d * math.log(math.tau) + float(np.sum(z * z)))  # transform/mod.rs:155-158
    logp_adapted = (s["logp"] + norm) - s["logdet"]
    assert abs(logp_adapted - standard_normal_logp) < 1e-12, (logp_adapted, standard_normal_logp)
    # and without the constants: logp(x) = -sum x^2 / (2 sigma2) = -1.5 = -|z|^2 / 2
    assert abs(s["logp"] - (-0.5 * float(np.sum(z * z)))) < 1e-12
    assert ham.transform()["id"] == 0  # -1 + one update (diagonal.rs:81,130)


def test_diag_round_trip(orc):
    # transform/mod.rs:253-318
    sigma2 = [2.0, 0.5, 3.0]
    ham = _exact_mass(orc, sigma2)
    x_orig = [1.5, -0.3, 2.1]
    p = ham.new_point()
    p.set_vec(p.X, x_orig)
    ham.init_from_untransformed(p)
    fwd = p.scalars()
    q = ham.new_point()
    q.set_vec(q.Z, p.vec(p.Z))
    ham.init_from_transformed(q)
    inv = q.scalars()
    np.testing.assert_allclose(q.vec(q.X), x_orig, atol=1e-12, rtol=0)
    assert abs(fwd["logp"] - inv["logp"]) < 1e-12
    assert abs(fwd["logdet"] - inv["logdet"]) < 1e-12
    np.testing.assert_allclose(q.vec(q.GZ), p.vec(p.GZ), atol=1e-12, rtol=0)


def test_diag_nonzero_mean(orc):
    # transform/mod.rs:321-377
    sigma2 = np.array([4.0, 1.0, 9.0])
    mu = np.array([3.0, -1.0, 2.0])
    ham = _exact_mass(orc, sigma2, draw_mean=mu)
    x = mu + np.sqrt(sigma2)
    p = ham.new_point()
    p.set_vec(p.X, x)
    ham.init_from_untransformed(p)
    np.testing.assert_allclose(p.vec(p.Z), [1.0, 1.0, 1.0], atol=1e-12, rtol=0)


def test_update_diag_draw_grad_rules(orc):
    # src/math/cpu_math.rs:671-708: sigma = (var_x/var_g)^(1/4); invalid ratios leave the entry untouched (fill None)
    model = _gauss(orc, [1.0, 1.0, 1.0, 1.0])
    ham = orc.Hamiltonian(model)
    ham.set_transform([2.0, 2.0, 2.0, 2.0], [0.0, 0.0, 0.0, 0.0])
    draw_var = np.array([16.0, 0.0, np.inf, 1e-100])
    grad_var = np.array([1.0, 1.0, 1.0, 1e100])
    draw_mean = np.array([1.0, 2.0, 3.0, 4.0])
    grad_mean = np.array([0.5, 0.5, 0.5, 0.5])
    ham.update_diag_draw_grad(draw_mean, grad_mean, draw_var, grad_var)
    t = ham.transform()
    # entry0: val = sqrt(16) = 4 -> std 2, inv_std 0.5 ; entry1: val=0 -> unchanged (2, 0.5) ; entry2: inf -> unchanged
    # entry3: val = 1e-100 -> clamp 1e-20 -> std 1e-10, inv_std 1e10
    np.testing.assert_allclose(t["stds"], [2.0, 2.0, 2.0, 1e-10], rtol=1e-15)
    np.testing.assert_allclose(t["inv_stds"], [0.5, 0.5, 0.5, 1e10], rtol=1e-15)
    np.testing.assert_allclose(t["mean"], t["stds"] ** 2 * grad_mean + draw_mean, rtol=1e-15)
    assert abs(t["logdet"] - float(np.sum(np.log(t["inv_stds"])))) < 1e-12
    assert t["id"] == 1


def test_update_diag_grad_rules(orc):
    # src/math/cpu_math.rs:710-738 + transform/diagonal.rs:133-154 (initialisation from the first gradient)
    model = _gauss(orc, [1.0, 1.0, 1.0])
    ham = orc.Hamiltonian(model)
    pos = np.array([1.0, -2.0, 0.5])
    grad = np.array([-4.0, 0.0, 1e30])
    ham.update_diag_grad(pos, grad, 1.0, (1e-20, 1e20))
    t = ham.transform()
    val = np.array([0.25, 1e20, 1e-20])  # 1/clamp(|g|)
    np.testing.assert_allclose(t["stds"], np.sqrt(val), rtol=1e-15)
    np.testing.assert_allclose(t["inv_stds"], np.sqrt(1.0 / val), rtol=1e-15)
    np.testing.assert_allclose(t["mean"], pos + t["stds"] ** 2 * grad, rtol=1e-15)
    assert t["id"] == 0


def test_set_transform(orc):
    model = _gauss(orc, [1.0, 1.0])
    ham = orc.Hamiltonian(model)
    ham.set_transform([2.0, 4.0], [1.0, -1.0])
    t = ham.transform()
    np.testing.assert_array_equal(t["inv_stds"], [0.5, 0.25])
    assert abs(t["logdet"] - (math.log(0.5) + math.log(0.25))) < 1e-15
