"""Pins the oracle's restatement of the geodesic / isokinetic integrator pieces (KineticEnergyKind::ExactNormal / Microcanonical):
the reference's own property tests for std_norm_flow / std_norm_grad_flow(_inplace) (src/math/util.rs:808-877: 32-ULP equality with
the scalar formulas, proptest f64::ANY inputs, epsilon in -10..10), the published ESH update (src/math/cpu_math.rs:505-551, Steeg &
Gallagher arXiv:2111.02434) restated in numpy, and integrator properties of Hamiltonian::leapfrog for the three kinds
(src/dynamics/transformed_hamiltonian.rs:160-258, 524-615): exactness on the standard normal, reversibility, unit-sphere momentum,
energy conservation to O(eps^2), and whole NUTS draws of the oracle with trajectory_kind = ExactNormal (sampler.rs:232, :757)."""
import math

import numpy as np
import pytest

from helpers import any_f64, assert_approx_eq, exact_fma, rel_err
from nuts_rs_b200 import _abi

NCASES = 200
KINDS = [_abi.NUTS_KINETIC_EUCLIDEAN, _abi.NUTS_KINETIC_EXACT_NORMAL, _abi.NUTS_KINETIC_MICROCANONICAL]


def test_std_norm_flow(orc):
    # util.rs:810-837: reference = p * cos + v * sin ; -(p) * sin + v * cos
    rng = np.random.default_rng(11)
    for _ in range(NCASES):
        n = int(rng.integers(0, 32))
        pos, vel = any_f64(rng, n), any_f64(rng, n)
        eps = float(rng.uniform(-10, 10))
        po, vo = orc.std_norm_flow(pos, vel, eps)
        s, c = math.sin(eps), math.cos(eps)
        with np.errstate(all="ignore"):
            want_p = pos * c + vel * s
            want_v = -pos * s + vel * c
        for a, b in zip(po, want_p):
            assert_approx_eq(a, b)
        for a, b in zip(vo, want_v):
            assert_approx_eq(a, b)


def test_std_norm_flow_is_a_rotation(orc):
    rng = np.random.default_rng(12)
    for n in (1, 3, 4, 5, 17, 100):
        pos, vel = rng.normal(size=n), rng.normal(size=n)
        po, vo = orc.std_norm_flow(pos, vel, 0.7)
        np.testing.assert_allclose(po**2 + vo**2, pos**2 + vel**2, rtol=1e-14)
        back_p, back_v = orc.std_norm_flow(po, vo, -0.7)
        np.testing.assert_allclose(back_p, pos, rtol=0, atol=1e-14)
        np.testing.assert_allclose(back_v, vel, rtol=0, atol=1e-14)
        # whole SIMD registers fuse the multiply-add, the last n % 4 elements do not (util.rs:541-572)
        s, c = math.sin(0.7), math.cos(0.7)
        body = (n // 4) * 4
        for i in range(n):
            if i < body:
                assert po[i] == exact_fma(pos[i], c, vel[i] * s)
                assert vo[i] == exact_fma(pos[i], -s, vel[i] * c)
            else:
                assert po[i] == pos[i] * c + vel[i] * s
                assert vo[i] == pos[i] * (-s) + vel[i] * c


@pytest.mark.parametrize("inplace", [False, True])
def test_std_norm_grad_flow(orc, inplace):
    # util.rs:839-877: reference = epsilon.mul_add(p + g, v)
    rng = np.random.default_rng(13 + inplace)
    for _ in range(NCASES):
        n = int(rng.integers(0, 32))
        pos, grad, vel = any_f64(rng, n), any_f64(rng, n), any_f64(rng, n)
        eps = float(rng.uniform(-10, 10))
        out = orc.std_norm_grad_flow(pos, grad, vel, eps, inplace=inplace)
        with np.errstate(all="ignore"):
            pg = pos + grad
        for o, a, v in zip(out, pg, vel):
            assert_approx_eq(o, exact_fma(eps, a, v))


def _esh_numpy(g, p, step):
    """cpu_math.rs:505-551 restated with numpy (pairwise sums instead of sequential ones: compare to 1e-13)."""
    n = g.size
    gn = math.sqrt(float(np.sum(g * g)))
    e = g / gn
    ue = float(np.dot(p, e))
    delta = step * gn / (n - 1)
    zeta = math.exp(-delta)
    raw = e * (1 - zeta) * (1 + zeta + ue * (1 - zeta)) + 2 * zeta * p
    new = raw / math.sqrt(float(np.sum(raw * raw)))
    dke = (delta - math.log(2) + math.log1p(ue + (1 - ue) * zeta * zeta)) * (n - 1)
    return new, dke


def test_esh_momentum_update(orc):
    rng = np.random.default_rng(15)
    for _ in range(NCASES):
        n = int(rng.integers(2, 40))
        g = rng.normal(size=n) * 10 ** rng.uniform(-2, 2)
        p = orc.array_normalize(rng.normal(size=n))
        assert abs(np.sum(p * p) - 1) < 1e-14
        step = float(rng.uniform(-1, 1))
        new, dke = orc.esh_momentum_update(g, p, step)
        want, want_dke = _esh_numpy(g, p, step)
        assert rel_err(new, want) < 1e-12
        assert abs(dke - want_dke) <= 1e-12 * max(1.0, abs(want_dke))
        assert abs(np.sum(new * new) - 1) < 1e-14  # stays on the unit sphere
        # the update is the exact flow of the momentum equation: a step back undoes it
        back, dke_back = orc.esh_momentum_update(g, new, -step)
        np.testing.assert_allclose(back, p, atol=1e-11 * max(1.0, math.exp(2 * abs(step) * np.linalg.norm(g) / (n - 1))))
        if abs(step) * np.linalg.norm(g) / (n - 1) < 2.0:  # beyond that 1 - p.e of the way back cancels catastrophically
            assert abs(dke + dke_back) <= 1e-9 * max(1.0, abs(dke))


def _start(orc, kind, d, model_kind=_abi.NUTS_LOGP_GAUSS_DIAG, seed=3, **kw):
    rng = np.random.default_rng(seed)
    if model_kind == _abi.NUTS_LOGP_GAUSS_DIAG:
        kw.setdefault("mu", 0.5)
        kw.setdefault("sigma", np.exp(np.linspace(-1, 1, d)))
    om = orc.Model(model_kind, d, **kw)
    h = orc.Hamiltonian(om)
    h.set_kinetic_energy_kind(kind)
    h.set_transform(np.exp(0.3 * rng.normal(size=d)), 0.1 * rng.normal(size=d))
    p, st = h.init_state(rng.normal(size=d))
    assert st == 0
    h.initialize_trajectory(p, True, 9, 1, 3)
    return om, h, p


def test_exact_normal_is_exact_on_the_standard_normal(orc):
    """KineticEnergyKind::ExactNormal docs (transformed_hamiltonian.rs:5-6, 28-36): the leapfrog is exact when the transformed
    posterior is a standard normal, at ANY step size - the energy error is rounding noise and the flow is a rotation."""
    d = 50
    om = orc.Model(_abi.NUTS_LOGP_GAUSS_ISO, d, mu=0.0)
    h = orc.Hamiltonian(om)
    h.set_kinetic_energy_kind(_abi.NUTS_KINETIC_EXACT_NORMAL)
    h.set_transform(np.ones(d), np.zeros(d))
    x0 = np.random.default_rng(1).normal(size=d)
    p, _ = h.init_state(x0)
    h.initialize_trajectory(p, True, 5, 1, 0)
    v0 = p.vec(p.V)
    for eps in (0.1, 1.0, 2.5, 7.0):
        cur = p
        for k in range(1, 6):
            cur, st, ee = h.leapfrog(cur, eps, 1)
            assert st == 0 and abs(ee) < 1e-11
            np.testing.assert_allclose(cur.vec(cur.Z), x0 * math.cos(k * eps) + v0 * math.sin(k * eps), atol=1e-12)
            np.testing.assert_allclose(cur.vec(cur.V), -x0 * math.sin(k * eps) + v0 * math.cos(k * eps), atol=1e-12)
    # the Euclidean leapfrog at the same step is far from exact
    h.set_kinetic_energy_kind(_abi.NUTS_KINETIC_EUCLIDEAN)
    _, _, ee = h.leapfrog(p, 1.0, 1)
    assert abs(ee) > 1e-3


@pytest.mark.parametrize("kind", KINDS)
def test_leapfrog_is_reversible_and_second_order(orc, kind):
    d = 20
    om, h, p = _start(orc, kind, d)
    eps = 0.05 if kind != _abi.NUTS_KINETIC_MICROCANONICAL else 0.05 / math.sqrt(d)  # :214-219 the step is scaled by sqrt(dim)
    cur = p
    for _ in range(8):
        cur, st, ee = h.leapfrog(cur, eps, 1)
        assert st == 0
    errs = [abs(ee)]
    sc = cur.scalars()
    assert sc["index_in_trajectory"] == 8
    if kind == _abi.NUTS_KINETIC_MICROCANONICAL:
        v = cur.vec(cur.V)
        assert abs(np.sum(v * v) - 1) < 1e-13  # :186-198 momentum stays on the unit sphere
    back = cur
    for _ in range(8):
        back, st, _ = h.leapfrog(back, eps, -1)
    assert back.scalars()["index_in_trajectory"] == 0
    np.testing.assert_allclose(back.vec(back.Z), p.vec(p.Z), atol=1e-11)
    np.testing.assert_allclose(back.vec(back.V), p.vec(p.V), atol=1e-11)
    if kind == _abi.NUTS_KINETIC_MICROCANONICAL:
        assert abs(back.scalars()["kinetic_energy"]) < 1e-10  # accumulated delta KE returns to 0
    # halving the step at fixed integration time divides the energy error by ~4
    cur = p
    for _ in range(16):
        cur, st, ee = h.leapfrog(cur, eps / 2, 1)
    errs.append(abs(ee))
    assert errs[1] < errs[0] / 2.5, errs


def test_microcanonical_divergence_is_two_sided(orc):
    """transformed_hamiltonian.rs:591-596: |energy error| >= max_energy_error diverges for Microcanonical, only a positive error
    for the other kinds."""
    d = 10
    for kind in KINDS:
        om, h, p = _start(orc, kind, d)
        e0 = p.scalars()["initial_energy"]
        nxt, st, ee = h.leapfrog(p, 1e-3, 1, energy_baseline=e0 + 50.0, max_energy_error=10.0)  # energy error ~ -50
        assert ee < -40
        assert st == (1 if kind == _abi.NUTS_KINETIC_MICROCANONICAL else 0)
        nxt, st, ee = h.leapfrog(p, 1e-3, 1, energy_baseline=e0 - 50.0, max_energy_error=10.0)  # ~ +50
        assert st == 1


def test_initialize_trajectory_kinds(orc):
    d = 12
    om, h, p = _start(orc, _abi.NUTS_KINETIC_EUCLIDEAN, d)
    v_e = p.vec(p.V)
    om2, h2, p2 = _start(orc, _abi.NUTS_KINETIC_MICROCANONICAL, d)
    v_m = p2.vec(p2.V)
    np.testing.assert_allclose(v_m, v_e / np.linalg.norm(v_e), rtol=1e-14)  # :699-702 same draws, normalised
    s = p2.scalars()
    assert s["kinetic_energy"] == 0.0 and s["initial_energy"] == -(s["logp"] + s["logdet"])  # :720-729
    om3, h3, p3 = _start(orc, _abi.NUTS_KINETIC_EXACT_NORMAL, d)
    np.testing.assert_array_equal(p3.vec(p3.V), v_e)
    assert p3.scalars()["kinetic_energy"] == p.scalars()["kinetic_energy"]


@pytest.mark.parametrize("kind", KINDS)
def test_whole_draws_with_trajectory_kind(orc, kind):
    """Whole NUTS draws of the oracle with NutsSettings::trajectory_kind set (sampler.rs:232, :757), on the target of the reference's
    behavioural check (src/adapt_strategy.rs:367-435: N(30, 1)^10 started at 1.5; not diverging, |x - 30| < 5 at the end)."""
    from nuts_rs_b200 import lib

    d, N = 10, 4
    settings = lib.DiagNutsSettings(num_tune=300, maxdepth=8, trajectory_kind=kind)
    om = orc.Model(_abi.NUTS_LOGP_GAUSS_ISO, d, mu=30.0)
    s = orc.Sampler(om, settings, seed=42, nchains=N)
    st = s.set_position(np.full((N, d), 1.5))
    assert (st == 0).all()
    draws, stats = s.draw(1300)
    assert (np.abs(draws[-1] - 30.0) < 5).all()
    assert not stats["diverging"][-1].any()
    post = draws[300:]
    assert abs(post.mean() - 30.0) < 0.05 and abs(post.std() - 1.0) < 0.08
    if kind == _abi.NUTS_KINETIC_EXACT_NORMAL:
        # the adapted diagonal makes the target a standard normal in z (var of draws = var of gradients): the geodesic flow is exact,
        # so there is no energy error and every proposal is accepted
        assert np.abs(stats["energy_error"][300:]).max() < 1e-9
        assert stats["mean_tree_accept"][300:].min() > 1 - 1e-9
