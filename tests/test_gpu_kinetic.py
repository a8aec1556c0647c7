"""GPU parity for KineticEnergyKind::ExactNormal / Microcanonical at Tier 1 (Math::std_norm_flow, std_norm_grad_flow(_inplace),
array_normalize, esh_momentum_update: reference src/math/math.rs:155-210) and Tier 2 (Hamiltonian::leapfrog / initialize_trajectory
with a kinetic energy kind: src/dynamics/transformed_hamiltonian.rs:160-258, 524-615, 687-736), through the C ABI against the CPU
oracle.  Elementwise flows are bit-identical (sin / cos of the step are evaluated on the host like the reference does); everything
behind a reduction agrees to 1e-12 (ESH) / the leapfrog tolerance of test_gpu_primitives (1e-11, funnel 1e-9)."""
import math

import numpy as np
import pytest

from helpers import any_f64, assert_approx_eq, rel_err
from nuts_rs_b200 import _abi

pytestmark = pytest.mark.gpu

SIZES = [2, 3, 4, 17, 100, 1000, 4567]
EXACT, MICRO = _abi.NUTS_KINETIC_EXACT_NORMAL, _abi.NUTS_KINETIC_MICROCANONICAL


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


@pytest.mark.parametrize("d", SIZES)
def test_std_norm_flows_bit_exact(L, orc, d):
    N = 5
    rng = np.random.default_rng(d)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    pos, grad, vel = rng.normal(size=(N, d)), rng.normal(size=(N, d)), rng.normal(size=(N, d))
    eps = rng.uniform(-10, 10, size=N)
    pp, pg, pv, pout = m.from_host(pos), m.from_host(grad), m.from_host(vel), m.new_array()
    # per-chain epsilon
    m.std_norm_flow(pp, pout, pv, eps)
    want = [orc.std_norm_flow(pos[c], vel[c], eps[c]) for c in range(N)]
    np.testing.assert_array_equal(pout.box_array(), np.stack([w[0] for w in want]))
    np.testing.assert_array_equal(pv.box_array(), np.stack([w[1] for w in want]))
    # broadcast epsilon + mask
    pv = m.from_host(vel)
    active = np.array([1, 0, 1, 1, 0], dtype=np.uint8)
    m.fill_array(pout, -7.0)
    m.std_norm_flow(pp, pout, pv, 0.3, active=active)
    got_p, got_v = pout.box_array(), pv.box_array()
    for c in range(N):
        if active[c]:
            wp, wv = orc.std_norm_flow(pos[c], vel[c], 0.3)
            np.testing.assert_array_equal(got_p[c], wp)
            np.testing.assert_array_equal(got_v[c], wv)
        else:
            assert (got_p[c] == -7.0).all()
            np.testing.assert_array_equal(got_v[c], vel[c])
    # gradient flow, out of place and in place
    pv = m.from_host(vel)
    m.std_norm_grad_flow(pp, pg, pv, pout, eps)
    np.testing.assert_array_equal(pout.box_array(), np.stack([orc.std_norm_grad_flow(pos[c], grad[c], vel[c], eps[c]) for c in range(N)]))
    m.std_norm_grad_flow_inplace(pp, pg, pv, -0.4)
    np.testing.assert_array_equal(pv.box_array(), np.stack([orc.std_norm_grad_flow(pos[c], grad[c], vel[c], -0.4, inplace=True) for c in range(N)]))
    m.close()


def test_std_norm_flows_any_f64(L, orc):
    """The reference's proptest domain (util.rs:808-877): arbitrary bit patterns, epsilon in -10..10, 32-ULP / NaN-inf equivalence."""
    N, d = 6, 37
    rng = np.random.default_rng(99)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    pos, grad, vel = (any_f64(rng, N * d).reshape(N, d) for _ in range(3))
    eps = rng.uniform(-10, 10, size=N)
    pp, pg, pv, pout = m.from_host(pos), m.from_host(grad), m.from_host(vel), m.new_array()
    m.std_norm_flow(pp, pout, pv, eps)
    got_p, got_v = pout.box_array(), pv.box_array()
    for c in range(N):
        wp, wv = orc.std_norm_flow(pos[c], vel[c], eps[c])
        for a, b in zip(got_p[c], wp):
            assert_approx_eq(a, b)
        for a, b in zip(got_v[c], wv):
            assert_approx_eq(a, b)
    pv = m.from_host(vel)
    m.std_norm_grad_flow(pp, pg, pv, pout, eps)
    got = pout.box_array()
    for c in range(N):
        for a, b in zip(got[c], orc.std_norm_grad_flow(pos[c], grad[c], vel[c], eps[c])):
            assert_approx_eq(a, b)
    m.close()


@pytest.mark.parametrize("d", SIZES)
def test_normalize_and_esh_momentum_update(L, orc, d):
    N = 4
    rng = np.random.default_rng(5 * d)
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    g = rng.normal(size=(N, d)) * 10 ** rng.uniform(-1, 1, size=(N, 1))
    p0 = rng.normal(size=(N, d))
    pg, pm = m.from_host(g), m.from_host(p0)
    m.array_normalize(pm)
    unit = pm.box_array()
    for c in range(N):
        assert rel_err(unit[c], orc.array_normalize(p0[c])) < 1e-13
    step = rng.uniform(-0.5, 0.5, size=N)
    dke = m.esh_momentum_update(pg, pm, step)
    got = pm.box_array()
    for c in range(N):
        want, want_dke = orc.esh_momentum_update(g[c], unit[c], step[c])
        assert rel_err(got[c], want) < 1e-12
        assert abs(dke[c] - want_dke) <= 1e-12 * max(1.0, abs(want_dke)) * d
        assert abs(np.sum(got[c] ** 2) - 1) < 1e-13
    # broadcast step, masked chains untouched
    before = pm.box_array()
    dke = m.esh_momentum_update(pg, pm, 0.05, active=np.array([1, 0, 1, 0], dtype=np.uint8))
    after = pm.box_array()
    np.testing.assert_array_equal(after[1], before[1])
    np.testing.assert_array_equal(after[3], before[3])
    assert dke[1] == 0.0 and dke[3] == 0.0
    want, want_dke = orc.esh_momentum_update(g[0], before[0], 0.05)
    assert rel_err(after[0], want) < 1e-12 and abs(dke[0] - want_dke) <= 1e-12 * max(1.0, abs(want_dke)) * d
    m.close()


def test_esh_needs_two_dimensions(L):
    m = L.CudaMath(2, 1, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    a, b = m.new_array(), m.new_array()
    with pytest.raises(L.NutsError):  # cpu_math.rs:514 assert!(n >= 2)
        m.esh_momentum_update(a, b, 0.1)
    p, _ = m.init_state(np.ones((2, 1)))
    with pytest.raises(L.NutsError):
        m.leapfrog(p, 0.1, kind=MICRO)
    with pytest.raises(L.NutsError):
        m.leapfrog(p, 0.1, kind=7)
    m.close()


MODELS = [
    ("iso", dict(kind=_abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)),
    ("diag", dict(kind=_abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma="logspace")),
    ("rank1", dict(kind=_abi.NUTS_LOGP_GAUSS_RANK1, mu=0.0, rank1_scale=0.5)),
    ("funnel", dict(kind=_abi.NUTS_LOGP_FUNNEL, funnel_scale=3.0)),
]


@pytest.mark.parametrize("kinetic", [EXACT, MICRO])
@pytest.mark.parametrize("name,spec", MODELS)
@pytest.mark.parametrize("d", [3, 10, 100, 1000])
def test_leapfrog_kinetic_parity(L, orc, kinetic, name, spec, d):
    """Same shape as test_gpu_primitives.test_leapfrog_parity: identical inputs, 5 steps forward and 5 backward, every plane and
    scalar of the point compared after every step."""
    N = 3
    kw = dict(spec)
    if isinstance(kw.get("sigma"), str):
        kw["sigma"] = np.exp(np.linspace(-1, 1, d))
    kind = kw.pop("kind")
    tol = 1e-9 if name == "funnel" else 1e-11
    rng = np.random.default_rng(7 * d + kinetic)
    m = L.CudaMath(N, d, kind, **kw)
    stds, mean = np.exp(0.3 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d))
    m.set_transform(stds, mean)
    x0 = rng.normal(size=(N, d))
    p, status = m.init_state(x0)
    assert (status == 0).all()
    m.initialize_trajectory(p, True, seed=9, chain_offset=0, counter=3, kind=kinetic)
    eps = 0.05 + 0.02 * rng.random(N)
    if kinetic == MICRO:
        eps = eps / math.sqrt(d)  # the Microcanonical step is scaled by sqrt(dim) (:214-219)
    om = orc.Model(kind, d, **kw)
    for c in range(N):
        h = orc.Hamiltonian(om)
        h.set_kinetic_energy_kind(kinetic)
        h.set_transform(stds[c], mean[c])
        op, ost = h.init_state(x0[c])
        assert ost == 0
        h.initialize_trajectory(op, True, 9, c + 1, 3)
        if kinetic == MICRO:
            assert rel_err(p.vec(p.V)[c], op.vec(op.V)) < 1e-13  # normalised: |v|^2 is summed in a different order
            assert abs(np.sum(p.vec(p.V)[c] ** 2) - 1) < 1e-13
            assert p.scalars()["kinetic_energy"][c] == 0.0
        else:
            np.testing.assert_array_equal(p.vec(p.V)[c], op.vec(op.V))
        sc, osc = p.scalars(), op.scalars()
        assert abs(sc["initial_energy"][c] - osc["initial_energy"]) <= 1e-12 * max(1.0, abs(osc["initial_energy"]))
        for direction in (1, -1):
            cur, ocur = p, op
            for step in range(5):
                nxt, st, ee = m.leapfrog(cur, eps, direction=direction, kind=kinetic)
                onxt, ost2, oee = h.leapfrog(ocur, eps[c], direction)
                assert st[c] == ost2  # the d = 1000 funnel diverges under the geodesic integrator: on both sides, at the same step
                for which in range(5):
                    assert rel_err(nxt.vec(which)[c], onxt.vec(which)) < tol, (which, step)
                s1, s2 = nxt.scalars(), onxt.scalars()
                assert s1["index_in_trajectory"][c] == s2["index_in_trajectory"] == direction * (step + 1)
                for key in ("logp", "kinetic_energy", "logdet", "initial_energy"):
                    assert abs(s1[key][c] - s2[key]) <= tol * max(1.0, abs(s2[key])) * (d if key == "kinetic_energy" and kinetic == MICRO else 1), key
                assert abs(ee[c] - oee) <= 1e-9 * max(1.0, abs(oee), abs(s2["initial_energy"]))
                assert m.is_turning(p, nxt)[c] == h.is_turning(op, onxt)
                cur, ocur = nxt, onxt
    m.close()


def test_exact_normal_is_exact_on_the_standard_normal(L):
    """transformed_hamiltonian.rs:28-36: with a standard-normal transformed posterior the geodesic leapfrog has no energy error at any
    step size; the Euclidean one at the same step does."""
    N, d = 4, 1000
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    m.set_transform(np.ones((N, d)), np.zeros((N, d)))
    x0 = np.random.default_rng(1).normal(size=(N, d))
    p, _ = m.init_state(x0)
    m.initialize_trajectory(p, True, 5, 0, 0, kind=EXACT)
    v0 = p.vec(p.V)
    eps = np.array([0.1, 1.0, 2.5, 7.0])
    cur = p
    for k in range(1, 5):
        cur, st, ee = m.leapfrog(cur, eps, kind=EXACT)
        assert (st == 0).all() and np.abs(ee).max() < 1e-9
        np.testing.assert_allclose(cur.vec(cur.Z), x0 * np.cos(k * eps)[:, None] + v0 * np.sin(k * eps)[:, None], atol=1e-12)
    _, _, ee = m.leapfrog(p, 1.0)
    assert np.abs(ee).min() > 1.0
    m.close()


def test_kinetic_divergence_rules_and_mask(L):
    """:591-596: Microcanonical diverges on |energy error| >= max, the others on energy error > max; masked chains untouched."""
    N, d = 4, 10
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    m.set_transform(np.ones((N, d)), np.zeros((N, d)))
    x0 = np.random.default_rng(2).normal(size=(N, d))
    for kinetic in (EXACT, MICRO):
        p, _ = m.init_state(x0)
        m.initialize_trajectory(p, True, 1, 0, 0, kind=kinetic)
        e0 = p.scalars()["initial_energy"]
        out = m.new_point()
        out.set_vec(out.Z, np.full((N, d), -7.0))
        _, st, ee = m.leapfrog(p, 1e-3, energy_baseline=e0 + 50.0, max_energy_error=10.0, kind=kinetic,
                               active=np.array([1, 1, 0, 1], dtype=np.uint8), out=out)
        assert (ee[[0, 1, 3]] < -40).all()
        assert (st[[0, 1, 3]] == (1 if kinetic == MICRO else 0)).all() and st[2] == 0
        assert (out.vec(out.Z)[2] == -7.0).all()
        _, st, ee = m.leapfrog(p, 1e-3, energy_baseline=e0 - 50.0, max_energy_error=10.0, kind=kinetic)
        assert (st == 1).all()
    m.close()


def test_whole_draw_sampler_rejects_other_kinds(L):
    """The Tier-3 engines are Euclidean: asking for another trajectory_kind fails loudly instead of sampling wrongly."""
    m = L.CudaMath(4, 10, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    with pytest.raises(L.NutsError):
        L.Sampler(m, L.DiagNutsSettings(trajectory_kind=EXACT), seed=1)
    m.close()
