"""GPU parity, Tier 1 + Tier 2: every batched Math-trait op and fused Hamiltonian op of libnuts_b200.so against the
CPU oracle, through the C ABI.  Elementwise ops must be bit-identical; reductions agree to 1e-12 relative
(order of summation differs: warp butterfly vs 4 SIMD accumulators); leapfrog to 1e-12."""
import numpy as np
import pytest

from helpers import any_f64, assert_approx_eq, rel_err
from nuts_rs_b200 import _abi

pytestmark = pytest.mark.gpu

SIZES = [1, 4, 16, 17, 100, 1000, 4567]  # reference benches/sample.rs:126 sizes plus the C2 dimension


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _math(L, N, d, kind=_abi.NUTS_LOGP_GAUSS_ISO, **kw):
    return L.CudaMath(N, d, kind, **kw)


@pytest.mark.parametrize("d", SIZES)
def test_elementwise_bit_exact(L, orc, d):
    N = 5
    rng = np.random.default_rng(d)
    m = _math(L, N, d, mu=0.0)
    x, y = rng.normal(size=(N, d)), rng.normal(size=(N, d))
    a = rng.normal(size=N)
    px, py, pout = m.from_host(x), m.from_host(y), m.new_array()
    m.axpy_out(px, py, a, pout)
    want = np.stack([orc.axpy_out(x[c], y[c], a[c]) for c in range(N)])
    np.testing.assert_array_equal(pout.box_array(), want)
    m.axpy(px, py, 0.37)
    np.testing.assert_array_equal(py.box_array(), np.stack([orc.axpy(x[c], y[c], 0.37) for c in range(N)]))
    m.array_mult(px, py, pout)
    np.testing.assert_array_equal(pout.box_array(), x * py.box_array())
    m.array_mult_inplace(pout, px)
    np.testing.assert_array_equal(pout.box_array(), x * (x * py.box_array()))
    m.array_recip(px, pout)
    np.testing.assert_array_equal(pout.box_array(), 1.0 / x)
    m.fill_array(pout, 2.5)
    assert (pout.box_array() == 2.5).all()
    m.copy_into(px, pout)
    np.testing.assert_array_equal(pout.box_array(), x)
    # masked axpy leaves inactive chains untouched
    before = py.box_array()
    active = np.array([1, 0, 1, 0, 1], dtype=np.uint8)
    m.axpy(px, py, 2.0, active=active)
    after = py.box_array()
    np.testing.assert_array_equal(after[1], before[1])
    np.testing.assert_array_equal(after[0], np.asarray(orc.axpy(x[0], before[0], 2.0)))
    m.close()


def test_elementwise_any_f64(L, orc):
    """The reference's proptest domain (prop::num::f64::ANY, src/math/util.rs:893-951): 32-ULP with NaN/inf tolerance."""
    N, d = 64, 9
    rng = np.random.default_rng(0)
    m = _math(L, N, d, mu=0.0)
    x, y = any_f64(rng, N * d).reshape(N, d), any_f64(rng, N * d).reshape(N, d)
    a = any_f64(rng, N)
    px, py, pout = m.from_host(x), m.from_host(y), m.new_array()
    m.axpy_out(px, py, a, pout)
    got = pout.box_array()
    for c in range(N):
        want = orc.axpy_out(x[c], y[c], a[c])
        for g, w in zip(got[c], want):
            assert_approx_eq(g, w)
    m.array_mult(px, py, pout)
    got = pout.box_array()
    with np.errstate(all="ignore"):
        want = x * y
    for g, w in zip(got.ravel(), want.ravel()):
        assert_approx_eq(g, w)
    m.close()


@pytest.mark.parametrize("d", SIZES)
def test_reductions(L, orc, d):
    N = 6
    rng = np.random.default_rng(100 + d)
    m = _math(L, N, d, mu=0.0)
    v = [rng.normal(size=(N, d)) for _ in range(5)]
    p = [m.from_host(a) for a in v]
    dot = m.array_vector_dot(p[0], p[1])
    for c in range(N):
        scale = float(np.dot(np.abs(v[0][c]), np.abs(v[1][c])))
        assert abs(dot[c] - orc.vector_dot(v[0][c], v[1][c])) <= 1e-13 * scale + 1e-300
    o1, o2 = m.scalar_prods3(p[0], p[1], p[2], p[3], p[4])
    q1, q2 = m.scalar_prods2(p[0], p[2], p[3], p[4])
    for c in range(N):
        w1, w2 = orc.scalar_prods3(v[0][c], v[1][c], v[2][c], v[3][c], v[4][c])
        s = np.abs(v[0][c] - v[1][c] + v[2][c])
        assert abs(o1[c] - w1) <= 1e-13 * float(np.dot(s, np.abs(v[3][c]))) + 1e-300
        assert abs(o2[c] - w2) <= 1e-13 * float(np.dot(s, np.abs(v[4][c]))) + 1e-300
        w1, w2 = orc.scalar_prods2(v[0][c], v[2][c], v[3][c], v[4][c])
        s = np.abs(v[0][c] + v[2][c])
        assert abs(q1[c] - w1) <= 1e-13 * float(np.dot(s, np.abs(v[3][c]))) + 1e-300
        assert abs(q2[c] - w2) <= 1e-13 * float(np.dot(s, np.abs(v[4][c]))) + 1e-300
    sq = m.sq_norm_sum(p[0], p[1])
    np.testing.assert_allclose(sq, [orc.lib().orc_sq_norm_sum(orc._p(v[0][c]), orc._p(v[1][c]), d) for c in range(N)], rtol=1e-13)
    pos = m.from_host(np.abs(v[0]) + 0.1)
    np.testing.assert_allclose(m.array_sum_ln(pos), np.sum(np.log(np.abs(v[0]) + 0.1), axis=1), rtol=1e-12, atol=1e-12)
    # finite / nonzero predicates
    bad = v[0].copy()
    bad[1, d // 2] = np.inf
    bad[2, 0] = np.nan
    bad[3, d - 1] = 0.0
    pb = m.from_host(bad)
    np.testing.assert_array_equal(m.array_all_finite(pb), [True, False, False, True, True, True])
    np.testing.assert_array_equal(m.array_all_finite_and_nonzero(pb), [True, False, False, False, True, True])
    m.close()


def test_integer_valued_reductions_exact(L):
    """Integer-valued inputs make every summation order exact: catches dropped / double-counted elements for any d."""
    for d in (1, 31, 32, 33, 255, 256, 257, 1000, 4567):
        N = 3
        rng = np.random.default_rng(d)
        x = rng.integers(-50, 50, (N, d)).astype(np.float64)
        y = rng.integers(-50, 50, (N, d)).astype(np.float64)
        m = _math(L, N, d, mu=0.0)
        px, py = m.from_host(x), m.from_host(y)
        np.testing.assert_array_equal(m.array_vector_dot(px, py), np.sum(x * y, axis=1))
        m.close()


def test_variance_and_mass_matrix_updates(L, orc):
    N, d = 4, 37
    rng = np.random.default_rng(5)
    m = _math(L, N, d, mu=0.0)
    mean, var, val = rng.normal(size=(N, d)), rng.random((N, d)), rng.normal(size=(N, d))
    pm, pv, pval = m.from_host(mean), m.from_host(var), m.from_host(val)
    scale = np.array([0.5, 0.25, 0.125, 1.0 / 7])
    m.array_update_variance(pm, pv, pval, scale)
    for c in range(N):
        wm, wv = orc.array_update_variance(mean[c], var[c], val[c], scale[c])
        np.testing.assert_array_equal(pm.box_array()[c], wm)
        np.testing.assert_array_equal(pv.box_array()[c], wv)
    # inv-std updates incl. invalid entries and clamping (reference src/math/cpu_math.rs:633-738)
    dv = rng.random((N, d)) * 10
    gv = rng.random((N, d)) * 10
    dv[0, 0], dv[1, 1], gv[2, 2], dv[3, 3], gv[3, 4] = 0.0, np.inf, 0.0, 1e-300, 1e300
    std0, inv0 = rng.random((N, d)) + 0.5, rng.random((N, d)) + 0.5
    O = orc.lib()
    for mode in ("draw_grad", "draw", "grad"):
        ps, pi = m.from_host(std0), m.from_host(inv0)
        pdv, pgv = m.from_host(dv), m.from_host(gv)
        ws, wi = std0.copy(), inv0.copy()
        for fill in (None, 2.0):
            if mode == "draw_grad":
                m.array_update_var_inv_std_draw_grad(pi, ps, pdv, pgv, fill, (1e-20, 1e20))
                for c in range(N):
                    O.orc_array_update_var_inv_std_draw_grad(orc._p(wi[c]), orc._p(ws[c]), orc._p(dv[c]), orc._p(gv[c]), fill is not None,
                                                             fill or 0.0, 1e-20, 1e20, d)
            elif mode == "draw":
                m.array_update_var_inv_std_draw(pi, ps, pdv, 0.1, fill, (1e-20, 1e20))
                for c in range(N):
                    O.orc_array_update_var_inv_std_draw(orc._p(wi[c]), orc._p(ws[c]), orc._p(dv[c]), 0.1, fill is not None, fill or 0.0,
                                                        1e-20, 1e20, d)
            else:
                m.array_update_var_inv_std_grad(pi, ps, pgv, 1.0, (1e-20, 1e20))
                for c in range(N):
                    O.orc_array_update_var_inv_std_grad(orc._p(wi[c]), orc._p(ws[c]), orc._p(gv[c]), 1.0, 1e-20, 1e20, d)
            np.testing.assert_array_equal(ps.box_array(), ws)
            np.testing.assert_array_equal(pi.box_array(), wi)
    m.close()


def test_array_gaussian_bit_identical(L, orc):
    """The velocity resample is bit-identical on CPU and GPU (deterministic log / sincos)."""
    for d in (1, 2, 7, 10, 100, 1001):
        N = 3
        m = _math(L, N, d, mu=0.0)
        stds = np.random.default_rng(d).random((N, d)) + 0.5
        dest, ps = m.new_array(), m.from_host(stds)
        m.array_gaussian(dest, ps, seed=42, chain_offset=5, counter=17)
        got = dest.box_array()
        for c in range(N):
            normals, newc = orc.fill_normal(42, 5 + c + 1, 17, d)
            assert newc == 17 + (d + 1) // 2
            np.testing.assert_array_equal(got[c], stds[c] * normals)
        m.close()


MODELS = [
    ("iso", dict(kind=_abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)),
    ("diag", dict(kind=_abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma="logspace")),
    ("rank1", dict(kind=_abi.NUTS_LOGP_GAUSS_RANK1, mu=0.0, rank1_scale=0.5)),
    ("funnel", dict(kind=_abi.NUTS_LOGP_FUNNEL, funnel_scale=3.0)),
]


def _model_kwargs(spec, d):
    kw = dict(spec)
    if isinstance(kw.get("sigma"), str):
        kw["sigma"] = np.exp(np.linspace(-1, 1, d))
    return kw


@pytest.mark.parametrize("name,spec", MODELS)
@pytest.mark.parametrize("d", [2, 10, 100, 1000, 3001])
def test_logp_array(L, orc, name, spec, d):
    N = 4
    kw = _model_kwargs(spec, d)
    kind = kw.pop("kind")
    m = L.CudaMath(N, d, kind, **kw)
    om = orc.Model(kind, d, **kw)
    x = np.random.default_rng(d).normal(size=(N, d))
    px, pg = m.from_host(x), m.new_array()
    logp, status = m.logp_array(px, pg)
    g = pg.box_array()
    for c in range(N):
        wl, wg = om.logp(x[c])
        assert abs(logp[c] - wl) <= 1e-12 * max(1.0, abs(wl))
        assert rel_err(g[c], wg) < 1e-12
    assert (status == 0).all()
    m.close()


@pytest.mark.parametrize("name,spec", MODELS)
@pytest.mark.parametrize("d", [3, 10, 100, 1000])
def test_leapfrog_parity(L, orc, name, spec, d):
    """Hamiltonian::leapfrog (reference src/dynamics/transformed_hamiltonian.rs:524-615) from identical inputs:
    init_state, initialize_trajectory (same seed => bit-identical velocity), 5 steps forward and 5 backward."""
    N = 3
    kw = _model_kwargs(spec, d)
    kind = kw.pop("kind")
    # the funnel gradient of x0 cancels O(d) terms (-(d-1)/2 + exp(-v)*S/2): summation-order noise is amplified, so it gets
    # the north-star trajectory tolerance (1e-9); the Gaussian targets stay at 1e-11
    tol = 1e-9 if name == "funnel" else 1e-11
    rng = np.random.default_rng(7 * d)
    m = L.CudaMath(N, d, kind, **kw)
    stds, mean = np.exp(0.3 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d))
    m.set_transform(stds, mean)
    x0 = rng.normal(size=(N, d))
    p, status = m.init_state(x0)
    assert (status == 0).all()
    m.initialize_trajectory(p, True, seed=9, chain_offset=0, counter=3)
    eps = 0.05 + 0.02 * rng.random(N)
    om = orc.Model(kind, d, **kw)
    for c in range(N):
        h = orc.Hamiltonian(om)
        h.set_transform(stds[c], mean[c])
        op, ost = h.init_state(x0[c])
        assert ost == 0
        h.initialize_trajectory(op, True, 9, c + 1, 3)
        np.testing.assert_array_equal(p.vec(p.V)[c], op.vec(op.V))  # velocity bit-identical
        np.testing.assert_array_equal(p.vec(p.Z)[c], op.vec(op.Z))
        sc, osc = p.scalars(), op.scalars()
        assert abs(sc["initial_energy"][c] - osc["initial_energy"]) <= 1e-12 * max(1.0, abs(osc["initial_energy"]))
        for direction in (1, -1):
            cur, ocur = p, op
            for step in range(5):
                nxt, st, ee = m.leapfrog(cur, eps, direction=direction)
                onxt, ost2, oee = h.leapfrog(ocur, eps[c], direction)
                assert st[c] == ost2
                for which in range(5):
                    assert rel_err(nxt.vec(which)[c], onxt.vec(which)) < tol, (which, step)
                s1, s2 = nxt.scalars(), onxt.scalars()
                assert s1["index_in_trajectory"][c] == s2["index_in_trajectory"] == direction * (step + 1)
                for key in ("logp", "kinetic_energy", "logdet", "initial_energy"):
                    assert abs(s1[key][c] - s2[key]) <= tol * max(1.0, abs(s2[key])), key
                assert abs(ee[c] - oee) <= 1e-9 * max(1.0, abs(oee), abs(s2["initial_energy"]))
                # is_turning agrees for (start, current)
                assert m.is_turning(p, nxt)[c] == h.is_turning(op, onxt)
                cur, ocur = nxt, onxt
    m.close()


@pytest.mark.parametrize("kind", [_abi.NUTS_LOGP_GAUSS_ISO, _abi.NUTS_LOGP_GAUSS_DIAG])
@pytest.mark.parametrize("d", [1, 3, 255, 511, 512, 513, 1000, 1024, 1025, 1537, 2000, 4567])
def test_leapfrog_tma_matches_register_path(L, kind, d, monkeypatch):
    """k_leapfrog_tma (input rows staged through shared memory by cp.async.bulk + mbarrier, the default for the elementwise targets)
    against k_leapfrog's register path (NUTS_B200_PLANE_TMA=0): every plane and scalar bit-identical, over chunk boundaries (512
    elements per stage), stage refills (d > 1024), odd tails, per-chain steps, both directions and a mask."""
    N = 6
    rng = np.random.default_rng(3 * d + kind)
    kw = dict(mu=0.5)
    if kind == _abi.NUTS_LOGP_GAUSS_DIAG:
        kw["sigma"] = np.exp(np.linspace(-1, 1, d))
    m = L.CudaMath(N, d, kind, **kw)
    m.set_transform(np.exp(0.3 * rng.normal(size=(N, d))), 0.1 * rng.normal(size=(N, d)))
    p, status = m.init_state(rng.normal(size=(N, d)))
    assert (status == 0).all()
    m.initialize_trajectory(p, True, seed=11, chain_offset=0, counter=0)
    eps = 0.05 + 0.02 * rng.random(N)
    direction = np.array([1, -1, 1, -1, -1, 1], dtype=np.int8)
    active = np.array([1, 1, 0, 1, 1, 1], dtype=np.uint8)
    results = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("NUTS_B200_PLANE_TMA", flag)
        cur, outs = p, []
        for step in range(3):
            nxt = m.new_point()
            nxt.set_vec(nxt.Z, np.full((N, d), -7.0))
            nxt, st, ee = m.leapfrog(cur, eps, direction=direction, active=active if step == 1 else None, out=nxt)
            outs.append(([nxt.vec(w) for w in range(5)], nxt.scalars(), st.copy(), ee.copy()))
            if step == 1:  # the masked chain keeps whatever the output point held
                assert (nxt.vec(nxt.Z)[2] == -7.0).all()
                nxt.set_vec(nxt.Z, np.where(active[:, None] != 0, nxt.vec(nxt.Z), cur.vec(cur.Z)))
            cur = nxt
        results[flag] = outs
    for a, b in zip(results["1"], results["0"]):
        for va, vb in zip(a[0], b[0]):
            np.testing.assert_array_equal(va, vb)
        for key in a[1]:
            np.testing.assert_array_equal(np.asarray(a[1][key])[active != 0], np.asarray(b[1][key])[active != 0], err_msg=key)
        np.testing.assert_array_equal(a[2], b[2])
        np.testing.assert_array_equal(a[3][active != 0], b[3][active != 0])
    m.close()


def test_last_kernel_ms(L):
    """nuts_ctx_last_kernel_ms: CUDA-event time of the kernel of the last nuts_leapfrog (bench.py's plane microbench reads it)."""
    m = L.CudaMath(64, 1000, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    with pytest.raises(L.NutsError):
        m.last_kernel_ms()
    m.set_transform(np.ones((64, 1000)), np.zeros((64, 1000)))
    p, _ = m.init_state(np.ones((64, 1000)))
    m.initialize_trajectory(p, True, 1, 0, 0)
    m.leapfrog(p, 0.1)
    ms = m.last_kernel_ms()
    assert 0.0 < ms < 50.0
    m.close()


def test_leapfrog_divergence_and_mask(L, orc):
    N, d = 4, 10
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    m.set_transform(np.ones((N, d)), np.zeros((N, d)))
    p, _ = m.init_state(np.full((N, d), 2.0))
    m.initialize_trajectory(p, True, 1, 0, 0)
    before = None
    # a huge step diverges (energy error > max_energy_error); masked chains are not touched
    out = m.new_point()
    out.set_vec(out.Z, np.full((N, d), -7.0))
    _, status, ee = m.leapfrog(p, 50.0, max_energy_error=1000.0, active=np.array([1, 1, 0, 1], dtype=np.uint8), out=out)
    assert status[0] == 1 and status[1] == 1 and status[3] == 1 and status[2] == 0
    assert (out.vec(out.Z)[2] == -7.0).all()
    assert (ee[[0, 1, 3]] > 1000).all()
    m.close()


def test_init_state_rejects_bad_points(L):
    """reference transformed_hamiltonian.rs:310-324,654-657: zero whitened gradient / non-finite position => BadInitGrad."""
    N, d = 3, 5
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=0.0)
    m.set_transform(np.ones((N, d)), np.zeros((N, d)))
    x = np.ones((N, d))
    x[1, 2] = 0.0  # gradient exactly zero at the mode coordinate
    x[2, 0] = np.inf
    _, status = m.init_state(x)
    np.testing.assert_array_equal(status, [0, 3, 3])
    m.close()


def test_branch_free_division_and_sqrt_are_ieee_exact(L):
    """device_common.cuh div_fast / sqrt_fast (the fast paths of the compiler's own sequences, used by the per-draw mass-matrix
    update so that the elements of a thread overlap): wherever their range test holds the result is bit-identical to `/` and
    sqrt(), i.e. to the reference's f64 arithmetic (src/math/cpu_math.rs:671-708); where it fails the engine falls back."""
    import ctypes as C

    lib = L.load()
    rng = np.random.default_rng(7)
    n = 1 << 20
    mags = 10.0 ** rng.uniform(-300, 300, size=n)
    a = np.concatenate([rng.normal(size=n) * mags, 10.0 ** rng.uniform(-25, 25, size=n), any_f64(rng, n // 4)])
    b = np.concatenate([rng.normal(size=n) * 10.0 ** rng.uniform(-300, 300, size=n), 10.0 ** rng.uniform(-25, 25, size=n), any_f64(rng, n // 4)])
    a[:n] = np.abs(a[:n])  # sqrt operands
    m = a.size
    out = [np.empty(m) for _ in range(4)]
    okd, oks = np.empty(m, dtype=np.uint8), np.empty(m, dtype=np.uint8)
    dp = _abi.c_double_p
    lib.nuts_debug_fast_math.argtypes = [dp] * 6 + [_abi.c_u8_p, _abi.c_u8_p, C.c_uint64]
    rc = lib.nuts_debug_fast_math(a.ctypes.data_as(dp), b.ctypes.data_as(dp), *[o.ctypes.data_as(dp) for o in out],
                                  okd.ctypes.data_as(_abi.c_u8_p), oks.ctypes.data_as(_abi.c_u8_p), m)
    assert rc == 0, lib.nuts_last_error()
    qf, qr, rf, rr = out
    d, s = okd.astype(bool), oks.astype(bool)
    assert d[n:2 * n].all() and s[n:2 * n].all()  # the whole range the mass-matrix update works in takes the fast path
    assert d.mean() > 0.5 and s.mean() > 0.5
    assert np.array_equal(qf[d].view(np.uint64), qr[d].view(np.uint64))
    assert np.array_equal(rf[s].view(np.uint64), rr[s].view(np.uint64))
    with np.errstate(all="ignore"):
        assert np.array_equal(qr[d].view(np.uint64), (a[d] / b[d]).view(np.uint64))  # and the CPU's IEEE division
        assert np.array_equal(rr[s].view(np.uint64), np.sqrt(a[s]).view(np.uint64))


@pytest.mark.parametrize("d", [10, 16, 1000, 1001])
def test_plane_device_pointer_reports_the_padded_row_stride(L, d):
    """nuts_plane_device_ptr: rows start on 128-byte lines (stride = dim rounded up to 16 doubles); a torch view built from the
    pointer and the stride sees exactly what read_from_slice wrote, for dim % 16 != 0 too."""
    import torch

    N = 5
    m = _math(L, N, d)
    x = np.random.default_rng(d).normal(size=(N, d))
    plane = m.new_array().read_from_slice(x)
    ptr, stride = plane.device_ptr()
    assert ptr != 0 and stride == (d + 15) // 16 * 16
    # copy the padded rows out with torch (device-to-device memcpy from the raw address) and compare
    import ctypes as C

    torch.cuda.synchronize()
    buf = torch.empty(N * stride, dtype=torch.float64, device="cuda")
    cudart = C.CDLL("libcudart.so.12")  # (already loaded by torch / libnuts_b200.so)
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rc = cudart.cudaMemcpy(buf.data_ptr(), ptr, N * stride * 8, 3)  # cudaMemcpyDeviceToDevice
    assert rc == 0
    got = buf.cpu().numpy().reshape(N, stride)
    np.testing.assert_array_equal(got[:, :d], x)
    np.testing.assert_array_equal(got[:, d:], 0.0)  # the padding is zero
    m.close()
