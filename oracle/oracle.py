"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/libnuts_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
The product (nuts_rs_b200) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from nuts_rs_b200 import _abi  # noqa: E402  (POD struct definitions of include/nuts_b200.h only)

_LIB = None
dp = _abi.c_double_p


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libnuts_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_logaddexp.restype = C.c_double
        L.orc_logaddexp.argtypes = [C.c_double, C.c_double]
        L.orc_vector_dot.restype = C.c_double
        L.orc_vector_dot.argtypes = [dp, dp, C.c_size_t]
        L.orc_axpy.argtypes = [dp, dp, C.c_double, C.c_size_t]
        L.orc_std_norm_flow.argtypes = [dp, dp, dp, C.c_double, C.c_size_t]
        L.orc_std_norm_grad_flow.argtypes = [dp, dp, dp, dp, C.c_double, C.c_size_t]
        L.orc_std_norm_grad_flow_inplace.argtypes = [dp, dp, dp, C.c_double, C.c_size_t]
        L.orc_array_normalize.argtypes = [dp, C.c_size_t]
        L.orc_esh_momentum_update.restype = C.c_double
        L.orc_esh_momentum_update.argtypes = [dp, dp, C.c_double, C.c_size_t]
        L.orc_ham_set_kinetic_energy_kind.argtypes = [C.c_void_p, C.c_int]
        L.orc_axpy_out.argtypes = [dp, dp, C.c_double, dp, C.c_size_t]
        L.orc_multiply.argtypes = [dp, dp, dp, C.c_size_t]
        L.orc_multiply_inplace.argtypes = [dp, dp, C.c_size_t]
        L.orc_scalar_prods3.argtypes = [dp, dp, dp, dp, dp, C.c_size_t, dp, dp]
        L.orc_scalar_prods2.argtypes = [dp, dp, dp, dp, C.c_size_t, dp, dp]
        L.orc_sq_norm_sum.restype = C.c_double
        L.orc_sq_norm_sum.argtypes = [dp, dp, C.c_size_t]
        L.orc_array_all_finite.argtypes = [dp, C.c_size_t]
        L.orc_array_all_finite_and_nonzero.argtypes = [dp, C.c_size_t]
        L.orc_array_sum_ln.restype = C.c_double
        L.orc_array_sum_ln.argtypes = [dp, C.c_size_t]
        L.orc_array_update_variance.argtypes = [dp, dp, dp, C.c_double, C.c_size_t]
        L.orc_array_update_var_inv_std_draw.argtypes = [dp, dp, dp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_size_t]
        L.orc_array_update_var_inv_std_draw_grad.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_size_t]
        L.orc_array_update_var_inv_std_grad.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_size_t]
        L.orc_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        L.orc_det_log.restype = C.c_double
        L.orc_det_log.argtypes = [C.c_double]
        L.orc_det_sincos2pi.argtypes = [C.c_double, dp, dp]
        L.orc_fill_normal.restype = C.c_uint64
        L.orc_fill_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, dp, C.c_size_t]
        L.orc_next_f64.restype = C.c_double
        L.orc_next_f64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_next_bool.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_uniform.restype = C.c_double
        L.orc_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double]
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.POINTER(_abi.LogpDesc), C.c_uint64]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_model_logp.restype = C.c_double
        L.orc_model_logp.argtypes = [C.c_void_p, dp, dp]
        L.orc_ham_create.restype = C.c_void_p
        L.orc_ham_create.argtypes = [C.c_void_p]
        L.orc_ham_destroy.argtypes = [C.c_void_p]
        L.orc_ham_set_step_size.argtypes = [C.c_void_p, C.c_double]
        L.orc_ham_set_transform.argtypes = [C.c_void_p, dp, dp]
        L.orc_ham_set_lowrank_transform.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, C.c_uint64]
        L.orc_ham_set_lowrank_transform.restype = C.c_int
        L.orc_apply_lowrank_transform.argtypes = [dp, dp, dp, dp, C.c_uint64, C.c_uint64]
        L.orc_sampler_set_lowrank_transform.argtypes = [C.c_void_p, dp, dp, C.c_uint64, dp, dp, C.POINTER(C.c_int32), dp, C.POINTER(C.c_uint8)]
        L.orc_sampler_set_lowrank_transform.restype = None
        L.orc_lowrank_spd_mean.argtypes = [dp, dp, C.c_uint64, dp]
        L.orc_lowrank_spd_mean.restype = C.c_int
        L.orc_lowrank_estimate_mass_matrix.argtypes = [dp, dp, C.c_uint64, C.c_uint64, C.c_double, dp, dp]
        L.orc_lowrank_estimate_mass_matrix.restype = C.c_int
        L.orc_lowrank_compute_update.argtypes = [dp, dp, C.c_uint64, C.c_uint64, C.c_double, C.c_double, dp, dp, dp, dp, dp]
        L.orc_lowrank_compute_update.restype = C.c_int64
        L.orc_ham_update_diag_draw_grad.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_ham_update_diag_grad.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_double]
        L.orc_ham_update_diag_draw.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_ham_get_transform.argtypes = [C.c_void_p, dp, dp, dp, dp, _abi.c_i64_p]
        L.orc_point_create.restype = C.c_void_p
        L.orc_point_create.argtypes = [C.c_void_p]
        L.orc_point_destroy.argtypes = [C.c_void_p]
        L.orc_point_get_vec.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_point_set_vec.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_point_get_scalars.argtypes = [C.c_void_p, _abi.c_i64_p, dp, dp, dp, dp, _abi.c_i64_p]
        L.orc_point_set_scalars.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64]
        L.orc_ham_init_state.argtypes = [C.c_void_p, C.c_void_p, dp]
        L.orc_ham_init_from_untransformed.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_ham_init_from_transformed.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_ham_initialize_trajectory.restype = C.c_uint64
        L.orc_ham_initialize_trajectory.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_ham_leapfrog.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, dp]
        L.orc_ham_is_turning.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sampler_create.restype = C.c_void_p
        L.orc_sampler_create.argtypes = [C.c_void_p, C.POINTER(_abi.NutsSettings), C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_sampler_destroy.argtypes = [C.c_void_p]
        L.orc_sampler_set_position.argtypes = [C.c_void_p, dp, _abi.c_i32_p, C.c_int]
        L.orc_sampler_draw.restype = C.c_uint64
        L.orc_sampler_draw.argtypes = [C.c_void_p, C.c_uint64, dp, C.POINTER(_abi.Stats), C.c_int]
        L.orc_sampler_get_state.argtypes = [C.c_void_p, dp, dp, dp, dp, _abi.c_u64_p]
        L.orc_sampler_set_step_size.argtypes = [C.c_void_p, dp]
        L.orc_sampler_get_chain_state.argtypes = [C.c_void_p, C.POINTER(_abi.ChainState)]
        L.orc_sampler_set_chain_state.argtypes = [C.c_void_p, C.POINTER(_abi.ChainState)]
        L.orc_sampler_counters.argtypes = [C.c_void_p, _abi.c_u64_p, _abi.c_u64_p]
        L.orc_settings_default.argtypes = [C.POINTER(_abi.NutsSettings)]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- primitives ----
def logaddexp(a, b):
    return lib().orc_logaddexp(a, b)


def axpy(x, y, a):
    x = _f64(x)
    y = _f64(y).copy()
    lib().orc_axpy(_p(x), _p(y), a, x.size)
    return y


def axpy_out(x, y, a):
    x = _f64(x)
    y = _f64(y)
    out = np.empty_like(x)
    lib().orc_axpy_out(_p(x), _p(y), a, _p(out), x.size)
    return out


def multiply(x, y):
    x = _f64(x)
    y = _f64(y)
    out = np.empty_like(x)
    lib().orc_multiply(_p(x), _p(y), _p(out), x.size)
    return out


def apply_lowrank_transform(vecs, vals, rhs):
    """(I + U (diag(vals) - I) U^T) rhs  (src/math/cpu_math.rs:332-377); vecs: [r, d]."""
    vals, rhs = _f64(vals), _f64(rhs)
    vecs = _f64(np.asarray(vecs, dtype=np.float64).reshape(len(vals), len(rhs)))
    out = np.empty_like(rhs)
    lib().orc_apply_lowrank_transform(_p(vecs), _p(vals), _p(rhs), _p(out), len(rhs), len(vals))
    return out


def lowrank_spd_mean(cov_draws, cov_grads):
    """spd_mean (src/transform/adapt/low_rank.rs:241-268)."""
    a, b = _f64(cov_draws), _f64(cov_grads)
    out = np.empty_like(a)
    ok = lib().orc_lowrank_spd_mean(_p(a), _p(b), a.shape[0], _p(out))
    return out if ok else None


def lowrank_estimate_mass_matrix(draws, grads, gamma):
    """estimate_mass_matrix (adapt/low_rank.rs:210-239): draws, grads [k, n]; returns (vals ascending, vecs [k, k] with eigenvector j in
    column j) or None."""
    d, g = _f64(draws), _f64(grads)
    k, n = d.shape
    vals, vecs = np.empty(k), np.empty((k, k))
    ok = lib().orc_lowrank_estimate_mass_matrix(_p(d), _p(g), k, n, gamma, _p(vals), _p(vecs))
    return (vals, vecs) if ok else None


def lowrank_compute_update(draws, grads, gamma=1e-5, eigval_cutoff=2.0):
    """LowRankMassMatrixStrategy::compute_update (adapt/low_rank.rs:73-131): draws, grads [n, dim] (oldest first).  Returns
    (stds, mean, vals [r], vecs [r, dim], mean_low_rank) or None."""
    d_, g_ = _f64(draws), _f64(grads)
    n, d = d_.shape
    cap = 2 * min(n, d)
    stds, mean, mu = np.empty(d), np.empty(d), np.empty(d)
    vals, vecs = np.empty(cap), np.empty((cap, d))
    r = lib().orc_lowrank_compute_update(_p(d_), _p(g_), n, d, gamma, eigval_cutoff, _p(stds), _p(mean), _p(vals), _p(vecs), _p(mu))
    if r < 0:
        return None
    return stds, mean, vals[:r].copy(), vecs[:r].copy(), mu


def std_norm_flow(pos, vel, epsilon):
    """util.rs:575 std_norm_flow: returns (pos_out, vel')."""
    pos, vel = _f64(pos), _f64(vel).copy()
    out = np.empty_like(pos)
    lib().orc_std_norm_flow(_p(pos), _p(out), _p(vel), epsilon, pos.size)
    return out, vel


def std_norm_grad_flow(pos, grad, vel, epsilon, inplace=False):
    """util.rs:650 / :726 std_norm_grad_flow(_inplace): returns vel + eps * (pos + grad)."""
    pos, grad, vel = _f64(pos), _f64(grad), _f64(vel)
    if inplace:
        v = vel.copy()
        lib().orc_std_norm_grad_flow_inplace(_p(pos), _p(grad), _p(v), epsilon, pos.size)
        return v
    out = np.empty_like(pos)
    lib().orc_std_norm_grad_flow(_p(pos), _p(grad), _p(vel), _p(out), epsilon, pos.size)
    return out


def array_normalize(v):
    v = _f64(v).copy()
    lib().orc_array_normalize(_p(v), v.size)
    return v


def esh_momentum_update(gradient, momentum, step_size):
    """cpu_math.rs:505-551: returns (momentum', kinetic energy change)."""
    g, m = _f64(gradient), _f64(momentum).copy()
    dke = lib().orc_esh_momentum_update(_p(g), _p(m), step_size, g.size)
    return m, dke


def vector_dot(a, b):
    a = _f64(a)
    b = _f64(b)
    return lib().orc_vector_dot(_p(a), _p(b), a.size)


def scalar_prods3(p1, n1, p2, x, y):
    p1, n1, p2, x, y = map(_f64, (p1, n1, p2, x, y))
    o1 = C.c_double()
    o2 = C.c_double()
    lib().orc_scalar_prods3(_p(p1), _p(n1), _p(p2), _p(x), _p(y), p1.size, C.byref(o1), C.byref(o2))
    return o1.value, o2.value


def scalar_prods2(p1, p2, x, y):
    p1, p2, x, y = map(_f64, (p1, p2, x, y))
    o1 = C.c_double()
    o2 = C.c_double()
    lib().orc_scalar_prods2(_p(p1), _p(p2), _p(x), _p(y), p1.size, C.byref(o1), C.byref(o2))
    return o1.value, o2.value


def array_update_variance(mean, variance, value, diff_scale):
    mean = _f64(mean).copy()
    variance = _f64(variance).copy()
    value = _f64(value)
    lib().orc_array_update_variance(_p(mean), _p(variance), _p(value), diff_scale, mean.size)
    return mean, variance


def philox(seed, stream, counter):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(seed, stream, counter, out)
    return list(out)


def fill_normal(seed, stream, counter, d):
    out = np.empty(d, dtype=np.float64)
    new_counter = lib().orc_fill_normal(seed, stream, counter, _p(out), d)
    return out, new_counter


def det_sincos2pi(u):
    s = C.c_double()
    c = C.c_double()
    lib().orc_det_sincos2pi(u, C.byref(s), C.byref(c))
    return s.value, c.value


class Model:
    def __init__(self, kind, dim, mu=None, sigma=None, rank1_scale=0.0, funnel_scale=3.0):
        self.dim = dim
        self.desc, self._keep = _abi.make_logp_desc(kind, dim, mu, sigma, rank1_scale, funnel_scale)
        self.h = lib().orc_model_create(C.byref(self.desc), dim)
        assert self.h

    def logp(self, x):
        x = _f64(x)
        g = np.empty_like(x)
        lp = lib().orc_model_logp(self.h, _p(x), _p(g))
        return lp, g

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_destroy(self.h)
            self.h = None


class Point:
    X, GX, Z, GZ, V = range(5)

    def __init__(self, ham):
        self.ham = ham
        self.h = lib().orc_point_create(ham.h)

    def vec(self, which):
        out = np.empty(self.ham.dim)
        lib().orc_point_get_vec(self.h, which, _p(out))
        return out

    def set_vec(self, which, v):
        v = _f64(v)
        lib().orc_point_set_vec(self.h, which, _p(v))

    def scalars(self):
        idx = C.c_int64()
        tid = C.c_int64()
        logp, logdet, ke, e0 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        lib().orc_point_get_scalars(self.h, C.byref(idx), C.byref(logp), C.byref(logdet), C.byref(ke), C.byref(e0), C.byref(tid))
        return dict(index_in_trajectory=idx.value, logp=logp.value, logdet=logdet.value, kinetic_energy=ke.value,
                    initial_energy=e0.value, transform_id=tid.value)

    def set_scalars(self, index_in_trajectory, logp, logdet, kinetic_energy, initial_energy, transform_id):
        lib().orc_point_set_scalars(self.h, index_in_trajectory, logp, logdet, kinetic_energy, initial_energy, transform_id)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_point_destroy(self.h)
            self.h = None


class Hamiltonian:
    """TransformedHamiltonian<DiagMassMatrix> for ONE chain (white-box access)."""

    def __init__(self, model):
        self.model = model
        self.dim = model.dim
        self.h = lib().orc_ham_create(model.h)

    def set_kinetic_energy_kind(self, kind):
        """KineticEnergyKind (transformed_hamiltonian.rs:22-50): _abi.NUTS_KINETIC_*."""
        lib().orc_ham_set_kinetic_energy_kind(self.h, int(kind))

    def set_transform(self, stds, mean):
        stds, mean = _f64(stds), _f64(mean)
        lib().orc_ham_set_transform(self.h, _p(stds), _p(mean))

    def set_lowrank_transform(self, stds, mean, vals, vecs, mean_low_rank):
        """LowRankMassMatrix::update (src/transform/low_rank.rs:158-190).  vecs: [r, dim] (row k = eigenvector k).  Returns False when an
        input was not finite (the old transformation stays)."""
        stds, mean, vals, mu = _f64(stds), _f64(mean), _f64(vals), _f64(mean_low_rank)
        vecs = _f64(np.asarray(vecs, dtype=np.float64).reshape(len(vals), self.dim))
        return bool(lib().orc_ham_set_lowrank_transform(self.h, _p(stds), _p(mean), _p(vals), _p(vecs), _p(mu), len(vals)))

    def update_diag_draw_grad(self, draw_mean, grad_mean, draw_var, grad_var, fill=None, clamp=(1e-20, 1e20)):
        a = list(map(_f64, (draw_mean, grad_mean, draw_var, grad_var)))
        lib().orc_ham_update_diag_draw_grad(self.h, *map(_p, a), fill is not None, fill or 0.0, clamp[0], clamp[1])

    def update_diag_grad(self, position, gradient, fill=1.0, clamp=(1e-20, 1e20)):
        a = list(map(_f64, (position, gradient)))
        lib().orc_ham_update_diag_grad(self.h, *map(_p, a), fill, clamp[0], clamp[1])

    def update_diag_draw(self, draw_mean, draw_var, scale, fill=None, clamp=(1e-20, 1e20)):
        a = list(map(_f64, (draw_mean, draw_var)))
        lib().orc_ham_update_diag_draw(self.h, *map(_p, a), scale, fill is not None, fill or 0.0, clamp[0], clamp[1])

    def transform(self):
        d = self.dim
        stds, inv, mean = np.empty(d), np.empty(d), np.empty(d)
        logdet = C.c_double()
        tid = C.c_int64()
        lib().orc_ham_get_transform(self.h, _p(stds), _p(inv), _p(mean), C.byref(logdet), C.byref(tid))
        return dict(stds=stds, inv_stds=inv, mean=mean, logdet=logdet.value, id=tid.value)

    def new_point(self):
        return Point(self)

    def init_state(self, x):
        p = Point(self)
        x = _f64(x)
        st = lib().orc_ham_init_state(self.h, p.h, _p(x))
        return p, st

    def init_from_untransformed(self, p):
        lib().orc_ham_init_from_untransformed(self.h, p.h)

    def init_from_transformed(self, p):
        lib().orc_ham_init_from_transformed(self.h, p.h)

    def initialize_trajectory(self, p, resample, seed, stream, counter):
        return lib().orc_ham_initialize_trajectory(self.h, p.h, int(resample), seed, stream, counter)

    def leapfrog(self, start, step_size, direction=1, energy_baseline=None, max_energy_error=1000.0):
        out = Point(self)
        if energy_baseline is None:
            energy_baseline = start.scalars()["initial_energy"]
        ee = C.c_double()
        st = lib().orc_ham_leapfrog(self.h, start.h, out.h, step_size, direction, energy_baseline, max_energy_error, C.byref(ee))
        return out, st, ee.value

    def is_turning(self, p1, p2):
        return bool(lib().orc_ham_is_turning(self.h, p1.h, p2.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ham_destroy(self.h)
            self.h = None


def alloc_stats(n_draws, nchains):
    """Returns (Stats struct, dict of numpy arrays [n_draws, nchains])."""
    st = _abi.Stats()
    arrays = {}
    for name, dt in _abi.STAT_DTYPES.items():
        a = np.zeros((n_draws, nchains), dtype=dt)
        arrays[name] = a
        setattr(st, name, a.ctypes.data_as(dict(st._fields_)[name]))
    return st, arrays


class Sampler:
    """N independent NutsChains (one per RNG stream chain_id_offset + c + 1)."""

    def __init__(self, model, settings, seed, nchains, chain_id_offset=0, nthreads=1):
        self.model = model
        self.dim = model.dim
        self.nchains = nchains
        self.nthreads = nthreads
        self.settings = settings
        self.h = lib().orc_sampler_create(model.h, C.byref(settings), seed, chain_id_offset, nchains)

    def set_position(self, x):
        x = _f64(x).reshape(self.nchains, self.dim)
        status = np.zeros(self.nchains, dtype=np.int32)
        lib().orc_sampler_set_position(self.h, _p(x), status.ctypes.data_as(_abi.c_i32_p), self.nthreads)
        return status

    def draw(self, n_draws, want_draws=True):
        draws = np.empty((n_draws, self.nchains, self.dim)) if want_draws else None
        st, arrays = alloc_stats(n_draws, self.nchains)
        total = lib().orc_sampler_draw(self.h, n_draws, _p(draws) if want_draws else None, C.byref(st), self.nthreads)
        arrays["_total_leapfrogs"] = total
        return draws, arrays

    def state(self):
        N, d = self.nchains, self.dim
        pos, eps, stds, mean = np.empty((N, d)), np.empty(N), np.empty((N, d)), np.empty((N, d))
        ctr = np.empty(N, dtype=np.uint64)
        lib().orc_sampler_get_state(self.h, _p(pos), _p(eps), _p(stds), _p(mean), ctr.ctypes.data_as(_abi.c_u64_p))
        return dict(position=pos, step_size=eps, stds=stds, mean=mean, rng_counter=ctr)

    def set_lowrank_transform(self, stds, mean, vals, vecs, mean_low_rank, rank=None):
        """LowRankMassMatrix::update for every chain (vals [N, r], vecs [N, r, dim], rank [N]); the first update of a run re-runs the
        step size search.  Returns the per-chain accepted flags."""
        N, d = self.nchains, self.model.dim
        stds, mean, mu = (_f64(np.broadcast_to(a, (N, d))) for a in (stds, mean, mean_low_rank))
        vals = np.asarray(vals, dtype=np.float64)
        r = vals.shape[-1] if vals.ndim else 0
        vals = _f64(np.broadcast_to(vals.reshape((-1, r)) if vals.ndim < 2 else vals, (N, r)))
        vecs = np.asarray(vecs, dtype=np.float64)
        vecs = _f64(np.broadcast_to(vecs.reshape((-1, r, d)) if vecs.ndim < 3 else vecs, (N, r, d)))
        rk = None if rank is None else np.ascontiguousarray(np.broadcast_to(rank, (N,)), dtype=np.int32)
        ok = np.zeros(N, dtype=np.uint8)
        lib().orc_sampler_set_lowrank_transform(self.h, _p(stds), _p(mean), r, _p(vals) if r else None, _p(vecs) if r else None,
                                                None if rk is None else rk.ctypes.data_as(C.POINTER(C.c_int32)), _p(mu),
                                                ok.ctypes.data_as(C.POINTER(C.c_uint8)))
        return ok.astype(bool)

    def set_step_size(self, eps):
        eps = _f64(np.broadcast_to(eps, (self.nchains,)))
        lib().orc_sampler_set_step_size(self.h, _p(eps))

    def chain_state(self):
        st, arrays = _abi.alloc_chain_state(self.nchains, self.dim)
        lib().orc_sampler_get_chain_state(self.h, C.byref(st))
        return arrays

    def set_chain_state(self, arrays):
        st, keep = _abi.alloc_chain_state(self.nchains, self.dim, arrays)
        lib().orc_sampler_set_chain_state(self.h, C.byref(st))

    def counters(self):
        a = C.c_uint64()
        b = C.c_uint64()
        lib().orc_sampler_counters(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sampler_destroy(self.h)
            self.h = None
