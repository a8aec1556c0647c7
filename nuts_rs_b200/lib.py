"""ctypes loader + thin host-side mirror of the reference's interface for the accelerated path.

Names follow nuts-rs: `CudaMath` plays the role of a `Math` implementation (reference src/math/math.rs:15-314,
batched over chains), `DiagNutsSettings` is `nuts_rs::DiagNutsSettings` (src/sampler.rs:241-244), `Sampler.set_position`
/ `Sampler.draw` are `Chain::set_position` / `Chain::draw` (src/chain.rs:137-188) for every chain at once.
Everything goes through the C ABI of include/nuts_b200.h; there is no CPU fallback — a missing library or GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import c_double_p as dp

_HERE = os.path.dirname(os.path.abspath(__file__))
# NUTS_B200_LIB selects an instrumented build of the same library (e.g. the -DNB_PHASE_TIMING one) for experiments
LIB_PATH = os.environ.get("NUTS_B200_LIB") or os.path.join(_HERE, "libnuts_b200.so")
_LIB = None

EXPORTED_SYMBOLS = [
    "nuts_last_error", "nuts_device_available", "nuts_settings_default",
    "nuts_ctx_create", "nuts_ctx_destroy", "nuts_ctx_synchronize", "nuts_ctx_nchains", "nuts_ctx_dim", "nuts_ctx_stream",
    "nuts_plane_alloc", "nuts_plane_free", "nuts_plane_read_from_host", "nuts_plane_write_to_host", "nuts_plane_device_ptr",
    "nuts_axpy", "nuts_axpy_out", "nuts_array_mult", "nuts_array_mult_inplace", "nuts_array_recip", "nuts_fill_array",
    "nuts_copy_into", "nuts_array_vector_dot", "nuts_scalar_prods3", "nuts_scalar_prods2", "nuts_sq_norm_sum",
    "nuts_array_all_finite", "nuts_array_all_finite_and_nonzero", "nuts_array_sum_ln", "nuts_array_gaussian",
    "nuts_array_update_variance", "nuts_array_update_var_inv_std_draw", "nuts_array_update_var_inv_std_draw_grad",
    "nuts_array_update_var_inv_std_grad", "nuts_logp_array",
    "nuts_point_alloc", "nuts_point_free", "nuts_point_plane", "nuts_point_get_scalars", "nuts_point_set_scalars",
    "nuts_set_transform", "nuts_get_transform", "nuts_init_state", "nuts_initialize_trajectory", "nuts_leapfrog", "nuts_is_turning",
    "nuts_sampler_create", "nuts_sampler_destroy", "nuts_set_position", "nuts_draw", "nuts_draw_device",
    "nuts_sampler_counters", "nuts_sampler_last_timing", "nuts_sampler_get_state", "nuts_sampler_set_step_size",
    "nuts_host_alloc", "nuts_host_free", "nuts_sampler_last_draw_direct",
    "nuts_sampler_get_chain_state", "nuts_sampler_set_chain_state",
    "nuts_sampler_create_lowrank", "nuts_sampler_set_lowrank_transform", "nuts_sampler_set_grads_out",
    "nuts_eigs_create", "nuts_eigs_free", "nuts_apply_lowrank_transform", "nuts_apply_lowrank_transform_inplace", "nuts_set_lowrank_transform",
    "nuts_std_norm_flow", "nuts_std_norm_grad_flow", "nuts_std_norm_grad_flow_inplace", "nuts_array_normalize", "nuts_esh_momentum_update",
    "nuts_leapfrog_kinetic", "nuts_initialize_trajectory_kinetic", "nuts_ctx_last_kernel_ms",
    "nuts_set_position_masked", "nuts_comm_unique_id", "nuts_comm_create", "nuts_comm_destroy", "nuts_gather_draws_begin", "nuts_gather_draws_end",
]


class NutsError(RuntimeError):
    pass


def load():
    """Load libnuts_b200.so (built by `make -C nuts_rs_b200/csrc` / __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NutsError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nuts_last_error.restype = C.c_char_p
    L.nuts_ctx_nchains.restype = C.c_uint64
    L.nuts_ctx_dim.restype = C.c_uint64
    L.nuts_ctx_stream.restype = vp
    L.nuts_plane_device_ptr.restype = vp
    L.nuts_point_plane.restype = vp
    L.nuts_settings_default.restype = None
    L.nuts_settings_default.argtypes = [C.POINTER(_abi.NutsSettings)]
    L.nuts_ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint64, C.c_uint64, C.POINTER(_abi.LogpDesc)]
    L.nuts_ctx_destroy.argtypes = [vp]
    L.nuts_ctx_synchronize.argtypes = [vp]
    L.nuts_ctx_nchains.argtypes = [vp]
    L.nuts_ctx_dim.argtypes = [vp]
    L.nuts_ctx_stream.argtypes = [vp]
    L.nuts_plane_alloc.argtypes = [vp, C.POINTER(vp)]
    L.nuts_plane_free.argtypes = [vp, vp]
    L.nuts_plane_read_from_host.argtypes = [vp, vp, dp]
    L.nuts_plane_write_to_host.argtypes = [vp, vp, dp]
    L.nuts_plane_device_ptr.argtypes = [vp, _abi.c_u64_p]
    L.nuts_axpy.argtypes = [vp, vp, vp, dp, C.c_double, _abi.c_u8_p]
    L.nuts_axpy_out.argtypes = [vp, vp, vp, dp, C.c_double, vp, _abi.c_u8_p]
    L.nuts_array_mult.argtypes = [vp, vp, vp, vp]
    L.nuts_std_norm_flow.argtypes = [vp, vp, vp, vp, dp, C.c_double, _abi.c_u8_p]
    L.nuts_std_norm_grad_flow.argtypes = [vp, vp, vp, vp, vp, dp, C.c_double, _abi.c_u8_p]
    L.nuts_std_norm_grad_flow_inplace.argtypes = [vp, vp, vp, vp, dp, C.c_double, _abi.c_u8_p]
    L.nuts_array_normalize.argtypes = [vp, vp, _abi.c_u8_p]
    L.nuts_esh_momentum_update.argtypes = [vp, vp, vp, dp, C.c_double, _abi.c_u8_p, dp]
    L.nuts_ctx_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.nuts_leapfrog_kinetic.argtypes = [vp, C.c_int, vp, vp, dp, C.c_double, _abi.c_i8_p, dp, C.c_double, _abi.c_u8_p, _abi.c_i32_p, dp]
    L.nuts_initialize_trajectory_kinetic.argtypes = [vp, C.c_int, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
    L.nuts_array_mult_inplace.argtypes = [vp, vp, vp]
    L.nuts_array_recip.argtypes = [vp, vp, vp]
    L.nuts_fill_array.argtypes = [vp, vp, C.c_double]
    L.nuts_copy_into.argtypes = [vp, vp, vp]
    L.nuts_array_vector_dot.argtypes = [vp, vp, vp, dp]
    L.nuts_scalar_prods3.argtypes = [vp, vp, vp, vp, vp, vp, dp, dp]
    L.nuts_scalar_prods2.argtypes = [vp, vp, vp, vp, vp, dp, dp]
    L.nuts_sq_norm_sum.argtypes = [vp, vp, vp, dp]
    L.nuts_array_all_finite.argtypes = [vp, vp, _abi.c_u8_p]
    L.nuts_array_all_finite_and_nonzero.argtypes = [vp, vp, _abi.c_u8_p]
    L.nuts_array_sum_ln.argtypes = [vp, vp, dp]
    L.nuts_array_gaussian.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.nuts_array_update_variance.argtypes = [vp, vp, vp, vp, dp, C.c_double]
    L.nuts_array_update_var_inv_std_draw.argtypes = [vp, vp, vp, vp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
    L.nuts_array_update_var_inv_std_draw_grad.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.nuts_array_update_var_inv_std_grad.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, C.c_double]
    L.nuts_logp_array.argtypes = [vp, vp, vp, dp, _abi.c_i32_p]
    L.nuts_point_alloc.argtypes = [vp, C.POINTER(vp)]
    L.nuts_point_free.argtypes = [vp, vp]
    L.nuts_point_plane.argtypes = [vp, C.c_int]
    L.nuts_point_get_scalars.argtypes = [vp, vp, _abi.c_i64_p, dp, dp, dp, dp, _abi.c_i64_p]
    L.nuts_point_set_scalars.argtypes = [vp, vp, _abi.c_i64_p, dp, dp, dp, dp, _abi.c_i64_p]
    L.nuts_set_transform.argtypes = [vp, dp, dp]
    L.nuts_get_transform.argtypes = [vp, dp, dp, dp, dp, _abi.c_i64_p]
    L.nuts_eigs_create.argtypes = [vp, C.POINTER(vp), C.c_uint64, dp, dp, _abi.c_i32_p]
    L.nuts_sampler_create_lowrank.argtypes = [vp, C.POINTER(vp), C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    L.nuts_sampler_set_lowrank_transform.argtypes = [vp, dp, dp, C.c_uint64, dp, dp, _abi.c_i32_p, dp, C.POINTER(C.c_uint8)]
    L.nuts_sampler_set_grads_out.argtypes = [vp, C.c_void_p]
    L.nuts_eigs_free.argtypes = [vp, vp]
    L.nuts_apply_lowrank_transform.argtypes = [vp, vp, vp, vp]
    L.nuts_apply_lowrank_transform_inplace.argtypes = [vp, vp, vp]
    L.nuts_set_lowrank_transform.argtypes = [vp, dp, dp, C.c_uint64, dp, dp, _abi.c_i32_p, dp, C.POINTER(C.c_uint8)]
    L.nuts_init_state.argtypes = [vp, vp, dp, _abi.c_i32_p]
    L.nuts_initialize_trajectory.argtypes = [vp, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
    L.nuts_leapfrog.argtypes = [vp, vp, vp, dp, C.c_double, _abi.c_i8_p, dp, C.c_double, _abi.c_u8_p, _abi.c_i32_p, dp]
    L.nuts_is_turning.argtypes = [vp, vp, vp, _abi.c_u8_p]
    L.nuts_sampler_create.argtypes = [vp, C.POINTER(vp), C.POINTER(_abi.NutsSettings), C.c_uint64, C.c_uint64]
    L.nuts_sampler_destroy.argtypes = [vp]
    L.nuts_set_position.argtypes = [vp, dp, _abi.c_i32_p]
    L.nuts_set_position_masked.argtypes = [vp, dp, _abi.c_u8_p, _abi.c_i32_p]
    L.nuts_draw.argtypes = [vp, C.c_uint64, dp, C.POINTER(_abi.Stats)]
    L.nuts_draw_device.argtypes = [vp, C.c_uint64, vp]
    L.nuts_sampler_counters.argtypes = [vp, _abi.c_u64_p, _abi.c_u64_p]
    L.nuts_sampler_last_timing.argtypes = [vp, dp, _abi.c_u64_p]
    L.nuts_sampler_get_state.argtypes = [vp, dp, dp, dp, dp, _abi.c_u64_p]
    L.nuts_sampler_set_step_size.argtypes = [vp, dp]
    L.nuts_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.nuts_host_free.argtypes = [vp]
    L.nuts_sampler_last_draw_direct.argtypes = [vp, _abi.c_i32_p]
    L.nuts_sampler_get_chain_state.argtypes = [vp, C.POINTER(_abi.ChainState)]
    L.nuts_sampler_set_chain_state.argtypes = [vp, C.POINTER(_abi.ChainState)]
    L.nuts_comm_unique_id.argtypes = [_abi.c_u8_p]
    L.nuts_comm_create.argtypes = [C.POINTER(vp), C.c_int, _abi.c_u8_p, C.c_int, C.c_int]
    L.nuts_comm_destroy.argtypes = [vp]
    L.nuts_gather_draws_begin.argtypes = [vp, vp, vp, vp, C.c_uint64]
    L.nuts_gather_draws_end.argtypes = [vp, dp]
    _LIB = L
    return L


def device_available() -> bool:
    return load().nuts_device_available() == 0


def _check(rc):
    if rc != 0:
        raise NutsError(f"libnuts_b200 error {rc}: {load().nuts_last_error().decode()}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def DiagNutsSettings(**overrides) -> _abi.NutsSettings:
    """`DiagNutsSettings::default()` from the library itself, with top-level field overrides (num_tune=..., maxdepth=...)."""
    s = _abi.NutsSettings()
    load().nuts_settings_default(C.byref(s))
    for k, v in overrides.items():
        setattr(s, k, v)
    return s


class Eigs:
    """EigVectors + EigValues of all chains on the device (reference src/math/math.rs:17-18, 150-160)."""

    def __init__(self, math, vecs, vals, rank=None):
        N, d = math.nchains, math.dim
        vals = np.asarray(vals, dtype=np.float64)
        r = vals.shape[-1]
        vals = _f64(np.broadcast_to(vals.reshape((-1, r)) if vals.ndim < 2 else vals, (N, r)))
        vecs = np.asarray(vecs, dtype=np.float64)
        vecs = _f64(np.broadcast_to(vecs.reshape((-1, r, d)) if vecs.ndim < 3 else vecs, (N, r, d)))
        rk = None if rank is None else np.ascontiguousarray(np.broadcast_to(rank, (N,)), dtype=np.int32)
        self.math = math
        h = C.c_void_p()
        _check(load().nuts_eigs_create(math.h, C.byref(h), r, _p(vecs), _p(vals), None if rk is None else rk.ctypes.data_as(_abi.c_i32_p)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and self.math.h:
            load().nuts_eigs_free(self.math.h, self.h)
            self.h = None


class Plane:
    """`nchains` x Math::Vector on the device."""

    def __init__(self, math, handle=None, owned=True):
        self.math = math
        self.owned = owned
        if handle is None:
            h = C.c_void_p()
            _check(load().nuts_plane_alloc(math.h, C.byref(h)))
            handle = h
        self.h = handle

    def read_from_slice(self, src):
        src = _f64(src).reshape(self.math.nchains, self.math.dim)
        _check(load().nuts_plane_read_from_host(self.math.h, self.h, _p(src)))
        return self

    def device_ptr(self):
        """(device address, row stride in doubles): rows are padded - chain c starts at address + 8 * c * stride
        (nuts_plane_device_ptr; the way to share the plane with torch or another CUDA owner)."""
        stride = C.c_uint64(0)
        ptr = load().nuts_plane_device_ptr(self.h, C.byref(stride))
        return int(ptr or 0), int(stride.value)

    def box_array(self):
        out = np.empty((self.math.nchains, self.math.dim))
        _check(load().nuts_plane_write_to_host(self.math.h, self.h, _p(out)))
        return out

    def __del__(self):
        if getattr(self, "owned", False) and getattr(self, "h", None) and self.math.h:
            load().nuts_plane_free(self.math.h, self.h)
            self.h = None


class Point:
    """`nchains` x TransformedPoint (reference src/dynamics/transformed_hamiltonian.rs:56-77)."""

    X, GX, Z, GZ, V = range(5)

    def __init__(self, math):
        self.math = math
        h = C.c_void_p()
        _check(load().nuts_point_alloc(math.h, C.byref(h)))
        self.h = h

    def plane(self, which):
        return Plane(self.math, C.c_void_p(load().nuts_point_plane(self.h, which)), owned=False)

    def vec(self, which):
        return self.plane(which).box_array()

    def set_vec(self, which, v):
        self.plane(which).read_from_slice(v)

    def scalars(self):
        N = self.math.nchains
        idx, tid = np.zeros(N, dtype=np.int64), np.zeros(N, dtype=np.int64)
        logp, logdet, ke, e0 = (np.zeros(N) for _ in range(4))
        _check(load().nuts_point_get_scalars(self.math.h, self.h, idx.ctypes.data_as(_abi.c_i64_p), _p(logp), _p(logdet), _p(ke), _p(e0),
                                             tid.ctypes.data_as(_abi.c_i64_p)))
        return dict(index_in_trajectory=idx, logp=logp, logdet=logdet, kinetic_energy=ke, initial_energy=e0, transform_id=tid)

    def set_scalars(self, index_in_trajectory=None, logp=None, logdet=None, kinetic_energy=None, initial_energy=None, transform_id=None):
        N = self.math.nchains

        def f(a):
            return None if a is None else _f64(np.broadcast_to(a, (N,)))

        def i(a):
            return None if a is None else np.ascontiguousarray(np.broadcast_to(a, (N,)), dtype=np.int64)

        idx, tid = i(index_in_trajectory), i(transform_id)
        a = [f(logp), f(logdet), f(kinetic_energy), f(initial_energy)]
        _check(load().nuts_point_set_scalars(self.math.h, self.h, None if idx is None else idx.ctypes.data_as(_abi.c_i64_p), *map(_p, a),
                                             None if tid is None else tid.ctypes.data_as(_abi.c_i64_p)))

    def __del__(self):
        if getattr(self, "h", None) and self.math.h:
            load().nuts_point_free(self.math.h, self.h)
            self.h = None


class CudaMath:
    """Batched `Math` backend on one B200: `nchains` independent chains of dimension `dim` with a device-side logp."""

    def __init__(self, nchains, dim, kind, mu=None, sigma=None, rank1_scale=0.0, funnel_scale=3.0, device=0, user_params=None):
        self.h = None
        self.nchains = int(nchains)
        self._dim = int(dim)
        self.desc, self._keep = _abi.make_logp_desc(kind, dim, mu, sigma, rank1_scale, funnel_scale, user_params)
        h = C.c_void_p()
        _check(load().nuts_ctx_create(C.byref(h), device, nchains, dim, C.byref(self.desc)))
        self.h = h

    @property
    def dim(self):
        return self._dim

    def new_array(self):
        return Plane(self)

    def from_host(self, a):
        return Plane(self).read_from_slice(a)

    # -- Tier 1 --
    def _mask(self, active):
        if active is None:
            return None, None
        m = np.ascontiguousarray(active, dtype=np.uint8)
        return m, m.ctypes.data_as(_abi.c_u8_p)

    def _scal(self, a):
        if np.isscalar(a):
            return None, None, float(a)
        arr = _f64(a)
        return arr, _p(arr), 0.0

    def axpy(self, x, y, a, active=None):
        keep, ap, ab = self._scal(a)
        m, mp = self._mask(active)
        _check(load().nuts_axpy(self.h, x.h, y.h, ap, ab, mp))

    def axpy_out(self, x, y, a, out, active=None):
        keep, ap, ab = self._scal(a)
        m, mp = self._mask(active)
        _check(load().nuts_axpy_out(self.h, x.h, y.h, ap, ab, out.h, mp))

    def std_norm_flow(self, pos, pos_out, vel, epsilon, active=None):
        """Math::std_norm_flow (math.rs:155-161): pos_out = pos cos(eps) + vel sin(eps); vel = -pos sin(eps) + vel cos(eps)."""
        keep, ep, eb = self._scal(epsilon)
        m, mp = self._mask(active)
        _check(load().nuts_std_norm_flow(self.h, pos.h, pos_out.h, vel.h, ep, eb, mp))

    def std_norm_grad_flow(self, pos, grad, vel, vel_out, epsilon, active=None):
        """Math::std_norm_grad_flow (math.rs:162-169): vel_out = vel + eps * (pos + grad)."""
        keep, ep, eb = self._scal(epsilon)
        m, mp = self._mask(active)
        _check(load().nuts_std_norm_grad_flow(self.h, pos.h, grad.h, vel.h, vel_out.h, ep, eb, mp))

    def std_norm_grad_flow_inplace(self, pos, grad, vel, epsilon, active=None):
        keep, ep, eb = self._scal(epsilon)
        m, mp = self._mask(active)
        _check(load().nuts_std_norm_grad_flow_inplace(self.h, pos.h, grad.h, vel.h, ep, eb, mp))

    def array_normalize(self, v, active=None):
        m, mp = self._mask(active)
        _check(load().nuts_array_normalize(self.h, v.h, mp))

    def esh_momentum_update(self, gradient, momentum, step_size, active=None):
        """Math::esh_momentum_update (math.rs:183-210): momentum updated in place; returns the kinetic-energy change per chain."""
        keep, sp, sb = self._scal(step_size)
        m, mp = self._mask(active)
        out = np.zeros(self.nchains)
        _check(load().nuts_esh_momentum_update(self.h, gradient.h, momentum.h, sp, sb, mp, _p(out)))
        return out

    def array_mult(self, a1, a2, dest):
        _check(load().nuts_array_mult(self.h, a1.h, a2.h, dest.h))

    def array_mult_inplace(self, a1, a2):
        _check(load().nuts_array_mult_inplace(self.h, a1.h, a2.h))

    def array_recip(self, a, dest):
        _check(load().nuts_array_recip(self.h, a.h, dest.h))

    def fill_array(self, a, val):
        _check(load().nuts_fill_array(self.h, a.h, val))

    def copy_into(self, src, dst):
        _check(load().nuts_copy_into(self.h, src.h, dst.h))

    def _red(self, fn, *planes):
        out = np.empty(self.nchains)
        _check(fn(self.h, *[p.h for p in planes], _p(out)))
        return out

    def array_vector_dot(self, a1, a2):
        return self._red(load().nuts_array_vector_dot, a1, a2)

    def sq_norm_sum(self, x, y):
        return self._red(load().nuts_sq_norm_sum, x, y)

    def array_sum_ln(self, a):
        return self._red(load().nuts_array_sum_ln, a)

    def scalar_prods3(self, positive1, negative1, positive2, x, y):
        o1, o2 = np.empty(self.nchains), np.empty(self.nchains)
        _check(load().nuts_scalar_prods3(self.h, positive1.h, negative1.h, positive2.h, x.h, y.h, _p(o1), _p(o2)))
        return o1, o2

    def scalar_prods2(self, positive1, positive2, x, y):
        o1, o2 = np.empty(self.nchains), np.empty(self.nchains)
        _check(load().nuts_scalar_prods2(self.h, positive1.h, positive2.h, x.h, y.h, _p(o1), _p(o2)))
        return o1, o2

    def array_all_finite(self, a):
        out = np.zeros(self.nchains, dtype=np.uint8)
        _check(load().nuts_array_all_finite(self.h, a.h, out.ctypes.data_as(_abi.c_u8_p)))
        return out.astype(bool)

    def array_all_finite_and_nonzero(self, a):
        out = np.zeros(self.nchains, dtype=np.uint8)
        _check(load().nuts_array_all_finite_and_nonzero(self.h, a.h, out.ctypes.data_as(_abi.c_u8_p)))
        return out.astype(bool)

    def array_gaussian(self, dest, stds, seed, chain_offset, counter):
        _check(load().nuts_array_gaussian(self.h, dest.h, stds.h, seed, chain_offset, counter))

    def array_update_variance(self, mean, variance, value, diff_scale):
        keep, ap, ab = self._scal(diff_scale)
        _check(load().nuts_array_update_variance(self.h, mean.h, variance.h, value.h, ap, ab))

    def array_update_var_inv_std_draw(self, inv_std, std, draw_var, scale, fill_invalid, clamp):
        _check(load().nuts_array_update_var_inv_std_draw(self.h, inv_std.h, std.h, draw_var.h, scale, fill_invalid is not None,
                                                         fill_invalid or 0.0, clamp[0], clamp[1]))

    def array_update_var_inv_std_draw_grad(self, inv_std, std, draw_var, grad_var, fill_invalid, clamp):
        _check(load().nuts_array_update_var_inv_std_draw_grad(self.h, inv_std.h, std.h, draw_var.h, grad_var.h, fill_invalid is not None,
                                                              fill_invalid or 0.0, clamp[0], clamp[1]))

    def array_update_var_inv_std_grad(self, inv_std, std, gradient, fill_invalid, clamp):
        _check(load().nuts_array_update_var_inv_std_grad(self.h, inv_std.h, std.h, gradient.h, fill_invalid, clamp[0], clamp[1]))

    def logp_array(self, position, gradient):
        logp = np.empty(self.nchains)
        status = np.zeros(self.nchains, dtype=np.int32)
        _check(load().nuts_logp_array(self.h, position.h, gradient.h, _p(logp), status.ctypes.data_as(_abi.c_i32_p)))
        return logp, status

    # -- Tier 2 (Hamiltonian) --
    def new_point(self):
        return Point(self)

    def set_transform(self, stds, mean):
        stds = _f64(np.broadcast_to(stds, (self.nchains, self.dim)))
        mean = _f64(np.broadcast_to(mean, (self.nchains, self.dim)))
        _check(load().nuts_set_transform(self.h, _p(stds), _p(mean)))

    def set_lowrank_transform(self, stds, mean, vals, vecs, mean_low_rank, rank=None):
        """LowRankMassMatrix::update (reference src/transform/low_rank.rs:158-190) for every chain.  vals: [N, r] (or [r], the same
        for every chain), vecs: [N, r, dim] (or [r, dim]), rank: [N] eigenvectors actually used per chain (default r).  Returns the
        per-chain `accepted` flags (False: a non-finite input, that chain keeps its old transformation)."""
        N, d = self.nchains, self.dim
        stds = _f64(np.broadcast_to(stds, (N, d)))
        mean = _f64(np.broadcast_to(mean, (N, d)))
        mu = _f64(np.broadcast_to(mean_low_rank, (N, d)))
        vals = np.asarray(vals, dtype=np.float64)
        r = vals.shape[-1] if vals.ndim else 0
        vals = _f64(np.broadcast_to(vals.reshape((-1, r)) if vals.ndim < 2 else vals, (N, r)))
        vecs = np.asarray(vecs, dtype=np.float64)
        vecs = _f64(np.broadcast_to(vecs.reshape((-1, r, d)) if vecs.ndim < 3 else vecs, (N, r, d)))
        rk = None if rank is None else np.ascontiguousarray(np.broadcast_to(rank, (N,)), dtype=np.int32)
        ok = np.zeros(N, dtype=np.uint8)
        _check(load().nuts_set_lowrank_transform(self.h, _p(stds), _p(mean), r, _p(vals) if r else None, _p(vecs) if r else None,
                                                 None if rk is None else rk.ctypes.data_as(_abi.c_i32_p), _p(mu),
                                                 ok.ctypes.data_as(C.POINTER(C.c_uint8))))
        return ok.astype(bool)

    def new_eigs(self, vecs, vals, rank=None):
        """Math::new_eig_vectors + new_eig_values for every chain: vecs [N, r, dim] (or [r, dim]), vals [N, r] (or [r])."""
        return Eigs(self, vecs, vals, rank)

    def apply_lowrank_transform(self, eigs, rhs, dest):
        _check(load().nuts_apply_lowrank_transform(self.h, eigs.h, rhs.h, dest.h))

    def apply_lowrank_transform_inplace(self, eigs, rhs_and_dest):
        _check(load().nuts_apply_lowrank_transform_inplace(self.h, eigs.h, rhs_and_dest.h))

    def transform(self):
        N, d = self.nchains, self.dim
        stds, inv, mean = np.empty((N, d)), np.empty((N, d)), np.empty((N, d))
        logdet = np.empty(N)
        tid = np.empty(N, dtype=np.int64)
        _check(load().nuts_get_transform(self.h, _p(stds), _p(inv), _p(mean), _p(logdet), tid.ctypes.data_as(_abi.c_i64_p)))
        return dict(stds=stds, inv_stds=inv, mean=mean, logdet=logdet, id=tid)

    def init_state(self, position):
        p = Point(self)
        position = _f64(position).reshape(self.nchains, self.dim)
        status = np.zeros(self.nchains, dtype=np.int32)
        _check(load().nuts_init_state(self.h, p.h, _p(position), status.ctypes.data_as(_abi.c_i32_p)))
        return p, status

    def initialize_trajectory(self, point, resample, seed, chain_offset, counter, kind=_abi.NUTS_KINETIC_EUCLIDEAN):
        _check(load().nuts_initialize_trajectory_kinetic(self.h, int(kind), point.h, int(resample), seed, chain_offset, counter))

    def leapfrog(self, start, step_size, direction=None, energy_baseline=None, max_energy_error=1000.0, active=None, out=None,
                 kind=_abi.NUTS_KINETIC_EUCLIDEAN):
        """Hamiltonian::leapfrog for every chain; kind = KineticEnergyKind (_abi.NUTS_KINETIC_*)."""
        out = out or Point(self)
        keep, sp, sb = self._scal(step_size)
        d8 = None if direction is None else np.ascontiguousarray(np.broadcast_to(direction, (self.nchains,)), dtype=np.int8)
        base = None if energy_baseline is None else _f64(np.broadcast_to(energy_baseline, (self.nchains,)))
        m, mp = self._mask(active)
        status = np.zeros(self.nchains, dtype=np.int32)
        ee = np.empty(self.nchains)
        _check(load().nuts_leapfrog_kinetic(self.h, int(kind), start.h, out.h, sp, sb, None if d8 is None else d8.ctypes.data_as(_abi.c_i8_p),
                                            _p(base), max_energy_error, mp, status.ctypes.data_as(_abi.c_i32_p), _p(ee)))
        return out, status, ee

    def last_kernel_ms(self):
        """Device time of the kernel of the last leapfrog() call (CUDA events on the context's stream)."""
        ms = C.c_float()
        _check(load().nuts_ctx_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def is_turning(self, p1, p2):
        out = np.zeros(self.nchains, dtype=np.uint8)
        _check(load().nuts_is_turning(self.h, p1.h, p2.h, out.ctypes.data_as(_abi.c_u8_p)))
        return out.astype(bool)

    def synchronize(self):
        _check(load().nuts_ctx_synchronize(self.h))

    def close(self):
        if self.h:
            load().nuts_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        pass  # explicit close(); planes / points may outlive python GC order


class HostBuffer:
    """Page-locked, device-mapped host memory (nuts_host_alloc) viewed as a numpy f64 array.  Passed as `out=` to
    Sampler.draw it is written directly by the draw kernel (no staging copy).  Keep the object alive while `array` is used."""

    def __init__(self, shape, dtype=np.float64):
        self.ptr = C.c_void_p()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _check(load().nuts_host_alloc(C.byref(self.ptr), max(n, 1)))
        buf = (C.c_char * max(n, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            load().nuts_host_free(self.ptr)
            self.ptr = None


def alloc_stats(n_draws, nchains, names=None):
    st = _abi.Stats()
    arrays = {}
    ftypes = dict(st._fields_)
    for name, dt in _abi.STAT_DTYPES.items():
        if names is not None and name not in names:
            continue
        a = np.zeros((n_draws, nchains), dtype=dt)
        arrays[name] = a
        setattr(st, name, a.ctypes.data_as(ftypes[name]))
    return st, arrays


class Sampler:
    """All chains of one GPU: `settings.new_chain(...)` + `Chain::set_position` / `Chain::draw` for each of them."""

    def __init__(self, math, settings, seed, chain_id_offset=0, lowrank_rank_max=0):
        """lowrank_rank_max > 0: an engine with the low-rank transformation compiled in (nuts_sampler_create_lowrank); the
        transformation is installed with set_lowrank_transform (see nuts_rs_b200/lowrank.py for the host-side estimator)."""
        self.math = math
        self.nchains, self.dim = math.nchains, math.dim
        self.settings = settings
        self.chain_id_offset = chain_id_offset
        self.lowrank_rank_max = int(lowrank_rank_max)
        self._grads = None
        h = C.c_void_p()
        if lowrank_rank_max:
            _check(load().nuts_sampler_create_lowrank(math.h, C.byref(h), C.cast(C.byref(settings), C.c_void_p), seed, chain_id_offset, int(lowrank_rank_max)))
        else:
            _check(load().nuts_sampler_create(math.h, C.byref(h), C.byref(settings), seed, chain_id_offset))
        self.h = h

    def set_lowrank_transform(self, stds, mean, vals, vecs, mean_low_rank, rank=None):
        """LowRankMassMatrix::update (reference src/transform/low_rank.rs:158-190) for every chain: vals [N, r], vecs [N, r, dim],
        rank [N] (eigenvectors actually used per chain).  Returns the accepted flags (False: non-finite input, old transformation kept)."""
        N, d = self.nchains, self.dim
        stds, mean, mu = (_f64(np.broadcast_to(a, (N, d))) for a in (stds, mean, mean_low_rank))
        vals = np.asarray(vals, dtype=np.float64)
        r = vals.shape[-1] if vals.ndim else 0
        vals = _f64(np.broadcast_to(vals.reshape((-1, r)) if vals.ndim < 2 else vals, (N, r)))
        vecs = np.asarray(vecs, dtype=np.float64)
        vecs = _f64(np.broadcast_to(vecs.reshape((-1, r, d)) if vecs.ndim < 3 else vecs, (N, r, d)))
        rk = None if rank is None else np.ascontiguousarray(np.broadcast_to(rank, (N,)), dtype=np.int32)
        ok = np.zeros(N, dtype=np.uint8)
        _check(load().nuts_sampler_set_lowrank_transform(self.h, _p(stds), _p(mean), r, _p(vals) if r else None, _p(vecs) if r else None,
                                                         None if rk is None else rk.ctypes.data_as(_abi.c_i32_p), _p(mu),
                                                         ok.ctypes.data_as(C.POINTER(C.c_uint8))))
        return ok.astype(bool)

    def set_grads_out(self, host_buffer):
        """The gradient of logp at every draw of the following draw calls lands in `host_buffer` (a HostBuffer of shape
        [n_draws, nchains, dim]); None switches it off."""
        self._grads = host_buffer
        _check(load().nuts_sampler_set_grads_out(self.h, None if host_buffer is None else host_buffer.ptr))

    def set_position(self, position):
        position = _f64(position).reshape(self.nchains, self.dim)
        status = np.zeros(self.nchains, dtype=np.int32)
        _check(load().nuts_set_position(self.h, _p(position), status.ctypes.data_as(_abi.c_i32_p)))
        return status

    def set_position_with_retries(self, init_position, max_tries=500):
        """The reference's chain start (src/sampler.rs:1133-1143): `init_position(chain_ids) -> [len(chain_ids), dim]` is asked for a
        fresh starting point for every chain whose set_position failed (NutsError::BadInitGrad), up to `max_tries` times.
        Returns (status, tries): status 3 is left for chains that never found a valid point."""
        ids = np.arange(self.nchains)
        position = _f64(init_position(ids + self.chain_id_offset)).reshape(self.nchains, self.dim)
        status = self.set_position(position)
        tries = 1
        while (status != 0).any() and tries < max_tries:
            bad = np.nonzero(status != 0)[0]
            position[bad] = _f64(init_position(bad + self.chain_id_offset)).reshape(len(bad), self.dim)
            mask = np.zeros(self.nchains, dtype=np.uint8)
            mask[bad] = 1
            _check(load().nuts_set_position_masked(self.h, _p(position), mask.ctypes.data_as(_abi.c_u8_p), status.ctypes.data_as(_abi.c_i32_p)))
            tries += 1
        return status, tries

    def draw(self, n_draws, want_draws=True, stats=True, out=None):
        draws = None
        if want_draws:
            if out is not None:
                # the kernel (or the D2H copy) writes n_draws * nchains * dim doubles straight to this pointer
                if not isinstance(out, np.ndarray) or out.shape != (n_draws, self.nchains, self.dim) or out.dtype != np.float64 \
                        or not out.flags.c_contiguous or not out.flags.writeable:
                    raise NutsError(f"Sampler.draw: `out` must be a writeable C-contiguous float64 array of shape "
                                    f"{(n_draws, self.nchains, self.dim)}")
            draws = out if out is not None else np.empty((n_draws, self.nchains, self.dim))
        st, arrays = (alloc_stats(n_draws, self.nchains) if stats else (None, {}))
        _check(load().nuts_draw(self.h, n_draws, _p(draws), C.byref(st) if stats else None))
        return draws, arrays

    def draw_trace(self, n_draws, out=None):
        """`draw`, returned in the reference's trace schema ([chain, draw, ...], reference statistic names): see trace.py."""
        from . import trace

        done = self.counters()[1]
        draws, stats = self.draw(n_draws, out=out)
        return trace.to_trace(draws, stats, chain_offset=self.chain_id_offset, draw_offset=done)

    def draw_device(self, n_draws, draws_dev_ptr=None):
        _check(load().nuts_draw_device(self.h, n_draws, draws_dev_ptr))

    def counters(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(load().nuts_sampler_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_draw_direct(self):
        v = C.c_int32()
        _check(load().nuts_sampler_last_draw_direct(self.h, C.byref(v)))
        return bool(v.value)

    def last_timing(self):
        ms, n = C.c_double(), C.c_uint64()
        _check(load().nuts_sampler_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def state(self):
        N, d = self.nchains, self.dim
        pos, eps, stds, mean = np.empty((N, d)), np.empty(N), np.empty((N, d)), np.empty((N, d))
        ctr = np.empty(N, dtype=np.uint64)
        _check(load().nuts_sampler_get_state(self.h, _p(pos), _p(eps), _p(stds), _p(mean), ctr.ctypes.data_as(_abi.c_u64_p)))
        return dict(position=pos, step_size=eps, stds=stds, mean=mean, rng_counter=ctr)

    def set_step_size(self, eps):
        eps = _f64(np.broadcast_to(eps, (self.nchains,)))
        _check(load().nuts_sampler_set_step_size(self.h, _p(eps)))

    def chain_state(self):
        """Everything the chains carry from one draw to the next (nuts_chain_state_t) as a dict of numpy arrays: a checkpoint."""
        st, arrays = _abi.alloc_chain_state(self.nchains, self.dim)
        _check(load().nuts_sampler_get_chain_state(self.h, C.byref(st)))
        return arrays

    def set_chain_state(self, arrays):
        """Resume from a checkpoint taken with chain_state() (same model, settings, seed and chain_id_offset)."""
        st, keep = _abi.alloc_chain_state(self.nchains, self.dim, arrays)
        _check(load().nuts_sampler_set_chain_state(self.h, C.byref(st)))

    def close(self):
        if self.h:
            load().nuts_sampler_destroy(self.h)
            self.h = None


class Comm:
    """NCCL communicator of the draw gather (nuts_comm_*): one per process / GPU.  `exchange(id_bytes_or_None) -> bytes` moves the
    128-byte id from rank 0 to every rank over any host channel (e.g. torch.distributed.broadcast_object_list)."""

    def __init__(self, device, nranks, rank, exchange):
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _check(load().nuts_comm_unique_id(ident))
        raw = exchange(bytes(ident) if rank == 0 else None)
        ident = (C.c_uint8 * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        _check(load().nuts_comm_create(C.byref(h), device, ident, nranks, rank))
        self.h = h
        self.nranks, self.rank = nranks, rank

    def gather_begin(self, sampler, local_dev_ptr, gathered_dev_ptr, count):
        _check(load().nuts_gather_draws_begin(sampler.h, self.h, local_dev_ptr, gathered_dev_ptr, count))

    def gather_end(self):
        ms = C.c_double()
        _check(load().nuts_gather_draws_end(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            load().nuts_comm_destroy(self.h)
            self.h = None
