"""ctypes mirror of include/nuts_b200.h (POD structs + constants only; no library loading here).

Field names follow the reference's settings structs one for one:
NutsSettings (reference src/sampler.rs:199-239), EuclideanAdaptOptions (src/adapt_strategy.rs:41-69),
StepSizeSettings / StepSizeAdaptOptions (src/stepsize/adapt.rs:21-49,308-329), DualAverageOptions
(src/stepsize/dual_avg.rs:11-31), DiagAdaptExpSettings (src/transform/adapt/diagonal.rs:93-106).
"""
import ctypes as C

NUTS_LOGP_GAUSS_ISO = 0
NUTS_LOGP_GAUSS_DIAG = 1
NUTS_LOGP_GAUSS_RANK1 = 2
NUTS_LOGP_FUNNEL = 3
NUTS_LOGP_USER = 4

NUTS_STEPSIZE_DUAL_AVERAGE = 0
NUTS_STEPSIZE_ADAM = 1
NUTS_STEPSIZE_FIXED = 2
NUTS_KINETIC_EUCLIDEAN = 0
NUTS_KINETIC_EXACT_NORMAL = 1
NUTS_KINETIC_MICROCANONICAL = 2

NUTS_STATUS_OK = 0
NUTS_STATUS_DIVERGENT_ENERGY = 1
NUTS_STATUS_DIVERGENT_LOGP = 2
NUTS_STATUS_FATAL = 3

NUTS_ERR_NO_DEVICE = -3

c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_u64_p = C.POINTER(C.c_uint64)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)
c_i8_p = C.POINTER(C.c_int8)


class LogpDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("_pad", C.c_int32),
        ("mu_scalar", C.c_double),
        ("mu", c_double_p),
        ("sigma", c_double_p),
        ("rank1_scale", C.c_double),
        ("funnel_scale", C.c_double),
        ("user_params", c_double_p),
        ("n_user_params", C.c_uint64),
    ]


class DualAverageOptions(C.Structure):
    _fields_ = [("k", C.c_double), ("t0", C.c_double), ("gamma", C.c_double), ("max_step_size", C.c_double)]


class AdamOptions(C.Structure):
    _fields_ = [("beta1", C.c_double), ("beta2", C.c_double), ("epsilon", C.c_double), ("learning_rate", C.c_double)]


class StepSizeAdaptOptions(C.Structure):
    _fields_ = [
        ("method", C.c_int32),
        ("_pad", C.c_int32),
        ("fixed_step", C.c_double),
        ("dual_average", DualAverageOptions),
        ("adam", AdamOptions),
    ]


class StepSizeSettings(C.Structure):
    _fields_ = [
        ("target_accept", C.c_double),
        ("initial_step", C.c_double),
        ("has_jitter", C.c_int32),
        ("_pad", C.c_int32),
        ("jitter", C.c_double),
        ("adapt_options", StepSizeAdaptOptions),
    ]


class DiagAdaptExpSettings(C.Structure):
    _fields_ = [("store_mass_matrix", C.c_int32), ("use_grad_based_estimate", C.c_int32)]


class EuclideanAdaptOptions(C.Structure):
    _fields_ = [
        ("step_size_settings", StepSizeSettings),
        ("mass_matrix_options", DiagAdaptExpSettings),
        ("early_window", C.c_double),
        ("step_size_window", C.c_double),
        ("mass_matrix_switch_freq", C.c_uint64),
        ("early_mass_matrix_switch_freq", C.c_uint64),
        ("mass_matrix_update_freq", C.c_uint64),
        ("mass_matrix_window_growth", C.c_double),
    ]


class NutsSettings(C.Structure):
    _fields_ = [
        ("num_tune", C.c_uint64),
        ("num_draws", C.c_uint64),
        ("maxdepth", C.c_uint64),
        ("mindepth", C.c_uint64),
        ("store_gradient", C.c_int32),
        ("store_unconstrained", C.c_int32),
        ("store_transformed", C.c_int32),
        ("store_divergences", C.c_int32),
        ("max_energy_error", C.c_double),
        ("adapt_options", EuclideanAdaptOptions),
        ("check_turning", C.c_int32),
        ("has_target_integration_time", C.c_int32),
        ("target_integration_time", C.c_double),
        ("trajectory_kind", C.c_int32),
        ("_pad", C.c_int32),
        ("num_chains", C.c_uint64),
        ("seed", C.c_uint64),
        ("extra_doublings", C.c_uint64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("depth", c_u64_p),
        ("maxdepth_reached", c_u8_p),
        ("index_in_trajectory", c_i64_p),
        ("logp", c_double_p),
        ("energy", c_double_p),
        ("energy_error", c_double_p),
        ("diverging", c_u8_p),
        ("step_size", c_double_p),
        ("step_size_bar", c_double_p),
        ("mean_tree_accept", c_double_p),
        ("mean_tree_accept_sym", c_double_p),
        ("n_steps", c_u64_p),
        ("max_energy_error", c_double_p),
        ("tuning", c_u8_p),
        ("fisher_distance", c_double_p),
    ]


# nuts_chain_state_t: (field, numpy dtype, is a [N*d] vector) in struct order
CHAIN_STATE_FIELDS = (
    [(n, "float64", True) for n in ("position", "gradient", "transformed_position", "transformed_gradient")]
    + [("logp", "float64", False), ("point_logdet", "float64", False), ("point_transform_id", "int64", False)]
    + [(n, "float64", True) for n in ("stds", "inv_stds", "mean")]
    + [("mass_matrix_logdet", "float64", False), ("mass_matrix_id", "int64", False), ("step_size", "float64", False)]
    + [(n, "float64", False) for n in ("da_log_step", "da_log_step_adapted", "da_hbar", "da_mu")]
    + [("da_count", "uint64", False)]
    + [(n, "float64", True) for n in ("draw_mean", "draw_var", "grad_mean", "grad_var", "draw_mean_bg", "draw_var_bg", "grad_mean_bg",
                                      "grad_var_bg")]
    + [("foreground_count", "uint64", False), ("background_count", "uint64", False), ("tuning", "uint8", False),
       ("has_initial_mass_matrix", "uint8", False), ("last_update", "uint64", False), ("current_window_size", "uint64", False),
       ("draw_count", "uint64", False), ("rng_counter", "uint64", False), ("total_leapfrogs", "uint64", False), ("alive", "uint8", False)]
)
_CT = {"float64": c_double_p, "int64": c_i64_p, "uint64": c_u64_p, "uint8": c_u8_p}


class ChainState(C.Structure):
    _fields_ = [(n, _CT[dt]) for n, dt, _ in CHAIN_STATE_FIELDS]


def alloc_chain_state(nchains, dim, arrays=None):
    """(ChainState struct, dict of numpy arrays).  `arrays`: an existing dict (e.g. read from another sampler) to wrap."""
    import numpy as np

    st = ChainState()
    out = {}
    for name, dt, vec in CHAIN_STATE_FIELDS:
        shape = (nchains, dim) if vec else (nchains,)
        a = np.zeros(shape, dtype=dt) if arrays is None else np.ascontiguousarray(arrays[name], dtype=dt).reshape(shape)
        out[name] = a
        setattr(st, name, a.ctypes.data_as(_CT[dt]))
    return st, out


# numpy dtype per stat, in struct order
STAT_DTYPES = {
    "depth": "uint64",
    "maxdepth_reached": "uint8",
    "index_in_trajectory": "int64",
    "logp": "float64",
    "energy": "float64",
    "energy_error": "float64",
    "diverging": "uint8",
    "step_size": "float64",
    "step_size_bar": "float64",
    "mean_tree_accept": "float64",
    "mean_tree_accept_sym": "float64",
    "n_steps": "uint64",
    "max_energy_error": "float64",
    "tuning": "uint8",
    "fisher_distance": "float64",
}


def default_settings() -> NutsSettings:
    """DiagNutsSettings::default() (reference src/sampler.rs:507-531,630-634) built on the Python side."""
    s = NutsSettings()
    s.num_tune = 400
    s.num_draws = 1000
    s.maxdepth = 10
    s.mindepth = 0
    s.max_energy_error = 1000.0
    s.check_turning = 1
    s.num_chains = 6
    s.seed = 0
    s.extra_doublings = 0
    a = s.adapt_options
    a.early_window = 0.3
    a.step_size_window = 0.15
    a.mass_matrix_switch_freq = 80
    a.early_mass_matrix_switch_freq = 10
    a.mass_matrix_update_freq = 1
    a.mass_matrix_window_growth = 1.5
    a.mass_matrix_options.store_mass_matrix = 0
    a.mass_matrix_options.use_grad_based_estimate = 1
    ss = a.step_size_settings
    ss.target_accept = 0.8
    ss.initial_step = 0.1
    ss.has_jitter = 1
    ss.jitter = 0.1
    ss.adapt_options.method = NUTS_STEPSIZE_DUAL_AVERAGE
    ss.adapt_options.fixed_step = 0.0
    ss.adapt_options.dual_average.k = 0.75
    ss.adapt_options.dual_average.t0 = 10.0
    ss.adapt_options.dual_average.gamma = 0.05
    ss.adapt_options.dual_average.max_step_size = 3.141592653589793
    ss.adapt_options.adam.beta1, ss.adapt_options.adam.beta2 = 0.9, 0.999  # AdamOptions::default (src/stepsize/adam.rs:25-34)
    ss.adapt_options.adam.epsilon, ss.adapt_options.adam.learning_rate = 1e-8, 0.05
    return s


def make_logp_desc(kind, dim, mu=None, sigma=None, rank1_scale=0.0, funnel_scale=3.0, user_params=None):
    """Build a LogpDesc; returns (desc, keepalive) — keep `keepalive` referenced while the desc is in use."""
    import numpy as np

    d = LogpDesc()
    d.kind = int(kind)
    keep = []
    if mu is None:
        d.mu_scalar = 0.0
    elif np.isscalar(mu):
        d.mu_scalar = float(mu)
    else:
        m = np.ascontiguousarray(mu, dtype=np.float64)
        assert m.shape == (dim,)
        keep.append(m)
        d.mu = m.ctypes.data_as(c_double_p)
    if sigma is not None:
        sg = np.ascontiguousarray(sigma, dtype=np.float64)
        assert sg.shape == (dim,)
        keep.append(sg)
        d.sigma = sg.ctypes.data_as(c_double_p)
    d.rank1_scale = float(rank1_scale)
    d.funnel_scale = float(funnel_scale)
    if user_params is not None:
        up = np.ascontiguousarray(user_params, dtype=np.float64).ravel()
        keep.append(up)
        d.user_params = up.ctypes.data_as(c_double_p)
        d.n_user_params = up.size
    return d, keep
