/* nuts_b200.h — C ABI of libnuts_b200.so: the B200 (sm_100a) many-chain NUTS hot path.
 *
 * This is the drop-in boundary for ONE path of pymc-devs/nuts-rs: leapfrog + logp/grad inside the
 * NUTS tree-doubling loop, plus the per-draw diagonal mass-matrix / dual-averaging adaptation.
 * The reference keeps that path behind `pub trait Math` (reference src/math/math.rs:15-314), one
 * instance per chain, one `faer::Col<f64>` per vector.  Here every object is BATCHED over chains:
 * a "plane" is a device-resident [nchains x dim] f64 matrix (row = chain), i.e. `nchains` Math::Vector's.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in any signature (cudaStream_t is passed as void*).
 *   - every function returns 0 on success, <0 on a fatal error (text via nuts_last_error()).
 *   - per-chain recoverable conditions are reported through int32 status arrays:
 *       0 ok, 1 divergent (energy error), 2 divergent (non-finite logp/grad, "recoverable logp error"),
 *       3 fatal / bad initial point (NutsError::BadInitGrad, reference src/nuts.rs:14-23).
 *   - all host pointers may be pageable or pinned; device pointers are never exposed except through
 *     nuts_plane_device_ptr (for callers that already own CUDA memory, e.g. torch tensors).
 *   - threading contract = the reference's `&mut self`: one ctx / sampler is driven by one host thread;
 *     different ctx's (one per GPU) may be driven concurrently.
 *   - there is NO CPU fallback: every entry point fails with NUTS_ERR_NO_DEVICE when no sm_100 GPU exists.
 */
#ifndef NUTS_B200_H
#define NUTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUTS_OK 0
#define NUTS_ERR_INVALID -1
#define NUTS_ERR_CUDA -2
#define NUTS_ERR_NO_DEVICE -3
#define NUTS_ERR_UNSUPPORTED -4

#define NUTS_STATUS_OK 0
#define NUTS_STATUS_DIVERGENT_ENERGY 1
#define NUTS_STATUS_DIVERGENT_LOGP 2
#define NUTS_STATUS_FATAL 3

typedef struct nuts_ctx nuts_ctx_t;
typedef struct nuts_plane nuts_plane_t;     /* [nchains x dim] f64 on the device */
typedef struct nuts_point nuts_point_t;     /* batched TransformedPoint (reference src/dynamics/transformed_hamiltonian.rs:56-77) */
typedef struct nuts_sampler nuts_sampler_t; /* batched NutsChain (reference src/chain.rs:44-61) */

/* ---- device-side log densities: stand-ins for CpuLogpFunc::logp (reference src/math/cpu_math.rs:885-970) ---- */
enum {
  NUTS_LOGP_GAUSS_ISO = 0,   /* logp = -sum (x-mu)^2/2          reference src/math/test_logps.rs:49-58, benches/sample.rs:49-62 */
  NUTS_LOGP_GAUSS_DIAG = 1,  /* logp = -sum (x-mu_i)^2/(2 s_i^2) diagonal generalisation of the above                           */
  NUTS_LOGP_GAUSS_RANK1 = 2, /* Sigma = I + s*11^T              reference tests/sample_normal.rs:29-96                           */
  NUTS_LOGP_FUNNEL = 3,      /* Neal's funnel: x0=v~N(0,fs^2), x_i~N(0,e^v)  (BASELINE.json config 3; not in the reference)      */
  NUTS_LOGP_USER = 4         /* the density compiled in from a user header: include/nuts_user_logp.cuh (CpuLogpFunc::logp,
                                reference src/math/cpu_math.rs:885-970, incl. the recoverable / fatal LogpError channel)          */
};

typedef struct {
  int32_t kind;
  int32_t _pad;
  double mu_scalar;     /* mean used for every coordinate when `mu` is NULL */
  const double* mu;     /* optional host pointer [dim] */
  const double* sigma;  /* GAUSS_DIAG: host pointer [dim] of standard deviations (required) */
  double rank1_scale;   /* GAUSS_RANK1: s */
  double funnel_scale;  /* FUNNEL: prior sd of v (3.0 in BASELINE) */
  const double* user_params; /* USER: host pointer [n_user_params], copied to the device and handed to NutsUserLogp */
  uint64_t n_user_params;
} nuts_logp_desc_t;

/* ---- settings: NutsSettings<EuclideanAdaptOptions<DiagAdaptExpSettings>> field for field --------------------
 * reference src/sampler.rs:199-239 (NutsSettings), src/adapt_strategy.rs:41-69 (EuclideanAdaptOptions),
 * src/stepsize/adapt.rs:308-329 (StepSizeSettings), :21-49 (StepSizeAdaptOptions / Method),
 * src/stepsize/dual_avg.rs:11-31 (DualAverageOptions), src/stepsize/adam.rs:13-34 (AdamOptions), src/transform/adapt/diagonal.rs:93-106 (DiagAdaptExpSettings). */
enum { NUTS_STEPSIZE_DUAL_AVERAGE = 0, NUTS_STEPSIZE_ADAM = 1, NUTS_STEPSIZE_FIXED = 2 };
/* KineticEnergyKind (dynamics/transformed_hamiltonian.rs:22-50).  Tier 1 / Tier 2 serve all three (nuts_std_norm_*, nuts_esh_momentum_update,
 * nuts_leapfrog_kinetic, nuts_initialize_trajectory_kinetic); the Tier-3 whole-draw engines are Euclidean only. */
enum { NUTS_KINETIC_EUCLIDEAN = 0, NUTS_KINETIC_EXACT_NORMAL = 1, NUTS_KINETIC_MICROCANONICAL = 2 };

typedef struct {
  double k, t0, gamma, max_step_size;
} nuts_dual_average_options_t;

typedef struct { /* AdamOptions (src/stepsize/adam.rs:13-34); defaults 0.9, 0.999, 1e-8, 0.05 */
  double beta1, beta2, epsilon, learning_rate;
} nuts_adam_options_t;

typedef struct {
  int32_t method;    /* NUTS_STEPSIZE_* */
  int32_t _pad;
  double fixed_step; /* StepSizeAdaptMethod::Fixed(val) */
  nuts_dual_average_options_t dual_average;
  nuts_adam_options_t adam;
} nuts_step_size_adapt_options_t;

typedef struct {
  double target_accept;
  double initial_step;
  int32_t has_jitter; /* Option<f64>: 0 = None */
  int32_t _pad;
  double jitter;
  nuts_step_size_adapt_options_t adapt_options;
} nuts_step_size_settings_t;

typedef struct {
  int32_t store_mass_matrix;
  int32_t use_grad_based_estimate;
} nuts_diag_adapt_settings_t;

typedef struct {
  nuts_step_size_settings_t step_size_settings;
  nuts_diag_adapt_settings_t mass_matrix_options;
  double early_window;
  double step_size_window;
  uint64_t mass_matrix_switch_freq;
  uint64_t early_mass_matrix_switch_freq;
  uint64_t mass_matrix_update_freq;
  double mass_matrix_window_growth;
} nuts_euclidean_adapt_options_t;

typedef struct {
  uint64_t num_tune;
  uint64_t num_draws;
  uint64_t maxdepth;
  uint64_t mindepth;
  int32_t store_gradient;
  int32_t store_unconstrained;
  int32_t store_transformed;
  int32_t store_divergences;
  double max_energy_error;
  nuts_euclidean_adapt_options_t adapt_options;
  int32_t check_turning;
  int32_t has_target_integration_time; /* Option<f64> */
  double target_integration_time;
  int32_t trajectory_kind; /* NUTS_KINETIC_EUCLIDEAN */
  int32_t _pad;
  uint64_t num_chains;
  uint64_t seed;
  uint64_t extra_doublings;
} nuts_settings_t;

/* DiagNutsSettings::default()  (reference src/sampler.rs:507-531,630-634). */
void nuts_settings_default(nuts_settings_t* s);

/* Per-draw sampler statistics, SoA, each a HOST array [n_draws x nchains] (draw-major) or NULL to skip.
 * Names follow the reference's stat schema: NutsStats (src/chain.rs:215-231), stepsize Stats
 * (src/stepsize/adapt.rs:274-281), PointStats (src/dynamics/transformed_hamiltonian.rs:96-112),
 * HamiltonianStats (:497-505), GlobalStrategyStats (src/adapt_strategy.rs:248-257), Progress (src/sampler.rs:165-174). */
typedef struct {
  uint64_t* depth;
  uint8_t* maxdepth_reached;
  int64_t* index_in_trajectory;
  double* logp;
  double* energy;
  double* energy_error;
  uint8_t* diverging;
  double* step_size;            /* step size AFTER this draw's adaptation, like HamiltonianStats.step_size */
  double* step_size_bar;
  double* mean_tree_accept;
  double* mean_tree_accept_sym;
  uint64_t* n_steps;
  double* max_energy_error;
  uint8_t* tuning;
  double* fisher_distance;
} nuts_stats_t;

const char* nuts_last_error(void);
/* 0 when a usable sm_100 device is visible, NUTS_ERR_NO_DEVICE otherwise. */
int nuts_device_available(void);

/* ===================== Tier 0: lifecycle (Math::new_array / read_from_slice / write_to_slice, math.rs:24,94-105) */
int nuts_ctx_create(nuts_ctx_t** ctx, int device_id, uint64_t nchains, uint64_t dim, const nuts_logp_desc_t* model);
int nuts_ctx_destroy(nuts_ctx_t* ctx);
int nuts_ctx_synchronize(nuts_ctx_t* ctx);
uint64_t nuts_ctx_nchains(const nuts_ctx_t* ctx);
uint64_t nuts_ctx_dim(const nuts_ctx_t* ctx); /* Math::dim, math.rs:69 */
void* nuts_ctx_stream(nuts_ctx_t* ctx);       /* the cudaStream_t every call on this ctx is ordered on */
/* device time (CUDA events on the ctx stream, ms) of the kernel of the last nuts_leapfrog call: measurement hook of bench.py */
int nuts_ctx_last_kernel_ms(nuts_ctx_t* ctx, float* ms);
int nuts_plane_alloc(nuts_ctx_t* ctx, nuts_plane_t** plane);                     /* new_array (zero filled) */
int nuts_plane_free(nuts_ctx_t* ctx, nuts_plane_t* plane);
int nuts_plane_read_from_host(nuts_ctx_t* ctx, nuts_plane_t* dst, const double* src /*[N*d]*/); /* read_from_slice */
int nuts_plane_write_to_host(nuts_ctx_t* ctx, const nuts_plane_t* src, double* dst /*[N*d]*/);  /* write_to_slice / box_array */
double* nuts_plane_device_ptr(nuts_plane_t* plane, uint64_t* row_stride_elems);

/* ===================== Tier 1: batched Math-trait ops, 1:1 names (math.rs line cited per op) ===================
 * `a` is a HOST array [N] of per-chain scalars, or NULL with `a_bcast` used for every chain.
 * `active` is an optional HOST uint8 [N] mask (NULL = all chains). Reductions write HOST arrays [N]. */
int nuts_axpy(nuts_ctx_t*, const nuts_plane_t* x, nuts_plane_t* y, const double* a, double a_bcast, const uint8_t* active);                 /* :99  y = a*x + y (FMA) */
int nuts_axpy_out(nuts_ctx_t*, const nuts_plane_t* x, const nuts_plane_t* y, const double* a, double a_bcast, nuts_plane_t* out, const uint8_t* active); /* :98 */
/* geodesic / isokinetic integrator primitives (KineticEnergyKind::ExactNormal / Microcanonical); epsilon / step_size: HOST [N] or NULL + bcast.
 * std_norm_flow (:155-161): pos_out = pos cos(eps) + vel sin(eps), vel = -pos sin(eps) + vel cos(eps) (sin / cos evaluated on the host);
 * std_norm_grad_flow(_inplace) (:162-176): vel_out = vel + eps (pos + grad); array_normalize (:178-181): v /= |v|;
 * esh_momentum_update (:183-210): momentum updated in place on the unit sphere, kinetic_energy_change HOST [N]. */
int nuts_std_norm_flow(nuts_ctx_t*, const nuts_plane_t* pos, nuts_plane_t* pos_out, nuts_plane_t* vel, const double* epsilon, double epsilon_bcast, const uint8_t* active);
int nuts_std_norm_grad_flow(nuts_ctx_t*, const nuts_plane_t* pos, const nuts_plane_t* grad, const nuts_plane_t* vel, nuts_plane_t* vel_out, const double* epsilon, double epsilon_bcast, const uint8_t* active);
int nuts_std_norm_grad_flow_inplace(nuts_ctx_t*, const nuts_plane_t* pos, const nuts_plane_t* grad, nuts_plane_t* vel, const double* epsilon, double epsilon_bcast, const uint8_t* active);
int nuts_array_normalize(nuts_ctx_t*, nuts_plane_t* v, const uint8_t* active);
int nuts_esh_momentum_update(nuts_ctx_t*, const nuts_plane_t* gradient, nuts_plane_t* momentum, const double* step_size, double step_size_bcast, const uint8_t* active, double* kinetic_energy_change);
int nuts_array_mult(nuts_ctx_t*, const nuts_plane_t* a1, const nuts_plane_t* a2, nuts_plane_t* dest);    /* :123 */
int nuts_array_mult_inplace(nuts_ctx_t*, nuts_plane_t* a1, const nuts_plane_t* a2);                      /* :124 */
int nuts_array_recip(nuts_ctx_t*, const nuts_plane_t* a, nuts_plane_t* dest);                            /* :125 */
int nuts_fill_array(nuts_ctx_t*, nuts_plane_t* a, double val);                                           /* :119 */
int nuts_copy_into(nuts_ctx_t*, const nuts_plane_t* src, nuts_plane_t* dst);                             /* :97  */
int nuts_array_vector_dot(nuts_ctx_t*, const nuts_plane_t* a1, const nuts_plane_t* a2, double* out);     /* :212 */
int nuts_scalar_prods3(nuts_ctx_t*, const nuts_plane_t* positive1, const nuts_plane_t* negative1, const nuts_plane_t* positive2,
                       const nuts_plane_t* x, const nuts_plane_t* y, double* out1, double* out2);        /* :75-82 */
int nuts_scalar_prods2(nuts_ctx_t*, const nuts_plane_t* positive1, const nuts_plane_t* positive2,
                       const nuts_plane_t* x, const nuts_plane_t* y, double* out1, double* out2);        /* :84-90 */
int nuts_sq_norm_sum(nuts_ctx_t*, const nuts_plane_t* x, const nuts_plane_t* y, double* out);            /* :92  sum (x+y)^2 */
int nuts_array_all_finite(nuts_ctx_t*, const nuts_plane_t* a, uint8_t* out);                             /* :121 */
int nuts_array_all_finite_and_nonzero(nuts_ctx_t*, const nuts_plane_t* a, uint8_t* out);                 /* :122 */
int nuts_array_sum_ln(nuts_ctx_t*, const nuts_plane_t* a, double* out);                                  /* :113-117 */
/* dest[c,i] = stds[c,i] * normal(seed, stream = chain_offset + c + 1, counter..)  (:213-218); advances nothing:
 * the caller owns the counter; consumes ceil(dim/2) counter values starting at `counter`. */
int nuts_array_gaussian(nuts_ctx_t*, nuts_plane_t* dest, const nuts_plane_t* stds, uint64_t seed, uint64_t chain_offset, uint64_t counter);
int nuts_array_update_variance(nuts_ctx_t*, nuts_plane_t* mean, nuts_plane_t* variance, const nuts_plane_t* value,
                               const double* diff_scale /*[N] host, or NULL*/, double diff_scale_bcast);  /* :227-233 */
int nuts_array_update_var_inv_std_draw(nuts_ctx_t*, nuts_plane_t* inv_std, nuts_plane_t* std, const nuts_plane_t* draw_var,
                                       double scale, int has_fill, double fill_invalid, double clamp_lo, double clamp_hi);        /* :234-242 */
int nuts_array_update_var_inv_std_draw_grad(nuts_ctx_t*, nuts_plane_t* inv_std, nuts_plane_t* std, const nuts_plane_t* draw_var,
                                            const nuts_plane_t* grad_var, int has_fill, double fill_invalid, double clamp_lo, double clamp_hi); /* :243-251 */
int nuts_array_update_var_inv_std_grad(nuts_ctx_t*, nuts_plane_t* inv_std, nuts_plane_t* std, const nuts_plane_t* gradient,
                                       double fill_invalid, double clamp_lo, double clamp_hi);                                    /* :253-260 */
/* logp_array (:46-50): gradient plane written, logp[N] + status[N] to host. */
int nuts_logp_array(nuts_ctx_t*, const nuts_plane_t* position, nuts_plane_t* gradient, double* logp, int32_t* status);

/* ===================== Tier 2: fused Hamiltonian ops (reference src/dynamics/hamiltonian.rs:145-258) ============
 * A nuts_point_t is N TransformedPoints: 5 planes + per-chain scalars.  The diagonal transformation
 * (DiagMassMatrix, reference src/transform/diagonal.rs:9-17) lives in the ctx: stds, inv_stds, mean, logdet, id. */
int nuts_point_alloc(nuts_ctx_t*, nuts_point_t** point);
int nuts_point_free(nuts_ctx_t*, nuts_point_t* point);
/* which: 0 untransformed_position, 1 untransformed_gradient, 2 transformed_position, 3 transformed_gradient, 4 velocity */
nuts_plane_t* nuts_point_plane(nuts_point_t* point, int which);
/* scalars, HOST arrays [N] (any may be NULL) */
int nuts_point_get_scalars(nuts_ctx_t*, const nuts_point_t*, int64_t* index_in_trajectory, double* logp, double* logdet,
                           double* kinetic_energy, double* initial_energy, int64_t* transform_id);
int nuts_point_set_scalars(nuts_ctx_t*, nuts_point_t*, const int64_t* index_in_trajectory, const double* logp, const double* logdet,
                           const double* kinetic_energy, const double* initial_energy, const int64_t* transform_id);
/* DiagMassMatrix::set_transform (diagonal.rs:156-162): stds/mean HOST [N*d]; bumps id, recomputes inv_stds + logdet. */
int nuts_set_transform(nuts_ctx_t*, const double* stds, const double* mean);
int nuts_get_transform(nuts_ctx_t*, double* stds, double* inv_stds, double* mean, double* logdet /*[N]*/, int64_t* id /*[N]*/);
/* ---- Low-rank mass matrix (reference src/transform/low_rank.rs, src/math/math.rs:127-144) -----------------------
 * EigVectors + EigValues of every chain (Math::new_eig_vectors / new_eig_values, math.rs:150-160): chain c has rank[c] <= rank_max
 * eigenvectors; vecs is HOST [N][rank_max][dim] (eigenvector k of chain c at vecs[(c * rank_max + k) * dim]), vals HOST [N][rank_max]
 * (entries k >= rank[c] are ignored), rank HOST [N] or NULL for rank_max everywhere.  rank_max <= 64. */
typedef struct nuts_eigs nuts_eigs_t;
int nuts_eigs_create(nuts_ctx_t*, nuts_eigs_t** eigs, uint64_t rank_max, const double* vecs, const double* vals, const int32_t* rank);
int nuts_eigs_free(nuts_ctx_t*, nuts_eigs_t* eigs);
/* Math::apply_lowrank_transform (math.rs:131-137, cpu_math.rs:332-377): dest = (I + U (diag(vals) - I) U^T) rhs per chain; a chain
 * without eigenvectors copies rhs.  _inplace (math.rs:139-144): rhs_and_dest is updated in place. */
int nuts_apply_lowrank_transform(nuts_ctx_t*, const nuts_eigs_t* eigs, const nuts_plane_t* rhs, nuts_plane_t* dest);
int nuts_apply_lowrank_transform_inplace(nuts_ctx_t*, const nuts_eigs_t* eigs, nuts_plane_t* rhs_and_dest);
/* LowRankMassMatrix::update (low_rank.rs:158-190) for every chain: the Tier-2 transformation becomes
 *   F(z) = sigma * ((I + U (sqrt(lambda) - 1) U^T) z + mean_low_rank) + mean,  logdet = sum ln(1 / sigma) - 1/2 sum ln(lambda).
 * stds, mean, mean_low_rank: HOST [N*dim]; vals / vecs / rank as in nuts_eigs_create (vals = the raw eigenvalues lambda).
 * A chain with a non-finite input keeps its old transformation (low_rank.rs:168-173) and gets accepted[c] = 0 (accepted may be
 * NULL).  nuts_set_transform (diagonal) drops the correction again, like update_from_grad (low_rank.rs:143-156).
 * nuts_init_state / nuts_initialize_trajectory / nuts_leapfrog use the transformation that is current. */
int nuts_set_lowrank_transform(nuts_ctx_t*, const double* stds, const double* mean, uint64_t rank_max, const double* vals,
                               const double* vecs, const int32_t* rank, const double* mean_low_rank, uint8_t* accepted);

/* Hamiltonian::init_state (transformed_hamiltonian.rs:640-661): x -> logp, grad, z, grad_z; status 3 when check_all fails. */
int nuts_init_state(nuts_ctx_t*, nuts_point_t* point, const double* position /*HOST [N*d]*/, int32_t* status);
/* Hamiltonian::initialize_trajectory (:687-736): resample velocity from (seed, chain_offset+c+1, counter), re-whiten when the
 * transformation id changed, kinetic energy, index=0, initial_energy. */
int nuts_initialize_trajectory(nuts_ctx_t*, nuts_point_t* point, int resample_velocity, uint64_t seed, uint64_t chain_offset, uint64_t counter);
/* Hamiltonian::leapfrog (:524-615), Euclidean: eps[c] = dir[c] * step_size[c] (* step_size_factor = 1).
 * step_size HOST [N] or NULL+bcast; dir HOST int8 [N] (+1/-1) or NULL (= +1); energy_baseline HOST [N] or NULL (= start.initial_energy).
 * Writes `out` for every active chain (also divergent ones) and status[N] (0 ok / 1 / 2); energy_error[N] optional.
 * Elementwise targets (GAUSS_ISO / GAUSS_DIAG) on the diagonal transformation run k_leapfrog_tma: the five input rows are staged through
 * shared memory by cp.async.bulk + mbarrier (0.96-0.98 of the HBM copy peak; bit-identical to the register path, which
 * NUTS_B200_PLANE_TMA=0 selects). */
int nuts_leapfrog(nuts_ctx_t*, const nuts_point_t* start, nuts_point_t* out, const double* step_size, double step_size_bcast,
                  const int8_t* dir, const double* energy_baseline, double max_energy_error, const uint8_t* active,
                  int32_t* status, double* energy_error);
/* Hamiltonian::leapfrog for the other KineticEnergyKinds (:160-258, :582-596).  kind = NUTS_KINETIC_*; EUCLIDEAN forwards to
 * nuts_leapfrog.  EXACT_NORMAL: std_norm_grad_flow half-steps around the exact rotation std_norm_flow; MICROCANONICAL: ESH momentum
 * updates with sqrt(dim)-scaled steps, point.kinetic_energy = the accumulated kinetic-energy change, divergence when
 * |energy error| >= max_energy_error.  Diagonal transformation only (NUTS_ERR_UNSUPPORTED while a low-rank correction is set). */
int nuts_leapfrog_kinetic(nuts_ctx_t*, int kind, const nuts_point_t* start, nuts_point_t* out, const double* step_size,
                          double step_size_bcast, const int8_t* dir, const double* energy_baseline, double max_energy_error,
                          const uint8_t* active, int32_t* status, double* energy_error);
/* Hamiltonian::initialize_trajectory with the kind (:699-702 unit-sphere momentum, :720-729 kinetic_energy = 0 for MICROCANONICAL) */
int nuts_initialize_trajectory_kinetic(nuts_ctx_t*, int kind, nuts_point_t* point, int resample_velocity, uint64_t seed,
                                       uint64_t chain_offset, uint64_t counter);
/* Hamiltonian::is_turning (:617-638) */
int nuts_is_turning(nuts_ctx_t*, const nuts_point_t* state1, const nuts_point_t* state2, uint8_t* turning);

/* ===================== Tier 3: whole draws (Chain::set_position / Chain::draw, reference src/chain.rs:137-188) == */
int nuts_sampler_create(nuts_ctx_t*, nuts_sampler_t** sampler, const nuts_settings_t* settings, uint64_t seed, uint64_t chain_id_offset);
int nuts_sampler_destroy(nuts_sampler_t* sampler);
/* Chain::set_position for every chain: position HOST [N*d]; status[N] (0 ok, 3 bad initial point). */
int nuts_set_position(nuts_sampler_t*, const double* position, int32_t* status);
/* The same for the chains with mask[c] != 0 only (mask, status: HOST [N]; status entries of the other chains are left as they
 * are): the retry of bad initial points - the reference draws a fresh init_position up to 500 times for a chain whose
 * set_position fails (src/sampler.rs:1133-1143).  The chain starts over as a new NutsChain; its random stream continues. */
int nuts_set_position_masked(nuts_sampler_t*, const double* position, const uint8_t* mask, int32_t* status);
/* n_draws x Chain::draw for every chain.  draws_out: HOST [n_draws x N x d] (may be NULL); stats: HOST SoA (may be NULL).
 * Returns after the stream is synchronised. */
int nuts_draw(nuts_sampler_t*, uint64_t n_draws, double* draws_out, const nuts_stats_t* stats);
/* draws_out placement decides how the draws leave the GPU: a page-locked host buffer (nuts_host_alloc, cudaHostAlloc,
 * cudaHostRegister, torch pin_memory) or a device buffer is written DIRECTLY by the draw kernel (posted PCIe writes that
 * overlap the sampling; no staging copy); pageable host memory is staged through a device buffer and copied afterwards. */
int nuts_host_alloc(void** ptr, uint64_t bytes);   /* page-locked, device-mapped host memory (cudaHostAlloc) */
int nuts_host_free(void* ptr);
/* 1 when the last nuts_draw wrote the draws straight into the caller's buffer, 0 when it staged them. */
int nuts_sampler_last_draw_direct(nuts_sampler_t*, int32_t* direct);
/* Same, but draws stay on the device: draws_dev is a DEVICE pointer [n_draws x N x d] (may be NULL). Asynchronous on the ctx stream. */
int nuts_draw_device(nuts_sampler_t*, uint64_t n_draws, double* draws_dev);
/* total leapfrog steps (sum over chains, incl. init searches and divergent steps) and draws done so far. */
int nuts_sampler_counters(nuts_sampler_t*, uint64_t* total_leapfrogs, uint64_t* draws_done);
/* milliseconds spent inside the draw kernel during the last nuts_draw / nuts_draw_device call (CUDA events on the ctx stream),
 * and the number of kernel launches it made. */
int nuts_sampler_last_timing(nuts_sampler_t*, double* kernel_ms, uint64_t* launches);
/* white-box access used by the parity tests: HOST arrays, any may be NULL.
 * position [N*d], step_size [N], stds [N*d], mean [N*d], rng_counter [N]. */
int nuts_sampler_get_state(nuts_sampler_t*, double* position, double* step_size, double* stds, double* mean, uint64_t* rng_counter);
int nuts_sampler_set_step_size(nuts_sampler_t*, const double* step_size /*[N]*/);

/* ---- complete per-chain state between two draws: checkpoint / resume, and teacher-forced parity tests --------------
 * Everything a NutsChain (reference src/chain.rs:44-61) carries from one Chain::draw to the next, SoA over chains, HOST
 * arrays; vectors are dense [N*d], scalars [N].  In nuts_sampler_get_chain_state any member may be NULL (skipped); in
 * nuts_sampler_set_chain_state every member must be set.  A sampler restored with set_chain_state continues bit-identically
 * to the sampler the state was read from (same settings, seed, chain_id_offset and model).
 *   point        = NutsChain.state: TransformedPoint x, grad_x, z, grad_z, logp, logdet, transform_id (transformed_hamiltonian.rs:56-77)
 *   mass matrix  = DiagMassMatrix stds, inv_stds, mean, logdet, id (src/transform/diagonal.rs:9-17)
 *   step size    = Hamiltonian.step_size + DualAverage log_step, log_step_adapted, hbar, mu, count (src/stepsize/dual_avg.rs:33-41)
 *   estimators   = DiagAdaptStrategy's four RunningVariance: foreground / background x draw / grad, each mean + variance
 *                  accumulator + count (src/transform/adapt/diagonal.rs:17-55,108-118)
 *   schedule     = GlobalStrategy tuning, has_initial_mass_matrix, last_update, current_window_size (src/adapt_strategy.rs:27-39)
 *   chain        = draw_count (chain.rs:56), position of the chain's random stream, alive (0 after BadInitGrad), leapfrog total */
typedef struct {
  double *position, *gradient, *transformed_position, *transformed_gradient; /* [N*d] */
  double *logp, *point_logdet;                                               /* [N]   */
  int64_t* point_transform_id;
  double *stds, *inv_stds, *mean; /* [N*d] */
  double* mass_matrix_logdet;     /* [N]   */
  int64_t* mass_matrix_id;
  double* step_size;
  double *da_log_step, *da_log_step_adapted, *da_hbar, *da_mu;
  uint64_t* da_count;
  double *draw_mean, *draw_var, *grad_mean, *grad_var;             /* foreground estimators [N*d] */
  double *draw_mean_bg, *draw_var_bg, *grad_mean_bg, *grad_var_bg; /* background estimators [N*d] */
  uint64_t *foreground_count, *background_count;
  uint8_t *tuning, *has_initial_mass_matrix;
  uint64_t *last_update, *current_window_size;
  uint64_t *draw_count, *rng_counter, *total_leapfrogs;
  uint8_t* alive;
} nuts_chain_state_t;
int nuts_sampler_get_chain_state(nuts_sampler_t*, const nuts_chain_state_t* out);
int nuts_sampler_set_chain_state(nuts_sampler_t*, const nuts_chain_state_t* in);

/* ---- Low-rank mass matrix in whole draws (reference src/transform/low_rank.rs + src/transform/adapt/low_rank.rs).
 * nuts_sampler_create_lowrank builds the sampler on an engine with the low-rank transformation compiled in (dim <= 1024): every
 * leapfrog applies x = sigma * ((I + U (sqrt(lambda) - 1) U^T) z + mu_lr) + mean and grad_z = (I + U (sqrt(lambda) - 1) U^T)(sigma * grad_x),
 * i.e. two skinny products with the chain's U [rank x dim] and two team-wide reductions of `rank` values on top of the diagonal path.
 * The estimator (thin SVDs, pivoted QR and three eigendecompositions per update and chain, adapt/low_rank.rs:73-262) stays on the
 * host: the caller collects draws (nuts_draw) and gradients (nuts_sampler_set_grads_out), computes (stds, mean, vals, vecs,
 * mean_low_rank) and installs them with nuts_sampler_set_lowrank_transform == LowRankMassMatrix::update (low_rank.rs:158-190) for
 * every chain (arguments as nuts_set_lowrank_transform).  The first update of a run re-runs the step size search from the current
 * point (adapt_strategy.rs:204-214).  The device-side adaptation of such a sampler should be limited to the step size: set
 * mass_matrix_update_freq / early_mass_matrix_switch_freq / mass_matrix_switch_freq beyond num_tune (nuts_rs_b200/lowrank.py does).
 * Checkpoints: nuts_sampler_get/set_chain_state carry the diagonal part and all scalars; the caller keeps (vals, vecs, mean_low_rank) of
 * its last update and re-installs them BEFORE nuts_sampler_set_chain_state when it restores a run. */
int nuts_sampler_create_lowrank(nuts_ctx_t*, nuts_sampler_t** sampler, const nuts_settings_t* settings, uint64_t seed, uint64_t chain_id_offset,
                                uint64_t rank_max);
int nuts_sampler_set_lowrank_transform(nuts_sampler_t*, const double* stds, const double* mean, uint64_t rank_max, const double* vals,
                                       const double* vecs, const int32_t* rank, const double* mean_low_rank, uint8_t* accepted);
/* gradient of logp at every draw of the following nuts_draw / nuts_draw_device calls: [n_draws][nchains][dim], device memory or
 * page-locked host memory (nuts_host_alloc); NULL switches it off again */
int nuts_sampler_set_grads_out(nuts_sampler_t*, double* grads);

/* ===================== Multi-GPU: gather of the draws (the only exchange of the path) ====================
 * Chains are independent (reference src/sampler.rs:1094-1126: one Math / RNG stream / adaptation per chain), so a job is split
 * into contiguous blocks of chains, one sampler per GPU, with chain_id_offset = first global chain id of the block; nothing is
 * exchanged while sampling.  What the reference does with its shared trace (src/sampler.rs:1065) is here an NCCL all-gather of
 * the per-rank draw buffers over NVLink, issued on the communicator's OWN stream so that it overlaps the next batch of draws:
 *     nuts_draw_device(s, n, buf[k & 1]);                       // batch k on the sampler's stream
 *     nuts_gather_draws_begin(s, comm, buf[k & 1], all[k & 1], count);   // waits (event) for batch k only, then runs beside batch k+1
 *     ... next batch ...      nuts_gather_draws_end(comm, &ms);  // before buf[k & 1] / all[k & 1] are reused
 * libnccl is loaded at run time (dlopen "libnccl.so.2": the copy a host process such as PyTorch already carries is used);
 * without it the calls fail with NUTS_ERR_UNSUPPORTED.  One process per GPU; the 128-byte id travels by any host channel. */
typedef struct nuts_comm nuts_comm_t;
int nuts_comm_unique_id(uint8_t id[128]); /* rank 0 */
int nuts_comm_create(nuts_comm_t** comm, int device_id, const uint8_t id[128], int nranks, int rank);
int nuts_comm_destroy(nuts_comm_t* comm);
/* all-gather `count` doubles per rank: gathered_dev = [nranks][count] (rank-major) on every rank; both DEVICE pointers. */
int nuts_gather_draws_begin(nuts_sampler_t*, nuts_comm_t* comm, const double* local_dev, double* gathered_dev, uint64_t count);
/* wait for the gather in flight; elapsed_ms (optional) = its device time on the communicator's stream. */
int nuts_gather_draws_end(nuts_comm_t* comm, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* NUTS_B200_H */
