/* nuts_user_logp.cuh — the user-supplied DEVICE log density of libnuts_b200 (model kind NUTS_LOGP_USER).
 *
 * What it replaces: `CpuLogpFunc::logp(&mut self, position: &[f64], gradient: &mut [f64]) -> Result<f64, LogpError>`
 * (reference src/math/cpu_math.rs:885-970, called through Math::logp_array, src/math/math.rs:46-50).  On the device the
 * density cannot be a host callback: it is a header that is compiled INTO the engine (one more model variant next to the four
 * built-in targets, so the user code is inlined into the leapfrog and runs at the same speed):
 *
 *     make -C nuts_rs_b200/csrc USER_LOGP=/abs/path/my_model.cuh          (default: user_models/diag_gaussian.cuh)
 *
 * The header defines `struct NutsUserLogp` with the members below.  A chain's position is spread over the threads of a team,
 * so the density is written per ELEMENT, with up to two team-wide sums available before the element pass - the shape all the
 * built-in targets have (rank-1 Gaussian: sum (x - mu); funnel: x_0 and sum x_i^2).  Arithmetic rules of the library apply:
 * the engine is compiled with -fmad=false, write fma() where a fused multiply-add is wanted.
 *
 *   static constexpr int NUM_SUMS                       0, 1 or 2 team-wide sums needed by element()
 *   static void   sums(i, dim, x_i, params, acc[2])     add element i's contribution to acc[0 .. NUM_SUMS-1]
 *   static double element(i, dim, x_i, sums[2], params, &grad_i, &status)
 *                                                       element i's term of logp (the terms are summed over i); writes
 *                                                       d logp / d x_i.  Errors: set status to NUTS_USER_RECOVERABLE (the
 *                                                       leapfrog is reported as a divergence: LogpError::is_recoverable() == true,
 *                                                       src/math/math.rs:9-13, transformed_hamiltonian.rs:562-565) or NUTS_USER_FATAL
 *                                                       (the chain stops: NutsError::LogpFailure, src/nuts.rs:231).  A non-finite
 *                                                       logp or gradient is a divergence as well (energy test, :590-597).
 *   static double finish(dim, sums[2], params)          terms of logp that belong to no element (added once)
 *
 * `params` is the device copy of nuts_logp_desc_t::user_params (n_user_params doubles, layout up to the model). */
#ifndef NUTS_USER_LOGP_CUH
#define NUTS_USER_LOGP_CUH
#define NUTS_USER_OK 0
#define NUTS_USER_RECOVERABLE 1
#define NUTS_USER_FATAL 2
#endif
