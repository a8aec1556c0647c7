// Example / default user density: the diagonal Gaussian of BASELINE config 2 written against the user interface
// (include/nuts_user_logp.cuh).  params = [mu_0 .. mu_{d-1} | prec_0 .. prec_{d-1}], prec_i = 1 / sigma_i^2.
// Same arithmetic as the built-in NUTS_LOGP_GAUSS_DIAG (and the reference's test target, src/math/test_logps.rs:49-58 with a
// per-coordinate precision), so tests/test_gpu_user_logp.py can require identical draws.  A coordinate beyond
// +-user_limit (params[2d], optional: n_user_params == 2d + 1) raises a recoverable error, x_0 beyond 10 x that a fatal one -
// the two error kinds of LogpError (reference src/math/math.rs:9-13).
#pragma once
#include "../../../include/nuts_user_logp.cuh"

struct NutsUserLogp {
  static constexpr int NUM_SUMS = 0;
  __device__ static void sums(int, int, double, const double*, double (&)[2]) {}
  __device__ static double element(int i, int dim, double x, const double (&)[2], const double* params, double& grad, int& status) {
    const double diff = x - params[i];
    const double pd = diff * params[dim + i];
    grad = -pd;
    const double limit = params[2 * dim];  // 0 = no limit
    if (limit > 0.0) {
      if (i == 0 && fabs(x) > 10.0 * limit) status = NUTS_USER_FATAL;
      else if (fabs(x) > limit) status = NUTS_USER_RECOVERABLE;
    }
    return -(diff * pd / 2.);
  }
  __device__ static double finish(int, const double (&)[2], const double*) { return 0.0; }
};
