// TEST INFRASTRUCTURE — CPU restatement ("oracle") of the nuts-rs diag-NUTS hot path.
//
// This file restates, function by function, the reference algorithm for the path SURVEY.md §8 scopes
// (pymc-devs/nuts-rs @ 5332136, v0.18.3; every function cites the reference file:line it follows).  It is
// the checker for the CUDA path and the CPU baseline of bench.py.  It is NOT part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or call it.
//
// Parity pinning: the reference (Rust) cannot be compiled in this image (no cargo/rustc), so the oracle is
// pinned against the reference's own known-answer tests instead: the 32-ULP kernel properties and logaddexp
// identities (src/math/util.rs:752-968 + proptest-regressions/), the diagonal-transform KATs
// (src/transform/mod.rs:175-377) and the behavioural sampler tests (src/adapt_strategy.rs:367-435,
// src/nuts.rs:399-419, src/sampler.rs:1662-1692, tests/sample_normal.rs:205-226) — see tests/test_oracle_*.py.
// The random streams are "parity unpinned" by necessity (see rng_spec.hpp).
//
// Build with -ffp-contract=off: every fused multiply-add the reference performs is written as std::fma,
// every plain product/sum stays unfused, exactly as in src/math/util.rs.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <vector>

#include "rng_spec.hpp"

namespace oracle {

using Vec = std::vector<double>;

// ------------------------------------------------------------------------------------------------
// src/math/util.rs — SIMD kernels.  pulp dispatches on the host ISA; we restate the x86-64-v3 shape:
// LANES=4 f64 per SIMD register, 4 independent accumulators, then a scalar tail (util.rs:367-396).
// ------------------------------------------------------------------------------------------------
constexpr int LANES = 4;

// util.rs:6-19
inline double logaddexp(double a, double b) {
  if (a == b) return a + std::log(2.0);
  double diff = a - b;
  if (diff > 0.) return a + std::log1p(std::exp(-diff));
  if (diff < 0.) return b + std::log1p(std::exp(diff));
  return diff;  // NaN
}

// util.rs:21-70
inline void multiply(const double* x, const double* y, double* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = x[i] * y[i];
}
// util.rs:72-112
inline void multiply_inplace(double* out, const double* x, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = x[i] * out[i];
}
// util.rs:402-446   y = a*x + y, FMA in the SIMD body and `a.mul_add(x, y)` in the tail
inline void axpy(const double* x, double* y, double a, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = std::fma(a, x[i], y[i]);
}
// util.rs:448-505
inline void axpy_out(const double* x, const double* y, double a, double* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = std::fma(a, x[i], y[i]);
}

// util.rs:507-590  harmonic-oscillator flow of the standard normal: pos_out = p cos(eps) + v sin(eps), vel = -p sin(eps) + v cos(eps).
// SIMD body (whole registers of LANES): mul_add(p, c, v * s) / mul_add(p, -s, v * c); scalar tail: unfused products.
inline void std_norm_flow(const double* pos, double* pos_out, double* vel, double epsilon, size_t n) {
  const double eps_sin = std::sin(epsilon), eps_cos = std::cos(epsilon);
  const size_t body = (n / LANES) * LANES;
  for (size_t i = 0; i < body; ++i) {
    double p = pos[i], v = vel[i];
    pos_out[i] = std::fma(p, eps_cos, v * eps_sin);
    vel[i] = std::fma(p, -eps_sin, v * eps_cos);
  }
  for (size_t i = body; i < n; ++i) {
    double p = pos[i], v = vel[i];
    double new_po = p * eps_cos + v * eps_sin;
    vel[i] = p * (-eps_sin) + v * eps_cos;
    pos_out[i] = new_po;
  }
}
// util.rs:592-671  vel_out = vel + eps * (pos + grad): mul_add(eps, p + g, v) in the SIMD body, `v + epsilon * (p + g)` (two roundings)
// in the scalar tail
inline void std_norm_grad_flow(const double* pos, const double* grad, const double* vel, double* vel_out, double epsilon, size_t n) {
  const size_t body = (n / LANES) * LANES;
  for (size_t i = 0; i < body; ++i) vel_out[i] = std::fma(epsilon, pos[i] + grad[i], vel[i]);
  for (size_t i = body; i < n; ++i) vel_out[i] = vel[i] + epsilon * (pos[i] + grad[i]);
}
// util.rs:673-742
inline void std_norm_grad_flow_inplace(const double* pos, const double* grad, double* vel, double epsilon, size_t n) {
  std_norm_grad_flow(pos, grad, vel, vel, epsilon, n);
}
// cpu_math.rs:496-503  v := v / ||v||  (sequential sum of squares, one reciprocal, one product per element)
inline void array_normalize(double* v, size_t n) {
  double ss = 0.;
  for (size_t i = 0; i < n; ++i) ss += v[i] * v[i];
  const double inv = 1.0 / std::sqrt(ss);
  for (size_t i = 0; i < n; ++i) v[i] *= inv;
}
// cpu_math.rs:505-551  ESH (isokinetic) momentum update of Microcanonical HMC; returns the kinetic-energy change
inline double esh_momentum_update(const double* gradient, double* momentum, double step_size, size_t n) {
  double gg = 0.;
  for (size_t i = 0; i < n; ++i) gg += gradient[i] * gradient[i];
  const double grad_norm = std::sqrt(gg);
  const double inv_grad_norm = 1.0 / grad_norm;
  double momentum_proj = 0.;
  for (size_t i = 0; i < n; ++i) momentum_proj += momentum[i] * gradient[i] * inv_grad_norm;
  const double dims_m1 = (double)(n - 1);
  const double delta = step_size * grad_norm / dims_m1;
  const double zeta = std::exp(-delta);
  const double coeff_g = (1.0 - zeta) * (1.0 + zeta + momentum_proj * (1.0 - zeta));
  const double coeff_p = 2.0 * zeta;
  for (size_t i = 0; i < n; ++i) momentum[i] = coeff_g * (gradient[i] * inv_grad_norm) + coeff_p * momentum[i];
  array_normalize(momentum, n);
  const double arg = momentum_proj + (1.0 - momentum_proj) * zeta * zeta;
  return (delta - 0.6931471805599453094 + std::log1p(arg)) * dims_m1;
}

// util.rs:349-400
inline double vector_dot(const double* a, const double* b, size_t n) {
  size_t nsimd = n / LANES;  // whole SIMD registers
  size_t ngroups = nsimd / 4;
  double acc[4][LANES] = {};
  size_t pos = 0;
  for (size_t g = 0; g < ngroups; ++g)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < LANES; ++l, ++pos) acc[k][l] = std::fma(a[pos], b[pos], acc[k][l]);
  for (size_t s = ngroups * 4; s < nsimd; ++s)
    for (int l = 0; l < LANES; ++l, ++pos) acc[0][l] = std::fma(a[pos], b[pos], acc[0][l]);
  double red[LANES];
  for (int l = 0; l < LANES; ++l) red[l] = (acc[0][l] + acc[1][l]) + (acc[2][l] + acc[3][l]);
  double result = (red[0] + red[1]) + (red[2] + red[3]);
  for (; pos < n; ++pos) result += a[pos] * b[pos];
  return result;
}

// util.rs:221-347  ((p1 + p2) - n1) . x  and  . y ; tail uses p1 - n1 + p2 and unfused multiply-add
inline void scalar_prods3(const double* p1, const double* n1, const double* p2, const double* x, const double* y, size_t n,
                          double* out1, double* out2) {
  size_t nsimd = n / LANES, ngroups = nsimd / 4;
  double s1[4][LANES] = {}, s2[4][LANES] = {};
  size_t pos = 0;
  for (size_t g = 0; g < ngroups; ++g)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < LANES; ++l, ++pos) {
        double sum = (p1[pos] + p2[pos]) - n1[pos];
        s1[k][l] = std::fma(sum, x[pos], s1[k][l]);
        s2[k][l] = std::fma(sum, y[pos], s2[k][l]);
      }
  for (size_t s = ngroups * 4; s < nsimd; ++s)
    for (int l = 0; l < LANES; ++l, ++pos) {
      double sum = (p1[pos] + p2[pos]) - n1[pos];
      s1[0][l] = std::fma(sum, x[pos], s1[0][l]);
      s2[0][l] = std::fma(sum, y[pos], s2[0][l]);
    }
  double r1[LANES], r2[LANES];
  for (int l = 0; l < LANES; ++l) {
    r1[l] = (s1[0][l] + s1[1][l]) + (s1[2][l] + s1[3][l]);
    r2[l] = (s2[0][l] + s2[1][l]) + (s2[2][l] + s2[3][l]);
  }
  double o1 = (r1[0] + r1[1]) + (r1[2] + r1[3]);
  double o2 = (r2[0] + r2[1]) + (r2[2] + r2[3]);
  for (; pos < n; ++pos) {
    double sum = p1[pos] - n1[pos] + p2[pos];
    o1 += sum * x[pos];
    o2 += sum * y[pos];
  }
  *out1 = o1;
  *out2 = o2;
}

// util.rs:114-219  (p1 + p2) . x  and  . y
inline void scalar_prods2(const double* p1, const double* p2, const double* x, const double* y, size_t n, double* out1,
                          double* out2) {
  size_t nsimd = n / LANES, ngroups = nsimd / 4;
  double s1[4][LANES] = {}, s2[4][LANES] = {};
  size_t pos = 0;
  for (size_t g = 0; g < ngroups; ++g)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < LANES; ++l, ++pos) {
        double sum = p1[pos] + p2[pos];
        s1[k][l] = std::fma(sum, x[pos], s1[k][l]);
        s2[k][l] = std::fma(sum, y[pos], s2[k][l]);
      }
  for (size_t s = ngroups * 4; s < nsimd; ++s)
    for (int l = 0; l < LANES; ++l, ++pos) {
      double sum = p1[pos] + p2[pos];
      s1[0][l] = std::fma(sum, x[pos], s1[0][l]);
      s2[0][l] = std::fma(sum, y[pos], s2[0][l]);
    }
  double r1[LANES], r2[LANES];
  for (int l = 0; l < LANES; ++l) {
    r1[l] = (s1[0][l] + s1[1][l]) + (s1[2][l] + s1[3][l]);
    r2[l] = (s2[0][l] + s2[1][l]) + (s2[2][l] + s2[3][l]);
  }
  double o1 = (r1[0] + r1[1]) + (r1[2] + r1[3]);
  double o2 = (r2[0] + r2[1]) + (r2[2] + r2[3]);
  for (; pos < n; ++pos) {
    double sum = p1[pos] + p2[pos];
    o1 += sum * x[pos];
    o2 += sum * y[pos];
  }
  *out1 = o1;
  *out2 = o2;
}

// ------------------------------------------------------------------------------------------------
// src/math/cpu_math.rs — the CpuMath methods the diag path uses that are not plain forwards to util.rs
// ------------------------------------------------------------------------------------------------
// cpu_math.rs:235-243
inline double sq_norm_sum(const Vec& x, const Vec& y) {
  double s = 0.;
  for (size_t i = 0; i < x.size(); ++i) s += (x[i] + y[i]) * (x[i] + y[i]);
  return s;
}
// cpu_math.rs:283-287
inline bool array_all_finite(const Vec& a) {
  bool ok = true;
  for (double v : a) ok &= std::isfinite(v);
  return ok;
}
// cpu_math.rs:289-298
inline bool array_all_finite_and_nonzero(const Vec& a) {
  for (double v : a)
    if (!(std::isfinite(v) & (v != 0.))) return false;
  return true;
}
// cpu_math.rs:300-304
inline double array_sum_ln(const Vec& a) {
  double sum = 0.;
  for (double v : a) sum += std::log(v);
  return sum;
}
// cpu_math.rs:605-631  NOTE: both updates use the OLD mean (not textbook Welford)
inline void array_update_variance(Vec& mean, Vec& variance, const Vec& value, double diff_scale) {
  for (size_t i = 0; i < mean.size(); ++i) {
    double diff = value[i] - mean[i];
    mean[i] += diff * diff_scale;
    variance[i] += diff * diff;
  }
}
inline double clamp(double v, double lo, double hi) {  // f64::clamp
  if (v < lo) return lo;
  if (v > hi) return hi;
  return v;
}
// cpu_math.rs:633-669
inline void array_update_var_inv_std_draw(Vec& inv_std, Vec& std_, const Vec& draw_var, double scale,
                                          std::optional<double> fill_invalid, double lo, double hi) {
  for (size_t i = 0; i < std_.size(); ++i) {
    double dv = draw_var[i] * scale;
    if ((!std::isfinite(dv)) | (dv == 0.)) {
      if (fill_invalid) {
        std_[i] = std::sqrt(*fill_invalid);
        inv_std[i] = std::sqrt(1.0 / *fill_invalid);
      }
    } else {
      double val = clamp(dv, lo, hi);
      std_[i] = std::sqrt(val);
      inv_std[i] = std::sqrt(1.0 / val);
    }
  }
}
// cpu_math.rs:671-708
inline void array_update_var_inv_std_draw_grad(Vec& inv_std, Vec& std_, const Vec& draw_var, const Vec& grad_var,
                                               std::optional<double> fill_invalid, double lo, double hi) {
  for (size_t i = 0; i < std_.size(); ++i) {
    double val = std::sqrt(draw_var[i] / grad_var[i]);
    if ((!std::isfinite(val)) | (val == 0.)) {
      if (fill_invalid) {
        std_[i] = std::sqrt(*fill_invalid);
        inv_std[i] = std::sqrt(1.0 / *fill_invalid);
      }
    } else {
      val = clamp(val, lo, hi);
      std_[i] = std::sqrt(val);
      inv_std[i] = std::sqrt(1.0 / val);
    }
  }
}
// cpu_math.rs:710-738
inline void array_update_var_inv_std_grad(Vec& inv_std, Vec& std_, const Vec& gradient, double fill_invalid, double lo,
                                          double hi) {
  for (size_t i = 0; i < std_.size(); ++i) {
    double val = 1.0 / clamp(std::fabs(gradient[i]), lo, hi);
    if (!std::isfinite(val)) val = fill_invalid;
    std_[i] = std::sqrt(val);
    inv_std[i] = std::sqrt(1.0 / val);
  }
}

// ------------------------------------------------------------------------------------------------
// CpuLogpFunc stand-ins (SURVEY §8 a4'): the same closed-form targets the device evaluates.
// ------------------------------------------------------------------------------------------------
struct LogpFunc {
  size_t dim = 0;
  virtual ~LogpFunc() {}
  // returns logp, writes grad; non-finite results are the "recoverable error" channel (handled by the energy test)
  virtual double logp(const double* x, double* grad) = 0;
};

// src/math/test_logps.rs:49-58, benches/sample.rs:49-62, tests/sample_normal.rs:142-157
struct GaussIso : LogpFunc {
  Vec mu;
  double logp(const double* x, double* grad) override {
    double lp = 0.;
    for (size_t i = 0; i < dim; ++i) {
      double diff = x[i] - mu[i];
      lp -= diff * diff / 2.;
      grad[i] = -diff;
    }
    return lp;
  }
};
// diagonal generalisation: prec_i = 1/sigma_i^2 ;  grad = -(diff*prec) ; logp -= (diff*diff)*prec/2
struct GaussDiag : LogpFunc {
  Vec mu, prec;
  double logp(const double* x, double* grad) override {
    double lp = 0.;
    for (size_t i = 0; i < dim; ++i) {
      double diff = x[i] - mu[i];
      double pd = diff * prec[i];
      lp -= diff * pd / 2.;
      grad[i] = -pd;
    }
    return lp;
  }
};
// tests/sample_normal.rs:29-96
struct GaussRank1 : LogpFunc {
  Vec mu;
  double prec_rank1_coeff = 0.;
  double logp(const double* x, double* grad) override {
    double sum_diff = 0.;
    for (size_t i = 0; i < dim; ++i) sum_diff += x[i] - mu[i];
    double rank1_term = prec_rank1_coeff * sum_diff;
    double lp = 0.;
    for (size_t i = 0; i < dim; ++i) {
      double diff = x[i] - mu[i];
      double ptd = diff - rank1_term;
      grad[i] = -ptd;
      lp -= 0.5 * diff * ptd;
    }
    return lp;
  }
};
// Neal's funnel (BASELINE.json config 3): x0 = v ~ N(0, fs^2), x_i ~ N(0, e^v)
struct Funnel : LogpFunc {
  double fs = 3.0;
  double logp(const double* x, double* grad) override {
    double v = x[0];
    double S = 0.;
    for (size_t i = 1; i < dim; ++i) S = std::fma(x[i], x[i], S);
    double ev = std::exp(-v);
    double nm1 = (double)(dim - 1);
    double inv_var = 1.0 / (fs * fs);
    double half_ev_S = 0.5 * ev * S;
    double lp = -0.5 * v * v * inv_var - 0.5 * nm1 * v - half_ev_S;
    grad[0] = -v * inv_var - 0.5 * nm1 + half_ev_S;
    for (size_t i = 1; i < dim; ++i) grad[i] = -x[i] * ev;
    return lp;
  }
};

// ------------------------------------------------------------------------------------------------
// src/math/cpu_math.rs:332-425 — apply_lowrank_transform(_inplace):  dest = rhs + U ((vals - 1) .* (U^T rhs))
// vecs: r eigenvectors of length d, one after the other (the columns of the reference's d x r matrix).  faer's matmul leaves the
// summation order open; here every dot product runs left to right with FMAs.  rhs == dest is allowed (the in-place form).
// ------------------------------------------------------------------------------------------------
inline void apply_lowrank_transform(const Vec& vecs, const Vec& vals, const double* rhs, double* dest, size_t d) {
  const size_t r = vals.size();
  if (r == 0) {  // :339-342
    if (dest != rhs) std::copy(rhs, rhs + d, dest);
    return;
  }
  Vec scratch(r);
  for (size_t k = 0; k < r; ++k) {  // scratch = U^T rhs, then scratch[k] *= vals[k] - 1   (:353-366)
    double acc = 0.;
    for (size_t i = 0; i < d; ++i) acc = std::fma(vecs[k * d + i], rhs[i], acc);
    scratch[k] = acc * (vals[k] - 1.0);
  }
  for (size_t i = 0; i < d; ++i) {  // dest = rhs + U scratch   (:368-377)
    double acc = rhs[i];
    for (size_t k = 0; k < r; ++k) acc = std::fma(vecs[k * d + i], scratch[k], acc);
    dest[i] = acc;
  }
}

// src/transform/low_rank.rs:25-91 — InnerMatrix: U, lambda^{1/2}, lambda^{-1/2}, -1/2 sum ln(lambda), mu
struct LowRankInner {
  Vec vecs, vals_sqrt, vals_sqrt_inv, mu;
  double logdet_contribution = 0.;
  LowRankInner(const Vec& vals, const Vec& vecs_, const Vec& mu_) : vecs(vecs_), vals_sqrt(vals), vals_sqrt_inv(vals), mu(mu_) {
    for (double v : vals) logdet_contribution += -0.5 * std::log(v);   // :57
    for (double& v : vals_sqrt) v = std::sqrt(v);                      // :66
    for (size_t k = 0; k < vals.size(); ++k) vals_sqrt_inv[k] = 1.0 / vals_sqrt[k];  // :70
  }
};

// ------------------------------------------------------------------------------------------------
// src/transform/diagonal.rs — DiagMassMatrix; with `inner` set it is the LowRankMassMatrix of src/transform/low_rank.rs:97-404
// (diag + optional low-rank correction; every diagonal update drops the correction like update_from_grad, low_rank.rs:143-156)
// ------------------------------------------------------------------------------------------------
struct DiagMassMatrix {
  Vec mean, inv_stds, stds;
  double logdet = 0.;
  int64_t id = -1;
  std::shared_ptr<LowRankInner> inner;
  // LowRankMassMatrix::update (low_rank.rs:158-190): silently keeps the old transformation when an input is not finite
  bool update_lowrank(const Vec& stds_, const Vec& mean_, const Vec& vals, const Vec& vecs, const Vec& mean_low_rank) {
    auto finite = [](const Vec& v) {
      for (double x : v)
        if (!std::isfinite(x)) return false;
      return true;
    };
    if (!finite(stds_) || !finite(mean_) || !finite(vals) || !finite(vecs)) return false;
    const int64_t old_id = id;
    set_transform(stds_, mean_);  // self.diag.set_transform
    auto in = std::make_shared<LowRankInner>(vals, vecs, mean_low_rank);
    logdet = in->logdet_contribution + logdet;  // :187
    inner = in;
    id = old_id + 1;
    return true;
  }
  explicit DiagMassMatrix(size_t d) : mean(d, 0.), inv_stds(d, 0.), stds(d, 0.) {}  // diagonal.rs:73-83

  // diagonal.rs:85-105
  void update_diag_draw(const Vec& draw_mean, const Vec& draw_var, double scale, std::optional<double> fill, double lo,
                        double hi) {
    inner.reset();
    array_update_var_inv_std_draw(inv_stds, stds, draw_var, scale, fill, lo, hi);
    mean = draw_mean;
    logdet = array_sum_ln(inv_stds);
    id += 1;
  }
  // diagonal.rs:107-131
  void update_diag_draw_grad(const Vec& draw_mean, const Vec& grad_mean, const Vec& draw_var, const Vec& grad_var,
                             std::optional<double> fill, double lo, double hi) {
    inner.reset();
    array_update_var_inv_std_draw_grad(inv_stds, stds, draw_var, grad_var, fill, lo, hi);
    size_t d = stds.size();
    Vec var(d);
    multiply(stds.data(), stds.data(), var.data(), d);
    multiply(var.data(), grad_mean.data(), mean.data(), d);
    axpy(draw_mean.data(), mean.data(), 1.0, d);
    logdet = array_sum_ln(inv_stds);
    id += 1;
  }
  // diagonal.rs:133-154
  void update_diag_grad(const Vec& position, const Vec& gradient, double fill, double lo, double hi) {
    inner.reset();
    array_update_var_inv_std_grad(inv_stds, stds, gradient, fill, lo, hi);
    size_t d = stds.size();
    Vec var(d);
    multiply(stds.data(), stds.data(), var.data(), d);
    multiply(var.data(), gradient.data(), mean.data(), d);
    axpy(position.data(), mean.data(), 1.0, d);
    logdet = array_sum_ln(inv_stds);
    id += 1;
  }
  // diagonal.rs:156-162
  void set_transform(const Vec& stds_, const Vec& mean_) {
    inner.reset();
    stds = stds_;
    mean = mean_;
    for (size_t i = 0; i < stds.size(); ++i) inv_stds[i] = 1.0 / stds[i];
    logdet = array_sum_ln(inv_stds);
    id += 1;
  }
  // diagonal.rs:233-246  z = (x - mu) * inv_std   (axpy_out with a=-1, then multiply_inplace)
  // (low_rank.rs:326-348: then z -= mu_lr and the lambda^{-1/2} correction)
  void compute_transformed_position(const Vec& x, Vec& z) const {
    axpy_out(mean.data(), x.data(), -1.0, z.data(), x.size());
    multiply_inplace(z.data(), inv_stds.data(), x.size());
    if (inner) {
      axpy(inner->mu.data(), z.data(), -1.0, z.size());
      apply_lowrank_transform(inner->vecs, inner->vals_sqrt_inv, z.data(), z.data(), z.size());
    }
  }
  // diagonal.rs:248-256  x = z*std ; x += mu   (low_rank.rs:350-378: x = ((I + U(sqrt(lambda) - 1)U^T) z + mu_lr) * std + mean)
  void compute_untransformed_position(const Vec& z, Vec& x) const {
    if (inner) {
      apply_lowrank_transform(inner->vecs, inner->vals_sqrt, z.data(), x.data(), z.size());
      axpy(inner->mu.data(), x.data(), 1.0, z.size());
      multiply_inplace(x.data(), stds.data(), z.size());
    } else {
      multiply(z.data(), stds.data(), x.data(), z.size());
    }
    axpy(mean.data(), x.data(), 1.0, z.size());
  }
  // diagonal.rs:258-265   (low_rank.rs:380-398: then the sqrt(lambda) correction)
  void compute_transformed_gradient(const Vec& gx, Vec& gz) const {
    multiply(gx.data(), stds.data(), gz.data(), gx.size());
    if (inner) apply_lowrank_transform(inner->vecs, inner->vals_sqrt, gz.data(), gz.data(), gz.size());
  }
};

// ------------------------------------------------------------------------------------------------
// src/dynamics/transformed_hamiltonian.rs — TransformedPoint / TransformedHamiltonian (Euclidean only)
// ------------------------------------------------------------------------------------------------
struct TransformedPoint {  // :56-77
  Vec untransformed_position, untransformed_gradient, transformed_position, transformed_gradient, velocity;
  int64_t index_in_trajectory = 0;
  double logp = 0., logdet = 0., kinetic_energy = 0., initial_energy = 0.;
  int64_t transform_id = -1;
  double step_size_factor = 1.0;
  explicit TransformedPoint(size_t d)
      : untransformed_position(d, 0.), untransformed_gradient(d, 0.), transformed_position(d, 0.), transformed_gradient(d, 0.),
        velocity(d, 0.) {}
  double energy() const { return kinetic_energy - (logp + logdet); }       // :349-351
  double energy_error() const { return energy() - initial_energy; }        // hamiltonian.rs:134-136
  void update_kinetic_energy() {                                            // :260-262
    kinetic_energy = 0.5 * vector_dot(velocity.data(), velocity.data(), velocity.size());
  }
  bool check_untransformed() const {  // :300-308
    return array_all_finite(untransformed_gradient) && array_all_finite(untransformed_position);
  }
  bool check_all() const {  // :310-324
    return array_all_finite(transformed_position) && array_all_finite_and_nonzero(transformed_gradient) &&
           array_all_finite(untransformed_gradient) && array_all_finite(untransformed_position);
  }
};
using State = std::shared_ptr<TransformedPoint>;

// src/dynamics/state.rs:11-116 — StatePool: a State that loses its last reference goes back to the free list of the pool it
// came from and is handed out again by new_state() WITHOUT being cleared (every field is overwritten by whoever fills it:
// leapfrog :532-588, init_state :645-647), so a leapfrog allocates nothing once the pool is warm.
struct StatePool {
  struct Storage {
    size_t dim;
    std::vector<TransformedPoint*> free_states;
    ~Storage() {
      for (TransformedPoint* p : free_states) delete p;
    }
  };
  std::shared_ptr<Storage> storage;
  StatePool(size_t dim, size_t capacity) : storage(std::make_shared<Storage>()) {  // state.rs:27-32
    storage->dim = dim;
    storage->free_states.reserve(capacity);
  }
  State new_state() const {  // state.rs:34-42
    TransformedPoint* p;
    if (!storage->free_states.empty()) {
      p = storage->free_states.back();
      storage->free_states.pop_back();
    } else {
      p = new TransformedPoint(storage->dim);
    }
    std::weak_ptr<Storage> reuser = storage;
    return State(p, [reuser](TransformedPoint* q) {  // Drop for State, state.rs:105-116
      if (auto st = reuser.lock()) st->free_states.push_back(q);
      else delete q;
    });
  }
};

struct DivergenceInfo {  // hamiltonian.rs:26-35 (the fields the stats use)
  bool logp_error = false;
  int64_t start_idx = 0, end_idx = 0;
  double energy_error = std::numeric_limits<double>::quiet_NaN();
};

struct NutsOptions {  // src/nuts.rs:257-279
  uint64_t maxdepth = 10, mindepth = 0;
  bool check_turning = true;
  bool store_divergences = false;
  std::optional<double> target_integration_time;
  uint64_t extra_doublings = 0;
  double max_energy_error = 1000.0;
};

// src/stepsize/dual_avg.rs:84-166
struct RunningMean {
  double sum = 0.;
  uint64_t count = 0;
  void add(double v) { sum += v; count += 1; }
  double current() const { return sum / (double)count; }
  void reset() { sum = 0.; count = 0; }
};
struct AcceptanceRateCollector {
  double initial_energy = 0.;
  RunningMean mean, mean_sym;
  double max_energy_error = 0.;
  void register_leapfrog(const TransformedPoint& end, bool divergent) {  // :131-158
    if (divergent) {
      mean.add(0.);
      mean_sym.add(0.);
      max_energy_error = -std::numeric_limits<double>::infinity();
    } else {
      double diff = initial_energy - end.energy();
      mean.add(std::exp(std::fmin(diff, 0.)));
      mean_sym.add(2. * std::exp(std::fmin(diff, 0.)) / (1. + std::exp(diff)));
      if (std::fabs(diff) > std::fabs(max_energy_error)) max_energy_error = diff;
    }
  }
  void register_init(const TransformedPoint& state) {  // :160-165
    initial_energy = state.energy();
    mean.reset();
    mean_sym.reset();
    max_energy_error = 0.;
  }
};

struct SampleInfo {  // src/nuts.rs:46-57
  uint64_t depth = 0;
  std::optional<DivergenceInfo> divergence_info;
  bool reached_maxdepth = false;
};

// src/transform/adapt/diagonal.rs:57-84
struct DrawGradCollector {
  Vec draw, grad;
  bool is_good = true;
  explicit DrawGradCollector(size_t d) : draw(d, 0.), grad(d, 0.) {}
  void register_draw(const TransformedPoint& p, const SampleInfo& info) {
    draw = p.untransformed_position;
    grad = p.untransformed_gradient;
    int64_t idx = p.index_in_trajectory;
    if (info.divergence_info) is_good = std::llabs(idx) > 4;
    else is_good = idx != 0;
  }
};

// src/adapt_strategy.rs:286-350
struct CombinedCollector {
  AcceptanceRateCollector collector1;
  DrawGradCollector collector2;
  explicit CombinedCollector(size_t d) : collector2(d) {}
  void register_leapfrog(const TransformedPoint& end, bool divergent) { collector1.register_leapfrog(end, divergent); }
  void register_draw(const TransformedPoint& p, const SampleInfo& info) { collector2.register_draw(p, info); }
  void register_init(const TransformedPoint& p) { collector1.register_init(p); }
};

enum class Direction { Forward, Backward };

struct LeapfrogResult {
  enum Kind { Ok, Divergence } kind = Ok;
  State state;  // set for Ok (and for energy divergences, for white-box tests)
  DivergenceInfo info;
};

struct BadInitGrad : std::runtime_error {
  BadInitGrad() : std::runtime_error("Invalid initial point") {}
};

enum class KineticEnergyKind { Euclidean = 0, ExactNormal = 1, Microcanonical = 2 };  // :22-50

struct TransformedHamiltonian {
  size_t dim;
  LogpFunc* logp_func;
  Vec ones, zeros;
  double step_size = 0.;
  KineticEnergyKind kinetic_energy_kind = KineticEnergyKind::Euclidean;  // :381, default :434
  DiagMassMatrix transformation;
  StatePool pool;  // :427 StatePool::new(math, 10)
  uint64_t n_logp_evals = 0, n_leapfrogs = 0;

  TransformedHamiltonian(LogpFunc* f)  // :420-436
      : dim(f->dim), logp_func(f), ones(f->dim, 1.), zeros(f->dim, 0.), transformation(f->dim), pool(f->dim, 10) {}

  double logp_array(const Vec& x, Vec& grad) {  // cpu_math.rs:126-141
    n_logp_evals += 1;
    return logp_func->logp(x.data(), grad.data());
  }

  // :264-280 via diagonal.rs:182-194
  void init_from_untransformed_position(TransformedPoint& p) {
    p.logp = logp_array(p.untransformed_position, p.untransformed_gradient);
    transformation.compute_transformed_position(p.untransformed_position, p.transformed_position);
    transformation.compute_transformed_gradient(p.untransformed_gradient, p.transformed_gradient);
    p.logdet = transformation.logdet;
    p.transform_id = transformation.id;
  }
  // :282-298 via diagonal.rs:196-208
  void init_from_transformed_position(TransformedPoint& p) {
    transformation.compute_untransformed_position(p.transformed_position, p.untransformed_position);
    p.logp = logp_array(p.untransformed_position, p.untransformed_gradient);
    transformation.compute_transformed_gradient(p.untransformed_gradient, p.transformed_gradient);
    p.logdet = transformation.logdet;
    p.transform_id = transformation.id;
  }

  // :524-615 (Euclidean branches :178-184, :220-225, :245-247)
  template <class Collector>
  LeapfrogResult leapfrog(const State& start, Direction dir, double step_size_factor, double energy_baseline,
                          double max_energy_error, Collector& collector) {
    n_leapfrogs += 1;
    State out = pool.new_state();  // :532 self.pool().new_state(math)
    TransformedPoint& o = *out;
    const TransformedPoint& s = *start;
    o.initial_energy = s.initial_energy;
    o.transform_id = s.transform_id;
    int sign = dir == Direction::Forward ? 1 : -1;
    double epsilon = (double)sign * step_size * step_size_factor;
    o.step_size_factor = step_size_factor;
    const KineticEnergyKind kind = kinetic_energy_kind;
    const double sqrt_d = std::sqrt((double)dim);
    // first velocity half-step (:160-199)
    if (kind == KineticEnergyKind::ExactNormal) {
      std_norm_grad_flow(s.transformed_position.data(), s.transformed_gradient.data(), s.velocity.data(), o.velocity.data(),
                         epsilon / 2., dim);
    } else if (kind == KineticEnergyKind::Euclidean) {
      axpy_out(s.transformed_gradient.data(), s.velocity.data(), epsilon / 2., o.velocity.data(), dim);  // v_out = (eps/2)*grad_z + v
    } else {
      o.velocity = s.velocity;
      o.kinetic_energy = s.kinetic_energy + esh_momentum_update(s.transformed_gradient.data(), o.velocity.data(), sqrt_d * epsilon / 2., dim);
    }
    // position step (:201-228)
    if (kind == KineticEnergyKind::ExactNormal) {
      std_norm_flow(s.transformed_position.data(), o.transformed_position.data(), o.velocity.data(), epsilon, dim);
    } else {
      const double e = kind == KineticEnergyKind::Microcanonical ? epsilon * sqrt_d : epsilon;
      axpy_out(o.velocity.data(), s.transformed_position.data(), e, o.transformed_position.data(), dim);  // z_out = eps*v_out + z
    }
    init_from_transformed_position(o);
    // second velocity half-step (:230-258); Microcanonical keeps the accumulated delta KE, the others recompute 1/2 |v|^2 (:582-586)
    if (kind == KineticEnergyKind::ExactNormal) {
      std_norm_grad_flow_inplace(o.transformed_position.data(), o.transformed_gradient.data(), o.velocity.data(), epsilon / 2., dim);
    } else if (kind == KineticEnergyKind::Euclidean) {
      axpy(o.transformed_gradient.data(), o.velocity.data(), epsilon / 2., dim);
    } else {
      o.kinetic_energy = o.kinetic_energy + esh_momentum_update(o.transformed_gradient.data(), o.velocity.data(), sqrt_d * epsilon / 2., dim);
    }
    if (kind != KineticEnergyKind::Microcanonical) o.update_kinetic_energy();
    o.index_in_trajectory = s.index_in_trajectory + sign;
    double energy_error = o.energy() - energy_baseline;
    bool bad_energy = kind == KineticEnergyKind::Microcanonical ? std::fabs(energy_error) >= max_energy_error  // :591-596
                                                                : energy_error > max_energy_error;
    LeapfrogResult res;
    if (bad_energy | !std::isfinite(energy_error)) {
      res.kind = LeapfrogResult::Divergence;
      res.info.start_idx = s.index_in_trajectory;
      res.info.end_idx = o.index_in_trajectory;
      res.info.energy_error = energy_error;
      res.state = out;
      collector.register_leapfrog(o, true);
      return res;
    }
    collector.register_leapfrog(o, false);
    res.state = out;
    return res;
  }

  // :617-638
  bool is_turning(const TransformedPoint& s1, const TransformedPoint& s2) const {
    const TransformedPoint* start = &s2;
    const TransformedPoint* end = &s1;
    if (s1.index_in_trajectory < s2.index_in_trajectory) {
      start = &s1;
      end = &s2;
    }
    double t1, t2;
    scalar_prods3(end->transformed_position.data(), start->transformed_position.data(), zeros.data(), start->velocity.data(),
                  end->velocity.data(), dim, &t1, &t2);
    return (t1 < 0.) | (t2 < 0.);
  }

  // :640-661
  State init_state(const double* init) {
    State st = pool.new_state();
    std::copy(init, init + dim, st->untransformed_position.begin());
    init_from_untransformed_position(*st);
    if (!st->check_all()) throw BadInitGrad();
    return st;
  }
  // :663-685
  State init_state_untransformed(const double* x) {
    State st = pool.new_state();
    std::copy(x, x + dim, st->untransformed_position.begin());
    st->logp = logp_array(st->untransformed_position, st->untransformed_gradient);
    st->transform_id = -1;
    if (!st->check_untransformed()) throw BadInitGrad();
    return st;
  }
  // :687-736
  void initialize_trajectory(TransformedPoint& p, bool resample_velocity, Rng& rng) {
    if (resample_velocity) {
      // cpu_math.rs:561-577 array_gaussian(rng, velocity, ones): v[i] = 1.0 * normal, sequential in i
      rng.fill_normal(p.velocity.data(), dim);
      for (size_t i = 0; i < dim; ++i) p.velocity[i] = ones[i] * p.velocity[i];
      if (kinetic_energy_kind == KineticEnergyKind::Microcanonical) array_normalize(p.velocity.data(), dim);  // :699-702
    }
    if (transformation.id != p.transform_id) {
      // diagonal.rs:210-221 inv_transform_normalize: no logp evaluation
      transformation.compute_transformed_position(p.untransformed_position, p.transformed_position);
      transformation.compute_transformed_gradient(p.untransformed_gradient, p.transformed_gradient);
      p.logdet = transformation.logdet;
      p.transform_id = transformation.id;
    }
    if (kinetic_energy_kind == KineticEnergyKind::Microcanonical) p.kinetic_energy = 0.0;  // :720-729
    else p.update_kinetic_energy();
    p.index_in_trajectory = 0;
    p.initial_energy = p.energy();
  }
};

// ------------------------------------------------------------------------------------------------
// src/nuts.rs — NutsTree and draw()
// ------------------------------------------------------------------------------------------------
struct NutsTree {  // :60-78
  State left, right, draw;
  double log_size = 0.;
  uint64_t depth = 0;
  bool is_main = true;
};

struct ExtendResult {  // :80-91
  enum Kind { Ok, Turning, Diverging } kind = Ok;
  NutsTree tree;
  DivergenceInfo info;
};

template <class Collector>
struct TreeBuilder {
  TransformedHamiltonian& h;
  Rng& rng;
  Collector& collector;

  // :209-245
  bool single_step(const NutsTree& self, Direction direction, const NutsOptions& options, NutsTree* out, DivergenceInfo* info) {
    const State& start = direction == Direction::Forward ? self.right : self.left;
    LeapfrogResult r = h.leapfrog(start, direction, 1.0, start->initial_energy, options.max_energy_error, collector);
    if (r.kind == LeapfrogResult::Divergence) {
      *info = r.info;
      return false;
    }
    double log_size = -r.state->energy_error();
    *out = NutsTree{r.state, r.state, r.state, log_size, 0, false};
    return true;
  }

  // :172-207
  void merge_into(NutsTree& self, NutsTree&& other, Direction direction) {
    assert(self.depth == other.depth);
    assert(self.left->index_in_trajectory <= self.right->index_in_trajectory);
    if (direction == Direction::Forward) self.right = other.right;
    else self.left = other.left;
    double log_size = logaddexp(self.log_size, other.log_size);
    double self_log_size = self.is_main ? self.log_size : log_size;
    if ((other.log_size >= self_log_size) || rng.random_bool(std::exp(other.log_size - self_log_size))) self.draw = other.draw;
    self.depth += 1;
    self.log_size = log_size;
  }

  // :108-170
  ExtendResult extend(NutsTree self, Direction direction, const NutsOptions& options) {
    NutsTree other;
    DivergenceInfo info;
    if (!single_step(self, direction, options, &other, &info)) return ExtendResult{ExtendResult::Diverging, std::move(self), info};
    while (other.depth < self.depth) {
      ExtendResult r = extend(std::move(other), direction, options);
      if (r.kind == ExtendResult::Ok) other = std::move(r.tree);
      else if (r.kind == ExtendResult::Turning) return ExtendResult{ExtendResult::Turning, std::move(self), {}};
      else return ExtendResult{ExtendResult::Diverging, std::move(self), r.info};
    }
    const State& first = direction == Direction::Forward ? self.left : other.left;
    const State& last = direction == Direction::Forward ? other.right : self.right;
    bool turning = false;
    if (options.check_turning) {
      turning = h.is_turning(*first, *last);
      if (self.depth > 0) {
        if (!turning) turning = h.is_turning(*self.right, *other.right);
        if (!turning) turning = h.is_turning(*self.left, *other.left);
      }
    }
    merge_into(self, std::move(other), direction);
    return ExtendResult{turning ? ExtendResult::Turning : ExtendResult::Ok, std::move(self), {}};
  }
};

// :281-388
template <class Collector>
std::pair<State, SampleInfo> nuts_draw(State& init, Rng& rng, TransformedHamiltonian& h, const NutsOptions& options,
                                       Collector& collector) {
  h.initialize_trajectory(*init, true, rng);
  collector.register_init(*init);
  NutsTree tree{init, init, init, 0., 0, true};
  uint64_t mindepth = options.mindepth, maxdepth = options.maxdepth;
  if (options.target_integration_time) {
    double step_size = h.step_size;
    uint64_t max_steps = (uint64_t)std::ceil(*options.target_integration_time / step_size);
    mindepth = std::max((uint64_t)std::floor(std::log2((double)max_steps)), options.mindepth);
    maxdepth = std::min(std::max((uint64_t)std::ceil(std::log2((double)max_steps)), mindepth), options.maxdepth);
  }
  SampleInfo info;
  if (h.dim == 0) {
    info.depth = tree.depth;
    collector.register_draw(*init, info);
    return {init, info};
  }
  NutsOptions options_no_check = options;
  options_no_check.check_turning = false;
  TreeBuilder<Collector> tb{h, rng, collector};
  while (tree.depth < maxdepth) {
    Direction direction = rng.next_bool() ? Direction::Forward : Direction::Backward;  // hamiltonian.rs:111-119
    const NutsOptions& cur = tree.depth < mindepth ? options_no_check : options;
    ExtendResult r = tb.extend(std::move(tree), direction, cur);
    if (r.kind == ExtendResult::Ok) {
      tree = std::move(r.tree);
    } else if (r.kind == ExtendResult::Turning) {
      tree = std::move(r.tree);
      for (uint64_t k = 0; k < options.extra_doublings; ++k) {
        ExtendResult r2 = tb.extend(std::move(tree), direction, options_no_check);
        if (r2.kind == ExtendResult::Diverging) {
          tree = std::move(r2.tree);
          info.depth = tree.depth;
          info.divergence_info = r2.info;
          collector.register_draw(*tree.draw, info);
          return {tree.draw, info};
        }
        tree = std::move(r2.tree);
      }
      info.depth = tree.depth;
      collector.register_draw(*tree.draw, info);
      return {tree.draw, info};
    } else {
      tree = std::move(r.tree);
      info.depth = tree.depth;
      info.divergence_info = r.info;
      collector.register_draw(*tree.draw, info);
      return {tree.draw, info};
    }
  }
  info.depth = tree.depth;
  info.reached_maxdepth = true;
  collector.register_draw(*tree.draw, info);
  return {tree.draw, info};
}

// ------------------------------------------------------------------------------------------------
// src/stepsize/dual_avg.rs:11-81, src/stepsize/adapt.rs
// ------------------------------------------------------------------------------------------------
struct DualAverageOptions {
  double k = 0.75, t0 = 10., gamma = 0.05, max_step_size = 3.14159265358979323846;
};
struct DualAverage {
  double log_step, log_step_adapted, hbar, mu;
  uint64_t count;
  DualAverageOptions settings;
  DualAverage(DualAverageOptions s, double initial_step)
      : log_step(std::log(initial_step)), log_step_adapted(std::log(initial_step)), hbar(0.), mu(std::log(10. * initial_step)),
        count(1), settings(s) {}
  void advance(double accept_stat, double target) {  // :55-63
    double w = 1. / ((double)count + settings.t0);
    hbar = (1. - w) * hbar + w * (target - accept_stat);
    log_step = mu - hbar * std::sqrt((double)count) / settings.gamma;
    log_step = std::fmin(log_step, std::log(settings.max_step_size));
    double mk = std::pow((double)count, -settings.k);
    log_step_adapted = mk * log_step + (1. - mk) * log_step_adapted;
    count += 1;
  }
  double current_step_size() const { return std::exp(log_step); }
  double current_step_size_adapted() const { return std::exp(log_step_adapted); }
};

// src/stepsize/adam.rs
struct AdamOptions {
  double beta1 = 0.9, beta2 = 0.999, epsilon = 1e-8, learning_rate = 0.05;  // :25-34
};
inline double powi(double a, int b) {  // f64::powi -> llvm.powi -> compiler-rt __powidf2 (square and multiply)
  const bool recip = b < 0;
  double r = 1.0;
  for (;;) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.0 / r : r;
}
struct Adam {
  double log_step, m = 0., v = 0.;
  uint64_t t = 0;
  AdamOptions settings;
  Adam(AdamOptions s, double initial_step) : log_step(std::log(initial_step)), settings(s) {}  // :57-65
  void advance(double accept_stat, double target) {  // :71-97
    double gradient = accept_stat - target;
    t += 1;
    m = settings.beta1 * m + (1.0 - settings.beta1) * gradient;
    v = settings.beta2 * v + (1.0 - settings.beta2) * gradient * gradient;
    double m_hat = m / (1.0 - powi(settings.beta1, (int)t));
    double v_hat = v / (1.0 - powi(settings.beta2, (int)t));
    log_step += settings.learning_rate * m_hat / (std::sqrt(v_hat) + settings.epsilon);
  }
  double current_step_size() const { return std::exp(log_step); }
};

enum class StepSizeMethod { DualAverage = 0, Adam = 1, Fixed = 2 };
struct StepSizeSettings {  // adapt.rs:308-329
  double target_accept = 0.8, initial_step = 0.1;
  std::optional<double> jitter = 0.1;
  StepSizeMethod method = StepSizeMethod::DualAverage;
  double fixed_step = 0.;
  DualAverageOptions dual_average;
  AdamOptions adam;
};

struct StepSizeStrategy {  // adapt.rs:52-267
  std::optional<DualAverage> adaptation;
  std::optional<Adam> adam;  // Either::Right
  StepSizeSettings options;
  double last_mean_tree_accept = 0., last_sym_mean_tree_accept = 0., last_max_energy_error = 0.;
  uint64_t last_n_steps = 0;
  explicit StepSizeStrategy(StepSizeSettings o) : options(o) {  // :67-89
    if (o.method == StepSizeMethod::DualAverage) adaptation = DualAverage(o.dual_average, o.initial_step);
    if (o.method == StepSizeMethod::Adam) adam = Adam(o.adam, o.initial_step);
  }
  // :91-199
  void init(TransformedHamiltonian& h, const double* position, Rng& rng) {
    if (options.method == StepSizeMethod::Fixed) {
      h.step_size = options.fixed_step;
      return;
    }
    State state = h.init_state(position);
    h.initialize_trajectory(*state, true, rng);
    AcceptanceRateCollector collector;
    collector.register_init(*state);
    h.step_size = options.initial_step;
    LeapfrogResult r = h.leapfrog(state, Direction::Forward, 1.0, state->initial_energy, 1000.0, collector);
    if (r.kind != LeapfrogResult::Ok) return;
    double accept_stat = collector.mean.current();
    Direction dir = accept_stat > options.target_accept ? Direction::Forward : Direction::Backward;
    for (int it = 0; it < 100; ++it) {
      AcceptanceRateCollector c;
      c.register_init(*state);
      LeapfrogResult r2 = h.leapfrog(state, dir, 1.0, state->initial_energy, 1000.0, c);
      if (r2.kind != LeapfrogResult::Ok) {
        h.step_size = options.initial_step;
        return;
      }
      double acc = c.mean.current();
      if (dir == Direction::Forward) {
        if ((acc <= options.target_accept) | (h.step_size > 1e5)) {
          reset_adaptation(h.step_size);
          return;
        }
        h.step_size *= 2.;
      } else {
        if ((acc >= options.target_accept) | (h.step_size < 1e-10)) {
          reset_adaptation(h.step_size);
          return;
        }
        h.step_size /= 2.;
      }
    }
    h.step_size = options.initial_step;
  }
  void reset_adaptation(double step) {  // :155-171 / :177-193
    if (options.method == StepSizeMethod::Adam) adam = Adam(options.adam, step);
    else adaptation = DualAverage(options.dual_average, step);
  }
  void update(const AcceptanceRateCollector& c) {  // :201-209
    last_sym_mean_tree_accept = c.mean_sym.current();
    last_mean_tree_accept = c.mean.current();
    last_n_steps = c.mean.count;
    last_max_energy_error = c.max_energy_error;
  }
  void update_estimator_early() {  // :211-221
    if (adaptation) adaptation->advance(last_mean_tree_accept, options.target_accept);
    if (adam) adam->advance(last_mean_tree_accept, options.target_accept);
  }
  void update_estimator_late() {  // :223-233
    if (adaptation) adaptation->advance(last_sym_mean_tree_accept, options.target_accept);
    if (adam) adam->advance(last_sym_mean_tree_accept, options.target_accept);
  }
  void update_stepsize(Rng& rng, TransformedHamiltonian& h, bool use_best_guess) {  // :235-267
    double step_size;
    if (adam) step_size = adam->current_step_size();  // :256
    else if (!adaptation) step_size = options.fixed_step;
    else step_size = use_best_guess ? adaptation->current_step_size_adapted() : adaptation->current_step_size();
    if (options.jitter) {
      double j = rng.uniform(1.0 - *options.jitter, 1.0 + *options.jitter);
      h.step_size = step_size * j;
    } else {
      h.step_size = step_size;
    }
  }
  double step_size_bar() const {  // :278-290
    if (adam) return adam->current_step_size();
    return adaptation ? adaptation->current_step_size_adapted() : options.fixed_step;
  }
};

// ------------------------------------------------------------------------------------------------
// src/transform/adapt/diagonal.rs — RunningVariance and the diagonal adaptation Strategy
// ------------------------------------------------------------------------------------------------
struct RunningVariance {  // :17-55
  Vec mean, variance;
  uint64_t count = 0;
  explicit RunningVariance(size_t d) : mean(d, 0.), variance(d, 0.) {}
  void add_sample(const Vec& value) {
    count += 1;
    if (count == 1) mean = value;
    else array_update_variance(mean, variance, value, 1.0 / (double)count);
  }
};

struct DiagAdaptStrategy {  // :108-231
  size_t dim;
  RunningVariance exp_variance_draw, exp_variance_grad, exp_variance_grad_bg, exp_variance_draw_bg;
  bool use_grad_based_estimate = true;
  static constexpr double LOWER_LIMIT = 1e-20, UPPER_LIMIT = 1e20;
  DiagAdaptStrategy(size_t d, bool use_grad)
      : dim(d), exp_variance_draw(d), exp_variance_grad(d), exp_variance_grad_bg(d), exp_variance_draw_bg(d),
        use_grad_based_estimate(use_grad) {}
  void update_estimators(const DrawGradCollector& c) {  // :134-141
    if (c.is_good) {
      exp_variance_draw.add_sample(c.draw);
      exp_variance_grad.add_sample(c.grad);
      exp_variance_draw_bg.add_sample(c.draw);
      exp_variance_grad_bg.add_sample(c.grad);
    }
  }
  void do_switch() {  // :143-148
    exp_variance_draw = std::move(exp_variance_draw_bg);
    exp_variance_draw_bg = RunningVariance(dim);
    exp_variance_grad = std::move(exp_variance_grad_bg);
    exp_variance_grad_bg = RunningVariance(dim);
  }
  uint64_t current_count() const { return exp_variance_draw.count; }
  uint64_t background_count() const { return exp_variance_draw_bg.count; }
  bool adapt(DiagMassMatrix& m) const {  // :161-196
    if (current_count() < 3) return false;
    if (use_grad_based_estimate) {
      m.update_diag_draw_grad(exp_variance_draw.mean, exp_variance_grad.mean, exp_variance_draw.variance, exp_variance_grad.variance,
                              std::nullopt, LOWER_LIMIT, UPPER_LIMIT);
    } else {
      double scale = 1.0 / (double)exp_variance_draw.count;
      m.update_diag_draw(exp_variance_draw.mean, exp_variance_draw.variance, scale, std::nullopt, LOWER_LIMIT, UPPER_LIMIT);
    }
    return true;
  }
  void init(DiagMassMatrix& m, const TransformedPoint& point) {  // :209-231
    exp_variance_draw.add_sample(point.untransformed_position);
    exp_variance_draw_bg.add_sample(point.untransformed_position);
    exp_variance_grad.add_sample(point.untransformed_gradient);
    exp_variance_grad_bg.add_sample(point.untransformed_gradient);
    m.update_diag_grad(point.untransformed_position, point.untransformed_gradient, 1., 1e-20, 1e20);
  }
};

// ------------------------------------------------------------------------------------------------
// src/adapt_strategy.rs — GlobalStrategy
// ------------------------------------------------------------------------------------------------
struct EuclideanAdaptOptions {  // :41-69
  StepSizeSettings step_size_settings;
  bool use_grad_based_estimate = true;
  double early_window = 0.3, step_size_window = 0.15;
  uint64_t mass_matrix_switch_freq = 80, early_mass_matrix_switch_freq = 10, mass_matrix_update_freq = 1;
  double mass_matrix_window_growth = 1.5;
};

struct GlobalStrategy {
  StepSizeStrategy step_size;
  DiagAdaptStrategy mass_matrix_adapt;
  EuclideanAdaptOptions options;
  uint64_t num_tune, early_end, final_step_size_window;
  bool tuning = true, has_initial_mass_matrix = true;
  uint64_t last_update = 0, current_window_size;

  GlobalStrategy(size_t dim, EuclideanAdaptOptions o, uint64_t num_tune_)  // :77-98
      : step_size(o.step_size_settings), mass_matrix_adapt(dim, o.use_grad_based_estimate), options(o), num_tune(num_tune_) {
    double num_tune_f = (double)num_tune;
    uint64_t step_size_window = (uint64_t)(o.step_size_window * num_tune_f);
    early_end = (uint64_t)(o.early_window * num_tune_f);
    final_step_size_window = num_tune >= step_size_window ? num_tune - step_size_window : 0;
    current_window_size = o.mass_matrix_switch_freq;
  }
  // :100-119
  void init(TransformedHamiltonian& h, const double* position, Rng& rng) {
    State state = h.init_state_untransformed(position);
    mass_matrix_adapt.init(h.transformation, *state);
    step_size.init(h, position, rng);
  }
  // :121-222
  void adapt(TransformedHamiltonian& h, uint64_t draw, const CombinedCollector& collector, const State& state, Rng& rng) {
    step_size.update(collector.collector1);
    if (draw >= num_tune) {
      step_size.update_stepsize(rng, h, true);
      tuning = false;
      return;
    }
    if (draw < final_step_size_window) {
      bool is_early = draw < early_end;
      if (!is_early && draw == early_end) current_window_size = std::max(current_window_size, mass_matrix_adapt.background_count());
      uint64_t switch_freq = is_early ? options.early_mass_matrix_switch_freq : current_window_size;
      mass_matrix_adapt.update_estimators(collector.collector2);
      bool could_switch = mass_matrix_adapt.background_count() >= switch_freq;
      uint64_t next_window_size =
          is_early ? options.early_mass_matrix_switch_freq
                   : std::max(current_window_size + 1,
                              (uint64_t)std::round((double)current_window_size * options.mass_matrix_window_growth));
      bool is_late = next_window_size + draw > final_step_size_window;
      bool force_update = false;
      if (could_switch && !is_late) {
        mass_matrix_adapt.do_switch();
        force_update = true;
        if (!is_early) current_window_size = next_window_size;
      }
      bool did_change = false;
      if (force_update | (draw - last_update >= options.mass_matrix_update_freq)) did_change = mass_matrix_adapt.adapt(h.transformation);
      if (did_change) last_update = draw;
      if (is_late) step_size.update_estimator_late();
      else step_size.update_estimator_early();
      if (did_change & has_initial_mass_matrix) {
        has_initial_mass_matrix = false;
        Vec position = state->untransformed_position;
        step_size.init(h, position.data(), rng);
      } else {
        step_size.update_stepsize(rng, h, false);
      }
      return;
    }
    step_size.update_estimator_late();
    bool is_last = draw == num_tune - 1;
    step_size.update_stepsize(rng, h, is_last);
  }
};

// ------------------------------------------------------------------------------------------------
// src/chain.rs — NutsChain, plus the stats a draw exports
// ------------------------------------------------------------------------------------------------
struct DrawStats {
  uint64_t depth = 0;
  bool maxdepth_reached = false;
  int64_t index_in_trajectory = 0;
  double logp = 0., energy = 0., energy_error = 0.;
  bool diverging = false;
  double step_size = 0., step_size_bar = 0., mean_tree_accept = 0., mean_tree_accept_sym = 0.;
  uint64_t n_steps = 0;
  double max_energy_error = 0.;
  bool tuning = true;
  double fisher_distance = 0.;
};

struct NutsSettings {  // src/sampler.rs:199-239 + defaults :507-531,630-634
  uint64_t num_tune = 400, num_draws = 1000, maxdepth = 10, mindepth = 0;
  double max_energy_error = 1000.;
  EuclideanAdaptOptions adapt_options;
  bool check_turning = true;
  std::optional<double> target_integration_time;
  uint64_t num_chains = 6, seed = 0, extra_doublings = 0;
  KineticEnergyKind trajectory_kind = KineticEnergyKind::Euclidean;  // sampler.rs:232, default :528
};

struct NutsChain {  // chain.rs:44-61
  TransformedHamiltonian hamiltonian;
  CombinedCollector collector;
  NutsOptions options;
  Rng rng;
  State state;
  uint64_t chain, draw_count = 0;
  GlobalStrategy strategy;

  NutsChain(LogpFunc* f, const NutsSettings& s, uint64_t chain_id, Rng rng_)  // sampler.rs:745-772
      : hamiltonian(f), collector(f->dim), rng(rng_), chain(chain_id), strategy(f->dim, s.adapt_options, s.num_tune) {
    options.maxdepth = s.maxdepth;
    options.mindepth = s.mindepth;
    options.check_turning = s.check_turning;
    options.target_integration_time = s.target_integration_time;
    options.extra_doublings = s.extra_doublings;
    options.max_energy_error = s.max_energy_error;
    hamiltonian.kinetic_energy_kind = s.trajectory_kind;  // sampler.rs:757 TransformedHamiltonian::new(.., self.trajectory_kind)
    state = std::make_shared<TransformedPoint>(f->dim);
  }
  // chain.rs:137-149 ; throws BadInitGrad
  void set_position(const double* position) {
    strategy.init(hamiltonian, position, rng);
    state = hamiltonian.init_state(position);
  }
  // LowRankMassMatrixStrategy::update -> LowRankMassMatrix::update (adapt/low_rank.rs:54-71, low_rank.rs:158-190) with values the
  // caller computed, followed by what GlobalStrategy::adapt does after a mass-matrix change (adapt_strategy.rs:204-214): the first
  // change of a run re-initialises the step size from the current point.  The point is re-whitened by the next draw
  // (initialize_trajectory, transformation id changed).
  bool set_lowrank_transform(const Vec& stds, const Vec& mean, const Vec& vals, const Vec& vecs, const Vec& mean_low_rank) {
    if (!hamiltonian.transformation.update_lowrank(stds, mean, vals, vecs, mean_low_rank)) return false;
    if (strategy.has_initial_mass_matrix && strategy.tuning) {
      strategy.has_initial_mass_matrix = false;
      Vec position = state->untransformed_position;
      strategy.step_size.init(hamiltonian, position.data(), rng);
    }
    return true;
  }
  // chain.rs:151-188 (+ the stats of expanded_draw :190-204)
  DrawStats draw(double* position_out) {
    auto [st, info] = nuts_draw(state, rng, hamiltonian, options, collector);
    std::copy(st->untransformed_position.begin(), st->untransformed_position.end(), position_out);
    strategy.adapt(hamiltonian, draw_count, collector, st, rng);
    DrawStats ds;
    ds.depth = info.depth;
    ds.maxdepth_reached = info.reached_maxdepth;
    ds.diverging = info.divergence_info.has_value();
    ds.tuning = strategy.tuning;
    ds.step_size = hamiltonian.step_size;
    ds.n_steps = strategy.step_size.last_n_steps;
    ds.step_size_bar = strategy.step_size.step_size_bar();
    ds.mean_tree_accept = strategy.step_size.last_mean_tree_accept;
    ds.mean_tree_accept_sym = strategy.step_size.last_sym_mean_tree_accept;
    ds.max_energy_error = strategy.step_size.last_max_energy_error;
    draw_count += 1;
    state = st;
    ds.index_in_trajectory = st->index_in_trajectory;
    ds.logp = st->logp;
    ds.energy = st->energy();
    ds.energy_error = st->energy_error();
    ds.fisher_distance = sq_norm_sum(st->transformed_position, st->transformed_gradient);
    return ds;
  }
};

}  // namespace oracle
