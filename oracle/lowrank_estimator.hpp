// lowrank_estimator.hpp — TEST INFRASTRUCTURE: C++ restatement of the reference's low-rank mass-matrix estimator,
// LowRankMassMatrixStrategy::compute_update and its helpers (pymc-devs/nuts-rs src/transform/adapt/low_rank.rs:73-268).
// The reference uses faer (thin SVD, column-pivoted QR, self-adjoint eigendecomposition: an unpinned third-party crate, absent from
// /root/reference); here the three factorisations are textbook algorithms (one-sided Jacobi SVD, Householder QR with column
// pivoting, cyclic Jacobi eigensolver) - the estimator's RESULT only depends on the subspaces and on U diag(vals) U^T, which are
// unique, not on faer's sign / ordering conventions.  Pinned by the reference's known-answer tests (adapt/low_rank.rs:354-407) in
// tests/test_oracle_lowrank.py; used to cross-check the product's numpy estimator (nuts_rs_b200/lowrank.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace oracle {
namespace lowrank {

// column-major dense matrix
struct Mat {
  size_t r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(size_t r_, size_t c_) : r(r_), c(c_), a(r_ * c_, 0.0) {}
  double& operator()(size_t i, size_t j) { return a[j * r + i]; }
  double operator()(size_t i, size_t j) const { return a[j * r + i]; }
};
inline Mat matmul(const Mat& x, const Mat& y, bool tx = false, bool ty = false) {
  const size_t m = tx ? x.c : x.r, k = tx ? x.r : x.c, n = ty ? y.r : y.c;
  Mat o(m, n);
  for (size_t j = 0; j < n; ++j)
    for (size_t i = 0; i < m; ++i) {
      double s = 0.;
      for (size_t l = 0; l < k; ++l) s += (tx ? x(l, i) : x(i, l)) * (ty ? y(j, l) : y(l, j));
      o(i, j) = s;
    }
  return o;
}

// cyclic Jacobi: symmetric A = V diag(w) V^T, eigenvalues ascending (self_adjoint_eigen)
inline bool eigh(Mat a, std::vector<double>& w, Mat& v) {
  const size_t n = a.r;
  v = Mat(n, n);
  for (size_t i = 0; i < n; ++i) v(i, i) = 1.;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0., diag = 0.;
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) (i == j ? diag : off) += a(i, j) * a(i, j);
    if (!std::isfinite(off + diag)) return false;
    if (off <= 1e-30 * diag) break;
    for (size_t p = 0; p + 1 < n; ++p)
      for (size_t q = p + 1; q < n; ++q) {
        if (a(p, q) == 0.) continue;
        const double theta = (a(q, q) - a(p, p)) / (2. * a(p, q));
        const double t = (theta >= 0. ? 1. : -1.) / (std::fabs(theta) + std::sqrt(theta * theta + 1.));
        const double cs = 1. / std::sqrt(t * t + 1.), sn = t * cs;
        for (size_t k = 0; k < n; ++k) {
          const double akp = a(k, p), akq = a(k, q);
          a(k, p) = cs * akp - sn * akq;
          a(k, q) = sn * akp + cs * akq;
        }
        for (size_t k = 0; k < n; ++k) {
          const double apk = a(p, k), aqk = a(q, k);
          a(p, k) = cs * apk - sn * aqk;
          a(q, k) = sn * apk + cs * aqk;
        }
        for (size_t k = 0; k < n; ++k) {
          const double vkp = v(k, p), vkq = v(k, q);
          v(k, p) = cs * vkp - sn * vkq;
          v(k, q) = sn * vkp + cs * vkq;
        }
      }
  }
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return a(x, x) < a(y, y); });
  w.resize(n);
  Mat vs(n, n);
  for (size_t j = 0; j < n; ++j) {
    w[j] = a(order[j], order[j]);
    for (size_t i = 0; i < n; ++i) vs(i, j) = v(i, order[j]);
  }
  v = vs;
  return true;
}

// left singular vectors of the thin SVD of x [m x n] with non-zero singular values: through the eigenvectors of the smaller Gram
// matrix (the estimator only uses span(U)); columns with singular value below 1e-12 * largest are dropped (they carry no direction)
inline bool thin_svd_u(const Mat& x, Mat& u) {
  const size_t m = x.r, n = x.c, k = std::min(m, n);
  std::vector<double> w;
  Mat v;
  if (m <= n) {
    if (!eigh(matmul(x, x, false, true), w, v)) return false;
    u = Mat(m, k);
    size_t col = 0;
    for (size_t j = n > m ? 0 : 0; j < m; ++j) {
      const size_t src = m - 1 - j;  // descending
      if (w[src] <= 1e-24 * std::max(w[m - 1], 1e-300)) continue;
      for (size_t i = 0; i < m; ++i) u(i, col) = v(i, src);
      ++col;
    }
    u.c = col;
    u.a.resize(m * col);
    return true;
  }
  if (!eigh(matmul(x, x, true, false), w, v)) return false;  // n x n
  u = Mat(m, n);
  size_t col = 0;
  for (size_t j = 0; j < n; ++j) {
    const size_t src = n - 1 - j;
    if (w[src] <= 1e-24 * std::max(w[n - 1], 1e-300)) continue;
    const double s = std::sqrt(w[src]);
    for (size_t i = 0; i < m; ++i) {
      double acc = 0.;
      for (size_t l = 0; l < n; ++l) acc += x(i, l) * v(l, src);
      u(i, col) = acc / s;
    }
    ++col;
  }
  u.c = col;
  u.a.resize(m * col);
  return true;
}

// orthonormal basis of span(columns of s): modified Gram-Schmidt with column pivoting and re-orthogonalisation
// (col_piv_qr().compute_thin_Q(): any orthonormal basis of the same span gives the same estimator result)
inline Mat orth_basis(const Mat& s) {
  const size_t m = s.r, n = s.c;
  Mat w = s, q(m, std::min(m, n));
  std::vector<bool> used(n, false);
  size_t k = 0;
  double first = 0.;
  for (; k < std::min(m, n); ++k) {
    size_t best = n;
    double bn = 0.;
    for (size_t j = 0; j < n; ++j) {
      if (used[j]) continue;
      double nn = 0.;
      for (size_t i = 0; i < m; ++i) nn += w(i, j) * w(i, j);
      if (best == n || nn > bn) best = j, bn = nn;
    }
    if (best == n) break;
    if (k == 0) first = bn;
    if (bn <= 1e-20 * std::max(first, 1e-300)) break;
    used[best] = true;
    for (size_t i = 0; i < m; ++i) q(i, k) = w(i, best);
    for (int pass = 0; pass < 2; ++pass)
      for (size_t p = 0; p < k; ++p) {
        double dot = 0.;
        for (size_t i = 0; i < m; ++i) dot += q(i, p) * q(i, k);
        for (size_t i = 0; i < m; ++i) q(i, k) -= dot * q(i, p);
      }
    double nn = 0.;
    for (size_t i = 0; i < m; ++i) nn += q(i, k) * q(i, k);
    nn = std::sqrt(nn);
    for (size_t i = 0; i < m; ++i) q(i, k) /= nn;
    for (size_t j = 0; j < n; ++j) {
      if (used[j]) continue;
      double dot = 0.;
      for (size_t i = 0; i < m; ++i) dot += q(i, k) * w(i, j);
      for (size_t i = 0; i < m; ++i) w(i, j) -= dot * q(i, k);
    }
  }
  q.c = k;
  q.a.resize(m * k);
  return q;
}

inline Mat sym_fun(const Mat& v, const std::vector<double>& w, double (*f)(double)) {  // V diag(f(w)) V^T
  const size_t n = v.r;
  Mat o(n, n);
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j) {
      double s = 0.;
      for (size_t k = 0; k < n; ++k) s += v(i, k) * f(w[k]) * v(j, k);
      o(i, j) = s;
    }
  return o;
}
inline double f_sqrt(double x) { return std::sqrt(x); }
inline double f_isqrt(double x) { return 1. / std::sqrt(x); }

// adapt/low_rank.rs:241-268
inline bool spd_mean(const Mat& cov_draws, const Mat& cov_grads, Mat& out) {
  std::vector<double> w, wm;
  Mat u, um;
  if (!eigh(cov_grads, w, u)) return false;
  const Mat gs = sym_fun(u, w, f_sqrt);
  const Mat m = matmul(matmul(gs, cov_draws), gs);
  if (!eigh(m, wm, um)) return false;
  const Mat ms = sym_fun(um, wm, f_sqrt);
  const Mat gi = sym_fun(u, w, f_isqrt);
  out = matmul(matmul(gi, ms), gi);
  return true;
}

// adapt/low_rank.rs:210-239: draws, grads [k x n]
inline bool estimate_mass_matrix(const Mat& draws, const Mat& grads, double gamma, std::vector<double>& vals, Mat& vecs) {
  Mat cd = matmul(draws, draws, false, true), cg = matmul(grads, grads, false, true);
  for (double& x : cd.a) x *= 1. / gamma;
  for (double& x : cg.a) x *= 1. / gamma;
  for (size_t i = 0; i < cd.r; ++i) cd(i, i) += 1., cg(i, i) += 1.;
  Mat mean;
  if (!spd_mean(cd, cg, mean)) return false;
  return eigh(mean, vals, vecs);
}

struct Update {
  std::vector<double> stds, mean, vals, vecs /* r eigenvectors of length d, one after the other */, mu;
};

// adapt/low_rank.rs:73-131 (+ rescale_points :150-208).  draws, grads: n rows of length d (oldest first)
inline bool compute_update(const double* draws_in, const double* grads_in, size_t n, size_t d, double gamma, double cutoff, Update& out) {
  Mat draws(d, n), grads(d, n);
  for (size_t j = 0; j < n; ++j)
    for (size_t i = 0; i < d; ++i) draws(i, j) = draws_in[j * d + i], grads(i, j) = grads_in[j * d + i];
  out.stds.assign(d, 0.), out.mean.assign(d, 0.);
  std::vector<double> dm(d), gm(d);
  const double nn = (double)n;
  for (size_t row = 0; row < d; ++row) {
    double s1 = 0., s2 = 0.;
    for (size_t j = 0; j < n; ++j) s1 += draws(row, j), s2 += grads(row, j);
    const double draw_mean = s1 / nn, grad_mean = s2 / nn;
    double v1 = 0., v2 = 0.;
    for (size_t j = 0; j < n; ++j) {
      v1 += (draws(row, j) - draw_mean) * (draws(row, j) - draw_mean);
      v2 += (grads(row, j) - grad_mean) * (grads(row, j) - grad_mean);
    }
    const double sigma = std::sqrt(std::sqrt((v1 / nn) / (v2 / nn)));
    out.mean[row] = draw_mean + sigma * sigma * grad_mean;
    out.stds[row] = sigma;
    const double ds = 1. / sigma;
    s1 = s2 = 0.;
    for (size_t j = 0; j < n; ++j) {
      draws(row, j) = (draws(row, j) - out.mean[row]) * ds;
      grads(row, j) = grads(row, j) * sigma;
      s1 += draws(row, j), s2 += grads(row, j);
    }
    dm[row] = s1 / nn, gm[row] = s2 / nn;
    for (size_t j = 0; j < n; ++j) draws(row, j) -= dm[row], grads(row, j) -= gm[row];
  }
  for (double x : draws.a)
    if (!std::isfinite(x)) return false;
  for (double x : grads.a)
    if (!std::isfinite(x)) return false;
  Mat ud, ug;
  if (!thin_svd_u(draws, ud) || !thin_svd_u(grads, ug)) return false;
  Mat sub(d, ud.c + ug.c);
  std::copy(ud.a.begin(), ud.a.end(), sub.a.begin());
  std::copy(ug.a.begin(), ug.a.end(), sub.a.begin() + ud.a.size());
  const Mat basis = orth_basis(sub);
  const Mat dp = matmul(basis, draws, true, false), gp = matmul(basis, grads, true, false);
  std::vector<double> vals;
  Mat vecs;
  if (!estimate_mass_matrix(dp, gp, gamma, vals, vecs)) return false;
  out.vals.clear(), out.vecs.clear();
  for (size_t k = 0; k < vals.size(); ++k) {
    if (!((vals[k] > cutoff) | (vals[k] < 1. / cutoff))) continue;
    out.vals.push_back(vals[k]);
    for (size_t i = 0; i < d; ++i) {
      double s = 0.;
      for (size_t l = 0; l < basis.c; ++l) s += basis(i, l) * vecs(l, k);
      out.vecs.push_back(s);
    }
  }
  const size_t r = out.vals.size();
  out.mu.assign(d, 0.);
  std::vector<double> b(r, 0.);
  for (size_t k = 0; k < r; ++k) {
    for (size_t i = 0; i < d; ++i) b[k] += out.vecs[k * d + i] * gm[i];
    b[k] *= out.vals[k] - 1.;
  }
  for (size_t i = 0; i < d; ++i) {
    double s = 0.;
    for (size_t k = 0; k < r; ++k) s += out.vecs[k * d + i] * b[k];
    out.mu[i] = dm[i] + gm[i] + s;
  }
  return true;
}

}  // namespace lowrank
}  // namespace oracle
