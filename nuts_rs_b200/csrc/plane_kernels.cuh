// plane_kernels.cuh — Tier 1 (batched Math-trait ops) and Tier 2 (fused Hamiltonian ops) kernels over
// [nchains x dim] f64 planes.  One CTA of PK_THREADS threads per chain row, strided over dim (any dim);
// these kernels stream their operands from HBM once (coalesced 8-byte loads, 256 B per warp instruction).
// Reference semantics: src/math/math.rs (trait), src/math/cpu_math.rs + src/math/util.rs (CPU implementation).
#pragma once
#include "device_common.cuh"

namespace nb {

constexpr int PK_THREADS = 256;

struct RowArgs {
  int N, d, ld;
};

#define NB_ROW_PROLOGUE                                       \
  const int c = blockIdx.x;                                   \
  if (c >= A.N) return;                                       \
  if (active && !active[c]) return;                           \
  const size_t row = (size_t)c * A.ld;                        \
  (void)row;

// ---------------------------------------------------------------- elementwise
__global__ void __launch_bounds__(PK_THREADS) k_axpy(RowArgs A, const double* __restrict__ x, double* __restrict__ y,
                                                      const double* __restrict__ a, double a_bcast, const uint8_t* active) {
  NB_ROW_PROLOGUE
  const double av = a ? a[c] : a_bcast;
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) y[row + i] = fma(av, x[row + i], y[row + i]);  // util.rs:426-437
}
__global__ void __launch_bounds__(PK_THREADS) k_axpy_out(RowArgs A, const double* __restrict__ x, const double* __restrict__ y,
                                                          const double* __restrict__ a, double a_bcast, double* __restrict__ out,
                                                          const uint8_t* active) {
  NB_ROW_PROLOGUE
  const double av = a ? a[c] : a_bcast;
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) out[row + i] = fma(av, x[row + i], y[row + i]);  // util.rs:472-495
}
__global__ void __launch_bounds__(PK_THREADS) k_mult(RowArgs A, const double* x, const double* y, double* out) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) out[row + i] = x[row + i] * y[row + i];  // util.rs:46-49 (plain product)
}
__global__ void __launch_bounds__(PK_THREADS) k_recip(RowArgs A, const double* x, double* out) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) out[row + i] = 1.0 / x[row + i];  // cpu_math.rs:328-330
}
__global__ void __launch_bounds__(PK_THREADS) k_fill(RowArgs A, double* x, double val) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) x[row + i] = val;
}
__global__ void __launch_bounds__(PK_THREADS) k_copy(RowArgs A, const double* src, double* dst) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) dst[row + i] = src[row + i];
}
// host [N*d] <-> plane [N][ld]
__global__ void __launch_bounds__(PK_THREADS) k_pack(RowArgs A, const double* dense, double* plane) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) plane[row + i] = dense[(size_t)c * A.d + i];
}
__global__ void __launch_bounds__(PK_THREADS) k_unpack(RowArgs A, const double* plane, double* dense) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) dense[(size_t)c * A.d + i] = plane[row + i];
}
__global__ void __launch_bounds__(PK_THREADS) k_gaussian(RowArgs A, double* dest, const double* stds, uint64_t seed,
                                                          uint64_t chain_offset, uint64_t counter) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  const uint64_t stream = chain_offset + (uint64_t)c + 1;
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS)
    dest[row + i] = stds[row + i] * stream_normal(seed, stream, counter, (uint32_t)i);  // cpu_math.rs:561-577
}
__global__ void __launch_bounds__(PK_THREADS) k_update_variance(RowArgs A, double* mean, double* var, const double* value,
                                                                 const double* scale, double scale_bcast) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  const double sc = scale ? scale[c] : scale_bcast;
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {  // cpu_math.rs:625-629
    double diff = value[row + i] - mean[row + i];
    mean[row + i] += diff * sc;
    var[row + i] += diff * diff;
  }
}
// mode 0: draw (cpu_math.rs:633-669), 1: draw_grad (:671-708), 2: grad (:710-738)
__global__ void __launch_bounds__(PK_THREADS) k_update_var_inv_std(RowArgs A, int mode, double* inv_std, double* std_, const double* a,
                                                                    const double* b, double scale, int has_fill, double fill,
                                                                    double lo, double hi) {
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    if (mode == 2) {
      double val = 1.0 / clampd(fabs(a[row + i]), lo, hi);
      if (!isfinite(val)) val = fill;
      std_[row + i] = sqrt(val);
      inv_std[row + i] = sqrt(1.0 / val);
    } else {
      double val = mode == 0 ? a[row + i] * scale : sqrt(a[row + i] / b[row + i]);
      if ((!isfinite(val)) | (val == 0.0)) {
        if (has_fill) {
          std_[row + i] = sqrt(fill);
          inv_std[row + i] = sqrt(1.0 / fill);
        }
      } else {
        val = clampd(val, lo, hi);
        std_[row + i] = sqrt(val);
        inv_std[row + i] = sqrt(1.0 / val);
      }
    }
  }
}

// ---------------------------------------------------------------- per-chain reductions -> out[N] (device)
// op 0: dot(x,y)  1: sq_norm_sum (x+y)^2  2: sum ln(x)  3: all finite  4: all finite and nonzero
__global__ void __launch_bounds__(PK_THREADS) k_reduce1(RowArgs A, int op, const double* x, const double* y, double* out) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  double s[1] = {0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    double xv = x[row + i];
    if (op == 0) s[0] = fma(xv, y[row + i], s[0]);
    else if (op == 1) s[0] += (xv + y[row + i]) * (xv + y[row + i]);
    else if (op == 2) s[0] += log(xv);
    else if (op == 3) s[0] += isfinite(xv) ? 0.0 : 1.0;
    else s[0] += (isfinite(xv) & (xv != 0.0)) ? 0.0 : 1.0;
  }
  red.allreduce(s);
  if (threadIdx.x == 0) out[c] = (op >= 3) ? (s[0] == 0.0 ? 1.0 : 0.0) : s[0];
}
// scalar_prods3 (util.rs:221-347): ((p1 + p2) - n1).x and .y ;  n1 == nullptr gives scalar_prods2 (util.rs:114-219)
__global__ void __launch_bounds__(PK_THREADS) k_scalar_prods(RowArgs A, const double* p1, const double* n1, const double* p2,
                                                              const double* x, const double* y, double* out1, double* out2) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  double s[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    double sum = p1[row + i] + p2[row + i];
    if (n1) sum = sum - n1[row + i];
    s[0] = fma(sum, x[row + i], s[0]);
    s[1] = fma(sum, y[row + i], s[1]);
  }
  red.allreduce(s);
  if (threadIdx.x == 0) {
    out1[c] = s[0];
    out2[c] = s[1];
  }
}

// ---------------------------------------------------------------- log density over planes (Math::logp_array)
struct PointDev {  // batched TransformedPoint (transformed_hamiltonian.rs:56-77)
  double *x, *gx, *z, *gz, *v;
  long long* idx;
  double *logp, *logdet, *ke, *e0;
  long long* tid;
};
struct TransformDev {  // batched DiagMassMatrix (transform/diagonal.rs:9-17) + the optional low-rank correction of
                       // LowRankMassMatrix (transform/low_rank.rs:25-41, 97-111): null / rank 0 = pure diagonal
  double *stds, *inv_stds, *mean;
  double* logdet;
  long long* id;
  const double* lr_vecs;           // [N][lr_rmax][ld]: eigenvector k of chain c at ((c * lr_rmax + k) * ld)
  const double* lr_vals_sqrt;      // [N][lr_rmax] lambda^{1/2}
  const double* lr_vals_sqrt_inv;  // [N][lr_rmax] lambda^{-1/2}
  const double* lr_mu;             // [N][ld]
  const int* lr_rank;              // [N] number of eigenvectors of chain c (-1: the transformation of chain c is diagonal)
  int lr_rmax;
};
constexpr int LR_MAX_RANK = 64;  // eigenvectors per chain the kernels hold coefficients for

// Math::apply_lowrank_transform(_inplace) (math.rs:131-144, cpu_math.rs:332-425) for the row of one chain:
//   out = in + U ((vals - 1) .* (U^T in)),  U = r eigenvectors of length d (rows of `vecs`, stride ld).  in == out is allowed.
// Two passes over U (2 * r * d * 8 bytes from HBM / L2 per chain): the r dot products 8 at a time through the team reduction,
// then the rank-r update with the coefficients in shared memory.
__device__ __forceinline__ void lowrank_apply_row(const double* __restrict__ vecs, const double* __restrict__ vals, int r, int d, int ld,
                                                  const double* in, double* out, TeamReduce<PK_THREADS>& red, double* coef) {
  for (int k0 = 0; k0 < r; k0 += 8) {
    double part[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      const double xi = in[i];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (k0 + q < r) part[q] = fma(vecs[(size_t)(k0 + q) * ld + i], xi, part[q]);
    }
    red.allreduce(part);
    if (threadIdx.x < 8 && k0 + (int)threadIdx.x < r) coef[k0 + threadIdx.x] = part[threadIdx.x] * (vals[k0 + threadIdx.x] - 1.0);
  }
  __syncthreads();  // coefficients visible; every thread has finished reading `in` (in == out)
  for (int i = threadIdx.x; i < d; i += PK_THREADS) {
    double acc = in[i];
    for (int k = 0; k < r; ++k) acc = fma(vecs[(size_t)k * ld + i], coef[k], acc);
    out[i] = acc;
  }
  __syncthreads();  // `out` complete before the caller's next pass reads other elements of it; coef free for the next call
}
// -1: the chain's transformation has no inner matrix (pure diagonal); >= 0: LowRankMassMatrix::update was called with that many
// eigenvectors - with 0 of them the translation mu_lr still applies (low_rank.rs:337-345: `if let Some(inner)`)
__device__ __forceinline__ int lowrank_rank(const TransformDev& T, int c) {
  if (!T.lr_rank) return -1;
  const int r = T.lr_rank[c];
  return r < T.lr_rmax ? r : T.lr_rmax;
}
// z, gz of a chain from x, gx (LowRankMassMatrix::compute_transformed_position / _gradient, low_rank.rs:326-348, 380-398; with
// rank 0 DiagMassMatrix's, diagonal.rs:233-265)
__device__ __forceinline__ void whiten_row(const TransformDev& T, int c, size_t row, int d, int ld, const double* x, const double* gx, double* z,
                                           double* gz, TeamReduce<PK_THREADS>& red, double* coef) {
  const int r = lowrank_rank(T, c);
  for (int i = threadIdx.x; i < d; i += PK_THREADS) {
    double t = fma(-1.0, T.mean[row + i], x[i]);
    double zv = T.inv_stds[row + i] * t;
    if (r >= 0) zv = fma(-1.0, T.lr_mu[row + i], zv);  // axpy(mu, z, -1)
    z[i] = zv;
    gz[i] = gx[i] * T.stds[row + i];
  }
  if (r > 0) {
    __syncthreads();
    const double* U = T.lr_vecs + (size_t)c * T.lr_rmax * ld;
    lowrank_apply_row(U, T.lr_vals_sqrt_inv + (size_t)c * T.lr_rmax, r, d, ld, z, z, red, coef);
    lowrank_apply_row(U, T.lr_vals_sqrt + (size_t)c * T.lr_rmax, r, d, ld, gz, gz, red, coef);
  }
}

// strided model evaluation at x (row pointer): writes gx, returns logp to every thread
__device__ __forceinline__ double model_eval_row(const ModelDev& m, int d, const double* x, double* gx, TeamReduce<PK_THREADS>& red) {
  double a0 = 0.0, a1 = 0.0;
  if (m.kind == LOGP_GAUSS_RANK1) {
    double s[1] = {0.0};
    for (int i = threadIdx.x; i < d; i += PK_THREADS) s[0] += x[i] - m.mu[i];
    red.allreduce(s);
    a0 = m.rank1_coeff * s[0];
  } else if (m.kind == LOGP_FUNNEL) {
    double s[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      if (i == 0) s[1] = x[i];
      else s[0] = fma(x[i], x[i], s[0]);
    }
    red.allreduce(s);
    a0 = s[1];
    a1 = s[0];
  }
  if (m.kind == LOGP_USER) {  // the user density (include/nuts_user_logp.cuh); an error of either kind makes logp NaN
    double s[2] = {0.0, 0.0};
    if (NutsUserLogp::NUM_SUMS > 0) {
      for (int i = threadIdx.x; i < d; i += PK_THREADS) NutsUserLogp::sums(i, d, x[i], m.user, s);
      red.allreduce(s);
    }
    double p2[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      int status = NUTS_USER_OK;
      double g = 0.0;
      p2[0] += NutsUserLogp::element(i, d, x[i], s, m.user, g, status);
      gx[i] = g;
      if (status != NUTS_USER_OK) p2[1] = 1.0;
    }
    red.allreduce(p2);
    return p2[1] > 0.0 ? __longlong_as_double(0x7ff8000000000000ll) : p2[0] + NutsUserLogp::finish(d, s, m.user);
  }
  double lp[1] = {0.0};
  if (m.kind == LOGP_FUNNEL) {
    double ev = exp(-a0), nm1 = (double)(d - 1), half_ev_S = 0.5 * ev * a1;
    for (int i = threadIdx.x; i < d; i += PK_THREADS)
      gx[i] = i == 0 ? (-a0 * m.funnel_inv_var - 0.5 * nm1 + half_ev_S) : (-x[i] * ev);
    return -0.5 * a0 * a0 * m.funnel_inv_var - 0.5 * nm1 * a0 - half_ev_S;
  }
  for (int i = threadIdx.x; i < d; i += PK_THREADS) {
    double diff = x[i] - m.mu[i];
    if (m.kind == LOGP_GAUSS_ISO) {
      lp[0] -= diff * diff / 2.;
      gx[i] = -diff;
    } else if (m.kind == LOGP_GAUSS_DIAG) {
      double pd = diff * m.prec[i];
      lp[0] -= diff * pd / 2.;
      gx[i] = -pd;
    } else {
      double ptd = diff - a0;
      gx[i] = -ptd;
      lp[0] -= 0.5 * diff * ptd;
    }
  }
  red.allreduce(lp);
  return lp[0];
}

__global__ void __launch_bounds__(PK_THREADS) k_logp_array(RowArgs A, ModelDev m, const double* x, double* gx, double* logp, int* status) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  double lp = model_eval_row(m, A.d, x + row, gx + row, red);
  double bad[1] = {0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS)
    if (!isfinite(gx[row + i])) bad[0] = 1.0;
  red.allreduce(bad);
  if (threadIdx.x == 0) {
    logp[c] = lp;
    if (status) status[c] = (bad[0] != 0.0 || !isfinite(lp)) ? 2 : 0;
  }
}

// DiagMassMatrix::set_transform (diagonal.rs:156-162): stds/mean already written; inv_stds, logdet, id
__global__ void __launch_bounds__(PK_THREADS) k_set_transform(RowArgs A, TransformDev T, const uint8_t* active) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  NB_ROW_PROLOGUE
  double s[1] = {0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    double is = 1.0 / T.stds[row + i];
    T.inv_stds[row + i] = is;
    s[0] += log(is);
  }
  red.allreduce(s);
  if (threadIdx.x == 0) {
    T.logdet[c] = s[0];
    T.id[c] += 1;
  }
}

// Hamiltonian::init_state (transformed_hamiltonian.rs:640-661): x plane given
__global__ void __launch_bounds__(PK_THREADS) k_init_state(RowArgs A, ModelDev m, TransformDev T, PointDev p, int* status) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  __shared__ double coef[LR_MAX_RANK];
  double lp = model_eval_row(m, A.d, p.x + row, p.gx + row, red);
  double bad[1] = {0.0};
  __syncthreads();  // gx of the whole row written
  whiten_row(T, c, row, A.d, A.ld, p.x + row, p.gx + row, p.z + row, p.gz + row, red, coef);
  __syncthreads();
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    double xv = p.x[row + i], gxv = p.gx[row + i], zv = p.z[row + i], gv = p.gz[row + i];
    if (!(isfinite(zv) && isfinite(gv) && (gv != 0.0) && isfinite(gxv) && isfinite(xv))) bad[0] = 1.0;
  }
  red.allreduce(bad);
  if (threadIdx.x == 0) {
    p.logp[c] = lp;
    p.logdet[c] = T.logdet[c];
    p.tid[c] = T.id[c];
    if (status) status[c] = bad[0] != 0.0 ? 3 : 0;
  }
}

// Hamiltonian::initialize_trajectory (transformed_hamiltonian.rs:687-736)
// kind: KineticEnergyKind; Microcanonical (2) puts a resampled momentum on the unit sphere (:699-702) and starts the accumulated
// kinetic-energy change at 0 (:720-729)
__global__ void __launch_bounds__(PK_THREADS) k_initialize_trajectory(RowArgs A, TransformDev T, PointDev p, int resample, uint64_t seed,
                                                                       uint64_t chain_offset, uint64_t counter, int kind) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  const uint64_t stream = chain_offset + (uint64_t)c + 1;
  __shared__ double coef[LR_MAX_RANK];
  const bool rewhiten = T.id[c] != p.tid[c];
  double s[1] = {0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    double vv = resample ? 1.0 * stream_normal(seed, stream, counter, (uint32_t)i) : p.v[row + i];
    if (resample) p.v[row + i] = vv;
    s[0] = fma(vv, vv, s[0]);
  }
  if (rewhiten) whiten_row(T, c, row, A.d, A.ld, p.x + row, p.gx + row, p.z + row, p.gz + row, red, coef);  // inv_transform_normalize
  red.allreduce(s);
  if (threadIdx.x == 0) {
    if (rewhiten) {
      p.logdet[c] = T.logdet[c];
      p.tid[c] = T.id[c];
    }
    double ke = kind == 2 ? 0.0 : 0.5 * s[0];
    p.ke[c] = ke;
    p.idx[c] = 0;
    p.e0[c] = ke - (p.logp[c] + p.logdet[c]);
  }
  if (kind == 2 && resample) {  // Math::array_normalize (cpu_math.rs:496-503); every thread re-reads the elements it wrote
    const double inv = 1.0 / sqrt(s[0]);
    for (int i = threadIdx.x; i < A.d; i += PK_THREADS) p.v[row + i] *= inv;
  }
}

// Hamiltonian::leapfrog (transformed_hamiltonian.rs:524-615), Euclidean, all chains in one launch.
// HBM traffic per chain: reads z, v, gz, stds, mean (40*d B) ; writes z', v', x', gx', gz' (40*d B).
// (LRK = false: the low-rank branch is compiled out - 64 registers, 4 CTAs per SM; the host launches LRK = true only while some
// chain carries a low-rank correction)
template <bool LRK>
__global__ void __launch_bounds__(PK_THREADS) k_leapfrog(RowArgs A, ModelDev m, TransformDev T, PointDev s, PointDev o,
                                                          const double* step_size, double step_bcast, const int8_t* dir,
                                                          const double* baseline, double max_energy_error, const uint8_t* active,
                                                          int* status, double* energy_error_out) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  NB_ROW_PROLOGUE
  const int d = A.d;
  const int sign = dir ? (int)dir[c] : 1;
  const double eps = (double)sign * (step_size ? step_size[c] : step_bcast) * 1.0;
  const double eps_half = eps / 2.;
  double part[2] = {0.0, 0.0};
  double lp;
  __shared__ double coef[LRK ? LR_MAX_RANK : 1];
  const int lr = LRK ? lowrank_rank(T, c) : -1;
  const bool single_pass = ((m.kind == LOGP_GAUSS_ISO) | (m.kind == LOGP_GAUSS_DIAG)) && lr < 0;
  if (single_pass) {
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      double sg = T.stds[row + i];
      double vh = fma(eps_half, s.gz[row + i], s.v[row + i]);
      double zn = fma(eps, vh, s.z[row + i]);
      double t = zn * sg;
      double xn = fma(1.0, T.mean[row + i], t);
      double diff = xn - m.mu[i];
      double gxn;
      if (m.kind == LOGP_GAUSS_ISO) {
        part[0] -= diff * diff / 2.;
        gxn = -diff;
      } else {
        double pd = diff * m.prec[i];
        part[0] -= diff * pd / 2.;
        gxn = -pd;
      }
      double gn = gxn * sg;
      double vn = fma(eps_half, gn, vh);
      part[1] = fma(vn, vn, part[1]);
      o.z[row + i] = zn;
      o.x[row + i] = xn;
      o.gx[row + i] = gxn;
      o.gz[row + i] = gn;
      o.v[row + i] = vn;
    }
    red.allreduce(part);
    lp = part[0];
  } else if (lr < 0) {
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      double vh = fma(eps_half, s.gz[row + i], s.v[row + i]);
      double zn = fma(eps, vh, s.z[row + i]);
      double t = zn * T.stds[row + i];
      o.v[row + i] = vh;
      o.z[row + i] = zn;
      o.x[row + i] = fma(1.0, T.mean[row + i], t);
    }
    lp = model_eval_row(m, d, o.x + row, o.gx + row, red);
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      double gn = o.gx[row + i] * T.stds[row + i];
      double vn = fma(eps_half, gn, o.v[row + i]);
      o.gz[row + i] = gn;
      o.v[row + i] = vn;
      part[1] = fma(vn, vn, part[1]);
    }
    red.allreduce(part);
  } else {
    // low-rank transformation (low_rank.rs:350-398): x = ((I + U (sqrt(lambda) - 1) U^T) z + mu_lr) * sigma + mean,
    // grad_z = (I + U (sqrt(lambda) - 1) U^T) (grad_x * sigma): two more passes with a team reduction each
    const double* U = T.lr_vecs + (size_t)c * T.lr_rmax * A.ld;
    const double* lam = T.lr_vals_sqrt + (size_t)c * T.lr_rmax;
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      double vh = fma(eps_half, s.gz[row + i], s.v[row + i]);
      o.v[row + i] = vh;
      o.z[row + i] = fma(eps, vh, s.z[row + i]);
    }
    __syncthreads();
    lowrank_apply_row(U, lam, lr, d, A.ld, o.z + row, o.x + row, red, coef);
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      double xv = fma(1.0, T.lr_mu[row + i], o.x[row + i]);  // axpy(mu, x, 1)
      xv = xv * T.stds[row + i];                             // array_mult_inplace(x, stds)
      o.x[row + i] = fma(1.0, T.mean[row + i], xv);          // axpy(mean, x, 1)
    }
    __syncthreads();
    lp = model_eval_row(m, d, o.x + row, o.gx + row, red);
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += PK_THREADS) o.gz[row + i] = o.gx[row + i] * T.stds[row + i];
    __syncthreads();
    lowrank_apply_row(U, lam, lr, d, A.ld, o.gz + row, o.gz + row, red, coef);
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      double vn = fma(eps_half, o.gz[row + i], o.v[row + i]);
      o.v[row + i] = vn;
      part[1] = fma(vn, vn, part[1]);
    }
    red.allreduce(part);
  }
  if (threadIdx.x == 0) {
    double ke = 0.5 * part[1];
    double logdet = T.logdet[c];
    o.logp[c] = lp;
    o.logdet[c] = logdet;
    o.ke[c] = ke;
    o.e0[c] = s.e0[c];
    o.tid[c] = T.id[c];
    o.idx[c] = s.idx[c] + sign;
    double base = baseline ? baseline[c] : s.e0[c];
    double ee = (ke - (lp + logdet)) - base;
    if (energy_error_out) energy_error_out[c] = ee;
    if (status) status[c] = ((ee > max_energy_error) | !isfinite(ee)) ? 1 : 0;
  }
}

// ---------------------------------------------------------------- TMA bulk copies (cp.async.bulk + mbarrier; SASS UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completion is counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spin = 0; !ok; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (spin > (1 << 22)) __trap();  // a byte count that never completes must not hang the GPU
  }
}

// Hamiltonian::leapfrog for the elementwise targets (isotropic / diagonal Gaussian, diagonal transformation) with the five input rows
// (z, v, grad_z, sigma, mean) staged through shared memory by the TMA: one thread arms an mbarrier per stage and issues five
// cp.async.bulk copies of a TMA_CH-element chunk each (4 KB per vector, 20 KB per stage, two stages: a d <= 1024 row is completely in
// flight before the first wait - 40 KB per CTA, 160 KB per SM at 4 resident CTAs, against ~40 KB for k_leapfrog's register loads); the
// 256 threads wait on the stage, compute from shared memory and store the five output rows straight from registers.  Element ->
// thread mapping and the order of every thread's partial sums are those of k_leapfrog's single-pass branch, so the results are
// bit-identical to it (tests/test_gpu_primitives.py::test_leapfrog_tma_matches_register_path).
// HBM traffic per chain: 40*d B in, 40*d B out (the model's mu / precision vectors are shared by all chains and stay in L2).
constexpr int TMA_CH = 512;
constexpr int TMA_NST = 2;
__global__ void __launch_bounds__(PK_THREADS) k_leapfrog_tma(RowArgs A, ModelDev m, TransformDev T, PointDev s, PointDev o,
                                                              const double* step_size, double step_bcast, const int8_t* dir,
                                                              const double* baseline, double max_energy_error, const uint8_t* active,
                                                              int* status, double* energy_error_out) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  __shared__ alignas(128) double buf[TMA_NST][5][TMA_CH];
  __shared__ alignas(8) uint64_t full[TMA_NST];
  TeamReduce<PK_THREADS> red(scratch);
  NB_ROW_PROLOGUE
  const int d = A.d;
  const int nch = (d + TMA_CH - 1) / TMA_CH;
  const int sign = dir ? (int)dir[c] : 1;
  const double eps = (double)sign * (step_size ? step_size[c] : step_bcast) * 1.0;
  const double eps_half = eps / 2.;
  const bool diag = m.kind == LOGP_GAUSS_DIAG;
  const double* src[5] = {s.z + row, s.v + row, s.gz + row, T.stds + row, T.mean + row};
  auto issue = [&](int k) {  // one thread: chunk k into stage k % TMA_NST
    const int st = k % TMA_NST, c0 = k * TMA_CH;
    const int len = min(TMA_CH, d - c0);
    const uint32_t bytes = (uint32_t)((len + 1) & ~1) * 8u;  // rows are padded to a multiple of 16 doubles: the odd tail reads its neighbour
    mbar_expect_tx(&full[st], 5u * bytes);
#pragma unroll
    for (int q = 0; q < 5; ++q) bulk_g2s(&buf[st][q][0], src[q] + c0, bytes, &full[st]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int st = 0; st < TMA_NST; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
    for (int k = 0; k < nch && k < TMA_NST; ++k) issue(k);
  }
  __syncthreads();  // the barriers are initialised before anybody waits on them
  double part[2] = {0.0, 0.0};
  for (int k = 0; k < nch; ++k) {
    const int st = k % TMA_NST, c0 = k * TMA_CH;
    const int len = min(TMA_CH, d - c0);
    // this thread's model parameters first: L2 hits that are in flight while the bulk copy lands
    double mu_[TMA_CH / PK_THREADS], pr_[TMA_CH / PK_THREADS];
#pragma unroll
    for (int e = 0; e < TMA_CH / PK_THREADS; ++e) {
      const int li = threadIdx.x + e * PK_THREADS;
      mu_[e] = li < len ? m.mu[c0 + li] : 0.0;
      pr_[e] = (diag && li < len) ? m.prec[c0 + li] : 0.0;
    }
    mbar_wait(&full[st], (uint32_t)((k / TMA_NST) & 1));
#pragma unroll
    for (int e = 0; e < TMA_CH / PK_THREADS; ++e) {
      const int li = threadIdx.x + e * PK_THREADS;
      if (li < len) {
        const size_t gi = row + c0 + li;
        const double sg = buf[st][3][li];
        const double vh = fma(eps_half, buf[st][2][li], buf[st][1][li]);
        const double zn = fma(eps, vh, buf[st][0][li]);
        const double t = zn * sg;
        const double xn = fma(1.0, buf[st][4][li], t);
        const double diff = xn - mu_[e];
        double gxn;
        if (!diag) {
          part[0] -= diff * diff / 2.;
          gxn = -diff;
        } else {
          const double pd = diff * pr_[e];
          part[0] -= diff * pd / 2.;
          gxn = -pd;
        }
        const double gn = gxn * sg;
        const double vn = fma(eps_half, gn, vh);
        part[1] = fma(vn, vn, part[1]);
        o.z[gi] = zn;
        o.x[gi] = xn;
        o.gx[gi] = gxn;
        o.gz[gi] = gn;
        o.v[gi] = vn;
      }
    }
    if (k + TMA_NST < nch) {  // refill the stage once every thread has read it
      __syncthreads();
      if (threadIdx.x == 0) issue(k + TMA_NST);
    }
  }
  red.allreduce(part);
  if (threadIdx.x == 0) {
    const double lp = part[0];
    const double ke = 0.5 * part[1];
    const double logdet = T.logdet[c];
    o.logp[c] = lp;
    o.logdet[c] = logdet;
    o.ke[c] = ke;
    o.e0[c] = s.e0[c];
    o.tid[c] = T.id[c];
    o.idx[c] = s.idx[c] + sign;
    const double base = baseline ? baseline[c] : s.e0[c];
    const double ee = (ke - (lp + logdet)) - base;
    if (energy_error_out) energy_error_out[c] = ee;
    if (status) status[c] = ((ee > max_energy_error) | !isfinite(ee)) ? 1 : 0;
  }
}

// Math::apply_lowrank_transform / _inplace for every chain (Tier 1)
__global__ void __launch_bounds__(PK_THREADS) k_lowrank_apply(RowArgs A, const double* vecs, const double* vals, const int* rank, int rmax,
                                                               const double* in, double* out) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  __shared__ double coef[LR_MAX_RANK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  const int r = rank[c] < rmax ? rank[c] : rmax;
  if (r == 0) {  // cpu_math.rs:339-342: no eigenvectors = copy
    if (in != out)
      for (int i = threadIdx.x; i < A.d; i += PK_THREADS) out[row + i] = in[row + i];
    return;
  }
  lowrank_apply_row(vecs + (size_t)c * rmax * A.ld, vals + (size_t)c * rmax, r, A.d, A.ld, in + row, out + row, red, coef);
}
// LowRankMassMatrix::update (low_rank.rs:186-189): logdet = inner.logdet() + diag.logdet() for the chains whose update was accepted
__global__ void k_add_logdet(int N, double* logdet, const double* contribution) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < N) logdet[c] = contribution[c] + logdet[c];
}

// Hamiltonian::is_turning (transformed_hamiltonian.rs:617-638)
__global__ void __launch_bounds__(PK_THREADS) k_is_turning(RowArgs A, PointDev s1, PointDev s2, uint8_t* turning) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  const uint8_t* active = nullptr;
  NB_ROW_PROLOGUE
  const bool first_is_start = s1.idx[c] < s2.idx[c];
  const PointDev& st = first_is_start ? s1 : s2;
  const PointDev& en = first_is_start ? s2 : s1;
  double s[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    double sum = (en.z[row + i] + 0.0) - st.z[row + i];
    s[0] = fma(sum, st.v[row + i], s[0]);
    s[1] = fma(sum, en.v[row + i], s[1]);
  }
  red.allreduce(s);
  if (threadIdx.x == 0) turning[c] = ((s[0] < 0.) | (s[1] < 0.)) ? 1 : 0;
}

// ---------------------------------------------------------------- KineticEnergyKind::ExactNormal / Microcanonical (Tier 1 + Tier 2)
// The reference's SIMD kernels fuse the multiply-add in whole registers of 4 lanes and leave the last d % 4 elements to an unfused scalar
// tail (util.rs:541-572, 625-647); `body` = (d / 4) * 4 keeps both roundings so the planes match the CPU bit for bit.
// Math::std_norm_flow (math.rs:155-161, util.rs:507-590): sin / cos of epsilon are evaluated ON THE HOST (f64::sin / cos, like the reference)
__global__ void __launch_bounds__(PK_THREADS) k_std_norm_flow(RowArgs A, const double* __restrict__ pos, double* __restrict__ pos_out,
                                                               double* __restrict__ vel, const double* __restrict__ sn,
                                                               const double* __restrict__ cs, double sn_bcast, double cs_bcast,
                                                               const uint8_t* active) {
  NB_ROW_PROLOGUE
  const double es = sn ? sn[c] : sn_bcast, ec = cs ? cs[c] : cs_bcast;
  const int body = (A.d / 4) * 4;
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    const double p = pos[row + i], v = vel[row + i];
    double po, vo;
    if (i < body) {
      po = fma(p, ec, v * es);
      vo = fma(p, -es, v * ec);
    } else {
      po = p * ec + v * es;
      vo = p * (-es) + v * ec;
    }
    pos_out[row + i] = po;
    vel[row + i] = vo;
  }
}
// Math::std_norm_grad_flow / _inplace (math.rs:162-176, util.rs:592-742): vel_out = vel + eps * (pos + grad); vel_out may alias vel
__global__ void __launch_bounds__(PK_THREADS) k_std_norm_grad_flow(RowArgs A, const double* pos, const double* grad, const double* vel,
                                                                    double* vel_out, const double* eps, double eps_bcast,
                                                                    const uint8_t* active) {
  NB_ROW_PROLOGUE
  const double e = eps ? eps[c] : eps_bcast;
  const int body = (A.d / 4) * 4;
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) {
    const double pg = pos[row + i] + grad[row + i], v = vel[row + i];
    vel_out[row + i] = i < body ? fma(e, pg, v) : v + e * pg;
  }
}
// Math::array_normalize (math.rs:178-181, cpu_math.rs:496-503)
__global__ void __launch_bounds__(PK_THREADS) k_normalize(RowArgs A, double* v, const uint8_t* active) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  NB_ROW_PROLOGUE
  double s[1] = {0.0};
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) s[0] += v[row + i] * v[row + i];
  red.allreduce(s);
  const double inv = 1.0 / sqrt(s[0]);
  for (int i = threadIdx.x; i < A.d; i += PK_THREADS) v[row + i] *= inv;
}
// Math::esh_momentum_update (math.rs:183-210, cpu_math.rs:505-551) on one row: three team reductions (|g|^2, p.g, |p_raw|^2); every thread
// only touches the elements tid + k * PK_THREADS, so no barrier is needed between the passes.  Returns the kinetic-energy change.
__device__ __forceinline__ double esh_row(const double* g, double* mom, double step_size, int d, TeamReduce<PK_THREADS>& red) {
  double s[1] = {0.0};
  for (int i = threadIdx.x; i < d; i += PK_THREADS) s[0] += g[i] * g[i];
  red.allreduce(s);
  const double grad_norm = sqrt(s[0]);
  const double inv_grad_norm = 1.0 / grad_norm;
  s[0] = 0.0;
  for (int i = threadIdx.x; i < d; i += PK_THREADS) s[0] += mom[i] * g[i] * inv_grad_norm;
  red.allreduce(s);
  const double momentum_proj = s[0];
  const double dims_m1 = (double)(d - 1);
  const double delta = step_size * grad_norm / dims_m1;
  const double zeta = exp(-delta);
  const double coeff_g = (1.0 - zeta) * (1.0 + zeta + momentum_proj * (1.0 - zeta));
  const double coeff_p = 2.0 * zeta;
  s[0] = 0.0;
  for (int i = threadIdx.x; i < d; i += PK_THREADS) {
    const double pr = coeff_g * (g[i] * inv_grad_norm) + coeff_p * mom[i];
    mom[i] = pr;
    s[0] += pr * pr;
  }
  red.allreduce(s);
  const double inv = 1.0 / sqrt(s[0]);
  for (int i = threadIdx.x; i < d; i += PK_THREADS) mom[i] *= inv;
  const double arg = momentum_proj + (1.0 - momentum_proj) * zeta * zeta;
  return (delta - 0.6931471805599453094 + log1p(arg)) * dims_m1;
}
__global__ void __launch_bounds__(PK_THREADS) k_esh_momentum_update(RowArgs A, const double* grad, double* mom, const double* step,
                                                                     double step_bcast, const uint8_t* active, double* dke) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  NB_ROW_PROLOGUE
  const double r = esh_row(grad + row, mom + row, step ? step[c] : step_bcast, A.d, red);
  if (threadIdx.x == 0) dke[c] = r;
}

// Hamiltonian::leapfrog (transformed_hamiltonian.rs:524-615) for KIND 1 = ExactNormal (:169-177, :206-213, :237-244: half-steps with the
// gradient of the standard-normal part removed, exact rotation in between) and KIND 2 = Microcanonical (:186-198, :214-227, :248-256: ESH
// momentum updates, step sizes scaled by sqrt(dim), kinetic_energy = accumulated change).  Diagonal transformation; sn / cs = sine and
// cosine of the signed step dir[c] * step_size[c], evaluated on the host (KIND 1).  HBM traffic per chain like k_leapfrog's three-pass branch.
template <int KIND>
__global__ void __launch_bounds__(PK_THREADS) k_leapfrog_kinetic(RowArgs A, ModelDev m, TransformDev T, PointDev s, PointDev o,
                                                                  const double* step_size, double step_bcast, const int8_t* dir,
                                                                  const double* sn, const double* cs, const double* baseline,
                                                                  double max_energy_error, const uint8_t* active, int* status,
                                                                  double* energy_error_out) {
  __shared__ double scratch[2 * 32 * REDUCE_MAXK];
  TeamReduce<PK_THREADS> red(scratch);
  NB_ROW_PROLOGUE
  const int d = A.d;
  const int body = (d / 4) * 4;
  const int sign = dir ? (int)dir[c] : 1;
  const double eps = (double)sign * (step_size ? step_size[c] : step_bcast) * 1.0;
  const double eps_half = eps / 2.;
  const double sqrt_d = sqrt((double)d);
  double ke = 0.0;
  if (KIND == 1) {
    const double es = sn[c], ec = cs[c];
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      const double z = s.z[row + i], pg = z + s.gz[row + i], v = s.v[row + i];
      const double vh = i < body ? fma(eps_half, pg, v) : v + eps_half * pg;  // std_norm_grad_flow(z, grad_z, v, out.v, eps / 2)
      double zn, vr;                                                            // std_norm_flow(z, out.z, out.v, eps)
      if (i < body) {
        zn = fma(z, ec, vh * es);
        vr = fma(z, -es, vh * ec);
      } else {
        zn = z * ec + vh * es;
        vr = z * (-es) + vh * ec;
      }
      o.z[row + i] = zn;
      o.v[row + i] = vr;
      o.x[row + i] = fma(1.0, T.mean[row + i], zn * T.stds[row + i]);
    }
  } else {
    for (int i = threadIdx.x; i < d; i += PK_THREADS) o.v[row + i] = s.v[row + i];
    ke = s.ke[c] + esh_row(s.gz + row, o.v + row, sqrt_d * eps / 2., d, red);
    const double e = eps * sqrt_d;
    for (int i = threadIdx.x; i < d; i += PK_THREADS) {
      const double zn = fma(e, o.v[row + i], s.z[row + i]);
      o.z[row + i] = zn;
      o.x[row + i] = fma(1.0, T.mean[row + i], zn * T.stds[row + i]);
    }
  }
  const double lp = model_eval_row(m, d, o.x + row, o.gx + row, red);
  double part[1] = {0.0};
  for (int i = threadIdx.x; i < d; i += PK_THREADS) {
    const double gn = o.gx[row + i] * T.stds[row + i];
    o.gz[row + i] = gn;
    if (KIND == 1) {
      const double pg = o.z[row + i] + gn, v = o.v[row + i];
      const double vn = i < body ? fma(eps_half, pg, v) : v + eps_half * pg;  // std_norm_grad_flow_inplace(out.z, out.grad_z, out.v, eps / 2)
      o.v[row + i] = vn;
      part[0] = fma(vn, vn, part[0]);
    }
  }
  if (KIND == 1) {
    red.allreduce(part);
    ke = 0.5 * part[0];
  } else {
    ke = ke + esh_row(o.gz + row, o.v + row, sqrt_d * eps / 2., d, red);
  }
  if (threadIdx.x == 0) {
    const double logdet = T.logdet[c];
    o.logp[c] = lp;
    o.logdet[c] = logdet;
    o.ke[c] = ke;
    o.e0[c] = s.e0[c];
    o.tid[c] = T.id[c];
    o.idx[c] = s.idx[c] + sign;
    const double base = baseline ? baseline[c] : s.e0[c];
    const double ee = (ke - (lp + logdet)) - base;
    if (energy_error_out) energy_error_out[c] = ee;
    const bool bad = KIND == 2 ? fabs(ee) >= max_energy_error : ee > max_energy_error;  // :591-596
    if (status) status[c] = (bad | !isfinite(ee)) ? 1 : 0;
  }
}

}  // namespace nb
