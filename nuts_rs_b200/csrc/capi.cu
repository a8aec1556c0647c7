// capi.cu — host side of libnuts_b200.so: the C ABI declared in include/nuts_b200.h.
// No torch, no CPU fallback: every entry point needs an sm_100 device and fails loudly otherwise.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "../../include/nuts_b200.h"
#include "chain_engine_v2.cuh"
#include "plane_kernels.cuh"

using namespace nb;

// ---- per-configuration launchers (engine_inst.cu) ----
// Weak declarations: a development build may contain only a subset of the (tiling, model) matrix (make CONFIGS=.. MODELS=..).
#define NB_DECL1(TPC, EPT, MINB, MODEL)                                                                                                      \
  extern "C" __attribute__((weak)) cudaError_t nb_launch_chain_##TPC##_##EPT##_##MINB##_##MODEL(const EngineParams* p, int grid, cudaStream_t s); \
  extern "C" __attribute__((weak)) cudaError_t nb_occupancy_chain_##TPC##_##EPT##_##MINB##_##MODEL(int* blocks_per_sm, int* cta_threads, int* smf);
#define NB_DECL(TPC, EPT, MINB) NB_DECL1(TPC, EPT, MINB, 1) NB_DECL1(TPC, EPT, MINB, 2) NB_DECL1(TPC, EPT, MINB, 3) NB_DECL1(TPC, EPT, MINB, 4)
NB_DECL(32, 1, 16)
NB_DECL(32, 2, 16)
NB_DECL(32, 4, 16)
NB_DECL(32, 8, 12)
NB_DECL(32, 16, 8)
NB_DECL(32, 32, 8)
NB_DECL(32, 32, 7)
NB_DECL(64, 16, 4)
// SM_ALIGN (tag 21): all warp teams of an SM in ONE CTA, starting their work units together (chain_engine.cuh)
NB_DECL(32, 1, 21)
NB_DECL(32, 2, 21)
NB_DECL(32, 4, 21)
NB_DECL(32, 8, 21)
NB_DECL(32, 16, 21)
// SM_LOWRANK (tag 31): the engines of nuts_sampler_create_lowrank
NB_DECL(32, 1, 31)
NB_DECL(32, 4, 31)
NB_DECL(32, 16, 31)
NB_DECL(64, 16, 31)
NB_DECL(64, 16, 44)   // SM_EXACT + SM_ALIGN: the 4 resident 64-thread teams of an SM in one 256-thread CTA
NB_DECL(64, 16, 54)  // SM_EXACT variant of 64x16x4 (rows padded to 1024)
NB_DECL(64, 16, 58)  // SM_EXACT + SM_STAGE, model parameters through the read-only path (no shared-memory copy)
NB_DECL(64, 16, 59)  // SM_EXACT + SM_STAGE + SM_STAGE1: one staging buffer, model parameters in shared memory, 4 CTAs per SM
NB_DECL(64, 16, 55)  // + SM_NOGRAD, 5 resident CTAs per SM
NB_DECL(64, 16, 57)  // SM_EXACT + gradient in shared memory, 200 registers: 5 resident CTAs per SM
NB_DECL(64, 16, 56)  // + SM_NOGRAD, 6 resident CTAs per SM
NB_DECL(64, 16, 5)
NB_DECL(64, 16, 6)
NB_DECL(64, 16, 7)
NB_DECL(64, 16, 8)
NB_DECL(128, 8, 5)
NB_DECL(128, 8, 53)  // SM_EXACT, 3 CTAs per SM (168 registers)
NB_DECL(128, 8, 52)  // SM_EXACT, 2 CTAs per SM
NB_DECL(256, 4, 62)  // SM_EXACT, 2 CTAs per SM
NB_DECL(256, 4, 61)  // SM_EXACT, 1 CTA per SM
NB_DECL(128, 8, 4)
NB_DECL(256, 8, 2)
NB_DECL(512, 8, 1)
NB_DECL(1024, 8, 1)
NB_DECL(1024, 10, 1)
NB_DECL(1024, 16, 1)
// decoupled engine (engine_v2_inst.cu): third number = 100 + teams per CTA; diagonal / isotropic Gaussian only
NB_DECL1(64, 16, 107, 1)
NB_DECL1(64, 16, 117, 1)
NB_DECL1(64, 16, 144, 1)
NB_DECL1(512, 20, 161, 1)
NB_DECL1(480, 21, 171, 1)
NB_DECL1(480, 18, 181, 1)
NB_DECL1(64, 16, 155, 1)
NB_DECL1(64, 16, 127, 1)
// cluster engine (engine_cl_inst.cu): the team spans the CTAs of a thread-block cluster
NB_DECL(1024, 10, 41)      // 4 CTAs x 256 threads, all model variants
NB_DECL1(768, 14, 42, 1)   // 2 CTAs x 384 threads (experiments)

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return fail(NUTS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

struct EngineConfig {
  int tpc, ept, minb, max_d;
  // indexed by model variant - 1 (1 diagonal/isotropic Gaussian, 2 rank-1 Gaussian, 3 funnel, 4 user density)
  cudaError_t (*launch[4])(const EngineParams*, int, cudaStream_t);
  cudaError_t (*occupancy[4])(int*, int*, int*);
  int min_d = 0;  // kDecoupledLarge only: chosen for min_d < dim <= max_d
};
#define NB_CFG(TPC, EPT, MINB)                                                                                                             \
  {                                                                                                                                        \
    TPC, EPT, MINB, TPC* EPT,                                                                                                              \
        {nb_launch_chain_##TPC##_##EPT##_##MINB##_1, nb_launch_chain_##TPC##_##EPT##_##MINB##_2, nb_launch_chain_##TPC##_##EPT##_##MINB##_3, \
         nb_launch_chain_##TPC##_##EPT##_##MINB##_4},                                                                                      \
    {                                                                                                                                      \
      nb_occupancy_chain_##TPC##_##EPT##_##MINB##_1, nb_occupancy_chain_##TPC##_##EPT##_##MINB##_2, nb_occupancy_chain_##TPC##_##EPT##_##MINB##_3, \
          nb_occupancy_chain_##TPC##_##EPT##_##MINB##_4                                                                                    \
    }                                                                                                                                      \
  }
#define NB_CFG1(TPC, EPT, MINB) \
  { TPC, EPT, MINB, TPC* EPT, {nb_launch_chain_##TPC##_##EPT##_##MINB##_1, nullptr, nullptr, nullptr}, {nb_occupancy_chain_##TPC##_##EPT##_##MINB##_1, nullptr, nullptr, nullptr} }
// default choice: the first entry whose capacity (tpc*ept) covers dim.  Warp-per-chain up to dim 1024 (no barrier in
// the whole kernel, all 1024 chains of config 2 resident at once); CTA-per-chain above.
const EngineConfig kConfigs[] = {NB_CFG(32, 1, 16), NB_CFG(32, 2, 16), NB_CFG(32, 4, 16),  NB_CFG(32, 8, 12),   NB_CFG(32, 16, 8),   NB_CFG(64, 16, 4),
                                 NB_CFG(256, 8, 2), NB_CFG(512, 8, 1), NB_CFG(1024, 8, 1), NB_CFG(1024, 10, 1), NB_CFG(1024, 16, 1)};
// decoupled engine for large dims (chosen for min_d < dim <= max_d, elementwise targets)
const EngineConfig kDecoupledLarge[] = {
    {480, 18, 181, 480 * 18, {nb_launch_chain_480_18_181_1, nullptr, nullptr, nullptr}, {nb_occupancy_chain_480_18_181_1, nullptr, nullptr, nullptr}, 4096},
    {480, 21, 171, 480 * 21, {nb_launch_chain_480_21_171_1, nullptr, nullptr, nullptr}, {nb_occupancy_chain_480_21_171_1, nullptr, nullptr, nullptr}, 480 * 18}};
// cluster engine (a team = the 4 CTAs of a thread-block cluster): default for the non-elementwise targets at 4096 < dim <= 10240
const EngineConfig kClusterLarge = NB_CFG(1024, 10, 41);
// SM_ALIGN variants of the warp tilings (tag 21): the default unless the target's tree depths vary wildly (funnel) or
// NUTS_B200_ALIGN=0
const EngineConfig kAlignedConfigs[] = {NB_CFG(32, 1, 21), NB_CFG(32, 2, 21), NB_CFG(32, 4, 21), NB_CFG(32, 8, 21), NB_CFG(32, 16, 21)};
// engines with the low-rank transformation compiled in (nuts_sampler_create_lowrank)
const EngineConfig kLowRankConfigs[] = {NB_CFG(32, 1, 31), NB_CFG(32, 4, 31), NB_CFG(32, 16, 31), NB_CFG(64, 16, 31)};
// warm-up build of the exact 64x16 tiling: the four resident teams of an SM in ONE 256-thread CTA that starts its work units together
// (SM_ALIGN, named barriers per team).  During the warm-up every draw ends in ~1500 instructions of adaptation code that sweep the
// 32 KB instruction cache; teams in lock-step share those fetches: config 2's 400-draw warm-up 80.6 -> 73.1 ms with two draws per
// unit.  After the warm-up the plain build is 1-4 % faster (no alignment barrier), so the choice is made per launch.
const EngineConfig kExactTuneConfig = NB_CFG(64, 16, 44);
// SM_EXACT variants of default tilings (tag = 50 + min blocks), chosen automatically when dim nearly fills the tile
const EngineConfig kExactConfigs[] = {NB_CFG(64, 16, 54)};
// alternatives selectable with NUTS_B200_ENGINE="tpc,ept,minb" (tuning experiments)
const EngineConfig kExtraConfigs[] = {NB_CFG(64, 16, 59), NB_CFG(64, 16, 44), NB_CFG(32, 1, 21), NB_CFG(32, 2, 21), NB_CFG(32, 4, 21), NB_CFG(32, 8, 21), NB_CFG(32, 16, 21), NB_CFG(1024, 10, 41), NB_CFG1(768, 14, 42), NB_CFG(128, 8, 53), NB_CFG(128, 8, 52), NB_CFG(256, 4, 62), NB_CFG(256, 4, 61), NB_CFG(64, 16, 58), NB_CFG(64, 16, 57), NB_CFG(64, 16, 55), NB_CFG(64, 16, 56), NB_CFG(64, 16, 54), NB_CFG1(64, 16, 107), NB_CFG1(64, 16, 117), NB_CFG1(64, 16, 127), NB_CFG1(64, 16, 144), NB_CFG1(512, 20, 161), NB_CFG1(480, 21, 171), NB_CFG1(480, 18, 181), NB_CFG1(64, 16, 155), NB_CFG(64, 16, 5), NB_CFG(64, 16, 6), NB_CFG(32, 32, 8), NB_CFG(32, 32, 7), NB_CFG(128, 8, 4), NB_CFG(64, 16, 7), NB_CFG(64, 16, 8), NB_CFG(128, 8, 5)};

}  // namespace

struct nuts_plane {
  double* ptr;
  uint64_t ld;  // row stride in doubles (dim rounded up to 16: rows start on 128-byte lines)
};

struct nuts_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t N = 0, d = 0, ld = 0;
  int num_sms = 0;
  ModelDev model{};
  double* d_model_mu = nullptr;
  double* d_model_prec = nullptr;
  double* d_model_user = nullptr;  // NUTS_LOGP_USER: device copy of user_params
  // Tier-2 transformation (one DiagMassMatrix per chain)
  TransformDev T{};
  // low-rank correction of the Tier-2 transformation (allocated by the first nuts_set_lowrank_transform)
  double* lr_vecs = nullptr;
  double *lr_vals_sqrt = nullptr, *lr_vals_sqrt_inv = nullptr, *lr_mu = nullptr;
  int* lr_rank = nullptr;
  int lr_rmax = 0;
  bool lr_active = false;  // some chain may carry a low-rank correction (since the last diagonal nuts_set_transform)
  // scratch
  double* d_dense = nullptr;    // [N*d] staging for host <-> plane packing
  double* d_sc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [N] f64 scratch
  uint8_t* d_u8 = nullptr;      // [N]
  int8_t* d_i8 = nullptr;       // [N]
  int* d_i32 = nullptr;         // [N]
  long long* d_i64 = nullptr;   // [N]
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around the kernel of the last nuts_leapfrog (nuts_ctx_last_kernel_ms)
  bool ev_valid = false;
  RowArgs row_args() const { return RowArgs{(int)N, (int)d, (int)ld}; }
};

struct nuts_point {
  nuts_plane planes[5];
  PointDev dev{};
};

struct nuts_sampler {
  nuts_ctx* ctx = nullptr;
  nuts_settings_t settings{};
  EngineParams P{};
  const EngineConfig* cfg = nullptr;
  int model_variant = 0;  // index into cfg->launch
  int grid = 0;
  // launches that lie entirely inside the warm-up use this build when set (same tiling and memory layout, other CTA shape)
  const EngineConfig* cfg_tune = nullptr;
  int grid_tune = 0;
  int teams_per_cta = 1;
  uint64_t resident_teams = 1;
  std::vector<void*> allocations;
  double* d_init = nullptr;
  int* d_status = nullptr;
  unsigned int* d_ring_template = nullptr;  // 0, 1, 2, ..: the ready ring's sequence numbers at the start of a launch
  // device stats buffers (grown on demand): ONE allocation holding the 15 arrays + a page-locked host mirror, so the statistics
  // of a nuts_draw call leave in a single D2H copy
  uint64_t stats_capacity = 0;  // in draws
  StatsDev d_stats{};
  unsigned char* d_stats_blob = nullptr;
  unsigned char* h_stats_blob = nullptr;
  size_t stats_offset[15] = {};
  size_t stats_blob_bytes = 0;
  double* d_draws = nullptr;
  uint64_t draws_capacity = 0;  // in draws
  double* h_pinned = nullptr;   // pinned staging for D2H of draws
  uint64_t pinned_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_kernel_ms = 0.0;
  uint64_t last_launches = 0;
  uint64_t draws_done = 0;
  bool positioned = false;
  bool last_direct = false;  // the last nuts_draw wrote its draws straight into the caller's buffer
  // low-rank sampler (nuts_sampler_create_lowrank)
  uint64_t lowrank_rmax = 0;
  double *lr_vecs = nullptr, *lr_vals_sqrt = nullptr, *lr_vals_sqrt_inv = nullptr, *lr_mu = nullptr;
  int* lr_rank = nullptr;
};

namespace {

int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(NUTS_ERR_NO_DEVICE, "no CUDA device visible (%s); libnuts_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  return NUTS_OK;
}

template <class T>
int dev_alloc(T** p, size_t count) {
  CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  CUDA_TRY(cudaMemset(*p, 0, count * sizeof(T)));
  return NUTS_OK;
}

int upload_mask(nuts_ctx* ctx, const uint8_t* active, const uint8_t** out) {
  *out = nullptr;
  if (!active) return NUTS_OK;
  CUDA_TRY(cudaMemcpyAsync(ctx->d_u8, active, ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  *out = ctx->d_u8;
  return NUTS_OK;
}
int upload_f64(nuts_ctx* ctx, const double* host, int slot, const double** out) {
  *out = nullptr;
  if (!host) return NUTS_OK;
  CUDA_TRY(cudaMemcpyAsync(ctx->d_sc[slot], host, ctx->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  *out = ctx->d_sc[slot];
  return NUTS_OK;
}
int download_f64(nuts_ctx* ctx, int slot, double* host) {
  if (!host) return NUTS_OK;
  CUDA_TRY(cudaMemcpyAsync(host, ctx->d_sc[slot], ctx->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return NUTS_OK;
}
int sync(nuts_ctx* ctx) {
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return NUTS_OK;
}
#define TRY(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r != NUTS_OK) return _r; \
  } while (0)
#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

// Destroys a half-built object when a constructor-like entry point returns early; dismiss() on success.
template <class T, int (*Destroy)(T*)>
struct Guard {
  T* p;
  explicit Guard(T* q) : p(q) {}
  ~Guard() {
    if (p) Destroy(p);
  }
  void dismiss() { p = nullptr; }
};

template <class T>
int grow(T** p, size_t count) {
  if (*p) CUDA_TRY(cudaFree(*p));
  *p = nullptr;
  CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  return NUTS_OK;
}


}  // namespace

extern "C" {

const char* nuts_last_error(void) { return g_last_error.c_str(); }

int nuts_device_available(void) {
  int r = check_device();
  if (r != NUTS_OK) return r;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) return fail(NUTS_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(NUTS_ERR_NO_DEVICE, "device 0 is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  return NUTS_OK;
}

void nuts_settings_default(nuts_settings_t* s) {
  // DiagNutsSettings::default(): reference src/sampler.rs:507-531 (num_tune 400, num_chains 6, max_energy_error 1000 at :630-634)
  std::memset(s, 0, sizeof(*s));
  s->num_tune = 400;
  s->num_draws = 1000;
  s->maxdepth = 10;
  s->mindepth = 0;
  s->max_energy_error = 1000.0;
  s->check_turning = 1;
  s->num_chains = 6;
  s->seed = 0;
  s->extra_doublings = 0;
  s->trajectory_kind = NUTS_KINETIC_EUCLIDEAN;
  nuts_euclidean_adapt_options_t& a = s->adapt_options;  // src/adapt_strategy.rs:56-69
  a.early_window = 0.3;
  a.step_size_window = 0.15;
  a.mass_matrix_switch_freq = 80;
  a.early_mass_matrix_switch_freq = 10;
  a.mass_matrix_update_freq = 1;
  a.mass_matrix_window_growth = 1.5;
  a.mass_matrix_options.store_mass_matrix = 0;  // src/transform/adapt/diagonal.rs:99-106
  a.mass_matrix_options.use_grad_based_estimate = 1;
  a.step_size_settings.target_accept = 0.8;  // src/stepsize/adapt.rs:320-329
  a.step_size_settings.initial_step = 0.1;
  a.step_size_settings.has_jitter = 1;
  a.step_size_settings.jitter = 0.1;
  a.step_size_settings.adapt_options.method = NUTS_STEPSIZE_DUAL_AVERAGE;
  a.step_size_settings.adapt_options.adam = nuts_adam_options_t{0.9, 0.999, 1e-8, 0.05};  // AdamOptions::default (adam.rs:25-34)
  a.step_size_settings.adapt_options.dual_average = {0.75, 10., 0.05, 3.14159265358979323846};  // src/stepsize/dual_avg.rs:22-31
}

// ============================================================ Tier 0
int nuts_ctx_create(nuts_ctx_t** out, int device_id, uint64_t nchains, uint64_t dim, const nuts_logp_desc_t* model) {
  TRY(check_device());
  if (!out || !model || nchains == 0 || dim == 0) return fail(NUTS_ERR_INVALID, "nuts_ctx_create: bad arguments");
  if (nchains > (1ull << 30) || dim > (1ull << 24)) return fail(NUTS_ERR_INVALID, "nuts_ctx_create: nchains/dim too large");
  CUDA_TRY(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) return fail(NUTS_ERR_NO_DEVICE, "device %d is sm_%d%d; libnuts_b200 is built for sm_100a only", device_id, prop.major, prop.minor);
  nuts_ctx* ctx = new nuts_ctx();
  Guard<nuts_ctx, nuts_ctx_destroy> guard(ctx);  // every early return below releases what was allocated so far
  ctx->device = device_id;
  ctx->N = nchains;
  ctx->d = dim;
  ctx->ld = (dim + 15) / 16 * 16;  // rows start on 128-byte boundaries
  ctx->num_sms = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  // model parameters
  const size_t model_ld = std::max<size_t>(ctx->ld, 16384);  // zero padding up to the largest tile: SM_EXACT engines read whole tiles
  std::vector<double> mu(model_ld, 0.0), prec(model_ld, 0.0);
  for (uint64_t i = 0; i < dim; ++i) mu[i] = model->mu ? model->mu[i] : model->mu_scalar;
  ctx->model.kind = model->kind;
  ctx->model.dim = (int)dim;
  switch (model->kind) {
    case NUTS_LOGP_GAUSS_ISO:
      for (uint64_t i = 0; i < dim; ++i) prec[i] = 1.0;  // diff * 1.0 is exact: the DIAG kernels serve this target bit-identically
      break;
    case NUTS_LOGP_GAUSS_DIAG:
      if (!model->sigma) return fail(NUTS_ERR_INVALID, "GAUSS_DIAG needs sigma");
      for (uint64_t i = 0; i < dim; ++i) prec[i] = 1.0 / (model->sigma[i] * model->sigma[i]);
      break;
    case NUTS_LOGP_GAUSS_RANK1:
      ctx->model.rank1_coeff = model->rank1_scale / (1.0 + model->rank1_scale * (double)dim);  // tests/sample_normal.rs:36
      break;
    case NUTS_LOGP_FUNNEL:
      ctx->model.funnel_inv_var = 1.0 / (model->funnel_scale * model->funnel_scale);
      break;
    case NUTS_LOGP_USER:  // the density compiled in from the user header (include/nuts_user_logp.cuh)
      if (model->n_user_params > 0 && !model->user_params) return fail(NUTS_ERR_INVALID, "NUTS_LOGP_USER: user_params is NULL");
      TRY(dev_alloc(&ctx->d_model_user, std::max<size_t>(model->n_user_params, 1)));
      if (model->n_user_params > 0)
        CUDA_TRY(cudaMemcpy(ctx->d_model_user, model->user_params, model->n_user_params * sizeof(double), cudaMemcpyHostToDevice));
      ctx->model.user = ctx->d_model_user;
      break;
    default:
      return fail(NUTS_ERR_INVALID, "unknown logp kind %d", model->kind);
  }
  TRY(dev_alloc(&ctx->d_model_mu, model_ld));
  TRY(dev_alloc(&ctx->d_model_prec, model_ld));
  CUDA_TRY(cudaMemcpy(ctx->d_model_mu, mu.data(), model_ld * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(ctx->d_model_prec, prec.data(), model_ld * sizeof(double), cudaMemcpyHostToDevice));
  ctx->model.mu = ctx->d_model_mu;
  ctx->model.prec = ctx->d_model_prec;
  // transformation planes: DiagMassMatrix::new (diagonal.rs:73-83): zero vectors, logdet 0, id -1
  const size_t plane = nchains * ctx->ld;
  TRY(dev_alloc(&ctx->T.stds, plane));
  TRY(dev_alloc(&ctx->T.inv_stds, plane));
  TRY(dev_alloc(&ctx->T.mean, plane));
  TRY(dev_alloc(&ctx->T.logdet, nchains));
  TRY(dev_alloc(&ctx->T.id, nchains));
  CUDA_TRY(cudaMemset(ctx->T.id, 0xff, nchains * sizeof(long long)));  // -1
  TRY(dev_alloc(&ctx->d_dense, nchains * dim));
  for (int k = 0; k < 6; ++k) TRY(dev_alloc(&ctx->d_sc[k], nchains));
  TRY(dev_alloc(&ctx->d_u8, nchains));
  TRY(dev_alloc(&ctx->d_i8, nchains));
  TRY(dev_alloc(&ctx->d_i32, nchains));
  TRY(dev_alloc(&ctx->d_i64, nchains));
  CUDA_TRY(cudaEventCreate(&ctx->ev_k0));
  CUDA_TRY(cudaEventCreate(&ctx->ev_k1));
  guard.dismiss();
  *out = ctx;
  return NUTS_OK;
}

int nuts_ctx_destroy(nuts_ctx_t* ctx) {
  if (!ctx) return NUTS_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_model_mu);
  cudaFree(ctx->d_model_prec);
  cudaFree(ctx->d_model_user);
  cudaFree(ctx->T.stds);
  cudaFree(ctx->T.inv_stds);
  cudaFree(ctx->T.mean);
  cudaFree(ctx->T.logdet);
  cudaFree(ctx->T.id);
  cudaFree(ctx->lr_vecs);
  cudaFree(ctx->lr_vals_sqrt);
  cudaFree(ctx->lr_vals_sqrt_inv);
  cudaFree(ctx->lr_mu);
  cudaFree(ctx->lr_rank);
  cudaFree(ctx->d_dense);
  for (int k = 0; k < 6; ++k) cudaFree(ctx->d_sc[k]);
  cudaFree(ctx->d_u8);
  cudaFree(ctx->d_i8);
  cudaFree(ctx->d_i32);
  cudaFree(ctx->d_i64);
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NUTS_OK;
}
int nuts_ctx_synchronize(nuts_ctx_t* ctx) { return sync(ctx); }
uint64_t nuts_ctx_nchains(const nuts_ctx_t* ctx) { return ctx->N; }
uint64_t nuts_ctx_dim(const nuts_ctx_t* ctx) { return ctx->d; }
void* nuts_ctx_stream(nuts_ctx_t* ctx) { return (void*)ctx->stream; }
int nuts_ctx_last_kernel_ms(nuts_ctx_t* ctx, float* ms) {
  if (!ctx || !ms) return fail(NUTS_ERR_INVALID, "nuts_ctx_last_kernel_ms: NULL argument");
  if (!ctx->ev_valid) return fail(NUTS_ERR_INVALID, "nuts_ctx_last_kernel_ms: no nuts_leapfrog call on this context yet");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
  return NUTS_OK;
}

int nuts_plane_alloc(nuts_ctx_t* ctx, nuts_plane_t** plane) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  nuts_plane* p = new nuts_plane();
  p->ld = ctx->ld;
  int r = dev_alloc(&p->ptr, ctx->N * ctx->ld);
  if (r != NUTS_OK) {
    delete p;
    return r;
  }
  *plane = p;
  return NUTS_OK;
}
int nuts_plane_free(nuts_ctx_t* ctx, nuts_plane_t* plane) {
  if (!plane) return NUTS_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(plane->ptr);
  delete plane;
  return NUTS_OK;
}
static int plane_from_host(nuts_ctx* ctx, double* plane, const double* src) {
  CUDA_TRY(cudaMemcpyAsync(ctx->d_dense, src, ctx->N * ctx->d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_pack<<<(unsigned)ctx->N, PK_THREADS, 0, ctx->stream>>>(ctx->row_args(), ctx->d_dense, plane);
  CHECK_LAUNCH();
  return sync(ctx);
}
static int plane_to_host(nuts_ctx* ctx, const double* plane, double* dst, int ld = 0) {
  RowArgs ra = ctx->row_args();
  if (ld > 0) ra.ld = ld;
  k_unpack<<<(unsigned)ctx->N, PK_THREADS, 0, ctx->stream>>>(ra, plane, ctx->d_dense);
  CHECK_LAUNCH();
  CUDA_TRY(cudaMemcpyAsync(dst, ctx->d_dense, ctx->N * ctx->d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
int nuts_plane_read_from_host(nuts_ctx_t* ctx, nuts_plane_t* dst, const double* src) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  return plane_from_host(ctx, dst->ptr, src);
}
int nuts_plane_write_to_host(nuts_ctx_t* ctx, const nuts_plane_t* src, double* dst) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  return plane_to_host(ctx, src->ptr, dst);
}
double* nuts_plane_device_ptr(nuts_plane_t* plane, uint64_t* row_stride_elems) {
  if (row_stride_elems) *row_stride_elems = plane->ld;
  return plane->ptr;
}

// ============================================================ Tier 1
#define GRID (unsigned)ctx->N, PK_THREADS, 0, ctx->stream

int nuts_axpy(nuts_ctx_t* ctx, const nuts_plane_t* x, nuts_plane_t* y, const double* a, double a_bcast, const uint8_t* active) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const double* da;
  const uint8_t* dm;
  TRY(upload_f64(ctx, a, 0, &da));
  TRY(upload_mask(ctx, active, &dm));
  k_axpy<<<GRID>>>(ctx->row_args(), x->ptr, y->ptr, da, a_bcast, dm);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_axpy_out(nuts_ctx_t* ctx, const nuts_plane_t* x, const nuts_plane_t* y, const double* a, double a_bcast, nuts_plane_t* out,
                  const uint8_t* active) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const double* da;
  const uint8_t* dm;
  TRY(upload_f64(ctx, a, 0, &da));
  TRY(upload_mask(ctx, active, &dm));
  k_axpy_out<<<GRID>>>(ctx->row_args(), x->ptr, y->ptr, da, a_bcast, out->ptr, dm);
  CHECK_LAUNCH();
  return sync(ctx);
}
// sine and cosine of the per-chain angles on the HOST (the reference calls f64::sin / f64::cos once per vector, util.rs:580-581) -> slots 4, 5
static int upload_sincos(nuts_ctx* ctx, const double* angle, const double** dsn, const double** dcs) {
  *dsn = *dcs = nullptr;
  if (!angle) return NUTS_OK;
  std::vector<double> sc(2 * ctx->N);
  for (size_t c = 0; c < ctx->N; ++c) {
    sc[c] = std::sin(angle[c]);
    sc[ctx->N + c] = std::cos(angle[c]);
  }
  CUDA_TRY(cudaMemcpyAsync(ctx->d_sc[4], sc.data(), ctx->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_sc[5], sc.data() + ctx->N, ctx->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // sc is pageable and leaves scope
  *dsn = ctx->d_sc[4];
  *dcs = ctx->d_sc[5];
  return NUTS_OK;
}
int nuts_std_norm_flow(nuts_ctx_t* ctx, const nuts_plane_t* pos, nuts_plane_t* pos_out, nuts_plane_t* vel, const double* epsilon,
                       double epsilon_bcast, const uint8_t* active) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (pos == pos_out || pos == vel || pos_out == vel) return fail(NUTS_ERR_INVALID, "nuts_std_norm_flow: the three planes must be distinct");
  const double *dsn, *dcs;
  const uint8_t* dm;
  TRY(upload_sincos(ctx, epsilon, &dsn, &dcs));
  TRY(upload_mask(ctx, active, &dm));
  k_std_norm_flow<<<GRID>>>(ctx->row_args(), pos->ptr, pos_out->ptr, vel->ptr, dsn, dcs, std::sin(epsilon_bcast), std::cos(epsilon_bcast), dm);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_std_norm_grad_flow(nuts_ctx_t* ctx, const nuts_plane_t* pos, const nuts_plane_t* grad, const nuts_plane_t* vel,
                            nuts_plane_t* vel_out, const double* epsilon, double epsilon_bcast, const uint8_t* active) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const double* de;
  const uint8_t* dm;
  TRY(upload_f64(ctx, epsilon, 0, &de));
  TRY(upload_mask(ctx, active, &dm));
  k_std_norm_grad_flow<<<GRID>>>(ctx->row_args(), pos->ptr, grad->ptr, vel->ptr, vel_out->ptr, de, epsilon_bcast, dm);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_std_norm_grad_flow_inplace(nuts_ctx_t* ctx, const nuts_plane_t* pos, const nuts_plane_t* grad, nuts_plane_t* vel,
                                    const double* epsilon, double epsilon_bcast, const uint8_t* active) {
  return nuts_std_norm_grad_flow(ctx, pos, grad, vel, vel, epsilon, epsilon_bcast, active);
}
int nuts_array_normalize(nuts_ctx_t* ctx, nuts_plane_t* v, const uint8_t* active) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const uint8_t* dm;
  TRY(upload_mask(ctx, active, &dm));
  k_normalize<<<GRID>>>(ctx->row_args(), v->ptr, dm);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_esh_momentum_update(nuts_ctx_t* ctx, const nuts_plane_t* gradient, nuts_plane_t* momentum, const double* step_size,
                             double step_size_bcast, const uint8_t* active, double* kinetic_energy_change) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (ctx->d < 2) return fail(NUTS_ERR_INVALID, "nuts_esh_momentum_update: ESH dynamics requires at least 2 dimensions");  // cpu_math.rs:514
  const double* ds;
  const uint8_t* dm;
  TRY(upload_f64(ctx, step_size, 0, &ds));
  TRY(upload_mask(ctx, active, &dm));
  CUDA_TRY(cudaMemsetAsync(ctx->d_sc[2], 0, ctx->N * sizeof(double), ctx->stream));
  k_esh_momentum_update<<<GRID>>>(ctx->row_args(), gradient->ptr, momentum->ptr, ds, step_size_bcast, dm, ctx->d_sc[2]);
  CHECK_LAUNCH();
  TRY(download_f64(ctx, 2, kinetic_energy_change));
  return sync(ctx);
}
int nuts_array_mult(nuts_ctx_t* ctx, const nuts_plane_t* a1, const nuts_plane_t* a2, nuts_plane_t* dest) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_mult<<<GRID>>>(ctx->row_args(), a1->ptr, a2->ptr, dest->ptr);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_array_mult_inplace(nuts_ctx_t* ctx, nuts_plane_t* a1, const nuts_plane_t* a2) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_mult<<<GRID>>>(ctx->row_args(), a2->ptr, a1->ptr, a1->ptr);  // out = x * out (util.rs:94-97)
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_array_recip(nuts_ctx_t* ctx, const nuts_plane_t* a, nuts_plane_t* dest) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_recip<<<GRID>>>(ctx->row_args(), a->ptr, dest->ptr);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_fill_array(nuts_ctx_t* ctx, nuts_plane_t* a, double val) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_fill<<<GRID>>>(ctx->row_args(), a->ptr, val);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_copy_into(nuts_ctx_t* ctx, const nuts_plane_t* src, nuts_plane_t* dst) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_copy<<<GRID>>>(ctx->row_args(), src->ptr, dst->ptr);
  CHECK_LAUNCH();
  return sync(ctx);
}
static int reduce1(nuts_ctx* ctx, int op, const double* x, const double* y, double* out_f64, uint8_t* out_u8) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_reduce1<<<GRID>>>(ctx->row_args(), op, x, y, ctx->d_sc[0]);
  CHECK_LAUNCH();
  if (out_f64) {
    TRY(download_f64(ctx, 0, out_f64));
    return sync(ctx);
  }
  std::vector<double> tmp(ctx->N);
  TRY(download_f64(ctx, 0, tmp.data()));
  TRY(sync(ctx));
  for (uint64_t c = 0; c < ctx->N; ++c) out_u8[c] = tmp[c] != 0.0;
  return NUTS_OK;
}
int nuts_array_vector_dot(nuts_ctx_t* ctx, const nuts_plane_t* a1, const nuts_plane_t* a2, double* out) {
  return reduce1(ctx, 0, a1->ptr, a2->ptr, out, nullptr);
}
int nuts_sq_norm_sum(nuts_ctx_t* ctx, const nuts_plane_t* x, const nuts_plane_t* y, double* out) {
  return reduce1(ctx, 1, x->ptr, y->ptr, out, nullptr);
}
int nuts_array_sum_ln(nuts_ctx_t* ctx, const nuts_plane_t* a, double* out) { return reduce1(ctx, 2, a->ptr, nullptr, out, nullptr); }
int nuts_array_all_finite(nuts_ctx_t* ctx, const nuts_plane_t* a, uint8_t* out) { return reduce1(ctx, 3, a->ptr, nullptr, nullptr, out); }
int nuts_array_all_finite_and_nonzero(nuts_ctx_t* ctx, const nuts_plane_t* a, uint8_t* out) {
  return reduce1(ctx, 4, a->ptr, nullptr, nullptr, out);
}
int nuts_scalar_prods3(nuts_ctx_t* ctx, const nuts_plane_t* positive1, const nuts_plane_t* negative1, const nuts_plane_t* positive2,
                       const nuts_plane_t* x, const nuts_plane_t* y, double* out1, double* out2) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_scalar_prods<<<GRID>>>(ctx->row_args(), positive1->ptr, negative1->ptr, positive2->ptr, x->ptr, y->ptr, ctx->d_sc[0], ctx->d_sc[1]);
  CHECK_LAUNCH();
  TRY(download_f64(ctx, 0, out1));
  TRY(download_f64(ctx, 1, out2));
  return sync(ctx);
}
int nuts_scalar_prods2(nuts_ctx_t* ctx, const nuts_plane_t* positive1, const nuts_plane_t* positive2, const nuts_plane_t* x,
                       const nuts_plane_t* y, double* out1, double* out2) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_scalar_prods<<<GRID>>>(ctx->row_args(), positive1->ptr, nullptr, positive2->ptr, x->ptr, y->ptr, ctx->d_sc[0], ctx->d_sc[1]);
  CHECK_LAUNCH();
  TRY(download_f64(ctx, 0, out1));
  TRY(download_f64(ctx, 1, out2));
  return sync(ctx);
}
int nuts_array_gaussian(nuts_ctx_t* ctx, nuts_plane_t* dest, const nuts_plane_t* stds, uint64_t seed, uint64_t chain_offset, uint64_t counter) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_gaussian<<<GRID>>>(ctx->row_args(), dest->ptr, stds->ptr, seed, chain_offset, counter);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_array_update_variance(nuts_ctx_t* ctx, nuts_plane_t* mean, nuts_plane_t* variance, const nuts_plane_t* value,
                               const double* diff_scale, double diff_scale_bcast) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const double* ds;
  TRY(upload_f64(ctx, diff_scale, 0, &ds));
  k_update_variance<<<GRID>>>(ctx->row_args(), mean->ptr, variance->ptr, value->ptr, ds, diff_scale_bcast);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_array_update_var_inv_std_draw(nuts_ctx_t* ctx, nuts_plane_t* inv_std, nuts_plane_t* std_, const nuts_plane_t* draw_var,
                                       double scale, int has_fill, double fill_invalid, double clamp_lo, double clamp_hi) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_update_var_inv_std<<<GRID>>>(ctx->row_args(), 0, inv_std->ptr, std_->ptr, draw_var->ptr, nullptr, scale, has_fill, fill_invalid, clamp_lo,
                                 clamp_hi);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_array_update_var_inv_std_draw_grad(nuts_ctx_t* ctx, nuts_plane_t* inv_std, nuts_plane_t* std_, const nuts_plane_t* draw_var,
                                            const nuts_plane_t* grad_var, int has_fill, double fill_invalid, double clamp_lo,
                                            double clamp_hi) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_update_var_inv_std<<<GRID>>>(ctx->row_args(), 1, inv_std->ptr, std_->ptr, draw_var->ptr, grad_var->ptr, 0.0, has_fill, fill_invalid,
                                 clamp_lo, clamp_hi);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_array_update_var_inv_std_grad(nuts_ctx_t* ctx, nuts_plane_t* inv_std, nuts_plane_t* std_, const nuts_plane_t* gradient,
                                       double fill_invalid, double clamp_lo, double clamp_hi) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_update_var_inv_std<<<GRID>>>(ctx->row_args(), 2, inv_std->ptr, std_->ptr, gradient->ptr, nullptr, 0.0, 1, fill_invalid, clamp_lo, clamp_hi);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_logp_array(nuts_ctx_t* ctx, const nuts_plane_t* position, nuts_plane_t* gradient, double* logp, int32_t* status) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_logp_array<<<GRID>>>(ctx->row_args(), ctx->model, position->ptr, gradient->ptr, ctx->d_sc[0], ctx->d_i32);
  CHECK_LAUNCH();
  TRY(download_f64(ctx, 0, logp));
  if (status) CUDA_TRY(cudaMemcpyAsync(status, ctx->d_i32, ctx->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

// ============================================================ Tier 2
int nuts_point_alloc(nuts_ctx_t* ctx, nuts_point_t** point) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  nuts_point* p = new nuts_point();
  struct PointGuard {
    nuts_ctx* c;
    nuts_point* p;
    ~PointGuard() {
      if (p) nuts_point_free(c, p);
    }
  } guard{ctx, p};
  const size_t plane = ctx->N * ctx->ld;
  for (int k = 0; k < 5; ++k) {
    p->planes[k].ld = ctx->ld;
    TRY(dev_alloc(&p->planes[k].ptr, plane));
  }
  p->dev.x = p->planes[0].ptr;
  p->dev.gx = p->planes[1].ptr;
  p->dev.z = p->planes[2].ptr;
  p->dev.gz = p->planes[3].ptr;
  p->dev.v = p->planes[4].ptr;
  TRY(dev_alloc(&p->dev.idx, ctx->N));
  TRY(dev_alloc(&p->dev.logp, ctx->N));
  TRY(dev_alloc(&p->dev.logdet, ctx->N));
  TRY(dev_alloc(&p->dev.ke, ctx->N));
  TRY(dev_alloc(&p->dev.e0, ctx->N));
  TRY(dev_alloc(&p->dev.tid, ctx->N));
  CUDA_TRY(cudaMemset(p->dev.tid, 0xff, ctx->N * sizeof(long long)));  // transform_id = -1 (transformed_hamiltonian.rs:376)
  guard.p = nullptr;
  *point = p;
  return NUTS_OK;
}
int nuts_point_free(nuts_ctx_t* ctx, nuts_point_t* p) {
  if (!p) return NUTS_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int k = 0; k < 5; ++k) cudaFree(p->planes[k].ptr);
  cudaFree(p->dev.idx);
  cudaFree(p->dev.logp);
  cudaFree(p->dev.logdet);
  cudaFree(p->dev.ke);
  cudaFree(p->dev.e0);
  cudaFree(p->dev.tid);
  delete p;
  return NUTS_OK;
}
nuts_plane_t* nuts_point_plane(nuts_point_t* point, int which) {
  if (which < 0 || which > 4) return nullptr;
  return &point->planes[which];
}
int nuts_point_get_scalars(nuts_ctx_t* ctx, const nuts_point_t* p, int64_t* idx, double* logp, double* logdet, double* ke, double* e0,
                           int64_t* tid) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t n8 = ctx->N * 8;
  if (idx) CUDA_TRY(cudaMemcpyAsync(idx, p->dev.idx, n8, cudaMemcpyDeviceToHost, ctx->stream));
  if (logp) CUDA_TRY(cudaMemcpyAsync(logp, p->dev.logp, n8, cudaMemcpyDeviceToHost, ctx->stream));
  if (logdet) CUDA_TRY(cudaMemcpyAsync(logdet, p->dev.logdet, n8, cudaMemcpyDeviceToHost, ctx->stream));
  if (ke) CUDA_TRY(cudaMemcpyAsync(ke, p->dev.ke, n8, cudaMemcpyDeviceToHost, ctx->stream));
  if (e0) CUDA_TRY(cudaMemcpyAsync(e0, p->dev.e0, n8, cudaMemcpyDeviceToHost, ctx->stream));
  if (tid) CUDA_TRY(cudaMemcpyAsync(tid, p->dev.tid, n8, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
int nuts_point_set_scalars(nuts_ctx_t* ctx, nuts_point_t* p, const int64_t* idx, const double* logp, const double* logdet, const double* ke,
                           const double* e0, const int64_t* tid) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t n8 = ctx->N * 8;
  if (idx) CUDA_TRY(cudaMemcpyAsync(p->dev.idx, idx, n8, cudaMemcpyHostToDevice, ctx->stream));
  if (logp) CUDA_TRY(cudaMemcpyAsync(p->dev.logp, logp, n8, cudaMemcpyHostToDevice, ctx->stream));
  if (logdet) CUDA_TRY(cudaMemcpyAsync(p->dev.logdet, logdet, n8, cudaMemcpyHostToDevice, ctx->stream));
  if (ke) CUDA_TRY(cudaMemcpyAsync(p->dev.ke, ke, n8, cudaMemcpyHostToDevice, ctx->stream));
  if (e0) CUDA_TRY(cudaMemcpyAsync(p->dev.e0, e0, n8, cudaMemcpyHostToDevice, ctx->stream));
  if (tid) CUDA_TRY(cudaMemcpyAsync(p->dev.tid, tid, n8, cudaMemcpyHostToDevice, ctx->stream));
  return sync(ctx);
}
int nuts_set_transform(nuts_ctx_t* ctx, const double* stds, const double* mean) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(plane_from_host(ctx, ctx->T.stds, stds));
  TRY(plane_from_host(ctx, ctx->T.mean, mean));
  if (ctx->lr_rank) CUDA_TRY(cudaMemsetAsync(ctx->lr_rank, 0xff, ctx->N * sizeof(int), ctx->stream));  // inner = None (rank -1)
  ctx->lr_active = false;
  k_set_transform<<<GRID>>>(ctx->row_args(), ctx->T, nullptr);
  CHECK_LAUNCH();
  return sync(ctx);
}

// ---- low-rank mass matrix: EigVectors / EigValues of all chains, Math::apply_lowrank_transform, LowRankMassMatrix::update
struct nuts_eigs {
  double* vecs = nullptr;  // [N][rmax][ld]
  double* vals = nullptr;  // [N][rmax]
  int* rank = nullptr;     // [N]
  int rmax = 0;
};
// host [N][rmax][d] -> device [N][rmax][ld] (rows padded with zeros like every plane)
static int upload_vecs(nuts_ctx* ctx, double* dst, const double* vecs, uint64_t rmax) {
  CUDA_TRY(cudaMemsetAsync(dst, 0, ctx->N * rmax * ctx->ld * sizeof(double), ctx->stream));
  CUDA_TRY(cudaMemcpy2DAsync(dst, ctx->ld * sizeof(double), vecs, ctx->d * sizeof(double), ctx->d * sizeof(double), ctx->N * rmax,
                             cudaMemcpyHostToDevice, ctx->stream));
  return NUTS_OK;
}
int nuts_eigs_create(nuts_ctx_t* ctx, nuts_eigs_t** out, uint64_t rank_max, const double* vecs, const double* vals, const int32_t* rank) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!out || !vecs || !vals) return fail(NUTS_ERR_INVALID, "nuts_eigs_create: NULL argument");
  if (rank_max == 0 || rank_max > (uint64_t)LR_MAX_RANK) return fail(NUTS_ERR_INVALID, "nuts_eigs_create: rank_max must be 1 .. %d", LR_MAX_RANK);
  nuts_eigs* e = new nuts_eigs();
  struct G {
    nuts_ctx* c;
    nuts_eigs* e;
    ~G() {
      if (e) nuts_eigs_free(c, e);
    }
  } guard{ctx, e};
  e->rmax = (int)rank_max;
  TRY(dev_alloc(&e->vecs, ctx->N * rank_max * ctx->ld));
  TRY(dev_alloc(&e->vals, ctx->N * rank_max));
  TRY(dev_alloc(&e->rank, ctx->N));
  TRY(upload_vecs(ctx, e->vecs, vecs, rank_max));
  CUDA_TRY(cudaMemcpyAsync(e->vals, vals, ctx->N * rank_max * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<int> rk(ctx->N, (int)rank_max);
  if (rank)
    for (uint64_t c = 0; c < ctx->N; ++c) {
      if (rank[c] < 0 || (uint64_t)rank[c] > rank_max) return fail(NUTS_ERR_INVALID, "nuts_eigs_create: rank[%llu] = %d outside 0 .. rank_max", (unsigned long long)c, rank[c]);
      rk[c] = rank[c];
    }
  CUDA_TRY(cudaMemcpyAsync(e->rank, rk.data(), ctx->N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  TRY(sync(ctx));
  guard.e = nullptr;
  *out = e;
  return NUTS_OK;
}
int nuts_eigs_free(nuts_ctx_t* ctx, nuts_eigs_t* e) {
  if (!e) return NUTS_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(e->vecs);
  cudaFree(e->vals);
  cudaFree(e->rank);
  delete e;
  return NUTS_OK;
}
int nuts_apply_lowrank_transform(nuts_ctx_t* ctx, const nuts_eigs_t* e, const nuts_plane_t* rhs, nuts_plane_t* dest) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!e || !rhs || !dest) return fail(NUTS_ERR_INVALID, "nuts_apply_lowrank_transform: NULL argument");
  k_lowrank_apply<<<GRID>>>(ctx->row_args(), e->vecs, e->vals, e->rank, e->rmax, rhs->ptr, dest->ptr);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_apply_lowrank_transform_inplace(nuts_ctx_t* ctx, const nuts_eigs_t* e, nuts_plane_t* rhs_and_dest) {
  return nuts_apply_lowrank_transform(ctx, e, rhs_and_dest, rhs_and_dest);
}
int nuts_set_lowrank_transform(nuts_ctx_t* ctx, const double* stds, const double* mean, uint64_t rank_max, const double* vals,
                               const double* vecs, const int32_t* rank, const double* mean_low_rank, uint8_t* accepted) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!stds || !mean || !mean_low_rank) return fail(NUTS_ERR_INVALID, "nuts_set_lowrank_transform: NULL argument");
  if (rank_max > (uint64_t)LR_MAX_RANK) return fail(NUTS_ERR_INVALID, "nuts_set_lowrank_transform: rank_max must be <= %d", LR_MAX_RANK);
  if (rank_max > 0 && (!vals || !vecs)) return fail(NUTS_ERR_INVALID, "nuts_set_lowrank_transform: vals / vecs are NULL");
  const uint64_t N = ctx->N, d = ctx->d, R = std::max<uint64_t>(rank_max, 1);
  if (!ctx->lr_rank || (uint64_t)ctx->lr_rmax < R) {  // (re)allocate for the larger rank
    // (a larger rank than before: every chain falls back to its diagonal part until this call has installed the new correction)
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->lr_vecs), cudaFree(ctx->lr_vals_sqrt), cudaFree(ctx->lr_vals_sqrt_inv), cudaFree(ctx->lr_mu), cudaFree(ctx->lr_rank);
    ctx->lr_vecs = ctx->lr_vals_sqrt = ctx->lr_vals_sqrt_inv = ctx->lr_mu = nullptr;
    ctx->lr_rank = nullptr;
    ctx->lr_rmax = 0;
    ctx->lr_active = false;
    ctx->T.lr_vecs = ctx->T.lr_vals_sqrt = ctx->T.lr_vals_sqrt_inv = ctx->T.lr_mu = nullptr;
    ctx->T.lr_rank = nullptr;
    ctx->T.lr_rmax = 0;
    TRY(dev_alloc(&ctx->lr_vecs, N * R * ctx->ld));
    TRY(dev_alloc(&ctx->lr_vals_sqrt, N * R));
    TRY(dev_alloc(&ctx->lr_vals_sqrt_inv, N * R));
    TRY(dev_alloc(&ctx->lr_mu, N * ctx->ld));
    TRY(dev_alloc(&ctx->lr_rank, N));
    CUDA_TRY(cudaMemset(ctx->lr_rank, 0xff, N * sizeof(int)));  // -1: no inner matrix yet
    ctx->lr_rmax = (int)R;
  }
  const uint64_t RM = (uint64_t)ctx->lr_rmax;
  // per chain: finiteness (low_rank.rs:168-173), sqrt(lambda), 1 / sqrt(lambda), -1/2 sum ln(lambda) (low_rank.rs:55-71)
  std::vector<uint8_t> ok(N, 1);
  std::vector<int> rk(N, 0);
  std::vector<double> vs(N * RM, 1.0), vi(N * RM, 1.0), contrib(N, 0.0);
  auto finite = [](const double* p, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
      if (!std::isfinite(p[i])) return false;
    return true;
  };
  for (uint64_t c = 0; c < N; ++c) {
    const int r = rank_max == 0 ? 0 : (rank ? rank[c] : (int)rank_max);
    if (r < 0 || (uint64_t)r > rank_max) return fail(NUTS_ERR_INVALID, "nuts_set_lowrank_transform: rank[%llu] = %d outside 0 .. rank_max", (unsigned long long)c, r);
    bool good = finite(stds + c * d, d) && finite(mean + c * d, d);
    if (r > 0) good = good && finite(vals + c * rank_max, (uint64_t)r) && finite(vecs + c * rank_max * d, (uint64_t)r * d);
    ok[c] = good ? 1 : 0;
    rk[c] = r;
    for (int k = 0; k < r; ++k) {
      const double lam = vals[c * rank_max + k];
      contrib[c] += -0.5 * std::log(lam);
      const double sq = std::sqrt(lam);
      vs[c * RM + k] = sq;
      vi[c * RM + k] = 1.0 / sq;
    }
  }
  bool all_ok = true;
  for (uint8_t b : ok) all_ok = all_ok && b;
  const uint8_t* d_mask = nullptr;
  if (all_ok) {
    TRY(plane_from_host(ctx, ctx->T.stds, stds));
    TRY(plane_from_host(ctx, ctx->T.mean, mean));
    TRY(plane_from_host(ctx, ctx->lr_mu, mean_low_rank));
    if (rank_max > 0) {
      CUDA_TRY(cudaMemsetAsync(ctx->lr_vecs, 0, N * RM * ctx->ld * sizeof(double), ctx->stream));
      for (uint64_t c = 0; c < N; ++c)  // [c][rank_max][d] -> [c][RM][ld]
        CUDA_TRY(cudaMemcpy2DAsync(ctx->lr_vecs + c * RM * ctx->ld, ctx->ld * sizeof(double), vecs + c * rank_max * d, d * sizeof(double),
                                   d * sizeof(double), rank_max, cudaMemcpyHostToDevice, ctx->stream));
    }
  } else {
    // some chains keep their old transformation (low_rank.rs:168-173): row-by-row upload of the accepted ones, and the old rank,
    // eigenvalues and log-determinant of the others stay what they are
    std::vector<int> old_rank(N);
    std::vector<double> old_vs(N * RM), old_vi(N * RM);
    CUDA_TRY(cudaMemcpyAsync(old_rank.data(), ctx->lr_rank, N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(old_vs.data(), ctx->lr_vals_sqrt, N * RM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(old_vi.data(), ctx->lr_vals_sqrt_inv, N * RM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync(ctx));
    for (uint64_t c = 0; c < N; ++c) {
      if (!ok[c]) {
        rk[c] = old_rank[c];
        contrib[c] = 0.0;
        std::copy(old_vs.begin() + c * RM, old_vs.begin() + (c + 1) * RM, vs.begin() + c * RM);
        std::copy(old_vi.begin() + c * RM, old_vi.begin() + (c + 1) * RM, vi.begin() + c * RM);
        continue;
      }
      const size_t row = c * ctx->ld;
      CUDA_TRY(cudaMemcpyAsync(ctx->T.stds + row, stds + c * d, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(ctx->T.mean + row, mean + c * d, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      CUDA_TRY(cudaMemcpyAsync(ctx->lr_mu + row, mean_low_rank + c * d, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      CUDA_TRY(cudaMemsetAsync(ctx->lr_vecs + c * RM * ctx->ld, 0, RM * ctx->ld * sizeof(double), ctx->stream));
      if (rank_max > 0)
        CUDA_TRY(cudaMemcpy2DAsync(ctx->lr_vecs + c * RM * ctx->ld, ctx->ld * sizeof(double), vecs + c * rank_max * d, d * sizeof(double),
                                   d * sizeof(double), rank_max, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->d_u8, ok.data(), N, cudaMemcpyHostToDevice, ctx->stream));
    d_mask = ctx->d_u8;
  }
  CUDA_TRY(cudaMemcpyAsync(ctx->lr_vals_sqrt, vs.data(), N * RM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->lr_vals_sqrt_inv, vi.data(), N * RM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->lr_rank, rk.data(), N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_sc[3], contrib.data(), N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ctx->T.lr_vecs = ctx->lr_vecs;
  ctx->T.lr_vals_sqrt = ctx->lr_vals_sqrt;
  ctx->T.lr_vals_sqrt_inv = ctx->lr_vals_sqrt_inv;
  ctx->T.lr_mu = ctx->lr_mu;
  ctx->T.lr_rank = ctx->lr_rank;
  ctx->T.lr_rmax = ctx->lr_rmax;
  ctx->lr_active = true;
  k_set_transform<<<GRID>>>(ctx->row_args(), ctx->T, d_mask);  // diag.set_transform: inv_stds, sum ln(1 / sigma), id += 1
  CHECK_LAUNCH();
  k_add_logdet<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>((int)N, ctx->T.logdet, ctx->d_sc[3]);
  CHECK_LAUNCH();
  if (accepted) std::memcpy(accepted, ok.data(), N);
  return sync(ctx);
}
int nuts_get_transform(nuts_ctx_t* ctx, double* stds, double* inv_stds, double* mean, double* logdet, int64_t* id) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (stds) TRY(plane_to_host(ctx, ctx->T.stds, stds));
  if (inv_stds) TRY(plane_to_host(ctx, ctx->T.inv_stds, inv_stds));
  if (mean) TRY(plane_to_host(ctx, ctx->T.mean, mean));
  if (logdet) CUDA_TRY(cudaMemcpyAsync(logdet, ctx->T.logdet, ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (id) CUDA_TRY(cudaMemcpyAsync(id, ctx->T.id, ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
int nuts_init_state(nuts_ctx_t* ctx, nuts_point_t* point, const double* position, int32_t* status) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(plane_from_host(ctx, point->dev.x, position));
  k_init_state<<<GRID>>>(ctx->row_args(), ctx->model, ctx->T, point->dev, ctx->d_i32);
  CHECK_LAUNCH();
  if (status) CUDA_TRY(cudaMemcpyAsync(status, ctx->d_i32, ctx->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}
static int check_kinetic_kind(nuts_ctx* ctx, int kind, const char* who) {
  if (kind != NUTS_KINETIC_EUCLIDEAN && kind != NUTS_KINETIC_EXACT_NORMAL && kind != NUTS_KINETIC_MICROCANONICAL)
    return fail(NUTS_ERR_INVALID, "%s: unknown kinetic energy kind %d", who, kind);
  if (kind == NUTS_KINETIC_MICROCANONICAL && ctx->d < 2)
    return fail(NUTS_ERR_INVALID, "%s: ESH dynamics requires at least 2 dimensions", who);
  return NUTS_OK;
}
int nuts_initialize_trajectory_kinetic(nuts_ctx_t* ctx, int kind, nuts_point_t* point, int resample_velocity, uint64_t seed,
                                       uint64_t chain_offset, uint64_t counter) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(check_kinetic_kind(ctx, kind, "nuts_initialize_trajectory_kinetic"));
  k_initialize_trajectory<<<GRID>>>(ctx->row_args(), ctx->T, point->dev, resample_velocity, seed, chain_offset, counter, kind);
  CHECK_LAUNCH();
  return sync(ctx);
}
int nuts_initialize_trajectory(nuts_ctx_t* ctx, nuts_point_t* point, int resample_velocity, uint64_t seed, uint64_t chain_offset,
                               uint64_t counter) {
  return nuts_initialize_trajectory_kinetic(ctx, NUTS_KINETIC_EUCLIDEAN, point, resample_velocity, seed, chain_offset, counter);
}
int nuts_leapfrog_kinetic(nuts_ctx_t* ctx, int kind, const nuts_point_t* start, nuts_point_t* out, const double* step_size,
                          double step_size_bcast, const int8_t* dir, const double* energy_baseline, double max_energy_error,
                          const uint8_t* active, int32_t* status, double* energy_error) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  TRY(check_kinetic_kind(ctx, kind, "nuts_leapfrog_kinetic"));
  if (kind == NUTS_KINETIC_EUCLIDEAN)
    return nuts_leapfrog(ctx, start, out, step_size, step_size_bcast, dir, energy_baseline, max_energy_error, active, status, energy_error);
  if (start == out) return fail(NUTS_ERR_INVALID, "nuts_leapfrog_kinetic: out must not alias start");
  if (ctx->lr_active)
    return fail(NUTS_ERR_UNSUPPORTED, "nuts_leapfrog_kinetic: ExactNormal / Microcanonical run on the diagonal transformation only");
  const double *dstep, *dbase, *dsn = nullptr, *dcs = nullptr;
  const uint8_t* dm;
  if (kind == NUTS_KINETIC_EXACT_NORMAL) {  // the signed step of every chain exactly as the kernel forms it, then f64::sin / cos on the host
    std::vector<double> eps(ctx->N);
    for (size_t c = 0; c < ctx->N; ++c) eps[c] = (double)(dir ? (int)dir[c] : 1) * (step_size ? step_size[c] : step_size_bcast) * 1.0;
    TRY(upload_sincos(ctx, eps.data(), &dsn, &dcs));
  }
  TRY(upload_f64(ctx, step_size, 0, &dstep));
  TRY(upload_f64(ctx, energy_baseline, 1, &dbase));
  TRY(upload_mask(ctx, active, &dm));
  const int8_t* ddir = nullptr;
  if (dir) {
    CUDA_TRY(cudaMemcpyAsync(ctx->d_i8, dir, ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    ddir = ctx->d_i8;
  }
  CUDA_TRY(cudaMemsetAsync(ctx->d_i32, 0, ctx->N * sizeof(int), ctx->stream));
  if (kind == NUTS_KINETIC_EXACT_NORMAL)
    k_leapfrog_kinetic<1><<<GRID>>>(ctx->row_args(), ctx->model, ctx->T, start->dev, out->dev, dstep, step_size_bcast, ddir, dsn, dcs, dbase,
                                    max_energy_error, dm, ctx->d_i32, ctx->d_sc[2]);
  else
    k_leapfrog_kinetic<2><<<GRID>>>(ctx->row_args(), ctx->model, ctx->T, start->dev, out->dev, dstep, step_size_bcast, ddir, dsn, dcs, dbase,
                                    max_energy_error, dm, ctx->d_i32, ctx->d_sc[2]);
  CHECK_LAUNCH();
  if (status) CUDA_TRY(cudaMemcpyAsync(status, ctx->d_i32, ctx->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(download_f64(ctx, 2, energy_error));
  return sync(ctx);
}
int nuts_leapfrog(nuts_ctx_t* ctx, const nuts_point_t* start, nuts_point_t* out, const double* step_size, double step_size_bcast,
                  const int8_t* dir, const double* energy_baseline, double max_energy_error, const uint8_t* active, int32_t* status,
                  double* energy_error) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (start == out) return fail(NUTS_ERR_INVALID, "nuts_leapfrog: out must not alias start");
  const double *dstep, *dbase;
  const uint8_t* dm;
  TRY(upload_f64(ctx, step_size, 0, &dstep));
  TRY(upload_f64(ctx, energy_baseline, 1, &dbase));
  TRY(upload_mask(ctx, active, &dm));
  const int8_t* ddir = nullptr;
  if (dir) {
    CUDA_TRY(cudaMemcpyAsync(ctx->d_i8, dir, ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    ddir = ctx->d_i8;
  }
  CUDA_TRY(cudaMemsetAsync(ctx->d_i32, 0, ctx->N * sizeof(int), ctx->stream));
  // elementwise targets on the diagonal transformation: input rows staged by the TMA (bit-identical to the register path;
  // NUTS_B200_PLANE_TMA=0 selects the register path)
  const char* tma_env = std::getenv("NUTS_B200_PLANE_TMA");
  const bool tma = !(tma_env && tma_env[0] == '0') && !ctx->lr_active && (ctx->model.kind == LOGP_GAUSS_ISO || ctx->model.kind == LOGP_GAUSS_DIAG);
  CUDA_TRY(cudaEventRecord(ctx->ev_k0, ctx->stream));
  if (tma)
    k_leapfrog_tma<<<GRID>>>(ctx->row_args(), ctx->model, ctx->T, start->dev, out->dev, dstep, step_size_bcast, ddir, dbase, max_energy_error, dm,
                             ctx->d_i32, ctx->d_sc[2]);
  else if (ctx->lr_active)
    k_leapfrog<true><<<GRID>>>(ctx->row_args(), ctx->model, ctx->T, start->dev, out->dev, dstep, step_size_bcast, ddir, dbase, max_energy_error,
                               dm, ctx->d_i32, ctx->d_sc[2]);
  else
    k_leapfrog<false><<<GRID>>>(ctx->row_args(), ctx->model, ctx->T, start->dev, out->dev, dstep, step_size_bcast, ddir, dbase, max_energy_error,
                                dm, ctx->d_i32, ctx->d_sc[2]);
  CHECK_LAUNCH();
  CUDA_TRY(cudaEventRecord(ctx->ev_k1, ctx->stream));
  ctx->ev_valid = true;
  if (status) CUDA_TRY(cudaMemcpyAsync(status, ctx->d_i32, ctx->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(download_f64(ctx, 2, energy_error));
  return sync(ctx);
}
int nuts_is_turning(nuts_ctx_t* ctx, const nuts_point_t* state1, const nuts_point_t* state2, uint8_t* turning) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  k_is_turning<<<GRID>>>(ctx->row_args(), state1->dev, state2->dev, ctx->d_u8);
  CHECK_LAUNCH();
  CUDA_TRY(cudaMemcpyAsync(turning, ctx->d_u8, ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  return sync(ctx);
}

// ============================================================ Tier 3
static int sampler_alloc(nuts_sampler* s, void** p, size_t bytes) {
  CUDA_TRY(cudaMalloc(p, bytes));
  CUDA_TRY(cudaMemset(*p, 0, bytes));
  s->allocations.push_back(*p);
  return NUTS_OK;
}

static int cluster_size_of(int smf) { return (smf & SM_CL4) ? 4 : ((smf & SM_CL2) ? 2 : 1); }

// chain scalars at construction: Strategy::new -> DualAverage::new(initial_step) (stepsize/adapt.rs:67-89), DiagMassMatrix id -1
// (diagonal.rs:81), GlobalStrategy flags (adapt_strategy.rs:87-97)
static ChainState fresh_chain_state(const SettingsDev& S) {
  ChainState c;
  std::memset(&c, 0, sizeof(c));
  c.step_size = 0.0;
  c.pt_transform_id = -1;
  c.mm_id = -1;
  c.da_log_step = std::log(S.initial_step);
  c.da_log_step_adapted = std::log(S.initial_step);
  c.da_hbar = 0.0;
  c.da_mu = S.method == 1 ? 0.0 : std::log(10.0 * S.initial_step);  // Adam::new (adam.rs:57-65): m = v = 0, t = 0
  c.da_count = S.method == 1 ? 0 : 1;
  c.tuning = 1;
  c.has_initial_mass_matrix = 1;
  c.current_window_size = S.mm_switch_freq;
  c.is_good = 1;
  c.alive = 0;
  return c;
}

static int sampler_create_impl(nuts_ctx_t* ctx, nuts_sampler_t** out, const nuts_settings_t* st, uint64_t seed, uint64_t chain_id_offset,
                               uint64_t lowrank_rmax) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!st) return fail(NUTS_ERR_INVALID, "nuts_sampler_create: settings is NULL");
  if (st->trajectory_kind != NUTS_KINETIC_EUCLIDEAN) return fail(NUTS_ERR_UNSUPPORTED, "the whole-draw engines support KineticEnergyKind::Euclidean only (ExactNormal / Microcanonical: Tier 1 / Tier 2, nuts_leapfrog_kinetic)");
  const int method = st->adapt_options.step_size_settings.adapt_options.method;
  if (method != NUTS_STEPSIZE_DUAL_AVERAGE && method != NUTS_STEPSIZE_ADAM && method != NUTS_STEPSIZE_FIXED)
    return fail(NUTS_ERR_INVALID, "unknown step size method %d", method);
  if (st->maxdepth + st->extra_doublings > (uint64_t)MAX_DOUBLING_DEPTH)
    return fail(NUTS_ERR_UNSUPPORTED, "maxdepth + extra_doublings must be <= %d", MAX_DOUBLING_DEPTH);
  if (st->adapt_options.mass_matrix_window_growth < 1.0) return fail(NUTS_ERR_INVALID, "mass_matrix_window_growth must be >= 1");
  const EngineConfig* cfg = nullptr;
  const int s_variant = ctx->model.kind == NUTS_LOGP_GAUSS_RANK1 ? 1 : ctx->model.kind == NUTS_LOGP_FUNNEL ? 2 : ctx->model.kind == NUTS_LOGP_USER ? 3 : 0;
  if (lowrank_rmax > 0) {
    // low-rank mass matrix (SM_LOWRANK builds): the general two-pass leapfrog with the eigenvector reductions
    if (lowrank_rmax > (uint64_t)LR_MAX_RANK) return fail(NUTS_ERR_INVALID, "low-rank sampler: rank_max must be <= %d", LR_MAX_RANK);
    for (const EngineConfig& c : kLowRankConfigs)
      if ((uint64_t)c.max_d >= ctx->d && c.launch[s_variant]) {
        cfg = &c;
        break;
      }
    if (!cfg) return fail(NUTS_ERR_UNSUPPORTED, "no low-rank engine build covers dim %llu (built: up to 1024)", (unsigned long long)ctx->d);
  } else {
  if (const char* env = std::getenv("NUTS_B200_ENGINE")) {
    int tpc = 0, ept = 0, minb = 0;
    if (std::sscanf(env, "%d,%d,%d", &tpc, &ept, &minb) == 3) {
      for (const EngineConfig& c : kConfigs)
        if (c.tpc == tpc && c.ept == ept && c.minb == minb && (uint64_t)c.max_d >= ctx->d) cfg = &c;
      for (const EngineConfig& c : kExtraConfigs)
        if (c.tpc == tpc && c.ept == ept && c.minb == minb && (uint64_t)c.max_d >= ctx->d) cfg = &c;
      if (!cfg) return fail(NUTS_ERR_INVALID, "NUTS_B200_ENGINE=%s is not a built configuration that covers dim %llu", env, (unsigned long long)ctx->d);
    }
  }
  if (!cfg)
    for (const EngineConfig& c : kConfigs)
      if ((uint64_t)c.max_d >= ctx->d) {
        cfg = &c;
        break;
      }
  if (!cfg) return fail(NUTS_ERR_UNSUPPORTED, "dim %llu exceeds the largest register-resident configuration (16384)", (unsigned long long)ctx->d);
  if (!std::getenv("NUTS_B200_ENGINE")) {
    // dim ~ 10^4: a chain's (z, v, sigma, mu) plus the redundant scalar state of the register-resident engine no longer fit one SM
    // (1024x10 runs at 64 registers per thread and spills 2.5 KB); the decoupled engine keeps the scalar state in ONE leader warp
    for (const EngineConfig& c : kDecoupledLarge)
      if (ctx->model.kind != NUTS_LOGP_GAUSS_RANK1 && ctx->model.kind != NUTS_LOGP_FUNNEL && ctx->model.kind != NUTS_LOGP_USER && c.launch[0] && ctx->d > (uint64_t)c.min_d &&
          ctx->d <= (uint64_t)c.max_d && st->maxdepth + st->extra_doublings <= (uint64_t)V2_MAXD + 1) {
        cfg = &c;
        break;
      }
    // rank-1 / funnel / user targets at dim ~ 10^4: the register-resident single-CTA tilings (1024 x 8 / 10) run at 64 registers per
    // thread and spill 2.5 KB; the cluster engine spreads the chain over the four SMs of a thread-block cluster (255 registers)
    if (ctx->model.kind != NUTS_LOGP_GAUSS_ISO && ctx->model.kind != NUTS_LOGP_GAUSS_DIAG && ctx->d > 4096 &&
        ctx->d <= (uint64_t)kClusterLarge.max_d && kClusterLarge.launch[s_variant])
      cfg = &kClusterLarge;
    // warp tilings: the aligned build.  Measured on diagonal / rank-1 Gaussians (dim 20 ... 500): tuning +23 ... +81 %, sampling
    // +3 ... +30 % (instruction-fetch stalls were 21 % / 65 % of the samples); the funnel, whose trees range from 1 to 1023
    // leapfrogs, loses 2 ... 8 % to the waiting and keeps the unaligned tiling.
    {
      const char* env = std::getenv("NUTS_B200_ALIGN");
      const bool align = env ? std::atoi(env) != 0 : ctx->model.kind != NUTS_LOGP_FUNNEL;
      if (align)
        for (const EngineConfig& c : kAlignedConfigs)
          if (c.tpc == cfg->tpc && c.ept == cfg->ept && c.launch[s_variant]) cfg = &c;
    }
    // same tiling without bounds checks (SM_EXACT: rows zero-padded to tpc*ept) when the padding costs at most 7 % more traffic
    for (const EngineConfig& c : kExactConfigs)
      if (c.tpc == cfg->tpc && c.ept == cfg->ept && c.launch[0] && (uint64_t)c.max_d - ctx->d <= ctx->d * 7 / 100) cfg = &c;
  }

  }
  nuts_sampler* s = new nuts_sampler();
  Guard<nuts_sampler, nuts_sampler_destroy> guard(s);  // every early return below releases what was allocated so far
  s->ctx = ctx;
  s->settings = *st;
  s->cfg = cfg;
  EngineParams& P = s->P;
  P.N = (int)ctx->N;
  P.d = (int)ctx->d;
  int blocks_per_sm = 0, cta_threads = 0, smf = 0;
  s->model_variant = ctx->model.kind == NUTS_LOGP_GAUSS_RANK1 ? 1 : ctx->model.kind == NUTS_LOGP_FUNNEL ? 2 : ctx->model.kind == NUTS_LOGP_USER ? 3 : 0;
  if (!cfg->launch[s->model_variant] || !cfg->occupancy[s->model_variant]) {
    return fail(NUTS_ERR_UNSUPPORTED, "engine %dx%d for model kind %d is not part of this build", cfg->tpc, cfg->ept, ctx->model.kind);
  }
  CUDA_TRY(cfg->occupancy[s->model_variant](&blocks_per_sm, &cta_threads, &smf));
  const int blocks_per_sm_raw = blocks_per_sm;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  // SM_EXACT engines run their hot loops without bounds checks: every sampler row is zero-padded to tpc*ept elements
  P.ld = (smf & SM_EXACT) ? (int)std::max<uint64_t>(ctx->ld, (uint64_t)cfg->tpc * cfg->ept) : (int)ctx->ld;
  const bool decoupled = cfg->minb >= 100;  // chain_engine_v2.cuh
  if (decoupled && st->maxdepth + st->extra_doublings > (uint64_t)V2_MAXD + 1) {
    return fail(NUTS_ERR_UNSUPPORTED, "the decoupled engine needs maxdepth + extra_doublings <= %d", V2_MAXD + 1);
  }
  // checkpoint pool: 3 roles per pending level + the main tree's draw and its two ends; the decoupled engine hands slots out V2_K
  // leaves ahead
  P.P = (int)std::min<uint64_t>(MAX_SLOTS, 3 * (st->maxdepth + st->extra_doublings) + 6 + (decoupled ? V2_K : 0));
  P.model = ctx->model;
  P.seed = seed;
  P.chain_offset = chain_id_offset;
  SettingsDev& S = P.s;
  S.num_tune = st->num_tune;
  S.maxdepth = st->maxdepth;
  S.mindepth = st->mindepth;
  S.extra_doublings = st->extra_doublings;
  S.max_energy_error = st->max_energy_error;
  S.check_turning = st->check_turning;
  S.has_target_time = st->has_target_integration_time;
  S.target_time = st->target_integration_time;
  const nuts_euclidean_adapt_options_t& a = st->adapt_options;
  S.target_accept = a.step_size_settings.target_accept;
  S.initial_step = a.step_size_settings.initial_step;
  S.has_jitter = a.step_size_settings.has_jitter;
  S.jitter = a.step_size_settings.jitter;
  S.method = method;
  S.adam_beta1 = a.step_size_settings.adapt_options.adam.beta1;
  S.adam_beta2 = a.step_size_settings.adapt_options.adam.beta2;
  S.adam_epsilon = a.step_size_settings.adapt_options.adam.epsilon;
  S.adam_lr = a.step_size_settings.adapt_options.adam.learning_rate;
  S.fixed_step = a.step_size_settings.adapt_options.fixed_step;
  S.da_k = a.step_size_settings.adapt_options.dual_average.k;
  S.da_t0 = a.step_size_settings.adapt_options.dual_average.t0;
  S.da_gamma = a.step_size_settings.adapt_options.dual_average.gamma;
  S.da_max_step = a.step_size_settings.adapt_options.dual_average.max_step_size;
  S.use_grad_based = a.mass_matrix_options.use_grad_based_estimate;
  // GlobalStrategy::new (adapt_strategy.rs:77-98)
  const double num_tune_f = (double)st->num_tune;
  const uint64_t step_size_window = (uint64_t)(a.step_size_window * num_tune_f);
  S.early_end = (uint64_t)(a.early_window * num_tune_f);
  S.final_step_size_window = st->num_tune >= step_size_window ? st->num_tune - step_size_window : 0;
  if (st->num_tune > 0 && !(S.early_end < st->num_tune)) {
    return fail(NUTS_ERR_INVALID, "early_window must leave early_end < num_tune");
  }
  S.mm_switch_freq = a.mass_matrix_switch_freq;
  S.early_mm_switch_freq = a.early_mass_matrix_switch_freq;
  S.mm_update_freq = a.mass_matrix_update_freq;
  S.mm_window_growth = a.mass_matrix_window_growth;

  // persistent grid: one wave of resident CTAs
  const int cluster = cluster_size_of(smf);  // CTAs per team (thread-block cluster engines), else 1
  const int teams_per_cta = cluster > 1 ? 1 : (decoupled ? cfg->minb % 10 : cta_threads / cfg->tpc);  // decoupled tags end in the number of teams
  uint64_t resident_teams;
  if (cluster > 1) {
    // the occupancy hook of a cluster engine reports MINUS the number of clusters the device can hold at once
    resident_teams = std::min<uint64_t>(ctx->N, (uint64_t)std::max(1, -blocks_per_sm_raw));
    if (const char* env = std::getenv("NUTS_B200_GRID")) resident_teams = std::max<uint64_t>(1, std::min<uint64_t>(resident_teams, std::atoi(env)));
    s->grid = (int)resident_teams * cluster;
  } else {
    const uint64_t ctas_needed = (ctx->N + teams_per_cta - 1) / teams_per_cta;
    s->grid = (int)std::min<uint64_t>(ctas_needed, (uint64_t)blocks_per_sm * ctx->num_sms);
    // NUTS_B200_GRID caps the persistent grid (tests / compute-sanitizer: forces draw migration between teams on small workloads)
    if (const char* env = std::getenv("NUTS_B200_GRID")) s->grid = std::max(1, std::min(s->grid, std::atoi(env)));
    resident_teams = (uint64_t)s->grid * teams_per_cta;
  }
  s->teams_per_cta = teams_per_cta;
  s->resident_teams = resident_teams;
  if (cfg == &kExactConfigs[0] && lowrank_rmax == 0 && kExactTuneConfig.launch[s->model_variant] && !std::getenv("NUTS_B200_GRID") &&
      !(std::getenv("NUTS_B200_TUNE_ENGINE") && std::atoi(std::getenv("NUTS_B200_TUNE_ENGINE")) == 0)) {
    int b2 = 0, t2 = 0, f2 = 0;
    CUDA_TRY(kExactTuneConfig.occupancy[s->model_variant](&b2, &t2, &f2));
    const int teams2 = t2 / kExactTuneConfig.tpc;
    const uint64_t ctas2 = (ctx->N + teams2 - 1) / teams2;
    const int grid2 = (int)std::min<uint64_t>(ctas2, (uint64_t)std::max(b2, 0) * ctx->num_sms);
    // same pools, same padded rows: only usable when it keeps exactly the resident teams the memory was sized for
    if (b2 >= 1 && (f2 & SM_EXACT) && (uint64_t)grid2 * teams2 == resident_teams && ctx->N > resident_teams) {
      s->cfg_tune = &kExactTuneConfig;
      s->grid_tune = grid2;
    }
  }
  const size_t plane = ctx->N * (size_t)P.ld * sizeof(double);
  const size_t team_plane = resident_teams * (size_t)P.ld * sizeof(double);  // one row per resident team
  int r = NUTS_OK;
  auto A = [&](void** p, size_t bytes) {
    if (r == NUTS_OK) r = sampler_alloc(s, p, bytes);
  };
  A((void**)&P.x, plane);
  A((void**)&P.gx, plane);
  A((void**)&P.z, plane);
  A((void**)&P.gz, plane);
  A((void**)&P.v0, plane);
  A((void**)&P.stds, plane);
  A((void**)&P.inv_stds, plane);
  A((void**)&P.mean, plane);
  A((void**)&P.est, plane * 8);
  A((void**)&P.slots, team_plane * 2 * (size_t)P.P);  // checkpoints live inside one draw on one team: pools per TEAM, not per chain
  A((void**)&P.ends, team_plane * 3 * NB_END_BUFFERS);
  A((void**)&P.cs, ctx->N * sizeof(ChainState));
  // ready queue of the work units (chain_engine.cuh, unit_pop / unit_push): two tickets, the per-chain unit counters, and a ring
  // of a power of two >= N slots whose sequence numbers are reset from a template before every launch
  size_t ring = 1;
  while (ring < (size_t)ctx->N) ring *= 2;
  P.ring_mask = (unsigned)(ring - 1);
  A((void**)&P.queue, 2 * sizeof(unsigned int));
  A((void**)&P.done, ctx->N * sizeof(unsigned int));
  A((void**)&P.ring_seq, ring * sizeof(unsigned int));
  A((void**)&P.ring_chain, ring * sizeof(unsigned int));
  A((void**)&s->d_ring_template, ring * sizeof(unsigned int));
  A((void**)&s->d_init, ctx->N * ctx->d * sizeof(double));
  A((void**)&s->d_status, ctx->N * sizeof(int));
  A((void**)&P.phase_clocks, 16 * sizeof(unsigned long long));
  if (lowrank_rmax > 0) {
    A((void**)&s->lr_vecs, ctx->N * lowrank_rmax * (size_t)P.ld * sizeof(double));
    A((void**)&s->lr_vals_sqrt, ctx->N * lowrank_rmax * sizeof(double));
    A((void**)&s->lr_vals_sqrt_inv, ctx->N * lowrank_rmax * sizeof(double));
    A((void**)&s->lr_mu, plane);
    A((void**)&s->lr_rank, ctx->N * sizeof(int));
  }
  if (r != NUTS_OK) return r;
  if (lowrank_rmax > 0) {
    CUDA_TRY(cudaMemset(s->lr_rank, 0xff, ctx->N * sizeof(int)));  // -1: no low-rank part until the first update
    P.lr_vecs = s->lr_vecs;
    P.lr_vals_sqrt = s->lr_vals_sqrt;
    P.lr_vals_sqrt_inv = s->lr_vals_sqrt_inv;
    P.lr_mu = s->lr_mu;
    P.lr_rank = s->lr_rank;
    P.lr_rmax = (int)lowrank_rmax;
    s->lowrank_rmax = lowrank_rmax;
  }
  std::vector<ChainState> cs(ctx->N, fresh_chain_state(S));
  CUDA_TRY(cudaMemcpy(P.cs, cs.data(), cs.size() * sizeof(ChainState), cudaMemcpyHostToDevice));
  {
    std::vector<unsigned int> seq(ring);
    for (size_t i = 0; i < ring; ++i) seq[i] = (unsigned)i;
    CUDA_TRY(cudaMemcpy(s->d_ring_template, seq.data(), ring * sizeof(unsigned int), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaEventCreate(&s->ev0));
  CUDA_TRY(cudaEventCreate(&s->ev1));
  guard.dismiss();
  *out = s;
  return NUTS_OK;
}

int nuts_sampler_create(nuts_ctx_t* ctx, nuts_sampler_t** out, const nuts_settings_t* st, uint64_t seed, uint64_t chain_id_offset) {
  return sampler_create_impl(ctx, out, st, seed, chain_id_offset, 0);
}
int nuts_sampler_create_lowrank(nuts_ctx_t* ctx, nuts_sampler_t** out, const nuts_settings_t* st, uint64_t seed, uint64_t chain_id_offset,
                                uint64_t rank_max) {
  if (rank_max == 0) return fail(NUTS_ERR_INVALID, "nuts_sampler_create_lowrank: rank_max must be >= 1");
  return sampler_create_impl(ctx, out, st, seed, chain_id_offset, rank_max);
}

int nuts_sampler_destroy(nuts_sampler_t* s) {
  if (!s) return NUTS_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  for (void* p : s->allocations) cudaFree(p);
  auto F = [](void* p) {
    if (p) cudaFree(p);
  };
  F(s->d_stats_blob);
  if (s->h_stats_blob) cudaFreeHost(s->h_stats_blob);
  F(s->d_draws);
  if (s->h_pinned) cudaFreeHost(s->h_pinned);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
  return NUTS_OK;
}

static int launch_engine(nuts_sampler* s, bool tuning_build = false) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaMemsetAsync(s->P.queue, 0, 2 * sizeof(unsigned int), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(s->P.done, 0, ctx->N * sizeof(unsigned int), ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(s->P.ring_seq, s->d_ring_template, ((size_t)s->P.ring_mask + 1) * sizeof(unsigned int), cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(cudaEventRecord(s->ev0, ctx->stream));
  if (tuning_build && s->cfg_tune) CUDA_TRY(s->cfg_tune->launch[s->model_variant](&s->P, s->grid_tune, ctx->stream));
  else CUDA_TRY(s->cfg->launch[s->model_variant](&s->P, s->grid, ctx->stream));
  CUDA_TRY(cudaEventRecord(s->ev1, ctx->stream));
  s->last_launches += 1;
  return NUTS_OK;
}

int nuts_set_position(nuts_sampler_t* s, const double* position, int32_t* status) { return nuts_set_position_masked(s, position, nullptr, status); }

int nuts_set_position_masked(nuts_sampler_t* s, const double* position, const uint8_t* mask, int32_t* status) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!position) return fail(NUTS_ERR_INVALID, "nuts_set_position: position is NULL");
  if (mask && !s->positioned) return fail(NUTS_ERR_INVALID, "nuts_set_position_masked: call nuts_set_position for all chains first");
  CUDA_TRY(cudaMemcpyAsync(s->d_init, position, ctx->N * ctx->d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  s->P.init_mask = nullptr;
  if (mask) {
    CUDA_TRY(cudaMemcpyAsync(ctx->d_u8, mask, ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    s->P.init_mask = ctx->d_u8;
    // a chain that is initialised again starts from a fresh NutsChain (Settings::new_chain, sampler.rs:745-772): reset its record
    std::vector<ChainState> cs(ctx->N);
    CUDA_TRY(cudaMemcpyAsync(cs.data(), s->P.cs, cs.size() * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync(ctx));
    for (uint64_t c = 0; c < ctx->N; ++c)
      if (mask[c]) {
        const uint64_t rng = cs[c].rng_counter;  // the chain's random stream goes on (the reference re-draws from the same rng)
        cs[c] = fresh_chain_state(s->P.s);
        cs[c].rng_counter = rng;
      }
    CUDA_TRY(cudaMemcpyAsync(s->P.cs, cs.data(), cs.size() * sizeof(ChainState), cudaMemcpyHostToDevice, ctx->stream));
    if (status) CUDA_TRY(cudaMemcpyAsync(s->d_status, status, ctx->N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  s->P.mode = 0;
  s->P.init_position = s->d_init;
  s->P.status_out = s->d_status;
  s->P.n_draws = 0;
  s->P.draws_per_unit = 1;
  s->P.draws_out = nullptr;
  s->P.stats = StatsDev{};
  s->last_launches = 0;
  TRY(launch_engine(s));
  if (status) CUDA_TRY(cudaMemcpyAsync(status, s->d_status, ctx->N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  s->last_kernel_ms = ms;
  s->positioned = true;
  return NUTS_OK;
}

// element size of statistic k, in nuts_stats_t / StatsDev member order
static const size_t kStatElem[15] = {8, 1, 8, 8, 8, 8, 1, 8, 8, 8, 8, 8, 8, 1, 8};

static int ensure_stats(nuts_sampler* s, uint64_t n_draws) {
  if (n_draws <= s->stats_capacity) return NUTS_OK;
  const size_t n = n_draws * s->ctx->N;
  s->stats_capacity = 0;  // nothing usable until the regrow below has succeeded
  s->d_stats = StatsDev{};
  if (s->d_stats_blob) CUDA_TRY(cudaFree(s->d_stats_blob));
  s->d_stats_blob = nullptr;
  if (s->h_stats_blob) CUDA_TRY(cudaFreeHost(s->h_stats_blob));
  s->h_stats_blob = nullptr;
  size_t off = 0;
  for (int k = 0; k < 15; ++k) {
    s->stats_offset[k] = off;
    off += (n * kStatElem[k] + 255) / 256 * 256;
  }
  s->stats_blob_bytes = off;
  CUDA_TRY(cudaMalloc((void**)&s->d_stats_blob, off));
  CUDA_TRY(cudaMemset(s->d_stats_blob, 0, off));
  CUDA_TRY(cudaHostAlloc((void**)&s->h_stats_blob, off, cudaHostAllocDefault));
  unsigned char* b = s->d_stats_blob;
  StatsDev& d = s->d_stats;
  d.depth = (uint64_t*)(b + s->stats_offset[0]);
  d.maxdepth_reached = (uint8_t*)(b + s->stats_offset[1]);
  d.index_in_trajectory = (long long*)(b + s->stats_offset[2]);
  d.logp = (double*)(b + s->stats_offset[3]);
  d.energy = (double*)(b + s->stats_offset[4]);
  d.energy_error = (double*)(b + s->stats_offset[5]);
  d.diverging = (uint8_t*)(b + s->stats_offset[6]);
  d.step_size = (double*)(b + s->stats_offset[7]);
  d.step_size_bar = (double*)(b + s->stats_offset[8]);
  d.mean_tree_accept = (double*)(b + s->stats_offset[9]);
  d.mean_tree_accept_sym = (double*)(b + s->stats_offset[10]);
  d.n_steps = (uint64_t*)(b + s->stats_offset[11]);
  d.max_energy_error = (double*)(b + s->stats_offset[12]);
  d.tuning = (uint8_t*)(b + s->stats_offset[13]);
  d.fisher_distance = (double*)(b + s->stats_offset[14]);
  s->stats_capacity = n_draws;
  return NUTS_OK;
}

static int run_draws(nuts_sampler* s, uint64_t n_draws, double* draws_dev, bool want_stats) {
  nuts_ctx* ctx = s->ctx;
  if (!s->positioned) return fail(NUTS_ERR_INVALID, "nuts_draw: call nuts_set_position first");
  if (n_draws * ctx->N >= (1ull << 31)) return fail(NUTS_ERR_INVALID, "nuts_draw: n_draws x chains must stay below 2^31 per call; draw in batches");
  if (want_stats) TRY(ensure_stats(s, n_draws));
  s->P.mode = 1;
  s->P.init_position = nullptr;
  s->P.status_out = nullptr;
  s->P.n_draws = n_draws;
  bool tuning_build = false;
  {
    // draws per work unit: the whole call when every chain has its own team (no hand-over at all), else one draw - or a few
    // for tiny dims; NUTS_B200_DRAWS_PER_UNIT overrides (experiments)
    const uint64_t teams = s->resident_teams;
    // (measured: blocks of draws pay for tiny dims, where the hand-over is a large part of a draw: config 3 +4 % sampling, +10 %
    // tuning; config 5 shard (dim 100) -5 %; config 2 +-0)
    uint64_t b = ctx->N <= teams ? n_draws : (ctx->d <= 32 ? std::max<uint64_t>(1, std::min<uint64_t>(8, n_draws / 8)) : 1);
    // aligned warp tilings after the warm-up: two draws per unit halve the fences / queue traffic / alignment barriers of the
    // hand-over (config 5 sampling +7 %, dim 200 +4 %); during the warm-up one draw per unit stays better (-5 ... -9 % with two:
    // the trees still differ in length and an aligned CTA waits for its longest unit)
    if (b == 1 && ctx->N > teams && s->cfg->tpc == 32 && s->cfg->minb == 21 && s->draws_done >= s->settings.num_tune && n_draws >= 2) b = 2;
    // a launch inside the warm-up on the aligned warm-up build (kExactTuneConfig): two draws per unit there as well
    tuning_build = s->cfg_tune != nullptr && s->draws_done + n_draws <= s->settings.num_tune;
    if (tuning_build && b == 1 && n_draws >= 2) b = 2;
    if (const char* env = std::getenv("NUTS_B200_DRAWS_PER_UNIT")) b = std::max<uint64_t>(1, std::strtoull(env, nullptr, 10));
    s->P.draws_per_unit = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(b, 1), std::max<uint64_t>(n_draws, 1));
  }
  s->P.draws_out = draws_dev;
  s->P.stats = want_stats ? s->d_stats : StatsDev{};
  TRY(launch_engine(s, tuning_build));
  s->draws_done += n_draws;
  (void)ctx;
  return NUTS_OK;
}

int nuts_draw_device(nuts_sampler_t* s, uint64_t n_draws, double* draws_dev) {
  CUDA_TRY(cudaSetDevice(s->ctx->device));
  s->last_launches = 0;
  // the statistics are produced like in nuts_draw (the whole output of Chain::draw); they stay in the sampler's device buffers
  return run_draws(s, n_draws, draws_dev, true);
}

int nuts_draw(nuts_sampler_t* s, uint64_t n_draws, double* draws_out, const nuts_stats_t* stats) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (n_draws == 0) return NUTS_OK;
  s->last_launches = 0;
  const size_t per_draw = ctx->N * ctx->d;
  // Where the kernel writes the draws.  A pinned (page-locked, device-mapped) host buffer or a device buffer is written
  // DIRECTLY by the kernel: 8*d bytes per draw and chain leave as coalesced posted writes over PCIe / into HBM while the
  // chains keep stepping, so there is no device staging buffer and no serial D2H copy after the kernel.  Pageable host
  // memory goes through a device buffer and one cudaMemcpyAsync.  NUTS_B200_STAGED_D2H=1 forces the staged path (A/B timing).
  double* direct = nullptr;
  if (draws_out && !std::getenv("NUTS_B200_STAGED_D2H")) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, draws_out) == cudaSuccess) {
      if ((attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged) && attr.devicePointer) direct = (double*)attr.devicePointer;
      else if (attr.type == cudaMemoryTypeDevice && attr.device == ctx->device) direct = draws_out;
      else if (attr.type == cudaMemoryTypeDevice)
        return fail(NUTS_ERR_INVALID, "nuts_draw: draws_out is memory of device %d, the sampler runs on device %d", attr.device, ctx->device);
    } else {
      cudaGetLastError();
    }
  }
  if (draws_out && !direct && n_draws > s->draws_capacity) {
    TRY(grow(&s->d_draws, n_draws * per_draw));
    s->draws_capacity = n_draws;
  }
  TRY(run_draws(s, n_draws, draws_out ? (direct ? direct : s->d_draws) : nullptr, stats != nullptr));
  s->last_direct = direct != nullptr;
  if (draws_out && !direct)
    CUDA_TRY(cudaMemcpyAsync(draws_out, s->d_draws, n_draws * per_draw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (stats) {  // into the page-locked mirror: real asynchronous DMA, no pageable staging by the driver
    const size_t n = n_draws * ctx->N;
    for (int k = 0; k < 15; ++k)
      CUDA_TRY(cudaMemcpyAsync(s->h_stats_blob + s->stats_offset[k], s->d_stats_blob + s->stats_offset[k], n * kStatElem[k],
                               cudaMemcpyDeviceToHost, ctx->stream));
  }
  TRY(sync(ctx));
  if (stats) {
    const size_t n = n_draws * ctx->N;
    void* const dst[15] = {stats->depth, stats->maxdepth_reached, stats->index_in_trajectory, stats->logp, stats->energy,
                           stats->energy_error, stats->diverging, stats->step_size, stats->step_size_bar, stats->mean_tree_accept,
                           stats->mean_tree_accept_sym, stats->n_steps, stats->max_energy_error, stats->tuning, stats->fisher_distance};
    for (int k = 0; k < 15; ++k)
      if (dst[k]) std::memcpy(dst[k], s->h_stats_blob + s->stats_offset[k], n * kStatElem[k]);
  }
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  s->last_kernel_ms = ms;
  return NUTS_OK;
}

int nuts_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return fail(NUTS_ERR_INVALID, "nuts_host_alloc: ptr is NULL");
  TRY(check_device());
  CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
  return NUTS_OK;
}

int nuts_host_free(void* ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return NUTS_OK;
}

int nuts_sampler_last_draw_direct(nuts_sampler_t* s, int32_t* direct) {
  if (direct) *direct = s->last_direct ? 1 : 0;
  return NUTS_OK;
}

int nuts_sampler_counters(nuts_sampler_t* s, uint64_t* total_leapfrogs, uint64_t* draws_done) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<ChainState> cs(ctx->N);
  CUDA_TRY(cudaMemcpyAsync(cs.data(), s->P.cs, cs.size() * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  uint64_t tot = 0;
  for (auto& c : cs) tot += c.total_leapfrogs;
  if (total_leapfrogs) *total_leapfrogs = tot;
  if (draws_done) *draws_done = s->draws_done;
  return NUTS_OK;
}

int nuts_sampler_last_timing(nuts_sampler_t* s, double* kernel_ms, uint64_t* launches) {
  CUDA_TRY(cudaSetDevice(s->ctx->device));
  if (s->ev1) {
    CUDA_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last_kernel_ms = ms;
  }
  if (kernel_ms) *kernel_ms = s->last_kernel_ms;
  if (launches) *launches = s->last_launches;
  return NUTS_OK;
}

int nuts_sampler_get_state(nuts_sampler_t* s, double* position, double* step_size, double* stds, double* mean, uint64_t* rng_counter) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (position) TRY(plane_to_host(ctx, s->P.x, position, s->P.ld));
  if (stds) TRY(plane_to_host(ctx, s->P.stds, stds, s->P.ld));
  if (mean) TRY(plane_to_host(ctx, s->P.mean, mean, s->P.ld));
  if (step_size || rng_counter) {
    std::vector<ChainState> cs(ctx->N);
    CUDA_TRY(cudaMemcpyAsync(cs.data(), s->P.cs, cs.size() * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync(ctx));
    for (uint64_t c = 0; c < ctx->N; ++c) {
      if (step_size) step_size[c] = cs[c].step_size;
      if (rng_counter) rng_counter[c] = cs[c].rng_counter;
    }
  }
  return NUTS_OK;
}

// The whole state a chain carries from one draw to the next (include/nuts_b200.h nuts_chain_state_t).  Between two launches all
// of it lives in global memory: the planes of EngineParams + one ChainState record per chain.
int nuts_sampler_get_chain_state(nuts_sampler_t* s, const nuts_chain_state_t* o) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!o) return fail(NUTS_ERR_INVALID, "nuts_sampler_get_chain_state: out is NULL");
  const EngineParams& P = s->P;
  const size_t N = ctx->N;
  std::vector<ChainState> cs(N);
  CUDA_TRY(cudaMemcpyAsync(cs.data(), P.cs, N * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  auto plane = [&](const double* src, double* dst) { return dst ? plane_to_host(ctx, src, dst, P.ld) : NUTS_OK; };
  TRY(plane(P.x, o->position));
  TRY(plane(P.gx, o->gradient));
  TRY(plane(P.z, o->transformed_position));
  TRY(plane(P.gz, o->transformed_gradient));
  TRY(plane(P.stds, o->stds));
  TRY(plane(P.inv_stds, o->inv_stds));
  TRY(plane(P.mean, o->mean));
  // estimators: est[chain][set][4][ld]; the foreground set is chain-dependent (a switch flips the index), so gather per chain
  double* est_out[2][4] = {{o->draw_mean, o->draw_var, o->grad_mean, o->grad_var},
                           {o->draw_mean_bg, o->draw_var_bg, o->grad_mean_bg, o->grad_var_bg}};
  for (int bg = 0; bg < 2; ++bg)
    for (int w = 0; w < 4; ++w)
      if (est_out[bg][w])
        for (size_t c = 0; c < N; ++c) {
          const int set = bg ? 1 - cs[c].fg_set : cs[c].fg_set;
          // an estimator without samples is a fresh RunningVariance (zeros, transform/adapt/diagonal.rs:24-30): after a window switch the
          // engine just resets the count, the planes are overwritten by the first sample
          if ((bg ? cs[c].bg_count : cs[c].fg_count) == 0) {
            std::memset(est_out[bg][w] + c * ctx->d, 0, ctx->d * sizeof(double));
            continue;
          }
          CUDA_TRY(cudaMemcpyAsync(est_out[bg][w] + c * ctx->d, P.est + ((c * 2 + set) * 4 + w) * (size_t)P.ld, ctx->d * sizeof(double),
                                   cudaMemcpyDeviceToHost, ctx->stream));
        }
  TRY(sync(ctx));
  for (size_t c = 0; c < N; ++c) {
    const ChainState& k = cs[c];
#define PUT(field, val) \
  if (o->field) o->field[c] = (val)
    PUT(logp, k.logp);
    PUT(point_logdet, k.pt_logdet);
    PUT(point_transform_id, k.pt_transform_id);
    PUT(mass_matrix_logdet, k.mm_logdet);
    PUT(mass_matrix_id, k.mm_id);
    PUT(step_size, k.step_size);
    PUT(da_log_step, k.da_log_step);
    PUT(da_log_step_adapted, k.da_log_step_adapted);
    PUT(da_hbar, k.da_hbar);
    PUT(da_mu, k.da_mu);
    PUT(da_count, k.da_count);
    PUT(foreground_count, k.fg_count);
    PUT(background_count, k.bg_count);
    PUT(tuning, (uint8_t)(k.tuning != 0));
    PUT(has_initial_mass_matrix, (uint8_t)(k.has_initial_mass_matrix != 0));
    PUT(last_update, k.last_update);
    PUT(current_window_size, k.current_window_size);
    PUT(draw_count, k.draw_count);
    PUT(rng_counter, k.rng_counter);
    PUT(total_leapfrogs, k.total_leapfrogs);
    PUT(alive, (uint8_t)(k.alive != 0));
#undef PUT
  }
  return NUTS_OK;
}

int nuts_sampler_set_chain_state(nuts_sampler_t* s, const nuts_chain_state_t* in) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!in) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_chain_state: in is NULL");
  const void* const* members = reinterpret_cast<const void* const*>(in);
  for (size_t k = 0; k < sizeof(nuts_chain_state_t) / sizeof(void*); ++k)
    if (!members[k]) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_chain_state: member %zu of nuts_chain_state_t is NULL", k);
  EngineParams& P = s->P;
  const size_t N = ctx->N;
  RowArgs ra = ctx->row_args();
  ra.ld = P.ld;  // rows keep their zero padding: k_pack only writes the first d elements
  auto plane = [&](const double* src, double* dst) {
    CUDA_TRY(cudaMemcpyAsync(ctx->d_dense, src, N * ctx->d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_pack<<<(unsigned)N, PK_THREADS, 0, ctx->stream>>>(ra, ctx->d_dense, dst);
    CHECK_LAUNCH();
    return sync(ctx);
  };
  TRY(plane(in->position, P.x));
  TRY(plane(in->gradient, P.gx));
  TRY(plane(in->transformed_position, P.z));
  TRY(plane(in->transformed_gradient, P.gz));
  TRY(plane(in->stds, P.stds));
  TRY(plane(in->inv_stds, P.inv_stds));
  TRY(plane(in->mean, P.mean));
  std::vector<ChainState> cs(N);
  CUDA_TRY(cudaMemcpyAsync(cs.data(), P.cs, N * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  const double* est_in[2][4] = {{in->draw_mean, in->draw_var, in->grad_mean, in->grad_var},
                                {in->draw_mean_bg, in->draw_var_bg, in->grad_mean_bg, in->grad_var_bg}};
  for (size_t c = 0; c < N; ++c) {
    ChainState& k = cs[c];
    k.fg_set = 0;  // restored estimators: foreground in set 0, background in set 1
    for (int bg = 0; bg < 2; ++bg)
      for (int w = 0; w < 4; ++w)
        CUDA_TRY(cudaMemcpyAsync(P.est + ((c * 2 + bg) * 4 + w) * (size_t)P.ld, est_in[bg][w] + c * ctx->d, ctx->d * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
    k.logp = in->logp[c];
    k.pt_logdet = in->point_logdet[c];
    k.pt_transform_id = in->point_transform_id[c];
    k.mm_logdet = in->mass_matrix_logdet[c];
    k.mm_id = in->mass_matrix_id[c];
    k.step_size = in->step_size[c];
    k.da_log_step = in->da_log_step[c];
    k.da_log_step_adapted = in->da_log_step_adapted[c];
    k.da_hbar = in->da_hbar[c];
    k.da_mu = in->da_mu[c];
    k.da_count = in->da_count[c];
    k.fg_count = in->foreground_count[c];
    k.bg_count = in->background_count[c];
    k.tuning = in->tuning[c] ? 1 : 0;
    k.has_initial_mass_matrix = in->has_initial_mass_matrix[c] ? 1 : 0;
    k.last_update = in->last_update[c];
    k.current_window_size = in->current_window_size[c];
    k.draw_count = in->draw_count[c];
    k.rng_counter = in->rng_counter[c];
    k.total_leapfrogs = in->total_leapfrogs[c];
    k.alive = in->alive[c] ? 1 : 0;
  }
  CUDA_TRY(cudaMemcpyAsync(P.cs, cs.data(), N * sizeof(ChainState), cudaMemcpyHostToDevice, ctx->stream));
  TRY(sync(ctx));
  s->positioned = true;
  s->draws_done = 0;
  for (size_t c = 0; c < N; ++c) s->draws_done = std::max<uint64_t>(s->draws_done, in->draw_count[c]);
  return NUTS_OK;
}

// ============================================================ multi-GPU gather of the draws (NCCL, loaded at run time)
namespace {
struct NcclId {  // ncclUniqueId: 128 opaque bytes, passed by value
  char internal[128];
};
struct NcclApi {  // the five entry points used, with their (stable) C signatures
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int load_nccl() {
  if (g_nccl.handle) return NUTS_OK;
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(NUTS_ERR_UNSUPPORTED, "libnccl.so.2 not found (%s): the draw gather needs NCCL", dlerror());
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy) return fail(NUTS_ERR_UNSUPPORTED, "libnccl lacks an expected symbol");
  g_nccl = a;
  return NUTS_OK;
}
#define NCCL_TRY(expr)                                                                                                         \
  do {                                                                                                                         \
    int _r = (expr);                                                                                                           \
    if (_r != 0) return fail(NUTS_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); \
  } while (0)
}  // namespace

struct nuts_comm {
  void* comm = nullptr;
  int device = 0, nranks = 1, rank = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ready = nullptr, t0 = nullptr, t1 = nullptr;
  bool in_flight = false;
};

int nuts_comm_unique_id(uint8_t id[128]) {
  TRY(load_nccl());
  NcclId u;
  NCCL_TRY(g_nccl.GetUniqueId(&u));
  std::memcpy(id, u.internal, 128);
  return NUTS_OK;
}
int nuts_comm_create(nuts_comm_t** out, int device_id, const uint8_t id[128], int nranks, int rank) {
  TRY(check_device());
  TRY(load_nccl());
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(NUTS_ERR_INVALID, "nuts_comm_create: bad arguments");
  CUDA_TRY(cudaSetDevice(device_id));
  nuts_comm* c = new nuts_comm();
  Guard<nuts_comm, nuts_comm_destroy> guard(c);
  c->device = device_id;
  c->nranks = nranks;
  c->rank = rank;
  NcclId u;
  std::memcpy(u.internal, id, 128);
  NCCL_TRY(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreate(&c->t0));
  CUDA_TRY(cudaEventCreate(&c->t1));
  guard.dismiss();
  *out = c;
  return NUTS_OK;
}
int nuts_comm_destroy(nuts_comm_t* c) {
  if (!c) return NUTS_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->ready) cudaEventDestroy(c->ready);
  if (c->t0) cudaEventDestroy(c->t0);
  if (c->t1) cudaEventDestroy(c->t1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return NUTS_OK;
}
int nuts_gather_draws_begin(nuts_sampler_t* s, nuts_comm_t* c, const double* local_dev, double* gathered_dev, uint64_t count) {
  if (!s || !c || !local_dev || !gathered_dev) return fail(NUTS_ERR_INVALID, "nuts_gather_draws_begin: NULL argument");
  if (c->in_flight) return fail(NUTS_ERR_INVALID, "nuts_gather_draws_begin: a gather is in flight; call nuts_gather_draws_end first");
  CUDA_TRY(cudaSetDevice(c->device));
  // ordered after the draws already enqueued on the sampler's stream, but not after anything enqueued later
  CUDA_TRY(cudaEventRecord(c->ready, s->ctx->stream));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ready, 0));
  CUDA_TRY(cudaEventRecord(c->t0, c->stream));
  NCCL_TRY(g_nccl.AllGather(local_dev, gathered_dev, (size_t)count, /* ncclFloat64 */ 8, c->comm, c->stream));
  CUDA_TRY(cudaEventRecord(c->t1, c->stream));
  c->in_flight = true;
  return NUTS_OK;
}
int nuts_gather_draws_end(nuts_comm_t* c, double* elapsed_ms) {
  if (!c) return fail(NUTS_ERR_INVALID, "nuts_gather_draws_end: comm is NULL");
  CUDA_TRY(cudaSetDevice(c->device));
  if (elapsed_ms) *elapsed_ms = 0.0;
  if (!c->in_flight) return NUTS_OK;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->in_flight = false;
  if (elapsed_ms) {
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->t0, c->t1));
    *elapsed_ms = ms;
  }
  return NUTS_OK;
}

// test hook: the branch-free division / square root of device_common.cuh next to the library operators, element by element
// (tests/test_gpu_primitives.py checks bit-equality wherever the range test `ok` holds).  Host arrays of n doubles.
__global__ void k_debug_fast_math(const double* a, const double* b, double* q_fast, double* q_ref, double* r_fast, double* r_ref,
                                  unsigned char* ok_div, unsigned char* ok_sqrt, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool okd = true, oks = true;
  q_fast[i] = div_fast(a[i], b[i], okd);
  q_ref[i] = a[i] / b[i];
  r_fast[i] = sqrt_fast(a[i], oks);
  r_ref[i] = sqrt(a[i]);
  ok_div[i] = okd;
  ok_sqrt[i] = oks;
}
extern "C" int nuts_debug_fast_math(const double* a, const double* b, double* q_fast, double* q_ref, double* r_fast, double* r_ref,
                                    uint8_t* ok_div, uint8_t* ok_sqrt, uint64_t n) {
  TRY(check_device());
  double* d[6] = {};
  unsigned char* u[2] = {};
  int rc = NUTS_OK;
  for (int k = 0; k < 6 && rc == NUTS_OK; ++k) rc = dev_alloc(&d[k], n);
  for (int k = 0; k < 2 && rc == NUTS_OK; ++k) rc = dev_alloc(&u[k], n);
  if (rc == NUTS_OK && (cudaMemcpy(d[0], a, n * 8, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(d[1], b, n * 8, cudaMemcpyHostToDevice) != cudaSuccess))
    rc = fail(NUTS_ERR_CUDA, "nuts_debug_fast_math: H2D failed");
  if (rc == NUTS_OK) {
    k_debug_fast_math<<<(unsigned)((n + 255) / 256), 256>>>(d[0], d[1], d[2], d[3], d[4], d[5], u[0], u[1], n);
    double* out[4] = {q_fast, q_ref, r_fast, r_ref};
    for (int k = 0; k < 4; ++k)
      if (cudaMemcpy(out[k], d[2 + k], n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(NUTS_ERR_CUDA, "nuts_debug_fast_math: D2H failed");
    if (cudaMemcpy(ok_div, u[0], n, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(ok_sqrt, u[1], n, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = fail(NUTS_ERR_CUDA, "nuts_debug_fast_math: D2H failed");
  }
  for (double* p : d) cudaFree(p);
  for (unsigned char* p : u) cudaFree(p);
  return rc;
}

// debug: per-phase clock totals of NB_PHASE_TIMING builds (zeros otherwise); resets the counters
extern "C" int nuts_debug_phase_clocks(nuts_sampler_t* s, unsigned long long* out16) {
  CUDA_TRY(cudaSetDevice(s->ctx->device));
  CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
  CUDA_TRY(cudaMemcpy(out16, s->P.phase_clocks, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(s->P.phase_clocks, 0, 16 * sizeof(unsigned long long)));
  return NUTS_OK;
}

int nuts_sampler_set_step_size(nuts_sampler_t* s, const double* step_size) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<ChainState> cs(ctx->N);
  CUDA_TRY(cudaMemcpyAsync(cs.data(), s->P.cs, cs.size() * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  for (uint64_t c = 0; c < ctx->N; ++c) cs[c].step_size = step_size[c];
  CUDA_TRY(cudaMemcpyAsync(s->P.cs, cs.data(), cs.size() * sizeof(ChainState), cudaMemcpyHostToDevice, ctx->stream));
  return sync(ctx);
}

// ---- low-rank sampler: the transformation comes from the host estimator (LowRankMassMatrixStrategy::update -> LowRankMassMatrix::update)
int nuts_sampler_set_grads_out(nuts_sampler_t* s, double* grads) {
  if (s->lowrank_rmax == 0) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_grads_out: not a low-rank sampler");
  if (grads) {
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, grads);
    if (e != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeManaged)) {
      cudaGetLastError();
      return fail(NUTS_ERR_INVALID, "nuts_sampler_set_grads_out: the buffer must be device or page-locked host memory (nuts_host_alloc)");
    }
    if (at.type == cudaMemoryTypeHost) {
      void* dp = nullptr;
      CUDA_TRY(cudaHostGetDevicePointer(&dp, grads, 0));
      grads = (double*)dp;
    }
  }
  s->P.grads_out = grads;
  return NUTS_OK;
}

int nuts_sampler_set_lowrank_transform(nuts_sampler_t* s, const double* stds, const double* mean, uint64_t rank_max, const double* vals,
                                       const double* vecs, const int32_t* rank, const double* mean_low_rank, uint8_t* accepted) {
  nuts_ctx* ctx = s->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (s->lowrank_rmax == 0) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_lowrank_transform: create the sampler with nuts_sampler_create_lowrank");
  if (!s->positioned) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_lowrank_transform: call nuts_set_position first");
  if (!stds || !mean || !mean_low_rank) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_lowrank_transform: NULL argument");
  if (rank_max > s->lowrank_rmax) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_lowrank_transform: rank_max %llu exceeds the sampler's %llu", (unsigned long long)rank_max, (unsigned long long)s->lowrank_rmax);
  if (rank_max > 0 && (!vals || !vecs)) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_lowrank_transform: vals / vecs are NULL");
  const uint64_t N = ctx->N, d = ctx->d, RM = s->lowrank_rmax, ld = (uint64_t)s->P.ld;
  auto finite = [](const double* p, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
      if (!std::isfinite(p[i])) return false;
    return true;
  };
  std::vector<ChainState> cs(N);
  std::vector<int> rk(N);
  std::vector<double> vs(N * RM), vi(N * RM);
  CUDA_TRY(cudaMemcpyAsync(cs.data(), s->P.cs, N * sizeof(ChainState), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(rk.data(), s->lr_rank, N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(vs.data(), s->lr_vals_sqrt, N * RM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(vi.data(), s->lr_vals_sqrt_inv, N * RM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  TRY(sync(ctx));
  std::vector<double> inv(d);
  for (uint64_t c = 0; c < N; ++c) {
    const int r = rank_max == 0 ? 0 : (rank ? rank[c] : (int)rank_max);
    if (r < 0 || (uint64_t)r > rank_max) return fail(NUTS_ERR_INVALID, "nuts_sampler_set_lowrank_transform: rank[%llu] = %d outside 0 .. rank_max", (unsigned long long)c, r);
    bool good = finite(stds + c * d, d) && finite(mean + c * d, d);  // low_rank.rs:168-173: otherwise the old transformation stays
    if (r > 0) good = good && finite(vals + c * rank_max, (uint64_t)r) && finite(vecs + c * rank_max * d, (uint64_t)r * d);
    if (accepted) accepted[c] = good ? 1 : 0;
    if (!good || !cs[c].alive) continue;
    // diag.set_transform (diagonal.rs:156-162) + InnerMatrix::new (low_rank.rs:55-71) + logdet (low_rank.rs:186-187)
    double logdet = 0.0, contrib = 0.0;
    for (uint64_t i = 0; i < d; ++i) {
      inv[i] = 1.0 / stds[c * d + i];
      logdet += std::log(inv[i]);
    }
    for (int k = 0; k < r; ++k) {
      const double lam = vals[c * rank_max + k];
      contrib += -0.5 * std::log(lam);
      const double sq = std::sqrt(lam);
      vs[c * RM + k] = sq;
      vi[c * RM + k] = 1.0 / sq;
    }
    rk[c] = r;
    cs[c].mm_logdet = contrib + logdet;
    cs[c].mm_id += 1;
    const size_t row = c * ld;
    CUDA_TRY(cudaMemcpyAsync(s->P.stds + row, stds + c * d, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(s->P.inv_stds + row, inv.data(), d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(s->P.mean + row, mean + c * d, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(s->lr_mu + row, mean_low_rank + c * d, d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(s->lr_vecs + c * RM * ld, 0, RM * ld * sizeof(double), ctx->stream));
    if (r > 0)
      CUDA_TRY(cudaMemcpy2DAsync(s->lr_vecs + c * RM * ld, ld * sizeof(double), vecs + c * rank_max * d, d * sizeof(double), d * sizeof(double),
                                 (size_t)r, cudaMemcpyHostToDevice, ctx->stream));
    TRY(sync(ctx));  // `inv` is reused by the next chain
  }
  CUDA_TRY(cudaMemcpyAsync(s->P.cs, cs.data(), N * sizeof(ChainState), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(s->lr_rank, rk.data(), N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(s->lr_vals_sqrt, vs.data(), N * RM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(s->lr_vals_sqrt_inv, vi.data(), N * RM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // the first mass-matrix update of a run re-initialises the step size from the current point (adapt_strategy.rs:204-214)
  s->P.mode = 2;
  s->P.init_position = nullptr;
  s->P.init_mask = nullptr;
  s->P.status_out = s->d_status;
  s->P.n_draws = 0;
  s->P.draws_per_unit = 1;
  TRY(launch_engine(s));
  return sync(ctx);
}

}  // extern "C"
