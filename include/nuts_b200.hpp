// nuts_b200.hpp — C++17 host side over the C ABI (include/nuts_b200.h), mirroring the public surface of the reference for the
// accelerated path.  The reference is compiled code (Rust) and this image has no Rust toolchain, so the host mirror of
// `DiagNutsSettings` / `Model + Math` / `Chain` is written in C++ (header-only, no CUDA or torch types); INTEGRATION.md shows the
// Rust binding a nuts-rs maintainer would add over the same entry points.
//
//   reference                                              here
//   DiagNutsSettings { num_tune, maxdepth, .. }             nuts_b200::DiagNutsSettings      (src/sampler.rs:199-239, 507-531)
//   Model::math() -> CpuMath<impl CpuLogpFunc>              nuts_b200::CudaMath              (src/model.rs:18-33, src/math/cpu_math.rs:885-970)
//   Settings::new_chain(chain, math, rng) -> impl Chain     nuts_b200::Chains                (src/sampler.rs:745-772; ALL chains of one GPU)
//   Chain::set_position(&init) -> Result<()>                Chains::set_position             (src/chain.rs:137-149; per-chain status)
//   Chain::draw() -> (position, stats)  x n                 Chains::draw(n) -> Draws         (src/chain.rs:151-188, 215-231)
//   NutsError::BadInitGrad / LogpFailure                    status 3 per chain; nuts_b200::Error for fatal library errors
//   init retries of Sampler::new (500 fresh init points)    Chains::set_position_with_retries   (src/sampler.rs:1133-1143)
//   Sampler::{pause, resume, progress, abort, wait}         nuts_b200::Sampler (batch granularity)   (src/sampler.rs:1253-1552)
//   (no counterpart: chains live in one process)            Chains::checkpoint / restore = nuts_chain_state_t: stop and resume a run
//   CpuLogpFunc implemented by the user                     CudaMath::user(..): the density compiled in from include/nuts_user_logp.cuh
//
// Chains::draw hands back the draws draw-major ([draw][chain][dim], what one kernel launch produces); Draws::position(chain, draw)
// and the statistics accessors give the reference's per-chain view.  With `pinned = true` the draws live in page-locked memory and
// are written by the kernel directly.  There is no CPU fallback: without an sm_100 device every call throws Error(NUTS_ERR_NO_DEVICE).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "nuts_b200.h"

namespace nuts_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
  if (rc != NUTS_OK) throw Error(rc, std::string("libnuts_b200: ") + nuts_last_error());
}
inline bool device_available() { return nuts_device_available() == NUTS_OK; }

// DiagNutsSettings::default() with public fields, like the reference's struct-update syntax:
//   DiagNutsSettings s; s.num_tune = 1000; s.maxdepth = 3;
struct DiagNutsSettings : nuts_settings_t {
  DiagNutsSettings() { nuts_settings_default(this); }
};

// Device-side log density of `nchains` independent chains (the stand-ins for CpuLogpFunc::logp of SURVEY §8 a4').
class CudaMath {
 public:
  // logp = -sum (x - mu)^2 / 2                      (reference benches/sample.rs:49-62, src/math/test_logps.rs:49-58)
  static CudaMath normal(uint64_t nchains, uint64_t dim, double mu, int device = 0) {
    nuts_logp_desc_t m{};
    m.kind = NUTS_LOGP_GAUSS_ISO;
    m.mu_scalar = mu;
    return CudaMath(nchains, dim, m, device);
  }
  // logp = -sum (x_i - mu)^2 / (2 sigma_i^2)
  static CudaMath diag_normal(uint64_t nchains, uint64_t dim, double mu, const std::vector<double>& sigma, int device = 0) {
    if (sigma.size() != dim) throw Error(NUTS_ERR_INVALID, "diag_normal: sigma must have dim entries");
    nuts_logp_desc_t m{};
    m.kind = NUTS_LOGP_GAUSS_DIAG;
    m.mu_scalar = mu;
    m.sigma = sigma.data();
    return CudaMath(nchains, dim, m, device);
  }
  // Sigma = I + s 1 1^T                              (reference tests/sample_normal.rs:29-96)
  static CudaMath correlated_normal(uint64_t nchains, uint64_t dim, double mu, double s, int device = 0) {
    nuts_logp_desc_t m{};
    m.kind = NUTS_LOGP_GAUSS_RANK1;
    m.mu_scalar = mu;
    m.rank1_scale = s;
    return CudaMath(nchains, dim, m, device);
  }
  static CudaMath funnel(uint64_t nchains, uint64_t dim, double scale = 3.0, int device = 0) {
    nuts_logp_desc_t m{};
    m.kind = NUTS_LOGP_FUNNEL;
    m.funnel_scale = scale;
    return CudaMath(nchains, dim, m, device);
  }
  // the user-supplied device density the library was built with (include/nuts_user_logp.cuh; `make USER_LOGP=my_model.cuh`)
  static CudaMath user(uint64_t nchains, uint64_t dim, const std::vector<double>& params, int device = 0) {
    nuts_logp_desc_t m{};
    m.kind = NUTS_LOGP_USER;
    m.user_params = params.data();
    m.n_user_params = params.size();
    return CudaMath(nchains, dim, m, device);
  }
  CudaMath(uint64_t nchains, uint64_t dim, const nuts_logp_desc_t& model, int device = 0) : nchains_(nchains), dim_(dim) {
    check(nuts_ctx_create(&ctx_, device, nchains, dim, &model));
  }
  CudaMath(CudaMath&& o) noexcept : ctx_(o.ctx_), nchains_(o.nchains_), dim_(o.dim_) { o.ctx_ = nullptr; }
  CudaMath(const CudaMath&) = delete;
  CudaMath& operator=(const CudaMath&) = delete;
  ~CudaMath() {
    if (ctx_) nuts_ctx_destroy(ctx_);
  }
  uint64_t dim() const { return dim_; }  // Math::dim (src/math/math.rs:69)
  uint64_t nchains() const { return nchains_; }
  nuts_ctx_t* handle() const { return ctx_; }

 private:
  nuts_ctx_t* ctx_ = nullptr;
  uint64_t nchains_, dim_;
};

// What n x Chain::draw returned for every chain: positions + the NutsStats columns (src/chain.rs:215-231 and the flattened
// hamiltonian / adapt / point / divergence statistics).
class Draws {
 public:
  Draws(uint64_t n_draws, uint64_t nchains, uint64_t dim, bool pinned) : n_(n_draws), N_(nchains), d_(dim) {
    const uint64_t nd = n_ * N_ * d_;
    if (pinned) {
      void* p = nullptr;
      check(nuts_host_alloc(&p, (nd > 0 ? nd : 1) * sizeof(double)));
      pinned_ = static_cast<double*>(p);
    } else {
      pageable_.resize(nd);
    }
    const size_t m = n_ * N_;
    depth.resize(m), n_steps.resize(m), index_in_trajectory.resize(m);
    maxdepth_reached.resize(m), diverging.resize(m), tuning.resize(m);
    logp.resize(m), energy.resize(m), energy_error.resize(m), step_size.resize(m), step_size_bar.resize(m);
    mean_tree_accept.resize(m), mean_tree_accept_sym.resize(m), max_energy_error.resize(m), fisher_distance.resize(m);
  }
  Draws(Draws&& o) noexcept { *this = std::move(o); }
  Draws& operator=(Draws&& o) noexcept {
    if (this != &o) {
      release();
      n_ = o.n_, N_ = o.N_, d_ = o.d_, pinned_ = o.pinned_, o.pinned_ = nullptr;
      pageable_ = std::move(o.pageable_);
      depth = std::move(o.depth), n_steps = std::move(o.n_steps), index_in_trajectory = std::move(o.index_in_trajectory);
      maxdepth_reached = std::move(o.maxdepth_reached), diverging = std::move(o.diverging), tuning = std::move(o.tuning);
      logp = std::move(o.logp), energy = std::move(o.energy), energy_error = std::move(o.energy_error);
      step_size = std::move(o.step_size), step_size_bar = std::move(o.step_size_bar), mean_tree_accept = std::move(o.mean_tree_accept);
      mean_tree_accept_sym = std::move(o.mean_tree_accept_sym), max_energy_error = std::move(o.max_energy_error);
      fisher_distance = std::move(o.fisher_distance);
    }
    return *this;
  }
  Draws(const Draws&) = delete;
  Draws& operator=(const Draws&) = delete;
  ~Draws() { release(); }

  uint64_t n_draws() const { return n_; }
  uint64_t nchains() const { return N_; }
  uint64_t dim() const { return d_; }
  double* data() { return pinned_ ? pinned_ : pageable_.data(); }  // [draw][chain][dim]
  const double* data() const { return pinned_ ? pinned_ : pageable_.data(); }
  // the `position` of Chain::draw number `draw` of chain `chain` (dim doubles)
  const double* position(uint64_t chain, uint64_t draw) const { return data() + (draw * N_ + chain) * d_; }
  size_t at(uint64_t chain, uint64_t draw) const { return draw * N_ + chain; }  // index into the statistics columns

  std::vector<uint64_t> depth, n_steps;
  std::vector<int64_t> index_in_trajectory;
  std::vector<uint8_t> maxdepth_reached, diverging, tuning;
  std::vector<double> logp, energy, energy_error, step_size, step_size_bar, mean_tree_accept, mean_tree_accept_sym, max_energy_error,
      fisher_distance;

  nuts_stats_t view() {
    nuts_stats_t s{};
    s.depth = depth.data(), s.maxdepth_reached = maxdepth_reached.data(), s.index_in_trajectory = index_in_trajectory.data();
    s.logp = logp.data(), s.energy = energy.data(), s.energy_error = energy_error.data(), s.diverging = diverging.data();
    s.step_size = step_size.data(), s.step_size_bar = step_size_bar.data(), s.mean_tree_accept = mean_tree_accept.data();
    s.mean_tree_accept_sym = mean_tree_accept_sym.data(), s.n_steps = n_steps.data(), s.max_energy_error = max_energy_error.data();
    s.tuning = tuning.data(), s.fisher_distance = fisher_distance.data();
    return s;
  }

 private:
  void release() {
    if (pinned_) nuts_host_free(pinned_);
    pinned_ = nullptr;
  }
  uint64_t n_ = 0, N_ = 0, d_ = 0;
  double* pinned_ = nullptr;
  std::vector<double> pageable_;
};

// Everything the chains carry from one draw to the next (nuts_chain_state_t), owning its arrays: a checkpoint of a run.
struct ChainState {
  uint64_t nchains = 0, dim = 0;
  std::vector<double> position, gradient, transformed_position, transformed_gradient, logp, point_logdet, stds, inv_stds, mean,
      mass_matrix_logdet, step_size, da_log_step, da_log_step_adapted, da_hbar, da_mu, draw_mean, draw_var, grad_mean, grad_var,
      draw_mean_bg, draw_var_bg, grad_mean_bg, grad_var_bg;
  std::vector<int64_t> point_transform_id, mass_matrix_id;
  std::vector<uint64_t> da_count, foreground_count, background_count, last_update, current_window_size, draw_count, rng_counter,
      total_leapfrogs;
  std::vector<uint8_t> tuning, has_initial_mass_matrix, alive;
  ChainState(uint64_t n, uint64_t d) : nchains(n), dim(d) {
    for (auto* v : {&position, &gradient, &transformed_position, &transformed_gradient, &stds, &inv_stds, &mean, &draw_mean, &draw_var,
                    &grad_mean, &grad_var, &draw_mean_bg, &draw_var_bg, &grad_mean_bg, &grad_var_bg})
      v->resize(n * d);
    for (auto* v : {&logp, &point_logdet, &mass_matrix_logdet, &step_size, &da_log_step, &da_log_step_adapted, &da_hbar, &da_mu}) v->resize(n);
    for (auto* v : {&point_transform_id, &mass_matrix_id}) v->resize(n);
    for (auto* v : {&da_count, &foreground_count, &background_count, &last_update, &current_window_size, &draw_count, &rng_counter,
                    &total_leapfrogs})
      v->resize(n);
    for (auto* v : {&tuning, &has_initial_mass_matrix, &alive}) v->resize(n);
  }
  nuts_chain_state_t view() {
    nuts_chain_state_t s{};
    s.position = position.data(), s.gradient = gradient.data(), s.transformed_position = transformed_position.data();
    s.transformed_gradient = transformed_gradient.data(), s.logp = logp.data(), s.point_logdet = point_logdet.data();
    s.point_transform_id = point_transform_id.data(), s.stds = stds.data(), s.inv_stds = inv_stds.data(), s.mean = mean.data();
    s.mass_matrix_logdet = mass_matrix_logdet.data(), s.mass_matrix_id = mass_matrix_id.data(), s.step_size = step_size.data();
    s.da_log_step = da_log_step.data(), s.da_log_step_adapted = da_log_step_adapted.data(), s.da_hbar = da_hbar.data();
    s.da_mu = da_mu.data(), s.da_count = da_count.data(), s.draw_mean = draw_mean.data(), s.draw_var = draw_var.data();
    s.grad_mean = grad_mean.data(), s.grad_var = grad_var.data(), s.draw_mean_bg = draw_mean_bg.data(), s.draw_var_bg = draw_var_bg.data();
    s.grad_mean_bg = grad_mean_bg.data(), s.grad_var_bg = grad_var_bg.data(), s.foreground_count = foreground_count.data();
    s.background_count = background_count.data(), s.tuning = tuning.data(), s.has_initial_mass_matrix = has_initial_mass_matrix.data();
    s.last_update = last_update.data(), s.current_window_size = current_window_size.data(), s.draw_count = draw_count.data();
    s.rng_counter = rng_counter.data(), s.total_leapfrogs = total_leapfrogs.data(), s.alive = alive.data();
    return s;
  }
};

// All chains of one GPU: `settings.new_chain(chain_id, math, rng)` for chain ids chain_id_offset .. chain_id_offset + nchains - 1
// (random streams are keyed by the global chain id, src/sampler.rs:1105-1106, so shards reproduce the unsharded run).
class Chains {
 public:
  // lowrank_rank_max > 0: LowRankNutsSettings chains (src/sampler.rs:636-690) on an engine with the low-rank transformation
  // compiled in; the caller's estimator installs transformations with set_lowrank_transform (nuts_rs_b200/lowrank.py is the
  // reference's LowRankMassMatrixStrategy around this API)
  Chains(const CudaMath& math, const DiagNutsSettings& settings, uint64_t seed, uint64_t chain_id_offset = 0, uint64_t lowrank_rank_max = 0)
      : nchains_(math.nchains()), dim_(math.dim()) {
    if (lowrank_rank_max > 0) check(nuts_sampler_create_lowrank(math.handle(), &s_, &settings, seed, chain_id_offset, lowrank_rank_max));
    else check(nuts_sampler_create(math.handle(), &s_, &settings, seed, chain_id_offset));
  }
  // LowRankMassMatrix::update (src/transform/low_rank.rs:158-190) for every chain: stds / mean / mean_low_rank [nchains][dim],
  // vals [nchains][rank_max], vecs [nchains][rank_max][dim], rank [nchains] (or empty: rank_max everywhere).  Returns the accepted
  // flags (0: non-finite input, that chain keeps its transformation).
  std::vector<uint8_t> set_lowrank_transform(const std::vector<double>& stds, const std::vector<double>& mean, uint64_t rank_max,
                                             const std::vector<double>& vals, const std::vector<double>& vecs,
                                             const std::vector<int32_t>& rank, const std::vector<double>& mean_low_rank) {
    if (stds.size() != nchains_ * dim_ || mean.size() != stds.size() || mean_low_rank.size() != stds.size() ||
        vals.size() != nchains_ * rank_max || vecs.size() != nchains_ * rank_max * dim_ || (!rank.empty() && rank.size() != nchains_))
      throw Error(NUTS_ERR_INVALID, "set_lowrank_transform: array sizes do not match nchains / dim / rank_max");
    std::vector<uint8_t> accepted(nchains_);
    check(nuts_sampler_set_lowrank_transform(s_, stds.data(), mean.data(), rank_max, vals.data(), vecs.data(),
                                             rank.empty() ? nullptr : rank.data(), mean_low_rank.data(), accepted.data()));
    return accepted;
  }
  // gradients of the following draws into a device / page-locked buffer [n_draws][nchains][dim] (the estimator's input); nullptr: off
  void set_grads_out(double* grads) { check(nuts_sampler_set_grads_out(s_, grads)); }
  Chains(const Chains&) = delete;
  Chains& operator=(const Chains&) = delete;
  ~Chains() {
    if (s_) nuts_sampler_destroy(s_);
  }
  // Chain::set_position for every chain; positions [nchains][dim].  Per-chain status: 0 ok, 3 = NutsError::BadInitGrad.
  std::vector<int32_t> set_position(const double* positions) {
    std::vector<int32_t> status(nchains_);
    check(nuts_set_position(s_, positions, status.data()));
    return status;
  }
  std::vector<int32_t> set_position(const std::vector<double>& positions) {
    if (positions.size() != nchains_ * dim_) throw Error(NUTS_ERR_INVALID, "set_position: need nchains x dim values");
    return set_position(positions.data());
  }
  // The chain start of the reference's Sampler (src/sampler.rs:1133-1143): `init(chain_id, out[dim])` is asked for a fresh initial
  // point for every chain whose set_position failed, up to max_tries times; returns the final per-chain status.
  std::vector<int32_t> set_position_with_retries(const std::function<void(uint64_t, double*)>& init, uint64_t chain_id_offset = 0,
                                                 int max_tries = 500) {
    std::vector<double> pos(nchains_ * dim_);
    for (uint64_t c = 0; c < nchains_; ++c) init(chain_id_offset + c, pos.data() + c * dim_);
    std::vector<int32_t> status = set_position(pos);
    for (int t = 1; t < max_tries; ++t) {
      std::vector<uint8_t> mask(nchains_, 0);
      bool any = false;
      for (uint64_t c = 0; c < nchains_; ++c)
        if (status[c] != 0) {
          mask[c] = 1, any = true;
          init(chain_id_offset + c, pos.data() + c * dim_);
        }
      if (!any) break;
      check(nuts_set_position_masked(s_, pos.data(), mask.data(), status.data()));
    }
    return status;
  }
  // checkpoint / resume: the complete state between two draws; restore() on a fresh Chains (same model, settings, seed, offset)
  // continues bit-identically
  ChainState checkpoint() {
    ChainState st(nchains_, dim_);
    nuts_chain_state_t v = st.view();
    check(nuts_sampler_get_chain_state(s_, &v));
    return st;
  }
  void restore(ChainState& st) {
    if (st.nchains != nchains_ || st.dim != dim_) throw Error(NUTS_ERR_INVALID, "restore: checkpoint of a different shape");
    nuts_chain_state_t v = st.view();
    check(nuts_sampler_set_chain_state(s_, &v));
  }
  // n x Chain::draw for every chain
  Draws draw(uint64_t n_draws, bool pinned = true) {
    Draws out(n_draws, nchains_, dim_, pinned);
    nuts_stats_t st = out.view();
    check(nuts_draw(s_, n_draws, out.data(), &st));
    return out;
  }
  // leapfrog steps so far (all chains; incl. step-size searches and divergent steps) and draws per chain
  std::pair<uint64_t, uint64_t> counters() {
    uint64_t lf = 0, draws = 0;
    check(nuts_sampler_counters(s_, &lf, &draws));
    return {lf, draws};
  }
  bool last_draw_direct() {
    int32_t v = 0;
    check(nuts_sampler_last_draw_direct(s_, &v));
    return v != 0;
  }
  nuts_sampler_t* handle() const { return s_; }

 private:
  nuts_sampler_t* s_ = nullptr;
  uint64_t nchains_, dim_;
};

// What the reference's Sampler offers around its chains (src/sampler.rs:1253-1552): run in the background, pause / resume, progress,
// abort - here at the granularity of one batch of draws (one nuts_draw call for all chains of the GPU).  Every finished batch is
// handed to `sink` (the reference writes into its trace storage there); the sampler owns nothing but the control flow.
struct Progress {  // per-sampler view of src/sampler.rs:165-174 + ChainProgress :1554-1660
  uint64_t finished_draws = 0, total_draws = 0, tuning_draws = 0, divergences = 0, leapfrogs = 0;
  bool tuning = true, paused = false, finished = false;
};
class Sampler {
 public:
  using Sink = std::function<void(const Draws&, uint64_t first_draw)>;
  Sampler(Chains& chains, const DiagNutsSettings& settings, Sink sink, uint64_t batch = 50)
      : chains_(chains), total_(settings.num_tune + settings.num_draws), num_tune_(settings.num_tune), batch_(batch), sink_(std::move(sink)) {
    progress_.total_draws = total_;
    progress_.tuning_draws = num_tune_;
    worker_ = std::thread([this] { run(); });
  }
  ~Sampler() {
    abort();
    if (worker_.joinable()) worker_.join();
  }
  void pause() {  // takes effect after the batch in flight (Sampler::pause, src/sampler.rs:1469-1485)
    std::lock_guard<std::mutex> l(m_);
    paused_ = true;
  }
  void resume() {
    {
      std::lock_guard<std::mutex> l(m_);
      paused_ = false;
    }
    cv_.notify_all();
  }
  void abort() {
    {
      std::lock_guard<std::mutex> l(m_);
      abort_ = true;
      paused_ = false;
    }
    cv_.notify_all();
  }
  Progress progress() {
    std::lock_guard<std::mutex> l(m_);
    Progress p = progress_;
    p.paused = paused_;
    return p;
  }
  // blocks until all draws are done (or abort()); rethrows an error of the worker (Sampler::wait_timeout / finish)
  void wait() {
    if (worker_.joinable()) worker_.join();
    if (error_) std::rethrow_exception(error_);
  }

 private:
  void run() {
    try {
      uint64_t done = 0;
      while (done < total_) {
        {
          std::unique_lock<std::mutex> l(m_);
          cv_.wait(l, [this] { return !paused_ || abort_; });
          if (abort_) break;
        }
        const uint64_t n = std::min<uint64_t>(batch_, total_ - done);
        Draws d = chains_.draw(n);
        uint64_t div = 0;
        for (uint8_t f : d.diverging) div += f;
        const uint64_t lf = chains_.counters().first;
        if (sink_) sink_(d, done);
        done += n;
        std::lock_guard<std::mutex> l(m_);
        progress_.finished_draws = done;
        progress_.divergences += div;
        progress_.leapfrogs = lf;
        progress_.tuning = done < num_tune_;
      }
      std::lock_guard<std::mutex> l(m_);
      progress_.finished = progress_.finished_draws >= total_;
    } catch (...) {
      error_ = std::current_exception();
    }
  }
  Chains& chains_;
  uint64_t total_, num_tune_, batch_;
  Sink sink_;
  std::mutex m_;
  std::condition_variable cv_;
  bool paused_ = false, abort_ = false;
  Progress progress_;
  std::exception_ptr error_;
  std::thread worker_;
};

}  // namespace nuts_b200
