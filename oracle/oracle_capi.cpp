// TEST INFRASTRUCTURE — C entry points of the CPU oracle (nuts_oracle.hpp) for ctypes / bench.py.
// Not part of the product; see the header of nuts_oracle.hpp for who may call this.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "../include/nuts_b200.h"
#include "nuts_oracle.hpp"
#include "lowrank_estimator.hpp"

using namespace oracle;

namespace {
LogpFunc* make_model(int kind, size_t dim, double mu_scalar, const double* mu, const double* sigma, double rank1_scale,
                     double funnel_scale) {
  Vec m(dim, mu_scalar);
  if (mu) m.assign(mu, mu + dim);
  switch (kind) {
    case NUTS_LOGP_GAUSS_ISO: {
      auto* g = new GaussIso();
      g->dim = dim;
      g->mu = m;
      return g;
    }
    case NUTS_LOGP_GAUSS_DIAG: {
      auto* g = new GaussDiag();
      g->dim = dim;
      g->mu = m;
      g->prec.resize(dim);
      for (size_t i = 0; i < dim; ++i) g->prec[i] = 1.0 / (sigma[i] * sigma[i]);
      return g;
    }
    case NUTS_LOGP_GAUSS_RANK1: {
      auto* g = new GaussRank1();
      g->dim = dim;
      g->mu = m;
      g->prec_rank1_coeff = rank1_scale / (1.0 + rank1_scale * (double)dim);  // tests/sample_normal.rs:36
      return g;
    }
    case NUTS_LOGP_FUNNEL: {
      auto* g = new Funnel();
      g->dim = dim;
      g->fs = funnel_scale;
      return g;
    }
  }
  return nullptr;
}

NutsSettings convert_settings(const nuts_settings_t* s) {
  NutsSettings o;
  o.num_tune = s->num_tune;
  o.num_draws = s->num_draws;
  o.maxdepth = s->maxdepth;
  o.mindepth = s->mindepth;
  o.max_energy_error = s->max_energy_error;
  o.check_turning = s->check_turning != 0;
  if (s->has_target_integration_time) o.target_integration_time = s->target_integration_time;
  o.num_chains = s->num_chains;
  o.seed = s->seed;
  o.extra_doublings = s->extra_doublings;
  o.trajectory_kind = s->trajectory_kind == NUTS_KINETIC_EXACT_NORMAL     ? KineticEnergyKind::ExactNormal
                      : s->trajectory_kind == NUTS_KINETIC_MICROCANONICAL ? KineticEnergyKind::Microcanonical
                                                                          : KineticEnergyKind::Euclidean;
  const auto& a = s->adapt_options;
  o.adapt_options.early_window = a.early_window;
  o.adapt_options.step_size_window = a.step_size_window;
  o.adapt_options.mass_matrix_switch_freq = a.mass_matrix_switch_freq;
  o.adapt_options.early_mass_matrix_switch_freq = a.early_mass_matrix_switch_freq;
  o.adapt_options.mass_matrix_update_freq = a.mass_matrix_update_freq;
  o.adapt_options.mass_matrix_window_growth = a.mass_matrix_window_growth;
  o.adapt_options.use_grad_based_estimate = a.mass_matrix_options.use_grad_based_estimate != 0;
  auto& ss = o.adapt_options.step_size_settings;
  ss.target_accept = a.step_size_settings.target_accept;
  ss.initial_step = a.step_size_settings.initial_step;
  if (a.step_size_settings.has_jitter) ss.jitter = a.step_size_settings.jitter;
  else ss.jitter = std::nullopt;
  ss.method = a.step_size_settings.adapt_options.method == NUTS_STEPSIZE_FIXED  ? StepSizeMethod::Fixed
              : a.step_size_settings.adapt_options.method == NUTS_STEPSIZE_ADAM ? StepSizeMethod::Adam
                                                                                  : StepSizeMethod::DualAverage;
  ss.adam.beta1 = a.step_size_settings.adapt_options.adam.beta1;
  ss.adam.beta2 = a.step_size_settings.adapt_options.adam.beta2;
  ss.adam.epsilon = a.step_size_settings.adapt_options.adam.epsilon;
  ss.adam.learning_rate = a.step_size_settings.adapt_options.adam.learning_rate;
  ss.fixed_step = a.step_size_settings.adapt_options.fixed_step;
  ss.dual_average.k = a.step_size_settings.adapt_options.dual_average.k;
  ss.dual_average.t0 = a.step_size_settings.adapt_options.dual_average.t0;
  ss.dual_average.gamma = a.step_size_settings.adapt_options.dual_average.gamma;
  ss.dual_average.max_step_size = a.step_size_settings.adapt_options.dual_average.max_step_size;
  return o;
}

struct Ham {
  LogpFunc* model;
  TransformedHamiltonian h;
  explicit Ham(LogpFunc* m) : model(m), h(m) {}
};

struct Sampler {
  LogpFunc* model;
  size_t dim;
  uint64_t nchains;
  std::vector<std::unique_ptr<NutsChain>> chains;
  std::vector<int> alive;
};

struct NullCollector {
  void register_leapfrog(const TransformedPoint&, bool) {}
};
template <class F>
void parallel_chains(uint64_t n, int nthreads, F f) {
  if (nthreads <= 1) {
    for (uint64_t c = 0; c < n; ++c) f(c);
    return;
  }
  std::vector<std::thread> th;
  std::atomic<uint64_t> next{0};
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&] {
      for (;;) {
        uint64_t c = next.fetch_add(1);
        if (c >= n) break;
        f(c);
      }
    });
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

// ---------------- primitives (src/math/util.rs) ----------------
double orc_logaddexp(double a, double b) { return logaddexp(a, b); }
void orc_multiply(const double* x, const double* y, double* out, size_t n) { multiply(x, y, out, n); }
void orc_multiply_inplace(double* out, const double* x, size_t n) { multiply_inplace(out, x, n); }
void orc_axpy(const double* x, double* y, double a, size_t n) { axpy(x, y, a, n); }
void orc_axpy_out(const double* x, const double* y, double a, double* out, size_t n) { axpy_out(x, y, a, out, n); }
double orc_vector_dot(const double* a, const double* b, size_t n) { return vector_dot(a, b, n); }
void orc_std_norm_flow(const double* pos, double* pos_out, double* vel, double epsilon, size_t n) { std_norm_flow(pos, pos_out, vel, epsilon, n); }
void orc_std_norm_grad_flow(const double* pos, const double* grad, const double* vel, double* vel_out, double epsilon, size_t n) {
  std_norm_grad_flow(pos, grad, vel, vel_out, epsilon, n);
}
void orc_std_norm_grad_flow_inplace(const double* pos, const double* grad, double* vel, double epsilon, size_t n) {
  std_norm_grad_flow_inplace(pos, grad, vel, epsilon, n);
}
void orc_array_normalize(double* v, size_t n) { array_normalize(v, n); }
double orc_esh_momentum_update(const double* gradient, double* momentum, double step_size, size_t n) {
  return esh_momentum_update(gradient, momentum, step_size, n);
}
void orc_scalar_prods3(const double* p1, const double* n1, const double* p2, const double* x, const double* y, size_t n,
                       double* o1, double* o2) {
  scalar_prods3(p1, n1, p2, x, y, n, o1, o2);
}
void orc_scalar_prods2(const double* p1, const double* p2, const double* x, const double* y, size_t n, double* o1, double* o2) {
  scalar_prods2(p1, p2, x, y, n, o1, o2);
}
// ---------------- CpuMath extras (src/math/cpu_math.rs) ----------------
double orc_sq_norm_sum(const double* x, const double* y, size_t n) { return sq_norm_sum(Vec(x, x + n), Vec(y, y + n)); }
int orc_array_all_finite(const double* a, size_t n) { return array_all_finite(Vec(a, a + n)); }
int orc_array_all_finite_and_nonzero(const double* a, size_t n) { return array_all_finite_and_nonzero(Vec(a, a + n)); }
double orc_array_sum_ln(const double* a, size_t n) { return array_sum_ln(Vec(a, a + n)); }
void orc_array_update_variance(double* mean, double* variance, const double* value, double diff_scale, size_t n) {
  Vec m(mean, mean + n), v(variance, variance + n);
  array_update_variance(m, v, Vec(value, value + n), diff_scale);
  std::copy(m.begin(), m.end(), mean);
  std::copy(v.begin(), v.end(), variance);
}
void orc_array_update_var_inv_std_draw(double* inv_std, double* std_, const double* draw_var, double scale, int has_fill,
                                       double fill, double lo, double hi, size_t n) {
  Vec is(inv_std, inv_std + n), s(std_, std_ + n);
  array_update_var_inv_std_draw(is, s, Vec(draw_var, draw_var + n), scale, has_fill ? std::optional<double>(fill) : std::nullopt,
                                lo, hi);
  std::copy(is.begin(), is.end(), inv_std);
  std::copy(s.begin(), s.end(), std_);
}
void orc_array_update_var_inv_std_draw_grad(double* inv_std, double* std_, const double* draw_var, const double* grad_var,
                                            int has_fill, double fill, double lo, double hi, size_t n) {
  Vec is(inv_std, inv_std + n), s(std_, std_ + n);
  array_update_var_inv_std_draw_grad(is, s, Vec(draw_var, draw_var + n), Vec(grad_var, grad_var + n),
                                     has_fill ? std::optional<double>(fill) : std::nullopt, lo, hi);
  std::copy(is.begin(), is.end(), inv_std);
  std::copy(s.begin(), s.end(), std_);
}
void orc_array_update_var_inv_std_grad(double* inv_std, double* std_, const double* gradient, double fill, double lo, double hi,
                                       size_t n) {
  Vec is(inv_std, inv_std + n), s(std_, std_ + n);
  array_update_var_inv_std_grad(is, s, Vec(gradient, gradient + n), fill, lo, hi);
  std::copy(is.begin(), is.end(), inv_std);
  std::copy(s.begin(), s.end(), std_);
}

// ---------------- random streams (rng_spec.hpp) ----------------
void orc_philox(uint64_t seed, uint64_t stream, uint64_t counter, uint32_t* out4) {
  PhiloxBlock b = philox4x32_10(seed, stream, counter);
  std::memcpy(out4, b.r, 16);
}
double orc_det_log(double u) { return det_log(u); }
void orc_det_sincos2pi(double u, double* s, double* c) { det_sincos2pi(u, s, c); }
uint64_t orc_fill_normal(uint64_t seed, uint64_t stream, uint64_t counter, double* out, size_t d) {
  Rng r(seed, stream, counter);
  r.fill_normal(out, d);
  return r.counter;
}
double orc_next_f64(uint64_t seed, uint64_t stream, uint64_t counter) { return Rng(seed, stream, counter).next_f64(); }
int orc_next_bool(uint64_t seed, uint64_t stream, uint64_t counter) { return Rng(seed, stream, counter).next_bool(); }
double orc_uniform(uint64_t seed, uint64_t stream, uint64_t counter, double lo, double hi) {
  return Rng(seed, stream, counter).uniform(lo, hi);
}

// ---------------- models ----------------
void* orc_model_create(const nuts_logp_desc_t* d, uint64_t dim) {
  return make_model(d->kind, dim, d->mu_scalar, d->mu, d->sigma, d->rank1_scale, d->funnel_scale);
}
void orc_model_destroy(void* m) { delete (LogpFunc*)m; }
double orc_model_logp(void* m, const double* x, double* grad) { return ((LogpFunc*)m)->logp(x, grad); }

// ---------------- Hamiltonian / DiagMassMatrix white-box handles ----------------
void* orc_ham_create(void* model) { return new Ham((LogpFunc*)model); }
void orc_ham_destroy(void* h) { delete (Ham*)h; }
void orc_ham_set_step_size(void* h, double eps) { ((Ham*)h)->h.step_size = eps; }
void orc_ham_set_kinetic_energy_kind(void* h, int kind) { ((Ham*)h)->h.kinetic_energy_kind = (KineticEnergyKind)kind; }
void orc_ham_set_transform(void* h, const double* stds, const double* mean) {
  auto& H = ((Ham*)h)->h;
  H.transformation.set_transform(Vec(stds, stds + H.dim), Vec(mean, mean + H.dim));
}
// LowRankMassMatrix::update (src/transform/low_rank.rs:158-190); vecs = r eigenvectors of length dim, one after the other
int orc_ham_set_lowrank_transform(void* h, const double* stds, const double* mean, const double* vals, const double* vecs,
                                  const double* mean_low_rank, uint64_t r) {
  auto& H = ((Ham*)h)->h;
  size_t d = H.dim;
  return H.transformation.update_lowrank(Vec(stds, stds + d), Vec(mean, mean + d), Vec(vals, vals + r), Vec(vecs, vecs + r * d),
                                         Vec(mean_low_rank, mean_low_rank + d)) ? 1 : 0;
}
void orc_apply_lowrank_transform(const double* vecs, const double* vals, const double* rhs, double* dest, uint64_t d, uint64_t r) {
  apply_lowrank_transform(Vec(vecs, vecs + r * d), Vec(vals, vals + r), rhs, dest, d);
}
void orc_ham_update_diag_draw_grad(void* h, const double* draw_mean, const double* grad_mean, const double* draw_var,
                                   const double* grad_var, int has_fill, double fill, double lo, double hi) {
  auto& H = ((Ham*)h)->h;
  size_t d = H.dim;
  H.transformation.update_diag_draw_grad(Vec(draw_mean, draw_mean + d), Vec(grad_mean, grad_mean + d), Vec(draw_var, draw_var + d),
                                         Vec(grad_var, grad_var + d), has_fill ? std::optional<double>(fill) : std::nullopt, lo, hi);
}
void orc_ham_update_diag_grad(void* h, const double* position, const double* gradient, double fill, double lo, double hi) {
  auto& H = ((Ham*)h)->h;
  size_t d = H.dim;
  H.transformation.update_diag_grad(Vec(position, position + d), Vec(gradient, gradient + d), fill, lo, hi);
}
void orc_ham_update_diag_draw(void* h, const double* draw_mean, const double* draw_var, double scale, int has_fill, double fill,
                              double lo, double hi) {
  auto& H = ((Ham*)h)->h;
  size_t d = H.dim;
  H.transformation.update_diag_draw(Vec(draw_mean, draw_mean + d), Vec(draw_var, draw_var + d), scale,
                                    has_fill ? std::optional<double>(fill) : std::nullopt, lo, hi);
}
void orc_ham_get_transform(void* h, double* stds, double* inv_stds, double* mean, double* logdet, int64_t* id) {
  auto& T = ((Ham*)h)->h.transformation;
  if (stds) std::copy(T.stds.begin(), T.stds.end(), stds);
  if (inv_stds) std::copy(T.inv_stds.begin(), T.inv_stds.end(), inv_stds);
  if (mean) std::copy(T.mean.begin(), T.mean.end(), mean);
  if (logdet) *logdet = T.logdet;
  if (id) *id = T.id;
}

void* orc_point_create(void* h) { return new State(std::make_shared<TransformedPoint>(((Ham*)h)->h.dim)); }
void orc_point_destroy(void* p) { delete (State*)p; }
// which: 0 x, 1 grad_x, 2 z, 3 grad_z, 4 velocity
static Vec& point_vec(TransformedPoint& p, int which) {
  switch (which) {
    case 0: return p.untransformed_position;
    case 1: return p.untransformed_gradient;
    case 2: return p.transformed_position;
    case 3: return p.transformed_gradient;
    default: return p.velocity;
  }
}
void orc_point_get_vec(void* p, int which, double* out) {
  Vec& v = point_vec(**(State*)p, which);
  std::copy(v.begin(), v.end(), out);
}
void orc_point_set_vec(void* p, int which, const double* in) {
  Vec& v = point_vec(**(State*)p, which);
  std::copy(in, in + v.size(), v.begin());
}
void orc_point_get_scalars(void* p, int64_t* idx, double* logp, double* logdet, double* ke, double* e0, int64_t* tid) {
  TransformedPoint& t = **(State*)p;
  if (idx) *idx = t.index_in_trajectory;
  if (logp) *logp = t.logp;
  if (logdet) *logdet = t.logdet;
  if (ke) *ke = t.kinetic_energy;
  if (e0) *e0 = t.initial_energy;
  if (tid) *tid = t.transform_id;
}
void orc_point_set_scalars(void* p, int64_t idx, double logp, double logdet, double ke, double e0, int64_t tid) {
  TransformedPoint& t = **(State*)p;
  t.index_in_trajectory = idx;
  t.logp = logp;
  t.logdet = logdet;
  t.kinetic_energy = ke;
  t.initial_energy = e0;
  t.transform_id = tid;
}
// returns 0 ok, 3 bad init
int orc_ham_init_state(void* h, void* p, const double* x) {
  try {
    *(State*)p = ((Ham*)h)->h.init_state(x);
  } catch (BadInitGrad&) {
    return NUTS_STATUS_FATAL;
  }
  return 0;
}
void orc_ham_init_from_untransformed(void* h, void* p) { ((Ham*)h)->h.init_from_untransformed_position(**(State*)p); }
void orc_ham_init_from_transformed(void* h, void* p) { ((Ham*)h)->h.init_from_transformed_position(**(State*)p); }
uint64_t orc_ham_initialize_trajectory(void* h, void* p, int resample, uint64_t seed, uint64_t stream, uint64_t counter) {
  Rng r(seed, stream, counter);
  ((Ham*)h)->h.initialize_trajectory(**(State*)p, resample != 0, r);
  return r.counter;
}
// dir: +1 forward / -1 backward.  Writes `out` also on an energy divergence (white-box).  Returns status 0/1.
int orc_ham_leapfrog(void* h, void* start, void* out, double step_size, int dir, double energy_baseline, double max_energy_error,
                     double* energy_error) {
  auto& H = ((Ham*)h)->h;
  H.step_size = step_size;
  NullCollector c;
  LeapfrogResult r = H.leapfrog(*(State*)start, dir >= 0 ? Direction::Forward : Direction::Backward, 1.0, energy_baseline,
                                max_energy_error, c);
  *(State*)out = r.state;
  if (energy_error) *energy_error = r.state->energy() - energy_baseline;
  return r.kind == LeapfrogResult::Ok ? 0 : NUTS_STATUS_DIVERGENT_ENERGY;
}
int orc_ham_is_turning(void* h, void* p1, void* p2) { return ((Ham*)h)->h.is_turning(**(State*)p1, **(State*)p2); }

// ---------------- batched sampler: N independent NutsChains, one per stream (src/sampler.rs:1094-1126) ----------------
void* orc_sampler_create(void* model, const nuts_settings_t* settings, uint64_t seed, uint64_t chain_id_offset, uint64_t nchains) {
  auto* s = new Sampler();
  s->model = (LogpFunc*)model;
  s->dim = s->model->dim;
  s->nchains = nchains;
  NutsSettings ns = convert_settings(settings);
  for (uint64_t c = 0; c < nchains; ++c) {
    uint64_t gid = chain_id_offset + c;
    s->chains.emplace_back(new NutsChain(s->model, ns, gid, Rng(seed, gid + 1, 0)));
  }
  s->alive.assign(nchains, 0);
  return s;
}
void orc_sampler_destroy(void* s) { delete (Sampler*)s; }

void orc_sampler_set_position(void* sp, const double* position, int32_t* status, int nthreads) {
  auto* s = (Sampler*)sp;
  parallel_chains(s->nchains, nthreads, [&](uint64_t c) {
    int st = 0;
    try {
      s->chains[c]->set_position(position + c * s->dim);
      s->alive[c] = 1;
    } catch (BadInitGrad&) {
      st = NUTS_STATUS_FATAL;
      s->alive[c] = 0;
    }
    if (status) status[c] = st;
  });
}

// draws_out [n_draws x N x d] (may be NULL); stats SoA [n_draws x N] (members may be NULL). Returns total leapfrogs of these draws.
uint64_t orc_sampler_draw(void* sp, uint64_t n_draws, double* draws_out, const nuts_stats_t* stats, int nthreads) {
  auto* s = (Sampler*)sp;
  uint64_t N = s->nchains, d = s->dim;
  std::atomic<uint64_t> total{0};
  parallel_chains(N, nthreads, [&](uint64_t c) {
    if (!s->alive[c]) return;
    Vec pos(d);
    uint64_t steps = 0;
    for (uint64_t t = 0; t < n_draws; ++t) {
      DrawStats ds = s->chains[c]->draw(pos.data());
      steps += ds.n_steps;
      if (draws_out) std::copy(pos.begin(), pos.end(), draws_out + (t * N + c) * d);
      if (stats) {
        size_t k = t * N + c;
        if (stats->depth) stats->depth[k] = ds.depth;
        if (stats->maxdepth_reached) stats->maxdepth_reached[k] = ds.maxdepth_reached;
        if (stats->index_in_trajectory) stats->index_in_trajectory[k] = ds.index_in_trajectory;
        if (stats->logp) stats->logp[k] = ds.logp;
        if (stats->energy) stats->energy[k] = ds.energy;
        if (stats->energy_error) stats->energy_error[k] = ds.energy_error;
        if (stats->diverging) stats->diverging[k] = ds.diverging;
        if (stats->step_size) stats->step_size[k] = ds.step_size;
        if (stats->step_size_bar) stats->step_size_bar[k] = ds.step_size_bar;
        if (stats->mean_tree_accept) stats->mean_tree_accept[k] = ds.mean_tree_accept;
        if (stats->mean_tree_accept_sym) stats->mean_tree_accept_sym[k] = ds.mean_tree_accept_sym;
        if (stats->n_steps) stats->n_steps[k] = ds.n_steps;
        if (stats->max_energy_error) stats->max_energy_error[k] = ds.max_energy_error;
        if (stats->tuning) stats->tuning[k] = ds.tuning;
        if (stats->fisher_distance) stats->fisher_distance[k] = ds.fisher_distance;
      }
    }
    total += steps;
  });
  return total.load();
}

void orc_sampler_get_state(void* sp, double* position, double* step_size, double* stds, double* mean, uint64_t* rng_counter) {
  auto* s = (Sampler*)sp;
  for (uint64_t c = 0; c < s->nchains; ++c) {
    NutsChain& ch = *s->chains[c];
    size_t d = s->dim;
    if (position) std::copy(ch.state->untransformed_position.begin(), ch.state->untransformed_position.end(), position + c * d);
    if (step_size) step_size[c] = ch.hamiltonian.step_size;
    if (stds) std::copy(ch.hamiltonian.transformation.stds.begin(), ch.hamiltonian.transformation.stds.end(), stds + c * d);
    if (mean) std::copy(ch.hamiltonian.transformation.mean.begin(), ch.hamiltonian.transformation.mean.end(), mean + c * d);
    if (rng_counter) rng_counter[c] = ch.rng.counter;
  }
}
// The complete state a NutsChain carries between two draws (include/nuts_b200.h nuts_chain_state_t): read it out / overwrite
// it.  Teacher-forced parity tests inject the GPU's state after draw t here and compare draw t+1.
void orc_sampler_get_chain_state(void* sp, const nuts_chain_state_t* o) {
  auto* s = (Sampler*)sp;
  const size_t d = s->dim;
  auto putv = [&](double* dst, uint64_t c, const Vec& v) {
    if (dst) std::copy(v.begin(), v.end(), dst + c * d);
  };
  for (uint64_t c = 0; c < s->nchains; ++c) {
    NutsChain& ch = *s->chains[c];
    const TransformedPoint& p = *ch.state;
    const DiagMassMatrix& m = ch.hamiltonian.transformation;
    const DiagAdaptStrategy& a = ch.strategy.mass_matrix_adapt;
    putv(o->position, c, p.untransformed_position);
    putv(o->gradient, c, p.untransformed_gradient);
    putv(o->transformed_position, c, p.transformed_position);
    putv(o->transformed_gradient, c, p.transformed_gradient);
    putv(o->stds, c, m.stds);
    putv(o->inv_stds, c, m.inv_stds);
    putv(o->mean, c, m.mean);
    putv(o->draw_mean, c, a.exp_variance_draw.mean);
    putv(o->draw_var, c, a.exp_variance_draw.variance);
    putv(o->grad_mean, c, a.exp_variance_grad.mean);
    putv(o->grad_var, c, a.exp_variance_grad.variance);
    putv(o->draw_mean_bg, c, a.exp_variance_draw_bg.mean);
    putv(o->draw_var_bg, c, a.exp_variance_draw_bg.variance);
    putv(o->grad_mean_bg, c, a.exp_variance_grad_bg.mean);
    putv(o->grad_var_bg, c, a.exp_variance_grad_bg.variance);
#define PUT(field, val) \
  if (o->field) o->field[c] = (val)
    PUT(logp, p.logp);
    PUT(point_logdet, p.logdet);
    PUT(point_transform_id, p.transform_id);
    PUT(mass_matrix_logdet, m.logdet);
    PUT(mass_matrix_id, m.id);
    PUT(step_size, ch.hamiltonian.step_size);
    const auto& da = ch.strategy.step_size.adaptation;
    const auto& ad = ch.strategy.step_size.adam;  // Adam shares the record: log_step, m, v, t (log_step_adapted mirrors log_step)
    PUT(da_log_step, ad ? ad->log_step : da ? da->log_step : 0.);
    PUT(da_log_step_adapted, ad ? ad->log_step : da ? da->log_step_adapted : 0.);
    PUT(da_hbar, ad ? ad->m : da ? da->hbar : 0.);
    PUT(da_mu, ad ? ad->v : da ? da->mu : 0.);
    PUT(da_count, ad ? ad->t : da ? da->count : 0);
    PUT(foreground_count, a.exp_variance_draw.count);
    PUT(background_count, a.exp_variance_draw_bg.count);
    PUT(tuning, (uint8_t)ch.strategy.tuning);
    PUT(has_initial_mass_matrix, (uint8_t)ch.strategy.has_initial_mass_matrix);
    PUT(last_update, ch.strategy.last_update);
    PUT(current_window_size, ch.strategy.current_window_size);
    PUT(draw_count, ch.draw_count);
    PUT(rng_counter, ch.rng.counter);
    PUT(total_leapfrogs, ch.hamiltonian.n_leapfrogs);
    PUT(alive, (uint8_t)s->alive[c]);
#undef PUT
  }
}
void orc_sampler_set_chain_state(void* sp, const nuts_chain_state_t* in) {
  auto* s = (Sampler*)sp;
  const size_t d = s->dim;
  auto getv = [&](Vec& v, const double* src, uint64_t c) { std::copy(src + c * d, src + (c + 1) * d, v.begin()); };
  for (uint64_t c = 0; c < s->nchains; ++c) {
    NutsChain& ch = *s->chains[c];
    ch.state = ch.hamiltonian.pool.new_state();  // never write through a State that a collector may still share
    TransformedPoint& p = *ch.state;
    DiagMassMatrix& m = ch.hamiltonian.transformation;
    DiagAdaptStrategy& a = ch.strategy.mass_matrix_adapt;
    getv(p.untransformed_position, in->position, c);
    getv(p.untransformed_gradient, in->gradient, c);
    getv(p.transformed_position, in->transformed_position, c);
    getv(p.transformed_gradient, in->transformed_gradient, c);
    std::fill(p.velocity.begin(), p.velocity.end(), 0.);
    p.logp = in->logp[c];
    p.logdet = in->point_logdet[c];
    p.transform_id = in->point_transform_id[c];
    p.index_in_trajectory = 0;
    p.kinetic_energy = 0.;
    p.initial_energy = 0.;
    p.step_size_factor = 1.0;
    getv(m.stds, in->stds, c);
    getv(m.inv_stds, in->inv_stds, c);
    getv(m.mean, in->mean, c);
    m.logdet = in->mass_matrix_logdet[c];
    m.id = in->mass_matrix_id[c];
    ch.hamiltonian.step_size = in->step_size[c];
    auto& da = ch.strategy.step_size.adaptation;
    if (da) {
      da->log_step = in->da_log_step[c];
      da->log_step_adapted = in->da_log_step_adapted[c];
      da->hbar = in->da_hbar[c];
      da->mu = in->da_mu[c];
      da->count = in->da_count[c];
    }
    if (auto& ad = ch.strategy.step_size.adam) {
      ad->log_step = in->da_log_step[c];
      ad->m = in->da_hbar[c];
      ad->v = in->da_mu[c];
      ad->t = in->da_count[c];
    }
    getv(a.exp_variance_draw.mean, in->draw_mean, c);
    getv(a.exp_variance_draw.variance, in->draw_var, c);
    getv(a.exp_variance_grad.mean, in->grad_mean, c);
    getv(a.exp_variance_grad.variance, in->grad_var, c);
    getv(a.exp_variance_draw_bg.mean, in->draw_mean_bg, c);
    getv(a.exp_variance_draw_bg.variance, in->draw_var_bg, c);
    getv(a.exp_variance_grad_bg.mean, in->grad_mean_bg, c);
    getv(a.exp_variance_grad_bg.variance, in->grad_var_bg, c);
    a.exp_variance_draw.count = a.exp_variance_grad.count = in->foreground_count[c];
    a.exp_variance_draw_bg.count = a.exp_variance_grad_bg.count = in->background_count[c];
    ch.strategy.tuning = in->tuning[c] != 0;
    ch.strategy.has_initial_mass_matrix = in->has_initial_mass_matrix[c] != 0;
    ch.strategy.last_update = in->last_update[c];
    ch.strategy.current_window_size = in->current_window_size[c];
    ch.draw_count = in->draw_count[c];
    ch.rng.counter = in->rng_counter[c];
    ch.hamiltonian.n_leapfrogs = in->total_leapfrogs[c];
    s->alive[c] = in->alive[c] ? 1 : 0;
  }
}
// ---------------- low-rank estimator (src/transform/adapt/low_rank.rs:73-268)
// spd_mean of two n x n matrices (row-major = column-major: symmetric); returns 0 on failure
int orc_lowrank_spd_mean(const double* cov_draws, const double* cov_grads, uint64_t n, double* out) {
  oracle::lowrank::Mat a(n, n), b(n, n), o;
  std::copy(cov_draws, cov_draws + n * n, a.a.begin());
  std::copy(cov_grads, cov_grads + n * n, b.a.begin());
  if (!oracle::lowrank::spd_mean(a, b, o)) return 0;
  std::copy(o.a.begin(), o.a.end(), out);
  return 1;
}
// estimate_mass_matrix: draws, grads [k][n] row-major; vals [k] ascending, vecs [k][k] (eigenvector j = column j, row-major)
int orc_lowrank_estimate_mass_matrix(const double* draws, const double* grads, uint64_t k, uint64_t n, double gamma, double* vals, double* vecs) {
  oracle::lowrank::Mat d(k, n), g(k, n), v;
  for (uint64_t i = 0; i < k; ++i)
    for (uint64_t j = 0; j < n; ++j) d(i, j) = draws[i * n + j], g(i, j) = grads[i * n + j];
  std::vector<double> w;
  if (!oracle::lowrank::estimate_mass_matrix(d, g, gamma, w, v)) return 0;
  for (uint64_t i = 0; i < k; ++i) {
    vals[i] = w[i];
    for (uint64_t j = 0; j < k; ++j) vecs[i * k + j] = v(i, j);
  }
  return 1;
}
// compute_update: draws, grads [n][d]; outputs stds / mean / mu [d], vals [<= 2 n], vecs [<= 2 n][d]; returns the rank or -1
int64_t orc_lowrank_compute_update(const double* draws, const double* grads, uint64_t n, uint64_t d, double gamma, double cutoff, double* stds,
                                   double* mean, double* vals, double* vecs, double* mu) {
  oracle::lowrank::Update u;
  if (!oracle::lowrank::compute_update(draws, grads, n, d, gamma, cutoff, u)) return -1;
  std::copy(u.stds.begin(), u.stds.end(), stds);
  std::copy(u.mean.begin(), u.mean.end(), mean);
  std::copy(u.vals.begin(), u.vals.end(), vals);
  std::copy(u.vecs.begin(), u.vecs.end(), vecs);
  std::copy(u.mu.begin(), u.mu.end(), mu);
  return (int64_t)u.vals.size();
}
// per chain LowRankMassMatrix::update with caller-supplied values; layouts as nuts_sampler_set_lowrank_transform
void orc_sampler_set_lowrank_transform(void* sp, const double* stds, const double* mean, uint64_t rank_max, const double* vals,
                                       const double* vecs, const int32_t* rank, const double* mean_low_rank, uint8_t* accepted) {
  auto* s = (Sampler*)sp;
  const size_t d = s->dim;
  for (uint64_t c = 0; c < s->nchains; ++c) {
    if (!s->alive[c]) {
      if (accepted) accepted[c] = 0;
      continue;
    }
    const size_t r = rank_max == 0 ? 0 : (rank ? (size_t)rank[c] : (size_t)rank_max);
    const double* vc = vals ? vals + c * rank_max : nullptr;
    const double* uc = vecs ? vecs + c * rank_max * d : nullptr;
    bool ok = false;
    try {
      ok = s->chains[c]->set_lowrank_transform(Vec(stds + c * d, stds + (c + 1) * d), Vec(mean + c * d, mean + (c + 1) * d),
                                               r ? Vec(vc, vc + r) : Vec(), r ? Vec(uc, uc + r * d) : Vec(),
                                               Vec(mean_low_rank + c * d, mean_low_rank + (c + 1) * d));
    } catch (const BadInitGrad&) {
      s->alive[c] = 0;
    }
    if (accepted) accepted[c] = ok ? 1 : 0;
  }
}
void orc_sampler_set_step_size(void* sp, const double* step_size) {
  auto* s = (Sampler*)sp;
  for (uint64_t c = 0; c < s->nchains; ++c) s->chains[c]->hamiltonian.step_size = step_size[c];
}
// total leapfrog calls (trees + step-size searches) and logp evaluations, summed over chains
void orc_sampler_counters(void* sp, uint64_t* leapfrogs, uint64_t* logp_evals) {
  auto* s = (Sampler*)sp;
  uint64_t l = 0, e = 0;
  for (auto& ch : s->chains) {
    l += ch->hamiltonian.n_leapfrogs;
    e += ch->hamiltonian.n_logp_evals;
  }
  if (leapfrogs) *leapfrogs = l;
  if (logp_evals) *logp_evals = e;
}

void orc_settings_default(nuts_settings_t* s) {
  std::memset(s, 0, sizeof(*s));
  NutsSettings d;
  s->num_tune = d.num_tune;
  s->num_draws = d.num_draws;
  s->maxdepth = d.maxdepth;
  s->mindepth = d.mindepth;
  s->max_energy_error = d.max_energy_error;
  s->check_turning = 1;
  s->num_chains = d.num_chains;
  s->seed = d.seed;
  s->extra_doublings = 0;
  s->trajectory_kind = NUTS_KINETIC_EUCLIDEAN;
  auto& a = s->adapt_options;
  a.early_window = 0.3;
  a.step_size_window = 0.15;
  a.mass_matrix_switch_freq = 80;
  a.early_mass_matrix_switch_freq = 10;
  a.mass_matrix_update_freq = 1;
  a.mass_matrix_window_growth = 1.5;
  a.mass_matrix_options.store_mass_matrix = 0;
  a.mass_matrix_options.use_grad_based_estimate = 1;
  a.step_size_settings.target_accept = 0.8;
  a.step_size_settings.initial_step = 0.1;
  a.step_size_settings.has_jitter = 1;
  a.step_size_settings.jitter = 0.1;
  a.step_size_settings.adapt_options.method = NUTS_STEPSIZE_DUAL_AVERAGE;
  a.step_size_settings.adapt_options.adam = nuts_adam_options_t{0.9, 0.999, 1e-8, 0.05};
  a.step_size_settings.adapt_options.dual_average = {0.75, 10., 0.05, 3.14159265358979323846};
}

}  // extern "C"
