// chain_engine_v2.cuh — the DECOUPLED whole-draw engine: scalar tree logic and vector work of a chain run on different warps.
//
// Why: in the register-resident engine (chain_engine.cuh) every thread of a team executes the scalar tree logic redundantly
// (energies, exp / log1p of the multinomial weights, Philox, reference counts).  That costs ~110 registers per thread (=> 4
// chains per SM, 592 of the 1024 chains of config 2 resident, a 2-wave launch) and puts ~1.5 us of dependent scalar latency
// between two leapfrogs of a chain.  Here
//   * a TEAM of TPC threads only does vector work: leapfrog, checkpoint store and the U-turn dot products of the merges a leaf
//     completes - whose operands follow from the leaf index alone (binary counter) - bundled into one ring entry per leaf;
//   * the LEADER warp of the CTA runs the scalar side, one LANE per team (reference src/nuts.rs:108-245, 281-388 and
//     src/stepsize/dual_avg.rs:131-158): it consumes the entries up to V2_K leaves behind the vector warps, hands out the
//     checkpoint slots ahead of time and only talks back at doubling boundaries (direction, accept / abort) or to stop a doubling.
// A team needs ~130 registers per thread, so C = 7 teams + the leader share one SM: all 1024 chains of config 2 are resident
// in ONE wave (148 x 7 = 1036 teams), and a chain's critical path per leapfrog is the vector work alone.
//
// Results are bit-identical to the 64x16 tiling of chain_engine.cuh: same per-thread partial sums, same warp reduce-scatter,
// same order of the per-warp partials, same scalar formulas in the same order, same RNG consumption.
#pragma once
#include "chain_engine.cuh"

namespace nb {

__device__ __forceinline__ bool v2_turn_eval(double sP, double sQ, int dir) {
  return dir ? ((sP < 0.) | (sQ < 0.)) : ((sP > 0.) | (sQ > 0.));
}

// The leader's transcendental / RNG work goes through ONE copy of each routine (the kernel's hot code has to stay inside the
// 32 KB instruction cache next to the unrolled vector loops).
__device__ __forceinline__ uint2 v2_philox01(uint64_t seed, uint64_t stream, uint64_t counter) { return philox01(seed, stream, counter); }
__device__ __forceinline__ double v2_stream_f64(uint64_t seed, uint64_t stream, uint64_t counter) {  // == stream_f64
  const uint2 b = v2_philox01(seed, stream, counter);
  return u53((uint64_t)b.x | ((uint64_t)b.y << 32));
}
// NutsTree::merge_into weights (nuts.rs:189-203): log_size of the merged tree and whether the other tree's draw is taken.
// is_main: self_log_size is the OLD log_size of self, else that of the merged tree.  Everything by value: a reference to a
// member of the leader's state would pin that state in local memory.
struct V2Merge {
  double total;
  int take_other, consumed;  // consumed: random numbers drawn (0 or 1)
};
static __device__ __noinline__ V2Merge v2_merge_into(double self_ls, double other_ls, bool is_main, uint64_t seed, uint64_t stream, uint64_t rng) {
  V2Merge r;
  r.total = logaddexp(self_ls, other_ls);
  const double ref = is_main ? self_ls : r.total;
  r.consumed = 0;
  bool take = other_ls >= ref;
  if (!take) {
    take = v2_stream_f64(seed, stream, rng) < exp(other_ls - ref);
    r.consumed = 1;
  }
  r.take_other = take ? 1 : 0;
  return r;
}

// Scalar side of one chain's trees; lives in the registers of ONE lane of the leader warp.
template <int W>
struct LeaderLane {
  enum { ST_IDLE = 0, ST_LEAF = 1, ST_EXITED = 2 };
  const EngineParams& P;
  V2Ctl& c;
  const double* ring;  // [V2_K][W][V2_NV]
  int st;
  unsigned start_seen, epoch;
  uint64_t stream, rng;
  double E0, pt_logdet;
  // AcceptanceRateCollector
  double acc_sum, acc_sym_sum, max_ee;
  uint64_t acc_count;
  // main tree (weights: linear domain while `lin`, see device_common.cuh; the reference's log sizes otherwise)
  double ls_main, draw_energy;
  bool lin;
  int depth, idx_left, idx_right, draw_slot, draw_idx;
  int end_slot_left, end_slot_right;  // checkpoint slots that hold the two ends of the main tree (-1: the initial point)
  uint64_t free_mask, rc_lo, rc_hi;
  // doubling under construction
  int D, dir, check, idx_cur;
  unsigned nleaf, i;
  uint64_t mindepth, maxdepth, extra_left;
  bool extra_mode, reached_maxdepth, diverging;

  __device__ __forceinline__ LeaderLane(const EngineParams& p, V2Ctl& ctl, const double* r)
      : P(p), c(ctl), ring(r), st(ST_IDLE), start_seen(0), epoch(0) {}

  __device__ __forceinline__ int rc_get(int s) const { return (int)(((s < 32 ? rc_lo : rc_hi) >> (2 * (s & 31))) & 3ull); }
  __device__ __forceinline__ void rc_add(int s, int delta) {
    const uint64_t inc = (uint64_t)(long long)delta << (2 * (s & 31));
    if (s < 32) rc_lo += inc;
    else rc_hi += inc;
  }
  __device__ __forceinline__ int alloc_slot() {
    int s = __ffsll((long long)free_mask) - 1;
    free_mask &= ~(1ull << s);
    return s;
  }
  __device__ __forceinline__ void unref(int s) {
    rc_add(s, -1);
    if (rc_get(s) == 0) free_mask |= (1ull << s);
  }
  __device__ __forceinline__ bool rng_bool() { return (v2_philox01(P.seed, stream, rng++).x & 1u) != 0; }  // == stream_bool
  __device__ __forceinline__ double val(const double* e, int k) const {  // sum of the per-warp partials, warp 0 first
    double t = e[k];
#pragma unroll
    for (int w = 1; w < W; ++w) t += e[w * V2_NV + k];
    return t;
  }

  // AcceptanceRateCollector::register_leapfrog (dual_avg.rs:131-158); returns exp(E0 - energy) = the leaf's linear weight
  __device__ __forceinline__ double register_leapfrog(double energy, bool divergent) {
    double ed = 0.;
    if (divergent) {
      max_ee = -INFINITY;
    } else {
      double diff = E0 - energy;
      ed = accept_exp(diff);
      const double e = diff < 0. ? ed : 1.0;
      acc_sum += e;
      acc_sym_sum += 2. * e / (1. + ed);
      if (fabs(diff) > fabs(max_ee)) max_ee = diff;
    }
    acc_count += 1;
    return ed;
  }
  __device__ __forceinline__ double next_f64() { return v2_stream_f64(P.seed, stream, rng++); }
  // leave the linear domain for the rest of this draw (chain_engine.cuh weights_to_log_domain): pending levels = set bits of i
  __device__ __forceinline__ void weights_to_log_domain() {
    lin = false;
    ls_main = log_noinline(ls_main);
    if (i & 1u) c.A_ls[0] = c.A_log0;
#pragma unroll 1
    for (int l = 1; l < V2_NT; ++l)
      if ((i >> l) & 1u) c.A_ls[l] = log_noinline(c.A_ls[l]);
  }

  __device__ __forceinline__ void command(int kind, int accepted) {
    c.cmd_kind = kind;
    c.cmd_dir = dir;
    c.cmd_check = check;
    c.cmd_depth = depth;
    c.cmd_prev_accepted = accepted;
    epoch = v2_ld(&c.cmd_seq) + 1;
    c.cons = (epoch & 0xFFFFFu) << 12;
    __threadfence_block();
    c.cmd_seq = epoch;
  }

  // the tree is finished: result record + V2_CMD_TREE_DONE
  __device__ __forceinline__ void finish_tree(int accepted) {
    c.acc_sum = acc_sum;
    c.acc_sym_sum = acc_sym_sum;
    c.max_energy_error = max_ee;
    c.acc_count = acc_count;
    c.rng_out = rng;
    c.depth = depth;
    c.draw_slot = draw_slot;
    c.draw_idx = draw_idx;
    c.draw_energy = draw_energy;
    c.reached_maxdepth = reached_maxdepth ? 1 : 0;
    c.diverging = diverging ? 1 : 0;
    command(V2_CMD_TREE_DONE, accepted);
    st = ST_IDLE;
  }

  // nuts::draw loop head (nuts.rs:333-374): decide the next doubling or finish.  `accepted`: the previous doubling was merged.
  __device__ __forceinline__ void next_doubling(int accepted) {
    const SettingsDev& S = P.s;
    if (!extra_mode) {
      if (!((uint64_t)depth < maxdepth)) {
        reached_maxdepth = true;
        finish_tree(accepted);
        return;
      }
      dir = rng_bool() ? 1 : 0;  // hamiltonian.rs:111-119: true => Forward
      check = (S.check_turning && !((uint64_t)depth < mindepth)) ? 1 : 0;
    } else {
      if (extra_left == 0) {
        finish_tree(accepted);
        return;
      }
      extra_left -= 1;
      check = 0;
    }
    // NutsTree::extend prologue (chain_engine.cuh extend()): fresh slot pool except the main tree's draw
    D = depth;
    nleaf = 1u << D;
    i = 0;
    free_mask = P.P >= 64 ? ~0ull : ((1ull << P.P) - 1ull);
    rc_lo = rc_hi = 0;
    if (draw_slot >= 0) {
      free_mask &= ~(1ull << draw_slot);
      rc_add(draw_slot, 1);
    }
    if (end_slot_left >= 0) free_mask &= ~(1ull << end_slot_left);  // never referenced by the half under construction
    if (end_slot_right >= 0) free_mask &= ~(1ull << end_slot_right);
    idx_cur = dir ? idx_right : idx_left;
    const unsigned npub = nleaf < (unsigned)V2_K ? nleaf : (unsigned)V2_K;
    for (unsigned j = 0; j < npub; ++j) c.slot_ring[j & 7] = (signed char)alloc_slot();
    command(V2_CMD_DOUBLING, accepted);
    st = ST_LEAF;
  }

  // a draw starts: NutsTree::new (nuts.rs:94-105) + register_init
  __device__ __forceinline__ void begin_draw(int chain) {
    const SettingsDev& S = P.s;
    E0 = c.E0;
    pt_logdet = c.pt_logdet;
    rng = c.rng;
    const double step = c.step;
    stream = P.chain_offset + (uint64_t)chain + 1;
    acc_sum = 0.;
    acc_sym_sum = 0.;
    acc_count = 0;
    max_ee = 0.;
    lin = true;
    ls_main = 1.;  // log_size 0
    depth = 0;
    idx_left = idx_right = 0;
    draw_slot = -1;
    end_slot_left = end_slot_right = -1;
    draw_energy = E0;
    draw_idx = 0;
    mindepth = S.mindepth;
    maxdepth = S.maxdepth;
    if (S.has_target_time) {  // nuts.rs:300-320
      uint64_t max_steps = (uint64_t)ceil(S.target_time / step);
      mindepth = max((uint64_t)floor(log2((double)max_steps)), S.mindepth);
      maxdepth = min(max((uint64_t)ceil(log2((double)max_steps)), mindepth), S.maxdepth);
    }
    extra_mode = false;
    extra_left = 0;
    reached_maxdepth = false;
    diverging = false;
    dir = 0;
    check = 0;
    next_doubling(0);
  }

  // the doubling ended with ExtendResult r (nuts.rs:80-91) -> loop tail of nuts::draw
  __device__ __forceinline__ void doubling_done(int r, int accepted) {
    if (r == EXT_DIVERGING) {
      diverging = true;
      finish_tree(accepted);
      return;
    }
    if (!extra_mode && r == EXT_TURNING) {
      extra_mode = true;
      extra_left = P.s.extra_doublings;
    }
    next_doubling(accepted);
  }

  // consume the entry of leaf i (NutsTree::extend, nuts.rs:108-170, as the binary counter of chain_engine.cuh extend())
  __device__ __forceinline__ void process_leaf() {
    const double* e = ring + (size_t)(i % V2_K) * W * V2_NV;
    const int sign = dir ? 1 : -1;
    const double logp_new = val(e, 0), ke_new = 0.5 * val(e, 1);
    const double energy = ke_new - (logp_new + pt_logdet);
    const double energy_error = energy - E0;
    const bool divergent = (energy_error > P.s.max_energy_error) | !isfinite(energy_error);
    const double leaf_w = register_leapfrog(energy, divergent);
    if (divergent) {
      doubling_done(EXT_DIVERGING, 0);
      return;
    }
    idx_cur += sign;
    const int s = c.slot_ring[i & 7];
    rc_add(s, 3);  // roles: first-of-B, draw-of-B, last-of-B (the newest leaf)
    int B_first = s, B_draw = s, B_draw_idx = idx_cur;
    const double leaf_ls = -energy_error;  // log weight of the leaf
    if (lin && fabs(leaf_ls) > LIN_WEIGHT_LIMIT) weights_to_log_domain();
    double B_ls = lin ? leaf_w : leaf_ls, B_draw_energy = energy;
    int t = __ffs(~i) - 1;
    if (t > D) t = D;
    int g = 0;
    for (int l = 0; l < t; ++l) {
      const int Af = c.A_first[l], Al = c.A_last[l];
      bool turning = false;
      if (check) {
        if (l == 0) {
          turning = v2_turn_eval(val(e, 2), val(e, 3), dir);
        } else {
          const int o = 4 + 6 * g;
          turning = v2_turn_eval(val(e, o), val(e, o + 1), dir) | v2_turn_eval(val(e, o + 2), val(e, o + 3), dir) |
                    v2_turn_eval(val(e, o + 4), val(e, o + 5), dir);
          ++g;
        }
      }
      // merge_into, non-main (nuts.rs:172-207): self_log_size = log_size of the merged tree
      bool take_B;
      double total;
      if (lin) {  // same arithmetic as chain_engine.cuh extend()
        const double WA = c.A_ls[l], WB = B_ls;
        total = WA + WB;
        if (WA >= WB * 0x1p-40) take_B = next_f64() * total < WB;
        else take_B = lin_reference_shortcut(WA, WB) || (next_f64() * total < WB);
      } else {
        const V2Merge mg = v2_merge_into(c.A_ls[l], B_ls, false, P.seed, stream, rng);
        rng += (uint64_t)mg.consumed;
        take_B = mg.take_other != 0;
        total = mg.total;
      }
      if (take_B) {
        unref(c.A_draw[l]);
      } else {
        unref(B_draw);
        B_draw = c.A_draw[l];
        B_draw_energy = c.A_draw_energy[l];
        B_draw_idx = c.A_draw_idx[l];
      }
      unref(B_first);
      B_first = Af;
      unref(Al);
      B_ls = total;
      if (turning) {  // inner turn: the old tree is returned unchanged (nuts.rs:131-133)
        doubling_done(EXT_TURNING, 0);
        return;
      }
    }
    if (i + 1 < nleaf) {
      c.A_first[t] = (signed char)B_first;
      c.A_last[t] = (signed char)s;  // the last-of-B reference moves to the pending sub-tree
      c.A_ls[t] = B_ls;
      if (t == 0) c.A_log0 = leaf_ls;
      c.A_draw[t] = (signed char)B_draw;
      c.A_draw_energy[t] = B_draw_energy;
      c.A_draw_idx[t] = B_draw_idx;
      // hand out the slot of leaf i + V2_K, then release entry i
      if (i + V2_K < nleaf) c.slot_ring[(i + V2_K) & 7] = (signed char)alloc_slot();
      i += 1;
      __threadfence_block();
      c.cons = ((epoch & 0xFFFFFu) << 12) | i;
      return;
    }
    // top-level merge of the main tree (A) with the finished half (B)
    bool turning = false;
    if (check) {
      const int o = 4 + 6 * g;
      turning = v2_turn_eval(val(e, o), val(e, o + 1), dir);
      if (D > 0) turning = turning | v2_turn_eval(val(e, o + 2), val(e, o + 3), dir) | v2_turn_eval(val(e, o + 4), val(e, o + 5), dir);
    }
    bool take;  // is_main: self_log_size = old log_size
    double total;
    if (lin) {
      take = (B_ls >= ls_main) || (next_f64() * ls_main < B_ls);
      total = ls_main + B_ls;
    } else {
      const V2Merge mg = v2_merge_into(ls_main, B_ls, true, P.seed, stream, rng);
      rng += (uint64_t)mg.consumed;
      take = mg.take_other != 0;
      total = mg.total;
    }
    if (take) {
      draw_slot = B_draw;
      draw_energy = B_draw_energy;
      draw_idx = B_draw_idx;
    }
    ls_main = total;
    depth += 1;
    if (dir) {
      idx_right = idx_cur;
      end_slot_right = s;
    } else {
      idx_left = idx_cur;
      end_slot_left = s;
    }
    doubling_done(turning ? EXT_TURNING : EXT_OK, 1);
  }

  // one polling step; returns true when it made progress
  __device__ __forceinline__ bool poll(int chain_of_team) {
    if (st == ST_IDLE) {
      if (v2_ld(&c.exit_flag)) {
        st = ST_EXITED;
        return true;
      }
      const unsigned ss = v2_ld(&c.start_seq);
      if (ss == start_seen) return false;
      start_seen = ss;
      __threadfence_block();
      begin_draw(chain_of_team);
      return true;
    }
    if (st == ST_LEAF) {
      const unsigned want = ((epoch & 0xFFFFFu) << 12) | (i + 1);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const unsigned pv = v2_ld(&c.prod[w]);
        if ((pv >> 12) != (want >> 12) || (pv & 0xFFFu) < (want & 0xFFFu)) return false;
      }
      __threadfence_block();
      process_leaf();
      return true;
    }
    return false;
  }
};

// shared memory of one CTA: C teams x [sigma | mean | TreeTables] , model [mu | prec] , C rings , C control blocks ,
// C reduction scratches , C chain ids
template <int TPC, int EPT, int C>
struct V2Layout {
  static constexpr int W = TPC / 32;
  // the CTA-wide copy of the model parameters only when it leaves room for the teams (dim ~ 10^4: 160 KB of sigma | mean alone)
  static constexpr bool MODEL_SHARED = (size_t)(2 * C + 2) * TPC * EPT * sizeof(double) <= 200 * 1024;
  // rows padded to TPC*EPT; sigma | mean in shared memory; grad_z is recomputed, never stored on chip
  static constexpr int SMF = SM_MASS | SM_EXACT | (MODEL_SHARED ? 0 : SM_MODEL_GLOBAL);
  static constexpr size_t team_bytes = team_smem_bytes<TPC, EPT, SMF>();
  static constexpr size_t off_model = (size_t)C * team_bytes;
  static constexpr size_t off_ring = off_model + (MODEL_SHARED ? 2 * (size_t)TPC * EPT * sizeof(double) : 0);
  static constexpr size_t ring_bytes = (size_t)V2_K * W * V2_NV * sizeof(double);
  static constexpr size_t off_ctl = off_ring + (size_t)C * ring_bytes;
  static constexpr size_t ctl_bytes = (sizeof(V2Ctl) + 15) / 16 * 16;
  static constexpr size_t off_scratch = off_ctl + (size_t)C * ctl_bytes;
  static constexpr size_t scratch_bytes = 2 * (size_t)W * REDUCE_MAXK * sizeof(double);
  static constexpr size_t off_chain = off_scratch + (size_t)C * scratch_bytes;
  // cp.async ring of the model parameters, one per team, only when they are not resident in shared memory
  static constexpr size_t off_mring = (off_chain + (size_t)C * 2 * sizeof(int) + 15) / 16 * 16;
  static constexpr size_t mring_bytes = MODEL_SHARED ? 0 : (size_t)V2_MRING * 2 * TPC * sizeof(double);
  static constexpr size_t total = off_mring + (size_t)C * mring_bytes;
};

// One kernel for both Chain::set_position (mode 0) and n_draws x Chain::draw (mode 1).  Warps 0 .. NL-1 = leaders, then C teams of TPC threads.
// NL leader warps share the teams (leader warp lw, lane j -> team lw + j * NL): a lane only advances when ITS team has an entry
// ready, so the lanes of one leader warp mostly run one at a time and a single leader warp saturates.
template <int TPC, int EPT, int C, int MODEL, int NL>
__global__ void NB_KERNEL_BOUNDS(32 * NL + C * TPC, 1) nuts_chain_kernel_v2(const __grid_constant__ EngineParams P) {
  using L = V2Layout<TPC, EPT, C>;
  constexpr int W = TPC / 32;
  constexpr int SMF = L::SMF;
  static_assert(C <= 15 && C <= 32, "one named barrier and one leader lane per team");
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  double* model_smem = reinterpret_cast<double*>(dyn_smem + L::off_model);
  // prologue: control blocks and the CTA-wide copy of the model parameters
  for (int k = threadIdx.x; k < (int)(C * L::ctl_bytes / 4); k += blockDim.x) reinterpret_cast<unsigned*>(dyn_smem + L::off_ctl)[k] = 0u;
  if (L::MODEL_SHARED)
    for (int k = threadIdx.x; k < TPC * EPT; k += blockDim.x) {
      model_smem[k] = k < P.d ? P.model.mu[k] : 0.0;
      model_smem[TPC * EPT + k] = k < P.d ? P.model.prec[k] : 0.0;
    }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp < NL) {
    // ---------------------------------------------------------------- leader: lane c = scalar side of team c
    const int lane = threadIdx.x & 31;
    const int my_team = warp + lane * NL;
    const int cidx = my_team < C ? my_team : 0;
    V2Ctl& ctl = *reinterpret_cast<V2Ctl*>(dyn_smem + L::off_ctl + (size_t)cidx * L::ctl_bytes);
    const volatile int* chain_of_team = reinterpret_cast<const volatile int*>(dyn_smem + L::off_chain) + 2 * cidx;
    const double* my_ring = reinterpret_cast<const double*>(dyn_smem + L::off_ring + (size_t)cidx * L::ring_bytes);
    LeaderLane<W> lead(P, ctl, my_ring);
    if (my_team >= C) lead.st = LeaderLane<W>::ST_EXITED;
#ifdef NB_PHASE_TIMING
    long long t_busy = 0, t_all0 = clock64(), n_polls = 0, n_prog = 0, n_multi = 0;
#endif
    for (;;) {
      bool progress = false;
#ifdef NB_PHASE_TIMING
      const long long t0 = clock64();
#endif
      if (lead.st != LeaderLane<W>::ST_EXITED) progress = lead.poll(*chain_of_team);
      const unsigned alive = __ballot_sync(0xffffffffu, lead.st != LeaderLane<W>::ST_EXITED);
#ifdef NB_PHASE_TIMING
      const unsigned pm = __ballot_sync(0xffffffffu, progress);
      n_polls += 1;
      if (pm) {
        t_busy += clock64() - t0;
        n_prog += 1;
        n_multi += __popc(pm);
      }
#endif
      if (alive == 0u) break;
      if (!__any_sync(0xffffffffu, progress)) __nanosleep(NB_V2_SLEEP);
    }
#ifdef NB_PHASE_TIMING
    if (lane == 0 && P.phase_clocks) {  // leader statistics of this CTA: busy cycles, total cycles, polls, polls with progress, lane-events
      atomicAdd(P.phase_clocks + 0, (unsigned long long)t_busy);
      atomicAdd(P.phase_clocks + 1, (unsigned long long)(clock64() - t_all0));
      atomicAdd(P.phase_clocks + 2, (unsigned long long)n_polls);
      atomicAdd(P.phase_clocks + 3, (unsigned long long)n_prog);
      atomicAdd(P.phase_clocks + 4, (unsigned long long)n_multi);
    }
#endif
    return;
  }
  // ------------------------------------------------------------------ teams: vector side
  const int team = (warp - NL) / W;
  const int tid = threadIdx.x - 32 * NL - team * TPC;
  unsigned char* my_smem = dyn_smem + (size_t)team * L::team_bytes;
  double* team_smem = reinterpret_cast<double*>(my_smem);
  TreeTables& tables = *reinterpret_cast<TreeTables*>(my_smem + (smem_vectors<SMF>() * (size_t)TPC * EPT * sizeof(double)));
  V2Ctl& ctl = *reinterpret_cast<V2Ctl*>(dyn_smem + L::off_ctl + (size_t)team * L::ctl_bytes);
  volatile int* next_chain = reinterpret_cast<volatile int*>(dyn_smem + L::off_chain) + 2 * team;
  double* scratch = reinterpret_cast<double*>(dyn_smem + L::off_scratch + (size_t)team * L::scratch_bytes);
  MultiCtx mc;
  mc.off_model = (unsigned)L::off_model;
  mc.off_ctl = (unsigned)(L::off_ctl + (size_t)team * L::ctl_bytes);
  mc.off_ring = (unsigned)(L::off_ring + (size_t)team * L::ring_bytes);
  mc.off_mring = (unsigned)(L::off_mring + (size_t)team * L::mring_bytes);
  mc.bar_id = 1 + team;
  mc.pool = (int)blockIdx.x * C + team;
  mc.warp = (warp - NL) % W;
  unsigned cmd_seen = 0;
  // work units as in nuts_chain_kernel: one chain for set_position, ONE DRAW of one chain (draw-major) for draws
  const unsigned B = P.draws_per_unit;
  const unsigned blocks = P.mode == 0 ? 1u : ((unsigned)P.n_draws + B - 1u) / B;
  const unsigned total_units = (unsigned)P.N * blocks;
  for (;;) {
    if (tid == 0) {
      unsigned b;
      const unsigned c = unit_pop(P, total_units, b);
      next_chain[0] = (int)c;  // the leader lane reads the chain id from here (-1: no more work)
      next_chain[1] = (int)b;
    }
    bar_sync(mc.bar_id, TPC);
    const int chain = next_chain[0];
    const unsigned blk = (unsigned)next_chain[1];
    bar_sync(mc.bar_id, TPC);
    if (chain < 0) break;
    Engine<TPC, EPT, SMF, MODEL, true> E(P, chain, tid, scratch, team_smem, tables, &mc);
    if (P.mode == 0) {
      if (P.init_mask == nullptr || P.init_mask[chain] != 0) {
        const int status = cold_set_position<TPC, EPT, SMF, MODEL, true>(P, chain, tid, scratch, team_smem, &mc);
        if (tid == 0 && P.status_out) P.status_out[chain] = status;
      }
    } else {
      if (blk > 0) __threadfence();
      const uint64_t t_end = min((uint64_t)(blk + 1u) * B, (uint64_t)P.n_draws);
      for (uint64_t t = (uint64_t)blk * B; t < t_end; ++t) {
        E.load_hot();
        if (E.hs_alive) {
          E.run_draw_v2(t, cmd_seen);
        } else {
          if (t == 0) cold_fill_dead(P, chain, tid, TPC, 0);
          break;
        }
      }
      __threadfence();
      bar_sync(mc.bar_id, TPC);
      if (tid == 0) unit_push(P, (unsigned)chain, blk + 1u, blocks);
    }
    bar_sync(mc.bar_id, TPC);
#ifdef NB_PHASE_TIMING_TEAM
    if (tid == 0 && P.phase_clocks)
      for (int k = 0; k < 8; ++k) atomicAdd(P.phase_clocks + k, (unsigned long long)E.phase[k]);
#endif
  }
  if (tid == 0) {
    __threadfence_block();
    ctl.exit_flag = 1u;
  }
}

}  // namespace nb
