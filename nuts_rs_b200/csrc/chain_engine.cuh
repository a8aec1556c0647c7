// chain_engine.cuh — the whole-draw NUTS engine: one TEAM of TPC threads owns one chain for a whole
// nuts_draw() call (n_draws transitions incl. adaptation) with the phase-space point (z, v, grad_z) and the
// chain's diagonal mass matrix (sigma, mu) held in REGISTERS, EPT elements per thread.
//
// What it replaces in the reference (per chain, on one CPU core):
//   Chain::draw / set_position          src/chain.rs:137-188
//   nuts::draw, NutsTree::extend/...    src/nuts.rs:94-388           (recursion -> iterative binary counter, see extend())
//   TransformedHamiltonian::leapfrog,
//     is_turning, initialize_trajectory src/dynamics/transformed_hamiltonian.rs:524-736
//   DiagMassMatrix transform + updates  src/transform/diagonal.rs:85-265
//   AcceptanceRateCollector, DualAverage src/stepsize/dual_avg.rs:33-166
//   stepsize::Strategy (init search, jitter) src/stepsize/adapt.rs:91-267
//   RunningVariance / DiagAdaptStrategy src/transform/adapt/diagonal.rs:17-231
//   GlobalStrategy::adapt schedule      src/adapt_strategy.rs:121-222
//
// Execution model (B200-first, not a translation):
//   * chains are independent (reference src/sampler.rs:1094-1126), so there is no lock-step between chains at all:
//     a persistent grid pulls chain ids from an atomic queue; a team that finishes its chain takes the next one,
//     i.e. the "active mask" of the north star degenerates to "a finished chain frees its SM slice immediately".
//   * per leapfrog the state never leaves the register file; the only memory traffic is the 16*d-byte checkpoint
//     (z, v) of the new leaf and the checkpoint reads of the U-turn checks (served by L2 for the low tree levels).
//   * every vector (planes, checkpoints, estimators) uses the same element -> thread mapping (i = tid + j*TPC), so a
//     thread only ever re-reads global addresses it wrote itself: no barrier or fence is needed around checkpoints.
//   * all scalar tree / adaptation logic is executed redundantly by every thread of the team on bit-identical
//     reduction results (see TeamReduce), so no broadcast or extra barrier is ever needed.
#pragma once
#include "device_common.cuh"

// NB_MAXNREG (optional, per translation unit): exact register budget instead of the one ptxas derives from MIN_BLOCKS
#ifdef NB_MAXNREG
#define NB_KERNEL_BOUNDS(T, B) __maxnreg__(NB_MAXNREG)
#else
#define NB_KERNEL_BOUNDS(T, B) __launch_bounds__(T, B)
#endif
namespace nb {

constexpr int MAX_DOUBLING_DEPTH = 19;  // deepest new half has 2^19 leaves (checkpoint pool: 3 per level + 6 <= MAX_SLOTS)
constexpr int MAX_SLOTS = 64;
constexpr int ACC_RING = 32;  // leaves whose acceptance statistics are evaluated together (one exp per lane; = warp size)
constexpr int NB_END_BUFFERS = 3;  // main-tree endpoint buffers per chain: left, right + one pending (decoupled engine)

struct SettingsDev {
  uint64_t num_tune, maxdepth, mindepth, extra_doublings;
  double max_energy_error;
  int check_turning, has_target_time;
  double target_time;
  double target_accept, initial_step;
  int has_jitter, method;  // method: 0 dual average, 1 Adam, 2 fixed
  double adam_beta1, adam_beta2, adam_epsilon, adam_lr;
  double jitter, fixed_step;
  double da_k, da_t0, da_gamma, da_max_step;
  int use_grad_based, _pad;
  uint64_t early_end, final_step_size_window, mm_switch_freq, early_mm_switch_freq, mm_update_freq;
  double mm_window_growth;
};

// per-chain scalar state that survives between kernel launches
struct ChainState {
  double step_size;
  uint64_t rng_counter;
  uint64_t draw_count;
  double logp, pt_logdet;
  long long pt_transform_id;
  double mm_logdet;
  long long mm_id;
  double da_log_step, da_log_step_adapted, da_hbar, da_mu;
  uint64_t da_count;
  int tuning, has_initial_mass_matrix;
  uint64_t last_update, current_window_size;
  uint64_t fg_count, bg_count;
  int fg_set, is_good;
  double last_mean_tree_accept, last_sym_mean_tree_accept, last_max_energy_error;
  uint64_t last_n_steps;
  int alive, _pad;
  uint64_t total_leapfrogs, tree_leapfrogs;
};

struct StatsDev {  // device SoA, each [n_draws x N]
  uint64_t* depth;
  uint8_t* maxdepth_reached;
  long long* index_in_trajectory;
  double* logp;
  double* energy;
  double* energy_error;
  uint8_t* diverging;
  double* step_size;
  double* step_size_bar;
  double* mean_tree_accept;
  double* mean_tree_accept_sym;
  uint64_t* n_steps;
  double* max_energy_error;
  uint8_t* tuning;
  double* fisher_distance;
};

struct EngineParams {
  int N, d, ld, P;
  ModelDev model;
  SettingsDev s;
  uint64_t seed, chain_offset;
  double *x, *gx, *z, *gz, *v0;     // chain point planes [N][ld]
  double *stds, *inv_stds, *mean;   // DiagMassMatrix planes [N][ld]
  double* est;                      // [N][2 sets][4: draw_mean, draw_var, grad_mean, grad_var][ld]
  double* slots;                    // [teams][P][2: z, v][ld]   leaf checkpoints of the half under construction (one pool per resident team)
  double* ends;                     // [teams][NB_END_BUFFERS][3: z, v, grad_z][ld]   main-tree endpoints (v1: 0 left, 1 right)
  ChainState* cs;
  unsigned int* queue;              // [2] ready queue of this launch: pop tickets taken, push tickets taken
  unsigned int* done;               // [N] work units (blocks of draws) of this launch each chain has completed
  unsigned int* ring_seq;           // [ring_mask + 1] ready ring, slot sequence numbers (reset to 0, 1, 2, .. before every launch)
  unsigned int* ring_chain;         // [ring_mask + 1] ready ring, chain ids
  unsigned int ring_mask;           // ring size - 1; the ring holds a power of two >= N entries
  // mode 0: set_position ; mode 1: draw
  int mode, _pad;
  const double* init_position;      // [N][d] device
  const unsigned char* init_mask;   // [N] device or null: set_position only for chains with a non-zero entry (retry of bad initial points)
  int* status_out;                  // [N]
  uint64_t n_draws;
  uint32_t draws_per_unit;          // work unit of a draw launch = this many consecutive draws of one chain (>= 1)
  uint32_t _pad2;
  double* draws_out;                // [n_draws][N][d] device (may be null)
  double* grads_out;                // [n_draws][N][d] device (may be null): gradient of logp at every draw (low-rank estimator input)
  // low-rank correction of the mass matrix (SM_LOWRANK engines; null otherwise): see plane_kernels.cuh TransformDev
  const double* lr_vecs;            // [N][lr_rmax][ld]
  const double* lr_vals_sqrt;       // [N][lr_rmax]
  const double* lr_vals_sqrt_inv;   // [N][lr_rmax]
  const double* lr_mu;              // [N][ld]
  const int* lr_rank;               // [N]; -1 = no correction
  int lr_rmax, _pad3;
  StatsDev stats;
  uint64_t stats_offset;            // unused draws before this call inside the stats arrays (always 0 for now)
  unsigned long long* phase_clocks; // [16] debug phase timing (NB_PHASE_TIMING / NB_PHASE_TIMING_COLD builds), else unused
};

enum { EXT_OK = 0, EXT_TURNING = 1, EXT_DIVERGING = 2 };

// Optional phase timing (build with -DNB_PHASE_TIMING): per-phase clock64() totals of thread 0 of every team, added to
// EngineParams::phase_clocks[8] at the end of a chain.  0 init_trajectory, 1 leapfrog, 2 leaf bookkeeping + checkpoint store,
// 3 merges (turn checks), 4 doubling prologue/epilogue, 5 materialise, 6 adapt, 7 whole draw.
#ifdef NB_PHASE_TIMING_COLD
#define NB_COLD_T(k) cold_t[k] = clock64()
#define NB_COLD_T0 const long long cold_t0 = clock64()
#else
#define NB_COLD_T(k)
#define NB_COLD_T0
#endif
#ifdef NB_PHASE_TIMING
#define NB_T0(var) long long var = clock64()
#define NB_ACC(k, var) do { long long _n = clock64(); phase[k] += _n - var; var = _n; } while (0)
#else
#define NB_T0(var)
#define NB_ACC(k, var)
#endif

// The per-draw adaptation and Chain::set_position are COLD: they run in non-inlined functions on their own Engine
// instance and exchange the chain scalars by value, so that the hot tree builder (run_draw -> extend -> leapfrog, all
// force-inlined with one call site each) never has its address taken and its vectors really live in registers.
struct MultiCtx;  // decoupled engine only (several teams per CTA), see below
// The launch parameters as the cold functions see them: a private copy (COPY) or the kernel's own (see cold_adapt).
template <bool COPY>
struct ParamsCopy {
  const EngineParams P;
  __device__ __forceinline__ explicit ParamsCopy(const EngineParams& g) : P(g) {}
};
template <>
struct ParamsCopy<false> {
  const EngineParams& P;
  __device__ __forceinline__ explicit ParamsCopy(const EngineParams& g) : P(g) {}
};
template <int TPC, int EPT, int SMF, int MODEL, bool MULTI>
__device__ __noinline__ int cold_adapt(const EngineParams& P, int chain, int tid, double* scratch, double* team_smem, const MultiCtx* mc, int parity,
                                       uint64_t t, double acc_sum, double acc_sym_sum, uint64_t acc_count, double max_energy_error, bool is_good,
                                       int depth, bool reached_maxdepth, bool diverging, int draw_idx, double pt_energy,
                                       double pt_energy_error, double fisher);
template <int TPC, int EPT, int SMF, int MODEL, bool MULTI>
__device__ __noinline__ int cold_set_position(const EngineParams& P, int chain, int tid, double* scratch, double* team_smem, const MultiCtx* mc);
static __device__ __noinline__ void cold_fill_dead(const EngineParams& P, int chain, int tid, int tpc, uint64_t t0);

// ------------------------------------------------------------------------------------------------------------------------
// Decoupled engine (chain_engine_v2.cuh): protocol between the VECTOR warps of a team (leapfrogs, checkpoints, U-turn
// products; they never wait for a scalar result inside a doubling) and the LEADER warp of the CTA (lane c runs the scalar
// tree logic of team c: energies, divergence, multinomial draws, reference counts, U-turn verdicts, RNG).
// All of it lives in shared memory; flags are volatile words published after __threadfence_block().
// ------------------------------------------------------------------------------------------------------------------------
#ifndef NB_V2_SLEEP
#define NB_V2_SLEEP 64         // ns between two polls of a flag word (command wait, idle leader)
#endif
#ifndef NB_V2_SLEEP_CREDIT
#define NB_V2_SLEEP_CREDIT 32  // ns between two polls of the ring credit
#endif
constexpr int V2_K = 4;                  // leaf entries a team may be ahead of its leader lane
constexpr int V2_MAXD = 10;              // deepest doubling (2^10 leaves): maxdepth + extra_doublings <= V2_MAXD + 1
constexpr int V2_NV = 4 + 6 * V2_MAXD;   // values per leaf entry and warp: leapfrog sums + 6 per bundled merge (<= V2_MAXD merges)
constexpr int V2_NT = V2_MAXD + 2;       // per-level table size
constexpr int V2_MAXW = 16;              // warps per team
constexpr int V2_MRING = 4;              // stages of the model-parameter ring of the large-dim teams (one stage = one element per thread)
enum { V2_CMD_DOUBLING = 1, V2_CMD_TREE_DONE = 2 };

struct V2Ctl {
  // leader -> team
  volatile unsigned cmd_seq;    // bumped for every command; the value also is the epoch of the doubling it starts
  volatile unsigned cons;       // (epoch << 12) | leaves of the current doubling the leader has consumed
  int cmd_kind, cmd_dir, cmd_check, cmd_depth, cmd_prev_accepted;
  volatile signed char slot_ring[8];  // checkpoint slot of leaf i at [i & 7], published V2_K leaves ahead
  // team -> leader
  volatile unsigned prod[V2_MAXW];  // per warp: (epoch << 12) | leaves of the current doubling produced
  volatile unsigned start_seq;      // bumped when the start record below is valid (one per draw)
  volatile unsigned exit_flag;      // the team has left the kernel
  // start record (team -> leader): initialize_trajectory is done
  double E0, pt_logdet, step;
  unsigned long long rng;
  // the team's chain scalars that are not needed while the tree is built (parked here to keep them out of the registers)
  double pk_logp, pk_mm_logdet;
  long long pk_pt_tid, pk_mm_id;
  unsigned long long pk_total_lf, pk_tree_lf;
  // result record (leader -> team), valid with V2_CMD_TREE_DONE
  double acc_sum, acc_sym_sum, max_energy_error, draw_energy;
  unsigned long long acc_count, rng_out;
  int depth, draw_slot, draw_idx, reached_maxdepth, diverging;
  // leader lane's pending sub-trees, one per level (the leader is the only reader and writer)
  double A_ls[V2_NT], A_draw_energy[V2_NT];  // A_ls: linear weights (LeaderLane::lin) or log sizes
  double A_log0;                              // log weight of the single leaf pending at level 0 (exact, for the switch to the log domain)
  int A_draw_idx[V2_NT];
  signed char A_first[V2_NT], A_last[V2_NT], A_draw[V2_NT];
  // vector side: first / last checkpoint slot of the pending sub-trees, one private copy per warp
  signed char VA_first[V2_MAXW][V2_NT], VA_last[V2_MAXW][V2_NT];
};

// leaf entry ring of one team: double ring[V2_K][W][V2_NV], per-warp partial sums of one leaf:
// [0..3] logp, v.v, sP, sQ of the leapfrog; then 6 products per bundled merge (inner levels 1.., then the top-level merge)

// Byte OFFSETS into the CTA's dynamic shared memory, not pointers: the struct travels through local memory (the cold functions
// take its address), and a pointer loaded from memory is a generic pointer to the compiler (generic LD / ST); rebuilding the
// pointers from the `extern __shared__` base keeps every access an LDS / STS.
struct MultiCtx {
  unsigned off_model;  // [2][TPC*EPT] model mu | prec, one copy per CTA
  unsigned off_ctl;    // V2Ctl of the team
  unsigned off_ring;   // double [V2_K][W][V2_NV]
  unsigned off_mring;  // double [V2_MRING][2][TPC]: cp.async ring of the model parameters (SM_MODEL_GLOBAL teams)
  int bar_id, warp;    // the team's named barrier, this warp's index inside the team
  int pool;            // index of the team's checkpoint pool (blockIdx.x * teams per CTA + team)
};

// volatile load of a flag word in SHARED memory (a plain volatile access through a generic pointer would be a generic LD)
__device__ __forceinline__ unsigned v2_ld(const volatile unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(const_cast<const unsigned*>(p))) : "memory");
  return v;
}

// Tree bookkeeping tables of one chain, in shared memory (local-memory tables cost an L1 miss per access once the stacks of
// all resident threads exceed L1).  Every thread computes the same values; thread 0 of the team stores them.  An entry
// written after leaf i is first read during leaf i+1's merges, i.e. after the barrier / __syncwarp of that leaf's leapfrog
// reduction, and entries of different levels never alias, so no extra synchronisation is needed for CTA teams.
struct alignas(16) TreeTables {  // (16: the teams' shared-memory slices follow each other and hold double2-accessed vectors)
  // pending sub-trees of the half under construction, one per level
  double A_ls[MAX_DOUBLING_DEPTH], A_draw_energy[MAX_DOUBLING_DEPTH];
  int A_draw_idx[MAX_DOUBLING_DEPTH];
  signed char A_first[MAX_DOUBLING_DEPTH], A_last[MAX_DOUBLING_DEPTH], A_draw[MAX_DOUBLING_DEPTH];
};

// Evaluate the acceptance statistics of the n pending leaves and add them in leaf order.  The batch lives in REGISTERS: lane k of
// every warp of the team holds the energy difference of pending leaf k (`mine`), so each warp evaluates the whole batch on its own -
// one exp per lane - and walks through it with shuffles: no shared memory, no traffic between the warps of a team.
struct AccSums {
  double sum, sym, max_err;
};
static __device__ __noinline__ AccSums accept_batch(double mine, int n, double acc_sum, double acc_sym_sum, double max_energy_error) {
  const int lane = threadIdx.x & 31;
  double e = 0.0, s = 0.0;
  if (lane < n) {
    const double ed = accept_exp(mine);
    e = mine < 0. ? ed : 1.0;  // == exp(min(diff, 0)) bit for bit (diff is finite here), one exp instead of two
    s = 2. * e / (1. + ed);
  }
  AccSums r{acc_sum, acc_sym_sum, max_energy_error};
  for (int k = 0; k < n; ++k) {
    r.sum += __shfl_sync(0xffffffffu, e, k);
    r.sym += __shfl_sync(0xffffffffu, s, k);
    const double diff = __shfl_sync(0xffffffffu, mine, k);
    if (fabs(diff) > fabs(r.max_err)) r.max_err = diff;
  }
  return r;
}

// What lives in the team's dynamic shared memory (SMF bit flags); everything else is registers (or L1-cached global for the
// model parameters).  Layout: [sigma | mean] (SM_MASS) [model mu | model prec] (SM_MODEL) [grad_z] (SM_GRAD) TreeTables.
// SM_EXACT: every sampler row is padded with zeros to TPC*EPT elements (EngineParams::ld >= TPC*EPT), so the hot loops run
// without bounds checks: the padding lanes compute on zeros and stay zero.
// SM_NOGRAD (elementwise target only): the tree builder keeps no gradient vector at all; a leapfrog recomputes grad_z of its start
// point from z (5 flops per element) - 2*EPT registers less per thread, i.e. more resident teams per SM.
// SM_MODEL_GLOBAL (decoupled engine): no CTA-wide shared-memory copy of the model parameters (they do not fit for dim ~ 10^4);
// they are read through the read-only path from global memory, where capi.cu pads them with zeros to a multiple of 1024.
// SM_STAGE: four more vectors of shared memory into which cp.async (LDGSTS, 16 bytes per pair) stages the OLDEST operand of the
// upcoming U-turn checks - (z, v) of the first leaf of the pending sub-tree (buffer X), of the far end of the main tree (buffer Y)
// - while the leapfrog of the leaf that triggers the check is still running; see Engine::stage_pair().
// SM_CL2 / SM_CL4: the team spans the 2 / 4 CTAs of a thread-block cluster (TPC = threads of the whole team): every CTA holds its
// threads' elements of the shared-memory vectors, reductions go through distributed shared memory (TeamReduce, CL > 1).
// SM_ALIGN: the warp teams of a CTA start their work units together (a CTA barrier per unit): warps that run the same code at
// the same time share the 32 KB instruction cache of the SM - the engine's per-draw instruction working set is 80 - 140 KB.
enum { SM_MASS = 1, SM_MODEL = 2, SM_GRAD = 4, SM_EXACT = 8, SM_NOGRAD = 16, SM_MODEL_GLOBAL = 32, SM_STAGE = 64, SM_CL2 = 128, SM_CL4 = 256, SM_ALIGN = 512, SM_LOWRANK = 1024, SM_STAGE1 = 2048 };
// SM_STAGE1 (with SM_STAGE): ONE staging buffer pair instead of two (16 KB less shared memory per 1000-dim team: 4 CTAs per SM stay
// resident next to the mass matrix and the model parameters); the end-of-tree buffer Y does not exist.
// SM_LOWRANK: the chain's transformation may carry the low-rank correction of LowRankMassMatrix (reference src/transform/low_rank.rs):
// x = sigma * ((I + U (sqrt(lambda) - 1) U^T) z + mu_lr) + mean.  Every leapfrog then runs the general (two-pass) path with two more
// team-wide reductions of r values (U^T z and U^T (sigma * grad_x)); the elementwise shortcuts of the diagonal Gaussian are off.
template <int SMF>
__host__ __device__ constexpr int cluster_size() {
  return (SMF & SM_CL4) ? 4 : ((SMF & SM_CL2) ? 2 : 1);
}
template <int SMF>
__host__ __device__ constexpr int smem_vectors() {
  return ((SMF & SM_MASS) ? 2 : 0) + ((SMF & SM_MODEL) ? 2 : 0) + ((SMF & SM_GRAD) ? 1 : 0) + ((SMF & SM_STAGE) ? ((SMF & 2048) ? 2 : 4) : 0);
}
template <int TPC, int EPT, int SMF>
__host__ __device__ constexpr size_t team_smem_bytes() {
  return smem_vectors<SMF>() * (size_t)(TPC / cluster_size<SMF>()) * EPT * sizeof(double) + sizeof(TreeTables);  // per CTA
}

template <int TPC, int EPT, int SMF, int MODEL, bool MULTI = false>
struct Engine {
  const EngineParams& P;
  const int chain;
  const int tid;  // thread index inside the team
  static constexpr int CL = cluster_size<SMF>();  // CTAs the team spans (thread-block cluster)
  static constexpr int LT = TPC / CL;             // the team's threads in THIS CTA
  static_assert(CL == 1 || !MULTI, "cluster teams use the register-resident engine");
  const int ltid;                                 // thread index inside the CTA's share of the team
  // several CTA-sized teams in one CTA (SM_ALIGN with TPC > 32): the team synchronises on its own named barrier like the
  // teams of the decoupled engine
  static constexpr bool SUBT = (SMF & SM_ALIGN) != 0 && TPC > 32 && !MULTI && CL == 1;
  TeamReduce<LT, MULTI || SUBT, CL> red;
  const MultiCtx* const mc;  // MULTI only
  const int d, ld;
  const size_t row;  // chain * ld
  double* const slots_base;  // this chain's checkpoint pool / endpoint buffers (thread-offset included)
  double* const ends_base;

  // ---- register-resident vectors ----
  static constexpr bool MMS = (SMF & SM_MASS) != 0, MODS = (SMF & SM_MODEL) != 0, GS = (SMF & SM_GRAD) != 0, EXACT = (SMF & SM_EXACT) != 0,
                        LR = (SMF & SM_LOWRANK) != 0,
                        ELEMWISE = MODEL == LOGP_GAUSS_DIAG && !LR,  // grad_z of a leaf is an elementwise function of its z
                        NOG = (SMF & SM_NOGRAD) != 0 && ELEMWISE,
                        MSH = MODS || (MULTI && (SMF & SM_MODEL_GLOBAL) == 0);  // model parameters are read from shared memory
  static constexpr bool STAGE = (SMF & SM_STAGE) != 0 && (EPT % 2) == 0 && !MULTI;
  // (capi.cu pads the global model parameter arrays with zeros up to the largest tile)
  static_assert(!EXACT || MMS, "SM_EXACT needs a zero-padded copy of the mass matrix");
  static_assert(!LR || (!MULTI && !GS && !EXACT && CL == 1), "low-rank engines: plain register-resident tilings");
  __device__ __forceinline__ bool inb(int i) const { return EXACT || i < d; }  // element i exists (or is zero padding that may be touched)
  // Element -> thread mapping of every vector.  Even EPT: a thread owns PAIRS of adjacent elements (2*tid, 2*tid + 1, then the
  // same 2*TPC further on), so its accesses to planes, checkpoints and shared memory are 16-byte (double2: LDG.E.128 / STG.E.128 /
  // LDS.128, cp.async 16 B) and one Box-Muller pair of the momentum belongs to one thread.  Odd EPT (EPT = 1): element tid + j*TPC.
  static constexpr bool PAIR = (EPT % 2) == 0;
  static constexpr int NP = EPT / 2;
  __device__ __forceinline__ int eidx(int j) const { return PAIR ? (j >> 1) * (2 * TPC) + 2 * tid + (j & 1) : tid + j * TPC; }
  // the same element in the CTA's shared-memory vectors (equal to eidx unless the team spans a cluster)
  __device__ __forceinline__ int sidx(int j) const { return CL == 1 ? eidx(j) : (PAIR ? (j >> 1) * (2 * LT) + 2 * ltid + (j & 1) : ltid + j * LT); }
  double z[EPT], v[EPT];  // current phase-space point: whitened position, velocity
  double g_reg[GS ? 1 : EPT];  // whitened gradient: registers, or shared memory (GS) to fit more chains per SM
  // this chain's DiagMassMatrix (stds, mean): registers, or shared memory when MMS
  double sig[MMS ? 1 : EPT], mu[MMS ? 1 : EPT];
  double *sm_sig, *sm_mu, *sm_mmu, *sm_mprec, *sm_g, *sm_stage;
  TreeTables& T;

  // Chain scalars used on the hot path are plain members (registers).  `cs` is only touched by the cold functions, which
  // exchange it with global memory themselves: an aggregate copy of a struct that also holds hot fields would pin the
  // whole engine object in local memory (measured: 2 local accesses per counter update per leapfrog).
  double hs_step, hs_logp, hs_pt_logdet, hs_mm_logdet;
  long long hs_pt_tid, hs_mm_id;
  uint64_t hs_rng, hs_total_lf, hs_tree_lf, hs_draw_count;
  double hs_da_lsa;  // DualAverage::log_step_adapted (read-only on the hot path: the step size after tuning)
  int hs_alive;
  ChainState cs;  // cold path only
  uint64_t stream;

  // ---- per-draw collector (AcceptanceRateCollector, dual_avg.rs:112-166) ----
  double E0, acc_sum, acc_sym_sum, max_energy_error;
  uint64_t acc_count;

  // ---- main tree ----
  // Tree weights: linear domain (`lin`, see device_common.cuh): ls_main = sum of exp(-energy error) over the main tree's leaves,
  // TreeTables::A_ls[l >= 1] the same for the pending sub-trees, A_ls[0] the LOG weight of a single pending leaf (two leaves are
  // exponentiated together when they merge).  Reference log domain (!lin): log_size everywhere.
  double ls_main;
  bool lin;
  int depth;
  int idx_left, idx_right;                 // index_in_trajectory of the two ends
  bool init_left, init_right;              // the end still is the initial point
  int es_left, es_right;                   // checkpoint slot that holds (z, v) of that end (the last leaf of the half that was merged in)
  bool holds_left, holds_right;            // the registers currently hold that end
  int draw_slot;  // -1: the draw is the initial point
  double draw_energy;
  int draw_idx;

  // checkpoint pool: free-slot mask and 2-bit reference counts (first / last / draw roles => at most 3), all in registers
  uint64_t free_mask, rc_lo, rc_hi;
#ifdef NB_PHASE_TIMING
  long long phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

  // team_smem: this team's slice of dynamic shared memory (team_smem_bytes()); tables: where the TreeTables live
  __device__ __forceinline__ Engine(const EngineParams& p, int chain_, int tid_, double* scratch, double* team_smem, TreeTables& tables,
                                    const MultiCtx* mc_ = nullptr)
      : P(p), chain(chain_), tid(tid_), ltid(CL > 1 ? (int)threadIdx.x : tid_), red(scratch), mc(mc_), d(p.d), ld(p.ld), row((size_t)chain_ * p.ld),
        slots_base(p.slots + (size_t)pool_index(mc_) * p.P * 2 * p.ld), ends_base(p.ends + (size_t)pool_index(mc_) * NB_END_BUFFERS * 3 * p.ld),
        sm_sig(team_smem),
        sm_mu(team_smem + (MMS ? LT * EPT : 0)), sm_mmu(team_smem + (MMS ? 2 : 0) * LT * EPT),
        sm_mprec(team_smem + ((MMS ? 2 : 0) + (MODS ? 1 : 0)) * LT * EPT),
        sm_g(team_smem + ((MMS ? 2 : 0) + (MODS ? 2 : 0)) * LT * EPT),
        sm_stage(team_smem + ((MMS ? 2 : 0) + (MODS ? 2 : 0) + (GS ? 1 : 0)) * LT * EPT), T(tables) {
    stream = p.chain_offset + (uint64_t)chain_ + 1;  // reference src/sampler.rs:1106 set_stream(chain_id + 1)
    if (LR && p.lr_rank) {
      const int r = p.lr_rank[chain_];
      lr_r = r < p.lr_rmax ? r : p.lr_rmax;
    }
    if (SUBT) {
      red.bar_id = 1 + (int)threadIdx.x / TPC;
      red.warp = ((int)threadIdx.x % TPC) >> 5;
    }
    if (MULTI) {  // several teams per CTA: named barrier, per-team reduction scratch, CTA-wide copy of the model parameters
      red.bar_id = mc->bar_id;
      red.warp = mc->warp;
      extern __shared__ __align__(16) unsigned char nb_dyn_smem[];
      sm_mmu = reinterpret_cast<double*>(nb_dyn_smem + mc->off_model);
      sm_mprec = sm_mmu + TPC * EPT;
      // keep what the hot loops use in registers (*mc lives in local memory)
      v2_ctl = reinterpret_cast<V2Ctl*>(nb_dyn_smem + mc->off_ctl);
      v2_ring = reinterpret_cast<double*>(nb_dyn_smem + mc->off_ring);
      sm_mring = reinterpret_cast<double*>(nb_dyn_smem + mc->off_mring);
      v2_warp = mc->warp;
    }
  }
  V2Ctl* v2_ctl;    // MULTI only
  double* v2_ring;
  double* sm_mring = nullptr;
  int v2_warp;
  // Large-dim teams (SM_MODEL_GLOBAL): the model's mu / precision (2 x 8*d bytes, the same for every leapfrog) do not fit on chip
  // next to z, v (registers) and sigma, mean (shared memory) and come from L2 on every leapfrog.  Plain loads put that latency
  // on the critical path of every element (the registers are full: few loads can be in flight); cp.async streams them through
  // a small shared-memory ring V2_MRING - 1 elements ahead instead - every thread copies and reads only its own elements, so
  // completion is its own cp.async.wait_group and no barrier is involved.
  static constexpr bool MSTREAM = MULTI && (SMF & SM_MODEL_GLOBAL) != 0;
  __device__ __forceinline__ void mstream_issue(int j) {
    if (j < EPT) {
      const int i = eidx(j);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sm_mring + (size_t)((j % V2_MRING) * 2) * TPC + tid);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(P.model.mu + i) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * TPC), "l"(P.model.prec + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // (an empty group past the last element keeps the count uniform)
  }
  __device__ __forceinline__ void mstream_wait(int j, double& mmu, double& mprec) {
    asm volatile("cp.async.wait_group %0;" ::"n"(V2_MRING - 1) : "memory");
    const double* st = sm_mring + (size_t)((j % V2_MRING) * 2) * TPC + tid;
    mmu = st[0];
    mprec = st[TPC];
  }
  // Checkpoints and end buffers only live inside one draw, which runs on one team from start to end (draw_finish materialises the
  // chain point into the x / z planes): the pools belong to the resident TEAM, not to the chain (EngineParams::slots / ends are
  // sized grid * teams per CTA).
  static __device__ __forceinline__ int pool_index(const MultiCtx* m) {
    if (MULTI) return m->pool;
    if (CL > 1) return (int)blockIdx.x / CL;  // one team per cluster
    return (int)blockIdx.x * ((int)blockDim.x / TPC) + (int)threadIdx.x / TPC;
  }
  __device__ __forceinline__ double sg(int j) const { return MMS ? sm_sig[sidx(j)] : sig[MMS ? 0 : j]; }
  __device__ __forceinline__ double mn(int j) const { return MMS ? sm_mu[sidx(j)] : mu[MMS ? 0 : j]; }
  // model parameters of element i = tid + j*TPC
  // (MULTI: one copy per CTA shared by all its teams, filled by the kernel prologue)
  __device__ __forceinline__ double model_mu(int j, int i) const { return MSH ? sm_mmu[sidx(j)] : __ldg(P.model.mu + i); }
  __device__ __forceinline__ double model_prec(int j, int i) const { return MSH ? sm_mprec[sidx(j)] : __ldg(P.model.prec + i); }
  // element j of the whitened gradient (each thread only touches its own entries: no synchronisation)
  __device__ __forceinline__ double& G(int j) { return GS ? sm_g[sidx(j)] : g_reg[GS ? 0 : j]; }
  // The four per-element constants of the diagonal Gaussian leapfrog (sigma, mean of the mass matrix; mu, precision of the
  // model) for elements j, j + 1 of a PAIR (j even): one 16-byte access per vector instead of two 8-byte ones.
  struct PairConsts {
    double sg[2], mn[2], mm[2], pr[2];
  };
  __device__ __forceinline__ void pair_consts(int j, PairConsts& c) const {
    const int i0 = eidx(j), s0 = sidx(j);
    if (MMS) {
      const double2 a = *reinterpret_cast<const double2*>(sm_sig + s0), b = *reinterpret_cast<const double2*>(sm_mu + s0);
      c.sg[0] = a.x, c.sg[1] = a.y, c.mn[0] = b.x, c.mn[1] = b.y;
    } else {
      c.sg[0] = sig[MMS ? 0 : j], c.sg[1] = sig[MMS ? 0 : j + 1], c.mn[0] = mu[MMS ? 0 : j], c.mn[1] = mu[MMS ? 0 : j + 1];
    }
    double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
    if (inb(i0)) {  // the partner of an odd last element reads the zero padding of the parameter arrays
      if (MSH) a = *reinterpret_cast<const double2*>(sm_mmu + s0), b = *reinterpret_cast<const double2*>(sm_mprec + s0);
      else a = __ldg(reinterpret_cast<const double2*>(P.model.mu + i0)), b = __ldg(reinterpret_cast<const double2*>(P.model.prec + i0));
    }
    c.mm[0] = a.x, c.mm[1] = a.y, c.pr[0] = b.x, c.pr[1] = b.y;
  }
  __device__ __forceinline__ void load_model_params() {
    if (MODS && !MULTI) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        sm_mmu[sidx(j)] = i < d ? P.model.mu[i] : 0.0;
        sm_mprec[sidx(j)] = i < d ? P.model.prec[i] : 0.0;
      }
    }
  }
  // make table stores visible to the other lanes of a warp team (CTA teams use private tables)
  __device__ __forceinline__ void tsync() const {
    if (TPC == 32) __syncwarp();
  }
  // checkpoint traffic bypasses L1 (ld.cg / st.cg): it is streamed once, L1 is kept for the model parameters and tables.
  // PAIR mapping: 16-byte accesses (rows start on 128-byte lines and the first element of a pair has an even index; the
  // partner of an odd last element is zero padding of the row, which is never overwritten).
  __device__ __forceinline__ void load_cg(const double* __restrict__ src, double (&a)[EPT]) const {
    if (PAIR) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = eidx(2 * p);
        double2 t = make_double2(0.0, 0.0);
        if (inb(i0)) t = __ldcg(reinterpret_cast<const double2*>(src + i0));
        a[2 * p] = t.x;
        a[2 * p + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        a[j] = inb(i) ? __ldcg(src + i) : 0.0;
      }
    }
  }
  __device__ __forceinline__ void store_cg(double* __restrict__ dst, const double (&a)[EPT]) const {
    if (PAIR) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = eidx(2 * p);
        if (EXACT || i0 + 1 < d) __stcg(reinterpret_cast<double2*>(dst + i0), make_double2(a[2 * p], a[2 * p + 1]));
        else if (i0 < d) __stcg(dst + i0, a[2 * p]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (inb(i)) __stcg(dst + i, a[j]);
      }
    }
  }

  // ------------------------------------------------------------------ vector helpers (rows of the [N][ld] planes)
  __device__ __forceinline__ void load(const double* __restrict__ src, double (&a)[EPT]) const {
    if (PAIR) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = eidx(2 * p);
        double2 t = make_double2(0.0, 0.0);
        if (i0 < d) t = *reinterpret_cast<const double2*>(src + i0);
        a[2 * p] = t.x;
        a[2 * p + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        a[j] = i < d ? src[i] : 0.0;
      }
    }
  }
  __device__ __forceinline__ void store(double* __restrict__ dst, const double (&a)[EPT]) const {
    if (PAIR) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = eidx(2 * p);
        if (i0 + 1 < d) *reinterpret_cast<double2*>(dst + i0) = make_double2(a[2 * p], a[2 * p + 1]);
        else if (i0 < d) dst[i0] = a[2 * p];
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) dst[i] = a[j];
      }
    }
  }
  // a row of the dense [n_draws][N][d] output (stride d: 16-byte aligned only for even d)
  __device__ __forceinline__ void store_dense(double* __restrict__ dst, const double (&a)[EPT]) const {
    // the draw output is written once and not read again by the engine: streaming (evict-first) stores
    if (PAIR && (d & 1) == 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = eidx(2 * p);
        if (i0 < d) __stcs(reinterpret_cast<double2*>(dst + i0), make_double2(a[2 * p], a[2 * p + 1]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) __stcs(dst + i, a[j]);
      }
    }
  }
  __device__ __forceinline__ void load_g(const double* __restrict__ src, bool cg) {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = eidx(j);
      G(j) = i < d ? (cg ? __ldcg(src + i) : src[i]) : 0.0;
    }
  }
  __device__ __forceinline__ void store_g(double* __restrict__ dst, bool cg) {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = eidx(j);
      if (i < d) {
        if (cg) __stcg(dst + i, G(j));
        else dst[i] = G(j);
      }
    }
  }
  __device__ __forceinline__ double* slot_ptr(int s, int which) const { return slots_base + (size_t)((s * 2 + which) * ld); }
  __device__ __forceinline__ double* end_ptr(int dir, int which) const { return ends_base + (size_t)((dir * 3 + which) * ld); }
  __device__ __forceinline__ double* est_ptr(int set, int which) const {
    return P.est + (((size_t)chain * 2 + set) * 4 + which) * P.ld;
  }

  // ------------------------------------------------------------------ chain scalars <-> global memory
  __device__ __forceinline__ void load_hot() {
    const ChainState* g = P.cs + chain;
    hs_step = __ldcg(&g->step_size);
    hs_logp = __ldcg(&g->logp);
    hs_pt_logdet = __ldcg(&g->pt_logdet);
    hs_mm_logdet = __ldcg(&g->mm_logdet);
    hs_pt_tid = __ldcg(&g->pt_transform_id);
    hs_mm_id = __ldcg(&g->mm_id);
    hs_rng = __ldcg((const unsigned long long*)&g->rng_counter);
    hs_total_lf = __ldcg((const unsigned long long*)&g->total_leapfrogs);
    hs_tree_lf = __ldcg((const unsigned long long*)&g->tree_leapfrogs);
    hs_alive = __ldcg(&g->alive);
    hs_draw_count = __ldcg((const unsigned long long*)&g->draw_count);
    hs_da_lsa = __ldcg(&g->da_log_step_adapted);
  }
  __device__ __forceinline__ void team_sync() const {
    if (TPC > 32) red.barrier();
    else __syncwarp();
  }
  // publish the hot scalars (all threads hold identical values; thread 0 writes) and make them visible to the team
  __device__ __forceinline__ void store_hot() {
    if (tid == 0) {
      ChainState* g = P.cs + chain;
      g->step_size = hs_step;
      g->logp = hs_logp;
      g->pt_logdet = hs_pt_logdet;
      g->mm_logdet = hs_mm_logdet;
      g->pt_transform_id = hs_pt_tid;
      g->mm_id = hs_mm_id;
      g->rng_counter = hs_rng;
      g->total_leapfrogs = hs_total_lf;
      g->tree_leapfrogs = hs_tree_lf;
    }
    team_sync();
  }
  // cold functions: whole ChainState from / to global
  __device__ __forceinline__ void cold_load() {
    cs = P.cs[chain];
    hs_step = cs.step_size; hs_logp = cs.logp; hs_pt_logdet = cs.pt_logdet; hs_mm_logdet = cs.mm_logdet;
    hs_pt_tid = cs.pt_transform_id; hs_mm_id = cs.mm_id; hs_rng = cs.rng_counter; hs_total_lf = cs.total_leapfrogs;
    hs_tree_lf = cs.tree_leapfrogs; hs_alive = cs.alive;
  }
  __device__ __forceinline__ void cold_store() {
    cs.step_size = hs_step; cs.logp = hs_logp; cs.pt_logdet = hs_pt_logdet; cs.mm_logdet = hs_mm_logdet;
    cs.pt_transform_id = hs_pt_tid; cs.mm_id = hs_mm_id; cs.rng_counter = hs_rng; cs.total_leapfrogs = hs_total_lf;
    cs.tree_leapfrogs = hs_tree_lf; cs.alive = hs_alive;
    team_sync();  // every thread has finished reading the old record
    if (tid == 0) P.cs[chain] = cs;
    team_sync();
  }

  // ------------------------------------------------------------------ random stream
  // Every thread of the team consumes the same scalar random numbers, so a warp evaluates 32 consecutive Philox blocks at once
  // (lane k: event counter pc_base + k) and serves them by shuffle: one block evaluation per ~32 events instead of one per event.
  // A block is a pure function of (seed, stream, counter): values equal stream_bool / stream_f64 of rng_spec bit for bit.
#ifdef NB_PHASE_TIMING_COLD
  long long cold_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  uint64_t pc_base = 1ull << 63;  // no counter ever gets near: the first use refills
  uint32_t pc_r0 = 0, pc_r1 = 0;
  __device__ __forceinline__ void rng_words(uint32_t& r0, uint32_t& r1) {
    uint64_t off = hs_rng - pc_base;
    if (off >= 32ull) {  // warp-uniform
      pc_base = hs_rng;
      const uint2 b = philox01(P.seed, stream, hs_rng + (uint64_t)(threadIdx.x & 31));
      pc_r0 = b.x;
      pc_r1 = b.y;
      off = 0;
    }
    r0 = __shfl_sync(0xffffffffu, pc_r0, (int)off);
    r1 = __shfl_sync(0xffffffffu, pc_r1, (int)off);
    hs_rng += 1;
  }
  __device__ __forceinline__ bool rng_bool() {
    uint32_t r0, r1;
    rng_words(r0, r1);
    return (r0 & 1u) != 0;
  }
  __device__ __forceinline__ double rng_f64() {
    uint32_t r0, r1;
    rng_words(r0, r1);
    return u53((uint64_t)r0 | ((uint64_t)r1 << 32));
  }
  // array_gaussian(rng, v, ones): v[i] = 1.0 * normal   (cpu_math.rs:561-577)
  __device__ __forceinline__ void sample_velocity() {
    // Elements 2p and 2p+1 are one Box-Muller pair.  PAIR mapping: both belong to the same thread; values are exactly those of
    // stream_normal().  The loop is ROLLED (the Box-Muller body is ~250 instructions; unrolled over EPT it would sweep 30 KB of
    // code through the 32 KB instruction cache once per draw) with NB_BM_STEP (2) independent pairs per iteration, so that their
    // dependency chains (log, sin / cos polynomials, square root) interleave.  The normals go to the chain's v0 plane - where
    // initialize_trajectory stores the fresh velocity anyway - and every thread reads back the elements it wrote itself.
    double* v0 = P.v0 + row;
    if (PAIR) {
#ifndef NB_BM_STEP
#define NB_BM_STEP 2
#endif
      constexpr int STEP = (NP % NB_BM_STEP == 0) ? NB_BM_STEP : ((NP % 2 == 0) ? 2 : 1);
#pragma unroll 1
      for (int p = 0; p < NP; p += STEP) {
        double n[STEP][2];
        int i0[STEP];
#pragma unroll
        for (int u = 0; u < STEP; ++u) {
          i0[u] = eidx(2 * (p + u));
          // a pair beyond dim (or the odd last element's partner) is computed and dropped: no branch inside the chains
          stream_normal_pair(P.seed, stream, hs_rng, (uint32_t)(i0[u] < d ? i0[u] : 0), n[u][0], n[u][1]);
        }
#pragma unroll
        for (int u = 0; u < STEP; ++u) {
          if (i0[u] < d) v0[i0[u]] = 1.0 * n[u][0];
          if (i0[u] + 1 < d) v0[i0[u] + 1] = 1.0 * n[u][1];
        }
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) v0[i] = 1.0 * stream_normal(P.seed, stream, hs_rng, (uint32_t)i);
      }
    }
    load(v0, v);
    hs_rng += (uint64_t)((d + 1) / 2);
  }

  // ------------------------------------------------------------------ log density
  // (logp, grad) at x; `extra` rides along in the final reduction (used for the kinetic energy).
  // Elementwise formulas match oracle/nuts_oracle.hpp (GaussIso, GaussDiag, GaussRank1, Funnel).
  __device__ __forceinline__ void model_phase_a(const double (&x)[EPT], double& a0, double& a1) {
    const ModelDev& m = P.model;
    a0 = 0.0;
    a1 = 0.0;
    if (MODEL == LOGP_GAUSS_RANK1) {
      double s[1] = {0.0};
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) s[0] += x[j] - model_mu(j, i);
      }
      red.allreduce(s);
      a0 = m.rank1_coeff * s[0];  // rank1_term
    } else if (MODEL == LOGP_FUNNEL) {
      double s[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) {
          if (i == 0) s[1] = x[j];
          else s[0] = fma(x[j], x[j], s[0]);
        }
      }
      red.allreduce(s);
      a0 = s[1];  // v
      a1 = s[0];  // S
    } else if (MODEL == LOGP_USER && NutsUserLogp::NUM_SUMS > 0) {  // the team-wide sums the user's element pass asks for
      double s[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) NutsUserLogp::sums(i, d, x[j], m.user, s);
      }
      red.allreduce(s);
      a0 = s[0];
      a1 = s[1];
    }
  }
  // LOGP_USER error channel (LogpError, reference src/math/math.rs:9-13): what this thread's elements reported during the last
  // model_phase_b, and the draw-level verdict after the team reduction
  double ue_rec = 0.0, ue_fatal = 0.0;
  bool user_fatal = false;
  // elementwise gradient; returns this thread's partial of logp (sum-type models)
  __device__ __forceinline__ double model_phase_b(const double (&x)[EPT], double (&gx)[EPT], double a0, double a1, double& ev_out) {
    const ModelDev& m = P.model;
    double lp = 0.0;
    ev_out = 0.0;
    if (false /* ISO is served by the DIAG code path with prec = 1 */) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        gx[j] = 0.0;
        if (i < d) {
          double diff = x[j] - model_mu(j, i);
          lp -= diff * diff / 2.;
          gx[j] = -diff;
        }
      }
    } else if (MODEL == LOGP_GAUSS_DIAG) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        gx[j] = 0.0;
        if (i < d) {
          double diff = x[j] - model_mu(j, i);
          double pd = diff * model_prec(j, i);
          lp -= diff * pd / 2.;
          gx[j] = -pd;
        }
      }
    } else if (MODEL == LOGP_GAUSS_RANK1) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        gx[j] = 0.0;
        if (i < d) {
          double diff = x[j] - model_mu(j, i);
          double ptd = diff - a0;
          gx[j] = -ptd;
          lp -= 0.5 * diff * ptd;
        }
      }
    } else if (MODEL == LOGP_USER) {
      const double sums[2] = {a0, a1};
      ue_rec = 0.0;
      ue_fatal = 0.0;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        gx[j] = 0.0;
        if (i < d) {
          int status = NUTS_USER_OK;
          lp += NutsUserLogp::element(i, d, x[j], sums, m.user, gx[j], status);
          if (status == NUTS_USER_RECOVERABLE) ue_rec = 1.0;
          if (status == NUTS_USER_FATAL) ue_fatal = 1.0;
        }
      }
    } else {  // funnel
      double vv = a0;
      double ev = exp(-vv);
      ev_out = ev;
      double nm1 = (double)(d - 1);
      double half_ev_S = 0.5 * ev * a1;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        gx[j] = 0.0;
        if (i < d) {
          if (i == 0) gx[j] = -vv * m.funnel_inv_var - 0.5 * nm1 + half_ev_S;
          else gx[j] = -x[j] * ev;
        }
      }
    }
    return lp;
  }
  __device__ __forceinline__ double model_finish(double lp_sum, double a0, double a1, double ev) const {
    const ModelDev& m = P.model;
    if (MODEL == LOGP_FUNNEL) {
      double nm1 = (double)(d - 1);
      double half_ev_S = 0.5 * ev * a1;
      return -0.5 * a0 * a0 * m.funnel_inv_var - 0.5 * nm1 * a0 - half_ev_S;
    }
    if (MODEL == LOGP_USER) {
      const double sums[2] = {a0, a1};
      return lp_sum + NutsUserLogp::finish(d, sums, m.user);
    }
    return lp_sum;
  }
  // LOGP_USER: fold the team's error flags (reduced next to logp) into the result: any error makes logp NaN - the leapfrog is then
  // a divergence (non-finite energy, transformed_hamiltonian.rs:562-565, 590-597) - and a fatal one also ends the chain
  __device__ __forceinline__ double user_verdict(double logp, double n_rec, double n_fatal) {
    if (n_fatal > 0.0) user_fatal = true;
    return (n_rec > 0.0 || n_fatal > 0.0) ? __longlong_as_double(0x7ff8000000000000ll) : logp;
  }

  // One velocity-Verlet step of the diagonal Gaussian WITHOUT the team reduction: (z, v, g) advance in place, part = this thread's
  // partial sums of (logp', v'.v', sP, sQ).  Elementwise target: the whole step is ONE pass, nothing but accumulators outlives
  // an element.
  __device__ __forceinline__ void leapfrog_partials(double eps, double (&part)[4]) {
    const double eps_half = eps / 2.;
    part[0] = part[1] = part[2] = part[3] = 0.0;
    PairConsts pc;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int i = eidx(j);
      if (PAIR && (j & 1) == 0) pair_consts(j, pc);
      const double zp = z[j], vp = v[j];
      const double vh = fma(eps_half, G(j), vp);   // first_velocity_halfstep :178-184  axpy_out(grad, v, eps/2)
      const double zn = fma(eps, vh, zp);          // position_step :220-225            axpy_out(v', z, eps)
      const double sgm = PAIR ? pc.sg[j & 1] : sg(j);
      const double t = zn * sgm;                   // compute_untransformed_position    diagonal.rs:253-255
      const double xn = fma(1.0, PAIR ? pc.mn[j & 1] : mn(j), t);
      double gxn = 0.0;
      if (inb(i)) {
        const double diff = xn - (PAIR ? pc.mm[j & 1] : model_mu(j, i));
        const double pd = diff * (PAIR ? pc.pr[j & 1] : model_prec(j, i));
        part[0] -= diff * pd / 2.;
        gxn = -pd;
      }
      const double gn = gxn * sgm;                 // compute_transformed_gradient      diagonal.rs:258-265
      const double vn = fma(eps_half, gn, vh);     // second_velocity_halfstep :245-247 axpy(grad', v, eps/2)
      part[1] = fma(vn, vn, part[1]);              // update_kinetic_energy :260-262
      if (inb(i)) {
        const double delta = (zn + 0.0) - zp;
        part[2] = fma(delta, vp, part[2]);
        part[3] = fma(delta, vn, part[3]);
      }
      z[j] = zn;
      v[j] = vn;
      G(j) = gn;
    }
  }

  // ------------------------------------------------------------------ transformation: position / gradient maps
  // (I + U (diag(vals) - I) U^T) a for a register vector (Math::apply_lowrank_transform_inplace, cpu_math.rs:380-425): the r dot
  // products 8 at a time through the team reduction, the update fused behind each group (the dot products use the ORIGINAL a).
  // U is read through the read-only path: it only changes between launches.
  int lr_r = -1;  // eigenvectors of this chain; -1: its transformation has no low-rank part
  __device__ __forceinline__ void lowrank_apply(double (&a)[EPT], const double* __restrict__ vals) {
    if (!LR || lr_r <= 0) return;
    const double* U = P.lr_vecs + (size_t)chain * P.lr_rmax * ld;
    const double* lam = vals + (size_t)chain * P.lr_rmax;
    double acc[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) acc[j] = a[j];
    for (int k0 = 0; k0 < lr_r; k0 += 8) {
      double part[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (k0 + q < lr_r) {
          const double* u = U + (size_t)(k0 + q) * ld;
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            const int i = eidx(j);
            if (i < d) part[q] = fma(__ldg(u + i), a[j], part[q]);
          }
        }
      }
      red.allreduce(part);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (k0 + q < lr_r) {
          const double c = part[q] * (__ldg(lam + k0 + q) - 1.0);
          const double* u = U + (size_t)(k0 + q) * ld;
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            const int i = eidx(j);
            if (i < d) acc[j] = fma(__ldg(u + i), c, acc[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j) a[j] = acc[j];
  }
  // x = F(z): compute_untransformed_position (diagonal.rs:248-256; low_rank.rs:350-378)
  __device__ __forceinline__ void untransform_position(const double (&zz)[EPT], double (&x)[EPT]) {
    if (LR && lr_r >= 0) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) x[j] = zz[j];
      lowrank_apply(x, P.lr_vals_sqrt);
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int i = eidx(j);
        double t = fma(1.0, i < d ? __ldg(P.lr_mu + row + i) : 0.0, x[j]);  // axpy(mu_lr, x, 1)
        t = t * sg(j);                                                       // array_mult_inplace(x, stds)
        x[j] = fma(1.0, mn(j), t);                                           // axpy(mean, x, 1)
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        double t = zz[j] * sg(j);
        x[j] = fma(1.0, mn(j), t);
      }
    }
  }
  // grad_z = J_F^T grad_x: compute_transformed_gradient (diagonal.rs:258-265; low_rank.rs:380-398) into the gradient vector
  __device__ __forceinline__ void transform_gradient(const double (&gx)[EPT]) {
    if (LR && lr_r > 0) {
      double w[EPT];
#pragma unroll
      for (int j = 0; j < EPT; ++j) w[j] = gx[j] * sg(j);
      lowrank_apply(w, P.lr_vals_sqrt);
#pragma unroll
      for (int j = 0; j < EPT; ++j) G(j) = w[j];
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) G(j) = gx[j] * sg(j);
    }
  }

  // ------------------------------------------------------------------ leapfrog (transformed_hamiltonian.rs:524-615, Euclidean)
  // (z, v, g) <- one velocity-Verlet step of size eps in the whitened space; returns logp' and kinetic energy'.
  // When `with_prev` is set it also returns the U-turn products of the pair (previous leaf, new leaf):
  //   sP = (z' - z) . v ,  sQ = (z' - z) . v'   (same operation order as merge_turning; the previous leaf IS the register
  // state before the step, so the most frequent merge - two single leaves - needs no checkpoint load and no extra reduction).
  __device__ __forceinline__ void leapfrog(double eps, double& logp_out, double& ke_out, bool with_prev, double& sP, double& sQ) {
    const double eps_half = eps / 2.;
    hs_total_lf += 1;
    if (ELEMWISE) {
      double part[4];
      leapfrog_partials(eps, part);
      if (TPC > 32 || with_prev) {
        red.allreduce(part);  // CTA teams: always all four sums - ONE reduction routine on the leaf path (instruction cache)
      } else {
        double p2[2] = {part[0], part[1]};
        red.allreduce(p2);
        part[0] = p2[0];
        part[1] = p2[1];
      }
      logp_out = part[0];
      ke_out = 0.5 * part[1];
      sP = part[2];
      sQ = part[3];
      return;
    }
    double x[EPT], gx[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      v[j] = fma(eps_half, G(j), v[j]);
      z[j] = fma(eps, v[j], z[j]);
    }
    untransform_position(z, x);
    double a0, a1, ev;
    model_phase_a(x, a0, a1);
    double part[2];
    part[0] = model_phase_b(x, gx, a0, a1, ev);
    part[1] = 0.0;
    transform_gradient(gx);
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      v[j] = fma(eps_half, G(j), v[j]);
      part[1] = fma(v[j], v[j], part[1]);
    }
    if (MODEL == LOGP_USER) {
      double p4[4] = {part[0], part[1], ue_rec, ue_fatal};
      red.allreduce(p4);
      logp_out = user_verdict(model_finish(p4[0], a0, a1, ev), p4[2], p4[3]);
      ke_out = 0.5 * p4[1];
    } else {
      red.allreduce(part);
      logp_out = model_finish(part[0], a0, a1, ev);
      ke_out = 0.5 * part[1];
    }
    sP = 0.0;
    sQ = 0.0;
  }
  // whether leapfrog() delivers the products of (previous leaf, new leaf)
  static constexpr bool kFusedPrevCheck = ELEMWISE;

  // logp + gradient at the x plane -> gx plane; returns logp (Math::logp_array)
  __device__ __forceinline__ double eval_at_position(double (&x)[EPT], double (&gx)[EPT]) {
    double a0, a1, ev;
    model_phase_a(x, a0, a1);
    double part[1];
    part[0] = model_phase_b(x, gx, a0, a1, ev);
    if (MODEL == LOGP_USER) {
      double p3[3] = {part[0], ue_rec, ue_fatal};
      red.allreduce(p3);
      return user_verdict(model_finish(p3[0], a0, a1, ev), p3[1], p3[2]);
    }
    red.allreduce(part);
    return model_finish(part[0], a0, a1, ev);
  }

  __device__ __forceinline__ void load_mass_matrix() {
    if (MMS) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        sm_sig[sidx(j)] = i < d ? P.stds[row + i] : 0.0;
        sm_mu[sidx(j)] = i < d ? P.mean[row + i] : 0.0;
      }
      // each thread only ever reads back the entries it wrote itself: no barrier needed
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        sig[MMS ? 0 : j] = i < d ? P.stds[row + i] : 0.0;
        mu[MMS ? 0 : j] = i < d ? P.mean[row + i] : 0.0;
      }
    }
  }

  // Large tiles of the decoupled engine (dim ~ 10^4: z and v alone fill 2/3 of the thread's registers): the once-per-draw vector
  // code must not hold further vectors in registers - unrolled over EPT elements it spilled 2.3 KB per thread, and that local
  // traffic (L1 is carved out for shared memory) was 1.4 x the checkpoint traffic of the whole draw.  ROLL: those passes run as
  // rolled loops from and to the global planes, CH elements in flight per thread; the gradient of the chain point stays in its
  // plane (the first half-step of a doubling reads it from there).
  static constexpr bool ROLL = MULTI && EPT > 16 && MODEL == LOGP_GAUSS_DIAG;
  __device__ __forceinline__ double model_mu_at(int i) const { return MSH ? sm_mmu[i] : __ldg(P.model.mu + i); }
  __device__ __forceinline__ double model_prec_at(int i) const { return MSH ? sm_mprec[i] : __ldg(P.model.prec + i); }
  // whiten_from_planes() from plane to plane (x, gx -> z, gz); sigma / mean are in shared memory (SM_MASS)
  __device__ __forceinline__ bool whiten_planes_rolled() {
    const double *xp = P.x + row, *gp = P.gx + row, *ip = P.inv_stds + row;
    double *zp = P.z + row, *gzp = P.gz + row;
    double bad[1] = {0.0};
#pragma unroll 1
    for (int j0 = 0; j0 < EPT; j0 += CH) {
      double x[CH], gx[CH], is[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = eidx(j0 + q);
        const int ii = ((j0 + q < EPT) && i < d) ? i : 0;  // unconditional loads (a lane without an element re-reads element 0)
        x[q] = xp[ii];
        gx[q] = gp[ii];
        is[q] = ip[ii];
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = eidx(j0 + q);
        if ((j0 + q < EPT) && i < d) {
          const double t = fma(-1.0, sm_mu[i], x[q]);  // axpy_out(mean, x, -1)
          const double zz = is[q] * t;                 // multiply_inplace(z, inv_stds): out = x*out
          const double gg = gx[q] * sm_sig[i];
          if (!(isfinite(zz) && isfinite(gg) && (gg != 0.0) && isfinite(gx[q]) && isfinite(x[q]))) bad[0] = 1.0;
          zp[i] = zz;
          gzp[i] = gg;
        }
      }
    }
    red.allreduce(bad);
    return bad[0] == 0.0;
  }

  // compute_transformed_position / _gradient (diagonal.rs:233-246, 258-265) of the chain point from the x, gx planes
  // into the z, g registers; also refreshes the z / gz planes.  Returns check_all() (transformed_hamiltonian.rs:310-324).
  __device__ __forceinline__ bool whiten_from_planes() {
    double x[EPT], gx[EPT], is[EPT];
    load(P.x + row, x);
    load(P.gx + row, gx);
    load(P.inv_stds + row, is);
    double bad[1] = {0.0};
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = eidx(j);
      double t = fma(-1.0, mn(j), x[j]);  // axpy_out(mean, x, -1)
      z[j] = is[j] * t;                   // multiply_inplace(z, inv_stds): out = x*out
      if (LR && lr_r >= 0) z[j] = fma(-1.0, i < d ? __ldg(P.lr_mu + row + i) : 0.0, z[j]);  // axpy(mu_lr, z, -1) (low_rank.rs:337-339)
      if (i >= d) z[j] = 0.0;
    }
    if (LR) lowrank_apply(z, P.lr_vals_sqrt_inv);  // (low_rank.rs:340-345)
    transform_gradient(gx);
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = eidx(j);
      if (i < d) {
        bool ok = isfinite(z[j]) && isfinite(G(j)) && (G(j) != 0.0) && isfinite(gx[j]) && isfinite(x[j]);
        if (!ok) bad[0] = 1.0;
      } else {
        z[j] = 0.0;
        G(j) = 0.0;
      }
    }
    store(P.z + row, z);
    store_g(P.gz + row, false);
    red.allreduce(bad);
    return bad[0] == 0.0;
  }

  // ------------------------------------------------------------------ slot pool
  __device__ __forceinline__ int rc_get(int s) const { return (int)(((s < 32 ? rc_lo : rc_hi) >> (2 * (s & 31))) & 3ull); }
  __device__ __forceinline__ void rc_add(int s, int delta) {
    const uint64_t inc = (uint64_t)(long long)delta << (2 * (s & 31));  // two's complement add on the 2-bit field
    if (s < 32) rc_lo += inc;
    else rc_hi += inc;
  }
  __device__ __forceinline__ int alloc_slot() {
    int s = __ffsll((long long)free_mask) - 1;
    free_mask &= ~(1ull << s);
    return s;  // its count is 0
  }
  __device__ __forceinline__ void unref(int s) {
    rc_add(s, -1);
    if (rc_get(s) == 0) free_mask |= (1ull << s);
  }

  // ------------------------------------------------------------------ cp.async staging of checkpoint pairs (SM_STAGE)
  // Buffer b (0 = X, 1 = Y) holds (z, v) of one checkpoint: sm_stage[(2b + which) * TPC*EPT + element].  Every thread copies
  // exactly the 16-byte pairs it owns under the engine's element mapping and later reads only those, so - like everywhere in this
  // engine - no barrier is involved: completion is the thread's own cp.async.wait_all.  stg_src[b] remembers what was staged.
  const double* stg_src[2] = {nullptr, nullptr};
  __device__ __forceinline__ const double* stage_buf(int b, int which) const { return sm_stage + (size_t)(2 * b + which) * (LT * EPT); }
  __device__ __forceinline__ void stage_pair(int b, const double* zsrc, const double* vsrc) {
    if (!STAGE || (b == 1 && (SMF & SM_STAGE1) != 0)) return;
    stg_src[b] = zsrc;
    const unsigned dz = (unsigned)__cvta_generic_to_shared(stage_buf(b, 0)), dv = (unsigned)__cvta_generic_to_shared(stage_buf(b, 1));
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int i0 = eidx(2 * p);
      if (inb(i0)) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dz + 8u * (unsigned)i0), "l"(zsrc + i0) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dv + 8u * (unsigned)i0), "l"(vsrc + i0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void stage_wait() const {
    if (STAGE) asm volatile("cp.async.wait_all;" ::: "memory");
  }
  // (z, v) of a checkpoint from a staging buffer into the registers (the start state of a doubling after a change of direction)
  __device__ __forceinline__ void load_staged(int b) {
    stage_wait();
    const double *zs = stage_buf(b, 0), *vs = stage_buf(b, 1);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int i0 = eidx(2 * p);
      double2 a = make_double2(0.0, 0.0), c = make_double2(0.0, 0.0);
      if (inb(i0)) a = *reinterpret_cast<const double2*>(zs + i0), c = *reinterpret_cast<const double2*>(vs + i0);
      z[2 * p] = a.x, z[2 * p + 1] = a.y, v[2 * p] = c.x, v[2 * p + 1] = c.y;
    }
  }

  // ------------------------------------------------------------------ U-turn products (is_turning :617-638 -> scalar_prods3)
  // pair (P = earlier built, Q = later built): delta = zQ - zP ; sP = delta . vP ; sQ = delta . vQ.
  // Forward: turning <=> sP < 0 | sQ < 0.  Backward the reference orders the pair the other way round, which negates
  // delta exactly, so turning <=> sP > 0 | sQ > 0.  NaN never turns (reference :637).
  __device__ __forceinline__ bool turn_eval(double sP, double sQ, int dir) const {
    return dir ? ((sP < 0.) | (sQ < 0.)) : ((sP > 0.) | (sQ > 0.));
  }
  // checks of one merge: (Af, cur) always; when `full` also (Al, cur) and (Af, Bf).   cur = registers.
  __device__ __forceinline__ bool merge_turning(const double* Afz, const double* Afv, const double* Alz, const double* Alv,
                                                const double* Bfz, const double* Bfv, bool full, int dir) {
    if (!full && !ELEMWISE) {  // (the elementwise target takes the full path below and ignores the extra products)
      double s[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (inb(i)) {
          double delta = (z[j] + 0.0) - __ldcg(Afz + i);
          s[0] = fma(delta, __ldcg(Afv + i), s[0]);
          s[1] = fma(delta, v[j], s[1]);
        }
      }
      red.allreduce(s);
      return turn_eval(s[0], s[1], dir);
    }
    double s[6];
    if (STAGE) {  // (z, v) of A.first come from shared memory: prefetched by extend(), or staged now when nothing was
      int b = stg_src[0] == Afz ? 0 : (stg_src[1] == Afz ? 1 : -1);
      if (b < 0) {
        b = 0;
        stage_pair(0, Afz, Afv);
      }
      stage_wait();
      merge_products<true>(stage_buf(b, 0), stage_buf(b, 1), Alz, Alv, Bfz, Bfv, true, s);
    } else {
      merge_products<false>(Afz, Afv, Alz, Alv, Bfz, Bfv, true, s);
    }
    red.allreduce(s);
    if (!full) return turn_eval(s[0], s[1], dir);
    return turn_eval(s[0], s[1], dir) | turn_eval(s[2], s[3], dir) | turn_eval(s[4], s[5], dir);
  }

  // ------------------------------------------------------------------ AcceptanceRateCollector::register_leapfrog (dual_avg.rs:131-158)
  // The statistics (exp(min(diff, 0)), its symmetric variant, the signed maximum) only matter at the end of the draw, and their
  // exp + division are ~150 dependent instructions on the critical path of every leaf.  Each leaf just records its energy
  // difference; ACC_RING of them are evaluated together - one exp per lane - and then added in leaf order, so the sums are
  // bit-identical to the reference's sequential accumulation.
  int acc_pending = 0;
  double acc_mine = 0.0;  // energy difference of pending leaf (lane & 31)
  __device__ __forceinline__ void register_leapfrog(double energy, bool divergent) {
    if (divergent) {
      flush_accept();
      max_energy_error = -INFINITY;
    } else {
      if ((int)(threadIdx.x & 31) == acc_pending) acc_mine = E0 - energy;
      acc_pending += 1;
      if (acc_pending == ACC_RING) flush_accept();
    }
    acc_count += 1;
  }
  __device__ __forceinline__ void flush_accept() {
    if (acc_pending == 0) return;
    // by value: a member function that is not inlined would take `this` and pin the whole engine in local memory
    const AccSums r = accept_batch(acc_mine, acc_pending, acc_sum, acc_sym_sum, max_energy_error);
    acc_sum = r.sum;
    acc_sym_sum = r.sym;
    max_energy_error = r.max_err;
    acc_pending = 0;
  }

  // Leave the linear domain for the rest of this draw (a leaf weight could overflow / underflow): the pending sub-trees of the
  // half under construction (levels = the set bits of the leaf index i; level 0 is a log weight already) and the main tree.
  __device__ __forceinline__ void weights_to_log_domain(uint32_t i) {
    lin = false;
    ls_main = log_noinline(ls_main);
    tsync();
    if (ltid == 0) {  // (one writer per copy of the tables: every CTA of a cluster team keeps its own)
#pragma unroll 1
      for (int l = 1; l < MAX_DOUBLING_DEPTH; ++l)
        if ((i >> l) & 1u) T.A_ls[l] = log_noinline(T.A_ls[l]);
    }
    tsync();
    if (TPC > 32) red.local_barrier();
  }

  // ------------------------------------------------------------------ NutsTree::extend for the MAIN tree (nuts.rs:108-170)
  // The reference recursion builds the new half leaf by leaf and merges equal-depth sub-trees in post-order; that is
  // a binary counter: after leaf i (0-based) merge levels 0..t-1, t = number of trailing one bits of i.  Every merge
  // of (A = earlier built, B = later built) checks the span (A.first, B.last) and, when A.depth > 0, (A.last, B.last)
  // and (A.first, B.first) (nuts.rs:143-161), and ALWAYS runs merge_into before the verdict (nuts.rs:163-169), which
  // is also the reference's RNG order.  An inner Turning / Diverging discards the whole half (nuts.rs:131-136).
  __device__ __forceinline__ int extend(int dir, bool check) {
    NB_T0(tq);
    tsync();  // warp teams: every lane has finished reading the tables of the previous doubling
    const int D = depth;
    const uint32_t nleaf = 1u << D;
    const double eps = dir ? hs_step : -hs_step;
    const int sign = dir ? 1 : -1;
    free_mask = P.P >= 64 ? ~0ull : ((1ull << P.P) - 1ull);
    rc_lo = rc_hi = 0;
    if (draw_slot >= 0) {
      free_mask &= ~(1ull << draw_slot);
      rc_add(draw_slot, 1);
    }
    // The ends of the main tree are not copied: (z, v) of an end IS the checkpoint of the leaf that became that end; its slot
    // stays out of the pool like the slot of the main tree's draw.
    if (!init_left) free_mask &= ~(1ull << es_left);
    if (!init_right) free_mask &= ~(1ull << es_right);
    // start state = the end of the main tree in direction dir
    const bool near_init = dir ? init_right : init_left, far_init = dir ? init_left : init_right;
    const int es_near = dir ? es_right : es_left, es_far = dir ? es_left : es_right;
    const double* nearZ = near_init ? P.z + row : slot_ptr(es_near, 0);
    const double* nearV = near_init ? P.v0 + row : slot_ptr(es_near, 1);
    const double* farZ = far_init ? P.z + row : slot_ptr(es_far, 0);
    const double* farV = far_init ? P.v0 + row : slot_ptr(es_far, 1);
    // staging buffers: X is per merge; Y keeps an end of the main tree across doublings (its checkpoint stays out of the pool)
    stg_src[0] = nullptr;
    if (stg_src[1] != farZ && stg_src[1] != nearZ) stg_src[1] = nullptr;
    bool half_done = false;  // NOG: v already holds the first half-step of leaf 0 (done with the stored gradient of the initial point)
    if (NOG) {
      if (near_init) {
        const double eps_half = eps / 2.;
        if (!(dir ? holds_right : holds_left)) {
          load_cg(nearZ, z);
          load_cg(nearV, v);
          const double* gp = P.gz + row;
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            const int i = eidx(j);
            const double gl = inb(i) ? __ldcg(gp + i) : 0.0;
            v[j] = fma(eps_half, gl, v[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < EPT; ++j) v[j] = fma(eps_half, G(j), v[j]);  // gradient loaded by draw_begin()
        }
        half_done = true;
      } else if (!(dir ? holds_right : holds_left)) {
        load_cg(nearZ, z);
        load_cg(nearV, v);
      }
    } else if (!(dir ? holds_right : holds_left)) {
      if (STAGE && stg_src[1] == nearZ) {  // the other end was staged for the last top-level check
        load_staged(1);
      } else {
        load_cg(nearZ, z);
        load_cg(nearV, v);
      }
      if (near_init) {
        load_g(P.gz + row, true);
      } else if (ELEMWISE) {
        // elementwise target: the gradient of a leaf is a function of its z (the leapfrog's own instruction sequence)
#pragma unroll
        for (int j = 0; j < EPT; ++j) G(j) = grad_z_at(z[j], j, eidx(j));
      } else {
        load_g(end_ptr(dir, 2), true);
      }
    }
    holds_left = holds_right = false;
    int s_last = -1;
    int idx_cur = dir ? idx_right : idx_left;
    // the sub-tree B that the newest leaf belongs to
    int B_first = -1, B_draw = -1, B_draw_idx = 0;
    double B_ls = 0., B_draw_energy = 0.;

    for (uint32_t i = 0; i < nleaf; ++i) {
      // single_step (nuts.rs:209-245): one leapfrog from the previous leaf; baseline = initial energy
      double logp_new, ke_new, sP0, sQ0;
      NB_ACC(4, tq);
      // odd leaves merge with their predecessor first (level 0): its U-turn products come out of the leapfrog itself
      const bool fuse0 = kFusedPrevCheck && check && (i & 1u);
      if (STAGE && check) {
        // operands of the checks this leaf will trigger, fetched while its leapfrog runs: A.first of the level-1 merge (its slot
        // is known from the tables: the merges a leaf completes follow from its index), the far end for the top-level check
        if ((i & 3u) == 3u && D >= 2) stage_pair(0, slot_ptr(T.A_first[1], 0), slot_ptr(T.A_first[1], 1));
        if (i + 1 == nleaf && stg_src[1] != farZ) stage_pair(1, farZ, farV);
      }
      if (NOG) {
        double part[4];
        leapfrog_partials_nog(eps, half_done, part);
        half_done = false;
        hs_total_lf += 1;
        if (fuse0) {
          red.allreduce(part);
        } else {
          double p2[2] = {part[0], part[1]};
          red.allreduce(p2);
          part[0] = p2[0];
          part[1] = p2[1];
        }
        logp_new = part[0];
        ke_new = 0.5 * part[1];
        sP0 = part[2];
        sQ0 = part[3];
      } else {
        leapfrog(eps, logp_new, ke_new, fuse0, sP0, sQ0);
      }
      NB_ACC(1, tq);
      hs_tree_lf += 1;
      double energy = ke_new - (logp_new + hs_pt_logdet);
      double energy_error = energy - E0;
      bool divergent = (energy_error > P.s.max_energy_error) | !isfinite(energy_error);
      register_leapfrog(energy, divergent);
      if (divergent) return EXT_DIVERGING;
      idx_cur += sign;
      int s = alloc_slot();
      s_last = s;
      store_cg(slot_ptr(s, 0), z);
      store_cg(slot_ptr(s, 1), v);
      rc_add(s, 3);  // roles: first-of-B, draw-of-B, last-of-B (the newest leaf)
      B_first = s;
      B_draw = s;
      B_ls = -energy_error;  // log weight of the leaf
      B_draw_energy = energy;
      B_draw_idx = idx_cur;
      int t = __ffs(~i) - 1;  // trailing ones of i
      if (t > D) t = D;
      if (lin && fabs(B_ls) > LIN_WEIGHT_LIMIT) weights_to_log_domain(i);  // rare: the weight could leave the double range
      NB_ACC(2, tq);
      for (int l = 0; l < t; ++l) {
        const int Af = T.A_first[l], Al = T.A_last[l];
        bool turning = false;
        if (check) {
          if (l == 0 && fuse0) turning = turn_eval(sP0, sQ0, dir);
          else
            turning = merge_turning(slot_ptr(Af, 0), slot_ptr(Af, 1), slot_ptr(Al, 0), slot_ptr(Al, 1), slot_ptr(B_first, 0),
                                    slot_ptr(B_first, 1), l > 0, dir);
        }
        // merge_into, non-main (nuts.rs:172-207): self_log_size = log_size of the merged tree
        double total;
        bool take_B;
        if (lin) {
          double WA = T.A_ls[l], WB = B_ls;
          if (l == 0) {  // two single leaves: both exponentials in one basic block
            WA = exp_small(WA);
            WB = exp_small(WB);
          }
          total = WA + WB;
          // P(draw from B) = WB / (WA + WB); the reference's `other.log_size >= log_size` shortcut only exists for WA << WB
          if (WA >= WB * 0x1p-40) take_B = rng_f64() * total < WB;
          else take_B = lin_reference_shortcut(WA, WB) || (rng_f64() * total < WB);
        } else {
          const LogMerge mg = log_domain_merge(T.A_ls[l], B_ls, false);
          total = mg.total;
          take_B = mg.shortcut || (rng_f64() < mg.p);
        }
        if (take_B) {
          unref(T.A_draw[l]);
        } else {
          unref(B_draw);
          B_draw = T.A_draw[l];
          B_draw_energy = T.A_draw_energy[l];
          B_draw_idx = T.A_draw_idx[l];
        }
        unref(B_first);
        B_first = Af;
        unref(Al);
        B_ls = total;
        if (turning) return EXT_TURNING;  // inner turn: the old tree is returned unchanged (nuts.rs:131-133)
      }
      stg_src[0] = nullptr;  // buffer X only lives from the prefetch to the merges of the same leaf (slots are re-used afterwards)
      NB_ACC(3, tq);
      if (i + 1 < nleaf) {
        if (ltid == 0) {
          T.A_first[t] = (signed char)B_first;
          T.A_last[t] = (signed char)s;  // the last-of-B reference moves to the pending sub-tree
          T.A_ls[t] = B_ls;
          T.A_draw[t] = (signed char)B_draw;
          T.A_draw_energy[t] = B_draw_energy;
          T.A_draw_idx[t] = B_draw_idx;
        }
        tsync();
      }
    }
    // top-level merge of the main tree (A) with the finished half (B)
    // (measured: routing this check through the level loop to save one inlined copy of the U-turn pass costs the warp-team
    // configurations 7 % - pointer selects and a longer loop on every merge - for 8 % less code; not kept)
    bool turning = false;
    if (check) {
      turning = merge_turning(farZ, farV, nearZ, nearV, slot_ptr(B_first, 0), slot_ptr(B_first, 1), D > 0, dir);
    }
    double total;
    bool take;  // is_main: self_log_size = old log_size (nuts.rs:190)
    if (lin) {
      const double WB = D == 0 ? exp_small(B_ls) : B_ls;  // a single leaf still carries its log weight
      take = (WB >= ls_main) || (rng_f64() * ls_main < WB);
      total = ls_main + WB;
    } else {
      const LogMerge mg = log_domain_merge(ls_main, B_ls, true);
      total = mg.total;
      take = mg.shortcut || (rng_f64() < mg.p);
    }
    if (take) {
      draw_slot = B_draw;
      draw_energy = B_draw_energy;
      draw_idx = B_draw_idx;
    }
    ls_main = total;
    depth += 1;
    if (!ELEMWISE) store_g(end_ptr(dir, 2), true);  // gradient of the new end (not a cheap function of z here)
    if (dir) {
      idx_right = idx_cur;
      es_right = s_last;
      init_right = false;
      holds_right = true;
    } else {
      idx_left = idx_cur;
      es_left = s_last;
      init_left = false;
      holds_left = true;
    }
    NB_ACC(4, tq);
    return turning ? EXT_TURNING : EXT_OK;
  }

  // ------------------------------------------------------------------ DualAverage (dual_avg.rs:33-81)
  // (Adam, stepsize/adam.rs:55-115, lives in the same record: log_step, m = da_hbar, v = da_mu, t = da_count; log_step_adapted
  // mirrors log_step, which is what step_size_bar and the post-warm-up step size are for Adam, stepsize/adapt.rs:256, 288)
  __device__ __forceinline__ void da_new(double initial_step) {
    cs.da_log_step = log(initial_step);
    cs.da_log_step_adapted = log(initial_step);
    if (P.s.method == 1) {
      cs.da_hbar = 0.;
      cs.da_mu = 0.;
      cs.da_count = 0;
      return;
    }
    cs.da_hbar = 0.;
    cs.da_mu = log(10. * initial_step);
    cs.da_count = 1;
  }
  __device__ __forceinline__ void da_advance(double accept_stat) {
    if (P.s.method == 2) return;
    if (P.s.method == 1) {  // Adam::advance
      const double gradient = accept_stat - P.s.target_accept;
      cs.da_count += 1;
      cs.da_hbar = P.s.adam_beta1 * cs.da_hbar + (1.0 - P.s.adam_beta1) * gradient;
      cs.da_mu = P.s.adam_beta2 * cs.da_mu + (1.0 - P.s.adam_beta2) * gradient * gradient;
      const double m_hat = cs.da_hbar / (1.0 - powi_ref(P.s.adam_beta1, (int)cs.da_count));
      const double v_hat = cs.da_mu / (1.0 - powi_ref(P.s.adam_beta2, (int)cs.da_count));
      cs.da_log_step += P.s.adam_lr * m_hat / (sqrt(v_hat) + P.s.adam_epsilon);
      cs.da_log_step_adapted = cs.da_log_step;
      return;
    }
    double cnt = (double)cs.da_count;
    double w = 1. / (cnt + P.s.da_t0);
    cs.da_hbar = (1. - w) * cs.da_hbar + w * (P.s.target_accept - accept_stat);
    cs.da_log_step = cs.da_mu - cs.da_hbar * sqrt(cnt) / P.s.da_gamma;
    cs.da_log_step = fmin(cs.da_log_step, log(P.s.da_max_step));
    double mk = pow(cnt, -P.s.da_k);
    cs.da_log_step_adapted = mk * cs.da_log_step + (1. - mk) * cs.da_log_step_adapted;
    cs.da_count += 1;
  }
  // Strategy::update_stepsize (stepsize/adapt.rs:235-267)
  __device__ __forceinline__ void update_stepsize(bool use_best_guess) {
    double step = P.s.method == 2 ? P.s.fixed_step : (use_best_guess ? exp(cs.da_log_step_adapted) : exp(cs.da_log_step));
    if (P.s.has_jitter) {
      double lo = 1.0 - P.s.jitter, hi = 1.0 + P.s.jitter;
      double j = fma(hi - lo, rng_f64(), lo);
      hs_step = step * j;
    } else {
      hs_step = step;
    }
  }

  // ------------------------------------------------------------------ Strategy::init (stepsize/adapt.rs:91-199)
  // Doubling / halving search from the chain's current position (x, gx planes, hs_logp).  Returns false when
  // init_state fails check_all (NutsError::BadInitGrad).  Uses the ends[0] buffers as scratch for the start state.
  __device__ __forceinline__ bool stepsize_search() {
    if (P.s.method == 2) {
      hs_step = P.s.fixed_step;
      return true;
    }
    load_mass_matrix();
    if (!whiten_from_planes()) return false;  // init_state: same x => same logp / gradient, new whitening
    const double logdet = hs_mm_logdet;
    sample_velocity();  // initialize_trajectory(resample = true)
    double ke[1] = {0.0};
#pragma unroll
    for (int j = 0; j < EPT; ++j) ke[0] = fma(v[j], v[j], ke[0]);
    red.allreduce(ke);
    const double e0 = 0.5 * ke[0] - (hs_logp + logdet);
    double* sz = end_ptr(0, 0);
    double* sv = end_ptr(0, 1);
    double* sg = end_ptr(0, 2);
    store(sz, z);
    store(sv, v);
    store_g(sg, false);
    hs_step = P.s.initial_step;
    double lp, k;
    double u0, u1;
    leapfrog(hs_step, lp, k, false, u0, u1);
    double ee = (k - (lp + logdet)) - e0;
    if ((ee > 1000.0) | !isfinite(ee)) return true;
    double accept = exp(fmin(e0 - (k - (lp + logdet)), 0.)) / 1.0;
    const bool forward = accept > P.s.target_accept;
    for (int it = 0; it < 100; ++it) {
      load(sz, z);
      load(sv, v);
      load_g(sg, false);
      leapfrog(forward ? hs_step : -hs_step, lp, k, false, u0, u1);
      double en = k - (lp + logdet);
      ee = en - e0;
      if ((ee > 1000.0) | !isfinite(ee)) {
        hs_step = P.s.initial_step;
        return true;
      }
      accept = exp(fmin(e0 - en, 0.));
      if (forward) {
        if ((accept <= P.s.target_accept) | (hs_step > 1e5)) {
          da_new(hs_step);
          return true;
        }
        hs_step *= 2.;
      } else {
        if ((accept >= P.s.target_accept) | (hs_step < 1e-10)) {
          da_new(hs_step);
          return true;
        }
        hs_step /= 2.;
      }
    }
    hs_step = P.s.initial_step;
    return true;
  }

  // ------------------------------------------------------------------ the vector work of GlobalStrategy::adapt in ONE pass
  // RunningVariance::add_sample x4 (transform/adapt/diagonal.rs:32-44,134-141; array_update_variance cpu_math.rs:605-631) and
  // Strategy::adapt -> DiagMassMatrix::update_diag_draw_grad / update_diag_draw (transform/diagonal.rs:85-131,
  // cpu_math.rs:633-708) share their operands: the mass matrix is computed from the estimator values this draw has just updated.
  //   upd:      add the draw's (x, grad_x) to estimator sets 0 and 1, whose counts become n0 / n1 (count 1 = first sample)
  //   do_mm:    update stds / inv_stds / mean / logdet from set `mm_set` (the foreground set AFTER a window switch)
  // The elements of a thread are processed in chunks of CH: all loads of a chunk are issued before its first store (the compiler
  // may not move a load of one plane above a store to another), and the 2 divisions + 3 square roots per element run through the
  // branch-free fast paths of device_common.cuh so that the CH dependency chains interleave - with the library operators every
  // call ends in a slow-path branch and the elements of a thread run one after the other (the round-1 tuning phase spent 58 k of
  // its 100 k adaptation cycles per draw there).  sum ln(inv_std) is accumulated as ln(product) per chunk: one log instead of CH.
#ifndef NB_ADAPT_CHUNK
#define NB_ADAPT_CHUNK 2
#endif
  static constexpr int CH = EPT < 4 ? EPT : (EPT % NB_ADAPT_CHUNK == 0 ? NB_ADAPT_CHUNK : 4);  // (the same for both engines: bit-identical results)
  // (everything by value and statically indexed at the call sites: a reference parameter or a rolled loop over the chunk would
  // move the chunk's register arrays to local memory)
  static __device__ __noinline__ double2 mass_matrix_element_slow(double dv, double gv, double scale, bool grad_based, double s_old, double is_old) {
    double val = grad_based ? sqrt(dv / gv) : dv * scale;  // cpu_math.rs:695 / :658
    if (!((!isfinite(val)) | (val == 0.0))) {               // fill_invalid = None: leave untouched
      val = clampd(val, 1e-20, 1e20);
      return make_double2(sqrt(val), sqrt(1.0 / val));
    }
    return make_double2(s_old, is_old);
  }
  __device__ __forceinline__ void adapt_vector_pass(bool upd, uint64_t n0, uint64_t n1, int mm_set, bool do_mm, uint64_t fg_count) {
    const double* xp = P.x + row;
    const double* gp = P.gx + row;
    double* sd = P.stds + row;
    double* isd = P.inv_stds + row;
    double* mn = P.mean + row;
    double* const est0 = est_ptr(0, 0);  // plane (set, which) = est0 + (4 * set + which) * pl; locals: a store cannot alias them
    const size_t pl = (size_t)P.ld;
    const uint64_t nn[2] = {n0, n1};
    const double sc[2] = {1.0 / (double)n0, 1.0 / (double)n1};
    const double mscale = 1.0 / (double)fg_count;
    const bool grad_based = P.s.use_grad_based != 0;
    double ld[1] = {0.0};
#pragma unroll 1
    for (int j0 = 0; j0 < EPT; j0 += CH) {
      double x[CH], gx[CH], e[2][4][CH], s_new[CH], is_new[CH];
      bool live[CH];
      // ---- every load of the chunk, unconditionally (a lane without an element re-reads element 0: no branch per load);
      // PAIR mapping: the two elements of a pair in one 16-byte access (a row's padding is readable)
      constexpr int VS = (PAIR && CH % 2 == 0) ? 2 : 1;
      int at[CH];  // where element q is loaded from
      // (STREAM: the estimator planes are touched once per draw - evict-first, so that they do not push the checkpoints of
      // the running trees out of L2)
      auto ldv = [&](const double* __restrict__ p, double (&a)[CH], auto stream) {
        constexpr bool STREAM = decltype(stream)::value;
#pragma unroll
        for (int q = 0; q < CH; q += VS) {
          if constexpr (VS == 2) {
            const double2* p2 = reinterpret_cast<const double2*>(p + at[q]);
            const double2 t = STREAM ? __ldcs(p2) : *p2;
            a[q] = t.x;
            a[q + 1] = t.y;
          } else {
            a[q] = STREAM ? __ldcs(p + at[q]) : p[at[q]];
          }
        }
      };
      auto stv = [&](double* __restrict__ p, const double (&a)[CH], auto stream) {
        constexpr bool STREAM = decltype(stream)::value;
#pragma unroll
        for (int q = 0; q < CH; q += VS) {
          if constexpr (VS == 2) {
            if (live[q + 1]) {
              if (STREAM) __stcs(reinterpret_cast<double2*>(p + at[q]), make_double2(a[q], a[q + 1]));
              else *reinterpret_cast<double2*>(p + at[q]) = make_double2(a[q], a[q + 1]);
            } else if (live[q]) {
              p[at[q]] = a[q];
            }
          } else {
            if (live[q]) {
              if (STREAM) __stcs(p + at[q], a[q]);
              else p[at[q]] = a[q];
            }
          }
        }
      };
      constexpr std::true_type kStream{};
      constexpr std::false_type kKeep{};
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = eidx(j0 + q);
        live[q] = (j0 + q < EPT) && (i < d);
        at[q] = (VS == 2 && (q & 1)) ? at[q - 1] + 1 : (live[q] ? i : 0);
      }
      ldv(xp, x, kKeep);
      ldv(gp, gx, kKeep);
#pragma unroll
      for (int st = 0; st < 2; ++st)
#pragma unroll
        for (int w = 0; w < 4; ++w) ldv(est0 + (size_t)(4 * st + w) * pl, e[st][w], kStream);
      ldv(sd, s_new, kKeep);
      ldv(isd, is_new, kKeep);
#ifdef NB_PHASE_TIMING_COLD
      const long long vp_a = clock64();
      asm volatile("" ::"d"(x[0]), "d"(is_new[CH - 1]), "d"(e[1][3][CH - 1]), "d"(e[0][0][0]));  // wait for the chunk's loads
      const long long vp_b = clock64();
      cold_t[2] += vp_b - vp_a;
#endif
      // ---- RunningVariance::add_sample for the four estimators of both sets
      if (upd) {
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          if (nn[st] == 1) {
#pragma unroll
            for (int q = 0; q < CH; ++q) {
              e[st][0][q] = x[q];
              e[st][1][q] = 0.0;
              e[st][2][q] = gx[q];
              e[st][3][q] = 0.0;
            }
          } else {
#pragma unroll
            for (int q = 0; q < CH; ++q) {
              // array_update_variance (cpu_math.rs:605-631): both terms use the OLD mean
              const double diff = x[q] - e[st][0][q];
              e[st][0][q] = e[st][0][q] + diff * sc[st];
              e[st][1][q] = e[st][1][q] + diff * diff;
              const double diff1 = gx[q] - e[st][2][q];
              e[st][2][q] = e[st][2][q] + diff1 * sc[st];
              e[st][3][q] = e[st][3][q] + diff1 * diff1;
            }
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) stv(est0 + (size_t)(4 * st + w) * pl, e[st][w], kStream);
        }
      }
#ifdef NB_PHASE_TIMING_COLD
      const long long vp_c = clock64();
      cold_t[3] += vp_c - vp_b;
#endif
      // ---- DiagMassMatrix::update_diag_draw_grad / update_diag_draw from set mm_set
      if (do_mm) {
        double cand_s[CH], cand_is[CH];
        bool valid[CH];
        bool all_ok = true;
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const double dv = mm_set ? e[1][1][q] : e[0][1][q], gv = mm_set ? e[1][3][q] : e[0][3][q];
          bool ok1 = true, ok2 = true;
          double val = grad_based ? sqrt_fast(div_fast(dv, gv, ok1), ok1) : dv * mscale;  // cpu_math.rs:695 / :658
          valid[q] = !((!isfinite(val)) | (val == 0.0));                                 // fill_invalid = None: leave untouched
          val = clampd(val, 1e-20, 1e20);
          cand_s[q] = sqrt_fast(val, ok2);
          cand_is[q] = sqrt_fast(div_fast(1.0, val, ok2), ok2);
          // a failed range test only matters where its result is used: an invalid val (NaN, inf, 0) is discarded anyway
          all_ok = all_ok & (!live[q] | (ok1 & (!valid[q] | ok2)));
        }
        if (all_ok) {
#pragma unroll
          for (int q = 0; q < CH; ++q)
            if (valid[q]) {
              s_new[q] = cand_s[q];
              is_new[q] = cand_is[q];
            }
        } else {  // denormal / huge operands somewhere in the chunk: the library operators, element by element
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            const double dv = mm_set ? e[1][1][q] : e[0][1][q], gv = mm_set ? e[1][3][q] : e[0][3][q];
            if (live[q]) {
              const double2 r = mass_matrix_element_slow(dv, gv, mscale, grad_based, s_new[q], is_new[q]);
              s_new[q] = r.x;
              is_new[q] = r.y;
            }
          }
        }
        double prod = 1.0;  // inv_std in [1e-10, 1e10] after the clamp (or the initial 1 / sqrt|grad| clamp): no overflow for CH <= 4
        double mean[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const double dm = mm_set ? e[1][0][q] : e[0][0][q], gm = mm_set ? e[1][2][q] : e[0][2][q];
          mean[q] = dm;
          if (grad_based) {
            const double var = s_new[q] * s_new[q];  // array_mult(stds, stds, var)
            const double m = var * gm;               // array_mult(var, grad_mean, mean)
            mean[q] = fma(1.0, dm, m);               // axpy(draw_mean, mean, 1.0)
          }
          if (live[q]) prod *= is_new[q];
        }
        stv(sd, s_new, kKeep);
        stv(isd, is_new, kKeep);
        stv(mn, mean, kKeep);
        ld[0] += log(prod);  // array_sum_ln(inv_stds)
      }
#ifdef NB_PHASE_TIMING_COLD
      cold_t[5] += clock64() - vp_c;
#endif
    }
    if (do_mm) {
      red.allreduce(ld);
      hs_mm_logdet = ld[0];
      hs_mm_id += 1;
    }
  }

  // ------------------------------------------------------------------ GlobalStrategy::adapt (adapt_strategy.rs:121-222)
  __device__ __forceinline__ bool adapt(uint64_t draw) {
    // Strategy::update (stepsize/adapt.rs:201-209)
    cs.last_mean_tree_accept = acc_sum / (double)acc_count;
    cs.last_sym_mean_tree_accept = acc_sym_sum / (double)acc_count;
    cs.last_n_steps = acc_count;
    cs.last_max_energy_error = max_energy_error;
    const SettingsDev& S = P.s;
    if (draw >= S.num_tune) {
      update_stepsize(true);
      cs.tuning = 0;
      return true;
    }
    if (draw < S.final_step_size_window) {
      const bool is_early = draw < S.early_end;
      if (!is_early && draw == S.early_end) cs.current_window_size = max(cs.current_window_size, cs.bg_count);
      const uint64_t switch_freq = is_early ? S.early_mm_switch_freq : cs.current_window_size;
      NB_COLD_T(1);
      const bool upd = cs.is_good != 0;  // update_estimators (transform/adapt/diagonal.rs:134-141): counted here, added in the pass below
      if (upd) {
        cs.fg_count += 1;
        cs.bg_count += 1;
      }
      const uint64_t n_set0 = cs.fg_set == 0 ? cs.fg_count : cs.bg_count, n_set1 = cs.fg_set == 0 ? cs.bg_count : cs.fg_count;
      const bool could_switch = cs.bg_count >= switch_freq;
      const uint64_t next_window_size =
          is_early ? S.early_mm_switch_freq
                   : max(cs.current_window_size + 1, (uint64_t)round((double)cs.current_window_size * S.mm_window_growth));
      const bool is_late = next_window_size + draw > S.final_step_size_window;
      bool force_update = false;
      if (could_switch && !is_late) {
        cs.fg_set = 1 - cs.fg_set;  // Strategy::switch: foreground <- background, background <- empty
        cs.fg_count = cs.bg_count;
        cs.bg_count = 0;
        force_update = true;
        if (!is_early) cs.current_window_size = next_window_size;
      }
      // Strategy::adapt needs three samples in the (new) foreground estimator (transform/adapt/diagonal.rs:166-168)
      const bool did_change = (force_update | (draw - cs.last_update >= S.mm_update_freq)) && cs.fg_count >= 3;
      if (upd | did_change) adapt_vector_pass(upd, n_set0, n_set1, cs.fg_set, did_change, cs.fg_count);
      NB_COLD_T(4);
      if (did_change) cs.last_update = draw;
      if (is_late) da_advance(cs.last_sym_mean_tree_accept);
      else da_advance(cs.last_mean_tree_accept);
      if (did_change & (cs.has_initial_mass_matrix != 0)) {
        cs.has_initial_mass_matrix = 0;
        return stepsize_search();
      }
      update_stepsize(false);
      return true;
    }
    da_advance(cs.last_sym_mean_tree_accept);
    update_stepsize(draw == S.num_tune - 1);
    return true;
  }

  // ------------------------------------------------------------------ Chain::draw (chain.rs:151-188) + nuts::draw (nuts.rs:281-388)
  // initialize_trajectory (transformed_hamiltonian.rs:687-736) + collector.register_init (dual_avg.rs:160-165)
  __device__ __forceinline__ void draw_begin() {
    load_mass_matrix();
    if (ROLL) {
      if (hs_mm_id != hs_pt_tid) {
        whiten_planes_rolled();  // inv_transform_normalize: no logp evaluation
        hs_pt_logdet = hs_mm_logdet;
        hs_pt_tid = hs_mm_id;
      }
      load(P.z + row, z);  // the gradient stays in the gz plane
    } else if (hs_mm_id != hs_pt_tid) {
      whiten_from_planes();  // inv_transform_normalize: no logp evaluation
      hs_pt_logdet = hs_mm_logdet;
      hs_pt_tid = hs_mm_id;
    } else {
      load(P.z + row, z);
      load_g(P.gz + row, false);
    }
    sample_velocity();  // also stores the v0 plane
    double ke[1] = {0.0};
#pragma unroll
    for (int j = 0; j < EPT; ++j) ke[0] = fma(v[j], v[j], ke[0]);
    red.allreduce(ke);
    E0 = 0.5 * ke[0] - (hs_logp + hs_pt_logdet);
    acc_sum = 0.;
    acc_sym_sum = 0.;
    acc_count = 0;
    acc_pending = 0;
    max_energy_error = 0.;
  }

  // Materialise the selected draw - the chain point becomes (x, gx, z, gz, logp) of that leaf - then run the per-draw
  // adaptation + statistics (cold, through global memory).  z is read back from its checkpoint; x / logp / gradient are
  // recomputed by the same instruction sequence the leaf used, hence bit-identical to what the leapfrog produced.
  __device__ __forceinline__ void draw_finish(uint64_t t, bool diverging, bool reached_maxdepth) {
    const size_t N = (size_t)P.N;
    NB_T0(tm);
    // While the mass matrix adapts, the pass at the end of this draw reads the chain's 8 estimator planes and 3 mass-matrix planes:
    // last touched a whole draw ago, so they come from DRAM (all chains' planes + the checkpoint traffic of a draw exceed L2), one
    // dependent round trip per chunk.  Ask L2 for them now; the materialisation below covers the latency.
    if (!ROLL && hs_draw_count < P.s.final_step_size_window) {
      const int lines = (d + 15) >> 4;  // rows start on 128-byte lines
      for (int l = tid; l < lines; l += TPC) {
        const size_t o = (size_t)l << 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) prefetch_l2(est_ptr(k >> 2, k & 3) + o);
        prefetch_l2(P.stds + row + o);
        prefetch_l2(P.inv_stds + row + o);
        prefetch_l2(P.mean + row + o);
      }
    }
    double fisher[1] = {0.0};
    if (ROLL) {
      // plane to plane: z of the selected leaf -> (x, gx, z, gz) planes + the draw; logp and the Fisher distance ride along
      const bool moved = draw_slot >= 0;
      const double* zs = moved ? slot_ptr(draw_slot, 0) : P.z + row;
      double* out = P.draws_out ? P.draws_out + (t * N + chain) * (size_t)d : nullptr;
      double lp[1] = {0.0};
#pragma unroll 1
      for (int j0 = 0; j0 < EPT; j0 += CH) {
        double zz[CH], g0[CH], x0[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const int i = eidx(j0 + q);
          const int ii = ((j0 + q < EPT) && i < d) ? i : 0;  // unconditional loads
          zz[q] = __ldcg(zs + ii);
          g0[q] = P.gz[row + ii];
          x0[q] = P.x[row + ii];
        }
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const int i = eidx(j0 + q);
          if ((j0 + q < EPT) && i < d) {
            double gn = g0[q], xn = x0[q];
            if (moved) {
              const double sgm = sm_sig[i];
              const double tt = zz[q] * sgm;
              xn = fma(1.0, sm_mu[i], tt);
              const double diff = xn - model_mu_at(i);
              const double pd = diff * model_prec_at(i);
              lp[0] -= diff * pd / 2.;
              const double gxn = -pd;
              gn = gxn * sgm;
              P.x[row + i] = xn;
              P.gx[row + i] = gxn;
              P.z[row + i] = zz[q];
              P.gz[row + i] = gn;
            }
            if (out) out[i] = xn;
            fisher[0] += (zz[q] + gn) * (zz[q] + gn);  // sq_norm_sum (cpu_math.rs:235-243)
          }
        }
      }
      if (moved) {
        red.allreduce(lp);
        hs_logp = lp[0];
      }
    } else if (draw_slot >= 0) {
      double x[EPT], gx[EPT];
      load_cg(slot_ptr(draw_slot, 0), z);
      untransform_position(z, x);
      hs_logp = eval_at_position(x, gx);
      transform_gradient(gx);
      store(P.x + row, x);
      store(P.gx + row, gx);
      store(P.z + row, z);
      store_g(P.gz + row, false);
      if (P.draws_out) store_dense(P.draws_out + (t * N + chain) * (size_t)d, x);
      if (LR && P.grads_out) store_dense(P.grads_out + (t * N + chain) * (size_t)d, gx);
    } else {
      load(P.z + row, z);
      load_g(P.gz + row, false);
      if (P.draws_out) {
        double x[EPT];
        load(P.x + row, x);
        store_dense(P.draws_out + (t * N + chain) * (size_t)d, x);
      }
      if (LR && P.grads_out) {
        double gx[EPT];
        load(P.gx + row, gx);
        store_dense(P.grads_out + (t * N + chain) * (size_t)d, gx);
      }
    }
    if (!ROLL) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) fisher[0] += (z[j] + G(j)) * (z[j] + G(j));  // sq_norm_sum (cpu_math.rs:235-243)
    }
    red.allreduce(fisher);
    NB_ACC(5, tm);
    // ---- after the tuning phase GlobalStrategy::adapt only copies the collector statistics and draws the next jittered step size
    // (adapt_strategy.rs:126-137 -> stepsize/adapt.rs:201-209, 235-267): done right here, without the round trip of the whole
    // ChainState through the cold function (same arithmetic, same random stream: bit-identical to the cold path)
    if (hs_draw_count >= P.s.num_tune) {
      const double mta = acc_sum / (double)acc_count, msa = acc_sym_sum / (double)acc_count;
      const double step_bar = P.s.method == 2 ? P.s.fixed_step : exp(hs_da_lsa);
      if (P.s.has_jitter) {
        const double lo = 1.0 - P.s.jitter, hi = 1.0 + P.s.jitter;
        hs_step = step_bar * fma(hi - lo, rng_f64(), lo);
      } else {
        hs_step = step_bar;
      }
      if (tid == 0) {
        ChainState* g = P.cs + chain;
        g->last_mean_tree_accept = mta;
        g->last_sym_mean_tree_accept = msa;
        g->last_n_steps = acc_count;
        g->last_max_energy_error = max_energy_error;
        g->is_good = (diverging ? (abs(draw_idx) > 4) : (draw_idx != 0)) ? 1 : 0;
        g->tuning = 0;
        g->draw_count = hs_draw_count + 1;
        const StatsDev& st = P.stats;
        const size_t k = (size_t)t * N + chain;
        if (st.depth) st.depth[k] = (uint64_t)depth;
        if (st.maxdepth_reached) st.maxdepth_reached[k] = reached_maxdepth ? 1 : 0;
        if (st.index_in_trajectory) st.index_in_trajectory[k] = draw_idx;
        if (st.logp) st.logp[k] = hs_logp;
        if (st.energy) st.energy[k] = draw_energy;
        if (st.energy_error) st.energy_error[k] = draw_energy - E0;
        if (st.diverging) st.diverging[k] = diverging ? 1 : 0;
        if (st.step_size) st.step_size[k] = hs_step;
        if (st.step_size_bar) st.step_size_bar[k] = step_bar;
        if (st.mean_tree_accept) st.mean_tree_accept[k] = mta;
        if (st.mean_tree_accept_sym) st.mean_tree_accept_sym[k] = msa;
        if (st.n_steps) st.n_steps[k] = acc_count;
        if (st.max_energy_error) st.max_energy_error[k] = max_energy_error;
        if (st.tuning) st.tuning[k] = 0;
        if (st.fisher_distance) st.fisher_distance[k] = fisher[0];
      }
      store_hot();
      NB_ACC(6, tm);
      return;
    }
    // ---- adaptation + statistics: cold, through global memory
    store_hot();
    const int ret = cold_adapt<TPC, EPT, SMF, MODEL, MULTI>(P, chain, tid, red.scratch, sm_sig, mc, (red.parity & 1) | ((red.xparity & 1) << 1), t, acc_sum, acc_sym_sum, acc_count,
                                              max_energy_error, diverging ? (abs(draw_idx) > 4) : (draw_idx != 0), depth, reached_maxdepth,
                                              diverging, draw_idx, draw_energy, draw_energy - E0, fisher[0]);
    red.parity = ret & 1;
    red.xparity = (ret >> 1) & 1;
    // the chain scalars now live in ChainState again; the next work unit (possibly on another team) reloads them
    NB_ACC(6, tm);
  }

  __device__ __forceinline__ void run_draw(uint64_t t) {
    const SettingsDev& S = P.s;
    NB_T0(td);
    NB_T0(tw);
    draw_begin();
    NB_ACC(0, td);
    // NutsTree::new (nuts.rs:94-105): log_size 0 = weight 1
    stg_src[0] = stg_src[1] = nullptr;
    lin = true;
    ls_main = 1.;
    depth = 0;
    idx_left = idx_right = 0;
    init_left = init_right = true;
    es_left = es_right = -1;
    holds_left = holds_right = true;
    draw_slot = -1;
    draw_energy = E0;
    draw_idx = 0;
    uint64_t mindepth = S.mindepth, maxdepth = S.maxdepth;
    if (S.has_target_time) {  // nuts.rs:300-320
      uint64_t max_steps = (uint64_t)ceil(S.target_time / hs_step);
      mindepth = max((uint64_t)floor(log2((double)max_steps)), S.mindepth);
      maxdepth = min(max((uint64_t)ceil(log2((double)max_steps)), mindepth), S.maxdepth);
    }
    // nuts.rs:333-385 with ONE extend() call site (the whole tree builder is inlined here): the regular doubling loop, then
    // the optional extra doublings in the same direction without turn checks (nuts.rs:349-374).
    bool diverging = false, reached_maxdepth = false, extra_mode = false;
    uint64_t extra_left = 0;
    int dir = 0;
    for (;;) {
      bool check;
      if (!extra_mode) {
        if (!((uint64_t)depth < maxdepth)) {
          reached_maxdepth = true;
          break;
        }
        dir = rng_bool() ? 1 : 0;  // hamiltonian.rs:111-119: true => Forward
        check = S.check_turning && !((uint64_t)depth < mindepth);
      } else {
        if (extra_left == 0) break;
        extra_left -= 1;
        check = false;
      }
      const int r = extend(dir, check);
      if (r == EXT_DIVERGING) {
        diverging = true;
        break;
      }
      if (!extra_mode && r == EXT_TURNING) {
        extra_mode = true;
        extra_left = S.extra_doublings;
      }
    }
    flush_accept();
    if (MODEL == LOGP_USER && user_fatal) {
      // LogpError::is_recoverable() == false: NutsError::LogpFailure ends the chain (nuts.rs:231); this draw and all later ones of
      // the launch read NaN / "nothing happened", the chain is reported dead (nuts_chain_state_t::alive = 0)
      hs_alive = 0;
      store_hot();
      if (tid == 0) P.cs[chain].alive = 0;
      team_sync();
      cold_fill_dead(P, chain, tid, TPC, t);
      return;
    }
    draw_finish(t, diverging, reached_maxdepth);
    NB_ACC(7, tw);
  }

  // ================================================================== decoupled engine, vector side (MULTI only) =============
  // The team's warps run the vector work of one doubling without ever waiting for a scalar verdict: leapfrog, checkpoint
  // store, the U-turn products of every merge this leaf completes (their operands are structural: a binary counter), all
  // bundled into ONE ring entry per leaf that the leader lane consumes up to V2_K leaves later.  When the leader ends the
  // doubling early (inner U-turn, divergence) the leaves computed ahead are simply dropped.
  __device__ __forceinline__ double* endbuf_ptr(int buf, int which) const { return ends_base + (size_t)((buf * 3 + which) * ld); }

  // U-turn products of one merge, per-thread partials (no reduction): (Af, cur) always; when `full` also (Al, cur), (Af, Bf).
  // AF_SMEM: Afz / Afv point into a staging buffer (shared memory) instead of the checkpoint pool
  template <bool AF_SMEM>
  __device__ __forceinline__ void merge_products(const double* Afz, const double* Afv, const double* Alz, const double* Alv,
                                                 const double* Bfz, const double* Bfv, bool full, double (&s)[6]) {
    s[0] = s[1] = s[2] = s[3] = s[4] = s[5] = 0.0;
    if (!full) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (inb(i)) {
          double delta = (z[j] + 0.0) - __ldcg(Afz + i);
          s[0] = fma(delta, __ldcg(Afv + i), s[0]);
          s[1] = fma(delta, v[j], s[1]);
        }
      }
      return;
    }
    if (PAIR) {  // 16-byte checkpoint reads; the partner of an odd last element is zero padding on both sides (adds +0.0)
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = eidx(2 * p);
        if (inb(i0)) {
          const double2 afz = AF_SMEM ? *reinterpret_cast<const double2*>(Afz + i0) : __ldcg(reinterpret_cast<const double2*>(Afz + i0)),
                        afv = AF_SMEM ? *reinterpret_cast<const double2*>(Afv + i0) : __ldcg(reinterpret_cast<const double2*>(Afv + i0)),
                        alz = __ldcg(reinterpret_cast<const double2*>(Alz + i0)), alv = __ldcg(reinterpret_cast<const double2*>(Alv + i0)),
                        bfz = __ldcg(reinterpret_cast<const double2*>(Bfz + i0)), bfv = __ldcg(reinterpret_cast<const double2*>(Bfv + i0));
          merge_element(z[2 * p], v[2 * p], afz.x, afv.x, alz.x, alv.x, bfz.x, bfv.x, s);
          merge_element(z[2 * p + 1], v[2 * p + 1], afz.y, afv.y, alz.y, alv.y, bfz.y, bfv.y, s);
        }
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = eidx(j);
      if (inb(i)) merge_element(z[j], v[j], __ldcg(Afz + i), __ldcg(Afv + i), __ldcg(Alz + i), __ldcg(Alv + i), __ldcg(Bfz + i), __ldcg(Bfv + i), s);
    }
  }
  // the six U-turn products of one element: pairs (Af, cur), (Al, cur), (Af, Bf)
  static __device__ __forceinline__ void merge_element(double zc, double vc, double afz, double afv, double alz, double alv, double bfz,
                                                       double bfv, double (&s)[6]) {
    const double d1 = (zc + 0.0) - afz;
    s[0] = fma(d1, afv, s[0]);
    s[1] = fma(d1, vc, s[1]);
    const double d2 = (zc + 0.0) - alz;
    s[2] = fma(d2, alv, s[2]);
    s[3] = fma(d2, vc, s[3]);
    const double d3 = (bfz + 0.0) - afz;
    s[4] = fma(d3, afv, s[4]);
    s[5] = fma(d3, bfv, s[5]);
  }

  // gradient of the diagonal Gaussian in the whitened space at whitened position zz, element j: the very instruction sequence
  // the leapfrog uses for the new point, so recomputing it is bit-identical to having stored it
  __device__ __forceinline__ double grad_z_at(double zz, int j, int i) const {
    const double sgm = sg(j);
    const double t = zz * sgm;
    const double xn = fma(1.0, mn(j), t);
    double gxn = 0.0;
    if (inb(i)) {
      const double diff = xn - model_mu(j, i);
      const double pd = diff * model_prec(j, i);
      gxn = -pd;
    }
    return gxn * sgm;
  }
  // leapfrog_partials without a stored gradient vector: grad_z of the current point is recomputed from z (5 more flops per
  // element instead of a shared-memory load + store; shared-memory bandwidth is what bounds 7 resident chains per SM).
  // `half_done`: v already holds the first half-step (done with the LOADED gradient when the state came from memory: after a
  // mass-matrix change the stored gradient is not a function of the stored z, see whiten_from_planes).
  __device__ __forceinline__ void leapfrog_partials_nog(double eps, bool half_done, double (&part)[4]) {
    const double eps_half = eps / 2.;
    part[0] = part[1] = part[2] = part[3] = 0.0;
    PairConsts pc;
    if (MSTREAM) {
#pragma unroll
      for (int j = 0; j < V2_MRING - 1; ++j) mstream_issue(j);
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const int i = eidx(j);
      if (!MSTREAM && PAIR && (j & 1) == 0) pair_consts(j, pc);
      const double zp = z[j], vp = v[j];
      const double sgm = (!MSTREAM && PAIR) ? pc.sg[j & 1] : sg(j), mnj = (!MSTREAM && PAIR) ? pc.mn[j & 1] : mn(j);
      double mmu = 0.0, mprec = 0.0;
      if (MSTREAM) {
        mstream_issue(j + V2_MRING - 1);
        mstream_wait(j, mmu, mprec);
      } else if (inb(i)) {
        mmu = PAIR ? pc.mm[j & 1] : model_mu(j, i);
        mprec = PAIR ? pc.pr[j & 1] : model_prec(j, i);
      }
      double g0;
      {
        const double t = zp * sgm;
        const double x0 = fma(1.0, mnj, t);
        double gx0 = 0.0;
        if (inb(i)) {
          const double diff = x0 - mmu;
          const double pd = diff * mprec;
          gx0 = -pd;
        }
        g0 = gx0 * sgm;
      }
      const double vh1 = fma(eps_half, g0, vp);    // first_velocity_halfstep :178-184  axpy_out(grad, v, eps/2)
      const double vh = half_done ? vp : vh1;
      const double zn = fma(eps, vh, zp);          // position_step :220-225            axpy_out(v', z, eps)
      const double t = zn * sgm;                   // compute_untransformed_position    diagonal.rs:253-255
      const double xn = fma(1.0, mnj, t);
      double gxn = 0.0;
      if (inb(i)) {
        const double diff = xn - mmu;
        const double pd = diff * mprec;
        part[0] -= diff * pd / 2.;
        gxn = -pd;
      }
      const double gn = gxn * sgm;                 // compute_transformed_gradient      diagonal.rs:258-265
      const double vn = fma(eps_half, gn, vh);     // second_velocity_halfstep :245-247 axpy(grad', v, eps/2)
      part[1] = fma(vn, vn, part[1]);              // update_kinetic_energy :260-262
      if (inb(i)) {
        const double delta = (zn + 0.0) - zp;
        part[2] = fma(delta, vp, part[2]);
        part[3] = fma(delta, vn, part[3]);
      }
      z[j] = zn;
      v[j] = vn;
    }
  }

  struct VecTree {  // what the vector side knows about the main tree: all of it follows from the command sequence
    bool holds_left, holds_right;  // the registers currently hold that end
    int es_left, es_right;         // checkpoint slot holding (z, v) of that end; -1: the end still is the initial point
    int last_slot;                 // slot of the last leaf of the doubling that was just built
  };

  // entry i may be written once the leader has consumed leaf i - V2_K; false when a new command arrived instead (abort)
  __device__ __forceinline__ bool v2_wait_credit(const V2Ctl& c, unsigned epoch, unsigned i) const {
    for (;;) {
      if (v2_ld(&c.cmd_seq) != epoch) return false;
      if (i < (unsigned)V2_K) return true;
      const unsigned cv = v2_ld(&c.cons);
      if ((cv >> 12) == (epoch & 0xFFFFFu) && (cv & 0xFFFu) + (unsigned)V2_K > i) return true;
      __nanosleep(NB_V2_SLEEP_CREDIT);
    }
  }

  // one doubling (2^D leaves from the `dir` end of the main tree); returns true when the leader aborted it.
  // The ends of the main tree are not copied anywhere: (z, v) of an end IS the checkpoint of the leaf that became that end (the
  // leader keeps the two end slots out of the pool, like the slot of the main tree's draw), and its gradient is recomputed from z.
  __device__ __forceinline__ bool extend_vec(V2Ctl& c, unsigned epoch, int dir, bool check, int D, VecTree& vt) {
    static_assert(!MULTI || MODEL == LOGP_GAUSS_DIAG, "the decoupled engine covers the elementwise (diagonal Gaussian) target");
    static_assert(!MULTI || !GS, "the decoupled engine recomputes grad_z instead of keeping it");
    constexpr int W = TPC / 32;
    const int w = v2_warp, lane = threadIdx.x & 31;
    const uint32_t nleaf = 1u << D;
    const double eps = dir ? hs_step : -hs_step;
    const int es_near = dir ? vt.es_right : vt.es_left, es_far = dir ? vt.es_left : vt.es_right;
    bool half_done = false;
    if (es_near < 0) {
      // start from the initial point of the draw: first half-step with the gradient draw_begin() loaded (after a mass-matrix
      // change it is not a function of z, see whiten_from_planes)
      const double eps_half = eps / 2.;
      const bool held = dir ? vt.holds_right : vt.holds_left;
      if (!held) {
        load_cg(P.z + row, z);
        load_cg(P.v0 + row, v);
      }
      if (!held || ROLL) {  // ROLL: the gradient of the chain point is never held in registers
        const double* gp = P.gz + row;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          const int i = eidx(j);
          const double gl = inb(i) ? __ldcg(gp + i) : 0.0;
          v[j] = fma(eps_half, gl, v[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) v[j] = fma(eps_half, G(j), v[j]);
      }
      half_done = true;
    } else if (!(dir ? vt.holds_right : vt.holds_left)) {
      load_cg(slot_ptr(es_near, 0), z);
      load_cg(slot_ptr(es_near, 1), v);
    }
    vt.holds_left = vt.holds_right = false;
    NB_T0(tq);
    for (uint32_t i = 0; i < nleaf; ++i) {
      NB_ACC(4, tq);
      if (!v2_wait_credit(c, epoch, i)) return true;
      NB_ACC(0, tq);
      __threadfence_block();
      const int s = c.slot_ring[i & 7];
      double part[4];
      leapfrog_partials_nog(eps, half_done, part);
      half_done = false;
      NB_ACC(1, tq);
      store_cg(slot_ptr(s, 0), z);
      store_cg(slot_ptr(s, 1), v);
      double* e = v2_ring + ((size_t)(i % V2_K) * W + w) * V2_NV;
      {
        const double mine = warp_reduce_scatter8<4>(part);
        if (lane < 4) e[lane] = mine;
      }
      int t = __ffs(~i) - 1;  // trailing ones of i: the merges this leaf completes (levels 0 .. t-1)
      if (t > D) t = D;
      NB_ACC(2, tq);
      const bool last = i + 1 == nleaf;
      // merges whose products are loaded from checkpoints: inner levels 1 .. t-1 (level 0 = (previous leaf, this leaf) came out of
      // the leapfrog itself), then - after the last leaf - the top-level merge of the main tree with the finished half.
      // B's first leaf at level l is A_first[l-1] (the merged tree of the levels below), the new leaf itself at level 0.
      const int nm = check ? ((t > 1 ? t - 1 : 0) + (last ? 1 : 0)) : 0;
      for (int m = 0; m < nm; ++m) {
        const int l = m + 1;
        const bool top = l >= t;  // only the last iteration of the last leaf
        const int Bf = t > 0 ? c.VA_first[w][(top ? t : l) - 1] : s;
        const double *az, *av, *lz, *lv;
        if (!top) {
          const int Af = c.VA_first[w][l], Al = c.VA_last[w][l];
          az = slot_ptr(Af, 0);
          av = slot_ptr(Af, 1);
          lz = slot_ptr(Al, 0);
          lv = slot_ptr(Al, 1);
        } else {  // A = the main tree: first = far end, last = near end (for D == 0 both are the initial point; extras unused)
          az = es_far < 0 ? P.z + row : slot_ptr(es_far, 0);
          av = es_far < 0 ? P.v0 + row : slot_ptr(es_far, 1);
          lz = es_near < 0 ? P.z + row : slot_ptr(es_near, 0);
          lv = es_near < 0 ? P.v0 + row : slot_ptr(es_near, 1);
        }
        double sp[6];
        merge_products<false>(az, av, lz, lv, slot_ptr(Bf, 0), slot_ptr(Bf, 1), true, sp);
        const double mine = warp_reduce_scatter8<6>(sp);
        if (lane < 6) e[4 + 6 * m + lane] = mine;
      }
      if (!last) {
        __syncwarp();
        if (lane == 0) {
          c.VA_first[w][t] = (signed char)(t > 0 ? c.VA_first[w][t - 1] : s);
          c.VA_last[w][t] = (signed char)s;
        }
      } else {
        vt.last_slot = s;
      }
      __syncwarp();  // entry values and table stores of every lane are ordered before the flag
      if (lane == 0) {
        __threadfence_block();
        c.prod[w] = ((epoch & 0xFFFFFu) << 12) | (i + 1);
      }
      NB_ACC(3, tq);
    }
    return false;
  }

  __device__ __forceinline__ void run_draw_v2(uint64_t t, unsigned& cmd_seen) {
    V2Ctl& c = *v2_ctl;
    draw_begin();
    if (tid == 0) {  // start record: the leader lane takes over the scalar side of the tree
      c.E0 = E0;
      c.pt_logdet = hs_pt_logdet;
      c.step = hs_step;
      c.rng = hs_rng;
      c.pk_logp = hs_logp;
      c.pk_mm_logdet = hs_mm_logdet;
      c.pk_pt_tid = hs_pt_tid;
      c.pk_mm_id = hs_mm_id;
      c.pk_total_lf = hs_total_lf;
      c.pk_tree_lf = hs_tree_lf;
      __threadfence_block();
      c.start_seq = v2_ld(&c.start_seq) + 1;
    }
    VecTree vt;
    vt.holds_left = vt.holds_right = true;
    vt.es_left = vt.es_right = -1;
    vt.last_slot = -1;
    bool pending = false;
    int prev_dir = 0;
    NB_T0(tw);
    NB_T0(tc);
    for (;;) {
      unsigned cs;
      NB_ACC(4, tc);
      while ((cs = v2_ld(&c.cmd_seq)) == cmd_seen) __nanosleep(NB_V2_SLEEP);
      NB_ACC(5, tc);
      cmd_seen = cs;
      __threadfence_block();
      const int kind = c.cmd_kind, dir = c.cmd_dir, check = c.cmd_check, D = c.cmd_depth, accepted = c.cmd_prev_accepted;
      if (pending && accepted) {  // the finished half became part of the main tree: its last leaf is the new `prev_dir` end
        if (prev_dir) {
          vt.es_right = vt.last_slot;
          vt.holds_right = true;
        } else {
          vt.es_left = vt.last_slot;
          vt.holds_left = true;
        }
      }
      pending = false;
      if (kind == V2_CMD_TREE_DONE) break;
      pending = !extend_vec(c, cs, dir, check != 0, D, vt);
      prev_dir = dir;
    }
    team_sync();  // every warp of the team has left the tree
    // result record + the parked chain scalars
    acc_sum = c.acc_sum;
    acc_sym_sum = c.acc_sym_sum;
    max_energy_error = c.max_energy_error;
    acc_count = c.acc_count;
    hs_rng = c.rng_out;
    E0 = c.E0;
    hs_pt_logdet = c.pt_logdet;
    hs_logp = c.pk_logp;
    hs_mm_logdet = c.pk_mm_logdet;
    hs_pt_tid = c.pk_pt_tid;
    hs_mm_id = c.pk_mm_id;
    hs_tree_lf = c.pk_tree_lf + acc_count;
    hs_total_lf = c.pk_total_lf + acc_count;
    depth = c.depth;
    draw_slot = c.draw_slot;
    draw_idx = c.draw_idx;
    draw_energy = c.draw_energy;
    const bool diverging = c.diverging != 0, reached_maxdepth = c.reached_maxdepth != 0;
    NB_ACC(7, tw);
    draw_finish(t, diverging, reached_maxdepth);
  }

  // ------------------------------------------------------------------ Chain::set_position (chain.rs:137-149)
  __device__ __forceinline__ int run_set_position() {
    double x[EPT], gx[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {  // dense [N][d] rows: not 16-byte aligned for odd d
      const int i = eidx(j);
      x[j] = i < d ? P.init_position[(size_t)chain * d + i] : 0.0;
    }
    // GlobalStrategy::init -> init_state_untransformed (transformed_hamiltonian.rs:663-685)
    hs_logp = eval_at_position(x, gx);
    double bad[1] = {0.0};
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = eidx(j);
      if (i < d && !(isfinite(x[j]) && isfinite(gx[j]))) bad[0] = 1.0;
    }
    red.allreduce(bad);
    if (bad[0] != 0.0) return 3;
    if (MODEL == LOGP_USER && !isfinite(hs_logp)) return 3;  // the user density reported an error at the initial point
    store(P.x + row, x);
    store(P.gx + row, gx);
    // mass_matrix_adapt.init (transform/adapt/diagonal.rs:209-231): seed all four estimators, update_diag_grad
    cs.fg_set = 0;
    cs.fg_count = 1;
    cs.bg_count = 1;
    adapt_vector_pass(true, 1, 1, 0, false, 1);  // reads the x / gx planes stored above
    double ld[1] = {0.0};
    {
      double* sd = P.stds + row;
      double* isd = P.inv_stds + row;
      double* mn = P.mean + row;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        int i = eidx(j);
        if (i < d) {
          double val = 1.0 / clampd(fabs(gx[j]), 1e-20, 1e20);  // cpu_math.rs:710-738
          if (!isfinite(val)) val = 1.0;
          double s_new = sqrt(val), is_new = sqrt(1.0 / val);
          sd[i] = s_new;
          isd[i] = is_new;
          double var = s_new * s_new;
          double m = var * gx[j];
          mn[i] = fma(1.0, x[j], m);
          ld[0] += log(is_new);
        }
      }
    }
    red.allreduce(ld);
    hs_mm_logdet = ld[0];
    hs_mm_id += 1;
    // step_size.init
    if (P.s.method != 2) {
      if (!stepsize_search()) return 3;
    } else {
      hs_step = P.s.fixed_step;
    }
    // self.state = hamiltonian.init_state(position) (transformed_hamiltonian.rs:640-661)
    load_mass_matrix();
    if (!whiten_from_planes()) return 3;
    hs_pt_logdet = hs_mm_logdet;
    hs_pt_tid = hs_mm_id;
    return 0;
  }
};

// GlobalStrategy::adapt + the statistics of Chain::expanded_draw for one chain; returns the reduction parity (bit 0).
template <int TPC, int EPT, int SMF, int MODEL, bool MULTI>
__device__ __noinline__ int cold_adapt(const EngineParams& Pg, int chain, int tid, double* scratch, double* team_smem, const MultiCtx* mc, int parity,
                                       uint64_t t, double acc_sum, double acc_sym_sum, uint64_t acc_count, double max_energy_error, bool is_good,
                                       int depth, bool reached_maxdepth, bool diverging, int draw_idx, double pt_energy,
                                       double pt_energy_error, double fisher) {
  // A private copy of the launch parameters.  Through the reference (a generic pointer to the kernel's parameter space) every
  // store in this function could alias them as far as the compiler knows, and each plane pointer was re-read - an L2 round trip,
  // the loads bypass L1 - between two consecutive stores: 16 dependent round trips per chunk of the estimator pass, 15 in the
  // statistics block (measured: 71 k + 16 k of the 100 k cycles a tuning draw of config 2 spent here).
  const ParamsCopy<!MULTI> pc(Pg);
  const EngineParams& P = pc.P;
  TreeTables& tables = *reinterpret_cast<TreeTables*>(reinterpret_cast<unsigned char*>(team_smem) + (smem_vectors<SMF>() * (size_t)(TPC / cluster_size<SMF>()) * EPT * sizeof(double)));
  Engine<TPC, EPT, SMF, MODEL, MULTI> E(P, chain, tid, scratch, team_smem, tables, mc);
  E.red.parity = parity & 1;
  E.red.xparity = (parity >> 1) & 1;
  NB_COLD_T0;
  E.cold_load();
  E.acc_sum = acc_sum;
  E.acc_sym_sum = acc_sym_sum;
  E.acc_count = acc_count;
  E.max_energy_error = max_energy_error;
  E.cs.is_good = is_good ? 1 : 0;  // register_draw (transform/adapt/diagonal.rs:74-83)
  const bool ok = E.adapt(E.cs.draw_count);
  E.cs.draw_count += 1;
  if (!ok) {
    E.hs_alive = 0;
    cold_fill_dead(Pg, chain, tid, TPC, t + 1);  // (the reference: the copy's address must not escape)
  }
  if (tid == 0) {
    const StatsDev& st = P.stats;
    const size_t k = (size_t)t * (size_t)P.N + chain;
    if (st.depth) st.depth[k] = (uint64_t)depth;
    if (st.maxdepth_reached) st.maxdepth_reached[k] = reached_maxdepth ? 1 : 0;
    if (st.index_in_trajectory) st.index_in_trajectory[k] = draw_idx;
    if (st.logp) st.logp[k] = E.hs_logp;
    if (st.energy) st.energy[k] = pt_energy;
    if (st.energy_error) st.energy_error[k] = pt_energy_error;
    if (st.diverging) st.diverging[k] = diverging ? 1 : 0;
    if (st.step_size) st.step_size[k] = E.hs_step;
    if (st.step_size_bar) st.step_size_bar[k] = P.s.method == 2 ? P.s.fixed_step : exp(E.cs.da_log_step_adapted);
    if (st.mean_tree_accept) st.mean_tree_accept[k] = E.cs.last_mean_tree_accept;
    if (st.mean_tree_accept_sym) st.mean_tree_accept_sym[k] = E.cs.last_sym_mean_tree_accept;
    if (st.n_steps) st.n_steps[k] = E.cs.last_n_steps;
    if (st.max_energy_error) st.max_energy_error[k] = E.cs.last_max_energy_error;
    if (st.tuning) st.tuning[k] = E.cs.tuning ? 1 : 0;
    if (st.fisher_distance) st.fisher_distance[k] = fisher;
  }
  E.cold_store();
#ifdef NB_PHASE_TIMING_COLD
  if (tid == 0 && P.phase_clocks) {  // [8] cold_load .. schedule, [9] the vector pass, [10] dual averaging / step size / stats / cold_store, [15] total
    const long long t_end = clock64();
    const long long* c = E.cold_t;
    if (c[1]) {
      atomicAdd(P.phase_clocks + 8, (unsigned long long)(c[1] - cold_t0));
      atomicAdd(P.phase_clocks + 9, (unsigned long long)(c[4] - c[1]));
      atomicAdd(P.phase_clocks + 10, (unsigned long long)(t_end - c[4]));
      atomicAdd(P.phase_clocks + 11, (unsigned long long)c[2]);  // vector pass: waiting for a chunk's loads
      atomicAdd(P.phase_clocks + 12, (unsigned long long)c[3]);  // vector pass: estimator updates + their stores
      atomicAdd(P.phase_clocks + 13, (unsigned long long)c[5]);  // vector pass: mass-matrix elements + their stores
    }
    atomicAdd(P.phase_clocks + 15, (unsigned long long)(t_end - cold_t0));
  }
#endif
  return (E.red.parity & 1) | ((E.red.xparity & 1) << 1);
}

// Chain::set_position for one chain; returns the per-chain status (0 ok, 3 bad initial point).
template <int TPC, int EPT, int SMF, int MODEL, bool MULTI>
__device__ __noinline__ int cold_set_position(const EngineParams& Pg, int chain, int tid, double* scratch, double* team_smem, const MultiCtx* mc) {
  const ParamsCopy<!MULTI> pc(Pg);  // as in cold_adapt
  const EngineParams& P = pc.P;
  TreeTables& tables = *reinterpret_cast<TreeTables*>(reinterpret_cast<unsigned char*>(team_smem) + (smem_vectors<SMF>() * (size_t)(TPC / cluster_size<SMF>()) * EPT * sizeof(double)));
  Engine<TPC, EPT, SMF, MODEL, MULTI> E(P, chain, tid, scratch, team_smem, tables, mc);
  E.cold_load();
  const int status = E.run_set_position();
  E.hs_alive = status == 0 ? 1 : 0;
  E.cold_store();
  return status;
}

// After the host replaced a chain's transformation (nuts_sampler_set_lowrank_transform): the first mass-matrix update of a run
// re-initialises the step size from the current point (GlobalStrategy::adapt, adapt_strategy.rs:204-214 -> stepsize/adapt.rs:91-199);
// the point itself is re-whitened by the next draw (transformation id changed).  Returns 0, or 3 when the search failed.
template <int TPC, int EPT, int SMF, int MODEL, bool MULTI>
__device__ __noinline__ int cold_retransform(const EngineParams& Pg, int chain, int tid, double* scratch, double* team_smem, const MultiCtx* mc) {
  const ParamsCopy<!MULTI> pc(Pg);
  const EngineParams& P = pc.P;
  TreeTables& tables = *reinterpret_cast<TreeTables*>(reinterpret_cast<unsigned char*>(team_smem) + (smem_vectors<SMF>() * (size_t)(TPC / cluster_size<SMF>()) * EPT * sizeof(double)));
  Engine<TPC, EPT, SMF, MODEL, MULTI> E(P, chain, tid, scratch, team_smem, tables, mc);
  E.cold_load();
  int status = 0;
  if (E.hs_alive && E.cs.has_initial_mass_matrix != 0 && E.cs.tuning != 0) {
    E.cs.has_initial_mass_matrix = 0;
    if (!E.stepsize_search()) {
      status = 3;
      E.hs_alive = 0;
    }
  }
  E.cold_store();
  return status;
}

// Draws t0.. of this launch that a dead chain never produces: NaN positions, and statistics that say "nothing happened"
// (0 leapfrogs, depth 0, NaN floats) instead of whatever the buffers held before.
static __device__ __noinline__ void cold_fill_dead(const EngineParams& P, int chain, int tid, int tpc, uint64_t t0) {
  const double nan = __longlong_as_double(-1ll);
  for (uint64_t t = t0; t < P.n_draws; ++t) {
    if (P.draws_out) {
      double* dst = P.draws_out + (t * (size_t)P.N + chain) * (size_t)P.d;
      for (int i = tid; i < P.d; i += tpc) dst[i] = nan;
    }
    if (tid == 0) {
      const StatsDev& st = P.stats;
      const size_t k = (size_t)t * (size_t)P.N + chain;
      if (st.depth) st.depth[k] = 0;
      if (st.maxdepth_reached) st.maxdepth_reached[k] = 0;
      if (st.index_in_trajectory) st.index_in_trajectory[k] = 0;
      if (st.logp) st.logp[k] = nan;
      if (st.energy) st.energy[k] = nan;
      if (st.energy_error) st.energy_error[k] = nan;
      if (st.diverging) st.diverging[k] = 0;
      if (st.step_size) st.step_size[k] = nan;
      if (st.step_size_bar) st.step_size_bar[k] = nan;
      if (st.mean_tree_accept) st.mean_tree_accept[k] = nan;
      if (st.mean_tree_accept_sym) st.mean_tree_accept_sym[k] = nan;
      if (st.n_steps) st.n_steps[k] = 0;
      if (st.max_energy_error) st.max_energy_error[k] = nan;
      if (st.tuning) st.tuning[k] = 0;
      if (st.fisher_distance) st.fisher_distance[k] = nan;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- work units
// set_position: one unit per chain.  Draws: a unit is draws_per_unit consecutive draws of one chain; a chain's state lives in
// global memory between units, so ANY team can run a chain's next unit.  Units are handed out through a ready queue: the first
// N pop tickets are the chains themselves (unit 0 of every chain), and a team that completes a unit of a chain that has more
// to do pushes the chain into a bounded multi-producer / multi-consumer ring (a sequence number per slot); pop ticket N + T
// takes what push ticket T delivered.  First in, first out keeps the launch draw-major (all chains advance together, the launch
// ends after total_units / teams rounds rather than ceil(N / teams) whole chains), but a team never waits for one PARTICULAR
// chain: with the trees of the warm-up phase ranging from 1 to 1023 leapfrogs, the fixed order unit u = (draw u / N, chain u % N)
// of round 1 left the team that drew the successor of a long draw asleep until that draw ended (12 - 15 % of the tuning phase
// of config 2 in the ncu stall samples).  Release / acquire at gpu scope on the slot's sequence number orders the chain's planes
// (the engine's global loads bypass L1: -dlcm=cg).  Both functions are called by ONE thread of the team.
__device__ __forceinline__ unsigned unit_pop(const EngineParams& P, unsigned total_units, unsigned& blk) {
  const unsigned h = atomicAdd(P.queue, 1u);
  blk = 0;
  if (h >= total_units) return ~0u;
  if (h < (unsigned)P.N) return h;
  const unsigned T = h - (unsigned)P.N, slot = T & P.ring_mask;
  for (;;) {
    unsigned sq;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(sq) : "l"(P.ring_seq + slot) : "memory");
    if (sq == T + 1u) break;
    __nanosleep(100);
  }
  const unsigned chain = *reinterpret_cast<volatile unsigned*>(P.ring_chain + slot);
  blk = *reinterpret_cast<volatile unsigned*>(P.done + chain);
  const unsigned freed = T + P.ring_mask + 1u;  // the slot now waits for push ticket T + ring size
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(P.ring_seq + slot), "r"(freed) : "memory");
  return chain;
}
// (the caller has made the team's stores visible: __threadfence by every thread, then a team barrier)
__device__ __forceinline__ void unit_push(const EngineParams& P, unsigned chain, unsigned blocks_done, unsigned blocks) {
  *reinterpret_cast<volatile unsigned*>(P.done + chain) = blocks_done;
  if (blocks_done >= blocks) return;  // the chain's last unit of this launch
  const unsigned T = atomicAdd(P.queue + 1, 1u), slot = T & P.ring_mask;
  for (;;) {  // never spins in practice: at most N chains are in flight and the ring has >= N slots
    unsigned sq;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(sq) : "l"(P.ring_seq + slot) : "memory");
    if (sq == T) break;
    __nanosleep(100);
  }
  *reinterpret_cast<volatile unsigned*>(P.ring_chain + slot) = chain;
  const unsigned filled = T + 1u;
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(P.ring_seq + slot), "r"(filled) : "memory");
}

// One kernel for both Chain::set_position (mode 0) and n_draws x Chain::draw (mode 1).
// Dynamic shared memory: TEAMS x team_smem_bytes<TPC, EPT, SMF>().
template <int TPC, int EPT, int CTA_THREADS, int MIN_BLOCKS, int SMF, int MODEL>
__global__ void NB_KERNEL_BOUNDS(CTA_THREADS, MIN_BLOCKS) nuts_chain_kernel(const __grid_constant__ EngineParams P) {
  constexpr int TEAMS = CTA_THREADS / TPC;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  static_assert(TEAMS == 1 || TPC <= 32 || (SMF & SM_ALIGN), "several CTA-sized teams per CTA: SM_ALIGN builds only");
  constexpr int kScratch = TPC > 32 ? 2 * (TPC / 32) * REDUCE_MAXK : 1;
  __shared__ double scratch_all[TEAMS * kScratch];
  __shared__ int next_chain[TEAMS][2];
  const int team = threadIdx.x / TPC;
  const int tid = threadIdx.x % TPC;
  double* scratch = scratch_all + team * kScratch;
  auto team_bar = [&]() {
    if (TPC <= 32) __syncwarp();
    else if (TEAMS > 1) bar_sync(1 + team, TPC);
    else __syncthreads();
  };
  unsigned char* my_smem = dyn_smem + (size_t)team * team_smem_bytes<TPC, EPT, SMF>();
  double* team_smem = reinterpret_cast<double*>(my_smem);
  TreeTables& tables = *reinterpret_cast<TreeTables*>(my_smem + (smem_vectors<SMF>() * (size_t)(TPC / cluster_size<SMF>()) * EPT * sizeof(double)));
  {  // model parameters: the same for every chain, loaded once per team
    Engine<TPC, EPT, SMF, MODEL> E0(P, 0, tid, scratch, team_smem, tables);
    E0.load_model_params();
  }
  const unsigned B = P.draws_per_unit;
  const unsigned blocks = P.mode != 1 ? 1u : ((unsigned)P.n_draws + B - 1u) / B;
  const unsigned total_units = (unsigned)P.N * blocks;
  bool active = true;
  for (;;) {
    if ((SMF & SM_ALIGN) && TEAMS > 1) {
      // The barrier comes BEFORE the pop: a team waiting here holds no chain (it has pushed the one it finished), so a team whose
      // pop has to wait for a push can never be waiting for a chain that is parked at a barrier.  Every team takes part until
      // the queue is empty for all of them.
      if (!__syncthreads_or(active)) break;
      if (!active) continue;
    }
    if (tid == 0) {
      unsigned b;
      next_chain[team][0] = (int)unit_pop(P, total_units, b);
      next_chain[team][1] = (int)b;
    }
    team_bar();
    const int chain = next_chain[team][0];
    const unsigned blk = (unsigned)next_chain[team][1];  // block of B consecutive draws (a few draws per unit amortise the hand-over)
    team_bar();
    if (chain < 0) {
      if ((SMF & SM_ALIGN) && TEAMS > 1) {
        active = false;
        continue;
      }
      break;
    }
    Engine<TPC, EPT, SMF, MODEL> E(P, chain, tid, scratch, team_smem, tables);
    if (P.mode == 0) {
      if (P.init_mask == nullptr || P.init_mask[chain] != 0) {
        const int status = cold_set_position<TPC, EPT, SMF, MODEL, false>(P, chain, tid, scratch, team_smem, nullptr);
        if (tid == 0 && P.status_out) P.status_out[chain] = status;
      }
    } else if (P.mode == 2) {
      const int status = cold_retransform<TPC, EPT, SMF, MODEL, false>(P, chain, tid, scratch, team_smem, nullptr);
      if (tid == 0 && P.status_out) P.status_out[chain] = status;
    } else {
      if (blk > 0) __threadfence();  // (thread 0 acquired the chain in unit_pop; the team barrier above ordered the others after it)
      const uint64_t t_end = min((uint64_t)(blk + 1u) * B, (uint64_t)P.n_draws);
      for (uint64_t t = (uint64_t)blk * B; t < t_end; ++t) {
        E.load_hot();
        if (E.hs_alive) {
          E.run_draw(t);  // a chain that dies here has its remaining draws NaN-filled by cold_adapt
        } else {
          // draws a dead chain never produced read NaN (the output buffer may be host memory the kernel writes directly)
          if (t == 0) cold_fill_dead(P, chain, tid, TPC, 0);
          break;
        }
      }
      __threadfence();
      E.team_sync();
      if (tid == 0) unit_push(P, (unsigned)chain, blk + 1u, blocks);
    }
    team_bar();
#ifdef NB_PHASE_TIMING
    if (tid == 0 && P.phase_clocks)
      for (int k = 0; k < 8; ++k) atomicAdd(P.phase_clocks + k, (unsigned long long)E.phase[k]);
#endif
  }
}

// The same engine with the team spread over the CL CTAs of a thread-block cluster (SM_CL2 / SM_CL4): CL SMs work on ONE chain.
// For dim ~ 10^4 a single SM cannot hold a chain's vectors (z, v, sigma, mean, model parameters: 6 x 80 KB); a cluster of four
// holds all of them in registers and shared memory, every leapfrog touches HBM only for its checkpoint, and the register file
// leaves room for the redundant scalar state again (no leader warp, no spills).  Work units as in nuts_chain_kernel, fetched by
// CTA 0 of the cluster and handed to the peers through distributed shared memory.
// Launched with cluster dimension CL (cudaLaunchKernelEx); blockDim = TPC / CL; grid = clusters * CL.
template <int TPC, int EPT, int SMF, int MODEL>
__global__ void __launch_bounds__(TPC / cluster_size<SMF>(), 1) nuts_chain_kernel_cluster(const __grid_constant__ EngineParams P) {
  constexpr int CL = cluster_size<SMF>();
  constexpr int LT = TPC / CL;
  static_assert(CL > 1 && LT % 32 == 0 && LT > 32, "cluster teams: at least two warps per CTA");
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ double scratch[TeamReduce<LT, false, CL>::SCRATCH_DOUBLES];
  __shared__ unsigned next_unit[2];
  const unsigned crank = cluster_ctarank();
  const int tid = (int)crank * LT + (int)threadIdx.x;
  double* team_smem = reinterpret_cast<double*>(dyn_smem);
  TreeTables& tables = *reinterpret_cast<TreeTables*>(dyn_smem + (smem_vectors<SMF>() * (size_t)LT * EPT * sizeof(double)));
  {
    Engine<TPC, EPT, SMF, MODEL> E0(P, 0, tid, scratch, team_smem, tables);
    E0.load_model_params();
  }
  const unsigned B = P.draws_per_unit;
  const unsigned blocks = P.mode == 0 ? 1u : ((unsigned)P.n_draws + B - 1u) / B;
  const unsigned total_units = (unsigned)P.N * blocks;
  // distributed shared memory may only be touched once every CTA of the cluster is running (racecheck: "block that might not have
  // entered yet"); the exit side is covered by the two barriers of the last iteration
  cluster_barrier();
  for (;;) {
    if (tid == 0) {
      unsigned b;
      const unsigned c = unit_pop(P, total_units, b);
#pragma unroll
      for (int r = 0; r < CL; ++r) {
        st_cluster_u32(&next_unit[0], (unsigned)r, c);
        st_cluster_u32(&next_unit[1], (unsigned)r, b);
      }
    }
    cluster_barrier();
    const int chain = (int)next_unit[0];
    const unsigned blk = next_unit[1];
    cluster_barrier();  // everybody has read it before CTA 0 fetches the next one
    if (chain < 0) break;
    Engine<TPC, EPT, SMF, MODEL> E(P, chain, tid, scratch, team_smem, tables);
    if (P.mode == 0) {
      if (P.init_mask == nullptr || P.init_mask[chain] != 0) {
        const int status = cold_set_position<TPC, EPT, SMF, MODEL, false>(P, chain, tid, scratch, team_smem, nullptr);
        if (tid == 0 && P.status_out) P.status_out[chain] = status;
      }
    } else {
      if (blk > 0) __threadfence();
      const uint64_t t_end = min((uint64_t)(blk + 1u) * B, (uint64_t)P.n_draws);
      for (uint64_t t = (uint64_t)blk * B; t < t_end; ++t) {
        E.load_hot();
        if (E.hs_alive) {
          E.run_draw(t);
        } else {
          if (t == 0) cold_fill_dead(P, chain, tid, TPC, 0);
          break;
        }
      }
      __threadfence();
      cluster_barrier();
      if (tid == 0) unit_push(P, (unsigned)chain, blk + 1u, blocks);
    }
    cluster_barrier();
  }
}

}  // namespace nb
