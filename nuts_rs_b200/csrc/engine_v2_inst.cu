// engine_v2_inst.cu — one instantiation of the decoupled chain kernel (chain_engine_v2.cuh) per (threads-per-team,
// elements-per-thread, teams-per-CTA, leader warps) tuple: -DCFG_TPC=.. -DCFG_EPT=.. -DCFG_C=.. -DCFG_NL=.. -DCFG_TAG=.. -DCFG_MODEL=..  The
// launcher symbols follow the naming of engine_inst.cu with the tag (>= 100, last digit = teams per CTA) in place of "min blocks",
// so capi.cu's configuration table treats both kinds alike.
#include "chain_engine_v2.cuh"

#define NB_CAT2(a, b, c, d, e) a##_##b##_##c##_##d##_##e
#define NB_CAT(a, b, c, d, e) NB_CAT2(a, b, c, d, e)
#ifndef CFG_NL
#define CFG_NL 1
#endif
#define NB_KERNEL nb::nuts_chain_kernel_v2<CFG_TPC, CFG_EPT, CFG_C, CFG_MODEL, CFG_NL>

static constexpr int kSmf = nb::V2Layout<CFG_TPC, CFG_EPT, CFG_C>::SMF;
static constexpr size_t kSmem = nb::V2Layout<CFG_TPC, CFG_EPT, CFG_C>::total;
static constexpr int kThreads = 32 * CFG_NL + CFG_C * CFG_TPC;
static_assert(kSmem <= 227 * 1024, "the layout exceeds the 227 KB of dynamic shared memory a CTA may use on sm_100");

extern "C" cudaError_t NB_CAT(nb_launch_chain, CFG_TPC, CFG_EPT, CFG_TAG, CFG_MODEL)(const nb::EngineParams* p, int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured && kSmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(NB_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
  }
  configured = true;
  NB_KERNEL<<<grid, kThreads, kSmem, stream>>>(*p);
  return cudaGetLastError();
}

extern "C" cudaError_t NB_CAT(nb_occupancy_chain, CFG_TPC, CFG_EPT, CFG_TAG, CFG_MODEL)(int* blocks_per_sm, int* cta_threads, int* smf) {
  *smf = kSmf;
  *cta_threads = kThreads;
  if (kSmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(NB_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
  }
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, NB_KERNEL, kThreads, kSmem);
}
