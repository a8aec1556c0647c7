// engine_inst.cu — one instantiation of the chain kernel per (threads-per-chain, elements-per-thread) pair.
// Compiled once per configuration with -DCFG_TPC=.. -DCFG_EPT=.. -DCFG_CTA=.. -DCFG_MINB=.. -DCFG_MMS=.. -DCFG_MODEL=.. -DCFG_TAG=.. (see Makefile;
// the tag names the symbols: it equals MINB except for variants of one tiling that differ in their flags) so
// the configurations build in parallel.
#include "chain_engine.cuh"

#define NB_CAT2(a, b, c, d, e) a##_##b##_##c##_##d##_##e
#define NB_CAT(a, b, c, d, e) NB_CAT2(a, b, c, d, e)
#define NB_KERNEL nb::nuts_chain_kernel<CFG_TPC, CFG_EPT, CFG_CTA, CFG_MINB, CFG_MMS, CFG_MODEL>

static constexpr int kSmf = CFG_MMS;  // shared-memory / padding flags (chain_engine.cuh SM_*)
static constexpr size_t kSmem = (CFG_CTA / CFG_TPC) * nb::team_smem_bytes<CFG_TPC, CFG_EPT, CFG_MMS>();

extern "C" cudaError_t NB_CAT(nb_launch_chain, CFG_TPC, CFG_EPT, CFG_TAG, CFG_MODEL)(const nb::EngineParams* p, int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured && kSmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(NB_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
  }
  configured = true;
  NB_KERNEL<<<grid, CFG_CTA, kSmem, stream>>>(*p);
  return cudaGetLastError();
}

// resident CTAs per SM for this configuration (grid sizing of the persistent kernel)
extern "C" cudaError_t NB_CAT(nb_occupancy_chain, CFG_TPC, CFG_EPT, CFG_TAG, CFG_MODEL)(int* blocks_per_sm, int* cta_threads, int* smf) {
  *smf = kSmf;
  *cta_threads = CFG_CTA;
  if (kSmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(NB_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
  }
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, NB_KERNEL, CFG_CTA, kSmem);
}
