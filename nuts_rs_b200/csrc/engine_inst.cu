// engine_inst.cu — one instantiation of the chain kernel per (threads-per-chain, elements-per-thread) pair.
// Compiled once per configuration with -DCFG_TPC=.. -DCFG_EPT=.. -DCFG_CTA=.. -DCFG_MINB=.. (see Makefile) so the
// configurations build in parallel.
#include "chain_engine.cuh"

#define NB_CAT2(a, b, c) a##_##b##_##c
#define NB_CAT(a, b, c) NB_CAT2(a, b, c)

extern "C" cudaError_t NB_CAT(nb_launch_chain, CFG_TPC, CFG_EPT)(const nb::EngineParams* p, int grid, cudaStream_t stream) {
  nb::nuts_chain_kernel<CFG_TPC, CFG_EPT, CFG_CTA, CFG_MINB><<<grid, CFG_CTA, 0, stream>>>(*p);
  return cudaGetLastError();
}

// resident CTAs per SM for this configuration (grid sizing of the persistent kernel)
extern "C" cudaError_t NB_CAT(nb_occupancy_chain, CFG_TPC, CFG_EPT)(int* blocks_per_sm, int* cta_threads) {
  *cta_threads = CFG_CTA;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, nb::nuts_chain_kernel<CFG_TPC, CFG_EPT, CFG_CTA, CFG_MINB>,
                                                       CFG_CTA, 0);
}
