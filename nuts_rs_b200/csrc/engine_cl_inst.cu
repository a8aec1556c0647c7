// engine_cl_inst.cu — one instantiation of the CLUSTER chain kernel (chain_engine.cuh nuts_chain_kernel_cluster): a team of
// CFG_TPC threads spread over the CTAs of a thread-block cluster (cluster size from the SM_CL2 / SM_CL4 bit of CFG_MMS).
//   -DCFG_TPC=.. (threads of the whole team) -DCFG_EPT=.. -DCFG_MMS=.. -DCFG_MODEL=.. -DCFG_TAG=..
// Symbols follow engine_inst.cu, so capi.cu's configuration table treats all engine kinds alike.
#include "chain_engine.cuh"

#define NB_CAT2(a, b, c, d, e) a##_##b##_##c##_##d##_##e
#define NB_CAT(a, b, c, d, e) NB_CAT2(a, b, c, d, e)
#define NB_KERNEL nb::nuts_chain_kernel_cluster<CFG_TPC, CFG_EPT, CFG_MMS, CFG_MODEL>

static constexpr int kSmf = CFG_MMS;
static constexpr int kCl = nb::cluster_size<CFG_MMS>();
static constexpr int kThreads = CFG_TPC / kCl;
static constexpr size_t kSmem = nb::team_smem_bytes<CFG_TPC, CFG_EPT, CFG_MMS>();
static_assert(kSmem <= 227 * 1024, "a CTA's share of the team exceeds the 227 KB of dynamic shared memory");

static cudaError_t configure() {
  static bool configured = false;
  if (!configured && kSmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(NB_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
  }
  configured = true;
  return cudaSuccess;
}

static void fill_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, cudaStream_t stream) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// grid = clusters * cluster size
extern "C" cudaError_t NB_CAT(nb_launch_chain, CFG_TPC, CFG_EPT, CFG_TAG, CFG_MODEL)(const nb::EngineParams* p, int grid, cudaStream_t stream) {
  cudaError_t e = configure();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(cfg, attr, grid, stream);
  return cudaLaunchKernelEx(&cfg, NB_KERNEL, *p);
}

// blocks_per_sm = MINUS the number of clusters that can be resident on the device at once (cluster teams: the persistent grid is
// that many clusters, not CTAs per SM)
extern "C" cudaError_t NB_CAT(nb_occupancy_chain, CFG_TPC, CFG_EPT, CFG_TAG, CFG_MODEL)(int* blocks_per_sm, int* cta_threads, int* smf) {
  *smf = kSmf;
  *cta_threads = kThreads;
  cudaError_t e = configure();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_config(cfg, attr, kCl, nullptr);
  int clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&clusters, NB_KERNEL, &cfg);
  *blocks_per_sm = -clusters;
  return e;
}
