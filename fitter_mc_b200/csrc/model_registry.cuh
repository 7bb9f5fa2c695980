// Registry that binds each compiled-in model plug-in to its kernel instantiations.
#pragma once
#include <cuda_runtime.h>

#include "bootstrap_kernel.cuh"

namespace fmc {

struct ModelEntry {
  const char* name;
  int nparam, nprop;
  const char* pnames[16];
  const char* qnames[16];
  double init[16];
  double (*model_y)(const double*, double);   // host copies of the plug-in routines
  void (*props)(const double*, double*);
  cudaError_t (*launch)(const KernelArgs&, bool inject, bool smem, int grid, size_t shmem, cudaStream_t);
  cudaError_t (*occupancy)(bool inject, bool smem, size_t shmem, int* blocks_per_sm);
};

template <class M>
struct ModelBinding {
  static double host_model_y(const double* p, double x) { return M::model_y(p, x); }
  static void host_props(const double* p, double* q) { M::derived_properties(p, q); }

  template <bool INJECT, bool SMEM>
  static cudaError_t prep(size_t shmem) {
    return cudaFuncSetAttribute(bootstrap_fit_kernel<M, INJECT, SMEM>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
  }
  static cudaError_t launch(const KernelArgs& a, bool inject, bool smem, int grid, size_t shmem,
                            cudaStream_t st) {
    cudaError_t e;
#define FMC_LAUNCH(I, S)                                                         \
  do {                                                                           \
    e = prep<I, S>(shmem);                                                       \
    if (e != cudaSuccess) return e;                                              \
    bootstrap_fit_kernel<M, I, S><<<grid, FMC_THREADS, shmem, st>>>(a);          \
    return cudaGetLastError();                                                   \
  } while (0)
    if (inject) {
      if (smem) FMC_LAUNCH(true, true);
      else FMC_LAUNCH(true, false);
    } else {
      if (smem) FMC_LAUNCH(false, true);
      else FMC_LAUNCH(false, false);
    }
#undef FMC_LAUNCH
  }
  static cudaError_t occupancy(bool inject, bool smem, size_t shmem, int* nb) {
    cudaError_t e;
#define FMC_OCC(I, S)                                                                           \
  do {                                                                                          \
    e = prep<I, S>(shmem);                                                                      \
    if (e != cudaSuccess) return e;                                                             \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, bootstrap_fit_kernel<M, I, S>,     \
                                                         FMC_THREADS, shmem);                   \
  } while (0)
    if (inject) {
      if (smem) FMC_OCC(true, true);
      else FMC_OCC(true, false);
    } else {
      if (smem) FMC_OCC(false, true);
      else FMC_OCC(false, false);
    }
#undef FMC_OCC
  }
  static ModelEntry entry() {
    ModelEntry e{};
    const char* name = "";
    const char* pn[16] = {0};
    const char* qn[16] = {0};
    M::initialise_model(&name, &pn, &qn, e.init);
    e.name = name;
    e.nparam = M::NPARAM;
    e.nprop = M::NPROP;
    for (int i = 0; i < 16; ++i) {
      e.pnames[i] = pn[i];
      e.qnames[i] = qn[i];
    }
    e.model_y = host_model_y;
    e.props = host_props;
    e.launch = launch;
    e.occupancy = occupancy;
    return e;
  }
};

// one translation unit per model (kernels_model*.cu) so that they compile in parallel
ModelEntry fmc_entry_model0();
ModelEntry fmc_entry_model1();
ModelEntry fmc_entry_model2();
ModelEntry fmc_entry_model3();

}  // namespace fmc
