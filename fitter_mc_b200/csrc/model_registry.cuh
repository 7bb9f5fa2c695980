// Registry that binds each compiled-in model plug-in to its kernel instantiations.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

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
  bool analytic;  // can form exact Jacobian rows (generic model_y or its own model_grad)
  int naux;       // per-row auxiliary values the model declares (pass_eval.cuh), 0 = none
  void (*row_aux)(double x, double* aux);
  // stage A (uniform first pass) and stage B (refill iterations) of one nl2sno call;
  // jac = JAC_FD (nl2sno) or JAC_ANALYTIC (nl2sol + madj)
  cudaError_t (*launch_init)(const KernelArgs&, bool inject, bool smem, int jac, size_t shmem, cudaStream_t);
  // shared-Jacobian variant of stage A for the first call (all fits start at params_start)
  cudaError_t (*launch_jac0)(const KernelArgs&, double* rowJ, double* G0, cudaStream_t);
  cudaError_t (*launch_init_shared)(const KernelArgs&, const double* rowJ, const double* G0, bool inject,
                                    size_t shmem, cudaStream_t);
  cudaError_t (*launch_iter)(const KernelArgs&, bool inject, bool smem, int jac, int grid, size_t shmem,
                             cudaStream_t);
  cudaError_t (*occupancy)(bool inject, bool smem, int jac, size_t shmem, int* blocks_per_sm);
};

template <class M>
struct ModelBinding {
  static double host_model_y(const double* p, double x) { return M::model_y(p, x); }
  static void host_props(const double* p, double* q) { M::derived_properties(p, q); }
  static void host_row_aux(double x, double* aux) {
    if constexpr (model_naux<M>::value > 0) M::row_aux(x, aux);
  }

  static constexpr bool kAnalytic = can_analytic_v<M>;

  template <bool INJECT, bool SMEM, int JAC>
  static cudaError_t prep(size_t shmem) {
    cudaError_t e = cudaFuncSetAttribute(iterate_kernel<M, INJECT, SMEM, JAC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(init_pass_kernel<M, INJECT, SMEM, JAC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)shmem);
  }
  // run f(INJECT, SMEM, JAC) with the runtime flags turned into compile-time constants
  template <class F>
  static cudaError_t dispatch(bool inject, bool smem, int jac, F&& f) {
    using JF = std::integral_constant<int, JAC_FD>;
    using JA = std::integral_constant<int, JAC_ANALYTIC>;
    if (jac == JAC_ANALYTIC) {
      if constexpr (kAnalytic) {
        if (inject)
          return smem ? f(std::true_type{}, std::true_type{}, JA{}) : f(std::true_type{}, std::false_type{}, JA{});
        return smem ? f(std::false_type{}, std::true_type{}, JA{}) : f(std::false_type{}, std::false_type{}, JA{});
      } else {
        return cudaErrorNotSupported;  // a double-only model without model_grad has no derivative
      }
    }
    if (inject) return smem ? f(std::true_type{}, std::true_type{}, JF{}) : f(std::true_type{}, std::false_type{}, JF{});
    return smem ? f(std::false_type{}, std::true_type{}, JF{}) : f(std::false_type{}, std::false_type{}, JF{});
  }

  static cudaError_t launch_init(const KernelArgs& a, bool inject, bool smem, int jac, size_t shmem,
                                 cudaStream_t st) {
    const int grid = (int)((a.nres + FMC_THREADS - 1) / FMC_THREADS);
    return dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      init_pass_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>
          <<<grid, FMC_THREADS, shmem, st>>>(a);
      return cudaGetLastError();
    });
  }
  static cudaError_t launch_jac0(const KernelArgs& a, double* rowJ, double* G0, cudaStream_t st) {
    jac0_kernel<M><<<1, FMC_JAC0_THREADS, 0, st>>>(a, rowJ, G0);
    return cudaGetLastError();
  }
  static cudaError_t launch_init_shared(const KernelArgs& a, const double* rowJ, const double* G0, bool inject,
                                        size_t shmem, cudaStream_t st) {
    const int grid = (int)((a.nres + FMC_THREADS - 1) / FMC_THREADS);
    cudaError_t e;
    if (inject) {
      e = cudaFuncSetAttribute(init_shared_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
      if (e != cudaSuccess) return e;
      init_shared_kernel<M, true><<<grid, FMC_THREADS, shmem, st>>>(a, rowJ, G0);
    } else {
      e = cudaFuncSetAttribute(init_shared_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
      if (e != cudaSuccess) return e;
      init_shared_kernel<M, false><<<grid, FMC_THREADS, shmem, st>>>(a, rowJ, G0);
    }
    return cudaGetLastError();
  }
  static cudaError_t launch_iter(const KernelArgs& a, bool inject, bool smem, int jac, int grid, size_t shmem,
                                 cudaStream_t st) {
    return dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      iterate_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>
          <<<grid, FMC_THREADS, shmem, st>>>(a);
      return cudaGetLastError();
    });
  }
  static cudaError_t occupancy(bool inject, bool smem, int jac, size_t shmem, int* nb) {
    return dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      cudaError_t e = prep<decltype(I)::value, decltype(S)::value, decltype(J)::value>(shmem);
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          nb, iterate_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>, FMC_THREADS, shmem);
    });
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
    e.analytic = kAnalytic;
    e.naux = model_naux<M>::value;
    e.row_aux = host_row_aux;
    e.model_y = host_model_y;
    e.props = host_props;
    e.launch_init = launch_init;
    e.launch_iter = launch_iter;
    e.launch_jac0 = launch_jac0;
    e.launch_init_shared = launch_init_shared;
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
