// EXPERIMENTAL (off unless FMC_SHARED_FIRST_TRIAL=1; written after the round-1 GPU budget was
// spent, verified through its host image only -- tests/test_shared_first_trial.py).
//
// In the first nl2sno call every fit starts at params_start, so at its FIRST trial point the
// previous point x0 is params_start for every fit, and the J(x0)^T r(x) part of that pass
// (the secant update's `lky`, toms573.f90:1268-1293) can be formed from the Jacobian rows at
// params_start that jac0_kernel computes once per launch: lky_k = sum_i rowJ[i][1+k] r_i(x),
// 2 FP64 operations per parameter and data point instead of a second model evaluation with
// its derivatives (39 % of the FP64 work of a trial pass on C2).  Because fits in the refill
// kernel are at different iterations, that saving is only available while all fits are in step:
// trial1_kernel is a uniform kernel (static assignment, no refill) that re-derives each fit's
// first trial point from the initial sums (the solver step is cheap and deterministic), makes
// the pass there, and stores the p x p sums; iterate_kernel<..., TRIAL1 = true> re-derives the
// same point and consumes the stored sums instead of making the pass.
#pragma once
#include "bootstrap_kernel.cuh"

namespace fmc {

FMC_HD inline int trial1_len(int p) { return 2 + 2 * p + p * (p + 1) / 2; }

template <class M, bool INJECT>
__global__ void __launch_bounds__(FMC_THREADS, 1) trial1_kernel(const KernelArgs a, const double* __restrict__ rowJ,
                                                                double* __restrict__ out) {
  constexpr int P = M::NPARAM;
  constexpr int PP = P * (P + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RowData* rows = reinterpret_cast<RowData*>(smem_raw);
  double* sJ = reinterpret_cast<double*>(smem_raw + sizeof(RowData) * a.n_pad);
  {
    const double2* src = reinterpret_cast<const double2*>(a.rows);
    double2* dst = reinterpret_cast<double2*>(rows);
    for (int i = threadIdx.x; i < 2 * a.n_pad; i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < (1 + P) * a.n_pad; i += blockDim.x) sJ[i] = rowJ[i];
  }
  __syncthreads();
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.nres) return;

  // the same start as iterate_kernel's fetch() for call 0
  Nl2<P> s;
  s.reset_fit();
  s.begin(a.start);
  s.res_f = a.init[slot];
#pragma unroll
  for (int i = 0; i < P; ++i) s.res_g[i] = a.init[(1 + i) * a.cap + slot];
#pragma unroll
  for (int i = 0; i < PP; ++i) s.res_G[i] = a.init[(1 + P + i) * a.cap + slot];
  if (!s.resume()) {  // the call ends without a trial step: iterate_kernel will find the same
    out[slot] = 0.0;
    return;
  }

  PassPoint<P> pa;
  pa.set(s.x, s.hA);
  PassAcc<P> acc;
  acc.zero();
  const unsigned long long id = (unsigned long long)(a.first + slot);
  const double* zrow = INJECT ? a.z_inject + (unsigned long long)slot * a.n : nullptr;
  for (int i = 0; i < a.n_pad; i += 4) {
    float z[4];
    if (!INJECT) normals4(a.seed, id, (uint32_t)(i >> 2), z);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const RowData rd = rows[i + m];
      double zz;
      if (INJECT)
        zz = (i + m < a.n) ? zrow[i + m] : 0.0;
      else
        zz = (double)z[m];
      const double yfit = fma(zz, rd.err, rd.y);
      // residual and Jacobian row at the trial point: as row_accumulate<M, false>
      double m0, dma[P];
      row_model<M>(pa, rd.x, m0, dma, model_naux<M>::value > 0 ? a.aux + (i + m) * model_naux<M>::value : nullptr);
      const double r = (yfit - m0) * rd.rerr;
      const double t = rd.rerr * r;
      const double w2 = rd.rerr * rd.rerr;
      acc.f = fma(r, r, acc.f);
      int q = 0;
#pragma unroll
      for (int ii = 0; ii < P; ++ii) {
        acc.g[ii] = fma(dma[ii], t, acc.g[ii]);
        const double dw = dma[ii] * w2;
#pragma unroll
        for (int jj = 0; jj <= ii; ++jj) {
          acc.G[q] = fma(dw, dma[jj], acc.G[q]);
          ++q;
        }
      }
      // J(params_start)^T r(x) from the shared rows (rowJ[i][1+k] = dm_ik / err_i at params_start)
      const double* jr = sJ + (i + m) * (1 + P);
#pragma unroll
      for (int kk = 0; kk < P; ++kk) acc.lky[kk] = fma(jr[1 + kk], r, acc.lky[kk]);
    }
  }
  double f, g[P], G[PP], lky[P];
  pass_scale<P>(acc, s.hA, s.hJ, f, g, G, lky);  // hJ: the steps of the Jacobian at params_start
  out[slot] = 1.0;
  out[a.cap + slot] = f;
#pragma unroll
  for (int i = 0; i < P; ++i) out[(2 + i) * a.cap + slot] = g[i];
#pragma unroll
  for (int i = 0; i < PP; ++i) out[(2 + P + i) * a.cap + slot] = G[i];
#pragma unroll
  for (int i = 0; i < P; ++i) out[(2 + P + PP + i) * a.cap + slot] = lky[i];
}

struct Trial1Entry {
  cudaError_t (*launch_trial1)(const KernelArgs&, const double* rowJ, double* out, bool inject, size_t shmem,
                               cudaStream_t);
  cudaError_t (*launch_iter)(const KernelArgs&, bool inject, int grid, size_t shmem, cudaStream_t);
  cudaError_t (*prepare)(bool inject, size_t shmem_trial1, size_t shmem_iter, int* blocks_per_sm);
};

// forward differences, data in shared memory: the only configuration in which the shared rows exist
template <class M>
struct Trial1Binding {
  static cudaError_t launch_trial1(const KernelArgs& a, const double* rowJ, double* out, bool inject, size_t shmem,
                                   cudaStream_t st) {
    const int grid = (int)((a.nres + FMC_THREADS - 1) / FMC_THREADS);
    if (inject)
      trial1_kernel<M, true><<<grid, FMC_THREADS, shmem, st>>>(a, rowJ, out);
    else
      trial1_kernel<M, false><<<grid, FMC_THREADS, shmem, st>>>(a, rowJ, out);
    return cudaGetLastError();
  }
  static cudaError_t launch_iter(const KernelArgs& a, bool inject, int grid, size_t shmem, cudaStream_t st) {
    if (inject)
      iterate_kernel<M, true, true, JAC_FD, true><<<grid, FMC_THREADS, shmem, st>>>(a);
    else
      iterate_kernel<M, false, true, JAC_FD, true><<<grid, FMC_THREADS, shmem, st>>>(a);
    return cudaGetLastError();
  }
  static cudaError_t prepare(bool inject, size_t shmem_trial1, size_t shmem_iter, int* nb) {
    cudaError_t e;
    if (inject) {
      e = cudaFuncSetAttribute(trial1_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_trial1);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(iterate_kernel<M, true, true, JAC_FD, true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_iter);
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, iterate_kernel<M, true, true, JAC_FD, true>,
                                                           FMC_THREADS, shmem_iter);
    }
    e = cudaFuncSetAttribute(trial1_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_trial1);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(iterate_kernel<M, false, true, JAC_FD, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem_iter);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, iterate_kernel<M, false, true, JAC_FD, true>,
                                                         FMC_THREADS, shmem_iter);
  }
  static Trial1Entry entry() { return Trial1Entry{launch_trial1, launch_iter, prepare}; }
};

Trial1Entry fmc_trial1_entry(int model);  // kernels_trial1.cu

}  // namespace fmc
