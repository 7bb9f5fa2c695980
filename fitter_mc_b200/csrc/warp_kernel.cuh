// One fit per WARP: the alternative mapping BASELINE.json's north star asks to be weighed
// against one fit per thread ("one fit per thread or per warp, chosen by measurement").
//
// The 32 lanes of a warp share ONE resample.  A pass over the n data points is split between
// them -- lane l takes the four-point groups l, l+32, l+64, ... (a group is what one counter
// block of the generator serves, rng.cuh) -- and the per-lane partial sums (f, g, G, lky:
// 1 + 2p + p(p+1)/2 doubles) are combined by a butterfly of warp shuffles, which leaves the
// bitwise SAME totals in every lane.  Every lane then runs the (cheap, p x p) solver step on
// those totals redundantly: identical inputs, identical control flow, no divergence, and no
// lane ever waits for another fit.  Both nl2sno calls of fit_data and their initial passes run
// inside the one kernel; warps take the next resample from the shared counter.
//
// Compared with one fit per thread (bootstrap_kernel.cuh): 32x fewer fits in flight (so 32x
// less solver state), perfectly uniform control flow, rows read with unit stride across lanes
// (coalesced when the data do not fit in shared memory); but the solver step is executed 32
// times over and each pass pays 5 shuffle rounds per sum.  It pays when n x p is large (C4),
// not for the small fits of C1/C2 -- which is to be confirmed by measurement: this path is
// OPT-IN (fmc_set_fit_mapping) and, at the end of round 1, has been exercised only through its
// host emulation (tests/host_emul, tests/test_warp_mapping.py), not yet on a GPU.
//
// Sums are formed in a different order than in the per-thread kernels (32 partial sums, then
// a tree), so results agree with them to rounding, not bitwise.
#pragma once
#include "bootstrap_kernel.cuh"

namespace fmc {

#ifndef FMC_WARP_THREADS
#define FMC_WARP_THREADS 256
#endif

// butterfly all-reduce: afterwards every lane holds the same sum (a + b == b + a bitwise)
FMC_INLINE __device__ double warp_allsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int P>
FMC_INLINE __device__ void warp_allsum_acc(PassAcc<P>& acc, bool with_b) {
  acc.f = warp_allsum(acc.f);
#pragma unroll
  for (int i = 0; i < P; ++i) acc.g[i] = warp_allsum(acc.g[i]);
#pragma unroll
  for (int i = 0; i < P * (P + 1) / 2; ++i) acc.G[i] = warp_allsum(acc.G[i]);
  if (with_b) {
#pragma unroll
    for (int i = 0; i < P; ++i) acc.lky[i] = warp_allsum(acc.lky[i]);
  }
}

// The share of lane `lane` in one pass: groups lane, lane + 32, ... of four data points.
template <class M, bool INJECT, bool WITH_B, int JAC>
FMC_INLINE __device__ void warp_pass_loop(const RowData* __restrict__ rows, int n, int n_pad, int lane,
                                          unsigned long long seed, unsigned long long id,
                                          const double* __restrict__ zrow, const PassPoint<M::NPARAM>& pa,
                                          const PassPoint<M::NPARAM>& pb, PassAcc<M::NPARAM>& acc) {
  for (int i = 4 * lane; i < n_pad; i += 4 * 32) {
    float z[4];
    if (!INJECT) normals4(seed, id, (uint32_t)(i >> 2), z);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const RowData rd = rows[i + m];
      double zz;
      if (INJECT)
        zz = (i + m < n) ? zrow[i + m] : 0.0;
      else
        zz = (double)z[m];
      const double yfit = fma(zz, rd.err, rd.y);  // offset_data, bootstrap.f90:208
      row_accumulate<M, WITH_B, JAC>(pa, pb, rd.x, yfit, rd.rerr, acc);
    }
  }
}

template <class M, bool INJECT, bool SMEM, int JAC>
__global__ void __launch_bounds__(FMC_WARP_THREADS, 1) warp_fit_kernel(const KernelArgs a) {
  constexpr int P = M::NPARAM;
  constexpr int Q = M::NPROP;
  constexpr int NM = 2 * (P + Q);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[FMC_WARP_THREADS / 32];
  __shared__ unsigned int sh_rc[16];
  if (threadIdx.x < 16) sh_rc[threadIdx.x] = 0u;
  const RowData* rows = stage_rows<SMEM>(a, smem_raw);
  __syncthreads();
  const int lane = threadIdx.x & 31;

  Nl2<P> s;
  double msum[NM], mcmp[NM];  // lane 0 of each warp carries the warp's moment sums
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    msum[i] = 0.0;
    mcmp[i] = 0.0;
  }
  double cq[Q > 0 ? Q : 1];
  M::derived_properties(a.start, cq);
  unsigned int n_fit = 0, n_fcall = 0, n_gcall = 0, n_iter = 0, n_pass = 0;

  for (;;) {
    unsigned long long kk = 0;
    if (lane == 0) kk = atomicAdd(a.next_id, 1ULL);
    kk = __shfl_sync(0xffffffffu, kk, 0);
    if (kk >= (unsigned long long)a.nres) break;
    const long long slot = (long long)kk;
    const unsigned long long id = (unsigned long long)(a.first + slot);
    const double* zrow = INJECT ? a.z_inject + (unsigned long long)slot * a.n : nullptr;

    s.reset_fit();
    double xs[P];
#pragma unroll
    for (int i = 0; i < P; ++i) xs[i] = a.start[i];
    int cnt[10];
    for (int call = 0; call < 2; ++call) {  // fit_data: nl2sno twice (bootstrap.f90:219-225)
      s.begin(xs);
      int passes = 0;
      bool more;
      {
        // first evaluation of the call: residual + Jacobian at the starting point, D = 1
        double h1[P];
#pragma unroll
        for (int i = 0; i < P; ++i) h1[i] = JAC == JAC_ANALYTIC ? 1.0 : s.hA[i];
        PassPoint<P> pa;
        pa.set(s.x, h1);
        PassAcc<P> acc;
        acc.zero();
        warp_pass_loop<M, INJECT, false, JAC>(rows, a.n, a.n_pad, lane, a.seed, id, zrow, pa, pa, acc);
        warp_allsum_acc<P>(acc, false);
        if constexpr (JAC == JAC_ANALYTIC)
          pass_finish_analytic<P>(acc, s);
        else
          pass_finish<P>(acc, s);
        ++passes;
        more = s.resume();
      }
      while (more) {
        PassPoint<P> pa, pb;
        pa.set(s.x, s.hA);
        pb.set(s.x0, s.hJ);
        PassAcc<P> acc;
        acc.zero();
        warp_pass_loop<M, INJECT, true, JAC>(rows, a.n, a.n_pad, lane, a.seed, id, zrow, pa, pb, acc);
        warp_allsum_acc<P>(acc, true);
        if constexpr (JAC == JAC_ANALYTIC)
          pass_finish_analytic<P>(acc, s);
        else
          pass_finish<P>(acc, s);
        ++passes;
        more = s.resume();
        if (more && passes > 2 * (k::mxfcal + k::mxiter)) {  // safety cap on passes per call
          s.anomaly |= 8;
          s.rc = 9;
          more = false;
        }
      }
      cnt[5 * call + 0] = s.rc;
      cnt[5 * call + 1] = s.nfcall;
      cnt[5 * call + 2] = s.ngcall;
      cnt[5 * call + 3] = s.niter;
      cnt[5 * call + 4] = passes;
      if (s.anomaly != 0 && a.anom != nullptr && lane == 0) {  // same record format as iterate_kernel
        const unsigned long long na = atomicAdd(a.anom, 1ULL);
        if (na < (unsigned long long)ANOM_CAP) {
          a.anom[1 + 2 * na] = id;
          a.anom[2 + 2 * na] = (unsigned long long)(s.anomaly | (call << 8));
        }
      }
      s.anomaly = 0;
#pragma unroll
      for (int i = 0; i < P; ++i) xs[i] = s.x[i];
    }

    if (lane == 0) {  // every lane holds the same result; one of them records it
      double props[Q > 0 ? Q : 1];
      M::derived_properties(xs, props);  // bootstrap.f90:316
#pragma unroll
      for (int i = 0; i < P; ++i) {       // bootstrap.f90:317-318, on shifted values
        const double dv = xs[i] - a.start[i];
        kahan_add(msum[i], mcmp[i], dv);
        kahan_add(msum[P + i], mcmp[P + i], dv * dv);
      }
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        const double dv = props[i] - cq[i];
        kahan_add(msum[2 * P + i], mcmp[2 * P + i], dv);
        kahan_add(msum[2 * P + Q + i], mcmp[2 * P + Q + i], dv * dv);
      }
      ++n_fit;
      n_fcall += cnt[1] + cnt[6];
      n_gcall += cnt[2] + cnt[7];
      n_iter += cnt[3] + cnt[8];
      n_pass += cnt[4] + cnt[9];
      atomicAdd(&sh_rc[cnt[5] & 15], 1u);
      if (a.params_out)
        for (int i = 0; i < P; ++i) a.params_out[slot * P + i] = xs[i];
      if (a.props_out)
        for (int i = 0; i < Q; ++i) a.props_out[slot * Q + i] = props[i];
      if (a.counters_out)
        for (int i = 0; i < 10; ++i) a.counters_out[slot * 10 + i] = cnt[i];
    }
  }

  // ---- epilogue: block reduction of the per-warp sums (lanes other than 0 contribute zero)
  auto block_sum = [&](double v) -> double {
    v = warp_allsum(v);
    __syncthreads();
    if (lane == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
      for (int w = 0; w < FMC_WARP_THREADS / 32; ++w) t += red[w];
    return t;
  };
  for (int i = 0; i < NM; ++i) {
    const double t = block_sum(msum[i] - mcmp[i]);
    if (threadIdx.x == 0) atomicAdd(&a.acc[i], t);
  }
  const double tails[5] = {(double)n_fit, (double)n_fcall, (double)n_gcall, (double)n_iter, (double)n_pass};
  const int tail_idx[5] = {ACC_NFIT, ACC_NFCALL, ACC_NGCALL, ACC_NITER, ACC_NPASS};
  for (int i = 0; i < 5; ++i) {
    const double t = block_sum(tails[i]);
    if (threadIdx.x == 0) atomicAdd(&a.acc[NM + tail_idx[i]], t);
  }
  __syncthreads();
  if (threadIdx.x < 16 && sh_rc[threadIdx.x] != 0u)
    atomicAdd(&a.acc[NM + ACC_RC + threadIdx.x], (double)sh_rc[threadIdx.x]);
}

// Host-side binding of one model to its warp-kernel instantiations (kernels_warp.cu).
struct WarpEntry {
  cudaError_t (*launch)(const KernelArgs&, bool inject, bool smem, int jac, int grid, size_t shmem, cudaStream_t);
  cudaError_t (*occupancy)(bool inject, bool smem, int jac, size_t shmem, int* blocks_per_sm);
};

template <class M>
struct WarpBinding {
  static constexpr bool kAnalytic = can_analytic_v<M>;
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
        return cudaErrorNotSupported;
      }
    }
    if (inject) return smem ? f(std::true_type{}, std::true_type{}, JF{}) : f(std::true_type{}, std::false_type{}, JF{});
    return smem ? f(std::false_type{}, std::true_type{}, JF{}) : f(std::false_type{}, std::false_type{}, JF{});
  }
  static cudaError_t launch(const KernelArgs& a, bool inject, bool smem, int jac, int grid, size_t shmem,
                            cudaStream_t st) {
    return dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      warp_fit_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>
          <<<grid, FMC_WARP_THREADS, shmem, st>>>(a);
      return cudaGetLastError();
    });
  }
  static cudaError_t occupancy(bool inject, bool smem, int jac, size_t shmem, int* nb) {
    return dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      auto kern = warp_fit_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, kern, FMC_WARP_THREADS, shmem);
    });
  }
  static WarpEntry entry() { return WarpEntry{launch, occupancy}; }
};

WarpEntry fmc_warp_entry(int model);  // kernels_warp.cu

}  // namespace fmc
