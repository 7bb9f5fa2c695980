// One fit per WARP: the mapping for fits with many parameters and many data points (BASELINE.json
// config C4: p = 10, n = 10^4), weighed against one fit per thread by measurement
// (tools/gpu_mapping_bench.py).
//
// Why one fit per thread runs out of road at large p: the sums of a pass are 1 + 2p + p(p+1)/2
// doubles (76 at p = 10 = 152 registers), the two evaluation points with their step-dependent
// constants another ~90 registers, and the solver state 4.4 KB per fit -- the row loop of the
// per-thread kernel compiles to 255 registers with the sums themselves spilled (78 local-memory
// operations per data point), and 38 000 fits in flight keep 160 MB of solver state moving.
//
// Here the 32 lanes of a warp share ONE resample:
//  * rows are dealt to the lanes in groups of four (a group is what one counter block of the
//    generator serves, rng.cuh): lane l takes groups l, l + 32, l + 64, ...  The data are stored
//    once more in that order, one array per column ("lane-major": element s = (4 j + m) 32 + l is
//    row 4 (32 j + l) + m), so that every load of a warp is one contiguous 256-byte segment.  x and
//    1/err are staged in shared memory when they fit (all four columns when those fit too).
//  * a trial pass is made in TWO SWEEPS over the warp's rows.  Sweep A evaluates the model and its
//    differences at the trial point x: f, g = J^T r, G = J^T J (1 + p + p(p+1)/2 sums) and leaves
//    t_i = r_i / err_i in a per-warp scratch row (L2 resident, one coalesced store).  Sweep B
//    evaluates the differences at the previous point x0 and forms lky = J(x0)^T r(x) from the
//    stored t_i (p sums).  Neither sweep holds the sums and the point-dependent constants of the
//    other, which is what keeps the loops inside the register file; the FP64 work is the same as
//    in one sweep.
//  * the per-lane partial sums are combined by a butterfly of warp shuffles; lane 0 hands the totals
//    to the solver, whose state (4.4 KB at p = 10) lives in SHARED memory, one copy per warp, and
//    runs the p x p solver step alone (an instruction issued for one lane costs what it costs for
//    32); the other lanes then read the next evaluation point from shared memory by broadcast.  No
//    divergence between fits, and no lane ever waits for another fit.  Both nl2sno calls of
//    fit_data and their initial passes run inside the one kernel; warps take the next resample
//    from a shared counter.
//
// The deviates are the same function of (seed, resample, data point) as in the per-thread
// kernels, so both mappings fit the same resampled data sets; sums are formed in a different
// order (32 partial sums, then a tree), so results agree to rounding, not bitwise.
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

// The columns of the data in lane-major order (see above); pointers are generic: shared memory
// when the column was staged, global otherwise.
struct WarpCols {
  const double* x;
  const double* y;
  const double* err;
  const double* rerr;
  const double* aux;  // [NAUX][n_w] per-row auxiliary values of the model (pass_eval.cuh)
};

// the auxiliary values of row s, gathered from their columns
template <class M>
FMC_INLINE __device__ void warp_row_aux(const WarpCols& c, int n_w, int s, double (&ra)[model_naux<M>::value > 0 ? model_naux<M>::value : 1]) {
  constexpr int NAUX = model_naux<M>::value;
#pragma unroll
  for (int k = 0; k < NAUX; ++k) ra[k] = c.aux[(size_t)k * n_w + s];
}

// Sums of sweep A (f, g, G) and of sweep B (lky), kept apart so that each loop only carries its own.
template <int P>
struct SweepA {
  static constexpr int PP = P * (P + 1) / 2;
  double f, g[P], G[PP];
  FMC_INLINE __device__ void zero() {
    f = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) g[i] = 0.0;
#pragma unroll
    for (int i = 0; i < PP; ++i) G[i] = 0.0;
  }
  FMC_INLINE __device__ void allsum() {
    f = warp_allsum(f);
#pragma unroll
    for (int i = 0; i < P; ++i) g[i] = warp_allsum(g[i]);
#pragma unroll
    for (int i = 0; i < PP; ++i) G[i] = warp_allsum(G[i]);
  }
};

// Sweep A over the rows of lane `lane`: residual, Jacobian row and the sums at the point `pa`.
// STORE_T: leave t_i = r_i / err_i in tbuf (lane-major) for sweep B.
template <class M, bool INJECT, bool STORE_T, int JAC>
FMC_INLINE __device__ void sweep_a(const WarpCols& c, int n, int n_w, int lane, unsigned long long seed,
                                   unsigned long long id, const double* __restrict__ zrow,
                                   const PassPoint<M::NPARAM>& pa, double* __restrict__ tbuf,
                                   SweepA<M::NPARAM>& acc) {
  constexpr int P = M::NPARAM;
  constexpr int NAUX = model_naux<M>::value;
  const int ngroup = n_w >> 7;  // groups of 4 rows per lane
  for (int j = 0; j < ngroup; ++j) {
    const int blk = 32 * j + lane;
    float z[4];
    if (!INJECT) normals4(seed, id, (uint32_t)blk, z);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      // rows beyond n (padding) are skipped, not given a zero weight: a model value that overflows
      // there would turn the sums into NaN (inf x 0) instead of an ordinary "too big" residual
      if (4 * blk + m >= n) continue;
      const int s = (4 * j + m) * 32 + lane;
      const double xi = c.x[s], yi = c.y[s], ei = c.err[s], wi = c.rerr[s];
      const double zz = INJECT ? zrow[4 * blk + m] : (double)z[m];
      const double yfit = fma(zz, ei, yi);  // offset_data, bootstrap.f90:208
      double m0, dma[P];
      double ra[NAUX > 0 ? NAUX : 1];
      warp_row_aux<M>(c, n_w, s, ra);
      row_model<M, JAC>(pa, xi, m0, dma, ra);
      const double r = (yfit - m0) * wi;  // madr, bootstrap.f90:247
      const double t = wi * r;
      const double w2 = wi * wi;
      if (STORE_T) tbuf[s] = t;
      acc.f = fma(r, r, acc.f);
      int q = 0;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        acc.g[i] = fma(dma[i], t, acc.g[i]);
        const double dw = dma[i] * w2;
#pragma unroll
        for (int jj = 0; jj <= i; ++jj) {
          acc.G[q] = fma(dw, dma[jj], acc.G[q]);
          ++q;
        }
      }
    }
  }
}

// Sweep B: lky = J(pb)^T r(pa), from the t_i that sweep A left in tbuf.
template <class M, int JAC>
FMC_INLINE __device__ void sweep_b(const WarpCols& c, int n, int n_w, int lane, const PassPoint<M::NPARAM>& pb,
                                   const double* __restrict__ tbuf, double (&lky)[M::NPARAM]) {
  constexpr int P = M::NPARAM;
#pragma unroll
  for (int i = 0; i < P; ++i) lky[i] = 0.0;
#pragma unroll 4
  for (int s = lane; s < n_w; s += 32) {
    if (4 * (32 * (s >> 7) + lane) + ((s >> 5) & 3) >= n) continue;  // padding row (see sweep_a)
    const double xi = c.x[s];
    const double t = tbuf[s];
    double mb, dmb[P];
    double ra[model_naux<M>::value > 0 ? model_naux<M>::value : 1];
    warp_row_aux<M>(c, n_w, s, ra);
    row_model<M, JAC>(pb, xi, mb, dmb, ra);
#pragma unroll
    for (int i = 0; i < P; ++i) lky[i] = fma(dmb[i], t, lky[i]);
  }
#pragma unroll
  for (int i = 0; i < P; ++i) lky[i] = warp_allsum(lky[i]);
}

// Shared memory of one warp's solver state (16-byte granules).
template <int P>
FMC_HD constexpr size_t warp_solver_bytes() {
  return (sizeof(Nl2<P>) + 15) / 16 * 16;
}

// Arguments of the warp kernel beyond KernelArgs.
struct WarpArgs {
  const double* cols;   // [4 + NAUX][n_w] lane-major x, y, err, 1/err, auxiliary values (global)
  double* tscratch;     // [warps in the grid][n_w]
  int n_w;              // rows padded to a multiple of 128 (padding: err = 1/err = 0)
  int stage;            // staged in shared memory: 0 nothing, 2 x, 1/err and the auxiliary columns, 4 all
};

template <class M, bool INJECT, int JAC>
__global__ void __launch_bounds__(FMC_WARP_THREADS, 1) warp_fit_kernel(const KernelArgs a, const WarpArgs wa) {
  constexpr int P = M::NPARAM;
  constexpr int Q = M::NPROP;
  constexpr int PP = P * (P + 1) / 2;
  constexpr int NM = 2 * (P + Q);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[FMC_WARP_THREADS / 32];
  __shared__ unsigned int sh_rc[16];
  if (threadIdx.x < 16) sh_rc[threadIdx.x] = 0u;
  const int n_w = wa.n_w;
  WarpCols c;
  {
    // staged columns: x and 1/err first (both sweeps / the sums need them most), then y and err
    double* sm = reinterpret_cast<double*>(smem_raw + (FMC_WARP_THREADS / 32) * warp_solver_bytes<P>());
    const double* gx = wa.cols;
    const double* gy = wa.cols + n_w;
    const double* ge = wa.cols + 2 * (size_t)n_w;
    const double* gw = wa.cols + 3 * (size_t)n_w;
    const double* ga = wa.cols + 4 * (size_t)n_w;
    constexpr int NAUX = model_naux<M>::value;
    if (wa.stage >= 2) {
      for (int i = threadIdx.x; i < n_w; i += blockDim.x) {
        sm[i] = gx[i];
        sm[n_w + i] = gw[i];
      }
      for (int i = threadIdx.x; i < NAUX * n_w; i += blockDim.x) sm[2 * n_w + i] = ga[i];
      c.x = sm;
      c.rerr = sm + n_w;
      c.aux = sm + 2 * n_w;
    } else {
      c.x = gx;
      c.rerr = gw;
      c.aux = ga;
    }
    if (wa.stage >= 4) {
      double* sm2 = sm + (size_t)(2 + NAUX) * n_w;
      for (int i = threadIdx.x; i < n_w; i += blockDim.x) {
        sm2[i] = gy[i];
        sm2[n_w + i] = ge[i];
      }
      c.y = sm2;
      c.err = sm2 + n_w;
    } else {
      c.y = gy;
      c.err = ge;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double* tbuf = wa.tscratch + ((size_t)blockIdx.x * (FMC_WARP_THREADS / 32) + warp) * (size_t)n_w;

  // The solver state of the warp's fit lives in SHARED memory and lane 0 alone runs the solver step
  // between passes (an instruction issued for one lane costs what it costs for 32): one copy per fit
  // instead of 32 thread-local ones (4.4 KB each at p = 10 -- 1.1 MB per SM, i.e. every access an L2
  // round trip), and the other lanes read the next evaluation point from it by broadcast.
  Nl2<P>& s = *reinterpret_cast<Nl2<P>*>(smem_raw + (size_t)warp * warp_solver_bytes<P>());
  double msum[NM], mcmp[NM];  // lane 0 of each warp carries the warp's moment sums
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    msum[i] = 0.0;
    mcmp[i] = 0.0;
  }
  double cq[Q > 0 ? Q : 1];
  M::derived_properties(a.start, cq);
  unsigned int n_fit = 0, n_fcall = 0, n_gcall = 0, n_iter = 0, n_pass = 0;

  // one pass at s.x (and, WITH_B, the J(x0)^T r(x) part); lane 0 hands the totals to the solver
  auto pass = [&](auto with_b, unsigned long long id, const double* zrow) {
    constexpr bool WITH_B = decltype(with_b)::value;
    double h1[P];
    {
      PassPoint<P> pa;
#pragma unroll
      for (int i = 0; i < P; ++i) h1[i] = JAC == JAC_ANALYTIC ? 1.0 : s.hA[i];
      pa.set(s.x, h1);
      SweepA<P> acc;
      acc.zero();
      sweep_a<M, INJECT, WITH_B, JAC>(c, a.n, n_w, lane, a.seed, id, zrow, pa, tbuf, acc);
      acc.allsum();
      if (lane == 0) {
        s.res_f = acc.f;
#pragma unroll
        for (int i = 0; i < P; ++i) s.res_g[i] = JAC == JAC_ANALYTIC ? acc.g[i] : acc.g[i] * (1.0 / h1[i]);
        int q = 0;
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
          for (int jj = 0; jj <= i; ++jj) {
            s.res_G[q] = JAC == JAC_ANALYTIC ? acc.G[q] : acc.G[q] * ((1.0 / h1[i]) * (1.0 / h1[jj]));
            ++q;
          }
      }
    }
    if constexpr (WITH_B) {
      double hb[P];
#pragma unroll
      for (int i = 0; i < P; ++i) hb[i] = JAC == JAC_ANALYTIC ? 1.0 : s.hJ[i];
      PassPoint<P> pb;
      pb.set(s.x0, hb);
      double lky[P];
      sweep_b<M, JAC>(c, a.n, n_w, lane, pb, tbuf, lky);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < P; ++i) s.res_lky[i] = JAC == JAC_ANALYTIC ? lky[i] : lky[i] * (1.0 / hb[i]);
      }
    }
  };
  // the solver step: lane 0 consumes the pass; everybody learns whether another pass is wanted
  auto step = [&](int& passes) -> bool {
    int more = 0;
    if (lane == 0) {
      more = s.resume() ? 1 : 0;
      if (more && passes > 2 * (k::mxfcal + k::mxiter)) {  // safety cap on passes per call
        s.anomaly |= 8;
        s.rc = 9;
        more = 0;
      }
    }
    more = __shfl_sync(0xffffffffu, more, 0);  // (also orders lane 0's writes before the others' reads)
    return more != 0;
  };

  for (;;) {
    unsigned long long kk = 0;
    if (lane == 0) kk = atomicAdd(a.next_id, 1ULL);
    kk = __shfl_sync(0xffffffffu, kk, 0);
    if (kk >= (unsigned long long)a.nres) break;
    const long long slot = (long long)kk;
    const unsigned long long id = (unsigned long long)(a.first + slot);
    const double* zrow = INJECT ? a.z_inject + (unsigned long long)slot * a.n : nullptr;

    int cnt[10];  // (lane 0's copy is the one that is used)
    for (int call = 0; call < 2; ++call) {  // fit_data: nl2sno twice (bootstrap.f90:219-225)
      if (lane == 0) {
        if (call == 0) {
          s.reset_fit();
          s.begin(a.start);
        } else {
          double xs[P];
#pragma unroll
          for (int i = 0; i < P; ++i) xs[i] = s.x[i];
          s.begin(xs);
        }
      }
      __syncwarp();
      int passes = 0;
      bool more = true;
      while (more) {  // (one call site of the solver step: its code exists once)
        if (passes == 0)
          pass(std::false_type{}, id, zrow);  // first evaluation of the call: residual + Jacobian at the start, D = 1
        else
          pass(std::true_type{}, id, zrow);
        ++passes;
        more = step(passes);
      }
      if (lane == 0) {
        cnt[5 * call + 0] = s.rc;
        cnt[5 * call + 1] = s.nfcall;
        cnt[5 * call + 2] = s.ngcall;
        cnt[5 * call + 3] = s.niter;
        cnt[5 * call + 4] = passes;
        if (s.anomaly != 0 && a.anom != nullptr) {  // same record format as iterate_kernel
          const unsigned long long na = atomicAdd(a.anom, 1ULL);
          if (na < (unsigned long long)ANOM_CAP) {
            a.anom[1 + 2 * na] = id;
            a.anom[2 + 2 * na] = (unsigned long long)(s.anomaly | (call << 8));
          }
        }
        s.anomaly = 0;
      }
    }

    if (lane == 0) {  // one lane records the fit
      double xs[P];
#pragma unroll
      for (int i = 0; i < P; ++i) xs[i] = s.x[i];
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
    __syncwarp();
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
#pragma unroll 1
  for (int i = 0; i < NM; ++i) {
    const double t = block_sum(msum[i] - mcmp[i]);
    if (threadIdx.x == 0) atomicAdd(&a.acc[i], t);
  }
  const double tails[5] = {(double)n_fit, (double)n_fcall, (double)n_gcall, (double)n_iter, (double)n_pass};
  const int tail_idx[5] = {ACC_NFIT, ACC_NFCALL, ACC_NGCALL, ACC_NITER, ACC_NPASS};
#pragma unroll 1
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
  cudaError_t (*launch)(const KernelArgs&, const WarpArgs&, bool inject, int jac, int grid, size_t shmem,
                        cudaStream_t);
  cudaError_t (*occupancy)(bool inject, int jac, size_t shmem, int* blocks_per_sm);
  size_t solver_smem;  // shared memory for the per-warp solver states of one block
};

template <class M>
struct WarpBinding {
  static constexpr bool kAnalytic = can_analytic_v<M>;
  template <class F>
  static cudaError_t dispatch(bool inject, int jac, F&& f) {
    using JF = std::integral_constant<int, JAC_FD>;
    using JA = std::integral_constant<int, JAC_ANALYTIC>;
    if (jac == JAC_ANALYTIC) {
      if constexpr (kAnalytic) {
        return inject ? f(std::true_type{}, JA{}) : f(std::false_type{}, JA{});
      } else {
        return cudaErrorNotSupported;
      }
    }
    return inject ? f(std::true_type{}, JF{}) : f(std::false_type{}, JF{});
  }
  static cudaError_t launch(const KernelArgs& a, const WarpArgs& wa, bool inject, int jac, int grid, size_t shmem,
                            cudaStream_t st) {
    return dispatch(inject, jac, [&](auto I, auto J) {
      warp_fit_kernel<M, decltype(I)::value, decltype(J)::value><<<grid, FMC_WARP_THREADS, shmem, st>>>(a, wa);
      return cudaGetLastError();
    });
  }
  static cudaError_t occupancy(bool inject, int jac, size_t shmem, int* nb) {
    return dispatch(inject, jac, [&](auto I, auto J) {
      auto kern = warp_fit_kernel<M, decltype(I)::value, decltype(J)::value>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, kern, FMC_WARP_THREADS, shmem);
    });
  }
  static WarpEntry entry() {
    return WarpEntry{launch, occupancy, (FMC_WARP_THREADS / 32) * warp_solver_bytes<M::NPARAM>()};
  }
};

// kernels_warp.cu, one object per model
WarpEntry fmc_warp_entry_model0();
WarpEntry fmc_warp_entry_model1();
WarpEntry fmc_warp_entry_model2();
WarpEntry fmc_warp_entry_model3();
inline WarpEntry fmc_warp_entry(int model) {
  switch (model) {
    case 0: return fmc_warp_entry_model0();
    case 1: return fmc_warp_entry_model1();
    case 2: return fmc_warp_entry_model2();
    case 3: return fmc_warp_entry_model3();
  }
  return WarpEntry{nullptr, nullptr, 0};
}

}  // namespace fmc
