// The bootstrap hot loop as ONE persistent CUDA kernel for sm_100a.
//
// Replaces bootstrap.f90:309-327 (the !$OMP PARALLEL DO over nrand_points with its
// REDUCTION(+) and two CRITICAL sections) and everything it calls.
//
// Mapping (DESIGN.md has the measurements behind these choices):
//  * one fit per THREAD.  A thread owns one resample at a time and runs its NL2SOL state
//    machine (nl2_solver.cuh); when the fit ends it adds the result to its private moment
//    sums and takes the next resample id from a global counter.  Lanes of a warp are thus at
//    different iterations of different fits, but all of them always execute the same thing
//    next -- one streaming pass over the n data points -- so the warp never waits for its
//    slowest fit ("convergence mask + compaction" by refill).
//  * the data points (x, y, err, 1/err) sit in shared memory, read by broadcast (every lane
//    reads the same row); per-fit state is registers + thread-local memory; nothing of length
//    n is ever stored per fit.  HBM traffic is O(outputs).
//  * the Gaussian offsets are regenerated from the counter-based generator in every pass
//    (INT32/FP32/SFU work that issues in the slots the FP64 pipe leaves free).
//  * per-thread compensated (Kahan) sums of shifted values -> warp shuffle -> block ->
//    one atomicAdd per block into the device accumulator that the caller all-reduces.
#pragma once
#include <cuda_runtime.h>

#include "fmc_common.cuh"
#include "nl2_solver.cuh"
#include "pass_eval.cuh"
#include "rng.cuh"

namespace fmc {

#ifndef FMC_THREADS
#define FMC_THREADS 128
#endif
#ifndef FMC_MIN_BLOCKS
#define FMC_MIN_BLOCKS 3
#endif

struct __align__(16) RowData {  // one data point, 32 B: two 16-byte broadcast loads
  double x, y, err, rerr;
};

// Accumulator block (all doubles so that a single all-reduce combines GPUs):
//   [0,P) sum(p-c)  [P,2P) sum((p-c)^2)  [2P,2P+Q) sum(q-cq)  [2P+Q,2P+2Q) sum((q-cq)^2)
//   then ACC_NFIT, ACC_RC+0..15 (histogram of the 2nd-pass return code), ACC_NFCALL,
//   ACC_NGCALL, ACC_NITER, ACC_NPASS.
enum { ACC_NFIT = 0, ACC_RC = 1, ACC_NFCALL = 17, ACC_NGCALL = 18, ACC_NITER = 19, ACC_NPASS = 20, ACC_NTAIL = 21 };
FMC_HD inline int acc_len(int p, int q) { return 2 * (p + q) + ACC_NTAIL; }

struct KernelArgs {
  const RowData* rows;  // [n_pad], padded rows have rerr = err = 0
  int n;                // data points
  int n_pad;            // multiple of 4
  double start[16];     // params_start: first guess of every fit and shift of the moment sums
  unsigned long long seed;
  long long first;      // global id of resample 0 of this launch
  long long nres;       // resamples in this launch
  const double* z_inject;   // INJECT only: [nres][n] offsets in units of err
  double* params_out;       // optional [nres][P]
  double* props_out;        // optional [nres][Q]
  int* counters_out;        // optional [nres][10]: per pass rc, nfcall, ngcall, niter, passes
  unsigned long long* next_id;  // work counter, zeroed before launch
  double* acc;              // accumulator block, zeroed before launch
};

FMC_INLINE __device__ void kahan_add(double& s, double& c, double v) {
  const double y = v - c;
  const double t = s + y;
  c = (t - s) - y;
  s = t;
}

template <class M, bool INJECT, bool WITH_B>
FMC_INLINE __device__ void pass_loop(const RowData* __restrict__ rows, int n, int n_pad,
                                     unsigned long long seed, unsigned long long id,
                                     const double* __restrict__ zrow,
                                     const PassPoint<M::NPARAM>& pa, const PassPoint<M::NPARAM>& pb,
                                     PassAcc<M::NPARAM>& acc) {
  for (int i = 0; i < n_pad; i += 4) {
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
      row_accumulate<M, WITH_B>(pa, pb, rd.x, yfit, rd.rerr, acc);
    }
  }
}

template <class M, bool INJECT, bool SMEM>
__global__ void __launch_bounds__(FMC_THREADS, FMC_MIN_BLOCKS) bootstrap_fit_kernel(const KernelArgs a) {
  constexpr int P = M::NPARAM;
  constexpr int Q = M::NPROP;
  constexpr int NM = 2 * (P + Q);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[FMC_THREADS / 32];
  __shared__ unsigned int sh_rc[16];

  const RowData* rows;
  if (SMEM) {
    RowData* srows = reinterpret_cast<RowData*>(smem_raw);
    const double2* src = reinterpret_cast<const double2*>(a.rows);
    double2* dst = reinterpret_cast<double2*>(srows);
    for (int i = threadIdx.x; i < 2 * a.n_pad; i += blockDim.x) dst[i] = src[i];
    rows = srows;
  } else {
    rows = a.rows;
  }
  if (threadIdx.x < 16) sh_rc[threadIdx.x] = 0u;
  __syncthreads();

  Nl2<P> s;
  double msum[NM], mcmp[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    msum[i] = 0.0;
    mcmp[i] = 0.0;
  }
  double cq[Q > 0 ? Q : 1];
  M::derived_properties(a.start, cq);
  unsigned int n_fit = 0, n_fcall = 0, n_gcall = 0, n_iter = 0, n_pass = 0;
  int cnt[10];
  int pass_no = 0, fit_passes = 0;
  unsigned long long slot = 0;
  bool active = false;

  auto fetch = [&]() {
    const unsigned long long kk = atomicAdd(a.next_id, 1ULL);
    if (kk < (unsigned long long)a.nres) {
      slot = kk;
      s.reset_fit();
      s.begin(a.start);
      pass_no = 0;
      fit_passes = 0;
      active = true;
    } else {
      active = false;
    }
  };
  fetch();

  while (__any_sync(0xffffffffu, active)) {
    const bool anyb = __any_sync(0xffffffffu, active && s.need_b);
    if (active) {
      PassPoint<P> pa, pb;
      pa.set(s.x, s.hA);
      pb.set(s.x0, s.hJ);
      PassAcc<P> acc;
      acc.zero();
      const unsigned long long id = (unsigned long long)a.first + slot;
      const double* zrow = INJECT ? a.z_inject + slot * (unsigned long long)a.n : nullptr;
      if (anyb)
        pass_loop<M, INJECT, true>(rows, a.n, a.n_pad, a.seed, id, zrow, pa, pb, acc);
      else
        pass_loop<M, INJECT, false>(rows, a.n, a.n_pad, a.seed, id, zrow, pa, pb, acc);
      pass_finish<P>(acc, s);
      ++fit_passes;
      ++n_pass;

      if (!s.resume()) {
        n_fcall += s.nfcall;
        n_gcall += s.ngcall;
        n_iter += s.niter;
        cnt[5 * pass_no + 0] = s.rc;
        cnt[5 * pass_no + 1] = s.nfcall;
        cnt[5 * pass_no + 2] = s.ngcall;
        cnt[5 * pass_no + 3] = s.niter;
        cnt[5 * pass_no + 4] = fit_passes;
        fit_passes = 0;
        if (pass_no == 0) {
          // fit_data calls nl2sno a second time from the first answer (bootstrap.f90:223-225)
          pass_no = 1;
          s.begin(s.x);
        } else {
          double props[Q > 0 ? Q : 1];
          M::derived_properties(s.x, props);  // bootstrap.f90:316
#pragma unroll
          for (int i = 0; i < P; ++i) {       // bootstrap.f90:317-318, on shifted values
            const double dv = s.x[i] - a.start[i];
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
          atomicAdd(&sh_rc[s.rc & 15], 1u);
          if (a.params_out)
            for (int i = 0; i < P; ++i) a.params_out[slot * P + i] = s.x[i];
          if (a.props_out)
            for (int i = 0; i < Q; ++i) a.props_out[slot * Q + i] = props[i];
          if (a.counters_out)
            for (int i = 0; i < 10; ++i) a.counters_out[slot * 10 + i] = cnt[i];
          fetch();
        }
      }
    }
  }

  // ---- epilogue: warp shuffle + block reduction of the moment sums, one atomic per block
  auto block_sum = [&](double v) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
      for (int w = 0; w < FMC_THREADS / 32; ++w) t += red[w];
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

}  // namespace fmc
