// The bootstrap hot loop as a short pipeline of CUDA kernels for sm_100a.
//
// Replaces bootstrap.f90:309-327 (the !$OMP PARALLEL DO over nrand_points with its
// REDUCTION(+) and two CRITICAL sections) and everything it calls.
//
// Mapping (DESIGN.md has the measurements behind these choices):
//  * one fit per THREAD; nothing of length n is ever stored per fit.  The data points
//    (x, y, err, 1/err) sit in shared memory and are read by broadcast (every lane reads the
//    same row); per-fit solver state is registers + thread-local memory; the Gaussian offsets
//    are regenerated from the counter-based generator in every pass (INT32/FP32/SFU work that
//    issues in the slots the FP64 pipe leaves free).  HBM traffic is O(outputs) plus ~0.4 KB
//    of pipeline state per fit.
//  * fit_data's two nl2sno calls (bootstrap.f90:219-225) become four launches per chunk of
//    resamples:   init_pass(call 0) -> iterate(call 0) -> init_pass(call 1) -> iterate(call 1).
//    init_pass is the first residual+Jacobian evaluation of a call: identical work for every
//    fit, so it runs as a perfectly uniform kernel.  iterate runs the NL2SOL state machine
//    (nl2_solver.cuh) with REFILL: a thread whose fit ends takes the next one from a shared
//    counter, so lanes of a warp are at different iterations of different fits but always
//    execute the same thing next -- one trial pass over the n data points -- and a warp never
//    waits for its slowest fit ("convergence mask + compaction").
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
#define FMC_THREADS 256
#endif
#ifndef FMC_MIN_BLOCKS
#define FMC_MIN_BLOCKS 1
#endif
#ifndef FMC_PHILOX_ROUNDS
#define FMC_PHILOX_ROUNDS 10
#endif

struct __align__(16) RowData {  // one data point, 32 B: two 16-byte broadcast loads
  double x, y, err, rerr;
};

// Accumulator block (all doubles so that a single all-reduce combines GPUs):
//   [0,P) sum(p-c)  [P,2P) sum((p-c)^2)  [2P,2P+Q) sum(q-cq)  [2P+Q,2P+2Q) sum((q-cq)^2)
//   then ACC_NFIT, ACC_RC+0..15 (histogram of the 2nd-pass return code), ACC_NFCALL,
//   ACC_NGCALL, ACC_NITER, ACC_NPASS.
enum { ACC_NFIT = 0, ACC_RC = 1, ACC_NFCALL = 17, ACC_NGCALL = 18, ACC_NITER = 19, ACC_NPASS = 20, ACC_NTAIL = 21 };
enum { ANOM_CAP = 64 };  // anomaly log: [0] count, then (id, flags) pairs
enum { HIST_MAX_QUANT = 32 };  // binned quantities: P parameters then Q properties
FMC_HD inline int acc_len(int p, int q) { return 2 * (p + q) + ACC_NTAIL; }

struct KernelArgs {
  const RowData* rows;  // [n_pad], padded rows have rerr = err = 0
  const double* aux;    // [n_pad][NAUX] per-row auxiliary values of the model (pass_eval.cuh), or nullptr
  int n;                // data points
  int n_pad;            // multiple of 4
  double start[16];     // params_start: first guess of every fit and shift of the moment sums
  unsigned long long seed;
  long long first;      // global resample id of slot 0 of this chunk
  long long nres;       // fits in this chunk
  long long cap;        // stride of the per-fit state arrays below (chunk capacity)
  int call;             // 0 = first nl2sno call of fit_data, 1 = second (bootstrap.f90:222-225)
  const double* z_inject;   // INJECT only: [nres][n] offsets in units of err (chunk-local)
  // per-fit pipeline state, structure-of-arrays [field][slot] so that lanes coalesce
  double* xbuf;   // [P]       parameters after the first nl2sno call
  double* init;   // [1+P+PP]  sums of the initial pass: f, g, G (already scaled by 1/h)
  int* istate;    // [2]       mlstgd, lsvmin LCG state carried from call 0 to call 1
  int* cnt;       // [10]      per call: rc, nfcall, ngcall, niter, passes
  double* params_out;       // optional [nres][P]   (chunk-local)
  double* props_out;        // optional [nres][Q]
  int* counters_out;        // optional [nres][10]
  unsigned long long* next_id;  // work counter of this launch, zeroed beforehand
  unsigned long long* anom;     // anomaly log (fits that hit a safety cap), zeroed per launch
  double* acc;              // accumulator block, zeroed before the first chunk
};

// Optional binned distributions of the fitted parameters and properties (replaces the
// reference's one-text-line-per-resample histogram files, bootstrap.f90:319-325): quantity j
// has `bins` equal-width bins on [lo[j], lo[j] + bins / scale[j]), stored as
// hist[j * (bins + 2) + b] with b = 0 underflow (and NaN), 1..bins, bins + 1 overflow.
struct HistArgs {
  unsigned long long* hist;
  int bins;
  int np, nq;                // parameters, properties per fit
  const double* params;      // [nres][np]  fitted parameters of one chunk
  const double* props;       // [nres][nq]
  long long nres;
  double lo[HIST_MAX_QUANT];
  double scale[HIST_MAX_QUANT];
};

// Bin index of value v: the same three FP64 operations (subtract, multiply, truncate) on host
// and device, so a host re-binning of per-resample outputs reproduces the counts exactly.
FMC_HD inline int hist_bin(double v, double lo, double scale, int nbins) {
  const double t = (v - lo) * scale;
  if (!(t >= 0.0)) return 0;
  if (t >= (double)nbins) return nbins + 1;
  return 1 + (int)t;
}

FMC_INLINE __device__ void kahan_add(double& s, double& c, double v) {
  const double y = v - c;
  const double t = s + y;
  c = (t - s) - y;
  s = t;
}

// One streaming pass of one fit over the n data points.
// Rows beyond n (the padding to a multiple of four) are SKIPPED, not multiplied by a zero weight: a
// model value that overflows at the padding abscissa would otherwise turn the sums into NaN
// (inf x 0) where the reference sees an ordinary "too big" residual (toms573.f90:949-952).
template <class M, bool INJECT, bool WITH_B, int JAC = JAC_FD>
FMC_INLINE __device__ void pass_loop(const RowData* __restrict__ rows, const double* __restrict__ aux, int n,
                                     int n_pad, unsigned long long seed, unsigned long long id,
                                     const double* __restrict__ zrow,
                                     const PassPoint<M::NPARAM>& pa, const PassPoint<M::NPARAM>& pb,
                                     PassAcc<M::NPARAM>& acc) {
  constexpr int NAUX = model_naux<M>::value;
  auto row = [&](int r, float zf) {
    const RowData rd = rows[r];
    const double zz = INJECT ? zrow[r] : (double)zf;
    const double yfit = fma(zz, rd.err, rd.y);  // offset_data, bootstrap.f90:208
    const double* ra = NAUX > 0 ? aux + r * NAUX : nullptr;
    if (NAUX > 0) __builtin_assume(ra != nullptr);
    row_accumulate<M, WITH_B, JAC>(pa, pb, rd.x, yfit, rd.rerr, acc, ra);
  };
  // four data points per trip: one counter block of the generator yields their four deviates
  const int n4 = n & ~3;
  for (int i = 0; i < n4; i += 4) {
    float z[4];
    if (!INJECT) normals4(seed, id, (uint32_t)(i >> 2), z);
#pragma unroll
    for (int m = 0; m < 4; ++m) row(i + m, z[m]);
  }
  if (n4 < n) {  // the last, partial group
    float z[4];
    if (!INJECT) normals4(seed, id, (uint32_t)(n4 >> 2), z);
#pragma unroll
    for (int m = 0; m < 3; ++m)
      if (n4 + m < n) row(n4 + m, z[m]);
  }
  (void)n_pad;
}

// Rows (and, NAUX > 0, the per-row auxiliary values behind them) into shared memory.
template <bool SMEM, int NAUX = 0>
FMC_INLINE __device__ const RowData* stage_rows(const KernelArgs& a, unsigned char* smem_raw,
                                                const double** aux_out = nullptr) {
  if (!SMEM) {
    if (aux_out) *aux_out = a.aux;
    return a.rows;
  }
  RowData* srows = reinterpret_cast<RowData*>(smem_raw);
  const double2* src = reinterpret_cast<const double2*>(a.rows);
  double2* dst = reinterpret_cast<double2*>(srows);
  for (int i = threadIdx.x; i < 2 * a.n_pad; i += blockDim.x) dst[i] = src[i];
  if (NAUX > 0) {
    double* saux = reinterpret_cast<double*>(smem_raw + sizeof(RowData) * a.n_pad);
    for (int i = threadIdx.x; i < NAUX * a.n_pad; i += blockDim.x) saux[i] = a.aux[i];
    if (aux_out) *aux_out = saux;
  }
  __syncthreads();
  return srows;
}

// ---------------------------------------------------------------------------------------
// Stage A: the first pass of an nl2sno call (residual + Jacobian at the starting point,
// toms573.f90:695-733 with D = 1).  Every fit does exactly the same amount of work here and
// none needs the J(x0)^T r(x) part, so this stage runs as its own, perfectly uniform kernel:
// one fit per thread, static assignment, no solver code.
template <class M, bool INJECT, bool SMEM, int JAC = JAC_FD>
__global__ void __launch_bounds__(FMC_THREADS) init_pass_kernel(const KernelArgs a) {
  constexpr int P = M::NPARAM;
  constexpr int PP = P * (P + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const double* aux = nullptr;
  const RowData* rows = stage_rows<SMEM, model_naux<M>::value>(a, smem_raw, &aux);
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.nres) return;
  double x[P], h[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    x[i] = a.call == 0 ? a.start[i] : a.xbuf[i * a.cap + slot];
    h[i] = JAC == JAC_ANALYTIC ? 1.0 : k::dltfdj * fmax(fabs(x[i]), 1.0);  // toms573.f90:716 with D = 1
  }
  PassPoint<P> pa;
  pa.set(x, h);
  PassAcc<P> acc;
  acc.zero();
  const unsigned long long id = (unsigned long long)(a.first + slot);
  const double* zrow = INJECT ? a.z_inject + (unsigned long long)slot * a.n : nullptr;
  pass_loop<M, INJECT, false, JAC>(rows, aux, a.n, a.n_pad, a.seed, id, zrow, pa, pa, acc);
  double f, g[P], G[PP], lky[P];
  pass_scale<P>(acc, h, h, f, g, G, lky);
  a.init[slot] = f;
#pragma unroll
  for (int i = 0; i < P; ++i) a.init[(1 + i) * a.cap + slot] = g[i];
#pragma unroll
  for (int i = 0; i < PP; ++i) a.init[(1 + P + i) * a.cap + slot] = G[i];
}

// ---------------------------------------------------------------------------------------
// Stage A for the FIRST nl2sno call: every fit starts from the same point params_start with
// the same steps h, and the model differences dm_k do not involve the resampled y' at all
// (pass_eval.cuh), so the Jacobian rows J(x_start) -- and hence G = J^T J -- are the same for
// every resample.  jac0_kernel evaluates them once per launch; init_shared_kernel then needs,
// per fit and data point, only r = (y' - m0) w:  f += r^2,  g_k += (dm_k w) r  (8 FP64
// operations instead of a model evaluation).
//   rowJ[i][0] = m0_i,  rowJ[i][1+k] = dm_ik / err_i     (unscaled by 1/h_k)
//   One-off O(n p) work: a single block of FMC_JAC0_THREADS threads (the bound keeps the launch
//   valid whatever registers the plug-in's row_model needs).
#define FMC_JAC0_THREADS 256
template <class M>
__global__ void __launch_bounds__(FMC_JAC0_THREADS) jac0_kernel(const KernelArgs a, double* __restrict__ rowJ,
                                                                double* __restrict__ G0) {
  static_assert(M::NPARAM * (M::NPARAM + 1) / 2 <= FMC_JAC0_THREADS, "one thread per element of G0");
  constexpr int P = M::NPARAM;
  constexpr int PP = P * (P + 1) / 2;
  double x[P], h[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    x[i] = a.start[i];
    h[i] = k::dltfdj * fmax(fabs(x[i]), 1.0);
  }
  PassPoint<P> pt;
  pt.set(x, h);
  for (int i = threadIdx.x; i < a.n_pad; i += blockDim.x) {
    const RowData rd = a.rows[i];
    double m0, dm[P];
    constexpr int NAUX = model_naux<M>::value;
    row_model<M>(pt, rd.x, m0, dm, NAUX > 0 ? a.aux + i * NAUX : nullptr);
    rowJ[i * (1 + P)] = m0;
#pragma unroll
    for (int kk = 0; kk < P; ++kk) rowJ[i * (1 + P) + 1 + kk] = dm[kk] * rd.rerr;
  }
  __syncthreads();
  if (threadIdx.x < PP) {  // thread m sums element m of the lower triangle, rows in order
    int ki = 0, kj = 0, m = 0;
    for (int i = 0; i < P; ++i)
      for (int j = 0; j <= i; ++j) {
        if (m == (int)threadIdx.x) {
          ki = i;
          kj = j;
        }
        ++m;
      }
    double sum = 0.0;
    for (int i = 0; i < a.n_pad; ++i) sum = fma(rowJ[i * (1 + P) + 1 + ki], rowJ[i * (1 + P) + 1 + kj], sum);
    G0[threadIdx.x] = sum * ((1.0 / h[ki]) * (1.0 / h[kj]));
  }
}

template <class M, bool INJECT>
__global__ void __launch_bounds__(FMC_THREADS) init_shared_kernel(const KernelArgs a,
                                                                  const double* __restrict__ rowJ,
                                                                  const double* __restrict__ G0) {
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
  const unsigned long long id = (unsigned long long)(a.first + slot);
  const double* zrow = INJECT ? a.z_inject + (unsigned long long)slot * a.n : nullptr;
  double f = 0.0, g[P];
#pragma unroll
  for (int i = 0; i < P; ++i) g[i] = 0.0;
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
      const double* jr = sJ + (i + m) * (1 + P);
      const double r = (yfit - jr[0]) * rd.rerr;
      f = fma(r, r, f);
#pragma unroll
      for (int kk = 0; kk < P; ++kk) g[kk] = fma(jr[1 + kk], r, g[kk]);
    }
  }
  a.init[slot] = f;
#pragma unroll
  for (int i = 0; i < P; ++i)
    a.init[(1 + i) * a.cap + slot] = g[i] * (1.0 / (k::dltfdj * fmax(fabs(a.start[i]), 1.0)));
#pragma unroll
  for (int i = 0; i < PP; ++i) a.init[(1 + P + i) * a.cap + slot] = G0[i];
}

// ---------------------------------------------------------------------------------------
// Stage B: the NL2SOL iterations of one nl2sno call, one fit per thread with refill: a thread
// that finishes takes the next fit from a shared counter, so lanes of a warp are at different
// iterations of different fits but always execute the same thing next -- one trial pass
// (residual + speculative Jacobian at x, Jacobian at x0 for the secant update).
template <class M, bool INJECT, bool SMEM, int JAC = JAC_FD>
__global__ void __launch_bounds__(FMC_THREADS, FMC_MIN_BLOCKS) iterate_kernel(const KernelArgs a) {
  constexpr int P = M::NPARAM;
  constexpr int Q = M::NPROP;
  constexpr int PP = P * (P + 1) / 2;
  constexpr int NM = 2 * (P + Q);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[FMC_THREADS / 32];
  __shared__ unsigned int sh_rc[16];
  if (threadIdx.x < 16) sh_rc[threadIdx.x] = 0u;
  const double* aux = nullptr;
  const RowData* rows = stage_rows<SMEM, model_naux<M>::value>(a, smem_raw, &aux);
  __syncthreads();

  Nl2<P> s;
  s.anomaly = 0;
  s.spec = 1;  // a trial pass of this kernel always evaluates the Jacobian too: lanes of a warp make their passes together
  double msum[NM], mcmp[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    msum[i] = 0.0;
    mcmp[i] = 0.0;
  }
  double cq[Q > 0 ? Q : 1];
  M::derived_properties(a.start, cq);
  unsigned int n_fit = 0, n_fcall = 0, n_gcall = 0, n_iter = 0, n_pass = 0;
  int fit_passes = 0;
  long long slot = 0;
  bool active = false;

  // end of an nl2sno call: hand the state to the next stage, or (second call) to the sums
  auto finish_call = [&]() {
    if (s.anomaly != 0 && a.anom != nullptr) {
      const unsigned long long kk = atomicAdd(a.anom, 1ULL);
      if (kk < (unsigned long long)ANOM_CAP) {
        a.anom[1 + 2 * kk] = (unsigned long long)(a.first + slot);
        a.anom[2 + 2 * kk] = (unsigned long long)(s.anomaly | (a.call << 8));
      }
      s.anomaly = 0;
    }
    if (a.call == 0) {
#pragma unroll
      for (int i = 0; i < P; ++i) a.xbuf[i * a.cap + slot] = s.x[i];
      a.istate[slot] = s.mlstgd;
      a.istate[a.cap + slot] = s.lsvmin_ix;
      a.cnt[0 * a.cap + slot] = s.rc;
      a.cnt[1 * a.cap + slot] = s.nfcall;
      a.cnt[2 * a.cap + slot] = s.ngcall;
      a.cnt[3 * a.cap + slot] = s.niter;
      a.cnt[4 * a.cap + slot] = fit_passes;
      return;
    }
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
    int c0[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) c0[i] = a.cnt[i * a.cap + slot];
    ++n_fit;
    n_fcall += s.nfcall + c0[1];
    n_gcall += s.ngcall + c0[2];
    n_iter += s.niter + c0[3];
    n_pass += fit_passes + c0[4];
    atomicAdd(&sh_rc[s.rc & 15], 1u);
    if (a.params_out)
      for (int i = 0; i < P; ++i) a.params_out[slot * P + i] = s.x[i];
    if (a.props_out)
      for (int i = 0; i < Q; ++i) a.props_out[slot * Q + i] = props[i];
    if (a.counters_out) {
      int* co = a.counters_out + slot * 10;
      for (int i = 0; i < 5; ++i) co[i] = c0[i];
      co[5] = s.rc;
      co[6] = s.nfcall;
      co[7] = s.ngcall;
      co[8] = s.niter;
      co[9] = fit_passes;
    }
  };

  // take the next fit: load its starting point and the sums of its initial pass (stage A);
  // false = the work counter of this launch is exhausted
  auto load_next = [&]() -> bool {
    const unsigned long long kk = atomicAdd(a.next_id, 1ULL);
    if (kk >= (unsigned long long)a.nres) return false;
    slot = (long long)kk;
    double xs[P];
    if (a.call == 0) {
#pragma unroll
      for (int i = 0; i < P; ++i) xs[i] = a.start[i];
      s.reset_fit();
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) xs[i] = a.xbuf[i * a.cap + slot];
      s.mlstgd = a.istate[slot];
      s.lsvmin_ix = a.istate[a.cap + slot];
      s.anomaly = 0;
    }
    s.begin(xs);
    s.res_f = a.init[slot];
#pragma unroll
    for (int i = 0; i < P; ++i) s.res_g[i] = a.init[(1 + i) * a.cap + slot];
#pragma unroll
    for (int i = 0; i < PP; ++i) s.res_G[i] = a.init[(1 + P + i) * a.cap + slot];
    fit_passes = 1;
    return true;
  };
  active = load_next();

  // The solver step has ONE call site (its inlined code -- tens of KB -- exists once in the kernel;
  // two copies measurably thrash the 32 KB instruction cache when the fits diverge, C5): on the
  // first trip every lane consumes the initial sums of its first fit instead of making a pass.
  bool first_trip = true;
  while (__any_sync(0xffffffffu, active)) {
    if (active && !first_trip) {
      PassPoint<P> pa, pb;
      pa.set(s.x, s.hA);   // (the steps are not used by an analytic pass)
      pb.set(s.x0, s.hJ);
      PassAcc<P> acc;
      acc.zero();
      const unsigned long long id = (unsigned long long)(a.first + slot);
      const double* zrow = INJECT ? a.z_inject + (unsigned long long)slot * a.n : nullptr;
      pass_loop<M, INJECT, true, JAC>(rows, aux, a.n, a.n_pad, a.seed, id, zrow, pa, pb, acc);
      if constexpr (JAC == JAC_ANALYTIC)
        pass_finish_analytic<P>(acc, s);
      else
        pass_finish<P>(acc, s);
      ++fit_passes;
    }
    // ---- the solver step of every lane that has just made a pass (or taken its first fit), and of
    // the lanes whose fit ended and that took another one: the lanes that call resume() together pass
    // their mask, so that the routine can keep them converged (nl2_solver.cuh)
    bool need = active;
    for (;;) {
      const unsigned int lanes = __ballot_sync(0xffffffffu, need);
      if (lanes == 0u) break;
      if (need) {
        bool more = s.resume(lanes);
        if (more && fit_passes > 2 * (k::mxfcal + k::mxiter)) {  // safety cap on passes per call
          s.anomaly |= 8;
          s.rc = 9;
          more = false;
        }
        if (more) {
          need = false;
        } else {
          finish_call();  // (also a call that ended without a single trial step, e.g. residual too big at the start)
          if (!load_next()) {
            active = false;
            need = false;
          }
        }
      }
    }
    first_trip = false;
  }
  if (a.call == 0) return;

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
