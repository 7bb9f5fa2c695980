// Round-synchronous schedule of the bootstrap fits ("convergence mask + compaction" at the
// grid level; the alternative to the refill kernel of bootstrap_kernel.cuh, chosen by measurement
// per workload -- DESIGN.md section 2.8).
//
// Replaces the same reference code as iterate_kernel (bootstrap.f90:309-327, fit_data 215-226 and
// everything nl2sno calls), with the work of one fit split between two kernels that alternate:
//
//   round_solve_kernel   one thread per SLOT of a pool of C slots: consumes the sums of the pass just
//                        made (Nl2::resume -- assess, model switch, secant update, the next
//                        Levenberg-Marquardt / gqtstp step), and either asks for another pass or
//                        ends the nl2sno call.  A fit whose FIRST call ends continues in place with
//                        the second (bootstrap.f90:222-225); a fit whose second call ends hands its
//                        slot to the next fit of the chunk (work counter).  Slots that want a pass
//                        are appended to one of two lists: with the J(x0)^T r(x) part (every trial
//                        pass) or without (the first pass of the second call).
//   round_ctl_kernel     one thread: publishes the list lengths, clears the other pair of lists,
//                        and tells the CUDA graph's WHILE node whether another round is needed.
//   round_pass_kernel    one thread per list entry: the streaming pass over the n data points
//                        (pass_loop, identical arithmetic to the refill kernel), nothing else.
//
// Why: in the refill kernel the p x p solver step runs with 8 warps per SM whose lanes take
// different branches of ~100 KB of inlined code.  For well-behaved fits (C1, C2) that costs ~15 %;
// for the stress config C5 (23 % non-converging fits, 77 passes per fit) ncu shows 2.5 of 32 lanes
// active in the solver code, 4.2 "no instruction" stall cycles per issued instruction -- the
// solver's code evicts the pass loop from the 32 KB instruction cache -- and an FP64 pipe 30 % busy
// (profiles/r2_ncu_c5_summary.txt).  Here the pass kernel is 12 KB of straight-line code with all 32
// lanes busy, and the divergent solver code runs at full occupancy where its latencies overlap.
// Cost: the solver state (sizeof(Nl2<P>): 1.7 KB for p = 4) makes a round trip through HBM/L2 per
// pass, ~3.4 KB per fit and pass against ~190 000 instructions of pass work at n = 1000.
//
// The per-fit arithmetic is the refill kernel's, operation for operation (same pass_loop, same
// resume()), so a fit's result does not depend on the schedule (tests/test_round_schedule.py).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "model_registry.cuh"

namespace fmc {

#ifndef FMC_ROUND_SOLVE_THREADS
#define FMC_ROUND_SOLVE_THREADS 128
#endif
// Pass kernel: blocks of 128 threads.  Without the solver's state the row loop of a p <= 4 model
// needs ~160 registers and no stack (ptxas), so three blocks = 12 warps fit on an SM where the refill
// kernel has 8; p = 10 needs all 255 registers and gets two blocks.
#ifndef FMC_ROUND_PASS_THREADS
#define FMC_ROUND_PASS_THREADS 128
#endif

// ctl words
enum { RC_KINDS = 3 };  // what a slot asks for: 0 = pass with the J(x0)^T r(x) part, 1 = without, 2 = the residual sum alone
enum { RC_ROUND = 0, RC_COUNTS = 1, RC_DONE = 1 + 2 * RC_KINDS, RC_ROUNDS_TOTAL = 2 + 2 * RC_KINDS, RC_ERROR = 3 + 2 * RC_KINDS,
       RC_NWORDS = 4 + 2 * RC_KINDS };  // counts: [parity][kind]

struct RoundArgs {
  long long nslots;            // C: slots of the pool
  unsigned long long* state;   // [W][C]   solver state images (Nl2<P> as 64-bit words), word-major
  long long* slot_fit;         // [C]      chunk-local index of the fit in the slot, < 0 = idle
  int* slot_meta;              // [2][C]   0 = nl2sno call of fit_data the fit is in, 1 = passes of that call so far
  int* list;                   // [2][RC_KINDS][C] slots that want a pass: [parity of the round][kind][.]
  unsigned int* ctl;           // [RC_NWORDS]
  int spec;                    // Nl2::spec of the fits: 1 = Jacobian with every trial point, 0 = residual sum first
};

template <int P>
constexpr int state_words() {
  return (int)((sizeof(Nl2<P>) + 7) / 8);
}

// The evaluation points and the pass results are the leading members of Nl2<P>, so the pass kernel
// reads and writes them in place inside the state image (no second copy of them anywhere).
template <int P>
struct StateLayout {
  static constexpr int PP = P * (P + 1) / 2;
  static constexpr int X = 0, HA = P, X0 = 2 * P, HJ = 3 * P, RES_F = 4 * P, RES_G = 4 * P + 1,
                       RES_GG = 5 * P + 1, RES_LKY = 5 * P + 1 + PP;
  static_assert(offsetof(Nl2<P>, x) == 8 * X && offsetof(Nl2<P>, hA) == 8 * HA && offsetof(Nl2<P>, x0) == 8 * X0 &&
                    offsetof(Nl2<P>, hJ) == 8 * HJ && offsetof(Nl2<P>, res_f) == 8 * RES_F &&
                    offsetof(Nl2<P>, res_g) == 8 * RES_G && offsetof(Nl2<P>, res_G) == 8 * RES_GG &&
                    offsetof(Nl2<P>, res_lky) == 8 * RES_LKY,
                "round_pass_kernel addresses these members by word offset");
};

// ---------------------------------------------------------------------------------------
// Where the solver state of a slot lives while its step runs, and how many slots share a warp.
//
// The state (Nl2<P>: 1.5 KB for p = 4) is indexed dynamically all over the solver, so in a thread's
// own frame it is LOCAL memory: with 640 resident threads per SM that is 1 MB per SM against a 256 KB
// L1, every access of the divergent solver code an L2 round trip (ncu, C5: 16 long-scoreboard stall
// cycles per issued instruction, 2.2 GB of DRAM traffic per launch for 0.34 GB of state).  For small p
// the images are therefore staged in SHARED memory -- image j of a block at word j * STRIDE, STRIDE
// odd, so that lanes touching the same member (or different members) never share a bank -- filled by
// 8-byte cp.async straight from the word-major pool (all 32 lanes of a warp copy, coalesced, no
// registers in between) and written back the same way.
//
// Shared memory then bounds the slots in flight per SM (128 for p = 4: four warps), and the solver
// step is latency-bound, branchy code: ncu on C5 shows 4 of 32 lanes active per issued instruction.
// That is NOT 32 lanes each walking its own path -- giving a warp only LANES = 16 / 8 / 4 slots (its
// other lanes help with the copies and idle through the step; 2x / 4x / 8x the warps for the same
// shared memory) was measured SLOWER on C5 and C2 (5.66 / 5.24 / 4.57 against 5.80e5 fits/s on C5,
// profiles/r2_solve_lanes.json): a warp's instruction stream is set by its slowest lane (the number of
// More-Hebden iterations in lmstep), not by the sum over its lanes, so more, emptier warps execute
// more instructions in total and thrash the instruction cache again.  LANES stays 32.
// Large p (image > SMEM_STATE_MAX_WORDS) keeps the local frame: there the pass over 10^4 rows dwarfs
// the solver step anyway.
constexpr int SMEM_STATE_MAX_WORDS = 224;
#ifndef FMC_ROUND_SOLVE_LANES
#define FMC_ROUND_SOLVE_LANES 32
#endif
template <int P>
struct SolveCfg {
  static constexpr int W = (int)((sizeof(Nl2<P>) + 7) / 8);       // = state_words<P>()
  static constexpr bool IN_SMEM = W <= SMEM_STATE_MAX_WORDS;
  static constexpr int LANES = IN_SMEM ? FMC_ROUND_SOLVE_LANES : 32;   // slots per warp
  // one-warp blocks when a warp's images fill a quarter of the SM's shared memory: a block then ends with its warp
  static constexpr int THREADS = (IN_SMEM && LANES == 32) ? 32 : FMC_ROUND_SOLVE_THREADS;
  static constexpr int SLOTS = THREADS / 32 * LANES;                // slots per block
  static constexpr int STRIDE = W | 1;                              // words between neighbouring images (odd)
  static constexpr size_t SMEM_BYTES = IN_SMEM ? sizeof(unsigned long long) * SLOTS * STRIDE : 0;
};

FMC_INLINE __device__ void cp_async_8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
FMC_INLINE __device__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <class M, int JAC>
__global__ void __launch_bounds__(SolveCfg<M::NPARAM>::THREADS) round_solve_kernel(const KernelArgs* __restrict__ ap,
                                                                                   const RoundArgs ra) {
  constexpr int P = M::NPARAM;
  constexpr int PP = P * (P + 1) / 2;
  using Cfg = SolveCfg<P>;
  constexpr int W = Cfg::W;
  constexpr bool IN_SMEM = Cfg::IN_SMEM;
  const long long C = ra.nslots;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // slots of this warp: [wslot0, wslot0 + LANES), one per lane < LANES
  const long long wslot0 = ((long long)blockIdx.x * (Cfg::THREADS / 32) + warp) * Cfg::LANES;
  const long long slot = lane < Cfg::LANES ? wslot0 + lane : C;
  const unsigned int round = ra.ctl[RC_ROUND];
  // No lane leaves before the end: the lanes of a warp vote on who calls the solver step together.
  // round 0 of a chunk: every slot takes a fit; later rounds: only slots whose fit has just had a pass.
  long long fit = (slot < C && round != 0) ? ra.slot_fit[slot] : -1;
  enum { IDLE = 0, FETCH = 1, STEP = 2, PARK = 3 };
  int todo = slot >= C ? IDLE : (round == 0 ? FETCH : (fit >= 0 ? STEP : IDLE));
  if (IN_SMEM && !__any_sync(0xffffffffu, todo != IDLE)) return;  // (a warp of idle slots: nothing to stage)

  const long long cap = ap->cap, nres = ap->nres;
  extern __shared__ __align__(16) unsigned long long solve_smem[];
  union Image {
    Nl2<P> s;
    unsigned long long w[W];
    __device__ Image() {}
  };
  Image local_image;  // (unused, and removed by the compiler, when the images are in shared memory)
  unsigned long long* const wimg = solve_smem + (size_t)warp * Cfg::LANES * Cfg::STRIDE;  // the warp's images
  unsigned long long* const img = IN_SMEM ? wimg + (size_t)(lane < Cfg::LANES ? lane : 0) * Cfg::STRIDE : local_image.w;
  Nl2<P>& s = *reinterpret_cast<Nl2<P>*>(img);
  int call = 0, fit_passes = 0;
  if constexpr (IN_SMEM) {
    // all 32 lanes copy the warp's LANES images: element e = (word w, image j), j fastest -- consecutive
    // lanes read consecutive slots of one word (the pool is word-major)
    if (round != 0) {
      const int nvalid = (int)((C - wslot0 < (long long)Cfg::LANES) ? (C - wslot0) : (long long)Cfg::LANES);
      for (int e = lane; e < W * Cfg::LANES; e += 32) {
        const int w = e / Cfg::LANES, j = e % Cfg::LANES;
        if (j < nvalid) cp_async_8(wimg + (size_t)j * Cfg::STRIDE + w, ra.state + (long long)w * C + wslot0 + j);
      }
      cp_async_wait_all();
    }
    __syncwarp();
  } else if (todo == STEP) {
#pragma unroll 4
    for (int w = 0; w < W; ++w) img[w] = ra.state[(long long)w * C + slot];
  }
  if (todo == STEP) {
    call = ra.slot_meta[slot];
    fit_passes = ra.slot_meta[C + slot] + 1;  // the pass whose sums are in the image
  }
  for (;;) {
    if (todo == FETCH) {
      // ---- take the next fit of the chunk: starting point + the sums of its initial pass (stage A)
      const unsigned long long kk = atomicAdd(ap->next_id, 1ULL);
      if (kk >= (unsigned long long)nres) {
        ra.slot_fit[slot] = -1;
        todo = IDLE;
      } else {
        fit = (long long)kk;
        double xs[P];
#pragma unroll
        for (int i = 0; i < P; ++i) xs[i] = ap->start[i];
        s.reset_fit();
        s.spec = ra.spec;
        s.begin(xs);
        s.res_f = ap->init[fit];
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) s.res_g[i] = ap->init[(size_t)(1 + i) * cap + fit];
        FMC_SOLVER_LOOP
        for (int i = 0; i < PP; ++i) s.res_G[i] = ap->init[(size_t)(1 + P + i) * cap + fit];
        call = 0;
        fit_passes = 1;
        todo = STEP;
      }
    }
    const unsigned int lanes = __ballot_sync(0xffffffffu, todo == STEP);
    if (lanes == 0u) break;
    if (todo == STEP) {
      bool more = s.resume(lanes);
      if (more && fit_passes > 2 * (k::mxfcal + k::mxiter)) {  // safety cap on passes per call
        s.anomaly |= 8;
        s.rc = 9;
        more = false;
      }
      if (more) {
        todo = PARK;
      } else {
        // ---- end of an nl2sno call
        if (s.anomaly != 0 && ap->anom != nullptr) {
          const unsigned long long kk = atomicAdd(ap->anom, 1ULL);
          if (kk < (unsigned long long)ANOM_CAP) {
            ap->anom[1 + 2 * kk] = (unsigned long long)(ap->first + fit);
            ap->anom[2 + 2 * kk] = (unsigned long long)(s.anomaly | (call << 8));
          }
          s.anomaly = 0;
        }
        int* cnt = ap->cnt + (size_t)(5 * call) * cap + fit;
        cnt[0 * cap] = s.rc;
        cnt[1 * cap] = s.nfcall;
        cnt[2 * cap] = s.ngcall;
        cnt[3 * cap] = s.niter;
        cnt[4 * cap] = fit_passes;
        if (call == 0) {
          // fit_data calls nl2sno a second time from the answer of the first (bootstrap.f90:222-225):
          // the fit stays in its slot; iv(mlstgd) and lsvmin's generator carry over as in the reference
          double xs[P];
#pragma unroll
          for (int i = 0; i < P; ++i) xs[i] = s.x[i];
          s.begin(xs);
          call = 1;
          fit_passes = 0;
          todo = PARK;  // wants the first evaluation of the call: residual + Jacobian at x, D = 1, no x0 part
        } else {
#pragma unroll
          for (int i = 0; i < P; ++i) ap->xbuf[(size_t)i * cap + fit] = s.x[i];
          todo = FETCH;
        }
      }
    }
  }
  // ---- slots that want another pass: park the state, queue the slot
  if constexpr (IN_SMEM) {
    __syncwarp();
    const unsigned int parked = __ballot_sync(0xffffffffu, todo == PARK);
    for (int e = lane; e < W * Cfg::LANES; e += 32) {
      const int w = e / Cfg::LANES, j = e % Cfg::LANES;
      if ((parked >> j) & 1u) ra.state[(long long)w * C + wslot0 + j] = wimg[(size_t)j * Cfg::STRIDE + w];
    }
    if (todo != PARK) return;
  } else {
    if (todo != PARK) return;
#pragma unroll 8
    for (int w = 0; w < W; ++w) ra.state[(long long)w * C + slot] = img[w];
  }
  ra.slot_fit[slot] = fit;
  ra.slot_meta[slot] = call;
  ra.slot_meta[C + slot] = fit_passes;
  const unsigned int par = round & 1u;
  const unsigned int kind = !s.need_jac ? 2u : (s.need_b ? 0u : 1u);
  const unsigned int pos = atomicAdd(&ra.ctl[RC_COUNTS + RC_KINDS * par + kind], 1u);
  ra.list[(size_t)(RC_KINDS * par + kind) * C + pos] = (int)slot;
}

// ---------------------------------------------------------------------------------------
// One thread.  `handle` is the WHILE node's condition when the rounds run inside a CUDA graph
// (use_handle != 0); the host-driven fallback reads ctl instead.
template <class M>  // (a template only so that the header can be included by several translation units)
__global__ void round_ctl_kernel(const RoundArgs ra, cudaGraphConditionalHandle handle, int use_handle) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned int round = ra.ctl[RC_ROUND];
  const unsigned int par = round & 1u;
  unsigned int n = 0u;
  for (int kk = 0; kk < RC_KINDS; ++kk) n += ra.ctl[RC_COUNTS + RC_KINDS * par + kk];
  // a chunk needs at most (passes per call cap) x 2 calls rounds; whatever happens, the loop ends
  const bool runaway = n != 0u && round > 4u * (unsigned int)(k::mxfcal + k::mxiter) + 16u;
  if (runaway) {
    n = 0u;
    ra.ctl[RC_ERROR] = 1u;
  }
  for (int kk = 0; kk < RC_KINDS; ++kk) ra.ctl[RC_COUNTS + RC_KINDS * (par ^ 1u) + kk] = 0u;
  ra.ctl[RC_ROUND] = round + 1u;
  ra.ctl[RC_DONE] = n == 0u ? 1u : 0u;
  ra.ctl[RC_ROUNDS_TOTAL] += 1u;
  if (use_handle) cudaGraphSetConditional(handle, n != 0u ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------
template <class M, bool INJECT, bool SMEM, int JAC>
__global__ void __launch_bounds__(FMC_ROUND_PASS_THREADS, (M::NPARAM <= 6 ? 3 : 2)) round_pass_kernel(const KernelArgs* __restrict__ ap,
                                                                                 const RoundArgs ra) {
  constexpr int P = M::NPARAM;
  constexpr int PP = P * (P + 1) / 2;
  using L = StateLayout<P>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long C = ra.nslots;
  const unsigned int par = (ra.ctl[RC_ROUND] - 1u) & 1u;  // (the control kernel has advanced the counter)
  const unsigned int nA = ra.ctl[RC_COUNTS + RC_KINDS * par], nN = ra.ctl[RC_COUNTS + RC_KINDS * par + 1];
  const unsigned int blocksA = (nA + FMC_ROUND_PASS_THREADS - 1) / FMC_ROUND_PASS_THREADS;
  // blocks [0, blocksA) take the list "with B", the following ones the list "without"; the rest exit
  const bool with_b = blockIdx.x < blocksA;
  const unsigned int b = with_b ? blockIdx.x : blockIdx.x - blocksA;
  const unsigned int cnt = with_b ? nA : nN;
  if ((unsigned long long)b * FMC_ROUND_PASS_THREADS >= cnt) return;
  const KernelArgs a = *ap;
  const double* aux = nullptr;
  const RowData* rows = stage_rows<SMEM, model_naux<M>::value>(a, smem_raw, &aux);
  const unsigned int idx = b * FMC_ROUND_PASS_THREADS + threadIdx.x;
  if (idx >= cnt) return;
  const long long slot = ra.list[(size_t)(RC_KINDS * par + (with_b ? 0 : 1)) * C + idx];
  double* img = reinterpret_cast<double*>(ra.state) + slot;  // word w of the image: img[w * C]
  double x[P], h[P], x0[P], h0[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    x[i] = img[(long long)(L::X + i) * C];
    h[i] = img[(long long)(L::HA + i) * C];
    x0[i] = img[(long long)(L::X0 + i) * C];
    h0[i] = img[(long long)(L::HJ + i) * C];
  }
  PassPoint<P> pa, pb;
  pa.set(x, h);
  pb.set(x0, h0);
  PassAcc<P> acc;
  acc.zero();
  const long long fit = ra.slot_fit[slot];
  const unsigned long long id = (unsigned long long)(a.first + fit);
  const double* zrow = INJECT ? a.z_inject + (unsigned long long)fit * a.n : nullptr;
  if (with_b)
    pass_loop<M, INJECT, true, JAC>(rows, aux, a.n, a.n_pad, a.seed, id, zrow, pa, pb, acc);
  else
    pass_loop<M, INJECT, false, JAC>(rows, aux, a.n, a.n_pad, a.seed, id, zrow, pa, pb, acc);
  double f, g[P], G[PP], lky[P];
  if constexpr (JAC == JAC_ANALYTIC) {
    f = acc.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      g[i] = acc.g[i];
      lky[i] = acc.lky[i];
    }
#pragma unroll
    for (int i = 0; i < PP; ++i) G[i] = acc.G[i];
  } else {
    pass_scale<P>(acc, h, h0, f, g, G, lky);
  }
  img[(long long)L::RES_F * C] = f;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    img[(long long)(L::RES_G + i) * C] = g[i];
    img[(long long)(L::RES_LKY + i) * C] = lky[i];
  }
#pragma unroll
  for (int i = 0; i < PP; ++i) img[(long long)(L::RES_GG + i) * C] = G[i];
}

// ---------------------------------------------------------------------------------------
// The residual sum alone at a trial point (list kind 2; RoundArgs::spec = 0): the model value per row and
// f += r^2, accumulated in the order of pass_loop, so that the solver reads the same f as from a full
// pass.  No Jacobian: few registers, many resident warps.  Used where a Jacobian row costs many model
// evaluations' worth (p = 10: a full trial pass is ~10x this), so that the passes nobody needed the
// Jacobian of -- rejected steps, the last step of each nl2sno call -- cost a tenth.
#ifndef FMC_ROUND_FONLY_THREADS
#define FMC_ROUND_FONLY_THREADS 256
#endif
template <class M, bool INJECT, bool SMEM>
__global__ void __launch_bounds__(FMC_ROUND_FONLY_THREADS) round_fonly_kernel(const KernelArgs* __restrict__ ap,
                                                                              const RoundArgs ra) {
  constexpr int P = M::NPARAM;
  constexpr int NAUX = model_naux<M>::value;
  using L = StateLayout<P>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long C = ra.nslots;
  const unsigned int par = (ra.ctl[RC_ROUND] - 1u) & 1u;
  const unsigned int cnt = ra.ctl[RC_COUNTS + RC_KINDS * par + 2];
  if ((unsigned long long)blockIdx.x * FMC_ROUND_FONLY_THREADS >= cnt) return;
  const KernelArgs a = *ap;
  const double* aux = nullptr;
  const RowData* rows = stage_rows<SMEM, NAUX>(a, smem_raw, &aux);
  const unsigned int idx = blockIdx.x * FMC_ROUND_FONLY_THREADS + threadIdx.x;
  if (idx >= cnt) return;
  const long long slot = ra.list[(size_t)(RC_KINDS * par + 2) * C + idx];
  double* img = reinterpret_cast<double*>(ra.state) + slot;
  double x[P];
#pragma unroll
  for (int i = 0; i < P; ++i) x[i] = img[(long long)(L::X + i) * C];
  const long long fit = ra.slot_fit[slot];
  const unsigned long long id = (unsigned long long)(a.first + fit);
  const double* zrow = INJECT ? a.z_inject + (unsigned long long)fit * a.n : nullptr;
  double f = 0.0;
  auto row = [&](int r, float zf) {
    const RowData rd = rows[r];
    const double zz = INJECT ? zrow[r] : (double)zf;
    const double yfit = fma(zz, rd.err, rd.y);  // offset_data, bootstrap.f90:208
    const double* ra_ = NAUX > 0 ? aux + r * NAUX : nullptr;
    if (NAUX > 0) __builtin_assume(ra_ != nullptr);
    const double m0 = eval_model<M>(x, rd.x, ra_);
    const double res = (yfit - m0) * rd.rerr;   // madr, bootstrap.f90:247
    f = fma(res, res, f);
  };
  const int n4 = a.n & ~3;
  for (int i = 0; i < n4; i += 4) {
    float z[4];
    if (!INJECT) normals4(a.seed, id, (uint32_t)(i >> 2), z);
#pragma unroll
    for (int m = 0; m < 4; ++m) row(i + m, z[m]);
  }
  if (n4 < a.n) {
    float z[4];
    if (!INJECT) normals4(a.seed, id, (uint32_t)(n4 >> 2), z);
#pragma unroll
    for (int m = 0; m < 3; ++m)
      if (n4 + m < a.n) row(n4 + m, z[m]);
  }
  img[(long long)L::RES_F * C] = f;
}

// ---------------------------------------------------------------------------------------
// After the rounds of a chunk: derived properties, compensated moment sums of shifted values,
// optional per-resample outputs, return-code histogram and counters (bootstrap.f90:316-318),
// from the per-fit results the solve kernel left in xbuf / cnt.  Fits are dealt to threads by
// index, so a thread's sums do not depend on the order in which the fits finished.
template <class M>
__global__ void __launch_bounds__(256) round_moments_kernel(const KernelArgs* __restrict__ ap) {
  constexpr int P = M::NPARAM;
  constexpr int Q = M::NPROP;
  constexpr int NM = 2 * (P + Q);
  __shared__ double red[256 / 32];
  __shared__ unsigned int sh_rc[16];
  if (threadIdx.x < 16) sh_rc[threadIdx.x] = 0u;
  __syncthreads();
  const KernelArgs& a = *ap;
  const long long cap = a.cap, nres = a.nres;
  double msum[NM], mcmp[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    msum[i] = 0.0;
    mcmp[i] = 0.0;
  }
  double cq[Q > 0 ? Q : 1];
  M::derived_properties(a.start, cq);
  unsigned int n_fit = 0, n_fcall = 0, n_gcall = 0, n_iter = 0, n_pass = 0;
  for (long long fit = (long long)blockIdx.x * blockDim.x + threadIdx.x; fit < nres;
       fit += (long long)gridDim.x * blockDim.x) {
    double x[P], props[Q > 0 ? Q : 1];
#pragma unroll
    for (int i = 0; i < P; ++i) x[i] = a.xbuf[(size_t)i * cap + fit];
    M::derived_properties(x, props);  // bootstrap.f90:316
#pragma unroll
    for (int i = 0; i < P; ++i) {     // bootstrap.f90:317-318, on shifted values
      const double dv = x[i] - a.start[i];
      kahan_add(msum[i], mcmp[i], dv);
      kahan_add(msum[P + i], mcmp[P + i], dv * dv);
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const double dv = props[i] - cq[i];
      kahan_add(msum[2 * P + i], mcmp[2 * P + i], dv);
      kahan_add(msum[2 * P + Q + i], mcmp[2 * P + Q + i], dv * dv);
    }
    int c[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) c[i] = a.cnt[(size_t)i * cap + fit];
    ++n_fit;
    n_fcall += c[1] + c[6];
    n_gcall += c[2] + c[7];
    n_iter += c[3] + c[8];
    n_pass += c[4] + c[9];
    atomicAdd(&sh_rc[c[5] & 15], 1u);
    if (a.params_out)
      for (int i = 0; i < P; ++i) a.params_out[fit * P + i] = x[i];
    if (a.props_out)
      for (int i = 0; i < Q; ++i) a.props_out[fit * Q + i] = props[i];
    if (a.counters_out)
      for (int i = 0; i < 10; ++i) a.counters_out[fit * 10 + i] = c[i];
  }
  auto block_sum = [&](double v) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
      for (int w = 0; w < 256 / 32; ++w) t += red[w];
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

// ---------------------------------------------------------------------------------------
// Host-side binding of one model to its round-kernel instantiations (kernels_round.cu).
struct RoundEntry {
  int state_words;  // 64-bit words of one solver state image
  // enqueue one round (solve, control, pass) on `st` -- inside a stream capture or directly
  cudaError_t (*enqueue_round)(const KernelArgs* ap, const RoundArgs& ra, bool inject, bool smem, int jac,
                               size_t shmem, cudaGraphConditionalHandle handle, int use_handle, cudaStream_t st);
  cudaError_t (*enqueue_moments)(const KernelArgs* ap, int grid, cudaStream_t st);
  cudaError_t (*prepare)(bool inject, bool smem, int jac, size_t shmem, int* pass_blocks_per_sm);
};

template <class M>
struct RoundBinding {
  using MB = ModelBinding<M>;
  static cudaError_t enqueue_round(const KernelArgs* ap, const RoundArgs& ra, bool inject, bool smem, int jac,
                                   size_t shmem, cudaGraphConditionalHandle handle, int use_handle,
                                   cudaStream_t st) {
    constexpr int ST = SolveCfg<M::NPARAM>::THREADS, SS = SolveCfg<M::NPARAM>::SLOTS;
    const unsigned int sgrid = (unsigned int)((ra.nslots + SS - 1) / SS);
    // both lists together hold at most C slots: ceil(nA/T) + ceil(nN/T) <= C/T + 2 blocks
    const unsigned int pgrid = (unsigned int)(ra.nslots / FMC_ROUND_PASS_THREADS + 2);
    return MB::dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      round_solve_kernel<M, decltype(J)::value><<<sgrid, ST, SolveCfg<M::NPARAM>::SMEM_BYTES, st>>>(ap, ra);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      round_ctl_kernel<M><<<1, 32, 0, st>>>(ra, handle, use_handle);
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      round_pass_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>
          <<<pgrid, FMC_ROUND_PASS_THREADS, shmem, st>>>(ap, ra);
      e = cudaGetLastError();
      if (e != cudaSuccess || ra.spec != 0) return e;
      const unsigned int fgrid = (unsigned int)(ra.nslots / FMC_ROUND_FONLY_THREADS + 1);
      round_fonly_kernel<M, decltype(I)::value, decltype(S)::value><<<fgrid, FMC_ROUND_FONLY_THREADS, shmem, st>>>(ap, ra);
      return cudaGetLastError();
    });
  }
  static cudaError_t enqueue_moments(const KernelArgs* ap, int grid, cudaStream_t st) {
    round_moments_kernel<M><<<grid, 256, 0, st>>>(ap);
    return cudaGetLastError();
  }
  static cudaError_t prepare(bool inject, bool smem, int jac, size_t shmem, int* nb) {
    return MB::dispatch(inject, smem, jac, [&](auto I, auto S, auto J) {
      cudaError_t e = cudaSuccess;
      if (SolveCfg<M::NPARAM>::SMEM_BYTES > 0) {
        e = cudaFuncSetAttribute(round_solve_kernel<M, decltype(J)::value>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SolveCfg<M::NPARAM>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
      }
      e = cudaFuncSetAttribute(round_fonly_kernel<M, decltype(I)::value, decltype(S)::value>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
      if (e != cudaSuccess) return e;
      auto kern = round_pass_kernel<M, decltype(I)::value, decltype(S)::value, decltype(J)::value>;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, kern, FMC_ROUND_PASS_THREADS, shmem);
    });
  }
  static RoundEntry entry() { return RoundEntry{state_words<M::NPARAM>(), enqueue_round, enqueue_moments, prepare}; }
};

// kernels_round.cu, one object per model
RoundEntry fmc_round_entry_model0();
RoundEntry fmc_round_entry_model1();
RoundEntry fmc_round_entry_model2();
RoundEntry fmc_round_entry_model3();
inline RoundEntry fmc_round_entry(int model) {
  switch (model) {
    case 0: return fmc_round_entry_model0();
    case 1: return fmc_round_entry_model1();
    case 2: return fmc_round_entry_model2();
    case 3: return fmc_round_entry_model3();
  }
  return RoundEntry{0, nullptr, nullptr, nullptr};
}

}  // namespace fmc
