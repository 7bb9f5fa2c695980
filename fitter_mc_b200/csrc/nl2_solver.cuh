// Per-fit NL2SOL (adaptive nonlinear least squares) solver state machine for the
// B200 bootstrap engine -- one instance per thread, all state in registers /
// thread-local memory.
//
// What it replaces: `nl2sno` + `nl2itr` + `assess` + `lmstep` + `gqtstp` + helpers of the
// reference (toms573.f90:590-769, 772-1444, 1447-1952, 3397-3839, 2471-3039, ...), as
// called twice per resample by `fit_data` (bootstrap.f90:215-226).
//
// How it differs (B200-first design):
//  * The reference keeps the n x p Jacobian, its Householder factors, r, rsave and qtr
//    per thread (57 KB at n=1000, p=4).  Here NOTHING of length n is stored per fit.  The
//    solver talks to its caller through streaming "passes" over the data: a pass at a point
//    xa returns  sum r^2,  g = J(xa)^T r(xa),  G = J(xa)^T J(xa)  and, for a second point
//    xb, J(xb)^T r(xa).  Everything NL2SOL needs follows from those p x p quantities:
//      - qrfact/qapply (pivoted Householder QR of J, Q^T r)  ->  pivoted Cholesky of G,
//        qtr = R^-T P^T g      (same R up to row signs, which lmstep is invariant to)
//      - dupdat column norms                                  ->  sqrt(diag G)
//      - H = D^-1 (J^T J + S) D^-1 for gqtstp                 ->  from G directly
//      - lky = J(x0)^T r(x) for the secant update of S        ->  the xb part of the pass
//  * A trial pass evaluates the Jacobian at the trial point speculatively, so an accepted
//    step (the normal case) needs no second pass over the data.
//  * The covariance computation that the reference performs at convergence and never reads
//    (toms573.f90:1400-1418) is not performed; it cannot change x.
//  * The final gradient evaluation after a converged/limited step (toms573.f90:1055-1061,
//    1259-1262 -> 450) only feeds that covariance and is skipped too.
//
// Control flow keeps the reference's statement labels (L100, L180, L300, ...) so it can be
// reviewed against toms573.f90 and against the oracle side by side.
#pragma once
#include "fmc_common.cuh"

// The loops of the solver have compile-time trip counts (p, p(p+1)/2).  By default the compiler
// unrolls them, which is what keeps a per-THREAD solver's small arrays in registers.  The
// warp-per-fit kernels (warp_kernel.cuh) keep the solver state in shared memory, where dynamic
// indexing is free, and define FMC_SOLVER_ROLLED: unrolled for p = 10 the solver is 25 000
// instructions (400 KB) of branchy code against a 32 KB instruction cache -- measured (ncu,
// profiles/r2_ncu_c4_warp_unrolled_summary.txt) as 4.3 "no instruction" stall cycles per issued
// instruction, the solver step costing more than the pass over 10^4 data points it follows.
//
// The round schedule's solve kernel (round_kernels.cuh) defines it as well: there the state sits in
// shared memory too, and what limits the kernel is the SIZE of the code its divergent lanes walk
// through (ncu: 16 "no instruction" stall cycles per issued instruction with the unrolled, fully inlined
// 20 000-instruction solver).  In that mode the small dense helpers -- inlined at some fifty call sites,
// each with its copy of the FP64 square root and of v2norm's overflow-safe fallback -- become real
// functions (FMC_SOLVER_FN).
#if defined(FMC_SOLVER_ROLLED) && !defined(FMC_SOLVER_OUTLINE)
#define FMC_SOLVER_OUTLINE 1
#endif
#if defined(FMC_SOLVER_ROLLED) && defined(__CUDACC__)
#define FMC_SOLVER_LOOP _Pragma("unroll 1")
#else
#define FMC_SOLVER_LOOP
#endif
#if defined(FMC_SOLVER_OUTLINE) && defined(__CUDACC__)
#define FMC_SOLVER_FN inline __noinline__
#else
#define FMC_SOLVER_FN inline
#endif

namespace fmc {

// ----------------------------------------------------------------- small dense helpers
// (nl2_dense.inc) dotn, nrm2, ltsolve_t, ltsolve, chol_rows, symv -- the flavour of the common path ...
#define FMC_DENSE_NAME(f) f
#define FMC_DENSE_LOOP FMC_SOLVER_LOOP
#define FMC_DENSE_ATTR FMC_SOLVER_FN
#if defined(FMC_SOLVER_OUTLINE) && defined(__CUDACC__)
#define FMC_DENSE_ATTR_SMALL FMC_INLINE
#define FMC_DENSE_ATTR_SMALL_OL FMC_SOLVER_FN
#else
#define FMC_DENSE_ATTR_SMALL FMC_INLINE
#define FMC_DENSE_ATTR_SMALL_OL FMC_INLINE
#endif
#include "nl2_dense.inc"
#undef FMC_DENSE_NAME
#undef FMC_DENSE_LOOP
#undef FMC_DENSE_ATTR
#undef FMC_DENSE_ATTR_SMALL
#undef FMC_DENSE_ATTR_SMALL_OL
// ... and the flavour of the RARE, long-running paths: names f_r, used by gqtstp, lsvmin and the
// More-Hebden iterations of lmstep.  In the refill kernel (FMC_SOLVER_RARE_ROLLED, kernels_model*.cu) the
// common path -- assess, the secant update, factor, the Gauss-Newton part of a step -- stays unrolled and
// inlined, which is what a well-behaved fit executes (C1, C2), while these paths and all loops of lmstep
// are rolled, with the helpers as real functions: 21 800 -> 13 300 instructions for p = 4.  Measured
// (profiles/r2_refill_solver_flavours.json): C5 +31 % (4.39e5 -> 5.75e5 fits/s), C2 +0.2 %, C1 -5 %;
// rolling everything gives C5 +48 % but C2 -3 % and C1 -30 %; rolling factor/adopt_jacobian too, C1 -17 %.
#if defined(FMC_SOLVER_RARE_ROLLED) && !defined(FMC_SOLVER_ROLLED) && defined(__CUDACC__)
#define FMC_RARE_LOOP _Pragma("unroll 1")
#define FMC_RARE_FN __noinline__
#define FMC_DENSE_NAME(f) f##_r
#define FMC_DENSE_LOOP _Pragma("unroll 1")
#define FMC_DENSE_ATTR inline __noinline__
#define FMC_DENSE_ATTR_SMALL inline __noinline__
#define FMC_DENSE_ATTR_SMALL_OL inline __noinline__
#include "nl2_dense.inc"
#undef FMC_DENSE_NAME
#undef FMC_DENSE_LOOP
#undef FMC_DENSE_ATTR
#undef FMC_DENSE_ATTR_SMALL
#undef FMC_DENSE_ATTR_SMALL_OL
#else
#define FMC_RARE_LOOP FMC_SOLVER_LOOP
#define FMC_RARE_FN
#define dotn_r dotn
#define nrm2_r nrm2
#define ltsolve_t_r ltsolve_t
#define ltsolve_r ltsolve
#define chol_rows_r chol_rows
#define symv_r symv
#endif

// ----------------------------------------------------------------- the solver
enum Nl2Phase { PH_INIT = 0, PH_TRIAL = 1, PH_JAC = 2 };

template <int P>
struct Nl2 {
  static constexpr int PP = P * (P + 1) / 2;
  static constexpr int WLEN = 4 * P + 7 + PP;  // w, with lmat at w + 4P+7 (toms573.f90:929-931)
  // v(1:18) of the reference, 0-based
  enum {
    DGNORM = 0, DSTNRM, DST0, GTSTEP, STPPAR, NREDUC, PREDUC, RADIUS, RAD0,  // = vsave set (nvsave = 9)
    F, FDIF, FLSTGD, F0, GTSLST, PLSTGD, RADFAC, RELDX, DSTSAV, NV
  };

  // ---- pass interface: the caller evaluates at x (FD steps hA) and, if need_b, at x0
  // (FD steps hJ), then fills res_* and calls resume().
  double x[P], hA[P], x0[P], hJ[P];
  double res_f;       // sum_i r_i(x)^2
  double res_g[P];    // J(x)^T r(x)
  double res_G[PP];   // J(x)^T J(x), lower triangle by rows
  double res_lky[P];  // J(x0)^T r(x)
  int need_b;
  // Whether a trial point is evaluated with its Jacobian right away (spec = 1: an accepted step, the
  // normal case of a well-behaved fit, then needs no second pass) or with the residual sum alone first
  // (spec = 0: need_jac = 0 for the trial pass; only an accepted step that does not end the call asks
  // for the Jacobian pass, through the same path as a step that assess moved back to an earlier trial
  // point).  The solver takes the same decisions on the same numbers either way; what changes is how
  // many passes compute a Jacobian nobody reads -- rejected steps and the last step of every call.
  int spec, need_jac;

  // ---- NL2SOL state (names follow the reference's iv/v maps)
  double step[P], stlstg[P], d[P], g[P], dig[P], lky[P], tmp1[P], tmp2[P], qtr[P];
  double S[PP], G[PP], rfac[PP], hmat[PP], w[WLEN];
  double v[NV], vsave[9];
  double size_, wscale_;
  int ipivot[P];
  int ierr, irc, mlstgd, model, nfcall, nfgcal, radinc, restor, stage, swtch, toobig, xirc, cnvcod,
      kagqt, kalm, niter, ngcall, sused, hvalid, stpmod, phase, x_is_eval, rc, lsvmin_ix;
  // Safety net for a massively parallel run: a fit must never spin forever.  Each bit records a
  // loop that hit its (generous) cap; the fit then ends like a false convergence.  The caps are
  // far above anything the reference's own bounds (kalim = 12 / 50, mxfcal, mxiter) allow in
  // exact arithmetic, so they only bite on non-finite or inconsistent intermediate values.
  int anomaly;
#ifdef FMC_TRACE
  // Decision trace for tools/parity_attribution.py (host image only; the shipped kernels are never
  // compiled with FMC_TRACE): the same record as the oracle's, after every call of assess.
  double* trace;
  int trace_cap, trace_len;
#endif

  // ------------------------------------------------------------------ entry points
  // Per-fit reset of the state that the reference leaves to persist between fits
  // (iv(mlstgd), lsvmin's SAVEd LCG): makes every fit independent of scheduling order.
  FMC_HD void reset_fit() {
    mlstgd = 0;
    lsvmin_ix = 2;
    anomaly = 0;
    spec = 1;
  }

  // nl2sno fresh start (toms573.f90:688-700): D = 1 for the first finite-difference Jacobian.
  FMC_HD void begin(const double* xs) {
    FMC_SOLVER_LOOP
    for (int i = 0; i < P; ++i) {
      x[i] = xs[i];
      x0[i] = xs[i];
      d[i] = 1.0;
    }
    set_fd_steps();
    FMC_SOLVER_LOOP
    for (int i = 0; i < P; ++i) hJ[i] = hA[i];
    need_b = 0;
    need_jac = 1;
    phase = PH_INIT;
    rc = 0;
  }

  // Finite-difference steps for the next Jacobian (toms573.f90:716).
  FMC_HD void set_fd_steps() {
    FMC_SOLVER_LOOP
    for (int i = 0; i < P; ++i) hA[i] = k::dltfdj * fmax(fabs(x[i]), 1.0 / d[i]);
  }

  // Consume the results of the pass just made.  Returns true if another pass is wanted,
  // false when this nl2sno call is over (rc = the reference's iv(1): 3..10, 13).
  // `lanes`: device only -- the mask of the lanes of this warp that make this call together (0 = the
  // calling lane is on its own); see the comment at the definition.
  FMC_HD bool resume(unsigned int lanes = 0u);

  // ------------------------------------------------------------------ pieces
  FMC_HD void init_iteration_state();
  FMC_HD void adopt_jacobian();
  FMC_HD void factor();
  FMC_HD void build_h();
  FMC_HD void hess_times_step();
  FMC_HD void assess();
  FMC_HD void lmstep();
  FMC_HD void gqtstp();
  FMC_HD double lsvmin(const double* l, double* xv, double* yv);
  FMC_HD void slupdt(const double* wchmtd);
};

// nl2itr label 10 (toms573.f90:904-945)
template <int P>
FMC_HD void Nl2<P>::init_iteration_state() {
  niter = 0;
  nfcall = 1;
  ngcall = 1;
  nfgcal = 1;
  toobig = 0;
  cnvcod = 0;
  kalm = -1;
  radinc = 0;
  hvalid = 0;
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) d[i] = 0.0;  // v(dinit) = 0: D is then set entirely by dupdat
  v[RAD0] = 0.0;
  v[STPPAR] = 0.0;
  v[RADIUS] = k::lmax0 / (1.0 + k::phmxfc);
  model = 1;
  FMC_SOLVER_LOOP
  for (int i = 0; i < PP; ++i) S[i] = 0.0;
}

// nl2itr label 60 (toms573.f90:972-999): take over the Jacobian just evaluated at x.
template <int P>
FMC_HD void Nl2<P>::adopt_jacobian() {
  kalm = -1;
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) {
    g[i] = res_g[i];
    hJ[i] = hA[i];
  }
  FMC_SOLVER_LOOP
  for (int i = 0; i < PP; ++i) G[i] = res_G[i];
  // dupdat (toms573.f90:2434-2468) with ||J(:,i)|| = sqrt(G_ii)
  int s1 = -1;
  FMC_SOLVER_LOOP
  for (int i = 1; i <= P; ++i) {
    s1 += i;
    double sii = S[s1];
    double t = sqrt(G[s1]);
    if (sii > 0.0) t = sqrt(t * t + sii);
    if (t < k::jtinit) t = fmax(k::d0init, k::jtinit);
    d[i - 1] = fmax(k::dfac * d[i - 1], t);
  }
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) dig[i] = g[i] / d[i];
  v[DGNORM] = nrm2(P, dig);
}

// Replaces qrfact + qapply (toms573.f90:4360-4567, 4281-4357): Cholesky of G = J^T J with
// the same pivot rule (largest remaining column norm, first maximum wins).  rfac holds R
// "upper triangle by columns" exactly as lmstep expects it; qtr = first p entries of Q^T r.
template <int P>
FMC_HD void Nl2<P>::factor() {
  double A[P][P];
  {
    int m = 0;
    FMC_SOLVER_LOOP
    for (int i = 0; i < P; ++i) {
      FMC_SOLVER_LOOP
      for (int j = 0; j <= i; ++j) {
        A[i][j] = G[m];
        A[j][i] = G[m];
        ++m;
      }
    }
  }
  FMC_SOLVER_LOOP
  for (int j = 0; j < P; ++j) ipivot[j] = j + 1;
  FMC_SOLVER_LOOP
  for (int i = 0; i < PP; ++i) rfac[i] = 0.0;
  ierr = 0;
  int rank = P;
#define FMC_R(i, j) rfac[(j) * ((j) + 1) / 2 + (i)]  // 0-based (i <= j)
  FMC_SOLVER_LOOP
  for (int kk = 0; kk < P; ++kk) {
    double sigma = 0.0;
    int jbar = -1;
    FMC_SOLVER_LOOP
    for (int j = kk; j < P; ++j) {
      if (sigma >= A[j][j]) continue;
      sigma = A[j][j];
      jbar = j;
    }
    if (jbar < 0) {  // no nonzero pivot left (qrfact label 120)
      ierr = kk + 1;
      rank = kk;
      break;
    }
    if (jbar != kk) {
      int ti = ipivot[kk];
      ipivot[kk] = ipivot[jbar];
      ipivot[jbar] = ti;
      FMC_SOLVER_LOOP
      for (int i = 0; i < P; ++i) {
        double t = A[i][kk];
        A[i][kk] = A[i][jbar];
        A[i][jbar] = t;
      }
      FMC_SOLVER_LOOP
      for (int i = 0; i < P; ++i) {
        double t = A[kk][i];
        A[kk][i] = A[jbar][i];
        A[jbar][i] = t;
      }
      FMC_SOLVER_LOOP
      for (int i = 0; i < kk; ++i) {
        double t = FMC_R(i, kk);
        FMC_R(i, kk) = FMC_R(i, jbar);
        FMC_R(i, jbar) = t;
      }
    }
    double rkk = sqrt(A[kk][kk]);
    FMC_R(kk, kk) = rkk;
    FMC_SOLVER_LOOP
    for (int j = kk + 1; j < P; ++j) FMC_R(kk, j) = A[kk][j] / rkk;
    FMC_SOLVER_LOOP
    for (int i = kk + 1; i < P; ++i) {
      FMC_SOLVER_LOOP
      for (int j = i; j < P; ++j) {
        double t = A[i][j] - FMC_R(kk, i) * FMC_R(kk, j);
        A[i][j] = t;
        A[j][i] = t;
      }
    }
  }
  // qtr = R^-T P^T g on the leading `rank` rows
  FMC_SOLVER_LOOP
  for (int kk = 0; kk < P; ++kk) {
    if (kk >= rank) {
      qtr[kk] = 0.0;
      continue;
    }
    double t = g[ipivot[kk] - 1];
    FMC_SOLVER_LOOP
    for (int i = 0; i < kk; ++i) t -= FMC_R(i, kk) * qtr[i];
    qtr[kk] = t / FMC_R(kk, kk);
  }
#undef FMC_R
}

// H = D^-1 (J^T J + S) D^-1 (toms573.f90:1113-1120; the variant at 1125-1165 rebuilds the
// same matrix from the QR factors).
template <int P>
FMC_HD void Nl2<P>::build_h() {
  int m = 0;
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) {
    double t = 1.0 / d[i];
    FMC_SOLVER_LOOP
    for (int j = 0; j <= i; ++j) {
      hmat[m] = t * (G[m] + S[m]) / d[j];
      ++m;
    }
  }
}

// temp1 = hessian * step for the gradient tests (toms573.f90:1297-1326).
template <int P>
FMC_HD void Nl2<P>::hess_times_step() {
  if (stpmod == 2) {
    FMC_SOLVER_LOOP
    for (int i = 0; i < P; ++i) tmp2[i] = d[i] * step[i];
    symv(P, tmp1, hmat, tmp2);
    FMC_SOLVER_LOOP
    for (int i = 0; i < P; ++i) tmp1[i] = d[i] * tmp1[i];
    return;
  }
  // rptmul(2,...): perm * R^T * R * perm^T * step (toms573.f90:4595-4655)
  double z[P], y[P];
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) z[i] = step[ipivot[i] - 1];
#define FMC_R(i, j) rfac[(j) * ((j) + 1) / 2 + (i)]
  y[0] = z[0] * FMC_R(0, 0);
  FMC_SOLVER_LOOP
  for (int kk = 1; kk < P; ++kk) {
    double zk = z[kk];
    FMC_SOLVER_LOOP
    for (int i = 0; i < kk; ++i) y[i] += FMC_R(i, kk) * zk;
    y[kk] = zk * FMC_R(kk, kk);
  }
  z[0] = y[0] * FMC_R(0, 0);
  FMC_SOLVER_LOOP
  for (int i = 1; i < P; ++i) {
    double t = 0.0;
    FMC_SOLVER_LOOP
    for (int m = 0; m < i; ++m) t += FMC_R(m, i) * y[m];
    z[i] = y[i] * FMC_R(i, i) + t;
  }
#undef FMC_R
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) tmp1[ipivot[i] - 1] = z[i];
}

// slupdt (toms573.f90:4658-4694): S <- size*S + u w^T + w u^T so that S*step = y (= lky).
template <int P>
FMC_HD void Nl2<P>::slupdt(const double* wchmtd) {
  double* u = tmp1;
  double* ww = tmp2;
  double sdotwm = dotn(P, step, wchmtd);
  double denmin = k::cosmin * nrm2(P, step) * nrm2(P, wchmtd);
  wscale_ = 1.0;
  if (denmin != 0.0) wscale_ = fmin(1.0, fabs(sdotwm / denmin));
  double t = 0.0;
  if (sdotwm != 0.0) t = wscale_ / sdotwm;
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) ww[i] = t * wchmtd[i];
  symv(P, u, S, step);
  t = 0.5 * (size_ * dotn(P, step, u) - dotn(P, step, lky));
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) u[i] = t * ww[i] + lky[i] - size_ * u[i];
  int m = 0;
  FMC_SOLVER_LOOP
  for (int i = 0; i < P; ++i) {
    double ui = u[i], wi = ww[i];
    FMC_SOLVER_LOOP
    for (int j = 0; j <= i; ++j) {
      S[m] = size_ * S[m] + ui * ww[j] + wi * u[j];
      ++m;
    }
  }
}

// ----------------------------------------------------------------- assess
// toms573.f90:1447-1952.
template <int P>
FMC_HD void Nl2<P>::assess() {
  bool goodx = true;
  int i, nfc = nfcall;
  double emax = 0.0, gts, reldx1, rfac1 = 1.0, xmax;
  swtch = 0;
  restor = 0;
  i = irc;
  switch (i) {
    case 1: goto L20;
    case 2: goto L30;
    case 3: case 4: goto L10;
    case 5: goto L40;
    case 6: goto L280;
    case 7: case 8: case 9: case 10: case 11: goto L220;
    case 12: goto L170;
    default: irc = 13; return;
  }
L10:
  stage = 1;
  radinc = 0;
  v[FLSTGD] = v[F0];
  if (toobig == 0) goto L90;
  stage = -1;
  xirc = i;
  goto L60;
L20:
  if (model != mlstgd) goto L30;
  stage = k::stglim;
  radinc = -1;
  goto L90;
L30:
  stage = stage + 1;
L40:
  if (stage > 0) goto L50;
  if (toobig != 0) goto L60;
  stage = -stage;
  i = xirc;
  switch (i) {
    case 1: goto L20;
    case 2: goto L30;
    case 3: case 4: goto L90;
    case 5: goto L70;
    default: break;
  }
L50:
  if (toobig == 0) goto L70;
  if (radinc > 0) goto L80;
  stage = -stage;
  xirc = irc;
L60:
  v[RADFAC] = k::decfac;
  radinc = radinc - 1;
  irc = 5;
  return;
L70:
  if (v[F] < v[FLSTGD]) goto L90;
  if (model == mlstgd) goto L80;
  model = mlstgd;
  swtch = 1;
L80:
  if (v[FLSTGD] >= v[F0]) goto L90;
  restor = 1;
  v[F] = v[FLSTGD];
  v[PREDUC] = v[PLSTGD];
  v[GTSTEP] = v[GTSLST];
  if (swtch == 0) rfac1 = v[DSTNRM] / v[DSTSAV];
  v[DSTNRM] = v[DSTSAV];
  nfc = nfgcal;
  goodx = false;
L90: {
  // reldst (toms573.f90:4570-4592)
  double em = 0.0, xm = 0.0;
  FMC_SOLVER_LOOP
  for (int m = 0; m < P; ++m) {
    double t = fabs(d[m] * (x[m] - x0[m]));
    if (em < t) em = t;
    t = d[m] * (fabs(x[m]) + fabs(x0[m]));
    if (xm < t) xm = t;
  }
  reldx1 = 0.0;
  if (xm > 0.0) reldx1 = em / xm;
}
  if (goodx) goto L110;
  FMC_SOLVER_LOOP
  for (int m = 0; m < P; ++m) {
    step[m] = stlstg[m];
    x[m] = x0[m] + stlstg[m];
  }
  x_is_eval = 0;  // x is an earlier trial point again: its Jacobian is not the one just evaluated
L110:
  v[FDIF] = v[F0] - v[F];
  if (v[FDIF] > k::tuner2 * v[PREDUC]) goto L140;
  v[RELDX] = reldx1;
  if (v[F] < v[F0]) goto L120;
  mlstgd = model;
  v[FLSTGD] = v[F];
  v[F] = v[F0];
  FMC_SOLVER_LOOP
  for (int m = 0; m < P; ++m) x[m] = x0[m];
  x_is_eval = 0;
  restor = 1;
  goto L130;
L120:
  nfgcal = nfc;
L130:
  irc = 1;
  if (stage < k::stglim) goto L160;
  irc = 5;
  radinc = radinc - 1;
  goto L160;
L140:
  nfgcal = nfc;
  rfac1 = 1.0;
  if (goodx) v[RELDX] = reldx1;
  v[DSTSAV] = v[DSTNRM];
  if (v[FDIF] > v[PREDUC] * k::tuner1) goto L190;
  if (stage >= k::stglim) goto L150;
  irc = 2;
  goto L160;
L150:
  irc = 4;
L160:
  xirc = irc;
  emax = v[GTSTEP] + v[FDIF];
  v[RADFAC] = 0.5 * rfac1;
  if (emax < v[GTSTEP]) v[RADFAC] = rfac1 * fmax(k::rdfcmn, 0.5 * v[GTSTEP] / emax);
L170:
  if (v[RELDX] <= k::xftol) goto L180;
  irc = xirc;
  if (v[F] < v[F0]) goto L200;
  goto L230;
L180:
  irc = 12;
  goto L240;
L190:
  if (v[FDIF] < (-k::tuner3 * v[GTSTEP])) goto L210;
  if (radinc < 0) goto L210;
  if (restor == 1) goto L210;
  v[RADFAC] = k::rdfcmx;
  gts = v[GTSTEP];
  if (v[FDIF] < (0.5 / v[RADFAC] - 1.0) * gts) v[RADFAC] = fmax(k::incfac, 0.5 * gts / (gts + v[FDIF]));
  irc = 4;
  if (v[STPPAR] == 0.0) goto L230;
  irc = 5;
  radinc = radinc + 1;
L200:
  v[FLSTGD] = v[F];
  mlstgd = model;
  FMC_SOLVER_LOOP
  for (int m = 0; m < P; ++m) stlstg[m] = step[m];
  v[DSTSAV] = v[DSTNRM];
  nfgcal = nfc;
  v[PLSTGD] = v[PREDUC];
  v[GTSLST] = v[GTSTEP];
  goto L230;
L210:
  v[RADFAC] = 1.0;
  irc = 3;
  goto L230;
L220:
  irc = xirc;
  if (v[DSTSAV] >= 0.0) goto L240;
  irc = 12;
  goto L240;
L230:
  xirc = irc;
L240:
  if (fabs(v[F]) < k::afctol) irc = 10;
  if (0.5 * v[FDIF] > v[PREDUC]) return;
  emax = k::rfctol * fabs(v[F0]);
  if (v[DSTNRM] > k::lmax0 && v[PREDUC] <= emax) irc = 11;
  if (v[DST0] < 0.0) goto L250;
  i = 0;
  if ((v[NREDUC] > 0.0 && v[NREDUC] <= emax) || (v[NREDUC] == 0.0 && v[PREDUC] == 0.0)) i = 2;
  if (v[STPPAR] == 0.0 && v[RELDX] <= k::xctol && goodx) i = i + 1;
  if (i > 0) irc = i + 6;
L250:
  if (abs(irc - 3) > 2 && irc != 12) return;
  if (v[DSTNRM] > k::lmax0) goto L260;
  if (v[PREDUC] >= emax) return;
  if (v[DST0] <= 0.0) goto L270;
  if (0.5 * v[DST0] <= k::lmax0) return;
  goto L270;
L260:
  if (0.5 * v[DSTNRM] <= k::lmax0) return;
  xmax = k::lmax0 / v[DSTNRM];
  if (xmax * (2.0 - xmax) * v[PREDUC] >= emax) return;
L270:
  if (v[NREDUC] < 0.0) goto L290;
  v[GTSLST] = v[GTSTEP];
  v[DSTSAV] = v[DSTNRM];
  if (irc == 12) v[DSTSAV] = -v[DSTSAV];
  v[PLSTGD] = v[PREDUC];
  irc = 6;
  FMC_SOLVER_LOOP
  for (int m = 0; m < P; ++m) stlstg[m] = step[m];
  return;
L280:
  v[GTSTEP] = v[GTSLST];
  v[DSTNRM] = fabs(v[DSTSAV]);
  FMC_SOLVER_LOOP
  for (int m = 0; m < P; ++m) step[m] = stlstg[m];
  irc = xirc;
  if (v[DSTSAV] <= 0.0) irc = 12;
  v[NREDUC] = -v[PREDUC];
  v[PREDUC] = v[PLSTGD];
L290:
  if (-v[NREDUC] <= k::rfctol * fabs(v[F0])) irc = 11;
}

// ----------------------------------------------------------------- lmstep
// Levenberg-Marquardt step by the More-Hebden technique with fast Givens updates of
// [R; sqrt(alpha) D] (toms573.f90:3397-3839).  1-based aliases as in the reference.
template <int P>
FMC_HD void Nl2<P>::lmstep() {
  const double* dd = d - 1;
  const double* gg = g - 1;
  const int* ipiv = ipivot - 1;
  double* st = step - 1;
  double* ww = w - 1;
  const int p = P;
  int& ka = kalm;
  int dstsav, i, ip1, i1, j1, kk, kalim, l, lk0, phipin, pp1o2, res, res0, rmat, rmat0, uk0;
  double a, adi, alphak, b, dfacsq, dst, dtol, d1, d2, lk, oldphi, phi, phimax, phimin, psifac, rad,
      si, sj, sqrtak, t, twopsi, uk, wl;
  const double dfac = 256.0, ttol = 2.5, p001 = 1e-3;
  int nloop = 0;

  psifac = 0.0;
  alphak = 0.0;
  dst = 0.0;
  phi = 0.0;
  lk0 = p + 1;
  phipin = lk0 + 1;
  uk0 = phipin + 1;
  dstsav = uk0 + 1;
  rmat0 = dstsav;
  rmat = rmat0 + 1;
  pp1o2 = PP;
  res0 = pp1o2 + rmat0;
  res = res0 + 1;
  rad = v[RADIUS];
  if (rad > 0.0) psifac = k::epslon / ((8.0 * (k::phmnfc + 1.0) + 3.0) * (rad * rad));
  phimax = k::phmxfc * rad;
  phimin = k::phmnfc * rad;
  dtol = 1.0 / dfac;
  dfacsq = dfac * dfac;
  oldphi = 0.0;
  lk = 0.0;
  uk = 0.0;
  kalim = ka + 12;

  if (ka < 0) goto L10;
  if (ka == 0) goto L20;
  goto L310;

L10:
  ka = 0;
  kalim = 12;
  kk = p;
  if (ierr != 0) kk = abs(ierr) - 1;
  v[NREDUC] = 0.5 * dotn(kk, qtr, qtr);
L20:
  v[DST0] = -1.0;
  if (ierr != 0) goto L50;
  ltsolve_t(p, w, hmat, qtr);
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j1 = ipiv[i];
    st[i] = dd[j1] * ww[i];
  }
  dst = nrm2(p, step);
  v[DST0] = dst;
  phi = dst - rad;
  if (phi <= phimax) goto L350;
  if (ka > 0) goto L70;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j1 = ipiv[i];
    st[i] = dd[j1] * (st[i] / dst);
  }
  ltsolve(p, step, hmat, step);
  t = 1.0 / nrm2(p, step);
  ww[phipin] = (t / dst) * t;
  lk = phi * ww[phipin];
L50:
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) ww[i] = gg[i] / dd[i];
  v[DGNORM] = nrm2(p, w);
  uk = v[DGNORM] / rad;
  if (uk <= 0.0) goto L330;
  alphak = fabs(v[STPPAR]) * v[RAD0] / rad;
L70:
  if (++nloop > 100) {  // safety cap (the reference relies on kalim = ka + 12)
    anomaly |= 1;
    goto L370;
  }
  ka = ka + 1;
  FMC_RARE_LOOP
  for (i = 0; i < pp1o2; ++i) ww[rmat + i] = hmat[i];
  FMC_RARE_LOOP
  for (i = 0; i < p; ++i) ww[res + i] = qtr[i];
  if (alphak <= 0.0 || alphak < lk || alphak >= uk) alphak = uk * fmax(p001, sqrt(lk / uk));
  sqrtak = sqrt(alphak);
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) ww[i] = 1.0;

  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    l = i * (i + 1) / 2 + rmat0;
    wl = ww[l];
    d2 = 1.0;
    d1 = ww[i];
    j1 = ipiv[i];
    adi = sqrtak * dd[j1];
    if (adi >= fabs(wl)) goto L110;
  L90:
    a = adi / wl;
    b = d2 * a / d1;
    t = a * b + 1.0;
    if (t > ttol) goto L110;
    ww[i] = d1 / t;
    d2 = d2 / t;
    ww[l] = t * wl;
    a = -a;
    FMC_RARE_LOOP
    for (j1 = i; j1 <= p; ++j1) {
      l = l + j1;
      st[j1] = a * ww[l];
    }
    goto L130;
  L110:
    b = wl / adi;
    a = d1 * b / d2;
    t = a * b + 1.0;
    if (t > ttol) goto L90;
    ww[i] = d2 / t;
    d2 = d1 / t;
    ww[l] = t * adi;
    FMC_RARE_LOOP
    for (j1 = i; j1 <= p; ++j1) {
      l = l + j1;
      wl = ww[l];
      st[j1] = -wl;
      ww[l] = a * wl;
    }
  L130:
    if (i == p) goto L240;
    ip1 = i + 1;
    FMC_RARE_LOOP
    for (i1 = ip1; i1 <= p; ++i1) {
      l = i1 * (i1 + 1) / 2 + rmat0;
      wl = ww[l];
      si = st[i1 - 1];
      d1 = ww[i1];
      if (d1 >= dtol) goto L150;
      d1 = d1 * dfacsq;
      wl = wl / dfac;
      kk = l;
      FMC_RARE_LOOP
      for (j1 = i1; j1 <= p; ++j1) {
        kk = kk + j1;
        ww[kk] = ww[kk] / dfac;
      }
    L150:
      if (fabs(si) > fabs(wl)) goto L180;
      if (si == 0.0) continue;
    L160:
      a = si / wl;
      b = d2 * a / d1;
      t = a * b + 1.0;
      if (t > ttol) goto L180;
      ww[l] = t * wl;
      ww[i1] = d1 / t;
      d2 = d2 / t;
      FMC_RARE_LOOP
      for (j1 = i1; j1 <= p; ++j1) {
        l = l + j1;
        wl = ww[l];
        sj = st[j1];
        ww[l] = wl + b * sj;
        st[j1] = sj - a * wl;
      }
      goto L200;
    L180:
      b = wl / si;
      a = d1 * b / d2;
      t = a * b + 1.0;
      if (t > ttol) goto L160;
      ww[i1] = d2 / t;
      d2 = d1 / t;
      ww[l] = t * si;
      FMC_RARE_LOOP
      for (j1 = i1; j1 <= p; ++j1) {
        l = l + j1;
        wl = ww[l];
        sj = st[j1];
        ww[l] = a * wl + sj;
        st[j1] = b * sj - wl;
      }
    L200:
      if (d2 >= dtol) continue;
      d2 = d2 * dfacsq;
      FMC_RARE_LOOP
      for (kk = i1; kk <= p; ++kk) st[kk] = st[kk] / dfac;
    }
  }
L240:
  ltsolve_t_r(p, &ww[res], &ww[rmat], &ww[res]);
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j1 = ipiv[i];
    kk = res0 + i;
    t = ww[kk];
    st[j1] = -t;
    ww[kk] = t * dd[j1];
  }
  dst = nrm2_r(p, &ww[res]);
  phi = dst - rad;
  if (phi <= phimax && phi >= phimin) goto L370;
  if (oldphi == phi) goto L370;
  oldphi = phi;
  if (phi > 0.0) goto L270;
  if (ka >= kalim) goto L370;
  twopsi = alphak * dst * dst - dotn_r(p, step, g);
  if (alphak >= twopsi * psifac) goto L270;
  v[STPPAR] = -alphak;
  goto L380;
L260:
  if (phi < 0.0) uk = fmin(uk, alphak);
  goto L280;
L270:
  if (phi < 0.0) uk = alphak;
L280:
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j1 = ipiv[i];
    kk = res0 + i;
    st[i] = dd[j1] * (ww[kk] / dst);
  }
  ltsolve_r(p, step, &ww[rmat], step);
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) st[i] = st[i] / sqrt(ww[i]);
  t = 1.0 / nrm2_r(p, step);
  alphak = alphak + t * phi * t / rad;
  lk = fmax(lk, alphak);
  goto L70;
L310:
  lk = ww[lk0];
  uk = ww[uk0];
  if (v[DST0] > 0.0 && v[DST0] - rad <= phimax) goto L20;
  alphak = fabs(v[STPPAR]);
  dst = ww[dstsav];
  phi = dst - rad;
  t = v[DGNORM] / rad;
  if (rad > v[RAD0]) goto L320;
  uk = t;
  if (alphak <= 0.0) lk = 0.0;
  if (v[DST0] > 0.0) lk = fmax(lk, (v[DST0] - rad) * ww[phipin]);
  goto L260;
L320:
  if (alphak <= 0.0 || uk > t) uk = t;
  lk = 0.0;
  if (v[DST0] > 0.0) lk = fmax(lk, (v[DST0] - rad) * ww[phipin]);
  goto L260;
L330:
  v[STPPAR] = 0.0;
  dst = 0.0;
  lk = 0.0;
  uk = 0.0;
  v[GTSTEP] = 0.0;
  v[PREDUC] = 0.0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) st[i] = 0.0;
  goto L390;
L350:
  alphak = 0.0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j1 = ipiv[i];
    st[j1] = -ww[i];
  }
L370:
  v[STPPAR] = alphak;
L380:
  v[GTSTEP] = dotn(p, step, g);
  v[PREDUC] = 0.5 * (alphak * dst * dst - v[GTSTEP]);
L390:
  v[DSTNRM] = dst;
  ww[dstsav] = dst;
  ww[lk0] = lk;
  ww[uk0] = uk;
  v[RAD0] = rad;
}

// lsvmin (toms573.f90:3896-4047): over-estimate of the smallest singular value of L.
template <int P>
FMC_HD FMC_RARE_FN double Nl2<P>::lsvmin(const double* l_, double* x_, double* y_) {
  const double* l = l_ - 1;
  double* xv = x_ - 1;
  double* yv = y_ - 1;
  const int p = P;
  int ii = 0;
  FMC_RARE_LOOP
  for (int i = 1; i <= p; ++i) {
    xv[i] = 0.0;
    ii += i;
    if (l[ii] == 0.0) return 0.0;
  }
  if (lsvmin_ix % 9973 == 0) lsvmin_ix = 2;
  FMC_RARE_LOOP
  for (int j = p; j >= 1; --j) {
    lsvmin_ix = (3432 * lsvmin_ix) % 9973;
    double b = 0.5 * (1.0 + (double)lsvmin_ix / 9973.0);
    double xplus = (b - xv[j]);
    double xminus = (-b - xv[j]);
    double splus = fabs(xplus);
    double sminus = fabs(xminus);
    int jm1 = j - 1;
    int j0 = j * jm1 / 2;
    int jj = j0 + j;
    xplus = xplus / l[jj];
    xminus = xminus / l[jj];
    FMC_RARE_LOOP
    for (int i = 1; i <= jm1; ++i) {
      splus += fabs(xv[i] + l[j0 + i] * xplus);
      sminus += fabs(xv[i] + l[j0 + i] * xminus);
    }
    if (sminus > splus) xplus = xminus;
    xv[j] = xplus;
    FMC_RARE_LOOP
    for (int i = 1; i <= jm1; ++i) xv[i] += l[j0 + i] * xplus;
  }
  double t = 1.0 / nrm2_r(p, x_);
  FMC_RARE_LOOP
  for (int i = 1; i <= p; ++i) xv[i] = t * xv[i];
  FMC_RARE_LOOP
  for (int j = 1; j <= p; ++j) {
    double psj = 0.0;
    int j0 = j * (j - 1) / 2;
    FMC_RARE_LOOP
    for (int i = 1; i <= j - 1; ++i) psj += l[j0 + i] * yv[i];
    yv[j] = (xv[j] - psj) / l[j0 + j];
  }
  return 1.0 / nrm2_r(p, y_);
}

// ----------------------------------------------------------------- gqtstp
// Goldfeld-Quandt-Trotter step for the augmented model (toms573.f90:2471-3039).
template <int P>
FMC_HD FMC_RARE_FN void Nl2<P>::gqtstp() {
  const double* dd = d - 1;
  const double* dg = dig - 1;
  double* dihdi = hmat - 1;
  double* lmat = w + (4 * P + 7);
  double* l = lmat - 1;
  double* st = step - 1;
  double* ww = w - 1;
  const int p = P;
  int& ka = kagqt;
  int dggdmx, diag, diag0, dstsav, emax, emin, i, im1, inc, irc_, j, kk, kalim, k1, lk0, phipin, q, q0,
      uk0, xx, x00;
  double alphak, aki, akk, delta, dst, epso6, lk, oldphi, phi, phimax, phimin, psifac, rad, root, si, sk,
      sw, t, twopsi, t1, uk, wi;
  bool restrt;
  const double epsfac = 50.0, kappa = 2.0, p001 = 1e-3;
  const double dgxfac = epsfac * k::machep;
  int nloop = 0;
  bool have_q = kagqt >= 0;  // w(q) holds a solved step (of an earlier call for this H, or of this one)

  dst = 0.0;
  phi = 0.0;
  alphak = 0.0;
  uk = 0.0;
  lk = 0.0;
  delta = 0.0;
  twopsi = 0.0;
  dggdmx = p + 1;
  emax = dggdmx + 1;
  emin = emax + 1;
  lk0 = emin + 1;
  phipin = lk0 + 1;
  uk0 = phipin + 1;
  dstsav = uk0 + 1;
  diag0 = dstsav;
  diag = diag0 + 1;
  q0 = diag0 + p;
  q = q0 + 1;
  rad = v[RADIUS];
  phimax = k::phmxfc * rad;
  phimin = k::phmnfc * rad;
  psifac = 2.0 * k::epslon /
           (3.0 * (4.0 * (k::phmnfc + 1.0) * (kappa + 1.0) + kappa + 2.0) * (rad * rad));
  oldphi = 0.0;
  epso6 = k::epslon / 6.0;
  irc_ = 0;
  restrt = false;
  kalim = ka + 50;

  if (ka >= 0) goto L290;
  kk = 0;
  uk = -1.0;
  ka = 0;
  kalim = 50;
  j = 0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j += i;
    ww[diag0 + i] = dihdi[j];
  }
  t1 = 0.0;
  FMC_RARE_LOOP
  for (i = 1; i <= PP; ++i) {
    t = fabs(dihdi[i]);
    if (t1 < t) t1 = t;
  }
  ww[dggdmx] = t1;
L30:
  irc_ = chol_rows_r(p, lmat, hmat);
  if (irc_ == 0) goto L50;
  j = irc_ * (irc_ + 1) / 2;
  t = l[j];
  l[j] = 1.0;
  FMC_RARE_LOOP
  for (i = 1; i <= irc_; ++i) ww[i] = 0.0;
  ww[irc_] = 1.0;
  ltsolve_t_r(irc_, w, lmat, w);
  t1 = nrm2_r(irc_, w);
  lk = -t / t1 / t1;
  v[DST0] = -lk;
  if (restrt) goto L200;
  v[NREDUC] = 0.0;
  goto L60;
L50:
  lk = 0.0;
  ltsolve_r(p, &ww[q], lmat, dig);
  v[NREDUC] = 0.5 * dotn_r(p, &ww[q], &ww[q]);
  ltsolve_t_r(p, &ww[q], lmat, &ww[q]);
  have_q = true;
  dst = nrm2_r(p, &ww[q]);
  v[DST0] = dst;
  phi = dst - rad;
  if (phi <= phimax) goto L260;
  if (restrt) goto L200;
L60:
  v[DGNORM] = nrm2_r(p, dig);
  if (v[DGNORM] == 0.0) goto L430;
  kk = 0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    wi = 0.0;
    if (i != 1) {
      im1 = i - 1;
      FMC_RARE_LOOP
      for (j = 1; j <= im1; ++j) {
        kk = kk + 1;
        t = fabs(dihdi[kk]);
        wi = wi + t;
        ww[j] = ww[j] + t;
      }
    }
    ww[i] = wi;
    kk = kk + 1;
  }
  kk = 1;
  t1 = ww[diag] - ww[1];
  FMC_RARE_LOOP
  for (i = 2; i <= p; ++i) {
    j = diag0 + i;
    t = ww[j] - ww[i];
    if (t >= t1) continue;
    t1 = t;
    kk = i;
  }
  sk = ww[kk];
  j = diag0 + kk;
  akk = ww[j];
  k1 = kk * (kk - 1) / 2 + 1;
  inc = 1;
  t = 0.0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    if (i == kk) goto L120;
    aki = fabs(dihdi[k1]);
    si = ww[i];
    j = diag0 + i;
    t1 = 0.5 * (akk - ww[j] + si - aki);
    t1 = t1 + sqrt(t1 * t1 + sk * aki);
    if (t < t1) t = t1;
    if (i < kk) goto L130;
  L120:
    inc = i;
  L130:
    k1 = k1 + inc;
  }
  ww[emin] = akk - t;
  uk = v[DGNORM] / rad - ww[emin];
  kk = 1;
  t1 = ww[diag] + ww[1];
  FMC_RARE_LOOP
  for (i = 2; i <= p; ++i) {
    j = diag0 + i;
    t = ww[j] + ww[i];
    if (t <= t1) continue;
    t1 = t;
    kk = i;
  }
  sk = ww[kk];
  j = diag0 + kk;
  akk = ww[j];
  k1 = kk * (kk - 1) / 2 + 1;
  inc = 1;
  t = 0.0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    if (i == kk) goto L170;
    aki = fabs(dihdi[k1]);
    si = ww[i];
    j = diag0 + i;
    t1 = 0.5 * (ww[j] + si - aki - akk);
    t1 = t1 + sqrt(t1 * t1 + sk * aki);
    if (t < t1) t = t1;
    if (i < kk) goto L180;
  L170:
    inc = i;
  L180:
    k1 = k1 + inc;
  }
  ww[emax] = akk + t;
  lk = fmax(lk, v[DGNORM] / rad - ww[emax]);
  alphak = fabs(v[STPPAR]) * v[RAD0] / rad;
  if (irc_ != 0) goto L200;
  ltsolve_r(p, w, lmat, &ww[q]);
  t = nrm2_r(p, w);
  ww[phipin] = dst / t / t;
  lk = fmax(lk, phi * ww[phipin]);
L200:
  if (++nloop > 150) {  // safety cap (the reference relies on kalim = ka + 50)
    anomaly |= 2;
    if (have_q) goto L270;  // keep the last step that was actually solved for ...
    goto L430;              // ... or take no step at all
  }
  ka = ka + 1;
  if (-v[DST0] >= alphak || alphak < lk || alphak >= uk) alphak = uk * fmax(p001, sqrt(lk / uk));
  kk = 0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    kk = kk + i;
    dihdi[kk] = ww[diag0 + i] + alphak;
  }
  irc_ = chol_rows_r(p, lmat, hmat);
  if (irc_ == 0) goto L230;
  j = (irc_ * (irc_ + 1)) / 2;
  t = l[j];
  l[j] = 1.0;
  FMC_RARE_LOOP
  for (i = 1; i <= irc_; ++i) ww[i] = 0.0;
  ww[irc_] = 1.0;
  ltsolve_t_r(irc_, w, lmat, w);
  t1 = nrm2_r(irc_, w);
  lk = alphak - t / t1 / t1;
  v[DST0] = -lk;
  goto L200;
L230:
  ltsolve_r(p, &ww[q], lmat, dig);
  ltsolve_t_r(p, &ww[q], lmat, &ww[q]);
  have_q = true;
  dst = nrm2_r(p, &ww[q]);
  phi = dst - rad;
  if (phi <= phimax && phi >= phimin) goto L270;
  if (phi == oldphi) goto L270;
  oldphi = phi;
  if (phi > 0.0) goto L240;
  if (v[DST0] > 0.0) goto L240;
  delta = alphak + v[DST0];
  twopsi = alphak * dst * dst + dotn_r(p, dig, &ww[q]);
  if (delta < psifac * twopsi) goto L250;
L240:
  if (ka >= kalim) goto L270;
  ltsolve_r(p, w, lmat, &ww[q]);
  t1 = nrm2_r(p, w);
  if (phi < 0.0) uk = fmin(uk, alphak);
  alphak = alphak + (phi / t1) * (dst / t1) * (dst / rad);
  lk = fmax(lk, alphak);
  goto L200;
L250:
  if (delta > dgxfac * ww[dggdmx]) goto L330;
  goto L270;
L260:
  alphak = 0.0;
L270:
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) st[i] = -ww[q0 + i] / dd[i];
  v[GTSTEP] = -dotn_r(p, dig, &ww[q]);
  v[PREDUC] = 0.5 * (fabs(alphak) * dst * dst - v[GTSTEP]);
  goto L410;
L290:
  if (v[DST0] <= 0.0 || v[DST0] - rad > phimax) goto L310;
  restrt = true;
  ka = ka + 1;
  kk = 0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    kk = kk + i;
    dihdi[kk] = ww[diag0 + i];
  }
  uk = -1.0;
  goto L30;
L310:
  if (ka == 0) goto L50;
  dst = ww[dstsav];
  alphak = fabs(v[STPPAR]);
  phi = dst - rad;
  t = v[DGNORM] / rad;
  if (rad > v[RAD0]) goto L320;
  uk = t - ww[emin];
  lk = 0.0;
  if (alphak > 0.0) lk = ww[lk0];
  lk = fmax(lk, t - ww[emax]);
  if (v[DST0] > 0.0) lk = fmax(lk, (v[DST0] - rad) * ww[phipin]);
  goto L240;
L320:
  uk = t - ww[emin];
  if (alphak > 0.0) uk = fmin(uk, ww[uk0]);
  lk = fmax(fmax(0.0, -v[DST0]), t - ww[emax]);
  if (v[DST0] > 0.0) lk = fmax(lk, (v[DST0] - rad) * ww[phipin]);
  goto L240;
L330:
  alphak = -alphak;
  x00 = q0 + p;
  xx = x00 + 1;
  delta = kappa * delta;
  t = lsvmin(lmat, &ww[xx], w);
  kk = 0;
L340:
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) ww[i] = t * ww[i];
  ltsolve_t_r(p, w, lmat, w);
  t1 = 1.0 / nrm2_r(p, w);
  t = t1 * t;
  if (t <= delta) goto L370;
  if (kk > 30) goto L270;
  kk = kk + 1;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) ww[x00 + i] = t1 * ww[i];
  ltsolve_r(p, w, lmat, &ww[xx]);
  t = 1.0 / nrm2_r(p, w);
  goto L340;
L370:
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) ww[i] = t1 * ww[i];
  sw = dotn_r(p, &ww[q], w);
  t1 = (rad + dst) * (rad - dst);
  root = sqrt(sw * sw + t1);
  if (sw < 0.0) root = -root;
  si = t1 / (sw + root);
  v[PREDUC] = 0.5 * twopsi;
  t1 = 0.0;
  t = si * (alphak * sw - 0.5 * si * (alphak + t * dotn_r(p, &ww[xx], w)));
  if (t < epso6 * twopsi) goto L390;
  v[PREDUC] = v[PREDUC] + t;
  dst = rad;
  t1 = -si;
L390:
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j = q0 + i;
    ww[j] = t1 * ww[i] - ww[j];
    st[i] = ww[j] / dd[i];
  }
  v[GTSTEP] = dotn_r(p, dig, &ww[q]);
L410:
  v[DSTNRM] = dst;
  v[STPPAR] = alphak;
  ww[lk0] = lk;
  ww[uk0] = uk;
  v[RAD0] = rad;
  ww[dstsav] = dst;
  j = 0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) {
    j += i;
    dihdi[j] = ww[diag0 + i];
  }
  return;
L430:
  v[STPPAR] = 0.0;
  v[PREDUC] = 0.0;
  v[DSTNRM] = 0.0;
  v[GTSTEP] = 0.0;
  FMC_RARE_LOOP
  for (i = 1; i <= p; ++i) st[i] = 0.0;
}

// ----------------------------------------------------------------- nl2itr, resumable
//
// Control flow.  The reference's nl2itr is one routine full of GO TOs (toms573.f90:772-1444); a
// literal restatement is correct but, on a GPU, once the lanes of a warp take different branches in
// assess they run the REST of the routine -- factorisation, Levenberg-Marquardt step, gqtstp -- in
// separate little groups, because irreducible goto flow gives the compiler no point at which to bring
// them back together (measured on C5, ncu source view: 18 lanes enter the routine together, 3.2
// reach lmstep together).  The routine is therefore written as a state machine: each block between
// two of the reference's labels is a state, a GO TO is an assignment to `st`, and the lanes of a warp
// that are inside this call together (`lanes` = their mask) repeatedly agree on the SMALLEST pending
// state and run exactly that one.  States are numbered along the normal flow, so lanes that took a
// short cut wait for the others to catch up and every expensive block is entered by all the lanes
// that need it at once.  The per-lane sequence of operations -- hence every result -- is the same as
// in the goto form; on the host (and with lanes = 0) it is a plain sequential state machine.
enum Nl2State {
  ST_ENTRY = 0,   // label 20
  ST_ASSESS = 1,  // label 300 and the dispatch on irc that follows it
  ST_UPDATE = 2,  // label 385 ... 490: gradient, secant update
  ST_L100 = 3,    // label 100/130: start of an iteration
  ST_L160 = 4,    // label 160: evaluation limit
  ST_L180 = 5,    // label 180: which step routine
  ST_LM = 6,      // factor + lmstep
  ST_GQ = 7,      // build_h + gqtstp
  ST_L290 = 8,    // label 290: the new trial point
  ST_DONE = 9
};

template <int P>
FMC_HD bool Nl2<P>::resume(unsigned int lanes) {
  int l, spins = 0;
  double e, sttsst, t, t1;
  bool more = false;
  int st = (phase == PH_JAC) ? ST_UPDATE : ST_ENTRY;
  (void)lanes;

  for (;;) {
#if defined(__CUDA_ARCH__)
    if (lanes != 0u) {
      const int run = __reduce_min_sync(lanes, st);
      if (run == ST_DONE) break;
      if (st != run) continue;
    } else if (st == ST_DONE) {
      break;
    }
#else
    if (st == ST_DONE) break;
#endif
    switch (st) {
      case ST_ENTRY:
        // label 20 (toms573.f90:949-952): f = half the sum of squares; a residual norm above
        // rlimit (or an overflowed sum) is "too big".  NaN compares false, as in the reference.
        t = sqrt(res_f);
        if (phase == PH_INIT) init_iteration_state();
        if (t > k::rlimit) toobig = 1;
        if (toobig == 0) v[F] = 0.5 * res_f;
        if (phase == PH_TRIAL) {
          x_is_eval = spec;  // (spec = 0: the pass evaluated the residual sum alone)
          st = ST_ASSESS;
          break;
        }
        if (toobig != 0) {  // label 40
          rc = 13;
          st = ST_DONE;
          break;
        }
        adopt_jacobian();  // label 60
        st = ST_L100;
        break;

      case ST_L100:
        if (niter >= k::mxiter) {
          rc = 10;
          st = ST_DONE;
          break;
        }
        {
          int kk = niter;
          niter = kk + 1;
          if (kk != 0) {
            FMC_SOLVER_LOOP
            for (int i = 0; i < P; ++i) step[i] = d[i] * step[i];
            v[RADIUS] = v[RADFAC] * nrm2(P, step);
          }
        }
        // label 130
        v[F0] = v[F];
        kagqt = -1;
        irc = 4;
        hvalid = 0;
        sused = model;
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) x0[i] = x[i];
        st = ST_L160;
        break;

      case ST_L160:
        if (nfcall >= k::mxfcal) {
          // The reference evaluates one more gradient when f < f0 (toms573.f90:1055-1061); it is
          // only used by the covariance code, so the result x is already final.
          if (v[F] < v[F0]) ngcall++;
          rc = 9;
          st = ST_DONE;
          break;
        }
        st = ST_L180;
        break;

      case ST_L180:
        if (++spins > 40) {  // safety cap: step recomputed over and over without a new evaluation
          anomaly |= 4;
          rc = 8;
          st = ST_DONE;
          break;
        }
        st = (model == 2) ? ST_GQ : ST_LM;
        break;

      case ST_LM:
        if (kalm < 0) factor();
        if (!hvalid) {
          FMC_SOLVER_LOOP
          for (int i = 0; i < PP; ++i) hmat[i] = rfac[i];
          hvalid = 1;
        }
        lmstep();
        st = ST_L290;
        break;

      case ST_GQ:  // label 220
        if (!hvalid) {
          build_h();
          hvalid = 1;
        }
        gqtstp();
        st = ST_L290;
        break;

      case ST_L290:
        if (irc == 6) {
          st = ST_ASSESS;
          break;
        }
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) x[i] = step[i] + x0[i];
        nfcall = nfcall + 1;
        toobig = 0;
        set_fd_steps();
        need_b = spec;
        need_jac = spec;
        phase = PH_TRIAL;
        more = true;
        st = ST_DONE;
        break;

      case ST_ASSESS:  // label 300
        assess();
#ifdef FMC_TRACE
        if (trace && trace_len < trace_cap) {
          double* tr = trace + (size_t)trace_len * 10;
          double piv = 0.0;
          FMC_SOLVER_LOOP
          for (int q = 0; q < P; ++q) piv = piv * P + (ipivot[q] - 1);
          tr[0] = irc;
          tr[1] = sused;
          tr[2] = nfcall;
          tr[3] = niter;
          tr[4] = v[F];
          tr[5] = v[RELDX];
          tr[6] = v[RADIUS];
          tr[7] = piv;
          tr[8] = restor;
          tr[9] = v[NREDUC];
          ++trace_len;
        }
#endif
        if (swtch != 0) {
          hvalid = 0;
          sused = sused + 2;
          FMC_SOLVER_LOOP
          for (int i = 0; i < 9; ++i) v[i] = vsave[i];
        }
        l = irc - 4;
        stpmod = model;
        if (l > 0) {
          if (l == 1) {  // label 340
            v[RADIUS] = v[RADFAC] * v[DSTNRM];
            st = ST_L160;
            break;
          }
          if (l == 2) {  // label 360
            v[RADIUS] = k::lmax0;
            st = ST_L180;
            break;
          }
          if (l <= 8) {  // label 370
            // convergence or false convergence: x is final.  (The reference still fetches a gradient
            // when f < f0, for the covariance matrix only: toms573.f90:1259-1262.)
            if (v[F] < v[F0] && xirc != 14) ngcall++;
            rc = l;
            st = ST_DONE;
            break;
          }
          rc = 14;  // l == 9: bad parameters to assess (cannot occur)
          st = ST_DONE;
          break;
        }
        {
          // decide whether to change models (toms573.f90:1215-1235)
          e = v[PREDUC] - v[FDIF];
          symv(P, lky, S, step);
          sttsst = 0.5 * dotn(P, step, lky);
          if (model == 1) sttsst = -sttsst;
          if (fabs(e + sttsst) * k::fuzz >= fabs(e)) {  // label 330
            if (!(-3 < l)) {
              v[RADIUS] = v[RADFAC] * v[DSTNRM];
              st = ST_L160;
              break;
            }
          } else {
            model = 3 - model;
            if (model == 1) kagqt = -1;
            if (model == 2 && kalm > 0) kalm = 0;
            if (!(-2 < l)) {
              hvalid = 0;
              sused = sused + 2;
              FMC_SOLVER_LOOP
              for (int i = 0; i < 9; ++i) vsave[i] = v[i];
              st = ST_L160;  // label 350 only saves r, which is not kept here
              break;
            }
          }
        }
        // label 380
        if (!x_is_eval) {
          // assess went back to an earlier trial point of this iteration: its Jacobian and
          // J(x0)^T r(x) need a pass of their own.
          set_fd_steps();
          need_b = 1;
          need_jac = 1;
          phase = PH_JAC;
          more = true;
          st = ST_DONE;
          break;
        }
        st = ST_UPDATE;
        break;

      case ST_UPDATE:  // label 385
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) lky[i] = res_lky[i];  // lky = J(x0)^T r(x), toms573.f90:1268-1293
        if (irc == 3) hess_times_step();                  // label 410
        ngcall = ngcall + 1;                              // label 450
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) w[i] = g[i];          // g0
        adopt_jacobian();                                 // labels 50-60 at the new x
        // label 460
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) w[i] = -w[i] + g[i];  // g - g0
        if (irc == 3) {
          FMC_SOLVER_LOOP
          for (int i = 0; i < P; ++i) tmp1[i] = (tmp1[i] - w[i]) / d[i];
          if (nrm2(P, tmp1) <= v[DGNORM] * k::tuner4) {
            v[RADFAC] = k::incfac;
          } else if (dotn(P, g, step) >= v[GTSTEP] * k::tuner5) {
          } else {
            v[RADFAC] = k::incfac;
          }
        }
        // label 490: y = (J(x) - J(x0))^T r(x), sizing factor, secant update of S
        FMC_SOLVER_LOOP
        for (int i = 0; i < P; ++i) lky[i] = -lky[i] + g[i];
        symv(P, tmp1, S, step);
        t1 = fabs(dotn(P, step, tmp1));
        t = fabs(dotn(P, step, lky));
        size_ = 1.0;
        if (t < t1) size_ = t / t1;
        slupdt(w);
        st = ST_L100;
        break;

      default:
        st = ST_DONE;
        break;
    }
  }
  return more;
}

}  // namespace fmc
