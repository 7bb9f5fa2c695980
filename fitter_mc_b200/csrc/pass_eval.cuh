// One row of a streaming "pass" over the data: residual and forward-difference Jacobian row
// at an evaluation point, and the p x p sums the solver is fed with (see nl2_solver.cuh).
//
// Replaces, per data point i:
//   madr             r_i = (y'_i - model_y(params, x_i)) / err_i          bootstrap.f90:246-248
//   nl2sno label 30  J_ik = (r_i(x + h_k e_k) - r_i(x)) / h_k             toms573.f90:714-733
//   nl2itr label 60  g = J^T r (dotprd per column)                        toms573.f90:975-978
//   dupdat/qrfact    column norms and J^T J                               toms573.f90:2458, 4360-4567
//   nl2itr label 380 lky = J(x0)^T r(x)                                   toms573.f90:1268-1293
// Nothing of length n is written: the row is consumed as soon as it is formed.
#pragma once
#include "fmc_common.cuh"

namespace fmc {

// An evaluation point with its finite-difference steps, held in registers for a whole pass.
template <int P>
struct PassPoint {
  double x[P];   // parameters
  double xh[P];  // x[k] + h[k]: the k-th perturbed coordinate (toms573.f90:718)
  FMC_HD FMC_INLINE void set(const double* xs, const double* h) {
#pragma unroll
    for (int kk = 0; kk < P; ++kk) {
      x[kk] = xs[kk];
      xh[kk] = xs[kk] + h[kk];
    }
  }
};

template <int P>
struct PassAcc {
  static constexpr int PP = P * (P + 1) / 2;
  double f;
  double g[P];
  double G[PP];
  double lky[P];
  FMC_HD FMC_INLINE void zero() {
    f = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      g[i] = 0.0;
      lky[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < PP; ++i) G[i] = 0.0;
  }
};

// Residual and UNSCALED Jacobian row at `pt` for the data point (xi, yfit, 1/err):
// D[k] = r(x + h_k e_k) - r(x).  The division by h_k of toms573.f90:729-732 is a constant
// factor per column, so it is applied once per pass to the p x p sums (pass_finish) instead
// of n times per column -- same value up to one rounding per term, 2p fewer FP64 ops per row.
// Because the
// model is inlined p+1 times with all but one coordinate identical, common subexpressions
// (e.g. exp(-al*x) of the default model for every column except al's) are evaluated once.
template <class M>
FMC_HD FMC_INLINE void row_eval(const PassPoint<M::NPARAM>& pt, double xi, double yfit, double rerr,
                                double& r0, double (&D)[M::NPARAM]) {
  constexpr int P = M::NPARAM;
  r0 = (yfit - M::model_y(pt.x, xi)) * rerr;
#pragma unroll
  for (int kk = 0; kk < P; ++kk) {
    double xs[P];
#pragma unroll
    for (int m = 0; m < P; ++m) xs[m] = (m == kk) ? pt.xh[m] : pt.x[m];
    const double rk = (yfit - M::model_y(xs, xi)) * rerr;
    D[kk] = rk - r0;
  }
}

// Add one data point to the pass sums.  WITH_B: also J(b)^T r(a) for the secant update.
template <class M, bool WITH_B>
FMC_HD FMC_INLINE void row_accumulate(const PassPoint<M::NPARAM>& a, const PassPoint<M::NPARAM>& b,
                                      double xi, double yfit, double rerr, PassAcc<M::NPARAM>& acc) {
  constexpr int P = M::NPARAM;
  double ra, Ja[P];
  row_eval<M>(a, xi, yfit, rerr, ra, Ja);
  acc.f = fma(ra, ra, acc.f);
  int m = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    acc.g[i] = fma(Ja[i], ra, acc.g[i]);
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      acc.G[m] = fma(Ja[i], Ja[j], acc.G[m]);
      ++m;
    }
  }
  if (WITH_B) {
    double rb, Jb[P];
    row_eval<M>(b, xi, yfit, rerr, rb, Jb);
#pragma unroll
    for (int i = 0; i < P; ++i) acc.lky[i] = fma(Jb[i], ra, acc.lky[i]);
  }
}

// Apply the 1/h_k column factors to the sums of a finished pass and hand them to the solver.
template <int P, class Solver>
FMC_HD FMC_INLINE void pass_finish(const PassAcc<P>& acc, Solver& s) {
  double ia[P], ib[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    ia[i] = 1.0 / s.hA[i];
    ib[i] = 1.0 / s.hJ[i];
  }
  s.res_f = acc.f;
  int m = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    s.res_g[i] = acc.g[i] * ia[i];
    s.res_lky[i] = acc.lky[i] * ib[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      s.res_G[m] = acc.G[m] * (ia[i] * ia[j]);
      ++m;
    }
  }
}

}  // namespace fmc
