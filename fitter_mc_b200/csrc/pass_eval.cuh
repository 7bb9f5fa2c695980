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
//
// Arithmetic form (all in FP64).  With m0 = model_y(x; x_i), mk = model_y(x + h_k e_k; x_i) and
// w = 1/err_i:
//     r_i(x + h_k e_k) - r_i(x) = (y' - mk) w - (y' - m0) w = (m0 - mk) w
// so the Jacobian row is formed from the model differences dm_k = m0 - mk, which do not
// involve the resampled y' at all.  The reference subtracts the two rounded residuals
// instead; both carry the same ~1e-8 relative forward-difference noise, from rounding in the
// model values themselves, so neither is the more accurate.  The 1/h_k column factors are
// constants of a pass and are applied once to the p x p sums (pass_finish), not n times.
// Per row this needs  (p+1) model values, p subtractions, and for the sums
//     t = w r,  f += r r,  g_k += dm_k t,  G_kl += (dm_k w^2) dm_l,  lky_k += dmB_k t.
#pragma once
#include <type_traits>
#include <utility>

#include "dual.cuh"
#include "fmc_common.cuh"
#include "pert.cuh"

namespace fmc {

// Does the plug-in provide a model_y that is generic in its scalar type (see pert.cuh)?
template <class M, class = void>
struct has_generic_model : std::false_type {};
template <class M>
struct has_generic_model<M, std::void_t<decltype(M::model_y(std::declval<const Pert*>(), 0.0))>> : std::true_type {};
#ifdef FMC_FD_PLAIN
template <class M>
constexpr bool use_pert_v = false;  // force the p+1 plain evaluations (validation builds)
#else
template <class M>
constexpr bool use_pert_v = has_generic_model<M>::value;
#endif

// How the Jacobian rows are formed.
//   JAC_FD        forward differences with nl2sno's steps (toms573.f90:714-733): the reference's path
//   JAC_ANALYTIC  exact derivatives, what nl2sol + a user madj give (toms573.f90:35-587, README:35-42):
//                 from the plug-in's own model_grad if it has one, else from the generic model_y
//                 evaluated on dual numbers (dual.cuh).  The step arrays are then all ones.
enum { JAC_FD = 0, JAC_ANALYTIC = 1 };

template <class M, class = void>
struct has_model_grad : std::false_type {};
template <class M>
struct has_model_grad<M, std::void_t<decltype(M::model_grad(std::declval<const double*>(), 0.0,
                                                            std::declval<double*>()))>> : std::true_type {};
template <class M, class = void>
struct has_dual_model : std::false_type {};
template <class M>
struct has_dual_model<M, std::void_t<decltype(M::model_y(std::declval<const Dual*>(), 0.0))>> : std::true_type {};
template <class M>
constexpr bool can_analytic_v = has_model_grad<M>::value || has_dual_model<M>::value;

// Optional per-row auxiliary values: functions of the data abscissa x_i alone that the model needs
// at every evaluation (e.g. log x_i for a power law x^(-d) = exp(-d log x_i)).  A plug-in declares
//     static constexpr int NAUX = k;
//     static void row_aux(double x, double* aux);                       // k values for one data point
//     template <class T> static T model_y(const T* params, double x, const double* aux);
// next to the reference-shaped model_y(params, x); the engine evaluates row_aux ONCE per data
// point when the data are set (fmc_set_data) and keeps the values beside the rows, so data-only
// transcendentals leave the (data points x passes x fits) loop.
template <class M, class = void>
struct model_naux : std::integral_constant<int, 0> {};
template <class M>
struct model_naux<M, std::void_t<decltype(M::NAUX)>> : std::integral_constant<int, M::NAUX> {};

template <class M, class T>
FMC_HD FMC_INLINE T eval_model(const T* params, double xi, const double* aux) {
  if constexpr (model_naux<M>::value > 0)
    return M::model_y(params, xi, aux);
  else
    return M::model_y(params, xi);
}

// An evaluation point with its finite-difference steps, held in registers for a whole pass.
template <int P>
struct PassPoint {
  double x[P];  // parameters
  double h[P];  // forward-difference steps (toms573.f90:716)
  FMC_HD FMC_INLINE void set(const double* xs, const double* hs) {
#pragma unroll
    for (int kk = 0; kk < P; ++kk) {
      x[kk] = xs[kk];
      h[kk] = hs[kk];
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

// Model value m0 = model_y(x; xi) and model differences dm[k] = m0 - model_y(x + h_k e_k; xi).
//  * generic plug-in (template <class T> model_y): ONE evaluation of the model's expensive
//    parts; the p differences come from perturbation arithmetic (pert.cuh).
//  * double-only plug-in: p+1 plain evaluations, as nl2sno does (toms573.f90:714-733).  The
//    model is inlined p+1 times with all but one coordinate identical, so the compiler still
//    shares common subexpressions (e.g. exp(-al*x) for every column except al's).
//  * JAC_ANALYTIC: dm[k] = -d model_y / d params_k (the sign convention of the differences).
//  `aux`: the row's auxiliary values (model_naux above); nullptr = evaluate them here (host images
//  and one-off kernels; the fit kernels always pass the stored values).
template <class M, int JAC = JAC_FD>
FMC_HD FMC_INLINE void row_model(const PassPoint<M::NPARAM>& pt, double xi, double& m0,
                                 double (&dm)[M::NPARAM], const double* aux = nullptr) {
  constexpr int P = M::NPARAM;
  constexpr int NAUX = model_naux<M>::value;
  double aux_here[NAUX > 0 ? NAUX : 1];
  if constexpr (NAUX > 0) {
    if (aux == nullptr) {
      M::row_aux(xi, aux_here);
      aux = aux_here;
    }
  }
  m0 = eval_model<M>(pt.x, xi, aux);
  if constexpr (JAC == JAC_ANALYTIC && has_model_grad<M>::value) {
    double gr[P];
    M::model_grad(pt.x, xi, gr);
#pragma unroll
    for (int kk = 0; kk < P; ++kk) dm[kk] = -gr[kk];
  } else if constexpr (JAC == JAC_ANALYTIC && has_dual_model<M>::value) {
#pragma unroll
    for (int kk = 0; kk < P; ++kk) {
      Dual xs[P];
#pragma unroll
      for (int m = 0; m < P; ++m) xs[m] = Dual(pt.x[m], (m == kk) ? 1.0 : 0.0, m == kk);
      const Dual mk = eval_model<M>(xs, xi, aux);
      dm[kk] = -mk.d;
    }
  } else if constexpr (use_pert_v<M>) {
#pragma unroll
    for (int kk = 0; kk < P; ++kk) {
      Pert xs[P];
#pragma unroll
      for (int m = 0; m < P; ++m) xs[m] = Pert(pt.x[m], (m == kk) ? pt.h[m] : 0.0, m == kk);
      const Pert mk = eval_model<M>(xs, xi, aux);
      dm[kk] = -mk.d;
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < P; ++kk) {
      double xs[P];
#pragma unroll
      for (int m = 0; m < P; ++m) xs[m] = (m == kk) ? pt.x[m] + pt.h[m] : pt.x[m];
      dm[kk] = m0 - eval_model<M>(xs, xi, aux);
    }
  }
}

// Add one data point to the pass sums.  WITH_B: also J(b)^T r(a) for the secant update.
template <class M, bool WITH_B, int JAC = JAC_FD>
FMC_HD FMC_INLINE void row_accumulate(const PassPoint<M::NPARAM>& a, const PassPoint<M::NPARAM>& b,
                                      double xi, double yfit, double rerr, PassAcc<M::NPARAM>& acc,
                                      const double* aux = nullptr) {
  constexpr int P = M::NPARAM;
  constexpr int NAUX = model_naux<M>::value;
  double aux_here[NAUX > 0 ? NAUX : 1];
  if constexpr (NAUX > 0) {  // (once for both points)
    if (aux == nullptr) {
      M::row_aux(xi, aux_here);
      aux = aux_here;
    }
  }
  double m0, dma[P];
  row_model<M, JAC>(a, xi, m0, dma, aux);
  const double r = (yfit - m0) * rerr;  // madr, bootstrap.f90:247
  const double t = rerr * r;
  const double w2 = rerr * rerr;
  acc.f = fma(r, r, acc.f);
  int m = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    acc.g[i] = fma(dma[i], t, acc.g[i]);
    const double dw = dma[i] * w2;
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      acc.G[m] = fma(dw, dma[j], acc.G[m]);
      ++m;
    }
  }
  if (WITH_B) {
    double mb, dmb[P];
    row_model<M, JAC>(b, xi, mb, dmb, aux);
#pragma unroll
    for (int i = 0; i < P; ++i) acc.lky[i] = fma(dmb[i], t, acc.lky[i]);
  }
}

// Apply the 1/h_k column factors to the sums of a finished pass: res = (f, g[P], G[PP], lky[P]).
template <int P>
FMC_HD FMC_INLINE void pass_scale(const PassAcc<P>& acc, const double* hA, const double* hB, double& res_f,
                                  double* res_g, double* res_G, double* res_lky) {
  double ia[P], ib[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    ia[i] = 1.0 / hA[i];
    ib[i] = 1.0 / hB[i];
  }
  res_f = acc.f;
  int m = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    res_g[i] = acc.g[i] * ia[i];
    res_lky[i] = acc.lky[i] * ib[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      res_G[m] = acc.G[m] * (ia[i] * ia[j]);
      ++m;
    }
  }
}

template <int P, class Solver>
FMC_HD FMC_INLINE void pass_finish(const PassAcc<P>& acc, Solver& s) {
  pass_scale<P>(acc, s.hA, s.hJ, s.res_f, s.res_g, s.res_G, s.res_lky);
}

// The same for an analytic pass: the sums already are J^T r and J^T J, nothing to scale.
template <int P, class Solver>
FMC_HD FMC_INLINE void pass_finish_analytic(const PassAcc<P>& acc, Solver& s) {
  s.res_f = acc.f;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    s.res_g[i] = acc.g[i];
    s.res_lky[i] = acc.lky[i];
  }
#pragma unroll
  for (int i = 0; i < P * (P + 1) / 2; ++i) s.res_G[i] = acc.G[i];
}

}  // namespace fmc
