// Perturbation arithmetic: forward differences without re-evaluating the model.
//
// nl2sno builds its Jacobian by forward differences (toms573.f90:714-733): for every
// parameter k it evaluates the residuals again at x + h_k e_k, h_k ~ 1.5e-8 |x_k|, and
// subtracts.  That costs p extra model evaluations per data point -- including the model's
// transcendental functions -- to obtain a difference whose leading digits cancel.
//
// A `Pert` value carries a quantity at the base point (v) together with its CHANGE (d) when
// the one perturbed parameter moves by its step h.  Every operation propagates d with the
// exact algebraic identity where one exists (+, -, *, /) and with a Taylor expansion in the
// tiny relative perturbation that is exact to working precision otherwise (exp, log, pow,
// sqrt, sin, cos: relative truncation error of d below 1e-16 for the step sizes NL2SOL uses;
// see expm1_small below).  The result's d is therefore
//     m(x + h_k e_k) - m(x)
// -- the very numerator of the reference's forward difference, with the SAME h and hence the
// same O(h f'') truncation error; this is not analytic differentiation -- but obtained
// without the second transcendental evaluation and without the cancellation.  It differs
// from the reference's computed numerator only by that numerator's own rounding noise.
//
// `on` says whether d can be nonzero.  After the kernel's per-parameter loop is unrolled
// and the model inlined, `on` is a compile-time constant everywhere, so operations on
// unperturbed quantities compile to their plain double form and only the data flow that
// actually depends on the perturbed parameter costs anything.
//
// A model plug-in opts in simply by being generic in its scalar type:
//     template <class T> static T model_y(const T* params, double x)
// (see models/model_user.cuh).  A model written for `double` only still works: the engine
// then falls back to p+1 plain evaluations per data point (pass_eval.cuh).
#pragma once
#include "fast_math.cuh"
#include "fmc_common.cuh"

namespace fmc {

struct Pert {
  double v;  // value at the base point
  double d;  // value(x + h_k e_k) - value(x)
  bool on;   // false: d == 0 identically
  FMC_HD FMC_INLINE Pert() : v(0.0), d(0.0), on(false) {}
  FMC_HD FMC_INLINE Pert(double value) : v(value), d(0.0), on(false) {}
  FMC_HD FMC_INLINE Pert(double value, double delta, bool active) : v(value), d(delta), on(active) {}
};

// ---- exact identities -------------------------------------------------------------------
FMC_HD FMC_INLINE Pert operator-(const Pert& a) { return Pert(-a.v, a.on ? -a.d : 0.0, a.on); }
FMC_HD FMC_INLINE Pert operator+(const Pert& a) { return a; }

FMC_HD FMC_INLINE Pert operator+(const Pert& a, const Pert& b) {
  const double d = (a.on && b.on) ? a.d + b.d : (a.on ? a.d : (b.on ? b.d : 0.0));
  return Pert(a.v + b.v, d, a.on || b.on);
}
FMC_HD FMC_INLINE Pert operator-(const Pert& a, const Pert& b) {
  const double d = (a.on && b.on) ? a.d - b.d : (a.on ? a.d : (b.on ? -b.d : 0.0));
  return Pert(a.v - b.v, d, a.on || b.on);
}
FMC_HD FMC_INLINE Pert operator*(const Pert& a, const Pert& b) {
  // (a+da)(b+db) - ab = a db + b da + da db
  double d = 0.0;
  if (a.on && b.on)
    d = fma(a.v, b.d, fma(b.v, a.d, a.d * b.d));
  else if (a.on)
    d = a.d * b.v;
  else if (b.on)
    d = a.v * b.d;
  return Pert(a.v * b.v, d, a.on || b.on);
}
FMC_HD FMC_INLINE Pert operator/(const Pert& a, const Pert& b) {
  // (a+da)/(b+db) - a/b = (da - (a/b) db) / (b + db)
  const double q = a.v / b.v;
  double d = 0.0;
  if (b.on)
    d = ((a.on ? a.d : 0.0) - q * b.d) / (b.v + b.d);
  else if (a.on)
    d = a.d / b.v;
  return Pert(q, d, a.on || b.on);
}
FMC_HD FMC_INLINE Pert operator+(const Pert& a, double b) { return Pert(a.v + b, a.d, a.on); }
FMC_HD FMC_INLINE Pert operator+(double a, const Pert& b) { return Pert(a + b.v, b.d, b.on); }
FMC_HD FMC_INLINE Pert operator-(const Pert& a, double b) { return Pert(a.v - b, a.d, a.on); }
FMC_HD FMC_INLINE Pert operator-(double a, const Pert& b) { return Pert(a - b.v, b.on ? -b.d : 0.0, b.on); }
FMC_HD FMC_INLINE Pert operator*(const Pert& a, double b) { return Pert(a.v * b, a.on ? a.d * b : 0.0, a.on); }
FMC_HD FMC_INLINE Pert operator*(double a, const Pert& b) { return Pert(a * b.v, b.on ? a * b.d : 0.0, b.on); }
FMC_HD FMC_INLINE Pert operator/(const Pert& a, double b) { return Pert(a.v / b, a.on ? a.d / b : 0.0, a.on); }
FMC_HD FMC_INLINE Pert operator/(double a, const Pert& b) { return Pert(a) / b; }

// ---- expansions in the tiny perturbation -------------------------------------------------
// expm1(t) and log1p(u) for the tiny arguments that arise here: t = (argument of the function) x
// (relative step), with NL2SOL's relative step 1.5e-8, so |t| ~ 1e-8 ... 1e-7 for arguments of
// order 1 ... 10.  Two terms, t + t^2/2, have relative truncation error t^2/6: 2e-17 at
// |t| = 1e-8 -- below FP64 rounding -- and still 2e-11 at |t| = 1e-5.  For comparison, the
// quantity being reproduced, the reference's computed numerator r(x+h) - r(x), carries a
// rounding noise of 1e-16/|t| (1e-8 ... 1e-11 over the same range) and differs from the
// derivative it stands for by t/2.  -DFMC_PERT_TERMS=4 keeps two more terms (error t^4/120);
// it costs 2 FMAs per transcendental and row (measured on C2: 3.7 % of the step) and changes
// the C2 ensemble mean by 4e-12 relative (profiles/r1_pert_terms.json).
#ifndef FMC_PERT_TERMS
#define FMC_PERT_TERMS 2
#endif
FMC_HD FMC_INLINE double expm1_small(double t) {
#if FMC_PERT_TERMS == 2
  return fma(0.5 * t, t, t);
#else
  return fma(t * t, fma(t, fma(t, 1.0 / 24.0, 1.0 / 6.0), 0.5), t);
#endif
}
FMC_HD FMC_INLINE double log1p_small(double u) {
#if FMC_PERT_TERMS == 2
  return fma(-0.5 * u, u, u);
#else
  return fma(u * u, fma(u, fma(u, -0.25, 1.0 / 3.0), -0.5), u);
#endif
}

FMC_HD FMC_INLINE Pert exp(const Pert& a) {
  const double e = exp(a.v);  // fmc::exp(double)
  return Pert(e, a.on ? e * expm1_small(a.d) : 0.0, a.on);
}
FMC_HD FMC_INLINE Pert log(const Pert& a) {
  return Pert(::log(a.v), a.on ? log1p_small(a.d / a.v) : 0.0, a.on);
}
FMC_HD FMC_INLINE Pert sqrt(const Pert& a) {
  const double s = ::sqrt(a.v);
  // sqrt(a+da) - sqrt(a) = da / (sqrt(a+da) + sqrt(a)), exact
  return Pert(s, a.on ? a.d / (::sqrt(a.v + a.d) + s) : 0.0, a.on);
}
FMC_HD FMC_INLINE Pert sin(const Pert& a) {
  double s, c;
#if defined(__CUDA_ARCH__)
  sincos(a.v, &s, &c);
#else
  s = ::sin(a.v);
  c = ::cos(a.v);
#endif
  if (!a.on) return Pert(s);
  const double t = a.d;  // sin(a+t)-sin(a) = s (cos t - 1) + c sin t
  const double cm1 = t * t * fma(t * t, 1.0 / 24.0, -0.5);
  const double st = fma(t * t, -t / 6.0, t);
  return Pert(s, fma(s, cm1, c * st), true);
}
FMC_HD FMC_INLINE Pert cos(const Pert& a) {
  double s, c;
#if defined(__CUDA_ARCH__)
  sincos(a.v, &s, &c);
#else
  s = ::sin(a.v);
  c = ::cos(a.v);
#endif
  if (!a.on) return Pert(c);
  const double t = a.d;  // cos(a+t)-cos(a) = c (cos t - 1) - s sin t
  const double cm1 = t * t * fma(t * t, 1.0 / 24.0, -0.5);
  const double st = fma(t * t, -t / 6.0, t);
  return Pert(c, fma(c, cm1, -(s * st)), true);
}
// a^b = exp(b ln a): (a+da)^(b+db) / a^b = exp(db ln a + (b+db) log1p(da/a))
FMC_HD FMC_INLINE Pert pow(const Pert& a, const Pert& b) {
  const double p = ::pow(a.v, b.v);
  if (!a.on && !b.on) return Pert(p);
  double t = 0.0;
  if (b.on) t = b.d * ::log(a.v);
  if (a.on) t = fma(b.v + (b.on ? b.d : 0.0), log1p_small(a.d / a.v), t);
  return Pert(p, p * expm1_small(t), true);
}
FMC_HD FMC_INLINE Pert pow(double a, const Pert& b) { return pow(Pert(a), b); }
FMC_HD FMC_INLINE Pert pow(const Pert& a, double b) { return pow(a, Pert(b)); }

// plain-double overloads so that a generic model body resolves the same names for T = double
FMC_HD FMC_INLINE double log(double x) { return ::log(x); }
FMC_HD FMC_INLINE double sqrt(double x) { return ::sqrt(x); }
FMC_HD FMC_INLINE double sin(double x) { return ::sin(x); }
FMC_HD FMC_INLINE double cos(double x) { return ::cos(x); }
FMC_HD FMC_INLINE double pow(double x, double y) { return ::pow(x, y); }

}  // namespace fmc
