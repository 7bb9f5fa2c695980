// Dual numbers: the analytic Jacobian without a hand-written `madj`.
//
// The reference tells users who want analytic derivatives to "introduce a subroutine madj for
// evaluating the Jacobian matrix and replace the call to nl2sno with a call to nl2sol"
// (README:35-42; nl2sol is toms573.f90:35-587).  Here the same generic model_y that serves the
// forward-difference path (pert.cuh) is evaluated on `Dual` values -- a quantity v together
// with its EXACT partial derivative d with respect to the one parameter being differentiated --
// so the Jacobian row d model_y / d params_k comes out of the model formula itself by the chain
// rule.  A plug-in may instead supply its own
//     static void model_grad(const double* params, double x, double* dydp)
// (the madj of the README), which then takes precedence (pass_eval.cuh).
//
// As for Pert, `on` is a compile-time constant after unrolling, so only the data flow that
// depends on the differentiated parameter costs anything.
#pragma once
#include "fast_math.cuh"
#include "fmc_common.cuh"

namespace fmc {

struct Dual {
  double v;  // value
  double d;  // d value / d params_k
  bool on;   // false: d == 0 identically
  FMC_HD FMC_INLINE Dual() : v(0.0), d(0.0), on(false) {}
  FMC_HD FMC_INLINE Dual(double value) : v(value), d(0.0), on(false) {}
  FMC_HD FMC_INLINE Dual(double value, double deriv, bool active) : v(value), d(deriv), on(active) {}
};

FMC_HD FMC_INLINE Dual operator-(const Dual& a) { return Dual(-a.v, a.on ? -a.d : 0.0, a.on); }
FMC_HD FMC_INLINE Dual operator+(const Dual& a) { return a; }
FMC_HD FMC_INLINE Dual operator+(const Dual& a, const Dual& b) {
  const double d = (a.on && b.on) ? a.d + b.d : (a.on ? a.d : (b.on ? b.d : 0.0));
  return Dual(a.v + b.v, d, a.on || b.on);
}
FMC_HD FMC_INLINE Dual operator-(const Dual& a, const Dual& b) {
  const double d = (a.on && b.on) ? a.d - b.d : (a.on ? a.d : (b.on ? -b.d : 0.0));
  return Dual(a.v - b.v, d, a.on || b.on);
}
FMC_HD FMC_INLINE Dual operator*(const Dual& a, const Dual& b) {
  double d = 0.0;
  if (a.on && b.on)
    d = fma(a.v, b.d, b.v * a.d);
  else if (a.on)
    d = a.d * b.v;
  else if (b.on)
    d = a.v * b.d;
  return Dual(a.v * b.v, d, a.on || b.on);
}
FMC_HD FMC_INLINE Dual operator/(const Dual& a, const Dual& b) {
  const double q = a.v / b.v;
  double d = 0.0;
  if (b.on)
    d = ((a.on ? a.d : 0.0) - q * b.d) / b.v;
  else if (a.on)
    d = a.d / b.v;
  return Dual(q, d, a.on || b.on);
}
FMC_HD FMC_INLINE Dual operator+(const Dual& a, double b) { return Dual(a.v + b, a.d, a.on); }
FMC_HD FMC_INLINE Dual operator+(double a, const Dual& b) { return Dual(a + b.v, b.d, b.on); }
FMC_HD FMC_INLINE Dual operator-(const Dual& a, double b) { return Dual(a.v - b, a.d, a.on); }
FMC_HD FMC_INLINE Dual operator-(double a, const Dual& b) { return Dual(a - b.v, b.on ? -b.d : 0.0, b.on); }
FMC_HD FMC_INLINE Dual operator*(const Dual& a, double b) { return Dual(a.v * b, a.on ? a.d * b : 0.0, a.on); }
FMC_HD FMC_INLINE Dual operator*(double a, const Dual& b) { return Dual(a * b.v, b.on ? a * b.d : 0.0, b.on); }
FMC_HD FMC_INLINE Dual operator/(const Dual& a, double b) { return Dual(a.v / b, a.on ? a.d / b : 0.0, a.on); }
FMC_HD FMC_INLINE Dual operator/(double a, const Dual& b) { return Dual(a) / b; }

FMC_HD FMC_INLINE Dual exp(const Dual& a) {
  const double e = exp(a.v);  // fmc::exp(double)
  return Dual(e, a.on ? e * a.d : 0.0, a.on);
}
FMC_HD FMC_INLINE Dual log(const Dual& a) { return Dual(::log(a.v), a.on ? a.d / a.v : 0.0, a.on); }
FMC_HD FMC_INLINE Dual sqrt(const Dual& a) {
  const double s = ::sqrt(a.v);
  return Dual(s, a.on ? 0.5 * a.d / s : 0.0, a.on);
}
FMC_HD FMC_INLINE Dual sin(const Dual& a) {
  if (!a.on) return Dual(::sin(a.v));
  return Dual(::sin(a.v), ::cos(a.v) * a.d, true);
}
FMC_HD FMC_INLINE Dual cos(const Dual& a) {
  if (!a.on) return Dual(::cos(a.v));
  return Dual(::cos(a.v), -(::sin(a.v) * a.d), true);
}
// d(a^b) = a^b (b' ln a + b a'/a)
FMC_HD FMC_INLINE Dual pow(const Dual& a, const Dual& b) {
  const double p = ::pow(a.v, b.v);
  if (!a.on && !b.on) return Dual(p);
  double t = 0.0;
  if (b.on) t = b.d * ::log(a.v);
  if (a.on) t = fma(b.v, a.d / a.v, t);
  return Dual(p, p * t, true);
}
FMC_HD FMC_INLINE Dual pow(double a, const Dual& b) { return pow(Dual(a), b); }
FMC_HD FMC_INLINE Dual pow(const Dual& a, double b) { return pow(a, Dual(b)); }

}  // namespace fmc
