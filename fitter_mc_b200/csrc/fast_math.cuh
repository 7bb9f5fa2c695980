// Branch-free FP64 exp for the model plug-ins (device side).
//
// Inside namespace fmc an unqualified `exp(x)` -- which is how the model routines in
// models/*.cuh are written, exactly like the Fortran EXP() of bootstrap.f90:395 -- resolves to
// fmc::exp below.  On the host it is std::exp.  On the device it replaces CUDA's libdevice exp
// with an equivalent (< 1 ulp on the polynomial, same range reduction) formulation chosen for
// the fit kernel's inner loop, which is bound by FP64 issue and dependent-DFMA latency:
//   * no branch (libdevice's exp has a slow-path branch => BSSY/BSYNC/BRA per call),
//   * coefficients come from the constant bank as DFMA operands (libdevice rematerialises them
//     with ~22 MOV/UMOV per call when registers are scarce),
//   * even/odd split of the polynomial: dependent chain of 12 instead of 15 FP64 ops.
// Domain: results are exact-rounded-ish for |x| <= 708; beyond that the scale factor saturates
// (exp(-800) returns ~1e-308 instead of 0, exp(+800) ~1e308 instead of inf), +-inf give NaN.
// None of that matters on this path: such values are "too big" residuals either way.
#pragma once
#include "fmc_common.cuh"

namespace fmc {

#if defined(__CUDACC__)
// q(r) = (exp(r) - 1 - r) / r^2 on |r| <= ln2/2, degree 9 (Chebyshev-node interpolant computed
// in 60-digit arithmetic; measured max error of the whole polynomial stage 0.84 ulp).
static __constant__ double kExpQ[10] = {
    5.00000000000000111e-01, 1.66666666666666685e-01, 4.16666666666241289e-02, 8.33333333333006153e-03,
    1.38888889172137167e-03, 1.98412698630536177e-04, 2.48015212959543761e-05, 2.75572684599970641e-06,
    2.76200884454097462e-07, 2.51003854955103203e-08};
static __constant__ double kExpR[3] = {1.4426950408889634, -6.93147180369123816490e-01,
                                       -1.90821492927058770002e-10};  // log2(e), -ln2_hi, -ln2_lo
#endif

FMC_HD FMC_INLINE double exp(double x) {
#if defined(__CUDA_ARCH__)
  const double t = fma(x, kExpR[0], 6755399441055744.0);  // round(x*log2e) in the low word
  int kk = __double2loint(t);
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, kExpR[1], x);
  r = fma(kd, kExpR[2], r);
  const double r2 = r * r;
  double e = fma(kExpQ[8], r2, kExpQ[6]);
  double o = fma(kExpQ[9], r2, kExpQ[7]);
  e = fma(e, r2, kExpQ[4]);
  o = fma(o, r2, kExpQ[5]);
  e = fma(e, r2, kExpQ[2]);
  o = fma(o, r2, kExpQ[3]);
  e = fma(e, r2, kExpQ[0]);
  o = fma(o, r2, kExpQ[1]);
  const double q = fma(r, o, e);
  const double p = 1.0 + fma(r2, q, r);  // in [0.70, 1.42]
  kk = max(-1021, min(kk, 1022));
  return __hiloint2double(__double2hiint(p) + (kk << 20), __double2loint(p));
#else
  return ::exp(x);
#endif
}

}  // namespace fmc
