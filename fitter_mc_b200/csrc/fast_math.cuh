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
// Domain: results are exact-rounded-ish for |x| <= 708; beyond that the result SATURATES for every
// finite or infinite argument (exp(-800), exp(-1e300), exp(-inf) return ~1e-308 instead of 0;
// exp(+800), exp(+inf) ~1e308 instead of inf); NaN gives NaN.  On this path such values are
// "too big" residuals (large side) or negligible terms (small side) either way.
#pragma once
#include <cstring>

#include "fmc_common.cuh"

namespace fmc {

// q(r) = (exp(r) - 1 - r) / r^2 on |r| <= ln2/2, degree 9 (Chebyshev-node interpolant computed
// in 60-digit arithmetic; measured max error of the whole polynomial stage 0.84 ulp).
#define FMC_EXP_Q                                                                                      \
  {5.00000000000000111e-01, 1.66666666666666685e-01, 4.16666666666241289e-02, 8.33333333333006153e-03, \
   1.38888889172137167e-03, 1.98412698630536177e-04, 2.48015212959543761e-05, 2.75572684599970641e-06, \
   2.76200884454097462e-07, 2.51003854955103203e-08}
#define FMC_EXP_R {1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10}  // log2(e), -ln2_hi, -ln2_lo
#if defined(__CUDACC__)
static __constant__ double kExpQ[10] = FMC_EXP_Q;  // device: DFMA operands straight from the constant bank
static __constant__ double kExpR[3] = FMC_EXP_R;
#endif
namespace host_tables {
static constexpr double kExpQ[10] = FMC_EXP_Q;
static constexpr double kExpR[3] = FMC_EXP_R;
}  // namespace host_tables

namespace bits {  // the two 32-bit words of a double, on the device and (for tests/host_emul) on the host
FMC_HD FMC_INLINE int hi(double v) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(v);
#else
  long long b;
  memcpy(&b, &v, 8);
  return (int)(b >> 32);
#endif
}
FMC_HD FMC_INLINE int lo(double v) {
#if defined(__CUDA_ARCH__)
  return __double2loint(v);
#else
  long long b;
  memcpy(&b, &v, 8);
  return (int)(b & 0xffffffffLL);
#endif
}
FMC_HD FMC_INLINE double make(int h, int l) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(h, l);
#else
  const long long b = (long long)(((unsigned long long)(unsigned)h << 32) | (unsigned long long)(unsigned)l);
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}
}  // namespace bits

// The device formulation, callable on the host too so that tests/host_emul can check it against
// 80-digit arithmetic (fmc::exp below is what the model plug-ins call).
FMC_HD FMC_INLINE double exp_fast(double x) {
#if !defined(__CUDA_ARCH__)
  using host_tables::kExpQ;
  using host_tables::kExpR;
#endif
  // Saturate |x| >= 1024 (and +-inf) to +-1024 BEFORE the integer part is extracted: round(x*log2e)
  // is read from the low 32-bit word of t below, which wraps once |x*log2e| >= 2^31 -- without
  // this, exp(-2.08e9) came out as ~1e307 instead of ~0.  Integer instructions only (the FP64
  // pipe is the kernel's bound); NaN passes through.
  {
    const int hx = bits::hi(x);  // (the low word is kept: any value in [1024, 1024 + 2^-20) will do)
    const bool big = (unsigned)((hx & 0x7fffffff) - 0x40900000) <= (unsigned)(0x7ff00000 - 0x40900000);
    x = bits::make(big ? (int)(((unsigned)hx & 0x80000000u) | 0x40900000u) : hx, bits::lo(x));
  }
  const double t = fma(x, kExpR[0], 6755399441055744.0);  // round(x*log2e) in the low word
  int kk = bits::lo(t);
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
  kk = kk < -1021 ? -1021 : (kk > 1022 ? 1022 : kk);
  return bits::make(bits::hi(p) + (kk << 20), bits::lo(p));
}

FMC_HD FMC_INLINE double exp(double x) {
#if defined(__CUDA_ARCH__) || defined(FMC_HOST_EXP_FAST)  // (the latter: host images that mimic the device)
  return exp_fast(x);
#else
  return ::exp(x);
#endif
}

}  // namespace fmc
