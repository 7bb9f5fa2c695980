// Common macros and NL2SOL default constants for the B200 bootstrap-fitting engine.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define FMC_HD __host__ __device__
#define FMC_INLINE __forceinline__
#else
// tests/host_emul compiles the solver logic with the host compiler so that it can be
// checked against the oracle in a container without a GPU.  The shipped library
// (libfmc_b200.so) is always built by nvcc and never takes this branch.
#define FMC_HD
#define FMC_INLINE inline
#endif

namespace fmc {

// Values of `dfault` (toms573.f90:2354-2398) for IEEE binary64, machep = 2^-52.
// bootstrap.f90:220-225 calls dfault before each nl2sno and changes nothing but the
// print flags, so these are compile-time constants here.
namespace k {
constexpr double machep = 2.220446049250313e-16;
constexpr double sqteps = 1.4901161193847656e-08;  // sqrt(machep) = 2^-26
constexpr double afctol = 1e-20;
constexpr double cosmin = 1e-6;   // max(1e-6, 1e2*machep)
constexpr double decfac = 0.5;
constexpr double dfac = 0.6;
constexpr double dltfdj = sqteps;
constexpr double d0init = 1.0;
constexpr double epslon = 0.1;
constexpr double fuzz = 1.5;
constexpr double incfac = 2.0;
constexpr double jtinit = 1e-6;
constexpr double lmax0 = 100.0;
constexpr double phmnfc = -0.1;
constexpr double phmxfc = 0.1;
constexpr double rdfcmn = 0.1;
constexpr double rdfcmx = 4.0;
constexpr double rfctol = 1e-10;  // max(1e-10, machep^(2/3) = 3.67e-11)
constexpr double rlimit = 1.3407807929942596e+154;  // sqrt(HUGE)
constexpr double tuner1 = 0.1;
constexpr double tuner2 = 1e-4;
constexpr double tuner3 = 0.75;
constexpr double tuner4 = 0.5;
constexpr double tuner5 = 0.75;
constexpr double xctol = sqteps;
constexpr double xftol = 1e2 * machep;
constexpr int mxfcal = 200;
constexpr int mxiter = 150;
constexpr int stglim = 2;
constexpr double sqteta = 1.4916681462400413e-154;  // sqrt(TINY), rmdcon(2)
}  // namespace k

}  // namespace fmc
