// Classical nonlinear least-squares test problems (More, Garbow & Hillstrom, ACM TOMS 7 (1981)
// 17-41 -- the set the NL2SOL paper reports on), written as model plug-ins so that the
// PRODUCT's solver code (nl2_solver.cuh + pass_eval.cuh, host build) can be run on them by the
// CPU tests.  Each residual r_i(p) is expressed as  y_i - model_y(p, t_i)  with err_i = 1; the
// data (t_i, y_i) come from the test (tests/test_classic_problems.py), which also defines the
// same residuals independently in numpy for the oracle.  TEST CODE, not shipped.
#pragma once
#include "../../fitter_mc_b200/csrc/fast_math.cuh"
#include "../../fitter_mc_b200/csrc/fmc_common.cuh"
#include "../../fitter_mc_b200/csrc/dual.cuh"
#include "../../fitter_mc_b200/csrc/pert.cuh"

namespace fmc {

#define MGH_COMMON(P)                                                                \
  static constexpr int NPARAM = P;                                                   \
  static constexpr int NPROP = 1;                                                    \
  static void derived_properties(const double* p, double* q) { q[0] = p[0]; }

// Rosenbrock: r1 = 10 (x2 - x1^2), r2 = 1 - x1.  t = 0, 1; y = (0, 1).
struct MghRosenbrock {
  MGH_COMMON(2)
  template <class T>
  static T model_y(const T* p, double t) {
    if (t < 0.5) return -10.0 * (p[1] - p[0] * p[0]);
    return p[0];
  }
};

// Powell singular: r = (x1 + 10 x2, sqrt5 (x3 - x4), (x2 - 2 x3)^2, sqrt10 (x1 - x4)^2).  y = 0.
struct MghPowellSingular {
  MGH_COMMON(4)
  template <class T>
  static T model_y(const T* p, double t) {
    if (t < 0.5) return -(p[0] + 10.0 * p[1]);
    if (t < 1.5) return -2.23606797749979 * (p[2] - p[3]);
    if (t < 2.5) {
      const T u = p[1] - 2.0 * p[2];
      return -(u * u);
    }
    const T u = p[0] - p[3];
    return -3.1622776601683795 * (u * u);
  }
};

// Freudenstein & Roth: r1 = -13 + x1 + ((5 - x2) x2 - 2) x2,  r2 = -29 + x1 + ((x2 + 1) x2 - 14) x2.
struct MghFreudensteinRoth {
  MGH_COMMON(2)
  template <class T>
  static T model_y(const T* p, double t) {
    if (t < 0.5) return -(p[0] + ((5.0 - p[1]) * p[1] - 2.0) * p[1]);   // y = -13
    return -(p[0] + ((p[1] + 1.0) * p[1] - 14.0) * p[1]);                // y = -29
  }
};

// Bard: y_i - (x1 + u / (v x2 + w x3)),  u = i, v = 16 - i, w = min(u, v);  t = i.
struct MghBard {
  MGH_COMMON(3)
  template <class T>
  static T model_y(const T* p, double t) {
    const double u = t, v = 16.0 - t, w = u < v ? u : v;
    return p[0] + u / (v * p[1] + w * p[2]);
  }
};

// Kowalik & Osborne: y_i - x1 (u^2 + u x2) / (u^2 + u x3 + x4);  t = u_i.
struct MghKowalikOsborne {
  MGH_COMMON(4)
  template <class T>
  static T model_y(const T* p, double t) {
    return p[0] * (t * t + t * p[1]) / (t * t + t * p[2] + p[3]);
  }
};

// Meyer: x1 exp(x2 / (t + x3)) - y_i,  t = 45 + 5 i.
struct MghMeyer {
  MGH_COMMON(3)
  template <class T>
  static T model_y(const T* p, double t) {
    return p[0] * exp(p[1] / (t + p[2]));
  }
};

// Osborne 1: y_i - (x1 + x2 exp(-t x4) + x3 exp(-t x5)),  t = 10 (i - 1).
struct MghOsborne1 {
  MGH_COMMON(5)
  template <class T>
  static T model_y(const T* p, double t) {
    return p[0] + p[1] * exp(-t * p[3]) + p[2] * exp(-t * p[4]);
  }
};

// Box three-dimensional: exp(-t x1) - exp(-t x2) - x3 (exp(-t) - exp(-10 t)),  t = 0.1 i;  y = 0.
struct MghBox3D {
  MGH_COMMON(3)
  template <class T>
  static T model_y(const T* p, double t) {
    const double c = ::exp(-t) - ::exp(-10.0 * t);
    return -(exp(-t * p[0]) - exp(-t * p[1]) - p[2] * c);
  }
};

// Brown & Dennis: (x1 + t x2 - e^t)^2 + (x3 + x4 sin t - cos t)^2,  t = i / 5;  y = 0.
// A large-residual problem: the secant (S) part of NL2SOL's model matters here.
struct MghBrownDennis {
  MGH_COMMON(4)
  template <class T>
  static T model_y(const T* p, double t) {
    const T a = p[0] + t * p[1] - ::exp(t);
    const T b = p[2] + p[3] * ::sin(t) - ::cos(t);
    return -(a * a + b * b);
  }
};

// Jennrich & Sampson: 2 + 2 i - (exp(i x1) + exp(i x2));  t = i;  y_i = 2 + 2 i.  Large residual.
struct MghJennrichSampson {
  MGH_COMMON(2)
  template <class T>
  static T model_y(const T* p, double t) {
    return exp(t * p[0]) + exp(t * p[1]);
  }
};

// NIST StRD Misra1a and BoxBOD: y_i - b1 (1 - exp(-b2 x_i)).
struct NistSaturation {
  MGH_COMMON(2)
  template <class T>
  static T model_y(const T* p, double t) {
    return p[0] * (1.0 - exp(-p[1] * t));
  }
};

// NIST StRD DanWood: y_i - b1 x_i^b2.
struct NistDanWood {
  MGH_COMMON(2)
  template <class T>
  static T model_y(const T* p, double t) {
    return p[0] * pow(t, p[1]);
  }
};

// NIST StRD Lanczos3: y_i - (b1 exp(-b2 x) + b3 exp(-b4 x) + b5 exp(-b6 x)), six parameters.
struct NistLanczos {
  MGH_COMMON(6)
  template <class T>
  static T model_y(const T* p, double t) {
    return p[0] * exp(-p[1] * t) + p[2] * exp(-p[3] * t) + p[4] * exp(-p[5] * t);
  }
};

#undef MGH_COMMON

}  // namespace fmc
