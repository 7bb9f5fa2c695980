// Benchmark models declared in BASELINE.md (configs C2/C3, C4, C5), written through the
// same three plug-in routines as models/model_user.cuh.  They are selected by model id
// through the C ABI; id 0 is always the user's model.
#pragma once
#include "../fmc_common.cuh"
#include "../fast_math.cuh"
#include "../pert.cuh"

namespace fmc {

// C2/C3: the declared 4-parameter extension of the reference model, y = a + b*exp(-al*x) + c*x.
struct ModelTest4 {
  static constexpr int NPARAM = 4;
  static constexpr int NPROP = 1;
  static void initialise_model(const char** model_name, const char* (*param_name)[16],
                               const char* (*prop_name)[16], double* params_init) {
    *model_name = "test4";
    (*param_name)[0] = "a ";
    (*param_name)[1] = "b ";
    (*param_name)[2] = "al";
    (*param_name)[3] = "c ";
    (*prop_name)[0] = "ab";
    params_init[0] = 0.0;
    params_init[1] = 2.0;
    params_init[2] = 1.0;
    params_init[3] = 0.0;
  }
  template <class T>
  FMC_HD static FMC_INLINE T model_y(const T* params, double x) {
    return params[0] + params[1] * exp(-params[2] * x) + params[3] * x;
  }
  FMC_HD static FMC_INLINE void derived_properties(const double* params, double* props) {
    props[0] = params[0] + params[1];
  }
};

// C4: constant + three Gaussian peaks, 10 parameters; properties = the three peak areas.
struct ModelPeaks10 {
  static constexpr int NPARAM = 10;
  static constexpr int NPROP = 3;
  static void initialise_model(const char** model_name, const char* (*param_name)[16],
                               const char* (*prop_name)[16], double* params_init) {
    static const char* pn[10] = {"a0", "A1", "m1", "w1", "A2", "m2", "w2", "A3", "m3", "w3"};
    static const char* qn[3] = {"I1", "I2", "I3"};
    static const double truth[10] = {1.0, 5.0, 2.5, 0.4, 3.0, 5.0, 0.6, 4.0, 7.5, 0.5};
    *model_name = "peaks10";
    for (int i = 0; i < 10; ++i) {
      (*param_name)[i] = pn[i];
      params_init[i] = truth[i] * 1.05;
    }
    for (int i = 0; i < 3; ++i) (*prop_name)[i] = qn[i];
  }
  template <class T>
  FMC_HD static FMC_INLINE T model_y(const T* params, double x) {
    T y = params[0];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      // 1/w is a function of the parameters alone: the compiler lifts it (and its perturbation) out
      // of the loop over the data points, which then has no division left
      const T rw = 1.0 / params[3 + 3 * kk];
      const T u = (x - params[2 + 3 * kk]) * rw;
      y = y + params[1 + 3 * kk] * exp(-0.5 * (u * u));
    }
    return y;
  }
  FMC_HD static FMC_INLINE void derived_properties(const double* params, double* props) {
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) props[kk] = params[1 + 3 * kk] * params[3 + 3 * kk] * 2.5066282746310002;
  }
};

// C5: stiff exponential + power law, y = a*exp(-b*x) + c*x^(-d); properties ln2/b and a/c.
struct ModelStiff4 {
  static constexpr int NPARAM = 4;
  static constexpr int NPROP = 2;
  static void initialise_model(const char** model_name, const char* (*param_name)[16],
                               const char* (*prop_name)[16], double* params_init) {
    *model_name = "stiff4";
    (*param_name)[0] = "a ";
    (*param_name)[1] = "b ";
    (*param_name)[2] = "c ";
    (*param_name)[3] = "d ";
    (*prop_name)[0] = "th";
    (*prop_name)[1] = "ac";
    for (int i = 0; i < 4; ++i) params_init[i] = 1.0;
  }
  // x^(-d) = exp(-d log x): log x_i belongs to the data, not to the fit, so it is declared as a
  // per-row auxiliary value (pass_eval.cuh) -- evaluated once per data point when the data are
  // set instead of once per data point, pass and fit.
  static constexpr int NAUX = 1;
  FMC_HD static FMC_INLINE void row_aux(double x, double* aux) { aux[0] = log(x); }
  template <class T>
  FMC_HD static FMC_INLINE T model_y(const T* params, double x, const double* aux) {
    return params[0] * exp(-params[1] * x) + params[2] * exp(-params[3] * aux[0]);
  }
  template <class T>
  FMC_HD static FMC_INLINE T model_y(const T* params, double x) {  // the reference's argument list
    double aux[NAUX];
    row_aux(x, aux);
    return model_y(params, x, aux);
  }
  FMC_HD static FMC_INLINE void derived_properties(const double* params, double* props) {
    props[0] = 0.6931471805599453 / params[1];
    props[1] = params[0] / params[2];
  }
};

}  // namespace fmc
