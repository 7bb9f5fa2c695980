// THE MODEL PLUG-IN: edit this file to fit a different model, then rebuild
// (python -c "import __graft_entry__ as g; g.build()"  or  make -C fitter_mc_b200).
//
// It holds the same three routines the reference asks its users to edit
// ("YOU NEED TO CHANGE THIS SUBROUTINE IF YOU HAVE A NEW MODEL", bootstrap.f90:365-407),
// with the same argument lists, as host/device callbacks that are inlined into the CUDA
// fit kernel (edit-and-recompile, exactly like the reference; no function pointers):
//
//   initialise_model                    bootstrap.f90:365-387  -> sets nparam, nprop, names, params_init
//   model_y(params, x)                  bootstrap.f90:390-396  -> REAL(dp) FUNCTION model_y(params,x)
//   derived_properties(params, props)   bootstrap.f90:399-407  -> SUBROUTINE derived_properties(params,props)
//
// As shipped it is the reference's own model: y = a + b*exp(-al*x), property ab = a + b.
#pragma once
#include "../fmc_common.cuh"
#include "../fast_math.cuh"
#include "../pert.cuh"

namespace fmc {

struct ModelUser {
  // YOU NEED TO CHANGE THESE IF YOU HAVE A NEW MODEL.
  static constexpr int NPARAM = 3;  // Number of parameters in model.
  static constexpr int NPROP = 1;   // Number of parameter-dependent derived properties to report.

  // YOU NEED TO CHANGE THIS ROUTINE IF YOU HAVE A NEW MODEL.
  static void initialise_model(const char** model_name, const char* (*param_name)[16],
                               const char* (*prop_name)[16], double* params_init) {
    *model_name = "test";
    (*param_name)[0] = "a ";  // Parameter names (CHARACTER(2) in the reference).
    (*param_name)[1] = "b ";
    (*param_name)[2] = "al";
    (*prop_name)[0] = "ab";   // Derived-property names.
    params_init[0] = 0.0;     // Initial parameter values.
    params_init[1] = 2.0;
    params_init[2] = 1.0;
  }

  // The model function.
  // YOU NEED TO CHANGE THIS ROUTINE IF YOU HAVE A NEW MODEL.
  // Keep it generic in the scalar type T (write the formula once, as you would for double):
  // the engine then gets nl2sno's forward differences from one evaluation (pert.cuh).  A
  // plain `static double model_y(const double* params, double x)` also works, just slower.
  template <class T>
  FMC_HD static FMC_INLINE T model_y(const T* params, double x) {
    return params[0] + params[1] * exp(-params[2] * x);
  }

  // Useful or interesting functions of the parameter values, which will have their mean and
  // standard error evaluated.
  // YOU NEED TO CHANGE THIS ROUTINE IF YOU HAVE A NEW MODEL.
  FMC_HD static FMC_INLINE void derived_properties(const double* params, double* props) {
    props[0] = params[0] + params[1];
  }
};

}  // namespace fmc
