// Experimental shared-first-trial kernels (trial1_kernel.cuh) of the compiled-in models, in
// their own translation unit so that the default kernels are compiled exactly as before.
#include "models/model_user.cuh"
#include "models/models_bench.cuh"
#include "trial1_kernel.cuh"

namespace fmc {

Trial1Entry fmc_trial1_entry(int model) {
  switch (model) {
    case 0: return Trial1Binding<ModelUser>::entry();
    case 1: return Trial1Binding<ModelTest4>::entry();
    case 2: return Trial1Binding<ModelPeaks10>::entry();
    case 3: return Trial1Binding<ModelStiff4>::entry();
  }
  return Trial1Entry{nullptr, nullptr, nullptr};
}

}  // namespace fmc
