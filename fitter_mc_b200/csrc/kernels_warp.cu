// Warp-per-fit kernels (warp_kernel.cuh) of the compiled-in models, in their own translation
// unit so that the thread-per-fit kernels are compiled exactly as before.
#include "models/model_user.cuh"
#include "models/models_bench.cuh"
#include "warp_kernel.cuh"

namespace fmc {

WarpEntry fmc_warp_entry(int model) {
  switch (model) {
    case 0: return WarpBinding<ModelUser>::entry();
    case 1: return WarpBinding<ModelTest4>::entry();
    case 2: return WarpBinding<ModelPeaks10>::entry();
    case 3: return WarpBinding<ModelStiff4>::entry();
  }
  return WarpEntry{nullptr, nullptr};
}

}  // namespace fmc
