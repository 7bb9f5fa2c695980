// Warp-per-fit kernels (warp_kernel.cuh) of the compiled-in models.  Compiled once per model
// (-DFMC_WARP_MODEL=k, see the Makefile) so that the instantiations build in parallel and the
// thread-per-fit kernels are compiled exactly as before.
#include "models/model_user.cuh"
#include "models/models_bench.cuh"
#include "warp_kernel.cuh"

namespace fmc {

#if FMC_WARP_MODEL == 0
WarpEntry fmc_warp_entry_model0() { return WarpBinding<ModelUser>::entry(); }
#elif FMC_WARP_MODEL == 1
WarpEntry fmc_warp_entry_model1() { return WarpBinding<ModelTest4>::entry(); }
#elif FMC_WARP_MODEL == 2
WarpEntry fmc_warp_entry_model2() { return WarpBinding<ModelPeaks10>::entry(); }
#elif FMC_WARP_MODEL == 3
WarpEntry fmc_warp_entry_model3() { return WarpBinding<ModelStiff4>::entry(); }
#else
#error "FMC_WARP_MODEL must be 0..3"
#endif

}  // namespace fmc
