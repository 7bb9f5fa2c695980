// The stress model of BASELINE config C5: its fits diverge (5 ... 500 passes each), the warps of an SM are in the solver
// a third of the time, and what the refill kernel is then short of is instruction cache for the pass loop (ncu:
// profiles/r2_refill_kernel_regions_c5.txt).  So this model is built with the COMPACT solver -- every loop rolled, helpers
// out of line, the flavour of the round schedule's solve kernel -- instead of the rare-paths-rolled one of the other models:
// C5 +18 %, at -3 % for the well-behaved fits of the same model (profiles/r2_refill_solver_flavours.json).
#define FMC_SOLVER_ROLLED 1
#include "model_registry.cuh"
#include "models/models_bench.cuh"
namespace fmc { ModelEntry fmc_entry_model3() { return ModelBinding<ModelStiff4>::entry(); } }
