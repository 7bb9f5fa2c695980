#define FMC_SOLVER_RARE_ROLLED 1  // rare solver paths rolled and out of line (nl2_solver.cuh)
#include "model_registry.cuh"
#include "models/models_bench.cuh"
namespace fmc { ModelEntry fmc_entry_model3() { return ModelBinding<ModelStiff4>::entry(); } }
