// Model 0 = the user's plug-in (models/model_user.cuh; as shipped, the reference's model).
#define FMC_SOLVER_RARE_ROLLED 1  // rare solver paths rolled and out of line (nl2_solver.cuh)
#include "model_registry.cuh"
#include "models/model_user.cuh"
namespace fmc { ModelEntry fmc_entry_model0() { return ModelBinding<ModelUser>::entry(); } }
