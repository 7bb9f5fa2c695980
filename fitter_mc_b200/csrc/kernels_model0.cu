// Model 0 = the user's plug-in (models/model_user.cuh; as shipped, the reference's model).
#include "model_registry.cuh"
#include "models/model_user.cuh"
namespace fmc { ModelEntry fmc_entry_model0() { return ModelBinding<ModelUser>::entry(); } }
