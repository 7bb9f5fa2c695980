#include "model_registry.cuh"
#include "models/models_bench.cuh"
namespace fmc { ModelEntry fmc_entry_model3() { return ModelBinding<ModelStiff4>::entry(); } }
