#include "model_registry.cuh"
#include "models/models_bench.cuh"
namespace fmc { ModelEntry fmc_entry_model1() { return ModelBinding<ModelTest4>::entry(); } }
