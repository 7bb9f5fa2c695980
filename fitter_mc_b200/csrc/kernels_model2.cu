#include "model_registry.cuh"
#include "models/models_bench.cuh"
namespace fmc { ModelEntry fmc_entry_model2() { return ModelBinding<ModelPeaks10>::entry(); } }
