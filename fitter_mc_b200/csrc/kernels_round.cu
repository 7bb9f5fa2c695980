// Round-schedule kernels (round_kernels.cuh) of the compiled-in models.  Compiled once per model
// (-DFMC_ROUND_MODEL=k, see the Makefile) so that the instantiations build in parallel and the
// refill kernels are compiled exactly as before.
// The solver is compiled with rolled loops and out-of-line helpers here (nl2_solver.cuh, FMC_SOLVER_ROLLED):
// the solve kernel keeps its state in shared (or local) memory where dynamic indexing costs nothing, and
// what it is short of is instruction cache.  No other kernel of this translation unit contains solver code.
#if !defined(FMC_ROUND_SOLVER_UNROLLED)
#define FMC_SOLVER_ROLLED 1
#else
#define FMC_SOLVER_OUTLINE 1  // (experiment: unrolled loops, out-of-line helpers)
#endif
#include "models/model_user.cuh"
#include "models/models_bench.cuh"
#include "round_kernels.cuh"

namespace fmc {

#if FMC_ROUND_MODEL == 0
RoundEntry fmc_round_entry_model0() { return RoundBinding<ModelUser>::entry(); }
#elif FMC_ROUND_MODEL == 1
RoundEntry fmc_round_entry_model1() { return RoundBinding<ModelTest4>::entry(); }
#elif FMC_ROUND_MODEL == 2
RoundEntry fmc_round_entry_model2() { return RoundBinding<ModelPeaks10>::entry(); }
#elif FMC_ROUND_MODEL == 3
RoundEntry fmc_round_entry_model3() { return RoundBinding<ModelStiff4>::entry(); }
#else
#error "FMC_ROUND_MODEL must be 0..3"
#endif

}  // namespace fmc
