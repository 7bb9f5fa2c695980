/* fmc_b200 -- C ABI of the B200-native bootstrap-fitting engine (libfmc_b200.so).
 *
 * This is the drop-in boundary for the hot path of ndd21/Fitter_MC: the Monte-Carlo
 * (parametric bootstrap) loop of `monte_carlo_fit_data`, bootstrap.f90:297-330, i.e. per
 * resample  offset_data(.TRUE.) -> fit_data (two nl2sno calls) -> derived_properties ->
 * accumulation of sum p, sum p^2, sum prop, sum prop^2.  The reference has no FFI for this
 * path (its plug-in surface is source-level, bootstrap.f90:365-407); these entry points are
 * what a Fortran `INTERFACE ... BIND(C)` block, a C++ driver or ctypes binds instead.  See
 * INTEGRATION.md for the Fortran binding and the five-line change to bootstrap.f90.
 *
 * Conventions
 *  - plain pointers and sizes; every array is caller-owned, contiguous, host memory unless
 *    the name says `_dev`; the library keeps nothing after a call returns except inside an
 *    fmc_context the caller created.
 *  - return value 0 = success, negative = error (fmc_strerror).  The Fortran shim turns a
 *    non-zero code into errstop('MONTE_CARLO_FIT_DATA', fmc_strerror(rc)), the reference's
 *    error convention (utils.f90:26-37).  Solver non-convergence is NOT an error: the
 *    reference never inspects nl2sno's return code (bootstrap.f90:215-226) and averages
 *    every fit; so does this library (the codes are returned as a histogram instead).
 *  - there is no CPU fallback: every compute entry point fails with FMC_ERR_NO_DEVICE
 *    when no CUDA device is usable.
 *  - not re-entrant per context; use one context per host thread / per GPU.  The one-call
 *    entry points (fmc_bootstrap_run, fmc_bootstrap_run_binned, fmc_fit_once,
 *    fmc_generate_offsets) share one cached context per GPU and are to be called from one
 *    host thread at a time -- like the routine they replace, which the reference calls once,
 *    outside any parallel region (fitter_mc.f90:36).
 *  - results are a deterministic function of (data, params_start, seed, global resample id):
 *    every resample draws its Gaussian offsets from a counter-based generator keyed by its
 *    global id, so any split of [first_counter, first_counter+nresample) over launches,
 *    GPUs or ranks gives the same fits (sums agree up to summation order).
 */
#ifndef FMC_B200_H
#define FMC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FMC_OK 0
#define FMC_ERR_NO_DEVICE (-1)   /* no usable CUDA device / driver */
#define FMC_ERR_CUDA (-2)        /* a CUDA call failed; fmc_last_cuda_error() has the text */
#define FMC_ERR_BAD_MODEL (-3)   /* model id out of range */
#define FMC_ERR_BAD_ARG (-4)     /* null pointer, ndata < nparam, nresample < 0, err_y <= 0 ... */
#define FMC_ERR_NO_DATA (-5)     /* fmc_launch before fmc_set_data */
#define FMC_ERR_STATE (-6)       /* fmc_finish without a launch in flight */
#define FMC_ERR_ALLOC (-7)       /* host or device allocation failed */
#define FMC_ERR_NCCL (-8)        /* NCCL could not be loaded or an NCCL call failed */
#define FMC_ERR_UNSUPPORTED (-9) /* an optional feature that this build of the library does not contain */

#define FMC_MAX_PARAM 16
#define FMC_DEFAULT_SEED 31415ULL /* the reference's seed, utils.f90:7 */

typedef struct fmc_context fmc_context; /* opaque: device buffers + stream of one GPU */

/* ---------------------------------------------------------------- model plug-in trio
 * Replaces SUBROUTINE initialise_model / FUNCTION model_y / SUBROUTINE derived_properties
 * (bootstrap.f90:365-387, 390-396, 399-407).  The routines themselves live in
 * fitter_mc_b200/csrc/models/model_user.cuh (model id 0; edit and rebuild, as with the
 * reference) and are compiled into the kernels; these are their host-callable images.
 * param_name / prop_name are the reference's CHARACTER(2) names, blank padded, NUL added. */
int fmc_model_count(void);
int fmc_initialise_model(int model, int* nparam, int* nprop, double* params_init,
                         char (*param_name)[3], char (*prop_name)[3], char* model_name32);
double fmc_model_y(int model, const double* params, double x);
int fmc_derived_properties(int model, const double* params, double* props);

/* ---------------------------------------------------------------- one-shot entry points */

/* The Monte-Carlo loop, bootstrap.f90:297-330 (zero sums -> loop -> the caller applies
 * 328-341).  Resamples [first_counter, first_counter + nresample) are split contiguously
 * over `ngpus` devices (0 = all visible), one host thread per GPU; the per-GPU accumulator
 * blocks are combined by ONE ncclAllReduce over NVLink (fmc_last_combine_method() == FMC_COMBINE_NCCL;
 * replaces the OpenMP REDUCTION(+) of bootstrap.f90:310-311 across GPUs).  Inputs: x ascending as after sort_data (bootstrap.f90:155-194),
 * err_y > 0 (bootstrap.f90:134-138).
 *   sum_p[nparam], sum_p2[nparam], sum_prop[nprop], sum_prop2[nprop]  raw sums, so that the
 *       caller's lines 328-341 apply unchanged (computed from compensated shifted sums).
 *   z_inject     NULL, or [nresample][ndata] offsets in units of err_y used INSTEAD of the
 *                generator (parity tests feed the same resamples to the oracle).
 *   params_out   NULL or [nresample][nparam]; props_out NULL or [nresample][nprop] (what the
 *                reference writes to histogram_<prop>.dat, bootstrap.f90:319-325).
 *   rc_count     NULL or [16]: histogram of nl2sno's return code iv(1) of the second call
 *                (3 x-conv, 4 rel-f, 5 both, 6 abs-f, 7 singular, 8 false, 9/10 limits).
 *   totals       NULL or [4]: sum over fits of nfcall, ngcall, niter, passes over the data. */
int fmc_bootstrap_run(int model, int ndata, const double* x, const double* y, const double* err_y,
                      const double* params_start, long long nresample, unsigned long long seed,
                      long long first_counter, int ngpus, const double* z_inject,
                      double* sum_p, double* sum_p2, double* sum_prop, double* sum_prop2,
                      double* params_out, double* props_out, long long* rc_count, long long* totals);

/* The un-resampled fit of bootstrap.f90:288-291: offset_data(.FALSE.); fit_data(params).
 * params is in/out; rc2[2] (may be NULL) receives nl2sno's return code of both calls. */
int fmc_fit_once(int model, int ndata, const double* x, const double* y, const double* err_y,
                 double* params, int* rc2);

/* Test hook: the raw output words of the device's counter-based generator (Philox4x32-10, rng.cuh).
 * counters_keys[n][6] = counter words c0..c3 then key words k0, k1; out[n][4].  Lets the published
 * known-answer vectors be checked on the GPU itself. */
int fmc_debug_philox4x32(const unsigned int* counters_keys, int n, unsigned int* out);

/* The Gaussian offsets (in units of err_y) the engine uses for resamples
 * [first_counter, first_counter+nresample): z_out[nresample][ndata].  Replaces
 * `rang(ndata, offset)` (utils.f90:345-364) as a reproducible per-sample generator. */
int fmc_generate_offsets(unsigned long long seed, long long first_counter, long long nresample,
                         int ndata, double* z_out);

/* How the last fmc_bootstrap_run / fmc_bootstrap_run_binned call combined its GPUs' accumulators. */
#define FMC_COMBINE_NONE 0 /* one GPU: nothing to combine */
#define FMC_COMBINE_NCCL 1 /* one ncclAllReduce(sum, double) over the GPUs' streams */
#define FMC_COMBINE_HOST 2 /* NCCL not loadable on this host: blocks added on the host */
int fmc_last_combine_method(void);

/* ---------------------------------------------------------------- NCCL combine, one rank per GPU
 * For callers that run one process (or thread) per GPU -- bench.py under torchrun, an MPI build of the
 * Fortran driver: rank 0 obtains an id, the caller broadcasts its 128 bytes by its own means
 * (MPI_Bcast, torch.distributed), every rank binds its context, and after each fmc_launch one
 * fmc_allreduce enqueues the single ncclAllReduce of the accumulator block (and of the binned counts,
 * when enabled) on the context's stream.  NCCL (libnccl.so.2) is loaded at run time.
 *   fmc_nccl_available   0 = not loadable, else NCCL's version code (or 1).
 *   fmc_allreduce        no-op for a context that was never bound or is the only rank. */
int fmc_nccl_available(void);
int fmc_nccl_unique_id(char* id128);
int fmc_nccl_init_rank(fmc_context* ctx, const char* id128, int nranks, int rank);
int fmc_allreduce(fmc_context* ctx);

/* ---------------------------------------------------------------- context API (one GPU)
 * For callers that keep data resident, overlap transfers, shard over ranks or combine
 * accumulators with their own collective (bench.py, the multi-GPU path). */
int fmc_create(fmc_context** ctx, int device);
int fmc_destroy(fmc_context* ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t) instead of the context's own. */
int fmc_set_stream(fmc_context* ctx, void* cuda_stream);
/* Copy the data to the GPU (also precomputes 1/err_y, bootstrap.f90:140). */
int fmc_set_data(fmc_context* ctx, int model, int ndata, const double* x, const double* y,
                 const double* err_y);
/* Accumulate into a caller-owned device buffer of fmc_accumulator_len() doubles (e.g. a
 * torch tensor that is then all-reduced with NCCL).  NULL = the context's own buffer. */
int fmc_set_accumulator_dev(fmc_context* ctx, double* acc_dev);
int fmc_accumulator_len(const fmc_context* ctx);
/* Copy the accumulator block (after the work queued on the context's stream, e.g. fmc_launch and
 * fmc_allreduce) to the host; decode it with fmc_decode_accumulator. */
int fmc_get_accumulator(fmc_context* ctx, double* acc_host);
/* Enqueue: zero the accumulator, run the fit kernel for `nresample` resamples.  Asynchronous.
 * z_inject_dev: NULL or DEVICE pointer [nresample][ndata].  keep_per_resample != 0 keeps
 * per-resample params/props/counters on the device for fmc_finish. */
int fmc_launch(fmc_context* ctx, const double* params_start, long long nresample,
               unsigned long long seed, long long first_counter, const double* z_inject_dev,
               int keep_per_resample);
/* Wait for the launch and read back.  Any output pointer may be NULL.
 * counters_out [nresample][10]: per nl2sno call (rc, nfcall, ngcall, niter, passes). */
int fmc_finish(fmc_context* ctx, double* sum_p, double* sum_p2, double* sum_prop, double* sum_prop2,
               double* params_out, double* props_out, int* counters_out, long long* rc_count,
               long long* totals);
/* Wait for the launch without reading anything back (the caller owns the accumulator, e.g.
 * after all-reducing it across ranks). */
int fmc_wait(fmc_context* ctx);
/* Turn an accumulator block (host copy, possibly summed over GPUs) into the raw sums. */
int fmc_decode_accumulator(int model, const double* params_start, const double* acc_host,
                           double* sum_p, double* sum_p2, double* sum_prop, double* sum_prop2,
                           long long* nfit, long long* rc_count, long long* totals);
/* Fits of the last launch whose solver hit one of its safety caps (a fit may never spin forever
 * on the GPU; such a fit ends like a false convergence).  Returns their number (0 in any normal
 * run) and fills up to `cap` global resample ids and flag words (bit0 lmstep, bit1 gqtstp, bit2
 * step recomputation, bit3 passes; bits 8+ = which nl2sno call).  Call after fmc_finish/fmc_wait. */
int fmc_last_anomalies(fmc_context* ctx, long long* ids, int* flags, int cap);

/* ---- analytic Jacobian ----------------------------------------------
 * The reference fits with nl2sno, which differences the residuals (toms573.f90:590-769), and
 * tells users who want analytic derivatives to write a `madj` and call nl2sol instead
 * (README:35-42; toms573.f90:35-587).  FMC_JACOBIAN_ANALYTIC is that switch: the same NL2SOL
 * iteration fed with exact Jacobian rows, taken from the plug-in's model_grad if it has one,
 * else from its generic model_y by dual numbers (csrc/dual.cuh).  Opt-in, because results
 * move at the 1e-8 level relative to the forward-difference path (the size of the forward-
 * difference truncation error).  ctx = NULL sets the mode of the one-call entry points
 * (fmc_bootstrap_run, fmc_bootstrap_run_binned, fmc_fit_once).  A launch in analytic mode on
 * a model without derivatives fails with FMC_ERR_BAD_ARG. */
#define FMC_JACOBIAN_FORWARD_DIFFERENCE 0
#define FMC_JACOBIAN_ANALYTIC 1
int fmc_set_jacobian_mode(fmc_context* ctx, int mode);
int fmc_model_has_analytic_jacobian(int model);

/* ---- how fits are mapped onto the GPU ---------------------------------------------------
 * FMC_FIT_PER_THREAD (default): one resample per thread, threads refill from a shared
 * counter; the measured choice for the small fits of this path.
 * FMC_FIT_PER_WARP: one resample per warp -- the data points of a pass are split over the 32
 * lanes and the p x p sums combined by warp shuffles; meant for large ndata x nparam.
 * Measured on B200 (round 2, profiles/r2_mapping_*.json): 0.47-0.70 of the thread-per-fit rate on
 * C4 (ndata 10^4, nparam 10) and 0.47 on C5, so it is NOT part of the default build
 * (`make WARP=1` compiles it in); without it the call returns FMC_ERR_UNSUPPORTED.
 * ctx = NULL sets the mapping of the one-call entry points. */
#define FMC_FIT_PER_THREAD 0
#define FMC_FIT_PER_WARP 1
int fmc_set_fit_mapping(fmc_context* ctx, int mapping);

/* ---- how the iterations of the fits are scheduled ----------------------------------------
 * FMC_SCHEDULE_REFILL: one persistent kernel per nl2sno call; a thread runs passes and solver
 * steps of its fit and takes the next fit from a counter when it ends.
 * FMC_SCHEDULE_ROUNDS: passes and solver steps in separate kernels that alternate inside a CUDA
 * graph WHILE node; fits that want another pass are compacted into lists between rounds
 * (DESIGN.md 2.8).  Same solver source and the same per-fit sequence of operations; the two compile it
 * differently, so a small share of the fits differs in the last bits (tests/test_round_schedule.py).
 * Under the round schedule with nparam > 6 a trial point gets its residual sum alone first and its Jacobian only
 * if the step is accepted and the call goes on (bit-identical fits, fewer Jacobians; FMC_ROUND_SPECULATE=0|1).
 * FMC_SCHEDULE_AUTO (default): by measurement -- rounds when the data rows do not fit in shared memory
 * or nparam > 6 (C4: +19 %), refill otherwise (C1, C2, and C5 -- strongly diverging fits -- once the model's
 * translation unit is built with the compact solver, csrc/kernels_model3.cu: 8.2e5 against 6.4e5 fits/s).
 * The environment variable
 * FMC_SCHEDULE=refill|rounds overrides the default.  ctx = NULL sets it for the one-call entry points.
 * fmc_get_schedule returns the setting, or, for AUTO, what the last launch resolved it to. */
#define FMC_SCHEDULE_REFILL 0
#define FMC_SCHEDULE_ROUNDS 1
#define FMC_SCHEDULE_AUTO 2
int fmc_set_schedule(fmc_context* ctx, int schedule);
int fmc_get_schedule(const fmc_context* ctx);

/* ---- binned distributions -------------------------------------------
 * Replaces the reference's histogram_<prop>.dat dumps -- one list-directed text line per
 * resample and property, written inside an OMP CRITICAL (bootstrap.f90:301-308, 319-325,
 * 348-355) -- which cannot scale to 1e8 resamples.  The fit kernel bins every fitted
 * parameter and derived property as the fit finishes; the counts are 64-bit integers, so the
 * result is exact and independent of the order in which fits finish or GPUs are combined.
 *
 * Quantity j (j < nparam: parameter j, else property j - nparam) has `nbins` equal-width bins
 * on [lo[j], hi[j]); counts are laid out [nparam + nprop][nbins + 2] with column 0 =
 * underflow (and NaN), 1..nbins the bins, nbins + 1 = overflow; bin(v) =
 * 1 + (int)((v - lo) * (nbins / (hi - lo))).  nbins = 0 switches binning off.  Needs
 * fmc_set_data first (the model fixes the number of quantities); the counts are zeroed at the
 * start of every fmc_launch, like the accumulator. */
int fmc_set_histogram(fmc_context* ctx, int nbins, const double* lo, const double* hi);
/* Number of 64-bit counts, (nparam + nprop) * (nbins + 2); 0 when binning is off. */
long long fmc_histogram_len(const fmc_context* ctx);
/* Let the counts live in caller-owned DEVICE memory (fmc_histogram_len() uint64), e.g. a
 * torch int64 tensor that is all-reduced across ranks afterwards.  NULL = internal buffer. */
int fmc_set_histogram_dev(fmc_context* ctx, unsigned long long* counts_dev);
/* Wait for the launch and copy the counts of this GPU to the host. */
int fmc_get_histogram(fmc_context* ctx, unsigned long long* counts_host);
/* Quantiles of ONE quantity from its nbins + 2 counts: the value below which a fraction
 * probs[i] of the binned resamples lies, linear within a bin.  A quantile that falls in the
 * underflow / overflow column is reported as lo / hi.  Returns the number of resamples that
 * fell outside [lo, hi) -- widen the range if that is not negligible -- or < 0 on bad input. */
long long fmc_histogram_quantiles(const unsigned long long* counts, int nbins, double lo, double hi,
                                  const double* probs, int nprobs, double* out);

/* One-call form: fmc_bootstrap_run without per-resample arrays but with the binned
 * distributions of all nparam + nprop quantities, summed over the GPUs used.  This is what a
 * make_histogram=.TRUE. run of the reference (bootstrap.f90:301-325) binds for large N. */
int fmc_bootstrap_run_binned(int model, int ndata, const double* x, const double* y,
                             const double* err_y, const double* params_start, long long nresample,
                             unsigned long long seed, long long first_counter, int ngpus,
                             double* sum_p, double* sum_p2, double* sum_prop, double* sum_prop2,
                             long long* rc_count, long long* totals, int nbins,
                             const double* bin_lo, const double* bin_hi, unsigned long long* counts);

/* Device time of the last fit kernel (CUDA events on the launching stream), milliseconds. */
float fmc_last_kernel_ms(fmc_context* ctx);
/* Number of this library's kernels the last fmc_launch enqueued (jac0, init_shared/init_pass, iterate,
 * warp_fit, hist -- per chunk of 2^21 resamples). */
long long fmc_last_launch_count(const fmc_context* ctx);
/* Grid geometry chosen for the last launch: blocks, threads per block, blocks per SM, SMs. */
int fmc_last_launch_geometry(const fmc_context* ctx, int* out4);

/* bootstrap.f90:328-341: mean and Bessel-corrected standard deviation from the raw sums. */
void fmc_finalise_moments(int k, long long nresample, const double* sum, const double* sum2,
                          double* mean, double* err);

/* FP64 roofline denominator: runs register-resident DFMA chains (iters x 128 FMAs per thread,
 * 8 independent chains, every SM full) and returns the achieved TFLOP/s (FMA = 2 flops). */
int fmc_measure_dfma_peak(int device, int iters, double* tflops, float* ms_out);

const char* fmc_strerror(int code);
const char* fmc_last_cuda_error(void);
const char* fmc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FMC_B200_H */
