/* TEST INFRASTRUCTURE ONLY -- parity oracle, not product code.
 *
 * C ABI of the CPU restatement of the reference's bootstrap hot path
 * (bootstrap.f90:197-250, 288-341, 365-407; utils.f90:338-364; ranlux.f;
 * toms573.f90 nl2sno and callees).  Loaded with ctypes from tests/, from
 * __graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference
 * legs only.  Nothing under fitter_mc_b200/ may include or link this.
 *
 * Pin status: RANLUX pinned by ranlux.f:399-473; the solver is unpinned by the
 * reference (no golden vectors, no Fortran compiler here) -- see
 * nl2sol_port.hpp.
 */
#ifndef FMC_ORACLE_H
#define FMC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RANLUX + rang (ranlux.f, utils.f90:338-364) ---- */
void fmc_oracle_rluxgo(int lux, int seed, int k1, int k2);
void fmc_oracle_rlux_reset_default(void); /* back to NOTYET=.TRUE. (default seed 314159265, lux 3) */
void fmc_oracle_ranlux(double* rvec, int len);
long long fmc_oracle_rluxat_kount(void);
void fmc_oracle_initialise_rng(void);     /* rluxgo(3, 31415, 0, 0) */
void fmc_oracle_rang(int n, double* r);

/* ---- model plug-in trio (bootstrap.f90:365-407) ----
 * model 0 = the reference's default model; 1..3 are the declared benchmark
 * extensions of BASELINE.md (C2/C3 4-parameter, C4 10-parameter, C5 stiff). */
#define FMC_ORACLE_MAX_P 16
int fmc_oracle_model_count(void);
/* returns 0 on success; names are blank-padded CHARACTER(2) + NUL */
int fmc_oracle_initialise_model(int model, int* nparam, int* nprop, double* params_init,
                                char (*param_name)[3], char (*prop_name)[3], char* model_name32);
double fmc_oracle_model_y(int model, const double* params, double x);
void fmc_oracle_derived_properties(int model, const double* params, double* props);

/* ---- flags ---- */
#define FMC_ORACLE_COVARIANCE 1   /* keep the covclc dead work (reference-equivalent cost) */
#define FMC_ORACLE_COPY_JAC 2     /* keep the per-call allocate+copy of j(nn,p) */
#define FMC_ORACLE_LSVMIN_PER_FIT 4 /* reset lsvmin's LCG to 2 at the start of each fit (order-independent results) */
#define FMC_ORACLE_HISTOGRAM 8    /* write props to /dev/null-like text sink under a critical section (cost model of bootstrap.f90:319-325) */

#define FMC_ORACLE_SAFETY_CAP 32  /* bound gqtstp's retry loop at 1000 rounds (the reference's is unbounded and can spin
                                     forever on ill-conditioned fits -- tests/test_robustness.py); for checks on the C5
                                     stress data, where a hung checker would hang the test-suite */
#define FMC_ORACLE_ANALYTIC_JAC 16 /* fit with nl2sol + an analytic madj instead of nl2sno (README:35-42) */

/* counters[pass][k]: k = 0 iv(1) return code, 1 nfcall, 2 ngcall, 3 niter, 4 nfcov, 5 ngcov */
#define FMC_ORACLE_NCOUNTERS 6

/* fit_data (bootstrap.f90:215-226): two nl2sno passes on (x, y_fit, 1/err). */
int fmc_oracle_fit_data(int model, int n, const double* x, const double* y_fit, const double* err,
                        double* params, int flags, int* counters /* [2][6] or NULL */);

/* The same with a decision trace (tools/parity_attribution.py): after every call of assess
 * (toms573.f90:1191) one record of 10 doubles -- iv(irc), model used, nfcall, niter, f, reldx, radius,
 * pivot order, stppar, nreduc -- into trace[pass][trace_cap][10]; trace_len[2] = records per call. */
int fmc_oracle_fit_trace(int model, int n, const double* x, const double* y_fit, const double* err,
                         double* params, int flags, int* counters, double* trace, int trace_cap, int* trace_len);

/* residual vector madr (bootstrap.f90:229-250) and chi^2 (bootstrap.f90:452-458) */
void fmc_oracle_madr(int model, int n, const double* x, const double* y_fit, const double* err,
                     const double* params, double* r);

/* madj: the analytic Jacobian of madr, j[k*n + i] = d r_i / d params_k (what README:35-42
 * asks a user of nl2sol to write), from hand-derived formulas for the compiled-in models. */
void fmc_oracle_madj(int model, int n, const double* x, const double* err, const double* params,
                     double* j);

/* rng_mode for the bootstrap loop */
#define FMC_ORACLE_RNG_RANLUX 0   /* the reference's serial RANLUX+polar stream (state = the global generator) */
#define FMC_ORACLE_RNG_INJECT 1   /* z[nresample][n] supplied by the caller */

/* monte_carlo_fit_data loop body (bootstrap.f90:297-330) for nresample resamples.
 * Outputs are raw sums (the host applies bootstrap.f90:328-341).  params_out /
 * props_out / counters_out may be NULL.  counters_out is [nresample][2][6].
 * rc_count[16] histograms iv(1) of the second pass.  Returns 0 on success. */
int fmc_oracle_bootstrap(int model, int n, const double* x, const double* y, const double* err,
                         const double* params_start, long long nresample, int rng_mode,
                         const double* z, int flags, int nthreads,
                         double* sum_p, double* sum_p2, double* sum_prop, double* sum_prop2,
                         double* params_out, double* props_out, int* counters_out,
                         long long* rc_count, long long* neval_total);

/* bootstrap.f90:328-341: mean and Bessel-corrected standard deviation from raw sums. */
void fmc_oracle_finalise_moments(int k, long long nresample, const double* sum, const double* sum2,
                                 double* mean, double* err);

#ifdef __cplusplus
}
#endif
#endif
