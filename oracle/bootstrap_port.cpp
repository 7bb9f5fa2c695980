// TEST INFRASTRUCTURE ONLY -- parity oracle, not product code.
//
// Restatement of the reference's bootstrap loop and model plug-in trio:
//   offset_data          bootstrap.f90:197-212
//   fit_data             bootstrap.f90:215-226   (dfault; iv(19:24)=0; nl2sno -- twice)
//   madr                 bootstrap.f90:229-250
//   monte_carlo_fit_data bootstrap.f90:297-341   (loop, OpenMP reduction, moments)
//   initialise_model / model_y / derived_properties   bootstrap.f90:365-407
#include "fmc_oracle.h"
#include "nl2sol_port.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- models
// Model 0 is the reference's (bootstrap.f90:365-407).  Models 1-3 are the
// benchmark extensions declared in BASELINE.md (C2/C3, C4, C5), written
// through the same three routines.
struct ModelDef {
  const char* name;
  int nparam, nprop;
  const char* pnames[FMC_ORACLE_MAX_P];
  const char* qnames[FMC_ORACLE_MAX_P];
  double init[FMC_ORACLE_MAX_P];
};

const ModelDef kModels[] = {
    {"test", 3, 1, {"a ", "b ", "al"}, {"ab"}, {0.0, 2.0, 1.0}},
    {"test4", 4, 1, {"a ", "b ", "al", "c "}, {"ab"}, {0.0, 2.0, 1.0, 0.0}},
    {"peaks10", 10, 3, {"a0", "A1", "m1", "w1", "A2", "m2", "w2", "A3", "m3", "w3"},
     {"I1", "I2", "I3"},
     {1.0 * 1.05, 5.0 * 1.05, 2.5 * 1.05, 0.4 * 1.05, 3.0 * 1.05, 5.0 * 1.05, 0.6 * 1.05,
      4.0 * 1.05, 7.5 * 1.05, 0.5 * 1.05}},
    {"stiff4", 4, 2, {"a ", "b ", "c ", "d "}, {"th", "ac"}, {1.0, 1.0, 1.0, 1.0}},
};
const int kNumModels = (int)(sizeof(kModels) / sizeof(kModels[0]));

inline double model_y(int model, const double* p, double x) {
  switch (model) {
    case 0:  // bootstrap.f90:395
      return p[0] + p[1] * std::exp(-p[2] * x);
    case 1:
      return p[0] + p[1] * std::exp(-p[2] * x) + p[3] * x;
    case 2: {
      double y = p[0];
      for (int k = 0; k < 3; ++k) {
        const double rw = 1.0 / p[3 + 3 * k];  // as written in csrc/models/models_bench.cuh
        double u = (x - p[2 + 3 * k]) * rw;
        y = y + p[1 + 3 * k] * std::exp(-0.5 * (u * u));
      }
      return y;
    }
    case 3:
      return p[0] * std::exp(-p[1] * x) + p[2] * std::exp(-p[3] * std::log(x));  // x^(-d), as in models_bench.cuh
    default:
      return 0.0;
  }
}

// Partial derivatives of model_y with respect to the parameters, derived by hand (the `madj`
// a user of nl2sol would write, README:35-42).
inline void model_grad(int model, const double* p, double x, double* g) {
  switch (model) {
    case 0: {
      const double e = std::exp(-p[2] * x);
      g[0] = 1.0;
      g[1] = e;
      g[2] = -p[1] * x * e;
      break;
    }
    case 1: {
      const double e = std::exp(-p[2] * x);
      g[0] = 1.0;
      g[1] = e;
      g[2] = -p[1] * x * e;
      g[3] = x;
      break;
    }
    case 2:
      g[0] = 1.0;
      for (int k = 0; k < 3; ++k) {
        const double A = p[1 + 3 * k], mu = p[2 + 3 * k], w = p[3 + 3 * k];
        const double rw = 1.0 / w, u = (x - mu) * rw, gs = std::exp(-0.5 * (u * u));
        g[1 + 3 * k] = gs;
        g[2 + 3 * k] = A * gs * u * rw;
        g[3 + 3 * k] = A * gs * u * u * rw;
      }
      break;
    case 3: {
      const double e = std::exp(-p[1] * x), pw = std::exp(-p[3] * std::log(x));
      g[0] = e;
      g[1] = -p[0] * x * e;
      g[2] = pw;
      g[3] = -p[2] * pw * std::log(x);
      break;
    }
    default:
      break;
  }
}

inline void derived_properties(int model, const double* p, double* props) {
  switch (model) {
    case 0:  // bootstrap.f90:406
    case 1:
      props[0] = p[0] + p[1];
      break;
    case 2:
      for (int k = 0; k < 3; ++k) props[k] = p[1 + 3 * k] * p[3 + 3 * k] * 2.5066282746310002;
      break;
    case 3:
      props[0] = 0.6931471805599453 / p[1];
      props[1] = p[0] / p[2];
      break;
    default:
      break;
  }
}

// ---------------------------------------------------------------- madr
struct FitCtx {
  int model;
  const double* x;
  const double* y_fit;  // data_y_offset
  const double* rerr;   // rec_errbar_y
  long long neval;
};

void madr(int n, int p, const double* params, int* nf, double* r, void* vctx) {
  (void)p;
  (void)nf;  // madr never sets nf = 0 (bootstrap.f90:229-250)
  FitCtx* c = (FitCtx*)vctx;
  for (int i = 0; i < n; ++i) r[i] = (c->y_fit[i] - model_y(c->model, params, c->x[i])) * c->rerr[i];
  c->neval++;
}

// madj: d r_i / d params_k = -(d model_y / d params_k) / err_i, column-major (toms573.f90:57-66)
void madj(int n, int p, const double* params, int* nf, double* j, void* vctx) {
  (void)nf;
  FitCtx* c = (FitCtx*)vctx;
  double g[FMC_ORACLE_MAX_P];
  for (int i = 0; i < n; ++i) {
    model_grad(c->model, params, c->x[i], g);
    for (int k = 0; k < p; ++k) j[(size_t)k * n + i] = -g[k] * c->rerr[i];
  }
}

struct Workspace {
  std::vector<int> iv;
  std::vector<double> v;
  Workspace(int n, int p) : iv(nl2::iv_len(p)), v(nl2::v_len(n, p)) {}
};

void fit_data(FitCtx& ctx, int n, int p, double* params, int flags, Workspace& ws, int* counters,
              double* trace = nullptr, int trace_cap = 0, int* trace_len = nullptr) {
  nl2::Options opt;
  opt.copy_jac = (flags & FMC_ORACLE_COPY_JAC) != 0;
  opt.gqt_retry_cap = (flags & FMC_ORACLE_SAFETY_CAP) ? 1000 : 0;
  int ix_local = 2;
  opt.lsvmin_ix = (flags & FMC_ORACLE_LSVMIN_PER_FIT) ? &ix_local : nullptr;
  int* iv = ws.iv.data();
  double* v = ws.v.data();
  for (int pass = 0; pass < 2; ++pass) {
    nl2::dfault(iv, v);
    for (int k = 19; k <= 24; ++k) iv[k - 1] = 0;  // "Make nl2sol silent", bootstrap.f90:221/224
    if (!(flags & FMC_ORACLE_COVARIANCE)) {
      // lean mode: covprt = covreq = 0 skips covclc (toms573.f90:1400); fitted
      // parameters are unchanged because covclc restores x and r exactly.
      iv[14 - 1] = 0;
      iv[15 - 1] = 0;
    }
    if (trace) {  // one trace per nl2sno call
      opt.trace = trace + (size_t)pass * trace_cap * nl2::TRACE_W;
      opt.trace_cap = trace_cap;
      opt.trace_len = trace_len + pass;
      trace_len[pass] = 0;
    }
    if (flags & FMC_ORACLE_ANALYTIC_JAC)
      nl2::nl2sol(n, p, params, madr, madj, iv, v, &ctx, opt);
    else
      nl2::nl2sno(n, p, params, madr, iv, v, &ctx, opt);
    if (counters) {
      int* c = counters + pass * FMC_ORACLE_NCOUNTERS;
      c[0] = iv[0];
      c[1] = iv[6 - 1];
      c[2] = iv[30 - 1];
      c[3] = iv[31 - 1];
      c[4] = iv[40 - 1];
      c[5] = iv[41 - 1];
    }
  }
}

}  // namespace

extern "C" {

int fmc_oracle_model_count(void) { return kNumModels; }

int fmc_oracle_initialise_model(int model, int* nparam, int* nprop, double* params_init,
                                char (*param_name)[3], char (*prop_name)[3], char* model_name32) {
  if (model < 0 || model >= kNumModels) return -1;
  const ModelDef& m = kModels[model];
  *nparam = m.nparam;
  *nprop = m.nprop;
  for (int i = 0; i < m.nparam; ++i) {
    if (params_init) params_init[i] = m.init[i];
    if (param_name) {
      std::memcpy(param_name[i], m.pnames[i], 2);
      param_name[i][2] = 0;
    }
  }
  for (int i = 0; i < m.nprop; ++i)
    if (prop_name) {
      std::memcpy(prop_name[i], m.qnames[i], 2);
      prop_name[i][2] = 0;
    }
  if (model_name32) std::snprintf(model_name32, 32, "%s", m.name);
  return 0;
}

double fmc_oracle_model_y(int model, const double* params, double x) { return model_y(model, params, x); }

void fmc_oracle_derived_properties(int model, const double* params, double* props) {
  derived_properties(model, params, props);
}

void fmc_oracle_madr(int model, int n, const double* x, const double* y_fit, const double* err,
                     const double* params, double* r) {
  for (int i = 0; i < n; ++i) r[i] = (y_fit[i] - model_y(model, params, x[i])) * (1.0 / err[i]);
}

void fmc_oracle_madj(int model, int n, const double* x, const double* err, const double* params, double* j) {
  const int p = kModels[model].nparam;
  double g[FMC_ORACLE_MAX_P];
  for (int i = 0; i < n; ++i) {
    model_grad(model, params, x[i], g);
    for (int k = 0; k < p; ++k) j[(size_t)k * n + i] = -g[k] * (1.0 / err[i]);
  }
}

int fmc_oracle_fit_data(int model, int n, const double* x, const double* y_fit, const double* err,
                        double* params, int flags, int* counters) {
  if (model < 0 || model >= kNumModels) return -1;
  const int p = kModels[model].nparam;
  if (n < p) return -2;
  std::vector<double> rerr(n);
  for (int i = 0; i < n; ++i) rerr[i] = 1.0 / err[i];  // bootstrap.f90:140
  FitCtx ctx{model, x, y_fit, rerr.data(), 0};
  Workspace ws(n, p);
  fit_data(ctx, n, p, params, flags, ws, counters);
  return 0;
}

int fmc_oracle_fit_trace(int model, int n, const double* x, const double* y_fit, const double* err,
                         double* params, int flags, int* counters, double* trace, int trace_cap, int* trace_len) {
  if (model < 0 || model >= kNumModels) return -1;
  const int p = kModels[model].nparam;
  if (n < p || !trace || !trace_len || trace_cap < 1) return -2;
  std::vector<double> rerr(n);
  for (int i = 0; i < n; ++i) rerr[i] = 1.0 / err[i];
  FitCtx ctx{model, x, y_fit, rerr.data(), 0};
  Workspace ws(n, p);
  fit_data(ctx, n, p, params, flags, ws, counters, trace, trace_cap, trace_len);
  return 0;
}

int fmc_oracle_bootstrap(int model, int n, const double* x, const double* y, const double* err,
                         const double* params_start, long long nresample, int rng_mode,
                         const double* z, int flags, int nthreads, double* sum_p, double* sum_p2,
                         double* sum_prop, double* sum_prop2, double* params_out,
                         double* props_out, int* counters_out, long long* rc_count,
                         long long* neval_total) {
  if (model < 0 || model >= kNumModels) return -1;
  const int p = kModels[model].nparam, q = kModels[model].nprop;
  if (n < p) return -2;
  if (rng_mode == FMC_ORACLE_RNG_INJECT && !z) return -3;
  std::vector<double> rerr(n);
  for (int i = 0; i < n; ++i) rerr[i] = 1.0 / err[i];
  for (int k = 0; k < p; ++k) sum_p[k] = sum_p2[k] = 0.0;
  for (int k = 0; k < q; ++k) sum_prop[k] = sum_prop2[k] = 0.0;
  long long rc_local[16] = {0};
  long long neval = 0;
  FILE* hist = nullptr;
  if (flags & FMC_ORACLE_HISTOGRAM) hist = std::fopen("/dev/null", "w");
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif

#pragma omp parallel num_threads(nthreads)
  {
    Workspace ws(n, p);
    std::vector<double> y_fit(n), offset(n), params(p), props(q > 0 ? q : 1);
    std::vector<double> sp(p, 0.0), sp2(p, 0.0), sq(q, 0.0), sq2(q, 0.0);
    long long rc_t[16] = {0};
    FitCtx ctx{model, x, y_fit.data(), rerr.data(), 0};
    int counters[2 * FMC_ORACLE_NCOUNTERS];
#pragma omp for schedule(static)
    for (long long i = 0; i < nresample; ++i) {
      for (int k = 0; k < p; ++k) params[k] = params_start[k];  // bootstrap.f90:313
      // offset_data(.TRUE.), bootstrap.f90:197-212
      if (rng_mode == FMC_ORACLE_RNG_RANLUX) {
#pragma omp critical(fmc_oracle_rng)
        fmc_oracle_rang(n, offset.data());
      } else {
        std::memcpy(offset.data(), z + (size_t)i * n, sizeof(double) * n);
      }
      for (int k = 0; k < n; ++k) y_fit[k] = y[k] + offset[k] * err[k];
      fit_data(ctx, n, p, params.data(), flags, ws, counters);
      derived_properties(model, params.data(), props.data());
      for (int k = 0; k < p; ++k) {
        sp[k] += params[k];
        sp2[k] += params[k] * params[k];
      }
      for (int k = 0; k < q; ++k) {
        sq[k] += props[k];
        sq2[k] += props[k] * props[k];
      }
      if (hist) {
#pragma omp critical(fmc_oracle_rng)  // the reference's two unnamed CRITICALs share one lock
        for (int k = 0; k < q; ++k) std::fprintf(hist, " %24.16E\n", props[k]);
      }
      int rc = counters[FMC_ORACLE_NCOUNTERS + 0];
      rc_t[(rc >= 0 && rc < 16) ? rc : 0]++;
      if (params_out) std::memcpy(params_out + (size_t)i * p, params.data(), sizeof(double) * p);
      if (props_out && q > 0) std::memcpy(props_out + (size_t)i * q, props.data(), sizeof(double) * q);
      if (counters_out)
        std::memcpy(counters_out + (size_t)i * 2 * FMC_ORACLE_NCOUNTERS, counters, sizeof(counters));
    }
#pragma omp critical(fmc_oracle_reduce)
    {
      for (int k = 0; k < p; ++k) {
        sum_p[k] += sp[k];
        sum_p2[k] += sp2[k];
      }
      for (int k = 0; k < q; ++k) {
        sum_prop[k] += sq[k];
        sum_prop2[k] += sq2[k];
      }
      for (int k = 0; k < 16; ++k) rc_local[k] += rc_t[k];
      neval += ctx.neval;
    }
  }
  if (hist) std::fclose(hist);
  if (rc_count) std::memcpy(rc_count, rc_local, sizeof(rc_local));
  if (neval_total) *neval_total = neval;
  return 0;
}

void fmc_oracle_finalise_moments(int k, long long nresample, const double* sum, const double* sum2,
                                 double* mean, double* err) {
  const double rtemp = 1.0 / (double)nresample;  // bootstrap.f90:328-330
  for (int i = 0; i < k; ++i) {
    double m = sum[i] * rtemp, m2 = sum2[i] * rtemp;
    mean[i] = m;
    if (nresample > 1) {
      double bessel = (double)nresample / (double)(nresample - 1);
      err[i] = std::sqrt(std::fmax((m2 - m * m) * bessel, 0.0));  // bootstrap.f90:334-337
    } else {
      err[i] = 0.0;
    }
  }
}

}  // extern "C"
