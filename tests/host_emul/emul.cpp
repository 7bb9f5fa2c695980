// Host build of the PRODUCT's solver logic (fitter_mc_b200/csrc/nl2_solver.cuh,
// pass_eval.cuh, rng.cuh, models) for unit tests in a container without a GPU.
//
// This is a test harness: it drives exactly the same per-fit state machine and per-row
// pass code the CUDA kernel uses, one fit after another on the CPU, so that its logic can
// be compared with the oracle here before a GPU run.  It is compiled only by the tests; the
// shipped library has no CPU path.
#include <cstring>
#include <vector>

#include "../../fitter_mc_b200/csrc/models/model_user.cuh"
#include "../../fitter_mc_b200/csrc/models/models_bench.cuh"
#include "../../fitter_mc_b200/csrc/nl2_solver.cuh"
#include "../../fitter_mc_b200/csrc/pass_eval.cuh"
#include "../../fitter_mc_b200/csrc/rng.cuh"
#include "mgh_models.hpp"

namespace fmc {
// Exercises every operation pert.cuh / dual.cuh define (the bundled models use only + - * / exp pow).
struct ModelAllOps {
  static constexpr int NPARAM = 3;
  static constexpr int NPROP = 1;
  template <class T>
  static T model_y(const T* p, double x) {
    const T a = sqrt(p[0] * x + 1.0) * sin(p[1] * x) - cos(p[2]) / (1.0 + p[0] * p[0]);
    const T b = log(p[1] + x) * pow(p[2], p[0]) + pow(x + 1.0, p[1]) - pow(p[0] + 2.0, 1.5);
    const T c = (-p[2] + 3.0 * x) / (x + p[1]) + 2.0 / p[0] - (+p[1]) + (0.5 - p[2]) * (x - p[0]) + exp(-p[0] * x);
    return a + b * 0.25 + c;
  }
  static void derived_properties(const double* p, double* q) { q[0] = p[0]; }
};

}  // namespace fmc

using namespace fmc;

namespace {

int g_spec = 1;  // Nl2::spec of the fits that follow (emul_set_spec): 1 = Jacobian with every trial point, 0 = residual sum first

template <class M, int JAC = JAC_FD>
void run_pass(Nl2<M::NPARAM>& s, int n, const double* x, const double* yfit, const double* rerr) {
  constexpr int P = M::NPARAM;
  PassPoint<P> a, b;
  a.set(s.x, s.hA);
  b.set(s.x0, s.hJ);
  PassAcc<P> acc;
  acc.zero();
  for (int i = 0; i < n; ++i) {
    if (s.need_b)
      row_accumulate<M, true, JAC>(a, b, x[i], yfit[i], rerr[i], acc);
    else
      row_accumulate<M, false, JAC>(a, b, x[i], yfit[i], rerr[i], acc);
  }
  if constexpr (JAC == JAC_ANALYTIC)
    pass_finish_analytic<P>(acc, s);
  else
    pass_finish<P>(acc, s);
}

// counters[pass][0..3] = rc, nfcall, ngcall, niter ; counters[pass][4] = passes over the data
template <class M, int JAC = JAC_FD>
int fit_one(int n, const double* x, const double* yfit, const double* err, double* params,
            int* counters, int* anomaly_out = nullptr) {
  constexpr int P = M::NPARAM;
  std::vector<double> rerr(n);
  for (int i = 0; i < n; ++i) rerr[i] = 1.0 / err[i];
  Nl2<P> s;
  std::memset(&s, 0, sizeof(s));
  s.reset_fit();
  s.spec = g_spec;  // (the host pass below always evaluates everything; spec = 0 only changes what the solver asks for and reads)
  for (int pass = 0; pass < 2; ++pass) {
    s.begin(params);
    int npass = 0;
    for (;;) {
      run_pass<M, JAC>(s, n, x, yfit, rerr.data());
      ++npass;
      if (!s.resume()) break;
      if (npass > 2 * (k::mxfcal + k::mxiter)) {  // the kernel's cap on passes per call
        s.anomaly |= 8;
        s.rc = 9;
        break;
      }
    }
    if (anomaly_out) *anomaly_out |= s.anomaly << (8 * pass);
    for (int i = 0; i < P; ++i) params[i] = s.x[i];
    if (counters) {
      int* c = counters + 5 * pass;
      c[0] = s.rc;
      c[1] = s.nfcall;
      c[2] = s.ngcall;
      c[3] = s.niter;
      c[4] = npass;
    }
  }
  return 0;
}

#ifdef FMC_TRACE
// fit_one with the solver's decision trace (tools/parity_attribution.py): trace[pass][cap][10]
template <class M, int JAC = JAC_FD>
int fit_one_trace(int n, const double* x, const double* yfit, const double* err, double* params, int* counters,
                  double* trace, int cap, int* len) {
  constexpr int P = M::NPARAM;
  std::vector<double> rerr(n);
  for (int i = 0; i < n; ++i) rerr[i] = 1.0 / err[i];
  Nl2<P> s;
  std::memset(&s, 0, sizeof(s));
  s.reset_fit();
  for (int pass = 0; pass < 2; ++pass) {
    s.begin(params);
    s.trace = trace + (size_t)pass * cap * 10;
    s.trace_cap = cap;
    s.trace_len = 0;
    int npass = 0;
    for (;;) {
      run_pass<M, JAC>(s, n, x, yfit, rerr.data());
      ++npass;
      if (!s.resume()) break;
      if (npass > 2 * (k::mxfcal + k::mxiter)) break;
    }
    len[pass] = s.trace_len;
    for (int i = 0; i < P; ++i) params[i] = s.x[i];
    int* c = counters + 5 * pass;
    c[0] = s.rc;
    c[1] = s.nfcall;
    c[2] = s.ngcall;
    c[3] = s.niter;
    c[4] = npass;
  }
  return 0;
}
#endif

template <class M, int JAC = JAC_FD>
int bootstrap(int n, const double* x, const double* y, const double* err, const double* start,
              long long nres, const double* z, unsigned long long seed, long long first,
              double* params_out, double* props_out, int* counters_out, int* anomalies_out = nullptr) {
  constexpr int P = M::NPARAM, Q = M::NPROP;
  std::vector<double> yfit(n), p(P);
  for (long long r = 0; r < nres; ++r) {
    if (z) {
      for (int i = 0; i < n; ++i) yfit[i] = fma(z[(size_t)r * n + i], err[i], y[i]);
    } else {
      for (int i = 0; i < n; i += 4) {
        float zz[4];
        normals4(seed, (uint64_t)(first + r), (uint32_t)(i / 4), zz);
        for (int m = 0; m < 4 && i + m < n; ++m) yfit[i + m] = fma((double)zz[m], err[i + m], y[i + m]);
      }
    }
    for (int i = 0; i < P; ++i) p[i] = start[i];
    int anom = 0;
    fit_one<M, JAC>(n, x, yfit.data(), err, p.data(), counters_out ? counters_out + 10 * r : nullptr, &anom);
    if (anomalies_out) anomalies_out[r] = anom;
    std::memcpy(params_out + (size_t)r * P, p.data(), sizeof(double) * P);
    if (props_out) M::derived_properties(p.data(), props_out + (size_t)r * Q);
  }
  return 0;
}


// ---- host image of the warp-per-fit mapping (csrc/warp_kernel.cuh): the rows of a pass are
// dealt to 32 "lanes" in groups of four (lane l: groups l, l+32, ...), each lane accumulates
// its own partial sums, and a butterfly (the order of __shfl_xor with offsets 16, 8, 4, 2, 1)
// combines them.  Returns the totals as lane 0 sees them and checks that all lanes agree.
template <int P>
bool butterfly(PassAcc<P>* lane_acc, bool with_b) {
  constexpr int PP = P * (P + 1) / 2;
  constexpr int NV = 1 + P + PP + P;
  double v[32][NV];
  for (int l = 0; l < 32; ++l) {
    int m = 0;
    v[l][m++] = lane_acc[l].f;
    for (int i = 0; i < P; ++i) v[l][m++] = lane_acc[l].g[i];
    for (int i = 0; i < PP; ++i) v[l][m++] = lane_acc[l].G[i];
    for (int i = 0; i < P; ++i) v[l][m++] = with_b ? lane_acc[l].lky[i] : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    double nv[32][NV];
    for (int l = 0; l < 32; ++l)
      for (int m = 0; m < NV; ++m) nv[l][m] = v[l][m] + v[l ^ o][m];
    std::memcpy(v, nv, sizeof(v));
  }
  bool same = true;
  for (int l = 1; l < 32; ++l) same = same && std::memcmp(v[l], v[0], sizeof(v[0])) == 0;
  int m = 0;
  lane_acc[0].f = v[0][m++];
  for (int i = 0; i < P; ++i) lane_acc[0].g[i] = v[0][m++];
  for (int i = 0; i < PP; ++i) lane_acc[0].G[i] = v[0][m++];
  for (int i = 0; i < P; ++i) lane_acc[0].lky[i] = v[0][m++];
  return same;
}

template <class M, int JAC>
bool warp_pass(Nl2<M::NPARAM>& s, bool first, int n, const double* x, const double* yfit, const double* rerr) {
  constexpr int P = M::NPARAM;
  const int n_pad = (n + 3) / 4 * 4;
  PassPoint<P> a, b;
  double h1[P];
  for (int i = 0; i < P; ++i) h1[i] = (first && JAC == JAC_ANALYTIC) ? 1.0 : s.hA[i];
  a.set(s.x, first ? h1 : s.hA);
  b.set(s.x0, s.hJ);
  PassAcc<P> lane_acc[32];
  for (int l = 0; l < 32; ++l) {
    lane_acc[l].zero();
    for (int i = 4 * l; i < n_pad; i += 4 * 32)
      for (int m = 0; m < 4; ++m) {
        const int r = i + m;
        if (r >= n) continue;  // padding rows contribute exactly zero on the device (rerr = 0)
        if (!first)
          row_accumulate<M, true, JAC>(a, b, x[r], yfit[r], rerr[r], lane_acc[l]);
        else
          row_accumulate<M, false, JAC>(a, a, x[r], yfit[r], rerr[r], lane_acc[l]);
      }
  }
  const bool same = butterfly<P>(lane_acc, !first);
  if constexpr (JAC == JAC_ANALYTIC)
    pass_finish_analytic<P>(lane_acc[0], s);
  else
    pass_finish<P>(lane_acc[0], s);
  return same;
}

// one fit_data through the warp mapping; returns false if the lanes ever disagreed
template <class M, int JAC>
bool fit_one_warp(int n, const double* x, const double* yfit, const double* err, double* params, int* counters) {
  constexpr int P = M::NPARAM;
  std::vector<double> rerr(n);
  for (int i = 0; i < n; ++i) rerr[i] = 1.0 / err[i];
  Nl2<P> s;
  std::memset(&s, 0, sizeof(s));
  s.reset_fit();
  bool same = true;
  for (int call = 0; call < 2; ++call) {
    s.begin(params);
    int passes = 0;
    same = warp_pass<M, JAC>(s, true, n, x, yfit, rerr.data()) && same;
    ++passes;
    bool more = s.resume();
    while (more) {
      same = warp_pass<M, JAC>(s, false, n, x, yfit, rerr.data()) && same;
      ++passes;
      more = s.resume();
      if (more && passes > 2 * (k::mxfcal + k::mxiter)) more = false;
    }
    for (int i = 0; i < P; ++i) params[i] = s.x[i];
    if (counters) {
      int* c = counters + 5 * call;
      c[0] = s.rc;
      c[1] = s.nfcall;
      c[2] = s.ngcall;
      c[3] = s.niter;
      c[4] = passes;
    }
  }
  return same;
}

template <class M, int JAC>
int bootstrap_warp(int n, const double* x, const double* y, const double* err, const double* start, long long nres,
                   const double* z, double* params_out, int* counters_out) {
  constexpr int P = M::NPARAM;
  std::vector<double> yfit(n), p(P);
  int disagreements = 0;
  for (long long r = 0; r < nres; ++r) {
    for (int i = 0; i < n; ++i) yfit[i] = fma(z[(size_t)r * n + i], err[i], y[i]);
    for (int i = 0; i < P; ++i) p[i] = start[i];
    if (!fit_one_warp<M, JAC>(n, x, yfit.data(), err, p.data(), counters_out ? counters_out + 10 * r : nullptr))
      ++disagreements;
    std::memcpy(params_out + (size_t)r * P, p.data(), sizeof(double) * P);
  }
  return disagreements;
}

// model value and forward differences dm[k] = m(x) - m(x + h_k e_k) through (a) the product's
// row_model (perturbation arithmetic for generic plug-ins) and (b) p+1 plain evaluations
template <class M>
void row_model_both(const double* x, const double* h, double xi, double* m0, double* dm_engine, double* dm_plain) {
  constexpr int P = M::NPARAM;
  PassPoint<P> pt;
  pt.set(x, h);
  double dm[P];
  row_model<M>(pt, xi, *m0, dm);
  for (int kk = 0; kk < P; ++kk) {
    dm_engine[kk] = dm[kk];
    double xs[P];
    for (int m = 0; m < P; ++m) xs[m] = (m == kk) ? x[m] + h[m] : x[m];
    dm_plain[kk] = *m0 - M::model_y(xs, xi);
  }
}

// d model_y / d params_k as the analytic path forms it (dual numbers or the plug-in's model_grad)
template <class M>
void row_grad(const double* x, double xi, double* grad) {
  constexpr int P = M::NPARAM;
  double ones[P];
  for (int i = 0; i < P; ++i) ones[i] = 1.0;
  PassPoint<P> pt;
  pt.set(x, ones);
  double m0, dm[P];
  row_model<M, JAC_ANALYTIC>(pt, xi, m0, dm);
  for (int i = 0; i < P; ++i) grad[i] = -dm[i];
}

// A double-only plug-in WITH its own madj (model_grad), and one without any derivative.
struct ModelHandGrad {
  static constexpr int NPARAM = 3;
  static constexpr int NPROP = 1;
  static double model_y(const double* p, double x) { return p[0] + p[1] * std::exp(-p[2] * x); }
  static void model_grad(const double* p, double x, double* g) {
    const double e = std::exp(-p[2] * x);
    g[0] = 1.0;
    g[1] = e;
    g[2] = -p[1] * x * e;
  }
  static void derived_properties(const double* p, double* q) { q[0] = p[0] + p[1]; }
};
struct ModelPlainOnly {
  static constexpr int NPARAM = 3;
  static constexpr int NPROP = 1;
  static double model_y(const double* p, double x) { return p[0] + p[1] * std::exp(-p[2] * x); }
  static void derived_properties(const double* p, double* q) { q[0] = p[0] + p[1]; }
};
static_assert(can_analytic_v<ModelUser> && can_analytic_v<ModelTest4> && can_analytic_v<ModelPeaks10> &&
                  can_analytic_v<ModelStiff4>,
              "the compiled-in models are generic, so they have dual-number derivatives");
static_assert(has_model_grad<ModelHandGrad>::value && !has_dual_model<ModelHandGrad>::value &&
                  can_analytic_v<ModelHandGrad>,
              "a plug-in's own model_grad is detected");
static_assert(!can_analytic_v<ModelPlainOnly> && !use_pert_v<ModelPlainOnly>, "double-only model: no derivative");

}  // namespace

extern "C" {

void emul_set_spec(int spec) { g_spec = spec ? 1 : 0; }

// jac = 0: forward differences (nl2sno), 1: analytic (nl2sol + madj).  model 100 = the
// double-only test plug-in with a hand-written model_grad (same formula as model 0).
int emul_bootstrap_jac(int model, int jac, int n, const double* x, const double* y, const double* err,
                       const double* start, long long nres, const double* z, double* params_out, double* props_out,
                       int* counters_out, int* anomalies_out) {
#define EMUL_CASE(ID, M)                                                                                          \
  case ID:                                                                                                        \
    return jac ? bootstrap<M, JAC_ANALYTIC>(n, x, y, err, start, nres, z, 0, 0, params_out, props_out,            \
                                            counters_out, anomalies_out)                                          \
               : bootstrap<M, JAC_FD>(n, x, y, err, start, nres, z, 0, 0, params_out, props_out, counters_out,   \
                                      anomalies_out);
  switch (model) {
    EMUL_CASE(0, ModelUser)
    EMUL_CASE(1, ModelTest4)
    EMUL_CASE(2, ModelPeaks10)
    EMUL_CASE(3, ModelStiff4)
    EMUL_CASE(100, ModelHandGrad)
  }
#undef EMUL_CASE
  return -1;
}

// One fit_data (two solver calls) of the product's solver on a classical test problem
// (mgh_models.hpp), jac = 0 forward differences / 1 analytic.  counters: [2][5].
int emul_fit_classic(int problem, int jac, int n, const double* t, const double* y, double* params, int* counters,
                     int* anomaly) {
  std::vector<double> err(n, 1.0);
  *anomaly = 0;
#define MGH_CASE(ID, M)                                                                       \
  case ID:                                                                                    \
    return jac ? fit_one<M, JAC_ANALYTIC>(n, t, y, err.data(), params, counters, anomaly)     \
               : fit_one<M, JAC_FD>(n, t, y, err.data(), params, counters, anomaly);
  switch (problem) {
    MGH_CASE(0, MghRosenbrock)
    MGH_CASE(1, MghPowellSingular)
    MGH_CASE(2, MghFreudensteinRoth)
    MGH_CASE(3, MghBard)
    MGH_CASE(4, MghKowalikOsborne)
    MGH_CASE(5, MghMeyer)
    MGH_CASE(6, MghOsborne1)
    MGH_CASE(7, MghBox3D)
    MGH_CASE(8, MghBrownDennis)
    MGH_CASE(9, MghJennrichSampson)
    MGH_CASE(10, NistSaturation)
    MGH_CASE(11, NistDanWood)
    MGH_CASE(12, NistLanczos)
  }
#undef MGH_CASE
  return -1;
}

// The product's gqtstp on a given scaled Hessian H = D^-1 (J^T J + S) D^-1 (packed lower triangle
// by rows), scaled gradient dig and radius, from a fresh start (ka = -1), p = 2.
int emul_gqtstp2(const double* hmat, const double* dig, const double* d, double radius, double* step,
                 int* ka_out, int* anomaly_out) {
  Nl2<2> s;
  std::memset(&s, 0, sizeof(s));
  s.reset_fit();
  for (int i = 0; i < 3; ++i) s.hmat[i] = hmat[i];
  for (int i = 0; i < 2; ++i) {
    s.dig[i] = dig[i];
    s.d[i] = d[i];
  }
  s.v[Nl2<2>::RADIUS] = radius;
  s.kagqt = -1;
  s.gqtstp();
  for (int i = 0; i < 2; ++i) step[i] = s.step[i];
  *ka_out = s.kagqt;
  *anomaly_out = s.anomaly;
  return 0;
}

// The warp-per-fit mapping (host image).  Returns the number of fits in which the 32 lanes did
// not end a butterfly with bitwise identical totals (must be 0), or -1 for a bad model id.
int emul_bootstrap_warp(int model, int jac, int n, const double* x, const double* y, const double* err,
                        const double* start, long long nres, const double* z, double* params_out,
                        int* counters_out) {
#define WARP_CASE(ID, M)                                                                                  \
  case ID:                                                                                                \
    return jac ? bootstrap_warp<M, JAC_ANALYTIC>(n, x, y, err, start, nres, z, params_out, counters_out)  \
               : bootstrap_warp<M, JAC_FD>(n, x, y, err, start, nres, z, params_out, counters_out);
  switch (model) {
    WARP_CASE(0, ModelUser)
    WARP_CASE(1, ModelTest4)
    WARP_CASE(2, ModelPeaks10)
    WARP_CASE(3, ModelStiff4)
  }
#undef WARP_CASE
  return -1;
}

#ifdef FMC_TRACE
int emul_fit_trace(int model, int jac, int n, const double* x, const double* yfit, const double* err, double* params,
                   int* counters, double* trace, int cap, int* len) {
#define TRACE_CASE(ID, M)                                                                               \
  case ID:                                                                                              \
    return jac ? fit_one_trace<M, JAC_ANALYTIC>(n, x, yfit, err, params, counters, trace, cap, len)     \
               : fit_one_trace<M, JAC_FD>(n, x, yfit, err, params, counters, trace, cap, len);
  switch (model) {
    TRACE_CASE(0, ModelUser)
    TRACE_CASE(1, ModelTest4)
    TRACE_CASE(2, ModelPeaks10)
    TRACE_CASE(3, ModelStiff4)
  }
#undef TRACE_CASE
  return -1;
}
#endif

int emul_row_grad(int model, const double* x, double xi, double* grad) {
  switch (model) {
    case 0: row_grad<ModelUser>(x, xi, grad); return 0;
    case 1: row_grad<ModelTest4>(x, xi, grad); return 0;
    case 2: row_grad<ModelPeaks10>(x, xi, grad); return 0;
    case 3: row_grad<ModelStiff4>(x, xi, grad); return 0;
    case 100: row_grad<ModelHandGrad>(x, xi, grad); return 0;
    case 101: row_grad<ModelAllOps>(x, xi, grad); return 0;
  }
  return -1;
}

int emul_fit(int model, int n, const double* x, const double* yfit, const double* err, double* params,
             int* counters) {
  switch (model) {
    case 0: return fit_one<ModelUser>(n, x, yfit, err, params, counters);
    case 1: return fit_one<ModelTest4>(n, x, yfit, err, params, counters);
    case 2: return fit_one<ModelPeaks10>(n, x, yfit, err, params, counters);
    case 3: return fit_one<ModelStiff4>(n, x, yfit, err, params, counters);
  }
  return -1;
}

int emul_bootstrap2(int model, int n, const double* x, const double* y, const double* err,
                    const double* start, long long nres, const double* z, unsigned long long seed,
                    long long first, double* params_out, double* props_out, int* counters_out,
                    int* anomalies_out) {
  switch (model) {
    case 0: return bootstrap<ModelUser>(n, x, y, err, start, nres, z, seed, first, params_out, props_out, counters_out, anomalies_out);
    case 1: return bootstrap<ModelTest4>(n, x, y, err, start, nres, z, seed, first, params_out, props_out, counters_out, anomalies_out);
    case 2: return bootstrap<ModelPeaks10>(n, x, y, err, start, nres, z, seed, first, params_out, props_out, counters_out, anomalies_out);
    case 3: return bootstrap<ModelStiff4>(n, x, y, err, start, nres, z, seed, first, params_out, props_out, counters_out, anomalies_out);
  }
  return -1;
}

int emul_bootstrap(int model, int n, const double* x, const double* y, const double* err,
                   const double* start, long long nres, const double* z, unsigned long long seed,
                   long long first, double* params_out, double* props_out, int* counters_out) {
  return emul_bootstrap2(model, n, x, y, err, start, nres, z, seed, first, params_out, props_out, counters_out, nullptr);
}

int emul_row_model(int model, const double* x, const double* h, double xi, double* m0, double* dm_engine,
                   double* dm_plain) {
  switch (model) {
    case 0: row_model_both<ModelUser>(x, h, xi, m0, dm_engine, dm_plain); return 0;
    case 1: row_model_both<ModelTest4>(x, h, xi, m0, dm_engine, dm_plain); return 0;
    case 2: row_model_both<ModelPeaks10>(x, h, xi, m0, dm_engine, dm_plain); return 0;
    case 3: row_model_both<ModelStiff4>(x, h, xi, m0, dm_engine, dm_plain); return 0;
    case 101: row_model_both<ModelAllOps>(x, h, xi, m0, dm_engine, dm_plain); return 0;
  }
  return -1;
}

int emul_uses_pert(int model) {
  switch (model) {
    case 0: return use_pert_v<ModelUser>;
    case 1: return use_pert_v<ModelTest4>;
    case 2: return use_pert_v<ModelPeaks10>;
    case 3: return use_pert_v<ModelStiff4>;
  }
  return -1;
}

void emul_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
  U4 c{c0, c1, c2, c3};
  U4 o = philox4x32_10(c, k0, k1);
  out[0] = o.x;
  out[1] = o.y;
  out[2] = o.z;
  out[3] = o.w;
}

void emul_normals(unsigned long long seed, long long id, int n, double* z) {
  for (int i = 0; i < n; i += 4) {
    float zz[4];
    normals4(seed, (uint64_t)id, (uint32_t)(i / 4), zz);
    for (int m = 0; m < 4 && i + m < n; ++m) z[i + m] = (double)zz[m];
  }
}

}  // extern "C"
