// TEST INFRASTRUCTURE ONLY -- C-ABI hooks that expose single routines of the
// NL2SOL restatement so tests can check them against numpy/scipy in isolation.
#include <cstring>
#include <vector>

#include "nl2sol_port.hpp"

extern "C" {

double fmc_oracle_v2norm(int p, const double* x) { return nl2::v2norm(p, x); }
double fmc_oracle_dotprd(int p, const double* x, const double* y) { return nl2::dotprd(p, x, y); }

// qrfact + qapply on a column-major n x p matrix (toms573.f90:4360, 4281).
// On return a holds the Householder vectors / strict upper R, alpha the diagonal of R,
// ipivot the 1-based permutation, r = Q**T r.
int fmc_oracle_qrfact(int n, int p, double* a, double* alpha, int* ipivot, double* r) {
  std::vector<double> sum(p);
  int ierr = 0;
  nl2::qrfact(n, p, a, n, alpha, ipivot, &ierr, 0, sum.data());
  if (r) nl2::qapply(n, p, a, n, r, ierr);
  return ierr;
}

int fmc_oracle_lsqrt(int p, double* l, const double* a) {
  int irc = 0;
  nl2::lsqrt(1, p, l, a, &irc);
  return irc;
}

void fmc_oracle_livmul(int p, double* x, const double* l, const double* y) { nl2::livmul(p, x, l, y); }
void fmc_oracle_litvmu(int p, double* x, const double* l, const double* y) { nl2::litvmu(p, x, l, y); }

// One trust-region step from scratch for a given Jacobian/residual/scale/radius.
// model 1: qrfact + qapply + lmstep (toms573.f90:1066-1097); model 2: gqtstp on
// H = D^-1 (J^T J + S) D^-1 (toms573.f90:1101-1172).  vout = v(1:9) after the call.
int fmc_oracle_tr_step(int model, int n, int p, const double* jac, const double* res,
                       const double* d, const double* s_packed, double radius, double stppar0,
                       double rad0, double* step, double* vout, int* ka_out) {
  std::vector<int> iv(nl2::iv_len(p));
  std::vector<double> v(100, 0.0);
  nl2::dfault(iv.data(), v.data());
  std::vector<double> j(jac, jac + (size_t)n * p), qtr(res, res + n), g(p), rd(p), sum(p);
  std::vector<int> ipiv(p);
  const int pp = p * (p + 1) / 2;
  for (int k = 0; k < p; ++k) g[k] = nl2::dotprd(n, res, &j[(size_t)k * n]);
  v[8 - 1] = radius;
  v[9 - 1] = rad0;
  v[5 - 1] = stppar0;
  int ka = -1;
  if (model == 1) {
    int ierr = 0;
    nl2::qrfact(n, p, j.data(), n, rd.data(), ipiv.data(), &ierr, 0, sum.data());
    nl2::qapply(n, p, j.data(), n, qtr.data(), ierr);
    std::vector<double> rmat(pp), w(p * (p + 5) / 2 + 4 + 8, 0.0);
    int k = 0;
    rmat[0] = rd[0];
    for (int i = 2; i <= p; ++i) {
      std::memcpy(&rmat[k + 1], &j[(size_t)(i - 1) * n], sizeof(double) * (i - 1));
      k += i;
      rmat[k] = rd[i - 1];
    }
    nl2::lmstep(d, g.data(), &ierr, ipiv.data(), &ka, p, qtr.data(), rmat.data(), step, v.data(), w.data());
  } else {
    std::vector<double> h(pp), l(pp), dig(p), w(4 * p + 7 + 8, 0.0);
    int k = 0;
    for (int i = 0; i < p; ++i)
      for (int m = 0; m <= i; ++m, ++k)
        h[k] = (1.0 / d[i]) * (nl2::dotprd(n, &j[(size_t)i * n], &j[(size_t)m * n]) + (s_packed ? s_packed[k] : 0.0)) / d[m];
    for (int i = 0; i < p; ++i) dig[i] = g[i] / d[i];
    int ix = 2;
    nl2::gqtstp(d, dig.data(), h.data(), &ka, l.data(), p, step, v.data(), w.data(), &ix);
  }
  std::memcpy(vout, v.data(), sizeof(double) * 9);
  *ka_out = ka;
  return 0;
}


// nl2sno on an arbitrary residual function supplied by the test (ctypes callback), called the
// way fit_data calls it (bootstrap.f90:219-225: dfault, silent, `ncalls` calls in a row, each
// restarting from the previous answer).  Lets the tests run the classical least-squares test
// problems with known minima (More, Garbow & Hillstrom 1981 -- the set the NL2SOL paper uses)
// through the restatement.  counters[call][0..3] = iv(1), nfcall, ngcall, niter; fval = v(f)
// = half the sum of squares at the final point.
typedef void (*fmc_resid_cb)(int n, int p, const double* x, double* r);

namespace {
struct CbCtx {
  fmc_resid_cb cb;
};
void cb_calcr(int n, int p, const double* x, int* nf, double* r, void* vctx) {
  (void)nf;
  ((CbCtx*)vctx)->cb(n, p, x, r);
}
}  // namespace

int fmc_oracle_nl2sno_generic(int n, int p, double* x, fmc_resid_cb cb, int ncalls, int* counters, double* fval) {
  std::vector<int> iv(nl2::iv_len(p));
  std::vector<double> v(nl2::v_len(n, p));
  CbCtx ctx{cb};
  nl2::Options opt;
  opt.copy_jac = false;
  int ix = 2;
  opt.lsvmin_ix = &ix;
  for (int call = 0; call < ncalls; ++call) {
    nl2::dfault(iv.data(), v.data());
    for (int k = 19; k <= 24; ++k) iv[k - 1] = 0;
    iv[14 - 1] = 0;  // covprt, covreq: no covariance (dead work on this path)
    iv[15 - 1] = 0;
    nl2::nl2sno(n, p, x, cb_calcr, iv.data(), v.data(), &ctx, opt);
    if (counters) {
      counters[4 * call + 0] = iv[0];
      counters[4 * call + 1] = iv[6 - 1];
      counters[4 * call + 2] = iv[30 - 1];
      counters[4 * call + 3] = iv[31 - 1];
    }
  }
  if (fval) *fval = v[10 - 1];  // v(f)
  return 0;
}
}  // extern "C"
