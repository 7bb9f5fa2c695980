// TEST INFRASTRUCTURE ONLY -- parity oracle, not product code.
//
// CPU restatement of the NL2SOL v2.2 (TOMS 573) finite-difference driver
// `nl2sno` and everything it reaches, as vendored by the reference in
// toms573.f90.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may link or call this.
//
// PARITY PIN STATUS: the reference ships no golden vectors for the solver
//  and no Fortran compiler exists in this image, so the
// solver part of the oracle is "parity unpinned" by the reference itself.  It
// is pinned instead by (i) routine-by-routine review against toms573.f90
// (file:line cited at every routine), (ii) agreement with scipy/MINPACK on
// identical data (tests/test_oracle_solver.py), (iii) the mcfit.py probe
// numbers in BASELINE.md and golden vectors generated from mcfit.py
// (tests/golden), (iv) the published minima of ten classical test problems
// and the NIST StRD certified values of Misra1a, DanWood and BoxBOD
// (tests/test_classic_problems.py).  The RANLUX part IS pinned by ranlux.f:399-473.
//
// Layout convention: the reference's 1-based IV()/V() maps (SURVEY.md
// Appendix B) are kept so every index constant can be checked by eye.
#pragma once
#include <cstdint>

namespace nl2 {

// calcr callback, toms573.f90:605-614.  nf<=0 on return means "x rejected".
typedef void (*calcr_fn)(int n, int p, const double* x, int* nf, double* r, void* ctx);

struct Options {
  bool covariance = true;   // covclc pass at convergence (dead work in the reference, toms573.f90:1400-1418)
  bool copy_jac = true;     // allocate/copy j(nn,p) on each nl2itr call (toms573.f90:876-884, 1434-1442)
  int gqt_retry_cap = 0;    // > 0: bound gqtstp's retry loop (unbounded in the reference: a latent non-termination)
  int* lsvmin_ix = nullptr; // LCG state of lsvmin (SAVEd `ix`, toms573.f90:3974); nullptr -> process-global
  // Decision trace (tools/parity_attribution.py): one record of TRACE_W doubles after every call of
  // assess (toms573.f90:1191): iv(irc), model used for the step, nfcall, niter, f at the trial point,
  // reldx, radius, pivot order of the current QR factorisation (digits base p), iv(restor), nreduc.
  double* trace = nullptr;
  int trace_cap = 0;        // records
  int* trace_len = nullptr;
};
enum { TRACE_W = 10 };

// IV / V lengths as allocated at bootstrap.f90:273.
inline int iv_len(int p) { return p + 60; }
inline int v_len(int n, int p) { return 93 + n * (p + 3) + p * (3 * p + 33) / 2; }

void dfault(int* iv, double* v);                                   // toms573.f90:2337-2400
void nl2sno(int n, int p, double* x, calcr_fn calcr, int* iv, double* v, void* ctx,
            const Options& opt);                                   // toms573.f90:590-769
// calcj callback, toms573.f90:57-66: j(i,k) = d r_i / d x_k, column-major n x p.  nf = 0 on
// return means "Jacobian cannot be computed at x".
typedef void (*calcj_fn)(int n, int p, const double* x, int* nf, double* j, void* ctx);
// The analytic-Jacobian driver the reference's README (35-42) points users to ("introduce a
// subroutine madj ... and replace the call to nl2sno with a call to nl2sol").
void nl2sol(int n, int p, double* x, calcr_fn calcr, calcj_fn calcj, int* iv, double* v, void* ctx,
            const Options& opt);                                   // toms573.f90:35-587

// exposed for unit tests
double dotprd(int p, const double* x, const double* y);            // :2403
double v2norm(int p, const double* x);                             // :4789
void qrfact(int m, int n, double* qr, int ldq, double* alpha, int* ipivot, int* ierr,
            int nopivk, double* sum);                              // :4360
void qapply(int n, int p, const double* j, int ldj, double* r, int ierr);  // :4281
void lsqrt(int n1, int n, double* l, const double* a, int* irc);   // :3842
void livmul(int n, double* x, const double* l, const double* y);   // :3367
void litvmu(int n, double* x, const double* l, const double* y);   // :3334
void lmstep(const double* d, const double* g, int* ierr, int* ipivot, int* ka, int p,
            const double* qtr, const double* r, double* step, double* v, double* w);  // :3397
void gqtstp(const double* d, const double* dig, double* dihdi, int* ka, double* l, int p,
            double* step, double* v, double* w, int* lsvmin_ix);   // :2471

}  // namespace nl2
