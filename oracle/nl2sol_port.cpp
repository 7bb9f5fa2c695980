// TEST INFRASTRUCTURE ONLY -- parity oracle, not product code.
// See nl2sol_port.hpp for scope, pin status and the indexing convention.
//
// Every routine below restates one routine of /root/reference/toms573.f90
// (NL2SOL v2.2 as vendored by the reference); the cited line ranges are the
// statement ranges it follows.  The control flow keeps the reference's
// statement labels (L10, L20, ...) so the two can be compared label by label.
// Compile with -ffp-contract=off: the restatement must not fuse a*b+c.
#include "nl2sol_port.hpp"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace nl2 {

namespace {

// machine_constants.f90:29-66 (rmdcon k = 1..6)
const double kTiny = DBL_MIN;          // TINY(1.d0)
const double kSqTiny = std::sqrt(DBL_MIN);
const double kMachep = DBL_EPSILON;    // EPSILON(1.d0)
const double kSqMachep = std::sqrt(DBL_EPSILON);
const double kSqHuge = std::sqrt(DBL_MAX);

// iv subscripts (toms573.f90:861-866, 1684-1685, 1993-1995)
enum {
  TOOBIG = 2, IRC = 3, MLSTGD = 4, MODEL = 5, NFCALL = 6, NFGCAL = 7, RADINC = 8, RESTOR = 9,
  STAGE = 10, STGLIM = 11, SWITCH = 12, XIRC = 13, COVPRT = 14, COVREQ = 15, DTYPE = 16,
  MXFCAL = 17, MXITER = 18, OUTLEV = 19, PARPRT = 20, PRUNIT = 21, SOLPRT = 22, STATPR = 23,
  X0PRT = 24, INITS = 25, COVMAT = 26, IV_D = 27, IV_G = 28, NGCALL = 30, NITER = 31, IERR = 32,
  IV_J = 33, CNVCOD = 34, KAGQT = 35, KALM = 36, LKY = 37, MODE = 38, NFCOV = 40, NGCOV = 41,
  DIG = 43, IV_H = 44, QTR = 49, IV_R = 50, RD = 51, RSAVE = 52, IV_S = 53, SAVEI = 54,
  STEP = 55, STLSTG = 56, SUSED = 57, LMAT = 58, IV_W = 59, X0 = 60, IPIVOT = 61, IPIV0 = 60
};
// v subscripts (toms573.f90:868-872, 1687-1690, 2351-2353)
enum {
  DGNORM = 1, DSTNRM = 2, DST0 = 3, GTSTEP = 4, STPPAR = 5, NREDUC = 6, PREDUC = 7, RADIUS = 8,
  RAD0 = 9, V_F = 10, FDIF = 11, FLSTGD = 12, F0 = 13, GTSLST = 14, PLSTGD = 15, RADFAC = 16,
  RELDX = 17, DSTSAV = 18, EPSLON = 19, PHMNFC = 20, PHMXFC = 21, DECFAC = 22, INCFAC = 23,
  RDFCMN = 24, RDFCMX = 25, TUNER1 = 26, TUNER2 = 27, TUNER3 = 28, TUNER4 = 29, TUNER5 = 30,
  AFCTOL = 31, RFCTOL = 32, XCTOL = 33, XFTOL = 34, LMAX0 = 35, DLTFDJ = 36, D0INIT = 37,
  DINIT = 38, JTINIT = 39, DLTFDC = 40, DFAC = 41, RLIMIT = 42, COSMIN = 43, DELTA0 = 44,
  FUZZ = 45, V_FX = 46, V_SIZE = 47, WSCALE = 48, XMSAVE = 49, V_DELTA = 50, VSAVE1 = 78,
  JTOL0 = 86, JTOL1 = 87
};
const int NVSAVE = 9;

#define IV(k) iv[(k)-1]
#define V(k) v[(k)-1]

inline void vcopy(int p, double* y, const double* x) { std::memmove(y, x, sizeof(double) * (size_t)(p > 0 ? p : 0)); }  // :4763
inline void vscopy(int p, double* y, double s) { for (int i = 0; i < p; ++i) y[i] = s; }  // :4776
inline void vaxpy(int p, double* w, double a, const double* x, const double* y) {        // :4750
  for (int i = 0; i < p; ++i) w[i] = a * x[i] + y[i];
}

int g_lsvmin_ix = 2;  // SAVEd LCG state of lsvmin, toms573.f90:3974

}  // namespace

// toms573.f90:2403-2431 -- inner product that skips products that would underflow.
double dotprd(int p, const double* x, const double* y) {
  double s = 0.0;
  if (p <= 0) return s;
  const double sqteta = kSqTiny;
  for (int i = 0; i < p; ++i) {
    double t = std::fmax(std::fabs(x[i]), std::fabs(y[i]));
    if (!(t > 1.0)) {
      if (t < sqteta) continue;
      t = (x[i] / sqteta) * y[i];
      if (std::fabs(t) < sqteta) continue;
    }
    s = s + x[i] * y[i];
  }
  return s;
}

// toms573.f90:4789-4838 -- scaled 2-norm, one divide per element.
double v2norm(int p, const double* x) {
  if (p <= 0) return 0.0;
  int i = 0;
  for (; i < p; ++i)
    if (x[i] != 0.0) break;
  if (i == p) return 0.0;
  double scale = std::fabs(x[i]);
  if (i == p - 1) return scale;
  double t = 1.0;
  const double sqteta = kSqTiny;
  for (int k = i + 1; k < p; ++k) {
    double xi = std::fabs(x[k]);
    if (!(xi > scale)) {
      double r = xi / scale;
      if (r > sqteta) t = t + r * r;
      continue;
    }
    double r = scale / xi;
    if (r <= sqteta) r = 0.0;
    t = 1.0 + t * r * r;
    scale = xi;
  }
  return scale * std::sqrt(t);
}

// toms573.f90:2337-2400 -- default iv / v values.
void dfault(int* iv, double* v) {
  IV(1) = 12;
  IV(COVPRT) = 1;
  IV(COVREQ) = 1;
  IV(DTYPE) = 1;
  IV(INITS) = 0;
  IV(MXFCAL) = 200;
  IV(MXITER) = 150;
  IV(OUTLEV) = 1;
  IV(PARPRT) = 1;
  IV(PRUNIT) = 6;  // imdcon(1), machine_constants.f90:20
  IV(SOLPRT) = 1;
  IV(STATPR) = 1;
  IV(X0PRT) = 1;
  const double machep = kMachep;
  V(AFCTOL) = 1e-20;
  if (machep > 1e-10) V(AFCTOL) = machep * machep;
  V(COSMIN) = std::fmax(1e-6, 1e2 * machep);
  V(DECFAC) = 0.5;
  const double sqteps = kSqMachep;
  V(DELTA0) = sqteps;
  V(DFAC) = 0.6;
  V(DINIT) = 0.0;
  const double mepcrt = std::pow(machep, 1.0 / 3.0);
  V(DLTFDC) = mepcrt;
  V(DLTFDJ) = sqteps;
  V(D0INIT) = 1.0;
  V(EPSLON) = 0.1;
  V(FUZZ) = 1.5;
  V(INCFAC) = 2.0;
  V(JTINIT) = 1e-6;
  V(LMAX0) = 100.0;
  V(PHMNFC) = -0.1;
  V(PHMXFC) = 0.1;
  V(RDFCMN) = 0.1;
  V(RDFCMX) = 4.0;
  V(RFCTOL) = std::fmax(1e-10, mepcrt * mepcrt);
  V(RLIMIT) = kSqHuge;
  V(TUNER1) = 0.1;
  V(TUNER2) = 1e-4;
  V(TUNER3) = 0.75;
  V(TUNER4) = 0.5;
  V(TUNER5) = 0.75;
  V(XCTOL) = sqteps;
  V(XFTOL) = 1e2 * machep;
}

// toms573.f90:3334-3364 -- solve (L**T) x = y, L packed lower-by-rows.
void litvmu(int n, double* x, const double* l, const double* y) {
  for (int i = 0; i < n; ++i) x[i] = y[i];
  int i0 = n * (n + 1) / 2;
  for (int ii = 1; ii <= n; ++ii) {
    int i = n + 1 - ii;
    double xi = x[i - 1] / l[i0 - 1];
    x[i - 1] = xi;
    if (i <= 1) return;
    i0 = i0 - i;
    if (xi == 0.0) continue;
    for (int j = 1; j <= i - 1; ++j) x[j - 1] = x[j - 1] - xi * l[i0 + j - 1];
  }
}

// toms573.f90:3367-3394 -- solve L x = y.
void livmul(int n, double* x, const double* l, const double* y) {
  int k;
  for (k = 1; k <= n; ++k) {
    if (y[k - 1] != 0.0) break;
    x[k - 1] = 0.0;
  }
  if (k > n) return;
  int j = k * (k + 1) / 2;
  x[k - 1] = y[k - 1] / l[j - 1];
  if (k >= n) return;
  k = k + 1;
  for (int i = k; i <= n; ++i) {
    double t = dotprd(i - 1, &l[j], x);
    j = j + i;
    x[i - 1] = (y[i - 1] - t) / l[j - 1];
  }
}

// toms573.f90:3842-3893 -- rows n1..n of the Cholesky factor, packed by rows.
void lsqrt(int n1, int n, double* l, const double* a, int* irc) {
  int i0 = n1 * (n1 - 1) / 2;
  for (int i = n1; i <= n; ++i) {
    double td = 0.0;
    if (i != 1) {
      int j0 = 0;
      for (int j = 1; j <= i - 1; ++j) {
        double t = 0.0;
        for (int k = 1; k <= j - 1; ++k) t = t + l[i0 + k - 1] * l[j0 + k - 1];
        int ij = i0 + j;
        j0 = j0 + j;
        t = (a[ij - 1] - t) / l[j0 - 1];
        l[ij - 1] = t;
        td = td + t * t;
      }
    }
    i0 = i0 + i;
    double t = a[i0 - 1] - td;
    if (t <= 0.0) {
      l[i0 - 1] = t;
      *irc = i;
      return;
    }
    l[i0 - 1] = std::sqrt(t);
  }
  *irc = 0;
}

namespace {

// toms573.f90:3296-3331 -- lin = L**-1 (used by the covariance pass only).
void linvrt(int n, double* lin, const double* l) {
  int np1 = n + 1;
  int j0 = n * np1 / 2;
  for (int ii = 1; ii <= n; ++ii) {
    int i = np1 - ii;
    lin[j0 - 1] = 1.0 / l[j0 - 1];
    if (i <= 1) return;
    int j1 = j0;
    int im1 = i - 1;
    for (int jj = 1; jj <= im1; ++jj) {
      double t = 0.0;
      j0 = j1;
      int k0 = j1 - jj;
      for (int k = 1; k <= jj; ++k) {
        t = t - l[k0 - 1] * lin[j0 - 1];
        j0 = j0 - 1;
        k0 = k0 + k - i;
      }
      lin[j0 - 1] = t / l[k0 - 1];
    }
    j0 = j0 - 1;
  }
}

// toms573.f90:4050-4083 -- a = lower triangle of (L**T) L.
void ltsqar(int n, double* a, const double* l) {
  int ii = 0;
  for (int i = 1; i <= n; ++i) {
    int i1 = ii + 1;
    ii = ii + i;
    int m = 1;
    if (i != 1) {
      int iim1 = ii - 1;
      for (int j = i1; j <= iim1; ++j) {
        double lj = l[j - 1];
        for (int k = i1; k <= j; ++k) {
          a[m - 1] = a[m - 1] + lj * l[k - 1];
          m = m + 1;
        }
      }
    }
    double lii = l[ii - 1];
    for (int j = i1; j <= ii; ++j) a[j - 1] = lii * l[j - 1];
  }
}

// toms573.f90:3896-4047 -- over-estimate of the smallest singular value of L.
double lsvmin(int p, const double* l_, double* x_, double* y_, int* ixp) {
  const double* l = l_ - 1;
  double* x = x_ - 1;
  double* y = y_ - 1;
  int& ix = *ixp;
  int ii = 0;
  for (int i = 1; i <= p; ++i) {
    x[i] = 0.0;
    ii = ii + i;
    if (l[ii] == 0.0) return 0.0;
  }
  if (ix % 9973 == 0) ix = 2;
  int pplus1 = p + 1;
  for (int jjj = 1; jjj <= p; ++jjj) {
    int j = pplus1 - jjj;
    ix = (3432 * ix) % 9973;
    double b = 0.5 * (1.0 + (double)ix / 9973.0);
    double xplus = (b - x[j]);
    double xminus = (-b - x[j]);
    double splus = std::fabs(xplus);
    double sminus = std::fabs(xminus);
    int jm1 = j - 1;
    int j0 = j * jm1 / 2;
    int jj = j0 + j;
    xplus = xplus / l[jj];
    xminus = xminus / l[jj];
    for (int i = 1; i <= jm1; ++i) {
      int ji = j0 + i;
      splus = splus + std::fabs(x[i] + l[ji] * xplus);
      sminus = sminus + std::fabs(x[i] + l[ji] * xminus);
    }
    if (sminus > splus) xplus = xminus;
    x[j] = xplus;
    for (int i = 1; i <= jm1; ++i) {
      int ji = j0 + i;
      x[i] = x[i] + l[ji] * xplus;
    }
  }
  double t = 1.0 / v2norm(p, x_);
  for (int i = 1; i <= p; ++i) x[i] = t * x[i];
  for (int j = 1; j <= p; ++j) {
    double psj = 0.0;
    int jm1 = j - 1;
    int j0 = j * jm1 / 2;
    for (int i = 1; i <= jm1; ++i) psj = psj + l[j0 + i] * y[i];
    int jj = j0 + j;
    y[j] = (x[j] - psj) / l[jj];
  }
  return 1.0 / v2norm(p, y_);
}

// toms573.f90:4570-4592
double reldst(int p, const double* d, const double* x, const double* x0) {
  double emax = 0.0, xmax = 0.0;
  for (int i = 0; i < p; ++i) {
    double t = std::fabs(d[i] * (x[i] - x0[i]));
    if (emax < t) emax = t;
    t = d[i] * (std::fabs(x[i]) + std::fabs(x0[i]));
    if (xmax < t) xmax = t;
  }
  double r = 0.0;
  if (xmax > 0.0) r = emax / xmax;
  return r;
}

// toms573.f90:4697-4727 -- y = S x, S symmetric packed lower-by-rows.
void slvmul(int p, double* y, const double* s, const double* x) {
  int j = 1;
  for (int i = 1; i <= p; ++i) {
    y[i - 1] = dotprd(i, &s[j - 1], x);
    j = j + i;
  }
  if (p <= 1) return;
  j = 1;
  for (int i = 2; i <= p; ++i) {
    double xi = x[i - 1];
    j = j + 1;
    for (int k = 1; k <= i - 1; ++k) {
      y[k - 1] = y[k - 1] + s[j - 1] * xi;
      j = j + 1;
    }
  }
}

// toms573.f90:4658-4694 -- secant update so that A*step = y.
void slupdt(double* a, double cosmin, int p, double size, const double* step, double* u, double* w,
            const double* wchmtd, double* wscale, const double* y) {
  double sdotwm = dotprd(p, step, wchmtd);
  double denmin = cosmin * v2norm(p, step) * v2norm(p, wchmtd);
  *wscale = 1.0;
  if (denmin != 0.0) *wscale = std::fmin(1.0, std::fabs(sdotwm / denmin));
  double t = 0.0;
  if (sdotwm != 0.0) t = *wscale / sdotwm;
  for (int i = 0; i < p; ++i) w[i] = t * wchmtd[i];
  slvmul(p, u, a, step);
  t = 0.5 * (size * dotprd(p, step, u) - dotprd(p, step, y));
  for (int i = 0; i < p; ++i) u[i] = t * w[i] + y[i] - size * u[i];
  int k = 0;
  for (int i = 0; i < p; ++i) {
    double ui = u[i], wi = w[i];
    for (int j = 0; j <= i; ++j) {
      a[k] = size * a[k] + ui * w[j] + wi * u[j];
      ++k;
    }
  }
}

// toms573.f90:2434-2468 -- scale-vector update.
void dupdat(double* d, const int* iv, const double* j, int ldj, int n, int p, const double* v) {
  int i = IV(DTYPE);
  if (i != 1) {
    if (IV(NITER) > 0) return;
  }
  double vdfac = V(DFAC);
  int d0 = JTOL0 + p;
  int s1 = IV(IV_S) - 1;
  for (i = 1; i <= p; ++i) {
    s1 = s1 + i;
    double sii = V(s1);
    double t = v2norm(n, j + (size_t)(i - 1) * ldj);
    if (sii > 0.0) t = std::sqrt(t * t + sii);
    int jtoli = JTOL0 + i;
    d0 = d0 + 1;
    if (t < V(jtoli)) t = std::fmax(V(d0), V(jtoli));
    d[i - 1] = std::fmax(vdfac * d[i - 1], t);
  }
}

// toms573.f90:4595-4655 -- products with the permuted R of qrfact.
void rptmul(int func, const int* ipivot, const double* j, int ldj, int p, const double* rd,
            const double* x, double* y, double* z) {
#define JJ(i, k) j[(size_t)((k)-1) * ldj + ((i)-1)]
  if (func <= 2) {
    for (int i = 1; i <= p; ++i) z[i - 1] = x[ipivot[i - 1] - 1];
    y[0] = z[0] * rd[0];
    for (int k = 2; k <= p; ++k) {
      double zk = z[k - 1];
      for (int i = 1; i <= k - 1; ++i) y[i - 1] = y[i - 1] + JJ(i, k) * zk;
      y[k - 1] = zk * rd[k - 1];
    }
    if (func <= 1) return;
  } else {
    for (int i = 0; i < p; ++i) y[i] = x[i];
  }
  z[0] = y[0] * rd[0];
  for (int i = 2; i <= p; ++i) z[i - 1] = y[i - 1] * rd[i - 1] + dotprd(i - 1, &JJ(1, i), y);
  for (int i = 1; i <= p; ++i) y[ipivot[i - 1] - 1] = z[i - 1];
#undef JJ
}

}  // namespace

// toms573.f90:4281-4357 -- apply the Householder vectors stored by qrfact.
void qapply(int n, int p, const double* j, int ldj, double* r, int ierr) {
  int k = p;
  if (ierr != 0) k = std::abs(ierr) - 1;
  for (int l = 1; l <= k; ++l) {
    int nl1 = n - l + 1;
    const double* col = j + (size_t)(l - 1) * ldj + (l - 1);
    double t = -dotprd(nl1, col, r + (l - 1));
    for (int i = 0; i < nl1; ++i) r[l - 1 + i] = r[l - 1 + i] + t * col[i];
  }
}

// toms573.f90:4360-4567 -- Householder QR with column pivoting on down-dated norms.
void qrfact(int m, int n, double* qr, int ldq, double* alpha, int* ipivot, int* ierr, int nopivk,
            double* sum) {
#define QR(i, k) qr[(size_t)((k)-1) * ldq + ((i)-1)]
  const double ufeta = kTiny, rktol = kSqMachep;
  const double p01 = 0.01, p99 = 0.99;
  int i, j, jbar, k, k1, minum, mk1;
  double alphak, beta, qrkk, qrkmax, sigma, temp, rktol1, sumj;
  *ierr = 0;
  rktol1 = p01 * rktol;
  for (j = 1; j <= n; ++j) {
    sum[j - 1] = v2norm(m, &QR(1, j));
    ipivot[j - 1] = j;
  }
  minum = m < n ? m : n;
  for (k = 1; k <= minum; ++k) {
    mk1 = m - k + 1;
    sigma = 0.0;
    jbar = 0;
    if (k <= nopivk) goto L50;
    for (j = k; j <= n; ++j) {
      if (sigma >= sum[j - 1]) continue;
      sigma = sum[j - 1];
      jbar = j;
    }
    if (jbar == 0) goto L120;
    if (jbar == k) goto L50;
    i = ipivot[k - 1];
    ipivot[k - 1] = ipivot[jbar - 1];
    ipivot[jbar - 1] = i;
    sum[jbar - 1] = sum[k - 1];
    sum[k - 1] = sigma;
    for (i = 1; i <= m; ++i) {
      sigma = QR(i, k);
      QR(i, k) = QR(i, jbar);
      QR(i, jbar) = sigma;
    }
  L50:
    qrkmax = 0.0;
    for (i = k; i <= m; ++i)
      if (std::fabs(QR(i, k)) > qrkmax) qrkmax = std::fabs(QR(i, k));
    if (qrkmax < ufeta) goto L110;
    alphak = v2norm(mk1, &QR(k, k)) / qrkmax;
    sigma = alphak * alphak;
    qrkk = QR(k, k);
    if (qrkk >= 0.0) alphak = -alphak;
    alpha[k - 1] = alphak * qrkmax;
    beta = qrkmax * std::sqrt(sigma - (qrkk * alphak / qrkmax));
    QR(k, k) = qrkk - alpha[k - 1];
    for (i = k; i <= m; ++i) QR(i, k) = QR(i, k) / beta;
    k1 = k + 1;
    if (k1 > n) continue;
    for (j = k1; j <= n; ++j) {
      temp = -dotprd(mk1, &QR(k, k), &QR(k, j));
      vaxpy(mk1, &QR(k, j), temp, &QR(k, k), &QR(k, j));
      if (k1 > m) continue;
      sumj = sum[j - 1];
      if (sumj < ufeta) continue;
      temp = std::fabs(QR(k, j) / sumj);
      if (temp < rktol1) continue;
      if (temp >= p99) {
        sum[j - 1] = v2norm(m - k, &QR(k1, j));
        continue;
      }
      sum[j - 1] = sumj * std::sqrt(1.0 - temp * temp);
    }
  }
  return;
L110:
  *ierr = -k;
  goto L130;
L120:
  *ierr = k;
L130:
  for (i = k; i <= n; ++i) {
    alpha[i - 1] = 0.0;
    if (i > k) vscopy(i - k, &QR(k, i), 0.0);
  }
#undef QR
}

// toms573.f90:3397-3839 -- Levenberg-Marquardt step, More-Hebden, fast Givens.
// All arrays are addressed 1-based through shifted aliases, as in the reference.
void lmstep(const double* d_, const double* g_, int* ierr_, int* ipivot_, int* ka_, int p,
            const double* qtr_, const double* r_, double* step_, double* v, double* w_) {
  const double* d = d_ - 1;
  const double* g = g_ - 1;
  const int* ipivot = ipivot_ - 1;
  const double* qtr = qtr_ - 1;
  double* step = step_ - 1;
  double* w = w_ - 1;
  int& ierr = *ierr_;
  int& ka = *ka_;
  int dstsav, i, ip1, i1, j1, k, kalim, l, lk0, phipin, pp1o2, res, res0, rmat, rmat0, uk0;
  double a, adi, alphak, b, dfacsq, dst, dtol, d1, d2, lk, oldphi, phi, phimax, phimin, psifac, rad,
      si, sj, sqrtak, t, twopsi, uk, wl;
  const double dfac = 256.0, eight = 8.0, half = 0.5, negone = -1.0, one = 1.0, p001 = 1e-3,
               three = 3.0, ttol = 2.5, zero = 0.0;

  psifac = 0.0;
  alphak = 0.0;
  dst = 0.0;
  phi = 0.0;
  lk0 = p + 1;
  phipin = lk0 + 1;
  uk0 = phipin + 1;
  dstsav = uk0 + 1;
  rmat0 = dstsav;
  rmat = rmat0 + 1;
  pp1o2 = p * (p + 1) / 2;
  res0 = pp1o2 + rmat0;
  res = res0 + 1;
  rad = V(RADIUS);
  if (rad > zero) psifac = V(EPSLON) / ((eight * (V(PHMNFC) + one) + three) * (rad * rad));
  phimax = V(PHMXFC) * rad;
  phimin = V(PHMNFC) * rad;
  dtol = one / dfac;
  dfacsq = dfac * dfac;
  oldphi = zero;
  lk = zero;
  uk = zero;
  kalim = ka + 12;

  if (ka < 0) goto L10;
  if (ka == 0) goto L20;
  goto L310;

L10:
  ka = 0;
  kalim = 12;
  k = p;
  if (ierr != 0) k = std::abs(ierr) - 1;
  V(NREDUC) = half * dotprd(k, qtr_, qtr_);

L20:
  V(DST0) = negone;
  if (ierr != 0) goto L50;
  litvmu(p, w_, r_, qtr_);
  for (i = 1; i <= p; ++i) {
    j1 = ipivot[i];
    step[i] = d[j1] * w[i];
  }
  dst = v2norm(p, step_);
  V(DST0) = dst;
  phi = dst - rad;
  if (phi <= phimax) goto L350;
  if (ka > 0) goto L70;
  for (i = 1; i <= p; ++i) {
    j1 = ipivot[i];
    step[i] = d[j1] * (step[i] / dst);
  }
  livmul(p, step_, r_, step_);
  t = one / v2norm(p, step_);
  w[phipin] = (t / dst) * t;
  lk = phi * w[phipin];

L50:
  for (i = 1; i <= p; ++i) w[i] = g[i] / d[i];
  V(DGNORM) = v2norm(p, w_);
  uk = V(DGNORM) / rad;
  if (uk <= zero) goto L330;
  alphak = std::fabs(V(STPPAR)) * V(RAD0) / rad;

L70:
  ka = ka + 1;
  vcopy(pp1o2, &w[rmat], r_);
  vcopy(p, &w[res], qtr_);
  if (alphak <= zero || alphak < lk || alphak >= uk) alphak = uk * std::fmax(p001, std::sqrt(lk / uk));
  sqrtak = std::sqrt(alphak);
  for (i = 1; i <= p; ++i) w[i] = one;

  for (i = 1; i <= p; ++i) {
    l = i * (i + 1) / 2 + rmat0;
    wl = w[l];
    d2 = one;
    d1 = w[i];
    j1 = ipivot[i];
    adi = sqrtak * d[j1];
    if (adi >= std::fabs(wl)) goto L110;
  L90:
    a = adi / wl;
    b = d2 * a / d1;
    t = a * b + one;
    if (t > ttol) goto L110;
    w[i] = d1 / t;
    d2 = d2 / t;
    w[l] = t * wl;
    a = -a;
    for (j1 = i; j1 <= p; ++j1) {
      l = l + j1;
      step[j1] = a * w[l];
    }
    goto L130;
  L110:
    b = wl / adi;
    a = d1 * b / d2;
    t = a * b + one;
    if (t > ttol) goto L90;
    w[i] = d2 / t;
    d2 = d1 / t;
    w[l] = t * adi;
    for (j1 = i; j1 <= p; ++j1) {
      l = l + j1;
      wl = w[l];
      step[j1] = -wl;
      w[l] = a * wl;
    }
  L130:
    if (i == p) goto L240;
    ip1 = i + 1;
    for (i1 = ip1; i1 <= p; ++i1) {
      l = i1 * (i1 + 1) / 2 + rmat0;
      wl = w[l];
      si = step[i1 - 1];
      d1 = w[i1];
      if (d1 >= dtol) goto L150;
      d1 = d1 * dfacsq;
      wl = wl / dfac;
      k = l;
      for (j1 = i1; j1 <= p; ++j1) {
        k = k + j1;
        w[k] = w[k] / dfac;
      }
    L150:
      if (std::fabs(si) > std::fabs(wl)) goto L180;
      if (si == zero) continue;
    L160:
      a = si / wl;
      b = d2 * a / d1;
      t = a * b + one;
      if (t > ttol) goto L180;
      w[l] = t * wl;
      w[i1] = d1 / t;
      d2 = d2 / t;
      for (j1 = i1; j1 <= p; ++j1) {
        l = l + j1;
        wl = w[l];
        sj = step[j1];
        w[l] = wl + b * sj;
        step[j1] = sj - a * wl;
      }
      goto L200;
    L180:
      b = wl / si;
      a = d1 * b / d2;
      t = a * b + one;
      if (t > ttol) goto L160;
      w[i1] = d2 / t;
      d2 = d1 / t;
      w[l] = t * si;
      for (j1 = i1; j1 <= p; ++j1) {
        l = l + j1;
        wl = w[l];
        sj = step[j1];
        w[l] = a * wl + sj;
        step[j1] = b * sj - wl;
      }
    L200:
      if (d2 >= dtol) continue;
      d2 = d2 * dfacsq;
      for (k = i1; k <= p; ++k) step[k] = step[k] / dfac;
    }
  }

L240:
  litvmu(p, &w[res], &w[rmat], &w[res]);
  for (i = 1; i <= p; ++i) {
    j1 = ipivot[i];
    k = res0 + i;
    t = w[k];
    step[j1] = -t;
    w[k] = t * d[j1];
  }
  dst = v2norm(p, &w[res]);
  phi = dst - rad;
  if (phi <= phimax && phi >= phimin) goto L370;
  if (oldphi == phi) goto L370;
  oldphi = phi;
  if (phi > zero) goto L270;
  if (ka >= kalim) goto L370;
  twopsi = alphak * dst * dst - dotprd(p, step_, g_);
  if (alphak >= twopsi * psifac) goto L270;
  V(STPPAR) = -alphak;
  goto L380;

L260:
  if (phi < zero) uk = std::fmin(uk, alphak);
  goto L280;
L270:
  if (phi < zero) uk = alphak;
L280:
  for (i = 1; i <= p; ++i) {
    j1 = ipivot[i];
    k = res0 + i;
    step[i] = d[j1] * (w[k] / dst);
  }
  livmul(p, step_, &w[rmat], step_);
  for (i = 1; i <= p; ++i) step[i] = step[i] / std::sqrt(w[i]);
  t = one / v2norm(p, step_);
  alphak = alphak + t * phi * t / rad;
  lk = std::fmax(lk, alphak);
  goto L70;

L310:
  lk = w[lk0];
  uk = w[uk0];
  if (V(DST0) > zero && V(DST0) - rad <= phimax) goto L20;
  alphak = std::fabs(V(STPPAR));
  dst = w[dstsav];
  phi = dst - rad;
  t = V(DGNORM) / rad;
  if (rad > V(RAD0)) goto L320;
  uk = t;
  if (alphak <= zero) lk = zero;
  if (V(DST0) > zero) lk = std::fmax(lk, (V(DST0) - rad) * w[phipin]);
  goto L260;
L320:
  if (alphak <= zero || uk > t) uk = t;
  lk = zero;
  if (V(DST0) > zero) lk = std::fmax(lk, (V(DST0) - rad) * w[phipin]);
  goto L260;

L330:
  V(STPPAR) = zero;
  dst = zero;
  lk = zero;
  uk = zero;
  V(GTSTEP) = zero;
  V(PREDUC) = zero;
  for (i = 1; i <= p; ++i) step[i] = zero;
  goto L390;

L350:
  alphak = zero;
  for (i = 1; i <= p; ++i) {
    j1 = ipivot[i];
    step[j1] = -w[i];
  }
L370:
  V(STPPAR) = alphak;
L380:
  V(GTSTEP) = dotprd(p, step_, g_);
  V(PREDUC) = half * (alphak * dst * dst - V(GTSTEP));
L390:
  V(DSTNRM) = dst;
  w[dstsav] = dst;
  w[lk0] = lk;
  w[uk0] = uk;
  V(RAD0) = rad;
}

// toms573.f90:2471-3039 -- Goldfeld-Quandt-Trotter step for the augmented model.
// Optional guard for the checker (Options::gqt_retry_cap): the retry after a failed factorisation at
// label 200 has no iteration limit in the reference (toms573.f90:2790-2812) and is a latent
// non-termination (tests/test_robustness.py).  0 = the reference's behaviour.  When the cap is hit the
// last solved step is kept, as the product does (nl2_solver.cuh, anomaly bit 2).
thread_local int t_gqt_retry_cap = 0;

void gqtstp(const double* d_, const double* dig_, double* dihdi_, int* ka_, double* l_, int p,
            double* step_, double* v, double* w_, int* lsvmin_ix) {
  int n_retry = 0;
  bool have_q = false;
  const double* d = d_ - 1;
  const double* dig = dig_ - 1;
  double* dihdi = dihdi_ - 1;
  double* l = l_ - 1;
  double* step = step_ - 1;
  double* w = w_ - 1;
  int& ka = *ka_;
  int dggdmx, diag, diag0, dstsav, emax, emin, i, im1, inc, irc, j, k, kalim, k1, lk0, phipin, q, q0,
      uk0, x, x0;
  double alphak, aki, akk, delta, dst, epso6, lk, oldphi, phi, phimax, phimin, psifac, rad, root, si,
      sk, sw, t, twopsi, t1, uk, wi;
  bool restrt;
  const double epsfac = 50.0, four = 4.0, half = 0.5, kappa = 2.0, negone = -1.0, one = 1.0,
               p001 = 1e-3, six = 6.0, three = 3.0, two = 2.0, zero = 0.0;
  const double dgxfac = epsfac * kMachep;

  dst = 0.0;
  phi = 0.0;
  alphak = 0.0;
  uk = 0.0;
  lk = 0.0;
  delta = 0.0;
  twopsi = 0.0;
  dggdmx = p + 1;
  emax = dggdmx + 1;
  emin = emax + 1;
  lk0 = emin + 1;
  phipin = lk0 + 1;
  uk0 = phipin + 1;
  dstsav = uk0 + 1;
  diag0 = dstsav;
  diag = diag0 + 1;
  q0 = diag0 + p;
  q = q0 + 1;
  rad = V(RADIUS);
  phimax = V(PHMXFC) * rad;
  phimin = V(PHMNFC) * rad;
  psifac = two * V(EPSLON) /
           (three * (four * (V(PHMNFC) + one) * (kappa + one) + kappa + two) * (rad * rad));
  oldphi = zero;
  epso6 = V(EPSLON) / six;
  irc = 0;
  restrt = false;
  kalim = ka + 50;

  if (ka >= 0) goto L290;

  k = 0;
  uk = negone;
  ka = 0;
  kalim = 50;
  j = 0;
  for (i = 1; i <= p; ++i) {
    j = j + i;
    k1 = diag0 + i;
    w[k1] = dihdi[j];
  }
  t1 = zero;
  j = p * (p + 1) / 2;
  for (i = 1; i <= j; ++i) {
    t = std::fabs(dihdi[i]);
    if (t1 < t) t1 = t;
  }
  w[dggdmx] = t1;

L30:
  lsqrt(1, p, l_, dihdi_, &irc);
  if (irc == 0) goto L50;
  j = irc * (irc + 1) / 2;
  t = l[j];
  l[j] = one;
  for (i = 1; i <= irc; ++i) w[i] = zero;
  w[irc] = one;
  litvmu(irc, w_, l_, w_);
  t1 = v2norm(irc, w_);
  lk = -t / t1 / t1;
  V(DST0) = -lk;
  if (restrt) goto L200;
  V(NREDUC) = zero;
  goto L60;

L50:
  lk = zero;
  livmul(p, &w[q], l_, dig_);
  V(NREDUC) = half * dotprd(p, &w[q], &w[q]);
  litvmu(p, &w[q], l_, &w[q]);
  have_q = true;
  dst = v2norm(p, &w[q]);
  V(DST0) = dst;
  phi = dst - rad;
  if (phi <= phimax) goto L260;
  if (restrt) goto L200;

L60:
  V(DGNORM) = v2norm(p, dig_);
  if (V(DGNORM) == zero) goto L430;
  k = 0;
  for (i = 1; i <= p; ++i) {
    wi = zero;
    if (i != 1) {
      im1 = i - 1;
      for (j = 1; j <= im1; ++j) {
        k = k + 1;
        t = std::fabs(dihdi[k]);
        wi = wi + t;
        w[j] = w[j] + t;
      }
    }
    w[i] = wi;
    k = k + 1;
  }

  k = 1;
  t1 = w[diag] - w[1];
  if (p > 1) {
    for (i = 2; i <= p; ++i) {
      j = diag0 + i;
      t = w[j] - w[i];
      if (t >= t1) continue;
      t1 = t;
      k = i;
    }
  }
  sk = w[k];
  j = diag0 + k;
  akk = w[j];
  k1 = k * (k - 1) / 2 + 1;
  inc = 1;
  t = zero;
  for (i = 1; i <= p; ++i) {
    if (i == k) goto L120;
    aki = std::fabs(dihdi[k1]);
    si = w[i];
    j = diag0 + i;
    t1 = half * (akk - w[j] + si - aki);
    t1 = t1 + std::sqrt(t1 * t1 + sk * aki);
    if (t < t1) t = t1;
    if (i < k) goto L130;
  L120:
    inc = i;
  L130:
    k1 = k1 + inc;
  }
  w[emin] = akk - t;
  uk = V(DGNORM) / rad - w[emin];

  k = 1;
  t1 = w[diag] + w[1];
  if (p > 1) {
    for (i = 2; i <= p; ++i) {
      j = diag0 + i;
      t = w[j] + w[i];
      if (t <= t1) continue;
      t1 = t;
      k = i;
    }
  }
  sk = w[k];
  j = diag0 + k;
  akk = w[j];
  k1 = k * (k - 1) / 2 + 1;
  inc = 1;
  t = zero;
  for (i = 1; i <= p; ++i) {
    if (i == k) goto L170;
    aki = std::fabs(dihdi[k1]);
    si = w[i];
    j = diag0 + i;
    t1 = half * (w[j] + si - aki - akk);
    t1 = t1 + std::sqrt(t1 * t1 + sk * aki);
    if (t < t1) t = t1;
    if (i < k) goto L180;
  L170:
    inc = i;
  L180:
    k1 = k1 + inc;
  }
  w[emax] = akk + t;
  lk = std::fmax(lk, V(DGNORM) / rad - w[emax]);
  alphak = std::fabs(V(STPPAR)) * V(RAD0) / rad;
  if (irc != 0) goto L200;
  livmul(p, w_, l_, &w[q]);
  t = v2norm(p, w_);
  w[phipin] = dst / t / t;
  lk = std::fmax(lk, phi * w[phipin]);

L200:
  if (t_gqt_retry_cap > 0 && ++n_retry > t_gqt_retry_cap) {
    if (have_q) goto L270;
    goto L430;
  }
  ka = ka + 1;
  if (-V(DST0) >= alphak || alphak < lk || alphak >= uk)
    alphak = uk * std::fmax(p001, std::sqrt(lk / uk));
  k = 0;
  for (i = 1; i <= p; ++i) {
    k = k + i;
    j = diag0 + i;
    dihdi[k] = w[j] + alphak;
  }
  lsqrt(1, p, l_, dihdi_, &irc);
  if (irc == 0) goto L230;
  j = (irc * (irc + 1)) / 2;
  t = l[j];
  l[j] = one;
  for (i = 1; i <= irc; ++i) w[i] = zero;
  w[irc] = one;
  litvmu(irc, w_, l_, w_);
  t1 = v2norm(irc, w_);
  lk = alphak - t / t1 / t1;
  V(DST0) = -lk;
  goto L200;

L230:
  livmul(p, &w[q], l_, dig_);
  litvmu(p, &w[q], l_, &w[q]);
  have_q = true;
  dst = v2norm(p, &w[q]);
  phi = dst - rad;
  if (phi <= phimax && phi >= phimin) goto L270;
  if (phi == oldphi) goto L270;
  oldphi = phi;
  if (phi > zero) goto L240;
  if (V(DST0) > zero) goto L240;
  delta = alphak + V(DST0);
  twopsi = alphak * dst * dst + dotprd(p, dig_, &w[q]);
  if (delta < psifac * twopsi) goto L250;

L240:
  if (ka >= kalim) goto L270;
  livmul(p, w_, l_, &w[q]);
  t1 = v2norm(p, w_);
  if (phi < zero) uk = std::fmin(uk, alphak);
  alphak = alphak + (phi / t1) * (dst / t1) * (dst / rad);
  lk = std::fmax(lk, alphak);
  goto L200;

L250:
  if (delta > dgxfac * w[dggdmx]) goto L330;
  goto L270;

L260:
  alphak = zero;
L270:
  for (i = 1; i <= p; ++i) step[i] = -w[q0 + i] / d[i];
  V(GTSTEP) = -dotprd(p, dig_, &w[q]);
  V(PREDUC) = half * (std::fabs(alphak) * dst * dst - V(GTSTEP));
  goto L410;

L290:
  if (V(DST0) <= zero || V(DST0) - rad > phimax) goto L310;
  restrt = true;
  ka = ka + 1;
  k = 0;
  for (i = 1; i <= p; ++i) {
    k = k + i;
    j = diag0 + i;
    dihdi[k] = w[j];
  }
  uk = negone;
  goto L30;

L310:
  if (ka == 0) goto L50;
  dst = w[dstsav];
  alphak = std::fabs(V(STPPAR));
  phi = dst - rad;
  t = V(DGNORM) / rad;
  if (rad > V(RAD0)) goto L320;
  uk = t - w[emin];
  lk = zero;
  if (alphak > zero) lk = w[lk0];
  lk = std::fmax(lk, t - w[emax]);
  if (V(DST0) > zero) lk = std::fmax(lk, (V(DST0) - rad) * w[phipin]);
  goto L240;
L320:
  uk = t - w[emin];
  if (alphak > zero) uk = std::fmin(uk, w[uk0]);
  lk = std::fmax(std::fmax(zero, -V(DST0)), t - w[emax]);
  if (V(DST0) > zero) lk = std::fmax(lk, (V(DST0) - rad) * w[phipin]);
  goto L240;

L330:
  alphak = -alphak;
  x0 = q0 + p;
  x = x0 + 1;
  delta = kappa * delta;
  t = lsvmin(p, l_, &w[x], w_, lsvmin_ix ? lsvmin_ix : &g_lsvmin_ix);
  k = 0;
L340:
  for (i = 1; i <= p; ++i) w[i] = t * w[i];
  litvmu(p, w_, l_, w_);
  t1 = one / v2norm(p, w_);
  t = t1 * t;
  if (t <= delta) goto L370;
  if (k > 30) goto L270;
  k = k + 1;
  for (i = 1; i <= p; ++i) w[x0 + i] = t1 * w[i];
  livmul(p, w_, l_, &w[x]);
  t = one / v2norm(p, w_);
  goto L340;

L370:
  for (i = 1; i <= p; ++i) w[i] = t1 * w[i];
  sw = dotprd(p, &w[q], w_);
  t1 = (rad + dst) * (rad - dst);
  root = std::sqrt(sw * sw + t1);
  if (sw < zero) root = -root;
  si = t1 / (sw + root);
  V(PREDUC) = half * twopsi;
  t1 = zero;
  t = si * (alphak * sw - half * si * (alphak + t * dotprd(p, &w[x], w_)));
  if (t < epso6 * twopsi) goto L390;
  V(PREDUC) = V(PREDUC) + t;
  dst = rad;
  t1 = -si;
L390:
  for (i = 1; i <= p; ++i) {
    j = q0 + i;
    w[j] = t1 * w[i] - w[j];
    step[i] = w[j] / d[i];
  }
  V(GTSTEP) = dotprd(p, dig_, &w[q]);

L410:
  V(DSTNRM) = dst;
  V(STPPAR) = alphak;
  w[lk0] = lk;
  w[uk0] = uk;
  V(RAD0) = rad;
  w[dstsav] = dst;
  j = 0;
  for (i = 1; i <= p; ++i) {
    j = j + i;
    k = diag0 + i;
    dihdi[j] = w[k];
  }
  return;

L430:
  V(STPPAR) = zero;
  V(PREDUC) = zero;
  V(DSTNRM) = zero;
  V(GTSTEP) = zero;
  for (i = 1; i <= p; ++i) step[i] = zero;
}

namespace {

// toms573.f90:1447-1952 -- assess a candidate step (accept / retry / converge).
void assess(const double* d, int* iv, int p, double* step, double* stlstg, double* v, double* x,
            const double* x0) {
  bool goodx;
  int i, nfc;
  double emax, gts, reldx1, rfac1, xmax;
  const double half = 0.5, one = 1.0, two = 2.0, zero = 0.0;

  nfc = IV(NFCALL);
  IV(SWITCH) = 0;
  IV(RESTOR) = 0;
  rfac1 = one;
  goodx = true;
  emax = zero;
  i = IV(IRC);
  switch (i) {
    case 1: goto L20;
    case 2: goto L30;
    case 3: case 4: goto L10;
    case 5: goto L40;
    case 6: goto L280;
    case 7: case 8: case 9: case 10: case 11: goto L220;
    case 12: goto L170;
    default: IV(IRC) = 13; goto L300;
  }

L10:
  IV(STAGE) = 1;
  IV(RADINC) = 0;
  V(FLSTGD) = V(F0);
  if (IV(TOOBIG) == 0) goto L90;
  IV(STAGE) = -1;
  IV(XIRC) = i;
  goto L60;

L20:
  if (IV(MODEL) != IV(MLSTGD)) goto L30;
  IV(STAGE) = IV(STGLIM);
  IV(RADINC) = -1;
  goto L90;

L30:
  IV(STAGE) = IV(STAGE) + 1;

L40:
  if (IV(STAGE) > 0) goto L50;
  if (IV(TOOBIG) != 0) goto L60;
  IV(STAGE) = -IV(STAGE);
  i = IV(XIRC);
  switch (i) {
    case 1: goto L20;
    case 2: goto L30;
    case 3: case 4: goto L90;
    case 5: goto L70;
    default: break;  // falls through to L50, as an out-of-range computed goto does
  }

L50:
  if (IV(TOOBIG) == 0) goto L70;
  if (IV(RADINC) > 0) goto L80;
  IV(STAGE) = -IV(STAGE);
  IV(XIRC) = IV(IRC);

L60:
  V(RADFAC) = V(DECFAC);
  IV(RADINC) = IV(RADINC) - 1;
  IV(IRC) = 5;
  goto L300;

L70:
  if (V(V_F) < V(FLSTGD)) goto L90;
  if (IV(MODEL) == IV(MLSTGD)) goto L80;
  IV(MODEL) = IV(MLSTGD);
  IV(SWITCH) = 1;

L80:
  if (V(FLSTGD) >= V(F0)) goto L90;
  IV(RESTOR) = 1;
  V(V_F) = V(FLSTGD);
  V(PREDUC) = V(PLSTGD);
  V(GTSTEP) = V(GTSLST);
  if (IV(SWITCH) == 0) rfac1 = V(DSTNRM) / V(DSTSAV);
  V(DSTNRM) = V(DSTSAV);
  nfc = IV(NFGCAL);
  goodx = false;

L90:
  reldx1 = reldst(p, d, x, x0);
  if (goodx) goto L110;
  for (i = 0; i < p; ++i) {
    step[i] = stlstg[i];
    x[i] = x0[i] + stlstg[i];
  }

L110:
  V(FDIF) = V(F0) - V(V_F);
  if (V(FDIF) > V(TUNER2) * V(PREDUC)) goto L140;
  V(RELDX) = reldx1;
  if (V(V_F) < V(F0)) goto L120;
  IV(MLSTGD) = IV(MODEL);
  V(FLSTGD) = V(V_F);
  V(V_F) = V(F0);
  vcopy(p, x, x0);
  IV(RESTOR) = 1;
  goto L130;
L120:
  IV(NFGCAL) = nfc;
L130:
  IV(IRC) = 1;
  if (IV(STAGE) < IV(STGLIM)) goto L160;
  IV(IRC) = 5;
  IV(RADINC) = IV(RADINC) - 1;
  goto L160;

L140:
  IV(NFGCAL) = nfc;
  rfac1 = one;
  if (goodx) V(RELDX) = reldx1;
  V(DSTSAV) = V(DSTNRM);
  if (V(FDIF) > V(PREDUC) * V(TUNER1)) goto L190;
  if (IV(STAGE) >= IV(STGLIM)) goto L150;
  IV(IRC) = 2;
  goto L160;
L150:
  IV(IRC) = 4;

L160:
  IV(XIRC) = IV(IRC);
  emax = V(GTSTEP) + V(FDIF);
  V(RADFAC) = half * rfac1;
  if (emax < V(GTSTEP)) V(RADFAC) = rfac1 * std::fmax(V(RDFCMN), half * V(GTSTEP) / emax);

L170:
  if (V(RELDX) <= V(XFTOL)) goto L180;
  IV(IRC) = IV(XIRC);
  if (V(V_F) < V(F0)) goto L200;
  goto L230;
L180:
  IV(IRC) = 12;
  goto L240;

L190:
  if (V(FDIF) < (-V(TUNER3) * V(GTSTEP))) goto L210;
  if (IV(RADINC) < 0) goto L210;
  if (IV(RESTOR) == 1) goto L210;
  V(RADFAC) = V(RDFCMX);
  gts = V(GTSTEP);
  if (V(FDIF) < (half / V(RADFAC) - one) * gts)
    V(RADFAC) = std::fmax(V(INCFAC), half * gts / (gts + V(FDIF)));
  IV(IRC) = 4;
  if (V(STPPAR) == zero) goto L230;
  IV(IRC) = 5;
  IV(RADINC) = IV(RADINC) + 1;

L200:
  V(FLSTGD) = V(V_F);
  IV(MLSTGD) = IV(MODEL);
  vcopy(p, stlstg, step);
  V(DSTSAV) = V(DSTNRM);
  IV(NFGCAL) = nfc;
  V(PLSTGD) = V(PREDUC);
  V(GTSLST) = V(GTSTEP);
  goto L230;

L210:
  V(RADFAC) = one;
  IV(IRC) = 3;
  goto L230;

L220:
  IV(IRC) = IV(XIRC);
  if (V(DSTSAV) >= zero) goto L240;
  IV(IRC) = 12;
  goto L240;

L230:
  IV(XIRC) = IV(IRC);
L240:
  if (std::fabs(V(V_F)) < V(AFCTOL)) IV(IRC) = 10;
  if (half * V(FDIF) > V(PREDUC)) goto L300;
  emax = V(RFCTOL) * std::fabs(V(F0));
  if (V(DSTNRM) > V(LMAX0) && V(PREDUC) <= emax) IV(IRC) = 11;
  if (V(DST0) < zero) goto L250;
  i = 0;
  if ((V(NREDUC) > zero && V(NREDUC) <= emax) || (V(NREDUC) == zero && V(PREDUC) == zero)) i = 2;
  if (V(STPPAR) == zero && V(RELDX) <= V(XCTOL) && goodx) i = i + 1;
  if (i > 0) IV(IRC) = i + 6;

L250:
  if (std::abs(IV(IRC) - 3) > 2 && IV(IRC) != 12) goto L300;
  if (V(DSTNRM) > V(LMAX0)) goto L260;
  if (V(PREDUC) >= emax) goto L300;
  if (V(DST0) <= zero) goto L270;
  if (half * V(DST0) <= V(LMAX0)) goto L300;
  goto L270;
L260:
  if (half * V(DSTNRM) <= V(LMAX0)) goto L300;
  xmax = V(LMAX0) / V(DSTNRM);
  if (xmax * (two - xmax) * V(PREDUC) >= emax) goto L300;
L270:
  if (V(NREDUC) < zero) goto L290;
  V(GTSLST) = V(GTSTEP);
  V(DSTSAV) = V(DSTNRM);
  if (IV(IRC) == 12) V(DSTSAV) = -V(DSTSAV);
  V(PLSTGD) = V(PREDUC);
  IV(IRC) = 6;
  vcopy(p, stlstg, step);
  goto L300;

L280:
  V(GTSTEP) = V(GTSLST);
  V(DSTNRM) = std::fabs(V(DSTSAV));
  vcopy(p, step, stlstg);
  IV(IRC) = IV(XIRC);
  if (V(DSTSAV) <= zero) IV(IRC) = 12;
  V(NREDUC) = -V(PREDUC);
  V(PREDUC) = V(PLSTGD);
L290:
  if (-V(NREDUC) <= V(RFCTOL) * std::fabs(V(F0))) IV(IRC) = 11;

L300:
  return;
}

// toms573.f90:1955-2334 -- covariance matrix (never read by the reference's
// bootstrap: dead work kept only so that CPU-baseline timing is cost-faithful).
void covclc(int* covirc_, const double* d_, int* iv, double* j, int ldj, int n, int p, double* r,
            double* v, double* x_) {
#define JM(i, k) j[(size_t)((k)-1) * ldj + ((i)-1)]
  int& covirc = *covirc_;
  const double* d = d_ - 1;
  double* x = x_ - 1;
  int cov, gp, gsave1, g1, hc, hmi, hpi, hpm, i, ipivi, ipivk, ip1, irc, k, kind, kl, l, m, mm1,
      mm1o2, pp1o2, qtr1, rd1, stpi, stpm, stp0, wl, w0, w1;
  double del, t, wk;
  bool havej;
  const double half = 0.5, negpt5 = -0.5, one = 1.0, two = 2.0, zero = 0.0;

  cov = IV(LMAT);
  covirc = 4;
  kind = IV(COVREQ);
  m = IV(MODE);
  if (m > 0) goto L10;
  IV(KAGQT) = -1;
  if (IV(KALM) > 0) IV(KALM) = 0;
  if (std::abs(kind) >= 3) goto L310;
  V(V_FX) = V(V_F);
  k = IV(RSAVE);
  vcopy(n, &V(k), r);
L10:
  if (m > p) goto L220;
  if (kind < 0) goto L110;

  // finite-difference Hessian from function and gradient values (kind >= 0;
  // not reachable through nl2sno, which forces kind < 0 at toms573.f90:687)
  gsave1 = IV(IV_W) + p;
  g1 = IV(IV_G);
  if (m > 0) goto L20;
  vcopy(p, &V(gsave1), &V(g1));
  IV(SWITCH) = IV(NFGCAL);
  goto L90;
L20:
  del = V(V_DELTA);
  x[m] = V(XMSAVE);
  if (IV(TOOBIG) == 0) goto L40;
  if (del * x[m] > zero) goto L30;
  IV(COVMAT) = -2;
  goto L210;
L30:
  del = negpt5 * del;
  goto L100;
L40:
  cov = IV(LMAT);
  gp = g1 + p - 1;
  for (i = g1; i <= gp; ++i) {
    V(i) = (V(i) - V(gsave1)) / del;
    gsave1 = gsave1 + 1;
  }
  k = cov + m * (m - 1) / 2;
  l = k + m - 2;
  if (m == 1) goto L70;
  for (i = k; i <= l; ++i) {
    V(i) = half * (V(i) + V(g1));
    g1 = g1 + 1;
  }
L70:
  l = l + 1;
  for (i = m; i <= p; ++i) {
    V(l) = V(g1);
    l = l + i;
    g1 = g1 + 1;
  }
L90:
  m = m + 1;
  IV(MODE) = m;
  if (m > p) goto L210;
  del = V(DELTA0) * std::fmax(one / d[m], std::fabs(x[m]));
  if (x[m] < zero) del = -del;
  V(XMSAVE) = x[m];
L100:
  x[m] = x[m] + del;
  V(V_DELTA) = del;
  covirc = 2;
  goto L390;

  // finite-difference Hessian from function values only
L110:
  stp0 = IV(IV_W) + p - 1;
  mm1 = m - 1;
  mm1o2 = m * mm1 / 2;
  if (m > 0) goto L120;
  IV(SAVEI) = 0;
  goto L200;
L120:
  i = IV(SAVEI);
  if (i > 0) goto L180;
  if (IV(TOOBIG) == 0) goto L140;
  stpm = stp0 + m;
  del = V(stpm);
  // The reference indexes x(xmsave) here (toms573.f90:2099, 2106), a slip for
  // v(xmsave); the branch needs an overflowing residual and is never taken on
  // this path, so the evident intent is restated.
  if (del * V(XMSAVE) > zero) goto L130;
  IV(COVMAT) = -2;
  goto L390;
L130:
  del = negpt5 * del;
  x[m] = V(XMSAVE) + del;
  V(stpm) = del;
  covirc = 1;
  goto L390;
L140:
  pp1o2 = p * (p - 1) / 2;
  cov = IV(LMAT);
  hpm = cov + pp1o2 + mm1;
  V(hpm) = V(V_F);
  hmi = cov + mm1o2;
  if (mm1 == 0) goto L160;
  hpi = cov + pp1o2;
  for (i = 1; i <= mm1; ++i) {
    V(hmi) = V(V_FX) - (V(V_F) + V(hpi));
    hmi = hmi + 1;
    hpi = hpi + 1;
  }
L160:
  V(hmi) = V(V_F) - two * V(V_FX);
  i = 1;
L170:
  IV(SAVEI) = i;
  stpi = stp0 + i;
  V(V_DELTA) = x[i];
  x[i] = x[i] + V(stpi);
  if (i == m) x[i] = V(XMSAVE) - V(stpi);
  covirc = 1;
  goto L390;
L180:
  x[i] = V(V_DELTA);
  if (IV(TOOBIG) == 0) goto L190;
  IV(COVMAT) = -2;
  goto L390;
L190:
  cov = IV(LMAT);
  stpi = stp0 + i;
  hmi = cov + mm1o2 + i - 1;
  stpm = stp0 + m;
  V(hmi) = (V(hmi) + V(V_F)) / (V(stpi) * V(stpm));
  i = i + 1;
  if (i <= m) goto L170;
  IV(SAVEI) = 0;
  x[m] = V(XMSAVE);
L200:
  m = m + 1;
  IV(MODE) = m;
  if (m > p) goto L210;
  del = V(DLTFDC) * std::fmax(one / d[m], std::fabs(x[m]));
  if (x[m] < zero) del = -del;
  V(XMSAVE) = x[m];
  x[m] = x[m] + del;
  stpm = stp0 + m;
  V(stpm) = del;
  covirc = 1;
  goto L390;

L210:
  k = IV(RSAVE);
  vcopy(n, r, &V(k));
  V(V_F) = V(V_FX);
  if (kind < 0) goto L220;
  IV(NFGCAL) = IV(SWITCH);
  qtr1 = IV(QTR);
  vcopy(n, &V(qtr1), r);
  if (IV(COVMAT) < 0) goto L390;
  covirc = 3;
  goto L390;

L220:
  cov = IV(LMAT);
  hc = cov;
  if (std::abs(kind) == 2) goto L230;
  hc = std::abs(IV(IV_H));
  IV(IV_H) = -hc;
L230:
  lsqrt(1, p, &V(hc), &V(cov), &irc);
  IV(COVMAT) = -1;
  if (irc != 0) goto L390;
  w1 = IV(IV_W) + p;
  if (std::abs(kind) > 1) goto L340;
  vscopy(p * (p + 1) / 2, &V(cov), zero);
  havej = IV(KALM) == (-1);
  m = p;
  if (havej) m = n;
  w0 = w1 - 1;
  rd1 = IV(RD);
  for (i = 1; i <= m; ++i) {
    if (havej) goto L250;
    vscopy(p, &V(w1), zero);
    ipivi = IPIV0 + i;
    l = w0 + IV(ipivi);
    V(l) = V(rd1);
    rd1 = rd1 + 1;
    if (i == p) goto L270;
    ip1 = i + 1;
    for (k = ip1; k <= p; ++k) {
      ipivk = IPIV0 + k;
      l = w0 + IV(ipivk);
      V(l) = JM(i, k);
    }
    goto L270;
  L250:
    l = w0;
    for (k = 1; k <= p; ++k) {
      l = l + 1;
      V(l) = JM(i, k);
    }
  L270:
    livmul(p, &V(w1), &V(hc), &V(w1));
    litvmu(p, &V(w1), &V(hc), &V(w1));
    kl = cov;
    for (k = 1; k <= p; ++k) {
      l = w0 + k;
      wk = V(l);
      for (l = 1; l <= k; ++l) {
        wl = w0 + l;
        V(kl) = V(kl) + wk * V(wl);
        kl = kl + 1;
      }
    }
  }
  goto L370;

L310:
  rd1 = IV(RD);
  if (IV(KALM) != (-1)) goto L320;
  qtr1 = IV(QTR);
  vcopy(n, &V(qtr1), r);
  w1 = IV(IV_W) + p;
  qrfact(n, p, j, ldj, &V(rd1), &IV(IPIVOT), &IV(IERR), 0, &V(w1));
  IV(KALM) = -2;
L320:
  IV(COVMAT) = -1;
  if (IV(IERR) != 0) goto L390;
  cov = IV(LMAT);
  hc = std::abs(IV(IV_H));
  IV(IV_H) = -hc;
  l = hc;
  for (i = 1; i <= p; ++i) {
    if (i > 1) vcopy(i - 1, &V(l), &JM(1, i));
    l = l + i - 1;
    V(l) = V(rd1);
    l = l + 1;
    rd1 = rd1 + 1;
  }
L340:
  linvrt(p, &V(hc), &V(hc));
  ltsqar(p, &V(hc), &V(hc));
  if (hc == cov) goto L370;
  for (i = 1; i <= p; ++i) {
    m = IPIV0 + i;
    ipivi = IV(m);
    kl = cov - 1 + ipivi * (ipivi - 1) / 2;
    for (k = 1; k <= i; ++k) {
      m = IPIV0 + k;
      ipivk = IV(m);
      l = kl + ipivk;
      if (ipivk > ipivi) l = l + (ipivk - ipivi) * (ipivk + ipivi - 3) / 2;
      V(l) = V(hc);
      hc = hc + 1;
    }
  }
L370:
  IV(COVMAT) = cov;
  t = V(V_F) / (half * (double)((n - p) > 1 ? (n - p) : 1));
  k = cov - 1 + p * (p + 1) / 2;
  for (i = cov; i <= k; ++i) V(i) = t * V(i);
L390:
  return;
#undef JM
}

// toms573.f90:772-1444 -- the reverse-communication iteration driver.
// `j` is the n x p Jacobian (column-major, leading dimension nn).
void nl2itr_body(double* d_, int* iv, double* j, int n, int nn, int p, double* r, double* v,
                 double* x, const Options& opt) {
#define JCOL(k) (j + (size_t)((k)-1) * nn)
#define JM(i, k) j[(size_t)((k)-1) * nn + ((i)-1)]
  double* d = d_ - 1;
  int dig1, g1, g01, h0, h1, i, im1, ipivi, ipivk, ipiv1, ipk, k, km1, l, lky1, lmat1, lstgst, m,
      pp1o2, qtr1, rdk, rd0, rd1, rsave1, smh, sstep, step1, stpmod, s1, temp1, temp2, w1, x01;
  double e, rdof1, sttsst, t, t1;
  const double half = 0.5, negone = -1.0, one = 1.0, zero = 0.0;

  pp1o2 = p * (p + 1) / 2;
  i = IV(1);
  if (i == 1) goto L20;
  if (i == 2) goto L50;

  // parchk (toms573.f90:4086-4278): with the defaults of dfault every range
  // check passes and nothing is printed (prunit = 0); its only other effect,
  // a stash of defaults in iv(21:)/v(33:), lands in slots that the
  // initialisation below overwrites.  A fresh start (iv(1) = 12) goes to 10.
  if (IV(1) != 12) {
    i = IV(1) - 2;
    if (i >= 1 && i <= 6) goto L300;
    if (i == 7 || i == 9) goto L150;
    if (i == 8) goto L100;
    if (i != 10) goto L590;
  }

  // L10: initialisation and storage allocation
  IV(NITER) = 0;
  IV(NFCALL) = 1;
  IV(NGCALL) = 1;
  IV(NFGCAL) = 1;
  IV(MODE) = -1;
  IV(STGLIM) = 2;
  IV(TOOBIG) = 0;
  IV(CNVCOD) = 0;
  IV(COVMAT) = 0;
  IV(NFCOV) = 0;
  IV(NGCOV) = 0;
  IV(KALM) = -1;
  IV(RADINC) = 0;
  IV(IV_S) = JTOL1 + 2 * p;
  IV(X0) = IV(IV_S) + pp1o2;
  IV(STEP) = IV(X0) + p;
  IV(STLSTG) = IV(STEP) + p;
  IV(DIG) = IV(STLSTG) + p;
  IV(IV_G) = IV(DIG) + p;
  IV(LKY) = IV(IV_G) + p;
  IV(RD) = IV(LKY) + p;
  IV(RSAVE) = IV(RD) + p;
  IV(QTR) = IV(RSAVE) + n;
  IV(IV_H) = IV(QTR) + n;
  IV(IV_W) = IV(IV_H) + pp1o2;
  IV(LMAT) = IV(IV_W) + 4 * p + 7;
  if (V(DINIT) >= zero) vscopy(p, d_, V(DINIT));
  if (V(JTINIT) > zero) vscopy(p, &V(JTOL1), V(JTINIT));
  i = JTOL1 + p;
  if (V(D0INIT) > zero) vscopy(p, &V(i), V(D0INIT));
  V(RAD0) = zero;
  V(STPPAR) = zero;
  V(RADIUS) = V(LMAX0) / (one + V(PHMXFC));
  IV(MODEL) = 1;
  if (IV(INITS) == 2) IV(MODEL) = 2;
  s1 = IV(IV_S);
  if (IV(INITS) == 0) vscopy(pp1o2, &V(s1), zero);

L20:
  t = v2norm(n, r);
  if (t > V(RLIMIT)) IV(TOOBIG) = 1;
  if (IV(TOOBIG) != 0) goto L30;
  V(V_F) = half * (t * t);
L30:
  if (IV(MODE) < 0) goto L40;
  if (IV(MODE) == 0) goto L300;
  goto L540;

L40:
  if (IV(TOOBIG) == 0) goto L60;
  IV(1) = 13;
  goto L580;

L50:
  if (IV(NFGCAL) != 0) goto L60;
  IV(1) = 15;
  goto L580;

L60:
  IV(KALM) = -1;
  g1 = IV(IV_G);
  for (i = 1; i <= p; ++i) {
    V(g1) = dotprd(n, r, JCOL(i));
    g1 = g1 + 1;
  }
  if (IV(MODE) > 0) goto L520;
  if (IV(DTYPE) > 0) dupdat(d_, iv, j, nn, n, p, v);
  rsave1 = IV(RSAVE);
  vcopy(n, &V(rsave1), r);
  qtr1 = IV(QTR);
  vcopy(n, &V(qtr1), r);
  g1 = IV(IV_G);
  dig1 = IV(DIG);
  k = dig1;
  for (i = 1; i <= p; ++i) {
    V(k) = V(g1) / d[i];
    k = k + 1;
    g1 = g1 + 1;
  }
  V(DGNORM) = v2norm(p, &V(dig1));
  if (IV(CNVCOD) != 0) goto L510;
  if (IV(MODE) == 0) goto L460;
  IV(MODE) = 0;

  // ----- main loop -----
L90:  // itsmry prints nothing (prunit = 0, toms573.f90:3071-3072)
L100:
  k = IV(NITER);
  if (k < IV(MXITER)) goto L110;
  IV(1) = 10;
  goto L580;
L110:
  IV(NITER) = k + 1;
  if (k == 0) goto L130;
  step1 = IV(STEP);
  for (i = 1; i <= p; ++i) {
    V(step1) = d[i] * V(step1);
    step1 = step1 + 1;
  }
  step1 = IV(STEP);
  V(RADIUS) = V(RADFAC) * v2norm(p, &V(step1));

L130:
  x01 = IV(X0);
  V(F0) = V(V_F);
  IV(KAGQT) = -1;
  IV(IRC) = 4;
  IV(IV_H) = -std::abs(IV(IV_H));
  IV(SUSED) = IV(MODEL);
  vcopy(p, &V(x01), x);

L140:  // stopx() is always .false. here (toms573.f90:30, 4744-4745)
  goto L160;

L150:
  if (V(V_F) >= V(F0)) goto L160;
  V(RADFAC) = one;
  k = IV(NITER);
  goto L110;

L160:
  if (IV(NFCALL) < IV(MXFCAL) + IV(NFCOV)) goto L180;
  IV(1) = 9;
  // L170:
  if (V(V_F) >= V(F0)) goto L580;
  IV(CNVCOD) = IV(1);
  goto L450;

L180:
  step1 = IV(STEP);
  w1 = IV(IV_W);
  if (IV(MODEL) == 2) goto L220;

  // Levenberg-Marquardt step
  qtr1 = IV(QTR);
  if (IV(KALM) >= 0) goto L190;
  rd1 = IV(RD);
  if (-1 == IV(KALM)) qrfact(n, p, j, nn, &V(rd1), &IV(IPIVOT), &IV(IERR), 0, &V(w1));
  qapply(n, p, j, nn, &V(qtr1), IV(IERR));
L190:
  h1 = IV(IV_H);
  if (h1 > 0) goto L210;
  h1 = -h1;
  IV(IV_H) = h1;
  k = h1;
  rd1 = IV(RD);
  V(k) = V(rd1);
  if (p == 1) goto L210;
  for (i = 2; i <= p; ++i) {
    vcopy(i - 1, &V(k + 1), JCOL(i));
    k = k + i;
    rd1 = rd1 + 1;
    V(k) = V(rd1);
  }
L210:
  g1 = IV(IV_G);
  lmstep(d_, &V(g1), &IV(IERR), &IV(IPIVOT), &IV(KALM), p, &V(qtr1), &V(h1), &V(step1), v, &V(w1));
  goto L290;

  // Goldfeld-Quandt-Trotter step (augmented model)
L220:
  if (IV(IV_H) > 0) goto L280;
  h1 = -IV(IV_H);
  IV(IV_H) = h1;
  s1 = IV(IV_S);
  if (-1 != IV(KALM)) goto L250;
  for (i = 1; i <= p; ++i) {
    t = one / d[i];
    for (k = 1; k <= i; ++k) {
      V(h1) = t * (dotprd(n, JCOL(i), JCOL(k)) + V(s1)) / d[k];
      h1 = h1 + 1;
      s1 = s1 + 1;
    }
  }
  goto L280;

L250:
  smh = s1 - h1;
  h0 = h1 - 1;
  ipiv1 = IV(IPIVOT);
  t1 = one / d[ipiv1];
  rd0 = IV(RD) - 1;
  rdof1 = V(rd0 + 1);
  for (i = 1; i <= p; ++i) {
    l = IPIV0 + i;
    ipivi = IV(l);
    h1 = h0 + ipivi * (ipivi - 1) / 2;
    l = h1 + ipivi;
    m = l + smh;
    t = one / d[ipivi];
    rdk = rd0 + i;
    e = V(rdk) * V(rdk);
    if (i > 1) e = e + dotprd(i - 1, JCOL(i), JCOL(i));
    V(l) = (e + V(m)) * (t * t);
    if (i == 1) continue;
    l = h1 + ipiv1;
    if (ipivi < ipiv1) l = l + ((ipiv1 - ipivi) * (ipiv1 + ipivi - 3)) / 2;
    m = l + smh;
    V(l) = t * (rdof1 * JM(1, i) + V(m)) * t1;
    if (i == 2) continue;
    im1 = i - 1;
    for (k = 2; k <= im1; ++k) {
      ipk = IPIV0 + k;
      ipivk = IV(ipk);
      l = h1 + ipivk;
      if (ipivi < ipivk) l = l + ((ipivk - ipivi) * (ipivk + ipivi - 3)) / 2;
      m = l + smh;
      km1 = k - 1;
      rdk = rd0 + k;
      V(l) = t * (dotprd(km1, JCOL(i), JCOL(k)) + V(rdk) * JM(k, i) + V(m)) / d[ipivk];
    }
  }

L280:
  h1 = IV(IV_H);
  dig1 = IV(DIG);
  lmat1 = IV(LMAT);
  gqtstp(d_, &V(dig1), &V(h1), &IV(KAGQT), &V(lmat1), p, &V(step1), v, &V(w1), opt.lsvmin_ix);

L290:
  if (IV(IRC) == 6) goto L300;
  x01 = IV(X0);
  step1 = IV(STEP);
  vaxpy(p, x, one, &V(step1), &V(x01));
  IV(NFCALL) = IV(NFCALL) + 1;
  IV(1) = 1;
  IV(TOOBIG) = 0;
  goto L590;

L300:
  step1 = IV(STEP);
  lstgst = IV(STLSTG);
  x01 = IV(X0);
  assess(d_, iv, p, &V(step1), &V(lstgst), v, x, &V(x01));
  if (opt.trace && opt.trace_len && *opt.trace_len < opt.trace_cap) {
    double* t = opt.trace + (size_t)(*opt.trace_len) * TRACE_W;
    double piv = 0.0;
    for (int q = 0; q < p; ++q) piv = piv * p + ((&IV(IPIVOT))[q] - 1);
    t[0] = IV(IRC);
    t[1] = IV(SUSED);
    t[2] = IV(NFCALL);
    t[3] = IV(NITER);
    t[4] = V(V_F);
    t[5] = V(RELDX);
    t[6] = V(RADIUS);
    t[7] = piv;
    t[8] = IV(RESTOR);
    t[9] = V(NREDUC);
    ++*opt.trace_len;
  }
  if (IV(SWITCH) == 0) goto L310;
  IV(IV_H) = -std::abs(IV(IV_H));
  IV(SUSED) = IV(SUSED) + 2;
  vcopy(NVSAVE, &V(1), &V(VSAVE1));
L310:
  if (IV(RESTOR) == 0) goto L320;
  rsave1 = IV(RSAVE);
  vcopy(n, r, &V(rsave1));
L320:
  l = IV(IRC) - 4;
  stpmod = IV(MODEL);
  if (l > 0) {
    switch (l) {
      case 1: goto L340;
      case 2: goto L360;
      case 3: case 4: case 5: case 6: case 7: case 8: goto L370;
      case 9: goto L500;
      case 10: goto L460;
      default: break;
    }
  }

  // decide whether to change models
  e = V(PREDUC) - V(FDIF);
  sstep = IV(LKY);
  s1 = IV(IV_S);
  slvmul(p, &V(sstep), &V(s1), &V(step1));
  sttsst = half * dotprd(p, &V(step1), &V(sstep));
  if (IV(MODEL) == 1) sttsst = -sttsst;
  if (std::fabs(e + sttsst) * V(FUZZ) >= std::fabs(e)) goto L330;

  IV(MODEL) = 3 - IV(MODEL);
  if (IV(MODEL) == 1) IV(KAGQT) = -1;
  if (IV(MODEL) == 2 && IV(KALM) > 0) IV(KALM) = 0;
  if (-2 < l) goto L380;
  IV(IV_H) = -std::abs(IV(IV_H));
  IV(SUSED) = IV(SUSED) + 2;
  vcopy(NVSAVE, &V(VSAVE1), &V(1));
  goto L350;

L330:
  if (-3 < l) goto L380;
  V(RADIUS) = V(RADFAC) * V(DSTNRM);
  goto L140;

L340:
  V(RADIUS) = V(RADFAC) * V(DSTNRM);
L350:
  if (V(V_F) >= V(F0)) goto L140;
  rsave1 = IV(RSAVE);
  vcopy(n, &V(rsave1), r);
  goto L140;

L360:
  V(RADIUS) = V(LMAX0);
  goto L180;

L370:
  IV(CNVCOD) = l;
  if (V(V_F) >= V(F0)) goto L510;
  if (IV(XIRC) == 14) goto L510;
  IV(XIRC) = 14;

L380:
  IV(COVMAT) = 0;
  lky1 = IV(LKY);
  if (IV(KALM) >= 0) goto L400;
  for (i = 1; i <= p; ++i) {
    V(lky1) = dotprd(n, JCOL(i), r);
    lky1 = lky1 + 1;
  }
  goto L410;
L400:
  qtr1 = IV(QTR);
  vcopy(n, &V(qtr1), r);
  qapply(n, p, j, nn, &V(qtr1), IV(IERR));
  rd1 = IV(RD);
  temp1 = IV(STLSTG);
  rptmul(3, &IV(IPIVOT), j, nn, p, &V(rd1), &V(qtr1), &V(lky1), &V(temp1));

L410:
  if (IV(IRC) != 3) goto L450;
  step1 = IV(STEP);
  temp1 = IV(STLSTG);
  temp2 = IV(X0);
  if (stpmod == 2) goto L420;
  rd1 = IV(RD);
  rptmul(2, &IV(IPIVOT), j, nn, p, &V(rd1), &V(step1), &V(temp1), &V(temp2));
  goto L450;
L420:
  h1 = IV(IV_H);
  k = temp2;
  for (i = 1; i <= p; ++i) {
    V(k) = d[i] * V(step1);
    k = k + 1;
    step1 = step1 + 1;
  }
  slvmul(p, &V(temp1), &V(h1), &V(temp2));
  for (i = 1; i <= p; ++i) {
    V(temp1) = d[i] * V(temp1);
    temp1 = temp1 + 1;
  }

L450:
  IV(NGCALL) = IV(NGCALL) + 1;
  g1 = IV(IV_G);
  g01 = IV(IV_W);
  vcopy(p, &V(g01), &V(g1));
  IV(1) = 2;
  goto L590;

L460:
  g01 = IV(IV_W);
  g1 = IV(IV_G);
  vaxpy(p, &V(g01), negone, &V(g01), &V(g1));
  step1 = IV(STEP);
  temp1 = IV(STLSTG);
  temp2 = IV(X0);
  if (IV(IRC) != 3) goto L490;
  k = temp1;
  l = g01;
  for (i = 1; i <= p; ++i) {
    V(k) = (V(k) - V(l)) / d[i];
    k = k + 1;
    l = l + 1;
  }
  if (v2norm(p, &V(temp1)) <= V(DGNORM) * V(TUNER4)) goto L480;
  if (dotprd(p, &V(g1), &V(step1)) >= V(GTSTEP) * V(TUNER5)) goto L490;
L480:
  V(RADFAC) = V(INCFAC);

L490:
  lky1 = IV(LKY);
  vaxpy(p, &V(lky1), negone, &V(lky1), &V(g1));
  s1 = IV(IV_S);
  slvmul(p, &V(temp1), &V(s1), &V(step1));
  t1 = std::fabs(dotprd(p, &V(step1), &V(temp1)));
  t = std::fabs(dotprd(p, &V(step1), &V(lky1)));
  V(V_SIZE) = one;
  if (t < t1) V(V_SIZE) = t / t1;
  slupdt(&V(s1), V(COSMIN), p, V(V_SIZE), &V(step1), &V(temp1), &V(temp2), &V(g01), &V(WSCALE),
         &V(lky1));
  IV(1) = 2;
  goto L90;

L500:
  IV(1) = 14;
  goto L580;

L510:
  if (IV(COVREQ) == 0 && IV(COVPRT) == 0) goto L570;
  if (IV(COVMAT) != 0) goto L570;
  if (IV(CNVCOD) >= 7) goto L570;
  IV(MODE) = 0;
L520:
  covclc(&i, d_, iv, j, nn, n, p, r, v, x);
  switch (i) {
    case 1:
    case 2:
      IV(NFCOV) = IV(NFCOV) + 1;
      IV(NFCALL) = IV(NFCALL) + 1;
      IV(RESTOR) = i;
      IV(1) = 1;
      goto L590;
    case 3:
      goto L550;
    case 4:
      IV(MODE) = 0;
      if (IV(NITER) == 0) IV(MODE) = -1;
      goto L570;
    default:
      goto L540;  // a SELECT CASE with no match falls out of the construct
  }

L540:
  if (IV(RESTOR) == 1 || IV(TOOBIG) != 0) goto L520;
  IV(NFGCAL) = IV(NFCALL);
L550:
  IV(NGCOV) = IV(NGCOV) + 1;
  IV(NGCALL) = IV(NGCALL) + 1;
  IV(1) = 2;
  goto L590;

L570:
  IV(1) = IV(CNVCOD);
  IV(CNVCOD) = 0;
L580:  // itsmry: silent
L590:
  return;
#undef JCOL
#undef JM
}

void nl2itr(double* d, int* iv, double* jac, int n, int nn, int p, double* r, double* v, double* x,
            const Options& opt) {
  if (!opt.copy_jac) {
    nl2itr_body(d, iv, jac, n, nn, p, r, v, x, opt);
    return;
  }
  // toms573.f90:876-884 / 1434-1442: the reference allocates j(nn,p), copies
  // jac in, works on j and copies it back on every call.
  double* jcopy = (double*)std::malloc(sizeof(double) * (size_t)nn * p);
  std::memcpy(jcopy, jac, sizeof(double) * (size_t)nn * p);
  nl2itr_body(d, iv, jcopy, n, nn, p, r, v, x, opt);
  std::memcpy(jac, jcopy, sizeof(double) * (size_t)nn * p);
  std::free(jcopy);
}

}  // namespace

// toms573.f90:590-769 -- forward-difference Jacobian + reverse-communication loop.
void nl2sno(int n, int p, double* x, calcr_fn calcr, int* iv, double* v, void* ctx,
            const Options& opt) {
  t_gqt_retry_cap = opt.gqt_retry_cap;
  const double hfac = 1e3, negpt5 = -0.5, one = 1.0, zero = 0.0;
  bool strted;
  int dk, d1, i, j1, j1k, k, nf, rn, r1;
  double h, xk;

  IV(NITER) = 0;  // "Initialize the iteration counter to zero here. (NDD)", :673
  d1 = 94 + 2 * n + p * (3 * p + 31) / 2;
  IV(IV_D) = d1;
  r1 = d1 + p;
  IV(IV_R) = r1;
  j1 = r1 + n;
  IV(IV_J) = j1;
  rn = j1 - 1;
  if (IV(1) == 0) dfault(iv, v);
  IV(COVREQ) = -std::abs(IV(COVREQ));
  if (IV(COVPRT) != 0 && IV(COVREQ) == 0) IV(COVREQ) = -1;
  strted = true;
  if (IV(1) != 12) goto L80;
  strted = false;
  IV(NFCALL) = 1;
  IV(NFGCAL) = 1;
  if (IV(DTYPE) > 0) vscopy(p, &V(d1), one);
  if (V(DINIT) > zero) vscopy(p, &V(d1), V(DINIT));

L10:
  nf = IV(NFCALL);
  calcr(n, p, x, &nf, &V(r1), ctx);
  if (strted) goto L20;
  if (nf > 0) goto L30;
  IV(1) = 13;
  goto L100;
L20:
  if (nf <= 0) IV(TOOBIG) = 1;
  goto L80;

L30:
  j1k = j1;
  dk = d1;
  for (k = 1; k <= p; ++k) {
    xk = x[k - 1];
    h = V(DLTFDJ) * std::fmax(std::fabs(xk), one / V(dk));
    dk = dk + 1;
    for (;;) {
      x[k - 1] = xk + h;
      nf = IV(NFGCAL);
      calcr(n, p, x, &nf, &V(j1k), ctx);
      if (nf > 0) break;
      h = negpt5 * h;
      if (std::fabs(h) >= hfac * kMachep) continue;
      IV(1) = 15;
      goto L100;
    }
    x[k - 1] = xk;
    for (i = r1; i <= rn; ++i) {
      V(j1k) = (V(j1k) - V(i)) / h;
      j1k = j1k + 1;
    }
  }
  strted = true;

L80:
  nl2itr(&V(d1), iv, &V(j1), n, n, p, &V(r1), v, x, opt);
  if (IV(1) < 2) goto L10;
  if (IV(1) == 2) goto L30;
L100:
  return;
}

// ---------------------------------------------------------------- nl2sol
// toms573.f90:35-587 (executable part 512-551).  Same reverse-communication loop as nl2sno
// with the finite-difference block replaced by one call of the user's calcj.
void nl2sol(int n, int p, double* x, calcr_fn calcr, calcj_fn calcj, int* iv, double* v, void* ctx,
            const Options& opt) {
  t_gqt_retry_cap = opt.gqt_retry_cap;
  bool strted;
  int d1, j1, nf, r1;

  d1 = 94 + 2 * n + p * (3 * p + 31) / 2;  // :512
  IV(IV_D) = d1;
  r1 = d1 + p;
  IV(IV_R) = r1;
  j1 = r1 + n;
  IV(IV_J) = j1;
  strted = true;
  if (IV(1) != 0 && IV(1) != 12) goto L40;  // :519
  strted = false;
  IV(NFCALL) = 1;
  IV(NFGCAL) = 1;

L10:
  nf = IV(NFCALL);
  calcr(n, p, x, &nf, &V(r1), ctx);
  if (strted) goto L20;
  if (nf > 0) goto L30;
  IV(1) = 13;
  goto L70;  // :528-529 (itsmry prints only)
L20:
  if (nf <= 0) IV(TOOBIG) = 1;
  goto L40;

L30:
  calcj(n, p, x, &IV(NFGCAL), &V(j1), ctx);  // :535
  if (IV(NFGCAL) == 0) {
    IV(1) = 15;
    goto L70;
  }
  strted = true;

L40:
  nl2itr(&V(d1), iv, &V(j1), n, n, p, &V(r1), v, x, opt);
  if (IV(1) < 2) goto L10;
  if (IV(1) == 2) goto L30;
L70:
  return;
}

}  // namespace nl2
