#include "report.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>

namespace fmc_host {

// gfortran prints list-directed REAL(8) as G25.17E3: F-form with 17 significant digits when
// 0.1 <= |v| < 1e17 (padded with 5 blanks on the right where the exponent would be), else
// ES-form 0.dddE+xxx-like "d.ddddddddddddddddE+xxx".
std::string list_directed_real(double v) {
  char buf[64];
  if (v == 0.0) return "   0.0000000000000000    ";
  const double a = std::fabs(v);
  if (a >= 0.1 && a < 1e17) {
    int intdigits = a < 1.0 ? 0 : (int)std::floor(std::log10(a)) + 1;
    int decimals = 17 - (intdigits > 0 ? intdigits : 0);
    if (intdigits == 0) decimals = 17;
    std::snprintf(buf, sizeof buf, "%.*f", decimals, v);
    // rounding may have produced one more integer digit (e.g. 9.99..->10.0): re-derive
    std::string s(buf);
    size_t digits = 0;
    for (char c : s)
      if (c >= '0' && c <= '9') ++digits;
    if (a >= 1.0 && digits > 17) s.erase(s.size() - (digits - 17));
    std::string out(20 > s.size() ? 20 - s.size() : 0, ' ');
    return out + s + "     ";
  }
  std::snprintf(buf, sizeof buf, "%.16E", v);
  // C gives d.ddddE+xx ; gfortran gives d.ddddE+xxx with a 3-digit exponent in this format
  std::string s(buf);
  size_t e = s.find('E');
  std::string mant = s.substr(0, e), ex = s.substr(e + 2);
  char sign = s[e + 1];
  while (ex.size() < 3) ex = "0" + ex;
  s = mant + "E" + sign + ex;
  std::string out(25 > s.size() ? 25 - s.size() : 0, ' ');
  return out + s;
}

std::string wordwrap(const std::string& text_in, int linelength) {
  std::string text = text_in;
  while (!text.empty() && text.back() == ' ') text.pop_back();
  if (text.empty()) return "\n";
  if (linelength < 2) linelength = 2;
  const size_t len = text.size();
  std::string out;
  size_t start = 0;
  bool first = true;
  for (;;) {
    size_t stop = start + (size_t)linelength - 1;  // inclusive, 0-based
    if (stop <= len - 1) {
      std::string seg = text.substr(start, (size_t)linelength);
      while (!seg.empty() && seg.back() == ' ') seg.pop_back();
      const size_t lastpos = seg.rfind(' ');
      if (lastpos != std::string::npos) stop = start + lastpos;
    } else {
      stop = len - 1;
    }
    std::string line = text.substr(start, stop - start + 1);
    while (!line.empty() && line.back() == ' ') line.pop_back();
    if (!first) {  // TRIM(ADJUSTL(temp)) for continuation lines
      const size_t p = line.find_first_not_of(' ');
      line = p == std::string::npos ? "" : line.substr(p);
    }
    out += " " + line + "\n";  // list-directed WRITE: leading blank
    first = false;
    if (stop == len - 1) break;
    start = stop + 1;
  }
  return out;
}

static int no_digits_int(int i) {  // utils.f90:221-236
  int j = i, k = 1;
  for (;;) {
    j = j / 10;
    if (j == 0) break;
    ++k;
  }
  return k;
}

std::string write_mean(double av, double std_err_in_mean, int err_prec) {
  if (std_err_in_mean <= 0.0) return "ERROR: NON-POSITIVE ERROR BAR!!!";
  if (err_prec < 1) return "ERROR: NON-POSITIVE PRECISION!!!";
  int lowest = (int)std::floor(std::log(std_err_in_mean) / std::log(10.0)) + 1 - err_prec;
  int err_quote = (int)std::lround(std_err_in_mean * std::pow(10.0, (double)(-lowest)));
  const int tenp = (int)std::lround(std::pow(10.0, err_prec));
  if (err_quote == tenp) {
    lowest += 1;
    err_quote /= 10;
  }
  if (err_quote >= tenp || err_quote < tenp / 10) return "ERROR: BUG IN WRITE_MEAN!!!";
  double av_quote = std::round(av * std::pow(10.0, (double)(-lowest))) * std::pow(10.0, (double)lowest);
  std::string sgn;
  if (av_quote < 0.0) {
    sgn = "-";
    av_quote = -av_quote;
  }
  if (std::trunc(av_quote) > (double)INT_MAX) return "ERROR: NUMBERS ARE TOO LARGE IN WRITE_MEAN!";
  const int int_part = (int)std::floor(av_quote);
  char buf[96];
  if (lowest < 0) {
    const double frac = std::round((av_quote - (double)int_part) * std::pow(10.0, (double)(-lowest)));
    if (frac > (double)INT_MAX) return "ERROR: NUMBERS ARE TOO LARGE IN WRITE_MEAN!";
    const int dec_part = (int)std::lround(frac);
    if (dec_part < 0) return "ERROR: BUG IN WRITE_MEAN! (2)";
    std::string zero_pad;
    for (int i = 1; i <= -lowest - no_digits_int(dec_part); ++i) zero_pad += '0';
    std::snprintf(buf, sizeof buf, "%s%d.%s%d(%d)", sgn.c_str(), int_part, zero_pad.c_str(), dec_part, err_quote);
  } else {
    std::snprintf(buf, sizeof buf, "%s%d(%d)", sgn.c_str(), int_part,
                  err_quote * (int)std::lround(std::pow(10.0, (double)lowest)));
  }
  return buf;
}

static double gammln(double xx) {  // utils.f90:313-335
  static const double cof[6] = {76.18009172947146,  -86.50532032941677,    24.01409824083091,
                                -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
  const double stp = 2.5066282746310005;
  double x = xx, y = x, tmp = x + 5.5;
  tmp = (x + 0.5) * std::log(tmp) - tmp;
  double ser = 1.000000000190015;
  for (int j = 0; j < 6; ++j) {
    y += 1.0;
    ser += cof[j] / y;
  }
  return tmp + std::log(stp * ser / x);
}

static double gser(double a, double x) {  // utils.f90:253-281
  const double eps = 6e-16;
  if (x <= 0.0) return 0.0;
  double ap = a, sum = 1.0 / a, del = sum;
  for (int n = 1; n <= 10000; ++n) {
    ap += 1.0;
    del = del * x / ap;
    sum += del;
    if (std::fabs(del) < std::fabs(sum) * eps) return sum * std::exp(-x + a * std::log(x) - gammln(a));
  }
  return NAN;
}

static double gcf(double a, double x) {  // utils.f90:284-310
  const double eps = 6e-16, fpmin = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / fpmin, d = 1.0 / b, h = d;
  for (int i = 1; i <= 10000; ++i) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < fpmin) d = fpmin;
    c = b + an / c;
    if (std::fabs(c) < fpmin) c = fpmin;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < eps) return std::exp(-x + a * std::log(x) - gammln(a)) * h;
  }
  return NAN;
}

double gammq(double a, double x) {
  if (x < 0.0 || a <= 0.0) return NAN;
  if (x < a + 1.0) return 1.0 - gser(a, x);
  return gcf(a, x);
}

// A Fortran list-directed REAL item: optional sign, digits with optional point, optional
// exponent with e/E/d/D (or a bare sign), terminated by blank, comma, slash or end of line.
static bool parse_real(const char*& p, double& v) {
  while (*p == ' ' || *p == '\t') ++p;
  const char* s = p;
  std::string t;
  if (*p == '+' || *p == '-') t += *p++;
  bool digits = false;
  while ((*p >= '0' && *p <= '9') || *p == '.') {
    if (*p != '.') digits = true;
    t += *p++;
  }
  if (!digits) {
    p = s;
    return false;
  }
  if (*p == 'e' || *p == 'E' || *p == 'd' || *p == 'D' || *p == 'q' || *p == 'Q') {
    const char* q = p + 1;
    std::string ex = "e";
    if (*q == '+' || *q == '-') ex += *q++;
    bool ed = false;
    while (*q >= '0' && *q <= '9') {
      ed = true;
      ex += *q++;
    }
    if (ed) {
      t += ex;
      p = q;
    }
  } else if ((*p == '+' || *p == '-') && p[1] >= '0' && p[1] <= '9') {  // 1.5+3 means 1.5e+3
    std::string ex = "e";
    ex += *p++;
    while (*p >= '0' && *p <= '9') ex += *p++;
    t += ex;
  }
  if (!(*p == '\0' || *p == ' ' || *p == '\t' || *p == ',' || *p == '/' || *p == '\r' || *p == '\n')) {
    p = s;
    return false;
  }
  v = std::strtod(t.c_str(), nullptr);
  return true;
}

bool isdataline(const std::string& line) {
  const char* p = line.c_str();
  double v;
  return parse_real(p, v);
}

static bool read3(const std::string& line, double* out) {
  const char* p = line.c_str();
  for (int k = 0; k < 3; ++k) {
    if (!parse_real(p, out[k])) return false;
    while (*p == ' ' || *p == '\t') ++p;
    if (*p == ',') ++p;
  }
  return true;
}

std::string read_data(const std::string& path, DataSet& out) {
  std::ifstream f(path);
  if (!f) return "Problem opening file " + path + ".";
  std::vector<std::string> lines;
  std::string line;
  while (std::getline(f, line)) {
    if (line.size() > 200) line.resize(200);  // CHARACTER(200), bootstrap.f90:59
    lines.push_back(line);
  }
  out = DataSet();
  bool any = false;
  for (auto& l : lines) {
    size_t p0 = l.find_first_not_of(" \t");
    std::string adj = p0 == std::string::npos ? "" : l.substr(p0);
    if (!isdataline(adj)) continue;
    any = true;
    double t[3];
    if (!read3(adj, t)) return "Problem reading " + path + ".\n Problem line: " + adj;
    if (t[2] <= 0.0) return "Found a non-positive error bar.\n Problem line: " + adj;  // bootstrap.f90:134-138
    out.x.push_back(t[0]);
    out.y.push_back(t[1]);
    out.err.push_back(t[2]);
  }
  if (!any) return "File " + path + " doesn't contain any data.";
  // sort_data (bootstrap.f90:155-194): stable ranking by x (mrgrnk is a merge sort)
  const size_t n = out.x.size();
  std::vector<size_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return out.x[a] < out.x[b]; });
  for (size_t i = 0; i < n; ++i)
    if (idx[i] != i) out.was_out_of_order = true;
  if (out.was_out_of_order) {
    DataSet s = out;
    for (size_t i = 0; i < n; ++i) {
      s.x[i] = out.x[idx[i]];
      s.y[i] = out.y[idx[i]];
      s.err[i] = out.err[idx[i]];
    }
    s.was_out_of_order = true;
    out = s;
  }
  for (size_t i = 1; i < n; ++i)
    if (std::fabs(out.x[i] - out.x[i - 1]) < 1e-13) {
      out.has_equal_x = true;
      out.equal_x_value = out.x[i];
      break;
    }
  return "";
}

}  // namespace fmc_host

// C hooks for the CPU tests (tests/test_host_driver.py)
extern "C" {
void fmc_host_write_mean(double av, double err, int prec, char* buf72) {
  std::snprintf(buf72, 72, "%s", fmc_host::write_mean(av, err, prec).c_str());
}
double fmc_host_gammq(double a, double x) { return fmc_host::gammq(a, x); }
void fmc_host_list_directed_real(double v, char* buf32) {
  std::snprintf(buf32, 32, "%s", fmc_host::list_directed_real(v).c_str());
}
int fmc_host_isdataline(const char* s) { return fmc_host::isdataline(s) ? 1 : 0; }
int fmc_host_wordwrap(const char* text, int linelength, char* out, int cap) {
  std::string s = fmc_host::wordwrap(text, linelength);
  std::snprintf(out, cap, "%s", s.c_str());
  return (int)s.size();
}
// returns n (>=0) or -1; flags bit0 = was out of order, bit1 = equal x found
int fmc_host_read_data(const char* path, double* x, double* y, double* err, int cap, int* flags, char* msg256) {
  fmc_host::DataSet d;
  std::string e = fmc_host::read_data(path, d);
  if (!e.empty()) {
    std::snprintf(msg256, 256, "%s", e.c_str());
    return -1;
  }
  const int n = (int)d.x.size();
  for (int i = 0; i < n && i < cap; ++i) {
    x[i] = d.x[i];
    y[i] = d.y[i];
    err[i] = d.err[i];
  }
  *flags = (d.was_out_of_order ? 1 : 0) | (d.has_equal_x ? 2 : 0);
  return n;
}
}
