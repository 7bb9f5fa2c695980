// TEST INFRASTRUCTURE ONLY -- parity oracle, not product code.
//
// Restatement of the reference's serial Gaussian stream:
//   RANLUX / RLUXGO   ranlux.f:116-152, 222-312  (subtract-with-borrow, luxury skipping)
//   rang              utils.f90:345-364          (Marsaglia polar method on pairs of uniforms)
//   initialise_rng    utils.f90:338-342, 6-7     (luxury level 3, seed 31415)
//
// PINNED by the reference's own known-answer vectors (ranlux.f:399-473), see
// tests/test_oracle_ranlux.py.
#include "fmc_oracle.h"

#include <cmath>
#include <cstdint>

namespace {

struct Ranlux {
  double seeds[25];  // 1-based
  int next[25];
  int i24 = 24, j24 = 10, in24 = 0, nskip = 199, luxlev = 3;
  long long kount = 0;
  double carry = 0.0, twom24 = 0.0, twom12 = 0.0;
  bool notyet = true;
};

Ranlux g_rng;

const int kNdskip[5] = {0, 24, 73, 199, 365};  // ranlux.f:72
const int kItwo24 = 1 << 24, kIcons = 2147483563, kJsdflt = 314159265;

inline void swb_step(Ranlux& s, double* uni_out) {  // ranlux.f:116-126 (== 139-149)
  double uni = s.seeds[s.j24] - s.seeds[s.i24] - s.carry;
  if (uni < 0.0) {
    uni = uni + 1.0;
    s.carry = s.twom24;
  } else {
    s.carry = 0.0;
  }
  s.seeds[s.i24] = uni;
  s.i24 = s.next[s.i24];
  s.j24 = s.next[s.j24];
  *uni_out = uni;
}

void rluxgo(Ranlux& s, int lux, int ins, int k1, int k2) {  // ranlux.f:222-312
  if (lux < 0) {
    s.luxlev = 3;
  } else if (lux <= 4) {
    s.luxlev = lux;
  } else if (lux < 24 || lux > 2000) {
    s.luxlev = 4;
  } else {
    s.luxlev = lux;
    for (int ilx = 0; ilx <= 4; ++ilx)
      if (lux == kNdskip[ilx] + 24) s.luxlev = ilx;
  }
  s.nskip = (s.luxlev <= 4) ? kNdskip[s.luxlev] : s.luxlev - 24;
  s.in24 = 0;
  int jseed = ins > 0 ? ins : kJsdflt;
  s.notyet = false;
  s.twom24 = 1.0;
  int iseeds[25];
  for (int i = 1; i <= 24; ++i) {
    s.twom24 = s.twom24 * 0.5;
    int k = jseed / 53668;
    jseed = 40014 * (jseed - k * 53668) - k * 12211;
    if (jseed < 0) jseed = jseed + kIcons;
    iseeds[i] = jseed % kItwo24;
  }
  s.twom12 = s.twom24 * 4096.0;
  for (int i = 1; i <= 24; ++i) {
    s.seeds[i] = (double)iseeds[i] * s.twom24;
    s.next[i] = i - 1;
  }
  s.next[1] = 24;
  s.i24 = 24;
  s.j24 = 10;
  s.carry = 0.0;
  if (s.seeds[24] == 0.0) s.carry = s.twom24;
  s.kount = k1;
  if (k1 + k2 != 0) {
    // restart by skipping k1 + 1e9*k2 numbers (ranlux.f:283-310)
    for (int outer = 1; outer <= k2 + 1; ++outer) {
      long long inner = (outer == k2 + 1) ? k1 : 1000000000LL;
      double u;
      for (long long isk = 1; isk <= inner; ++isk) swb_step(s, &u);
    }
    s.in24 = (int)(s.kount % (s.nskip + 24));
    if (k2 > 0) {
      int izip = 1000000000 % (s.nskip + 24);
      int izip2 = k2 * izip + s.in24;
      s.in24 = izip2 % (s.nskip + 24);
    }
    if (s.in24 > 23) s.in24 = 0;
  }
}

void ranlux(Ranlux& s, double* rvec, int lenv) {  // ranlux.f:76-158
  if (s.notyet) rluxgo(s, 3, kJsdflt, 0, 0);  // default initialisation, ranlux.f:76-113
  for (int ivec = 0; ivec < lenv; ++ivec) {
    double uni;
    swb_step(s, &uni);
    rvec[ivec] = uni;
    if (uni < s.twom12) {  // "padding" of small numbers, ranlux.f:129-133
      rvec[ivec] = rvec[ivec] + s.twom24 * s.seeds[s.j24];
      if (rvec[ivec] == 0.0) rvec[ivec] = s.twom24 * s.twom24;
    }
    s.in24 = s.in24 + 1;
    if (s.in24 == 24) {  // luxury skipping, ranlux.f:135-151
      s.in24 = 0;
      s.kount = s.kount + s.nskip;
      double u;
      for (int isk = 1; isk <= s.nskip; ++isk) swb_step(s, &u);
    }
  }
  s.kount = s.kount + lenv;
}

}  // namespace

extern "C" {

void fmc_oracle_rluxgo(int lux, int seed, int k1, int k2) { rluxgo(g_rng, lux, seed, k1, k2); }

void fmc_oracle_rlux_reset_default(void) { g_rng = Ranlux(); }

void fmc_oracle_ranlux(double* rvec, int len) { ranlux(g_rng, rvec, len); }

long long fmc_oracle_rluxat_kount(void) { return g_rng.kount; }

// utils.f90:338-342
void fmc_oracle_initialise_rng(void) { rluxgo(g_rng, 3, 31415, 0, 0); }

// utils.f90:345-364 -- polar method; for odd n the second deviate of the last pair is dropped.
void fmc_oracle_rang(int n, double* r) {
  for (int i = 1; i <= n; i += 2) {
    double ranx[2], v1, v2, rad;
    for (;;) {
      ranlux(g_rng, ranx, 2);
      v1 = 2.0 * ranx[0] - 1.0;
      v2 = 2.0 * ranx[1] - 1.0;
      rad = v1 * v1 + v2 * v2;
      if (rad <= 1.0 && rad > 0.0) break;
    }
    double f = std::sqrt(-2.0 * std::log(rad) / rad);
    r[i - 1] = v1 * f;
    if (i < n) r[i] = v2 * f;
  }
}

}  // extern "C"
