// Counter-based Gaussian generator: replaces the reference's single serial RANLUX stream
// (rang under !$OMP CRITICAL, bootstrap.f90:205-207; utils.f90:345-364; ranlux.f) so that
// every (resample, data point) pair has its own reproducible deviate, independent of
// which thread, GPU or launch computes it.
//
//   bits    = Philox4x32-10( key = (seed_lo, seed_hi), counter = (point/4, 0, resample_lo, resample_hi) )
//   z[4k..4k+3] = two Box-Muller pairs from the four 32-bit words of block k
//
// The deviates are evaluated in FP32 (by default with the SFU's lg2/rsq/sin/cos): they carry
// about the resolution of the reference's, which are built from 24-bit RANLUX uniforms, and
// the INT32/FP32/SFU work runs in issue slots the half-rate FP64 pipe leaves free, so
// regenerating z in every pass over the data costs no FP64 throughput and no memory.
#pragma once
#include <cstring>

#include "fmc_common.cuh"

#ifndef FMC_PHILOX_ROUNDS
#define FMC_PHILOX_ROUNDS 10
#endif

namespace fmc {

struct U4 {
  uint32_t x, y, z, w;
};

FMC_HD FMC_INLINE void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  hi = __umulhi(a, b);
  lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

// Philox4x32-10 (Salmon et al., SC'11).  Known answers (Random123 kat_vectors) are checked on the host
// image in tests/test_host_emul.py and on the device, bit for bit, in tests/test_gpu_rng.py.
FMC_HD FMC_INLINE U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < FMC_PHILOX_ROUNDS; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c.x, hi0, lo0);
    mulhilo32(0xCD9E8D57u, c.z, hi1, lo1);
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Box-Muller on two 32-bit words: u in (0,1], angle = 2*pi*v with v in [0,1).
// Branch-free: u is always a normal positive float (>= 2^-33), so ln(u) needs no special
// cases; ln(u) = e*ln2 + ln(m), m in [sqrt(1/2), sqrt(2)), ln(m) by the atanh series in
// s = (m-1)/(m+1) (|s| < 0.172, truncation error < 2e-9 relative).
FMC_HD FMC_INLINE float ln_pos_normal(float u) {
#if defined(__CUDA_ARCH__)
  int bits = __float_as_int(u);
#else
  int bits;
  memcpy(&bits, &u, 4);
#endif
  int e = ((bits - 0x3f3504f3) >> 23);           // exponent such that m = u * 2^-e is in [0.7071, 1.4142)
  int mb = bits - (e << 23);
#if defined(__CUDA_ARCH__)
  float m = __int_as_float(mb);
  const float s = __fdividef(m - 1.0f, m + 1.0f);
#else
  float m;
  memcpy(&m, &mb, 4);
  const float s = (m - 1.0f) / (m + 1.0f);
#endif
  const float s2 = s * s;
  float p = fmaf(s2, 0.2857142857f, 0.4f);        // 2/7, 2/5
  p = fmaf(p, s2, 0.6666666667f);                 // 2/3
  p = fmaf(p, s2, 2.0f);
  return fmaf((float)e, 0.6931471806f, p * s);
}

FMC_HD FMC_INLINE void normal_pair(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // 2^-32
#if defined(__CUDA_ARCH__) && !defined(FMC_ACCURATE_NORMALS)
  // SFU path (MUFU lg2/rsq/sin/cos, ~1e-6 relative): +8 % fits/s on C2 versus the polynomial
  // path below, which -DFMC_ACCURATE_NORMALS selects.  Either way the deviates are what
  // fmc_generate_offsets reports, bit for bit.
  const float tf = -1.3862943611f * __log2f(u);
  const float rf = (tf > 0.0f) ? tf * rsqrtf(tf) : 0.0f;
  float sf, cf;
  __sincosf(6.2831853072f * ((float)b * 2.3283064365386963e-10f), &sf, &cf);
  z0 = rf * cf;
  z1 = rf * sf;
  return;
#endif
  const float t = -2.0f * ln_pos_normal(u);                       // >= 0 (u <= 1)
#if defined(__CUDA_ARCH__)
  const float r = (t > 0.0f) ? t * rsqrtf(t) : 0.0f;
#else
  const float r = (t > 0.0f) ? sqrtf(t) : 0.0f;
#endif
  const float v2 = (float)b * 4.6566128730773926e-10f;            // 2 * b * 2^-32
  float s, c;
#if defined(__CUDA_ARCH__)
  sincospif(v2, &s, &c);
#else
  s = sinf(3.14159265358979323846f * v2);
  c = cosf(3.14159265358979323846f * v2);
#endif
  z0 = r * c;
  z1 = r * s;
}

// The four deviates of points 4*blk .. 4*blk+3 of resample `id`.
FMC_HD FMC_INLINE void normals4(uint64_t seed, uint64_t id, uint32_t blk, float (&z)[4]) {
  U4 c;
  c.x = blk;
  c.y = 0u;
  c.z = (uint32_t)id;
  c.w = (uint32_t)(id >> 32);
  U4 o = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  normal_pair(o.x, o.y, z[0], z[1]);
  normal_pair(o.z, o.w, z[2], z[3]);
}

}  // namespace fmc
