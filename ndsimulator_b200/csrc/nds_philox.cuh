// nds_philox.cuh -- Philox4x32-10 counter-based generator and Box-Muller normals.
//
// Replaces the reference's numpy.random.RandomState draws (engines/md.py:120,127-128 per step,
// engines/modify.py:45-47 for the initial velocities).  The Gaussian STREAM differs from MT19937 by
// construction (statistical parity only, BASELINE.json north_star); the generator itself is pinned by
// the Random123 known-answer vectors in tests/test_philox.py.
//
// Stream layout (results do not depend on the number of ranks or on launch boundaries):
//   key     = (seed_lo, seed_hi ^ domain)          domain 0: per-step noise, 1: initial velocities
//   counter = (call_lo, call_hi, walker_lo, walker_hi)   walker = GLOBAL walker id
//   one call yields SIX normals: its 128 bits are cut into six 21-bit fields (three Box-Muller pairs of
//   a radius variate and an angle).  32x32 -> 64-bit multiplies run at a quarter of the FP32 rate on sm_100
//   (profiles/microbench), so Philox is the most expensive part of a 2-D step.
//   fp32: u1 = (a + 1/2) 2^-21 (cell midpoint: the tail mass of every cell is right to first order), radius
//         cut-off sqrt(2 ln 2^22) = 5.52 sigma; fp32 keeps 24 bits at most anyway.
//   fp64: the parity-precision engine spends a SECOND Philox call (same counter, top bit of the call word set) to
//         extend every field: 53-bit radius variates u1 = (a 2^32 + e + 1/2) 2^-53 (cut-off 8.57 sigma) and
//         32 / 32 / 31-bit angles.
//   Steps are grouped so that a group of G steps uses whole calls (noise_group() in nds_kernels.cuh).
#pragma once
#include <cstdint>
#include "nds_f2.cuh"

namespace nds {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
    const uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
#else
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c.x, p1 = (uint64_t)PHILOX_M1 * c.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    c = u32x4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += PHILOX_W0;
    k1 += PHILOX_W1;
  }
  return c;
}

// Same function with the ten round keys precomputed (they are uniform over the whole run, so the
// kernels read them from the constant bank instead of bumping them per thread).
struct PhiloxKeys { uint32_t k0[10], k1[10]; };

__host__ __device__ inline PhiloxKeys philox_round_keys(uint32_t k0, uint32_t k1) {
  PhiloxKeys r;
  for (int i = 0; i < 10; i++) { r.k0[i] = k0 + (uint32_t)i * PHILOX_W0; r.k1[i] = k1 + (uint32_t)i * PHILOX_W1; }
  return r;
}

#ifdef __CUDACC__
__device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, const PhiloxKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
    const uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
    c = u32x4{hi1 ^ c.y ^ rk.k0[r], lo1, hi0 ^ c.w ^ rk.k1[r], lo0};
  }
  return c;
}

__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_approx(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_approx(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int NDS_NPC = 6;  // normals per Philox call

// Two standard normals from two 21-bit fields (Box-Muller).
// fp32: u1 in (0,1] feeds lg2.approx / sqrt.approx, the angle uses sin/cos.approx (MUFU pipe);
// fp64: full-precision log / sqrt / sincospi on the same 21-bit uniforms.
// Radius constants: r^2 = ra * lg2(2a + 1) + rb with (ra, rb) = (-2 ln2, 44 ln2) * scale^2; scale = 1 gives
// standard normals, scale = c2 gives the noise term c2 * xi of the Langevin kick without a multiply.
struct BmScale { float ra, rb; };
__host__ __device__ inline BmScale bm_scale(double scale) {
  return BmScale{(float)(-1.3862943611198906 * scale * scale), (float)(30.498475944637594 * scale * scale)};
}

__device__ __forceinline__ void box_muller21(uint32_t a, uint32_t b, float& z0, float& z1,
                                             BmScale k = BmScale{-1.3862943611198906f, 30.498475944637594f}) {
  // u1 = (a + 1/2) 2^-21 = (2a + 1) 2^-22 (never 0; midpoint of the cell):
  //   r^2 = -2 ln u1 = -2 ln2 * (lg2 (2a + 1) - 22)  -> one integer op, one I2F, one MUFU.LG2, one FFMA, one MUFU.SQRT
  // angle = 2 pi * b * 2^-21                          -> one I2F, one FMUL, MUFU.SIN + MUFU.COS
  const float n = (float)((a << 1) | 1u);  // exact: 22 significant bits
  // |.|: near n = 2^22 the FFMA may round to -1e-7; the absolute value is a free MUFU operand modifier
  const float r = sqrt_approx(fabsf(fmaf(lg2_approx(n), k.ra, k.rb)));
  const float ang = (float)b * 2.9960562263391724e-06f;  // 2 pi 2^-21
  z0 = r * cos_approx(ang);
  z1 = r * sin_approx(ang);
}

// fp64: 21-bit fields a, b extended by ea (32 bits) and eb (eb_bits bits) from the second Philox call
__device__ __forceinline__ void box_muller53(uint32_t a, uint32_t ea, uint32_t b, uint32_t eb, int eb_bits, double& z0, double& z1) {
  const double u1 = ((double)(((unsigned long long)a << 32) | ea) + 0.5) * 1.1102230246251565e-16;  // (0, 1), 2^-53
  const double u2 = (double)(((unsigned long long)b << eb_bits) | eb) * (eb_bits == 11 ? 2.3283064365386963e-10   // 2^-32
                                                                                      : 4.6566128730773926e-10);  // 2^-31
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// Fill z[0..5] with the six normals of Philox call `call` of one walker.  Fields: pair 0 = top 21 bits of
// words 0, 1; pair 1 = top 21 bits of words 2, 3; pair 2 = the leftovers (low 11 bits of word 0 | low 10
// bits of word 1, low 11 bits of word 2 | low 10 bits of word 3).
template <typename R>
__device__ __forceinline__ void normals6(const PhiloxKeys& rk, unsigned long long call,
                                         unsigned long long walker, R (&z)[NDS_NPC],
                                         BmScale k = BmScale{-1.3862943611198906f, 30.498475944637594f}) {
  const u32x4 r = philox4x32_10(
      u32x4{(uint32_t)call, (uint32_t)(call >> 32), (uint32_t)walker, (uint32_t)(walker >> 32)}, rk);
  box_muller21(r.x >> 11, r.y >> 11, z[0], z[1], k);
  box_muller21(r.z >> 11, r.w >> 11, z[2], z[3], k);
  box_muller21((r.x & 0x7FFu) | ((r.y & 0x3FFu) << 11), (r.z & 0x7FFu) | ((r.w & 0x3FFu) << 11), z[4], z[5], k);
}

template <>
__device__ __forceinline__ void normals6<double>(const PhiloxKeys& rk, unsigned long long call,
                                                 unsigned long long walker, double (&z)[NDS_NPC], BmScale) {
  const u32x4 r = philox4x32_10(
      u32x4{(uint32_t)call, (uint32_t)(call >> 32), (uint32_t)walker, (uint32_t)(walker >> 32)}, rk);
  const u32x4 e = philox4x32_10(
      u32x4{(uint32_t)call, (uint32_t)(call >> 32) | 0x80000000u, (uint32_t)walker, (uint32_t)(walker >> 32)}, rk);
  box_muller53(r.x >> 11, e.x, r.y >> 11, e.w >> 21, 11, z[0], z[1]);
  box_muller53(r.z >> 11, e.y, r.w >> 11, (e.w >> 10) & 0x7FFu, 11, z[2], z[3]);
  box_muller53((r.x & 0x7FFu) | ((r.y & 0x3FFu) << 11), e.z, (r.z & 0x7FFu) | ((r.w & 0x3FFu) << 11), e.w & 0x3FFu, 10, z[4], z[5]);
}

// f2: the walker pair (walker_a, walker_b), one Philox call each (neighbours, or any two compacted lanes)
template <typename R>
__device__ __forceinline__ void normals6_pair(const PhiloxKeys& rk, unsigned long long call,
                                              unsigned long long walker, unsigned long long, R (&z)[NDS_NPC],
                                              BmScale k = BmScale{-1.3862943611198906f, 30.498475944637594f}) {
  normals6<R>(rk, call, walker, z, k);
}
template <>
__device__ __forceinline__ void normals6_pair<f2>(const PhiloxKeys& rk, unsigned long long call,
                                                  unsigned long long walker_a, unsigned long long walker_b,
                                                  f2 (&z)[NDS_NPC], BmScale k) {
  float a[NDS_NPC], b[NDS_NPC];
  normals6<float>(rk, call, walker_a, a, k);
  normals6<float>(rk, call, walker_b, b, k);
#pragma unroll
  for (int j = 0; j < NDS_NPC; j++) z[j] = f2(a[j], b[j]);
}
#endif

}  // namespace nds
