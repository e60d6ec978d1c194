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
//   normal n of a walker (n = step * normals_per_step + j) comes from call n >> 2, lane n & 3;
//   lanes (0,1) and (2,3) are the two Box-Muller pairs of a call.
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

// Two standard normals from two 32-bit words (Box-Muller).
// fp32: u1 in (0,1] from 24 bits feeds lg2.approx / sqrt.approx, the angle uses sin/cos.approx
//       (MUFU pipe); fp64: full-precision log / sqrt / sincospi on 32-bit uniforms.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  // u1 = n * 2^-24 with n = (a >> 8) + 1 in [1, 2^24]  (never 0, never denormal):
  //   r^2 = -2 ln u1 = -2 ln2 * (lg2 n - 24)  -> one I2F, one MUFU.LG2, one FFMA, one MUFU.SQRT
  // angle = 2 pi * (b >> 8) * 2^-24                  -> one I2F, one FMUL, MUFU.SIN + MUFU.COS
  const float n = (float)((a >> 8) + 1u);
  // |.|: at n = 2^24 the FFMA may round to -1e-7; the absolute value is a free MUFU operand modifier
  const float r = sqrt_approx(fabsf(fmaf(lg2_approx(n), -1.3862943611198906f, 33.271064666877374f)));
  const float ang = (float)(b >> 8) * 3.7450702829239655e-07f;  // 2 pi 2^-24
  z0 = r * cos_approx(ang);
  z1 = r * sin_approx(ang);
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, double& z0, double& z1) {
  const double u1 = ((double)a + 1.0) * 2.3283064365386963e-10;  // (0, 1]
  const double u2 = (double)b * 2.3283064365386963e-10;          // [0, 1)
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// Fill z[0..3] with the four normals of Philox call `call` of one walker (for f2: of the walker pair
// (walker, walker + 1), one Philox call each).
template <typename R>
__device__ __forceinline__ void normals4(const PhiloxKeys& rk, unsigned long long call,
                                         unsigned long long walker, R (&z)[4]) {
  const u32x4 r = philox4x32_10(
      u32x4{(uint32_t)call, (uint32_t)(call >> 32), (uint32_t)walker, (uint32_t)(walker >> 32)}, rk);
  box_muller(r.x, r.y, z[0], z[1]);
  box_muller(r.z, r.w, z[2], z[3]);
}

template <>
__device__ __forceinline__ void normals4<f2>(const PhiloxKeys& rk, unsigned long long call,
                                             unsigned long long walker, f2 (&z)[4]) {
  float a[4], b[4];
  normals4<float>(rk, call, walker, a);
  normals4<float>(rk, call, walker + 1ull, b);
#pragma unroll
  for (int j = 0; j < 4; j++) z[j] = f2(a[j], b[j]);
}

// the same for a pair of walkers that are not neighbours (compacted committor lanes)
template <typename R>
__device__ __forceinline__ void normals4_pair(const PhiloxKeys& rk, unsigned long long call,
                                              unsigned long long walker, unsigned long long, R (&z)[4]) {
  normals4<R>(rk, call, walker, z);
}
template <>
__device__ __forceinline__ void normals4_pair<f2>(const PhiloxKeys& rk, unsigned long long call,
                                                  unsigned long long walker_a, unsigned long long walker_b,
                                                  f2 (&z)[4]) {
  float a[4], b[4];
  normals4<float>(rk, call, walker_a, a);
  normals4<float>(rk, call, walker_b, b);
#pragma unroll
  for (int j = 0; j < 4; j++) z[j] = f2(a[j], b[j]);
}
#endif

}  // namespace nds
