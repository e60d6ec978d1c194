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

#ifdef __CUDACC__
// Two standard normals from two 32-bit words (Box-Muller).
// fp32: u1 in (0,1] from 24 bits feeds lg2.approx / sqrt.approx, the angle uses sin/cos.approx
//       (MUFU pipe); fp64: full-precision log / sqrt / sincospi on 32-bit uniforms.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  // u1 = (a >> 8 + 1) * 2^-24 in (0, 1]; built with integer ops + one FFMA (no I2F on the XU pipe for
  // the mantissa trick would need |u| in [1,2); the conversion below is a single I2F.
  const float u1 = (float)((a >> 8) + 1u) * 5.9604644775390625e-08f;   // 2^-24
  const float u2 = (float)(b >> 8) * 5.9604644775390625e-08f;          // [0, 1)
  // r = sqrt(-2 ln u1) = sqrt(-2 ln2 * log2 u1)
  const float r = sqrtf(-1.3862943611198906f * __log2f(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
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

// Fill z[0..3] with the four normals of Philox call `call` of one walker.
template <typename R>
__device__ __forceinline__ void normals4(uint32_t k0, uint32_t k1, unsigned long long call,
                                         unsigned long long walker, R (&z)[4]) {
  const u32x4 r = philox4x32_10(
      u32x4{(uint32_t)call, (uint32_t)(call >> 32), (uint32_t)walker, (uint32_t)(walker >> 32)}, k0, k1);
  box_muller(r.x, r.y, z[0], z[1]);
  box_muller(r.z, r.w, z[2], z[3]);
}
#endif

}  // namespace nds
