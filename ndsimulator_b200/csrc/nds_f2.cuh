// nds_f2.cuh -- "two walkers per thread" real type for Blackwell packed FP32.
//
// sm_100 adds FFMA2 / FMUL2 / FADD2 (PTX fma.rn.f32x2 ...): one issue slot carries two FP32 operations.
// The fused step kernel is bound by instruction issue (ncu: 77 % issue-active, profiles/r01_c2_*), so
// instantiating the SAME templated physics with R = f2 -- every value a pair (walker 2t, walker 2t+1) --
// halves the FP32 issue slots per walker-step without a second implementation of the physics.
// MUFU and Philox work stays per walker.  Memory layout is unchanged: SoA float[d][M] read as f2[d][M/2].
#pragma once
#include <cuda_runtime.h>
#include <cmath>

namespace nds {

struct f2 {
  float2 v;
  __host__ __device__ f2() : v{0.f, 0.f} {}
  __host__ __device__ f2(float a) : v{a, a} {}
  __host__ __device__ f2(double a) : v{(float)a, (float)a} {}
  __host__ __device__ f2(int a) : v{(float)a, (float)a} {}
  __host__ __device__ f2(float a, float b) : v{a, b} {}
  __host__ __device__ explicit f2(float2 a) : v(a) {}
};

template <typename R> struct pack_of { static constexpr int value = 1; };
template <> struct pack_of<f2> { static constexpr int value = 2; };

#ifdef __CUDACC__
using ::fmaf;  // keep the scalar overload visible next to nds::fmaf(f2, f2, f2)
#define NDS_F2 __device__ __forceinline__
NDS_F2 f2 operator+(f2 a, f2 b) { return f2(__fadd2_rn(a.v, b.v)); }
NDS_F2 f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }
NDS_F2 f2 operator-(f2 a, f2 b) { return f2(__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))); }
NDS_F2 f2 operator*(f2 a, f2 b) { return f2(__fmul2_rn(a.v, b.v)); }
NDS_F2 f2 operator/(f2 a, f2 b) { return f2(__fdividef(a.v.x, b.v.x), __fdividef(a.v.y, b.v.y)); }
NDS_F2 f2& operator+=(f2& a, f2 b) { a = a + b; return a; }
NDS_F2 f2& operator-=(f2& a, f2 b) { a = a - b; return a; }
NDS_F2 f2& operator*=(f2& a, f2 b) { a = a * b; return a; }
NDS_F2 f2 fmaf(f2 a, f2 b, f2 c) { return f2(__ffma2_rn(a.v, b.v, c.v)); }
// comparisons exist only so that code paths that are compiled out for packed kernels still parse
NDS_F2 bool operator<(f2 a, f2 b) { return a.v.x < b.v.x; }
NDS_F2 bool operator>(f2 a, f2 b) { return a.v.x > b.v.x; }
#endif

}  // namespace nds
