// Pipe-throughput microbenchmark for the step kernels' instruction mix (sm_100a).
// Each kernel runs ITER iterations of an unrolled body with 8 independent accumulator chains per thread;
// reports warp-instructions per clock per SM sub-partition (SMSP).  Build: nvcc -arch=sm_100a -O3 pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITER 4096
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, float c0, float c1, unsigned seed, long long* cyc) {
  float a[NACC]; float2 p[NACC]; unsigned u[NACC]; float m[NACC];
  for (int i = 0; i < NACC; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = make_float2(a[i], a[i] + 1.f); u[i] = seed + threadIdx.x * 7 + i; m[i] = a[i] * 0.01f; }
  const float2 cc0 = make_float2(c0, c0), cc1 = make_float2(c1, c1);
  long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (MODE == 0) a[i] = fmaf(a[i], c0, c1);                                   // FFMA r, UR, UR
      if (MODE == 1) p[i] = __ffma2_rn(p[i], cc0, cc1);                           // FFMA2 r, UR, UR
      if (MODE == 2) { p[i] = __ffma2_rn(p[i], cc0, cc1); a[i] = fmaf(a[i], c0, c1); }        // 1 packed : 1 scalar
      if (MODE == 3) { p[i] = __ffma2_rn(p[i], cc0, cc1); a[i] = fmaf(a[i], c0, c1); m[i] = fmaf(m[i], c0, c1); }  // 1:2
      if (MODE == 4) { unsigned long long w = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(w >> 32) ^ (unsigned)w; }  // IMAD.WIDE + LOP3
      if (MODE == 5) { p[i] = __ffma2_rn(p[i], cc0, cc1); unsigned long long w = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(w >> 32) ^ (unsigned)w; }
      if (MODE == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i])); }  // MUFU
      if (MODE == 7) { p[i] = __ffma2_rn(p[i], cc0, cc1); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i])); }
      if (MODE == 8) { a[i] = fmaf(a[i], m[i], c1); }                              // FFMA r, r, UR  (2 reg sources)
      if (MODE == 9) { p[i] = __ffma2_rn(p[i], p[(i + 1) % NACC], cc1); }          // FFMA2 r, r, UR
      if (MODE == 10) { a[i] = fmaf(a[i], c0, c1); unsigned long long w = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(w >> 32) ^ (unsigned)w; }
      if (MODE == 11) { p[i] = __fmul2_rn(p[i], cc0); }
      if (MODE == 13) { u[i] ^= u[(i + 1) % NACC]; }
      if (MODE == 14) { unsigned lo, hi; asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(u[i])); u[i] = hi; }
      if (MODE == 15) { u[i] = __umulhi(u[i], 0xD2511F53u); }
      if (MODE == 16) { u[i] = u[i] * 0xD2511F53u + 1u; }
      if (MODE == 17) { a[i] = fmaf(a[i], c0, c1); u[i] = u[i] * 0xD2511F53u + 1u; }
      if (MODE == 18) { a[i] = fmaf(a[i], c0, c1); u[i] ^= u[(i + 1) % NACC]; }
      if (MODE == 12) { p[i] = __ffma2_rn(p[i], cc0, cc1); p[i] = __ffma2_rn(p[i], cc1, cc0); a[i] = fmaf(a[i], c0, c1); }  // 2 packed : 1 scalar
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < NACC; i++) s += a[i] + p[i].x + p[i].y + (float)u[i] + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_body, int warps_per_smsp) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 16 * 128 * 4 * 4); cudaMalloc(&cyc, 8);
  const int blocks_per_sm = warps_per_smsp;  // 128 threads = 4 warps = one per SMSP
  k<MODE><<<148 * blocks_per_sm, 128>>>(out, 1.0001f, 0.5f, 12345u, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * blocks_per_sm, 128>>>(out, 1.0001f, 0.5f, 12345u, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double winstr = (double)ITER * NACC * instr_per_body * warps_per_smsp;  // per SMSP
  (void)c;
  printf("%-30s warps/SMSP=%2d  %.3f ms  clk per body per warp-slot = %.2f  (warp-instr/clk/SMSP %.3f)\n", name, warps_per_smsp, ms,
         (ms - 0.004) * 1e-3 * 1.965e9 / ((double)ITER * NACC * warps_per_smsp), winstr / ((ms - 0.004) * 1e-3 * 1.965e9));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {12}) {
    run<0>("FFMA r,UR,UR", 1, w);
    run<8>("FFMA r,r,UR", 1, w);
    run<1>("FFMA2 r,UR,UR", 1, w);
    run<9>("FFMA2 r,r,UR", 1, w);
    run<11>("FMUL2 r,UR", 1, w);
    run<2>("FFMA2 + FFMA (1:1)", 2, w);
    run<3>("FFMA2 + 2 FFMA (1:2)", 3, w);
    run<12>("2 FFMA2 + FFMA (2:1)", 3, w);
    run<4>("IMAD.WIDE + LOP3", 2, w);
    run<5>("FFMA2 + IMAD.WIDE + LOP3", 3, w);
    run<10>("FFMA + IMAD.WIDE + LOP3", 3, w);
    run<13>("LOP3", 1, w);
    run<14>("IMAD.WIDE", 1, w);
    run<15>("IMAD.HI", 1, w);
    run<16>("IMAD lo", 1, w);
    run<17>("FFMA + IMAD lo", 2, w);
    run<18>("FFMA + LOP3", 2, w);
    run<6>("MUFU.EX2", 1, w);
    run<7>("FFMA2 + MUFU.EX2", 2, w);
  }
  return 0;
}
