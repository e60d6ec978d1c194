// nds_mtd_kernel.cuh -- shared-hill metadynamics step kernel (the C5 hot path).
//
// mtd_block<Spec, W, NW>: the walker x hill interaction of bias/gaus.py:236-261 is the whole cost of
// multiple-walker metadynamics (M walkers x H hills per MD step, H = M_total * depositions), so the
// mapping is the opposite of md_block:
//   * one CTA of NW*HS warps per SM; the hill table streams through shared memory in 32 KB tiles
//     with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), NBUF tiles in flight;
//   * LANES ARE HILLS: lane l of a warp owns records l, l+32*HS, ... of the tile (one LDS.128 each);
//     HS warps share the same walkers and split every tile between them (more warps per scheduler
//     without changing the walker -> CTA map), partial sums meet in shared memory once per step;
//   * each warp carries W walkers; their colvars sit in registers of every lane (warp-shuffle broadcast),
//     so one loaded hill is applied to W walkers: 8 FP32 ops + 1 MUFU.EX2 per walker-hill pair;
//   * per step a warp-shuffle butterfly reduces (V, g0, g1) of the W walkers; lane w < W owns walker w:
//     it finishes the force (potential + chain rule), integrates, draws Philox noise;
//   * K steps per launch; the host ends a launch where the next deposition is due (gaus.py:134).
// Restrictions (else the engine uses md_block): shared table, colvardim 2, method md, no frames.
#pragma once
#include "nds_kernels.cuh"

namespace nds {

constexpr int MTD_TILE_BYTES = 32768;  // one shared-memory tile: 2048 fp32 records / 1024 fp64 records
constexpr int MTD_NBUF = 3;            // tiles in flight
template <typename R> __host__ __device__ constexpr int mtd_tile_records() { return MTD_TILE_BYTES / (4 * (int)sizeof(R)); }

template <typename R> struct HillRec;
template <> struct HillRec<float> { float4 v; NDS_DEV float h0() const { return v.x; } NDS_DEV float h1() const { return v.y; } NDS_DEV float w() const { return v.z; } NDS_DEV float lw() const { return v.w; } };
template <> struct HillRec<double> { double4 v; NDS_DEV double h0() const { return v.x; } NDS_DEV double h1() const { return v.y; } NDS_DEV double w() const { return v.z; } };

// One pass over the whole table for the W walkers of this warp.
//   c0[w], c1[w]: colvars of walker w (identical in every lane).
//   out: lane w receives (Vm, g0, g1) of walker w in CANONICAL units (gaus.py:252-258, before .dot(J)).
template <typename R, int W, int NW, int HS, bool ENERGY>
NDS_DEV void mtd_hill_pass(const Consts<R>& K, const R* __restrict__ hills, long long n_hills,
                           long long n_tiles, unsigned char* smem, uint64_t* full, R* red, uint32_t& tile_counter,
                           const R (&c0)[W], const R (&c1)[W], R& Vm, R& g0, R& g1) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wi = warp % NW, half = warp / NW;
  constexpr uint32_t TILE_BYTES = MTD_TILE_BYTES;
  constexpr int MTD_TILE = mtd_tile_records<R>();
  R aV[W], a0[W], a1[W];
#pragma unroll
  for (int w = 0; w < W; w++) { aV[w] = (R)0; a0[w] = (R)0; a1[w] = (R)0; }
  // fp32: prescale so that the exponent is -(d0^2 + d1^2) in log2 units (one FMUL per coordinate per
  // hill-lane, amortised over W walkers); fp64: the reference's expression order.
  R p0 = (R)0, p1 = (R)0, s0[W], s1[W];
  if constexpr (is_f32<R>::value) {
    p0 = K.mtd_pre[0]; p1 = K.mtd_pre[1];
#pragma unroll
    for (int w = 0; w < W; w++) { s0[w] = c0[w] * p0; s1[w] = c1[w] * p1; }
  }
  // prefetch the first tiles of this pass (thread 0 of the CTA is the TMA producer)
  if (threadIdx.x == 0) {
    const long long pre = n_tiles < MTD_NBUF ? n_tiles : MTD_NBUF;
    for (long long t = 0; t < pre; t++) {
      const uint32_t b = (tile_counter + (uint32_t)t) % MTD_NBUF;
      mbar_expect_tx(&full[b], TILE_BYTES);
      tma_load_1d(smem + (size_t)b * TILE_BYTES, (const unsigned char*)hills + (size_t)t * TILE_BYTES, TILE_BYTES, &full[b]);
    }
  }
  for (long long t = 0; t < n_tiles; t++) {
    const uint32_t seq = tile_counter + (uint32_t)t;
    const uint32_t b = seq % MTD_NBUF;
    mbar_wait(&full[b], (seq / MTD_NBUF) & 1u);
    const HillRec<R>* tile = reinterpret_cast<const HillRec<R>*>(smem + (size_t)b * TILE_BYTES);
    // records at or beyond n_hills are never read: with the peer-memory exchange another rank may
    // already be storing the NEXT stride's records there while this launch is still running
    const long long left = n_hills - t * MTD_TILE;
    const int valid = left < MTD_TILE ? (int)left : MTD_TILE;
#pragma unroll 2
    for (int j = half * 32 + lane; j < valid; j += 32 * HS) {
      const HillRec<R> h = tile[j];
      if constexpr (is_f32<R>::value) {
        // Blackwell packed FP32 (FFMA2 / FADD2, sm_100 fma.rn.f32x2): two walkers per instruction,
        // an odd last walker in scalar form.  The height enters through its log2 (4th word of the
        // record, written at deposition):  w 2^-(d0^2+d1^2) = 2^-(d0^2 + d1^2 - log2 w),
        // so a walker-hill pair costs 6 FP32 lane-ops (3 packed issue slots) + 1 MUFU.EX2 and the loop
        // is bound by the MUFU pipe, not by instruction issue.
        const R hx = h.h0() * p0, hy = h.h1() * p1, nl = -h.lw();
        const float2 nhx = make_float2(-hx, -hx), nhy = make_float2(-hy, -hy), nlw = make_float2(nl, nl);
#pragma unroll
        for (int w = 0; w + 1 < W; w += 2) {
          const float2 d0 = __fadd2_rn(make_float2(s0[w], s0[w + 1]), nhx);
          const float2 d1 = __fadd2_rn(make_float2(s1[w], s1[w + 1]), nhy);
          const float2 e = __ffma2_rn(d1, d1, __ffma2_rn(d0, d0, nlw));
          const float2 tt = make_float2(ex2_approx(-e.x), ex2_approx(-e.y));
          if (ENERGY) { const float2 av = __fadd2_rn(make_float2(aV[w], aV[w + 1]), tt); aV[w] = av.x; aV[w + 1] = av.y; }
          const float2 n0 = __ffma2_rn(tt, d0, make_float2(a0[w], a0[w + 1]));
          const float2 n1 = __ffma2_rn(tt, d1, make_float2(a1[w], a1[w + 1]));
          a0[w] = n0.x; a0[w + 1] = n0.y; a1[w] = n1.x; a1[w + 1] = n1.y;
        }
        if constexpr (W % 2 == 1) {
          constexpr int w = W - 1;
          const R d0 = s0[w] - hx, d1 = s1[w] - hy;
          const R e = fmaf(d1, d1, fmaf(d0, d0, nl));
          const R tt = ex2_approx(-e);
          if (ENERGY) aV[w] += tt;
          a0[w] = fmaf(tt, d0, a0[w]);
          a1[w] = fmaf(tt, d1, a1[w]);
        }
      } else {
#pragma unroll
        for (int w = 0; w < W; w++) {
          const R d0 = c0[w] - h.h0(), d1 = c1[w] - h.h1();
          const R ex = (d0 * d0) / (R)2.0 * K.mtd_invsig2[0] + (d1 * d1) / (R)2.0 * K.mtd_invsig2[1];
          const R vr = h.w() * nds_exp(-ex);
          aV[w] += vr;
          a0[w] += vr * (d0 * K.mtd_invsig2[0]);
          a1[w] += vr * (d1 * K.mtd_invsig2[1]);
        }
      }
    }
    __syncthreads();  // every warp is done with buffer b
    if (threadIdx.x == 0 && t + MTD_NBUF < n_tiles) {
      mbar_expect_tx(&full[b], TILE_BYTES);
      tma_load_1d(smem + (size_t)b * TILE_BYTES, (const unsigned char*)hills + (size_t)(t + MTD_NBUF) * TILE_BYTES,
                  TILE_BYTES, &full[b]);
    }
  }
  tile_counter += (uint32_t)n_tiles;
  // warp-shuffle butterfly: afterwards every lane holds the totals; lane w keeps walker w's
  Vm = (R)0; g0 = (R)0; g1 = (R)0;
#pragma unroll
  for (int w = 0; w < W; w++) {
    R v = aV[w], x = a0[w], y = a1[w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (ENERGY) v += __shfl_xor_sync(0xffffffffu, v, o);
      x += __shfl_xor_sync(0xffffffffu, x, o);
      y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if (lane == w) { Vm = v; g0 = x; g1 = y; }
  }
  if constexpr (HS > 1) {
    // the HS warps that share these walkers add their partial sums in a fixed order, so every copy of a
    // walker sees bit-identical totals
    if (lane < W) {
      R* slot = red + ((half * NW + wi) * W + lane) * 3;
      slot[0] = Vm; slot[1] = g0; slot[2] = g1;
    }
    __syncthreads();
    if (lane < W) {
      R tv = (R)0, t0 = (R)0, t1 = (R)0;
#pragma unroll
      for (int hh = 0; hh < HS; hh++) {
        const R* slot = red + ((hh * NW + wi) * W + lane) * 3;
        tv += slot[0]; t0 += slot[1]; t1 += slot[2];
      }
      Vm = tv; g0 = t0; g1 = t1;
    }
    __syncthreads();
  }
  if constexpr (is_f32<R>::value) {
    // back to canonical units: (c - h)/sigma^2 = d' / (p sigma^2)
    g0 *= K.mtd_invsig2[0] / p0; g1 *= K.mtd_invsig2[1] / p1;
  }
}

template <class S, int W, int NW, int HS>
__global__ void __launch_bounds__(NW * HS * 32, 1)
mtd_block(const __grid_constant__ Consts<typename S::R> K, const __grid_constant__ BlockArgs<typename S::R> A) {
  using R = typename S::R;
  constexpr int ND = S::ND;
  constexpr int INTEG = S::INTEG;
  constexpr int NN = normals_per_step(INTEG, ND);
  constexpr int G = noise_group(NN);
  constexpr int CG = noise_calls(NN);
  static_assert(W <= 32, "one lane per walker of the warp");

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[MTD_NBUF];
  __shared__ R red[HS * NW * W * 3];
  constexpr int MTD_TILE = mtd_tile_records<R>();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long M = A.M;
  // the HS warps {wi, wi + NW, ...} carry the same W walkers (identical copies of their state)
  const long long w = ((long long)blockIdx.x * NW + (warp % NW)) * W + lane;  // walker of this lane
  const bool owner = lane < W && w < M;

  const int pot = S::POT >= 0 ? S::POT : K.pot;
  const int cv = S::CV >= 0 ? S::CV : K.cv;
  const int tcv = S::TCV >= 0 ? S::TCV : K.tcv;
  const int bias_mask = S::BIAS >= 0 ? S::BIAS : K.bias_mask;
  const int other_bias = bias_mask & ~BIAS_MTD;
  const bool ext_noise = S::NOISE >= 0 ? (S::NOISE == NDS_NOISE_EXTERNAL) : (K.noise == NDS_NOISE_EXTERNAL);
  const int cd = S::CV >= 0 ? cv_dim(S::CV, ND) : K.cd;  // 1 or 2; a 1-D colvar runs the 2-D hill pass with c1 = 0

  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < MTD_NBUF; b++) mbar_init(&full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t tile_counter = 0;
  const long long n_tiles = (A.n_hills + MTD_TILE - 1) / MTD_TILE;

  R x[ND], v[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) { x[d] = owner ? A.x[d * M + w] : (R)1; v[d] = owner ? A.v[d * M + w] : (R)0; }
  WalkerBias<R, ND> wb;
#pragma unroll
  for (int k = 0; k < ND; k++) { wb.k[k] = (R)0; wb.r0[k] = (R)0; }
  if ((bias_mask & BIAS_UMB_WALKER) && owner) {
#pragma unroll
    for (int k = 0; k < ND; k++)
      if (k < cd) { wb.k[k] = A.umb_k[k * M + w]; wb.r0[k] = A.umb_r0[k * M + w]; }
  }
  const unsigned long long gid = (unsigned long long)(A.walker_offset + (owner ? w : 0));

  if (A.need_init_v) {
    R z[NDS_NPC];
    normals6<R>(K.rk_init, 0ull, gid, z);
#pragma unroll
    for (int d = 0; d < ND; d++) v[d] = K.vel_scale * z[d];
  }

  // full evaluation at x: colvar + potential + other biases on the owner lanes, hill pass by the warp
  Eval<R, ND> e;
  auto evaluate = [&](bool energy) {
    cv_eval<R, ND>(K, cv, K.rot, x, e.cv);
    R c0[W], c1[W];
#pragma unroll
    for (int q = 0; q < W; q++) {
      c0[q] = __shfl_sync(0xffffffffu, e.cv.c[0], q);
      c1[q] = cd > 1 ? __shfl_sync(0xffffffffu, e.cv.c[ND > 1 ? 1 : 0], q) : (R)0;
    }
    R Vm, g[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) g[d] = (R)0;
    if (energy) mtd_hill_pass<R, W, NW, HS, true>(K, A.hills, A.n_hills, n_tiles, smem, full, red, tile_counter, c0, c1, Vm, g[0], g[1]);
    else mtd_hill_pass<R, W, NW, HS, false>(K, A.hills, A.n_hills, n_tiles, smem, full, red, tile_counter, c0, c1, Vm, g[0], g[1]);
    if (energy || (other_bias & BIAS_PE)) pot_eval<R, ND, true>(K, pot, tcv, x, e.pe, e.fp);
    else pot_eval<R, ND, false>(K, pot, tcv, x, e.pe, e.fp);
    R vm_other;
    if (other_bias)
      bias_eval<R, ND>(K, other_bias, cv, cd, x, e.cv, wb, nullptr, 0, 1, e.pe, e.fp, e.be, e.fb, vm_other);
    else {
      e.be = (R)0;
#pragma unroll
      for (int d = 0; d < ND; d++) e.fb[d] = (R)0;
    }
    if (A.n_hills > 0) {  // gaus.py:257-258: f = (VRarray . dcolv/sigma^2) . J
      R fm[ND];
      cv_pullback<R, ND>(cv, K.rot, x, e.cv, g, fm);
#pragma unroll
      for (int d = 0; d < ND; d++) e.fb[d] += fm[d];
    }
    e.vm = Vm;
    e.be += Vm;
  };

  evaluate(true);
  const bool writer = owner && (warp / NW) == 0;  // one copy of each walker writes back
  if (A.refresh_vr && writer) A.vr[w] = e.vm;
  R f[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) f[d] = e.fp[d] + e.fb[d];

  long long s = A.step0;
  const long long s_end = A.step0 + A.nsteps;
  while (s < s_end) {
    const long long grp = s / G;
    const int first = (int)(s - grp * G);
    R zb[G * NN > 0 ? G * NN : 1];
    if (NN > 0 && !ext_noise) {
#pragma unroll
      for (int c = 0; c < CG; c++) {
        R z6[NDS_NPC];
        normals6<R>(K.rk_step, (unsigned long long)(grp * CG + c), gid, z6);
#pragma unroll
        for (int j = 0; j < NDS_NPC; j++)
          if (NDS_NPC * c + j < G * NN) zb[NDS_NPC * c + j] = z6[j];
      }
    }
#pragma unroll
    for (int i = 0; i < G; i++) {
      if (i >= first && s < s_end) {
        R xi[ND], eta[ND];
        if (NN > 0) {
          if (ext_noise) {
            const double* row = A.noise + ((s - A.noise_first) * NN) * M + (owner ? w : 0);
#pragma unroll
            for (int d = 0; d < ND; d++) {
              xi[d] = (R)row[(long long)d * M];
              if (INTEG == NDS_INT_LANGEVIN2) eta[d] = (R)row[(long long)(ND + d) * M];
            }
          } else {
#pragma unroll
            for (int d = 0; d < ND; d++) {
              xi[d] = zb[i * NN + d];
              if (INTEG == NDS_INT_LANGEVIN2) eta[d] = zb[i * NN + ND + d];
            }
          }
        }
#pragma unroll
        for (int d = 0; d < ND; d++) {  // md.py:164-178
          R dx;
          if (INTEG == NDS_INT_LANGEVIN) {
            v[d] += K.dt / (R)2.0 * f[d] / K.m;
            v[d] = K.c1 * v[d] + K.c2 * xi[d];
            dx = v[d] * K.dt;
          } else if (INTEG == NDS_INT_LANGEVIN2) {
            v[d] += K.c1 * f[d] / K.m - K.c2 * v[d] + K.c3 * xi[d] - K.c4 * eta[d];
            dx = K.dt * v[d] + K.c5 * eta[d] * K.dt;
          } else {
            v[d] += K.dt / (R)2.0 * f[d] / K.m;
            dx = v[d] * K.dt;
          }
          x[d] = x[d] + dx;
        }
        evaluate(s + 1 == s_end);  // the last evaluation of the launch also yields the observables
#pragma unroll
        for (int d = 0; d < ND; d++) f[d] = e.fp[d] + e.fb[d];
#pragma unroll
        for (int d = 0; d < ND; d++) {  // md.py:190-198
          if (INTEG == NDS_INT_LANGEVIN) {
            v[d] += f[d] / (R)2.0 / K.m * K.dt;
            v[d] = K.c1 * v[d] + K.c2 * xi[d];
          } else if (INTEG == NDS_INT_LANGEVIN2) {
            v[d] += K.c1 * f[d] / K.m - K.c2 * v[d] + K.c3 * xi[d] - K.c4 * eta[d];
          } else {
            v[d] += f[d] / K.m * K.dt / (R)2.0;
          }
        }
        s += 1;
      }
    }
  }

  if (writer) {
    R v2 = (R)0;
#pragma unroll
    for (int d = 0; d < ND; d++) {
      A.x[d * M + w] = x[d];
      A.v[d * M + w] = v[d];
      A.f[d * M + w] = e.fp[d];
      A.fixf[d * M + w] = e.fb[d];
      v2 += v[d] * v[d];
    }
    A.colv[0 * M + w] = e.cv.c[0];
    if (cd > 1) A.colv[1 * M + w] = e.cv.c[ND > 1 ? 1 : 0];
    const R ke = v2 * K.m / (R)2.0;
    A.thermo[0 * M + w] = A.begin_mode ? ke / (R)ND * (R)2.0 / K.kB : ke / (R)ND / K.kB * (R)2.0;
    A.thermo[1 * M + w] = e.pe;
    A.thermo[2 * M + w] = ke;
    A.thermo[3 * M + w] = e.be;
    A.thermo[4 * M + w] = A.begin_mode ? ke + e.pe : e.pe + ke + e.be;
  }
}

template <class S, int W, int NW, int HS>
cudaError_t launch_mtd_block(const Consts<typename S::R>& K, const BlockArgs<typename S::R>& A,
                             cudaStream_t st) {
  using R = typename S::R;
  const size_t smem = (size_t)MTD_NBUF * MTD_TILE_BYTES;
  cudaError_t e = cudaFuncSetAttribute(mtd_block<S, W, NW, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long per_cta = (long long)W * NW;
  const unsigned grid = (unsigned)((A.M + per_cta - 1) / per_cta);
  mtd_block<S, W, NW, HS><<<grid, NW * HS * 32, smem, st>>>(K, A);
  return cudaGetLastError();
}

}  // namespace nds
