// nds_pair_kernel.cuh -- TWO LANES PER WALKER for small ensembles of the C5 grid-mode configuration
// (5-D Mueller-Brown in Proj5dto2d space, first-order Langevin, metadynamics bias from the accumulated grid).
//
// Same loop as md_block (NDRun.run body ndrun.py:227-251 -> MD.update md.py:112-231 -> Mueller5d.compute
// mueller.py:112-142 -> Gaussian.compute gaus.py:236-261 through grid_interp), for the regime BASELINE's C5 names:
// 8192 walkers per GPU.  One thread per walker leaves 336 of the 592 warp schedulers of a B200 idle there and the
// 256 busy ones run one warp each, bound by the LENGTH of the dependent instruction stream of a step (about 260
// instructions of physics + 70 of noise at IPC 0.27, profiles/r02_c5grid_raw.csv).  Here a walker is a lane PAIR and
// every part of the step splits in two with little replication:
//   lane A: dimensions 0, 1, 4 (they form colvar q0), Gaussians 0, 1 of the surface, grid column i,
//           the Philox call of EVEN steps;
//   lane B: dimensions 2, 3 (colvar q1), Gaussians 2, 3, grid column i + 1, the Philox call of ODD steps.
// Per step the lanes exchange: their colvar (1 shuffle), the colvar-space force components they owe each other
// (1 shuffle), and -- once per two steps -- the normals drawn for the partner's dimensions (3 shuffles).  Replicated
// on both lanes: the cell search and the Hermite basis of the interpolation (~45 instructions).  The noise STREAM
// is the one of md_block (one Philox4x32-10 call per walker and step, normals 0..4 of the call for dimensions 0..4),
// so results agree with md_block_wide up to the order of the floating-point sums of the force.
// The Philox / Box-Muller chain of the NEXT pair of steps is drawn in the same basic block as this pair's physics: the
// independent integer chain fills the bubbles of the dependent FP32 chain (41.9 -> 37.5 us per 100 steps).
// Launch entry evaluates the force with the loop's own arithmetic (sums of products are explicit fmaf chains, so
// every inlined copy rounds alike and launch boundaries do not change results) and refreshes VR after a deposition
// through grid_interp; launch exit (observables, state write-back) reuses the generic eval_all on both lanes of the
// pair; lane A stores.
#pragma once
#include <cstdlib>
#include "nds_variants.h"

namespace nds {

constexpr int NDS_PAIR_BLOCK = 128;        // threads per CTA = 64 walkers
// walkers per launch up to which the pair kernel is chosen: it runs 2 x 212 instructions per walker-step against
// md_block_wide's ~300, so it only pays while the GPU is short of warps -- measured per 100 steps, pair / wide:
// 8192 walkers 37.5 / 48.4 us, 16384: 40.7 / 45.5, 32768: 65.0 / 57.0, 65536: 106 / 91 (exp_pair_threshold.py)
constexpr long long NDS_PAIR_MAX = 16384;

template <class S>
__global__ void __launch_bounds__(NDS_PAIR_BLOCK)
md_pair_grid(const __grid_constant__ Consts<float> K, const __grid_constant__ BlockArgs<float> A) {
  static_assert(is_scalar_f32<typename S::R>::value && S::ND == 5 && S::CV == NDS_CV_PROJ5DTO2D &&
                    S::TCV == NDS_CV_PROJ5DTO2D && S::INTEG == NDS_INT_LANGEVIN &&
                    S::BIAS == (BIAS_MTD | BIAS_MTD_GRID) && S::METHOD == NDS_METHOD_MD &&
                    S::NOISE == NDS_NOISE_PHILOX && S::DUMP == 0 &&
                    (S::POT == NDS_POT_MUELLER5D || S::POT == NDS_POT_MUELLER5DSHL),
                "md_pair_grid: the C5 grid-mode configuration only");
  constexpr int ND = 5;
  constexpr int BM = BIAS_MTD | BIAS_MTD_GRID;
  const long long n = A.n_active > 0 ? A.n_active : A.M;
  const long long M = A.M;
  const int half = threadIdx.x & 1;  // 0: lane A, 1: lane B
  long long w = A.slot0 + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1);
  const bool valid = w < n;  // a pair beyond the ensemble shadows the last walker (it never stores)
  if (!valid) w = n - 1;
  const bool writer = valid && half == 0;
  const unsigned long long gid = (unsigned long long)(A.walker_offset + w);
  const GridView<float> gv{A.grid, 0, half ? nullptr : A.grid_outside};  // one count per walker

  // ---- this lane's share of the state: dimensions (0, 1, 4) or (2, 3, -)
  float x[3], v[3], f[3];
  x[0] = A.x[(half ? 2 : 0) * M + w]; x[1] = A.x[(half ? 3 : 1) * M + w]; x[2] = half ? 0.f : A.x[4 * M + w];
  v[0] = A.v[(half ? 2 : 0) * M + w]; v[1] = A.v[(half ? 3 : 1) * M + w]; v[2] = half ? 0.f : A.v[4 * M + w];
  const float cz = half ? 0.f : 1.0e-7f;  // weight of the third dimension in q0 (five.py:52) and in its Jacobian (five.py:66)
  // the two Gaussians of this lane (mueller.py:8-16 through fill_consts): fast fp32 form of four_gauss
  float gx[2], gy[2], ga[2], gb[2], gc[2], gl[2], sg[2];
#pragma unroll
  for (int m = 0; m < 2; m++) {
    const int i = 2 * half + m;
    gx[m] = K.gx[i]; gy[m] = K.gy[i]; ga[m] = K.fa[i]; gb[m] = K.fb[i]; gc[m] = K.fc[i]; gl[m] = K.flA[i];
    sg[m] = ((0x7 >> i) & 1) ? 1.f : -1.f;  // force = -A e^.. grad: +u G for the negative amplitudes (terms 0..2)
  }
  const int nx = K.grid_nx, ny = K.grid_ny;
  // the whole bias grid in shared memory when it fits (1-D TMA bulk copies, like md_block_wide): the node loads sit on
  // the critical path of a step, and an L2 hit costs ~10x a shared-memory load
  extern __shared__ __align__(128) unsigned char pair_smem[];
  const float* gsrc = A.grid;
  if (A.grid_smem) {
    __shared__ uint64_t grid_bar;
    grid_to_smem(pair_smem, A.grid, 16u * (uint32_t)(nx * ny), &grid_bar);
    gsrc = reinterpret_cast<const float*>(pair_smem);
  }
  const bool in_smem = A.grid_smem != 0;

  // total force on this lane's dimensions at the pair's current positions (potential + bias)
  auto force = [&]() {
    // ---- colvar of my group (five.py:50-57), the partner's through one shuffle
    // (sums of two or more products are written as explicit fmaf chains: force() is inlined at several places, the
    // compiler is free to pick which product of a*b + c*d it fuses, and results must not depend on which copy ran --
    // launch boundaries may fall anywhere)
    const float ss = fmaf(x[1], x[1], fmaf(x[0], x[0], cz * (x[2] * x[2])));
    const float iq = nds_rsqrt(ss), q = ss * iq;
    const float qp = __shfl_xor_sync(0xffffffffu, q, 1);
    const float q0 = half ? qp : q, q1 = half ? q : qp;
    // ---- bias grid: cell of (q0, q1); this lane loads column i + half (two nodes)
    const float u = (q0 - K.grid_lo[0]) * K.grid_invh[0], ww = (q1 - K.grid_lo[1]) * K.grid_invh[1];
    const bool out = !(u >= 0.f) || !(u <= (float)(nx - 1)) || !(ww >= 0.f) || !(ww <= (float)(ny - 1));
    int ci = (int)u, cj = (int)ww;
    ci = ci > nx - 2 ? nx - 2 : ci; cj = cj > ny - 2 ? ny - 2 : cj;
    ci = out ? 0 : ci; cj = out ? 0 : cj;
    const float t = u - (float)ci, sc = ww - (float)cj;
    float4 n0, n1;
    if (in_smem) {  // uniform over the launch
      n0 = grid_load<float, true>(gsrc, cj * nx + ci + half); n1 = grid_load<float, true>(gsrc, (cj + 1) * nx + ci + half);
    } else {
      n0 = grid_load<float, false>(gsrc, cj * nx + ci + half); n1 = grid_load<float, false>(gsrc, (cj + 1) * nx + ci + half);
    }
    if (out) {  // a point outside the box sees no bias and is counted (grid_fetch)
      n0 = make_float4(0.f, 0.f, 0.f, 0.f); n1 = n0;
      if (writer) atomicAdd(A.grid_outside, 1ull);
    }
    // ---- my two Gaussians of the surface in (q0, q1): mueller.py:30-44 (four_gauss, fp32 form)
    float fx = 0.f, fy = 0.f;
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const float dx = q0 - gx[m], dy = q1 - gy[m];
      const float Gx = fmaf(gb[m], dy, ga[m] * dx), Gy = fmaf(gb[m], dx, gc[m] * dy);
      const float su = sg[m] * nds_exp2(fmaf(dy, Gy, fmaf(dx, Gx, gl[m])));
      fx = fmaf(su, Gx, fx); fy = fmaf(su, Gy, fy);
    }
    // ---- my column of the bicubic Hermite interpolation (grid_combine)
    // along t only the basis functions of MY column are needed: h01(t) = h00(1 - t), h11(t) = -h10(1 - t), so one
    // set of formulas serves both lanes through tau = t (column i) or 1 - t (column i + 1)
    float as[2], bs[2], das[2], dbs[2];
    hermite_basis<float>(sc, as, bs, das, dbs);
    const float tau = half ? 1.f - t : t, ups = 1.f - tau, sgn = half ? -1.f : 1.f;
    const float uu = ups * ups, stu = sgn * tau;
    const float ap = fmaf(2.f, tau, 1.f) * uu, bp = stu * uu;
    const float dap = -6.f * (stu * ups), dbp = ups * fmaf(-3.f, tau, 1.f);
    const float vp = fmaf(bs[1], n1.z, fmaf(as[1], n1.x, fmaf(bs[0], n0.z, as[0] * n0.x)));
    const float yp = fmaf(bs[1], n1.w, fmaf(as[1], n1.y, fmaf(bs[0], n0.w, as[0] * n0.y)));
    const float vps = fmaf(dbs[1], n1.z, fmaf(das[1], n1.x, fmaf(dbs[0], n0.z, das[0] * n0.x)));
    const float yps = fmaf(dbs[1], n1.w, fmaf(das[1], n1.y, fmaf(dbs[0], n0.w, das[0] * n0.y)));
    const float dVt = fmaf(dbp, yp, dap * vp);
    const float dVs = fmaf(bp, yps, ap * vps);
    // colvar-space force, my part: surface + bias (g = -dV/dc)
    const float P0 = fmaf(-dVt, K.grid_invh[0], fx), P1 = fmaf(-dVs, K.grid_invh[1], fy);
    // A needs the total along q0, B along q1: each sends the component it does not need
    const float rp = __shfl_xor_sync(0xffffffffu, half ? P0 : P1, 1);
    const float G = (half ? P1 : P0) + rp;
    // ---- pull-back through the Jacobian of my colvar (five.py:59-71)
    const float i0 = iq * G;
    f[0] = x[0] * i0; f[1] = x[1] * i0; f[2] = cz * x[2] * i0;
  };

  // ---- launch entry: forces at the current positions (md.py:101-110 / gaus.py:138-146) -- the same arithmetic as
  //      inside the loop, so launch boundaries do not change results
  force();
  if (A.refresh_vr) {  // Gaussian.VR after a deposition (gaus.py:292-293): bias energy at the walker's colvar
    float xa[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) xa[d] = A.x[d * M + w];
    CvOut<float, ND> cvo;
    cv_eval<float, ND>(K, S::CV, K.rot, xa, cvo);
    float Vm, gq[ND];
    grid_interp<float, ND, false>(K, 2, A.grid, cvo.c, Vm, gq, nullptr);
    if (writer) A.vr[w] = Vm;
  }

  const long long s_end = A.step0 + A.nsteps;
  // one MD.update (md.py:164-198) of this lane's dimensions with the noise xi (= c2 * normal)
  auto step = [&](const float (&xi)[3]) {
    // ---- first half kick + drift, md.py:164-178
#pragma unroll
    for (int k = 0; k < 3; k++) {
      v[k] = fmaf(K.hdtm, f[k], v[k]);
      v[k] = fmaf(K.c1, v[k], xi[k]);
      x[k] = fmaf(v[k], K.dt, x[k]);
    }
    // ---- colvar, potential and bias at the new positions, md.py:179-188
    force();
    // ---- second half kick, md.py:190-198
#pragma unroll
    for (int k = 0; k < 3; k++) {
      v[k] = fmaf(K.hdtm, f[k], v[k]);
      v[k] = fmaf(K.c1, v[k], xi[k]);
    }
  };
  // the normals of the step pair (s2, s2 + 1): lane A draws the call of the even step, lane B that of the odd one,
  // and each hands the partner the normals of the partner's dimensions -- A owes (z2, z3, -), B owes (z0, z1, z4)
  auto draw = [&](long long s2, float (&xe)[3], float (&xo)[3]) {
    float z[NDS_NPC];
    normals6<float>(K.rk_step, (unsigned long long)(s2 + half), gid, z, K.bm_c2);  // c2 * xi (PRESCALED)
    const float r0 = __shfl_xor_sync(0xffffffffu, half ? z[0] : z[2], 1);
    const float r1 = __shfl_xor_sync(0xffffffffu, half ? z[1] : z[3], 1);
    const float r2 = __shfl_xor_sync(0xffffffffu, half ? z[4] : 0.f, 1);
    xe[0] = half ? r0 : z[0]; xe[1] = half ? r1 : z[1]; xe[2] = half ? 0.f : z[4];  // even step (drawn by A)
    xo[0] = half ? z[2] : r0; xo[1] = half ? z[3] : r1; xo[2] = half ? 0.f : r2;    // odd step (drawn by B)
  };
  long long s = A.step0;
  float xe[3], xo[3];
  if (s < s_end && (s & 1)) {  // the launch starts on an odd step: second half of its pair
    draw(s - 1, xe, xo);
    step(xo);
    s += 1;
  }
  // whole pairs, software-pipelined: the Philox / Box-Muller chain of the NEXT pair is independent of this pair's
  // physics, so it fills the issue slots the dependent FP32 chain leaves empty (one warp per scheduler)
  if (s + 1 < s_end) draw(s, xe, xo);
  while (s + 1 < s_end) {
    float ne[3], no[3];
    draw(s + 2, ne, no);  // (one pair beyond the launch at the end: drawn, never used)
    step(xe);
    step(xo);
#pragma unroll
    for (int k = 0; k < 3; k++) { xe[k] = ne[k]; xo[k] = no[k]; }
    s += 2;
  }
  if (s < s_end) {  // the launch ends on an even step: first half of its pair
    draw(s, xe, xo);
    step(xe);
  }

  // ---- launch exit: gather the pair's state, observables at the final positions (md.py:210-217,226-229)
  const float ox0 = __shfl_xor_sync(0xffffffffu, x[0], 1), ox1 = __shfl_xor_sync(0xffffffffu, x[1], 1);
  const float ox2 = __shfl_xor_sync(0xffffffffu, x[2], 1);
  const float ov0 = __shfl_xor_sync(0xffffffffu, v[0], 1), ov1 = __shfl_xor_sync(0xffffffffu, v[1], 1);
  const float ov2 = __shfl_xor_sync(0xffffffffu, v[2], 1);
  float xf[ND], vf[ND];
  WalkerBias<float, ND> wb;
#pragma unroll
  for (int k = 0; k < ND; k++) { wb.k[k] = 0.f; wb.r0[k] = 0.f; }
  xf[0] = half ? ox0 : x[0]; xf[1] = half ? ox1 : x[1]; xf[2] = half ? x[0] : ox0; xf[3] = half ? x[1] : ox1;
  xf[4] = half ? ox2 : x[2];
  vf[0] = half ? ov0 : v[0]; vf[1] = half ? ov1 : v[1]; vf[2] = half ? v[0] : ov0; vf[3] = half ? v[1] : ov1;
  vf[4] = half ? ov2 : v[2];
  Eval<float, ND> o;
  eval_all<float, ND, true>(K, S::POT, S::CV, S::TCV, 2, BM, xf, wb, nullptr, 0, 1, o, gv);
  if (!writer) return;
  float v2 = 0.f;
#pragma unroll
  for (int d = 0; d < ND; d++) {
    A.x[d * M + w] = xf[d];
    A.v[d * M + w] = vf[d];
    A.f[d * M + w] = o.fp[d];
    A.fixf[d * M + w] = o.fb[d];
    v2 += vf[d] * vf[d];
  }
  A.colv[0 * M + w] = o.cv.c[0];
  A.colv[1 * M + w] = o.cv.c[1];
  const float ke = v2 * K.m / 2.0f;  // md.py:227
  A.thermo[1 * M + w] = o.pe;
  A.thermo[3 * M + w] = o.be;
  A.thermo[0 * M + w] = ke / (float)ND / K.kB * 2.0f;  // md.py:228
  A.thermo[2 * M + w] = ke;
  A.thermo[4 * M + w] = o.pe + ke + o.be;  // md.py:229
}

// C5 grid-mode variant: the pair kernel for step launches of small ensembles, md_block_wide for everything else
// (NDRun.begin, initial velocities, large ensembles, NDS_NO_PAIR=1)
template <class S>
cudaError_t launch_md_pair_or_wide(const Consts<float>& K, const BlockArgs<float>& A, cudaStream_t st) {
  const long long n = (A.n_active > 0 ? A.n_active : A.M) - A.slot0;
  const bool off = getenv("NDS_NO_PAIR") != nullptr;
  if (off || n > NDS_PAIR_MAX || n < 1 || A.nsteps <= 0 || A.begin_mode || A.need_init_v || A.ids || A.noise)
    return launch_md_block_wide<S>(K, A, st);
  const unsigned grid = (unsigned)((2 * n + NDS_PAIR_BLOCK - 1) / NDS_PAIR_BLOCK);
  const size_t smem = A.grid_smem ? (size_t)16 * (size_t)(K.grid_nx * K.grid_ny) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(md_pair_grid<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  md_pair_grid<S><<<grid, NDS_PAIR_BLOCK, smem, st>>>(K, A);
  return cudaGetLastError();
}

template <class S>
Variant make_variant_pair(const char* name) {
  return Variant{VariantKey{prec_of<typename S::R>::value, S::ND, S::POT, S::CV, S::TCV, S::INTEG,
                            S::BIAS, S::METHOD, S::NOISE, S::DUMP, S::NBASIN},
                 name, (void*)&launch_md_pair_or_wide<S>, true};
}

}  // namespace nds
