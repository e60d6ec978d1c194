// nds_physics.cuh -- device code of the built-in collective variables, potentials and biases.
//
// Each function names the reference code it replaces (file:line in mir-group/NDSimulator).  All code is
// templated on the real type R: the fp64 instantiation keeps the reference's operation order (parity
// within 1e-10, tests/test_parity_gpu.py); the fp32 instantiation is free to use MUFU intrinsics and
// algebraically equivalent forms (statistical parity).
//
// Every `id` argument (potential / colvar id) is either a compile-time constant folded by inlining
// (hot-path kernel variants) or a value read from the constant block (generic kernels); the switch
// statements below serve both.
#pragma once
#include "nds_common.cuh"

namespace nds {

#define NDS_DEV __device__ __forceinline__

// is_f32: fp32 arithmetic (scalar float or the packed pair f2) -> MUFU / re-associated fast forms
template <typename R> struct is_f32 { static constexpr bool value = false; };
template <> struct is_f32<float> { static constexpr bool value = true; };
template <> struct is_f32<f2> { static constexpr bool value = true; };
template <typename R> struct is_scalar_f32 { static constexpr bool value = false; };
template <> struct is_scalar_f32<float> { static constexpr bool value = true; };

NDS_DEV float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
NDS_DEV float nds_exp(float x) { return ex2_approx(x * 1.4426950408889634f); }
NDS_DEV double nds_exp(double x) { return exp(x); }
NDS_DEV float nds_exp2(float x) { return ex2_approx(x); }
NDS_DEV float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
NDS_DEV float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
NDS_DEV float nds_sqrt(float x) { return sqrt_approx(x); }
NDS_DEV double nds_sqrt(double x) { return sqrt(x); }
NDS_DEV float nds_rcp(float x) { return rcp_approx(x); }
NDS_DEV float nds_rsqrt(float x) { return rsqrt_approx(x); }
NDS_DEV double nds_rsqrt(double x) { return 1.0 / sqrt(x); }
NDS_DEV double nds_rcp(double x) { return 1.0 / x; }
NDS_DEV f2 nds_exp(f2 x) { return f2(nds_exp(x.v.x), nds_exp(x.v.y)); }
NDS_DEV f2 nds_exp2(f2 x) { return f2(ex2_approx(x.v.x), ex2_approx(x.v.y)); }
NDS_DEV f2 nds_sqrt(f2 x) { return f2(sqrt_approx(x.v.x), sqrt_approx(x.v.y)); }
NDS_DEV f2 nds_rcp(f2 x) { return f2(rcp_approx(x.v.x), rcp_approx(x.v.y)); }
NDS_DEV f2 nds_rsqrt(f2 x) { return f2(rsqrt_approx(x.v.x), rsqrt_approx(x.v.y)); }

// colvardim of a colvar id (colvars/*.py class attribute `colvardim`)
__host__ __device__ constexpr int cv_dim(int cv, int ndim) {
  switch (cv) {
    case NDS_CV_ORIGINAL: return ndim;
    case NDS_CV_ROTATE2D: case NDS_CV_PROJ5DTO2D: case NDS_CV_PROJ5DTO2DV2: return 2;
    default: return 1;  // including NDS_CV_PATH, NDS_CV_PATH5DTO2D
  }
}

// ------------------------------------------------------------------------------------ colvars
// Value c[0..cd) plus the few intermediates the Jacobian needs (inverse norms), so that the pull-back
// f_j = sum_k g_k J_kj can be applied sparsely without materialising J.
template <typename R, int ND>
struct CvOut {
  R c[ND];
  R q[2];   // norms (Proj*) for the Jacobian
  R iq[2];  // their reciprocals (fp32 fast path: one MUFU.RSQ gives both the norm and 1/norm)
  R jp[ND]; // path colvars: the 1 x ndim Jacobian row (colvars/path.py:90-102,144-180)
};

// Branduardi path variable s(q) = sum_i ind_i w_i / sum_i w_i / N,  w_i = exp(-|q - r_i|^2 / (2 sigma^2))
// (shifted by the smallest exponent like path.py:75-77), and ds/dq.  The reference sums
// w_i w_j ind_i (r_i - r_j) over all pairs (path.py:95-100); the identical
//   sum_ij = D * sum_i ind_i w_i (r_i - rbar),  rbar = sum_j w_j r_j / D,  D = sum_j w_j
// needs one more pass instead of n^2 terms and keeps the same (r_i - r_j)-sized differences.
template <typename R, int PD>
NDS_DEV void path_eval(const Consts<R>& K, const R (&q)[PD], int pd, R& s, R (&ds)[PD]) {
  const int n = K.n_path;
  R dmin = (R)0;
  for (int i = 0; i < n; i++) {
    R d2 = (R)0;
#pragma unroll
    for (int d = 0; d < PD; d++)
      if (d < pd) { const R t = q[d] - K.path_ref[i * pd + d]; d2 += t * t * K.path_sig2_invh; }
    if (i == 0 || d2 < dmin) dmin = d2;
  }
  R D = (R)0, B = (R)0, C[PD];
#pragma unroll
  for (int d = 0; d < PD; d++) C[d] = (R)0;
  for (int i = 0; i < n; i++) {
    R d2 = (R)0;
#pragma unroll
    for (int d = 0; d < PD; d++)
      if (d < pd) { const R t = q[d] - K.path_ref[i * pd + d]; d2 += t * t * K.path_sig2_invh; }
    const R w = nds_exp(-(d2 - dmin));
    D += w;
    B += K.path_ind[i] * w;
#pragma unroll
    for (int d = 0; d < PD; d++)
      if (d < pd) C[d] += w * K.path_ref[i * pd + d];
  }
  s = B / D / K.path_N;
  R acc[PD];
#pragma unroll
  for (int d = 0; d < PD; d++) { acc[d] = (R)0; C[d] = C[d] / D; }
  for (int i = 0; i < n; i++) {
    R d2 = (R)0;
#pragma unroll
    for (int d = 0; d < PD; d++)
      if (d < pd) { const R t = q[d] - K.path_ref[i * pd + d]; d2 += t * t * K.path_sig2_invh; }
    const R iw = K.path_ind[i] * nds_exp(-(d2 - dmin));
#pragma unroll
    for (int d = 0; d < PD; d++)
      if (d < pd) acc[d] += iw * (K.path_ref[i * pd + d] - C[d]);
  }
#pragma unroll
  for (int d = 0; d < PD; d++) ds[d] = acc[d] * K.path_inv_sig2 / D / K.path_N;
}

// norm and reciprocal norm of a squared length
template <typename R>
NDS_DEV void norm_pair(R s, R& n, R& inv) {
  if constexpr (is_f32<R>::value) { inv = nds_rsqrt(s); n = s * inv; }
  else { n = nds_sqrt(s); inv = (R)0; }
}

template <typename R, int ND>
NDS_DEV void cv_eval(const Consts<R>& K, int cv, R rot, const R (&x)[ND], CvOut<R, ND>& o) {
  switch (cv) {
    case NDS_CV_ORIGINAL:  // colvars/misc.py:12-13
#pragma unroll
      for (int d = 0; d < ND; d++) o.c[d] = x[d];
      break;
    case NDS_CV_PATH:      // colvars/path.py:73-102 (path_dim == ndim)
      path_eval<R, ND>(K, x, ND, o.c[0], o.jp);
      break;
    default: break;
  }
  if constexpr (ND == 2) {
    switch (cv) {
      case NDS_CV_EXTRACTX: o.c[0] = x[0]; break;          // two.py:21-22
      case NDS_CV_EXTRACTY: o.c[0] = x[1]; break;          // two.py:10-11
      case NDS_CV_XMY: o.c[0] = x[0] - x[1]; break;        // two.py:32-33
      case NDS_CV_XPY: o.c[0] = x[0] + x[1]; break;        // two.py:43-44
      case NDS_CV_ROTATE2D: {                              // two.py:52-58
        const R r = rot;
        o.c[0] = x[0] * r + x[1] * r;
        o.c[1] = x[0] * r + x[1] * (-r);
      } break;
      case NDS_CV_PROJ2DTO1D:                              // two.py:67-69
        norm_pair<R>(x[0] * x[0] + x[1] * x[1], o.q[0], o.iq[0]);
        o.c[0] = o.q[0];
        break;
      default: break;
    }
  }
  if constexpr (ND == 5) {
    switch (cv) {
      case NDS_CV_PROJ5DTO1DX0: o.c[0] = x[0]; break;      // five.py:9-11
      case NDS_CV_PROJ5DTO1DX2: o.c[0] = x[2]; break;      // five.py:22-24
      case NDS_CV_PROJ5DTO1D:                              // five.py:34-38
        norm_pair<R>(x[0] * x[0] + x[1] * x[1] + (R)1e-7 * (x[4] * x[4]), o.q[0], o.iq[0]);
        o.c[0] = o.q[0];
        break;
      case NDS_CV_PROJ5DTO2D:                              // five.py:50-57
        norm_pair<R>(x[0] * x[0] + x[1] * x[1] + (R)1e-7 * (x[4] * x[4]), o.q[0], o.iq[0]);
        norm_pair<R>(x[2] * x[2] + x[3] * x[3], o.q[1], o.iq[1]);
        o.c[0] = o.q[0];
        o.c[1] = o.q[1];
        break;
      case NDS_CV_PROJ5DTO2DV2: {                          // five.py:77-81
        const R d = x[1] - x[4];
        norm_pair<R>(x[0] * x[0] + d * d, o.q[0], o.iq[0]);
        norm_pair<R>(x[2] * x[2] + x[3] * x[3], o.q[1], o.iq[1]);
        o.c[0] = o.q[0];
        o.c[1] = o.q[1];
      } break;
      case NDS_CV_PATH5DTO2D: {                            // path.py:121-180: path variable of Proj5dto2d(x)
        R q[2], dq[2];
        norm_pair<R>(x[0] * x[0] + x[1] * x[1] + (R)1e-7 * (x[4] * x[4]), o.q[0], o.iq[0]);
        norm_pair<R>(x[2] * x[2] + x[3] * x[3], o.q[1], o.iq[1]);
        q[0] = o.q[0]; q[1] = o.q[1];
        path_eval<R, 2>(K, q, 2, o.c[0], dq);
        // jab1 (1x2) . jab2 (2x5), path.py:160-169
        o.jp[0] = dq[0] * (x[0] / o.q[0]);
        o.jp[1] = dq[0] * (x[1] / o.q[0]);
        o.jp[2] = dq[1] * (x[2] / o.q[1]);
        o.jp[3] = dq[1] * (x[3] / o.q[1]);
        o.jp[4] = dq[0] * ((R)1.0e-7 * x[4] / o.q[0]);
      } break;
      default: break;
    }
  }
}

// f[j] = sum_k g[k] * J[k][j]  with J = colvar.jacobian(x)   (e.g. mueller.py:142, umb.py:31, gaus.py:258)
template <typename R, int ND>
NDS_DEV void cv_pullback(int cv, R rot, const R (&x)[ND], const CvOut<R, ND>& o, const R (&g)[ND],
                         R (&f)[ND]) {
#pragma unroll
  for (int d = 0; d < ND; d++) f[d] = (R)0;
  if (cv == NDS_CV_ORIGINAL) {  // identity Jacobian, misc.py:15-16
#pragma unroll
    for (int d = 0; d < ND; d++) f[d] = g[d];
    return;
  }
  if (cv == NDS_CV_PATH || cv == NDS_CV_PATH5DTO2D) {  // 1 x ndim Jacobian row stored by cv_eval
#pragma unroll
    for (int d = 0; d < ND; d++) f[d] = g[0] * o.jp[d];
    return;
  }
  if constexpr (ND == 2) {
    switch (cv) {
      case NDS_CV_EXTRACTX: f[0] = g[0]; break;                        // two.py:19
      case NDS_CV_EXTRACTY: f[1] = g[0]; break;                        // two.py:8
      case NDS_CV_XMY: f[0] = g[0]; f[1] = -g[0]; break;               // two.py:30
      case NDS_CV_XPY: f[0] = g[0]; f[1] = g[0]; break;                // two.py:41
      case NDS_CV_ROTATE2D: {                                          // two.py:53-54
        const R r = rot;
        f[0] = g[0] * r + g[1] * r;
        f[1] = g[0] * r + g[1] * (-r);
      } break;
      case NDS_CV_PROJ2DTO1D:                                          // two.py:71-73
        if constexpr (is_f32<R>::value) {
          const R i0 = g[0] * o.iq[0];
          f[0] = x[0] * i0; f[1] = x[1] * i0;
        } else {
          f[0] = g[0] * (x[0] / o.q[0]); f[1] = g[0] * (x[1] / o.q[0]);
        }
        break;
      default: break;
    }
  }
  if constexpr (ND == 5) {
    switch (cv) {
      case NDS_CV_PROJ5DTO1DX0: f[0] = g[0]; break;                    // five.py:7
      case NDS_CV_PROJ5DTO1DX2: f[2] = g[0]; break;                    // five.py:20
      case NDS_CV_PROJ5DTO1D:                                          // five.py:40-44 (2e-7)
        if constexpr (is_f32<R>::value) {
          const R i0 = o.iq[0] * g[0];
          f[0] = x[0] * i0; f[1] = x[1] * i0; f[4] = (R)2.0e-7 * x[4] * i0;
        } else {
          f[0] = g[0] * (x[0] / o.q[0]);
          f[1] = g[0] * (x[1] / o.q[0]);
          f[4] = g[0] * ((R)2.0 * (R)1.0e-7 * x[4] / o.q[0]);
        }
        break;
      case NDS_CV_PROJ5DTO2D:                                          // five.py:59-71
        if constexpr (is_f32<R>::value) {
          const R i0 = o.iq[0] * g[0], i1 = o.iq[1] * g[1];
          f[0] = x[0] * i0; f[1] = x[1] * i0; f[4] = (R)1.0e-7 * x[4] * i0;
          f[2] = x[2] * i1; f[3] = x[3] * i1;
        } else {
          f[0] = g[0] * (x[0] / o.q[0]);
          f[1] = g[0] * (x[1] / o.q[0]);
          f[4] = g[0] * ((R)1.0e-7 * x[4] / o.q[0]);
          f[2] = g[1] * (x[2] / o.q[1]);
          f[3] = g[1] * (x[3] / o.q[1]);
        }
        break;
      case NDS_CV_PROJ5DTO2DV2: {                                      // five.py:83-98
        const R d = x[1] - x[4];
        if constexpr (is_f32<R>::value) {
          const R i0 = o.iq[0] * g[0], i1 = o.iq[1] * g[1];
          f[0] = x[0] * i0; f[1] = d * i0; f[4] = -(d * i0);
          f[2] = x[2] * i1; f[3] = x[3] * i1;
        } else {
          f[0] = g[0] * (x[0] / o.q[0]);
          f[1] = g[0] * (d / o.q[0]);
          f[4] = g[0] * (-d / o.q[0]);
          f[2] = g[1] * (x[2] / o.q[1]);
          f[3] = g[1] * (x[3] / o.q[1]);
        }
      } break;
      default: break;
    }
  }
}

// ---------------------------------------------------------------------------------- potentials
// Sum of four anisotropic Gaussians in (q0, q1): V = sum A_i exp(a dx^2 + b dx dy + c dy^2) and the
// gradient force g = -dV/dq.   potentials/mueller.py:30-44, potentials/three_hole.py:36-50.
// BMASK: bit i set when term i has a non-zero cross coefficient b_i (class constants of the reference:
// Mueller 0b1100, ThreeHole 0); lets the fp32 form drop the b terms at compile time.
// NMASK: bit i set when the amplitude A_i is negative (class constants too: Mueller 0b0111, ThreeHole 0b1110).
template <typename R, bool ENERGY, int BMASK = 0xF, int NMASK = 0x0>
NDS_DEV void four_gauss(const Consts<R>& K, R q0, R q1, R& V, R& g0, R& g1) {
  if constexpr (is_f32<R>::value) {
    // fast form: G = k * grad(exponent) with k = log2(e)/2, so that  exponent*log2(e) = dx*Gx + dy*Gy
    // feeds MUFU.EX2 directly; |A/k| rides in the exponent as log2 and the sign of A is an operand
    // negation, so the force is +-2^s' * G.  8-10 FP32 ops + 1 MUFU per term.
    R fx = 0.f, fy = 0.f, e = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const R dx = q0 - K.gx[i], dy = q1 - K.gy[i];
      const bool cross = (BMASK >> i) & 1;
      const bool neg = (NMASK >> i) & 1;
      const R Gx = cross ? fmaf(K.fb[i], dy, K.fa[i] * dx) : K.fa[i] * dx;
      const R Gy = cross ? fmaf(K.fb[i], dx, K.fc[i] * dy) : K.fc[i] * dy;
      const R s = fmaf(dy, Gy, fmaf(dx, Gx, K.flA[i]));
      const R u = nds_exp2(s);  // |A_i| / k * exp(exponent)
      // force = -A_i e^.. grad(exponent): for a negative amplitude that is +u G
      if (i == 0) { fx = neg ? u * Gx : -u * Gx; fy = neg ? u * Gy : -u * Gy; }
      else { fx = neg ? fmaf(u, Gx, fx) : fmaf(-u, Gx, fx); fy = neg ? fmaf(u, Gy, fy) : fmaf(-u, Gy, fy); }
      if (ENERGY) e = (i == 0) ? (neg ? -u : u) : (neg ? e - u : e + u);
    }
    g0 = fx; g1 = fy;
    if (ENERGY) V = e * (R)0.72134752044448170368;  // k: A = fA * k
  } else {
    R v = 0, f0 = 0, f1 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const R dx = q0 - K.gx[i], dy = q1 - K.gy[i];
      const R tmp = K.gA[i] * nds_exp(K.ga[i] * (dx * dx) + K.gb[i] * dx * dy + K.gc[i] * (dy * dy));
      v += tmp;
      f0 -= tmp * ((R)2 * K.ga[i] * dx + K.gb[i] * dy);
      f1 -= tmp * ((R)2 * K.gc[i] * dy + K.gb[i] * dx);
    }
    V = v; g0 = f0; g1 = f1;
  }
}

// potential.compute(x) -> (V, f).  ENERGY=false lets the compiler drop the energy accumulation in the
// step loop (the reference evaluates it every step, md.py:180,210; it only feeds observables).
template <typename R, int ND, bool ENERGY>
NDS_DEV void pot_eval(const Consts<R>& K, int pot, int tcv, const R (&x)[ND], R& V, R (&f)[ND]) {
  V = (R)0;
#pragma unroll
  for (int d = 0; d < ND; d++) f[d] = (R)0;
  if (pot == NDS_POT_HARMONIC_QUARTIC) {  // extension: separable harmonic + quartic well
#pragma unroll
    for (int d = 0; d < ND; d++) {
      const R u = x[d] - K.hq_c[d], u2 = u * u;
      if (ENERGY) V += K.hq_k[d] * u2 / (R)2.0 + K.hq_q[d] * (u2 * u2) / (R)4.0;
      f[d] = -(K.hq_k[d] * u) - K.hq_q[d] * (u2 * u);
    }
    return;
  }
  if constexpr (ND == 1) {
    if (pot == NDS_POT_DOUBLEWELL1D) {  // double_well.py:16-25
      const R d1 = x[0] - K.dw_x1[0], d2 = x[0] - K.dw_x2[0];
      const R e1 = nds_exp(-(d1 * d1) * K.dw_invsig2), e2 = nds_exp(-(d2 * d2) * K.dw_invsig2);
      V = (R)-0.5 * (K.dw_K1 * e1 + K.dw_K2 * e2);
      f[0] = (-K.dw_K1 * e1 * d1 - K.dw_K2 * e2 * d2) * K.dw_invsig2;
    }
    // NDS_POT_FLAT1D: zero (flat.py:17-24)
  }
  if constexpr (ND == 2) {
    switch (pot) {
      case NDS_POT_MUELLER2D: {  // mueller.py:25-45
        four_gauss<R, ENERGY, 0xC, 0x7>(K, x[0], x[1], V, f[0], f[1]);
      } break;
      case NDS_POT_DOUBLEWELL: {  // double_well.py:44-53
        const R a0 = x[0] - K.dw_x1[0], a1 = x[1] - K.dw_x1[1];
        const R b0 = x[0] - K.dw_x2[0], b1 = x[1] - K.dw_x2[1];
        const R e1 = nds_exp(-(a0 * a0 + a1 * a1) * K.dw_invsig2);
        const R e2 = nds_exp(-(b0 * b0 + b1 * b1) * K.dw_invsig2);
        V = (R)-0.5 * (K.dw_K1 * e1 + K.dw_K2 * e2);
        f[0] = (-K.dw_K1 * e1 * a0 - K.dw_K2 * e2 * b0) * K.dw_invsig2;
        f[1] = (-K.dw_K1 * e1 * a1 - K.dw_K2 * e2 * b1) * K.dw_invsig2;
      } break;
      case NDS_POT_GAUSSIAN2D: {  // misc.py:7-26: A exp(a (x-20)^2 + c (y-20)^2)
        const R dx = x[0] - (R)20, dy = x[1] - (R)20;
        const R tmp = K.gA[0] * nds_exp(K.ga[0] * (dx * dx) + K.gc[0] * (dy * dy));
        V = tmp;
        f[0] = -tmp * (R)2 * K.ga[0] * dx;
        f[1] = -tmp * (R)2 * K.gc[0] * dy;
      } break;
      case NDS_POT_GAUSSIAN: {  // misc.py:32-43; the force uses x, not x - centre
        const R s2 = (R)100;
        const R V1 = -nds_exp(-((x[0] - (R)0.5) * (x[0] - (R)0.5) + x[1] * x[1]) / s2);
        const R V2 = -nds_exp(-((x[0] + (R)0.5) * (x[0] + (R)0.5) + x[1] * x[1]) / s2);
        V = V1 + V2;
        f[0] = V1 * (R)2 * x[0] / s2 + V2 * (R)2 * x[0] / s2;
        f[1] = V1 * (R)2 * x[1] / s2 + V2 * (R)2 * x[1] / s2;
      } break;
      default: break;  // NDS_POT_FLAT2D: zero (flat.py:5-14)
    }
  }
  if constexpr (ND == 2 || ND == 5) {
    const bool mueller_cv = (ND == 5) && (pot == NDS_POT_MUELLER5D || pot == NDS_POT_MUELLER5DSHL);
    const bool threehole = (ND == 2 && pot == NDS_POT_THREEHOLE2D) || (ND == 5 && pot == NDS_POT_THREEHOLE5D);
    if (mueller_cv || threehole) {
      // mueller.py:112-142 / three_hole.py:24-57,90-125: evaluate in true_colvar space, chain rule
      CvOut<R, ND> o;
      cv_eval<R, ND>(K, tcv, K.rot, x, o);
      R g[ND];
#pragma unroll
      for (int d = 0; d < ND; d++) g[d] = (R)0;
      R v = (R)0;
      if (threehole) four_gauss<R, ENERGY, 0x0, 0xE>(K, o.c[0], o.c[1], v, g[0], g[1]);
      else four_gauss<R, ENERGY, 0xC, 0x7>(K, o.c[0], o.c[1], v, g[0], g[1]);
      if (threehole) {  // quartic confinement, three_hole.py:51-55
        const R dx = o.c[0] - K.th_x1, dy = o.c[1] - K.th_y1;
        const R dx2 = dx * dx, dy2 = dy * dy;
        if (ENERGY) {
          v += K.th_D * (dx2 * dx2);
          if (ND == 2) v += K.th_E * (dy2 * dy2) + K.th_ref;
          else { v += K.th_E * (dy2 * dy2); v += K.th_ref; }
        }
        g[0] -= (R)4.0 * K.th_D * (dx2 * dx);
        g[1] -= (R)4.0 * K.th_E * (dy2 * dy);
      }
      V = v;
      cv_pullback<R, ND>(tcv, K.rot, x, o, g, f);
    }
  }
}

// ------------------------------------------------------------------------------------- biases
template <typename R, int ND>
struct WalkerBias {  // umbrella window of this walker (restraint 0) when BIAS_UMB_WALKER
  R k[ND], r0[ND];
};

// Hill sum of well-tempered metadynamics, fixed sigma: bias/gaus.py:236-261.
//   V = sum_i w_i exp(-sum_k (c_k - h_ik)^2 / (2 sigma_k^2)),  g_k = sum_i V_i (c_k - h_ik) / sigma_k^2
// Hill records are (h_0..h_{cd-1}, w) with stride K.hill_stride; shared table: record i at
// hills[i*stride]; per-walker tables: field j of record i of walker w at hills[(i*stride + j)*M + w].
template <typename R, int ND>
NDS_DEV void hill_sum(const Consts<R>& K, int cd, const R* __restrict__ hills, long long n_hills,
                      long long elem_stride, const R (&c)[ND], R& V, R (&g)[ND]) {
  R v = (R)0;
#pragma unroll
  for (int k = 0; k < ND; k++) g[k] = (R)0;
  const int rs = K.hill_stride;
  if (pack_of<R>::value == 1 && cd == 2 && elem_stride == 1) {
    // 16/32-byte records (h0, h1, w, pad): one vector load per hill, warp-uniform address (broadcast)
    if constexpr (is_scalar_f32<R>::value) {
      const float4* __restrict__ h4 = reinterpret_cast<const float4*>(hills);
#pragma unroll 4
      for (long long i = 0; i < n_hills; i++) {
        const float4 h = __ldg(h4 + i);
        const R d0 = c[0] - h.x, d1 = c[1] - h.y;
        const R s0 = d0 * K.mtd_invsig2[0], s1 = d1 * K.mtd_invsig2[1];
        const R ex = (R)0.5 * (d0 * s0 + d1 * s1);
        const R vr = h.z * nds_exp(-ex);
        v += vr;
        g[0] = fmaf(vr, s0, g[0]);
        g[1] = fmaf(vr, s1, g[1]);
      }
    } else if constexpr (pack_of<R>::value == 1) {
      const double2* __restrict__ h2 = reinterpret_cast<const double2*>(hills);
#pragma unroll 2
      for (long long i = 0; i < n_hills; i++) {
        const double2 ha = __ldg(h2 + 2 * i), hb = __ldg(h2 + 2 * i + 1);
        const R d0 = c[0] - ha.x, d1 = c[1] - ha.y;
        const R ex = (d0 * d0) / (R)2.0 * K.mtd_invsig2[0] + (d1 * d1) / (R)2.0 * K.mtd_invsig2[1];
        const R vr = hb.x * nds_exp(-ex);
        v += vr;
        g[0] += vr * (d0 * K.mtd_invsig2[0]);
        g[1] += vr * (d1 * K.mtd_invsig2[1]);
      }
    }
  } else {
    const bool adaptive = K.mtd_adaptive != 0;  // colvardim 1: 1/sigma^2 of each hill is field 2 of its record
    for (long long i = 0; i < n_hills; i++) {
      const R* h = hills + i * rs * elem_stride;
      R ex = (R)0, ds[ND];
#pragma unroll
      for (int k = 0; k < ND; k++)
        if (k < cd) {
          const R dc = c[k] - h[k * elem_stride];
          const R inv = adaptive ? h[2 * elem_stride] : K.mtd_invsig2[k];  // gaus.py:268-275 / :240-248
          ds[k] = dc * inv;
          ex += (dc * dc) / (R)2.0 * inv;
        }
      const R vr = h[K.hill_hidx * elem_stride] * nds_exp(-ex);
      v += vr;
#pragma unroll
      for (int k = 0; k < ND; k++)
        if (k < cd) g[k] += vr * ds[k];
    }
  }
  V = v;
}

// ---------------------------------------------------------------------------- grid-accumulated bias
// mtd_grid: the metadynamics bias is kept as a regular grid of (V, h0 V_0, h1 V_1, h0 h1 V_01) (value, first
// derivatives and the cross derivative, pre-multiplied by the spacings), accumulated hill by hill at deposition
// time (grid_accumulate_kernel) -- the reference's own incremental trick for its FES plot (gaus.py:380-393).  The
// step evaluates V and g = -dV/dc by bicubic Hermite interpolation: 4 vector loads + ~60 FMAs whatever the
// number of hills.  The interpolant is C1 and its gradient is the exact derivative of the interpolated energy, so
// the bias force stays conservative.  Error for a Gaussian of width sigma at spacing h: (h/sigma)^4 * 3/384 of the
// hill height in V (3e-5 at h = sigma/4).
// A point outside [lo, hi] sees no bias (V = 0, g = 0) and is counted in *outside.
template <typename R> struct Real4;
template <> struct Real4<float> { using type = float4; };
template <> struct Real4<double> { using type = double4; };
template <> struct Real4<f2> { using type = float4; };  // never used (no grid bias in packed kernels): keeps shared code well-formed

template <typename R, bool SMEM>
NDS_DEV typename Real4<R>::type grid_load(const R* grid, int idx) {
  using V4 = typename Real4<R>::type;
  if constexpr (SMEM && sizeof(R) == 4) {
    float4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(grid) + (unsigned)idx * 16u;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
  } else if constexpr (SMEM) {
    return reinterpret_cast<const V4*>(grid)[idx];
  } else {
    if constexpr (sizeof(R) == 4) return __ldg(reinterpret_cast<const float4*>(grid) + idx);
    else {
      const double2 a = __ldg(reinterpret_cast<const double2*>(grid) + 2 * idx);
      const double2 b = __ldg(reinterpret_cast<const double2*>(grid) + 2 * idx + 1);
      return double4{a.x, a.y, b.x, b.y};
    }
  }
}

// cubic Hermite basis on [0, 1]: a0 = h00, a1 = h01 (values), b0 = h10, b1 = h11 (slopes), and derivatives
template <typename R>
NDS_DEV void hermite_basis(R t, R (&a)[2], R (&b)[2], R (&da)[2], R (&db)[2]) {
  const R u = (R)1 - t;
  a[0] = ((R)1 + (R)2 * t) * (u * u);
  a[1] = (t * t) * ((R)3 - (R)2 * t);
  b[0] = t * (u * u);
  b[1] = -((t * t) * u);
  da[1] = (R)6 * (t * u);
  da[0] = -da[1];
  db[0] = u * ((R)1 - (R)3 * t);
  db[1] = t * ((R)3 * t - (R)2);
}

// The interpolation in two halves, so that the step kernel can issue the node loads right after the colvar is known
// and evaluate the potential while they are in flight (grid_fetch), then contract (grid_combine).
template <typename R>
struct GridFetch {
  typename Real4<R>::type n[4];  // nodes (j, i), (j + 1, i), (j, i + 1), (j + 1, i + 1); 1-D: nodes i, i + 1
  R t, s;                        // position inside the cell
  int state;                     // -1: not fetched (interpolate in place), 0: outside the box, 1: 1-D, 2: 2-D
};

template <typename R, int ND, bool SMEM>
NDS_DEV void grid_fetch(const Consts<R>& K, int cd, const R* grid, const R (&c)[ND], GridFetch<R>& f,
                        unsigned long long* outside) {
  f.state = 0;
  if constexpr (pack_of<R>::value == 1) {
    const int nx = K.grid_nx, ny = K.grid_ny;
    const R u = (c[0] - K.grid_lo[0]) * K.grid_invh[0];
    R w = (R)0;
    if (cd > 1) w = (c[ND > 1 ? 1 : 0] - K.grid_lo[1]) * K.grid_invh[1];
    // !(a >= b) forms: a NaN colvar counts as outside
    bool out = !(u >= (R)0) || !(u <= (R)(nx - 1));
    if (cd > 1) out = out || !(w >= (R)0) || !(w <= (R)(ny - 1));
    if (out) {
      if (outside) atomicAdd(outside, 1ull);
      return;
    }
    int i = (int)u;
    if (i > nx - 2) i = nx - 2;
    f.t = u - (R)i;
    if (cd == 1) {
      f.n[0] = grid_load<R, SMEM>(grid, i);
      f.n[1] = grid_load<R, SMEM>(grid, i + 1);
      f.state = 1;
      return;
    }
    int j = (int)w;
    if (j > ny - 2) j = ny - 2;
    f.s = w - (R)j;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      f.n[2 * p] = grid_load<R, SMEM>(grid, j * nx + i + p);
      f.n[2 * p + 1] = grid_load<R, SMEM>(grid, (j + 1) * nx + i + p);
    }
    f.state = 2;
  }
}

template <typename R, int ND>
NDS_DEV void grid_combine(const Consts<R>& K, const GridFetch<R>& f, R& V, R (&g)[ND]) {
  V = (R)0;
#pragma unroll
  for (int k = 0; k < ND; k++) g[k] = (R)0;
  if constexpr (pack_of<R>::value == 1) {
    if (f.state <= 0) return;
    R a[2], b[2], da[2], db[2];
    hermite_basis<R>(f.t, a, b, da, db);
    if (f.state == 1) {
      const auto G0 = f.n[0], G1 = f.n[1];
      V = a[0] * G0.x + b[0] * G0.y + a[1] * G1.x + b[1] * G1.y;
      const R dt = da[0] * G0.x + db[0] * G0.y + da[1] * G1.x + db[1] * G1.y;
      g[0] = -(dt * K.grid_invh[0]);
      return;
    }
    R as[2], bs[2], das[2], dbs[2];
    hermite_basis<R>(f.s, as, bs, das, dbs);
    R Vt = (R)0, dVt = (R)0, dVs = (R)0;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const auto G0 = f.n[2 * p];
      const auto G1 = f.n[2 * p + 1];
      // contraction along s at column i + p: value, t-slope, and their s-derivatives
      const R vp = as[0] * G0.x + bs[0] * G0.z + as[1] * G1.x + bs[1] * G1.z;
      const R yp = as[0] * G0.y + bs[0] * G0.w + as[1] * G1.y + bs[1] * G1.w;
      const R vps = das[0] * G0.x + dbs[0] * G0.z + das[1] * G1.x + dbs[1] * G1.z;
      const R yps = das[0] * G0.y + dbs[0] * G0.w + das[1] * G1.y + dbs[1] * G1.w;
      Vt += a[p] * vp + b[p] * yp;
      dVt += da[p] * vp + db[p] * yp;
      dVs += a[p] * vps + b[p] * yps;
    }
    V = Vt;
    g[0] = -(dVt * K.grid_invh[0]);
    if constexpr (ND > 1) g[1] = -(dVs * K.grid_invh[1]);
  }
}

template <typename R, int ND, bool SMEM>
NDS_DEV void grid_interp(const Consts<R>& K, int cd, const R* grid, const R (&c)[ND], R& V, R (&g)[ND],
                         unsigned long long* outside) {
  GridFetch<R> f;
  grid_fetch<R, ND, SMEM>(K, cd, grid, c, f, outside);
  grid_combine<R, ND>(K, f, V, g);
}

template <typename R>
struct GridView {  // accumulated bias grid seen by one evaluation (mtd_grid)
  const R* grid;   // 4 reals per point: shared-memory copy (smem = 1) or the global array
  int smem;
  unsigned long long* outside;
};

struct HillView {
  const void* base;      // hill table (engine precision)
  long long n;           // hills visible to this evaluation
  long long elem_stride; // 1 for the shared table, M for per-walker tables
};

// Sum of fix.compute(x0, col0) over the bias list (md.py:182-188): energy Vb, force fb[ND] in x space,
// and the metadynamics part Vm alone (what Gaussian.compute would store in VR on a no-arg call,
// gaus.py:292-293).  Order of accumulation: harmonic restraints, metadynamics, PE bias.
template <typename R, int ND>
NDS_DEV void bias_eval(const Consts<R>& K, int bias_mask, int cv, int cd, const R (&x)[ND],
                       const CvOut<R, ND>& o, const WalkerBias<R, ND>& wb, const R* hills,
                       long long n_hills, long long elem_stride, R pe, const R (&fpot)[ND], R& Vb,
                       R (&fb)[ND], R& Vm, const GridView<R> gv = GridView<R>{nullptr, 0, nullptr},
                       const GridFetch<R>* pre = nullptr) {
  Vb = (R)0;
  Vm = (R)0;
#pragma unroll
  for (int d = 0; d < ND; d++) fb[d] = (R)0;
  if (bias_mask & BIAS_UMB) {  // bias/umb.py:23-40
#pragma unroll
    for (int u = 0; u < NDS_MAX_UMB; u++) {
      if (u >= K.n_umb) break;
      R g[ND], e = (R)0;
#pragma unroll
      for (int k = 0; k < ND; k++) {
        g[k] = (R)0;
        if (k < cd) {
          const bool pw = (bias_mask & BIAS_UMB_WALKER) && u == 0;
          const R kk = pw ? wb.k[k] : K.umb_k[u][k];
          const R r0 = pw ? wb.r0[k] : K.umb_r0[u][k];
          const R dR = o.c[k] - r0;
          g[k] = -kk * dR;
          e += kk * (dR * dR);
        }
      }
      R f[ND];
      cv_pullback<R, ND>(cv, K.rot, x, o, g, f);
#pragma unroll
      for (int d = 0; d < ND; d++) fb[d] += f[d];
      Vb += (R)0.5 * e;
    }
  }
  if (bias_mask & BIAS_MTD) {
    R g[ND], f[ND];
    if (bias_mask & BIAS_MTD_GRID) {
      if (pre) grid_combine<R, ND>(K, *pre, Vm, g);  // nodes were requested before the potential evaluation
      else if (gv.smem) grid_interp<R, ND, true>(K, cd, gv.grid, o.c, Vm, g, gv.outside);
      else grid_interp<R, ND, false>(K, cd, gv.grid, o.c, Vm, g, gv.outside);
    } else {
      hill_sum<R, ND>(K, cd, hills, n_hills, elem_stride, o.c, Vm, g);
    }
    if (n_hills > 0 || (bias_mask & BIAS_MTD_GRID)) {
      cv_pullback<R, ND>(cv, K.rot, x, o, g, f);
#pragma unroll
      for (int d = 0; d < ND; d++) fb[d] += f[d];
    }
    Vb += Vm;
  }
  if (bias_mask & BIAS_PE) {  // bias/pe.py:30-51
    if (pe < K.pe_thred) {
#pragma unroll
      for (int d = 0; d < ND; d++) fb[d] += -fpot[d];
      Vb += K.pe_thred - pe;
    }
  }
}

// committor.py:118-142: inclusive boxes in colvar space; the last matching basin wins
// ZT: the engine is the minimiser (the only case in which temperature can be 0); NB: compile-time basin count
// or DYN.  T is the scalar type of one walker; the constant block may be the packed one (KT = Consts<f2>),
// whose entries carry the same value in both lanes.
NDS_DEV float lane0(float a) { return a; }
NDS_DEV double lane0(double a) { return a; }
NDS_DEV float lane0(f2 a) { return a.v.x; }

template <typename T, int ND, bool ZT = true, int NB = DYN, typename KT>
NDS_DEV int commit_test(const KT& K, int cd, const T (&c)[ND], T pe) {
  // at temperature 0 a basin additionally needs pe < emin (committor.py:128-131)
  if (ZT && K.zero_temperature && !(pe < lane0(K.emin))) return -1;
  const int n_basins = NB >= 0 ? NB : K.n_basins;
  // branch-free: every comparison is evaluated (constant-bank operands) and combined with bit ops, so a
  // warp never diverges on the test.  "outside" is (c < lo || c > hi) like the reference, so a NaN
  // colvar counts as inside, exactly as there.
  auto in_box = [&](int b) {
    unsigned out = 0u;
#pragma unroll
    for (int k = 0; k < ND; k++)
      if (k < cd) out |= (unsigned)(c[k] < lane0(K.crit[b][k][0])) | (unsigned)(c[k] > lane0(K.crit[b][k][1]));
    return out == 0u;
  };
  int ib = in_box(0) ? 0 : -1;
  if (n_basins > 1) ib = in_box(1) ? 1 : ib;  // uniform
  if (n_basins > 2) {  // rare: skipped by one uniform branch
#pragma unroll 1
    for (int b = 2; b < n_basins; b++) ib = in_box(b) ? b : ib;
  }
  return ib;
}

}  // namespace nds
