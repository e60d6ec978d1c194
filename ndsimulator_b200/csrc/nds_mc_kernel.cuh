// nds_mc_kernel.cuh -- batched Metropolis Monte Carlo (reference engines/mc.py:42-101, "next" row 4 of
// SURVEY.md section 8f): one thread per walker, K trial moves per launch, energies only.
//
// Per move: propose (one random coordinate or all of them, uniform in +-mc_bounds), wrap into the
// periodic box like numpy's %, evaluate potential + bias energy at the trial point, accept with
// min(1, exp(-dE/kBT)).  Uniforms come from Philox4x32-10 (domain 2: key1 ^ 2, counter = move index,
// global walker id) or from a caller-supplied table (parity tests).  The reference draws its proposals
// from the unseeded global numpy RNG, so only statistical parity exists for this engine.
#pragma once
#include "nds_kernels.cuh"

namespace nds {

// np.argmin(np.abs(Elist - e)), Elist = arange(101)/100*(Emax-Emin)+Emin (wang_landau.py:12-20); the
// nearest grid point is found among the neighbours of the analytic position, ties go to the lower index
template <typename R>
NDS_DEV int wl_bin(R e) {
  const R Emin = (R)-1.5, Emax = (R)0.2;
  const R t = (e - Emin) / (Emax - Emin) * (R)(NDS_WL_GRID - 1);
  int i0 = (int)floor((double)t) - 1;
  if (i0 < 0) i0 = 0;
  if (i0 > NDS_WL_GRID - 1) i0 = NDS_WL_GRID - 1;
  int best = i0;
  R bd = (R)0;
  for (int i = i0; i <= i0 + 3 && i < NDS_WL_GRID; i++) {
    const R El = (R)i / (R)(NDS_WL_GRID - 1.0) * (Emax - Emin) + Emin;
    R d = El - e;
    d = d < (R)0 ? -d : d;
    if (i == i0 || d < bd) { bd = d; best = i; }
  }
  return best;
}

NDS_DEV double u01(uint32_t w, double) { return ((double)w + 0.5) * 2.3283064365386963e-10; }
NDS_DEV float u01(uint32_t w, float) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f; }

template <typename R, int ND>
__global__ void __launch_bounds__(NDS_BLOCK)
mc_block(const __grid_constant__ Consts<R> K, const __grid_constant__ BlockArgs<R> A, long long* accepted,
         R* wl_hE /* [NDS_WL_GRID][M] per-walker histograms, Wang-Landau only */,
         double* mc_time /* [M] NDRun.time of each walker: advances on ACCEPTED moves only (ndrun.py:234-237) */,
         long long* mc_dep /* [M] Gaussian._dep_count of each walker's own table; [M] = overflow flag */,
         R* hills_rw /* per-walker hill tables, written by the in-kernel deposition */) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long M = A.M;
  if (w >= M) return;
  const int cd = K.cd, bias_mask = K.bias_mask;
  const bool ext = K.noise == NDS_NOISE_EXTERNAL;
  const int nu = K.mc_all ? ND + 1 : 3;
  const unsigned long long gid = (unsigned long long)(A.walker_offset + w);

  R x[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) x[d] = A.x[d * M + w];
  WalkerBias<R, ND> wb;
#pragma unroll
  for (int k = 0; k < ND; k++) { wb.k[k] = (R)0; wb.r0[k] = (R)0; }
  if (bias_mask & BIAS_UMB_WALKER) {
#pragma unroll
    for (int k = 0; k < ND; k++)
      if (k < cd) { wb.k[k] = A.umb_k[k * M + w]; wb.r0[k] = A.umb_r0[k * M + w]; }
  }
  // Metadynamics under MC: NDRun.time only advances on accepted moves (ndrun.py:234-237), so fix.update
  // (gaus.py:130-146) follows a clock of the walker's own: every walker keeps its own table (mtd_shared = 0,
  // checked by nds_create), deposits inside this kernel and sees dep hills.
  const bool mtd = (bias_mask & BIAS_MTD) != 0;
  const long long hstride = M;
  const R* hills = (mtd && hills_rw) ? hills_rw + w : nullptr;
  double tw = mtd ? mc_time[w] : 0.0;
  long long dep = mtd ? mc_dep[w] : 0;
  R vr = mtd ? A.vr[w] : (R)0;
  // MC.begin (mc.py:34-40): energies at the current point
  Eval<R, ND> cur;
  eval_all<R, ND, true>(K, K.pot, K.cv, K.tcv, cd, bias_mask, x, wb, hills, dep, hstride, cur);
  long long acc = accepted[w];
  const long long acc0 = acc;

  bool deposited = false, overflowed = false;
  R ke = A.thermo[2 * M + w], tot = A.thermo[4 * M + w];  // atoms.ke / atoms.totale follow the last event
  for (long long s = A.step0; s < A.step0 + A.nsteps; s++) {
    const bool dep_due = mtd && (long long)floor(tw / K.mtd_dep_freq64) + 1 > dep;
    if (dep_due && dep >= K.mtd_capacity && !overflowed) {
      // no room for the hill the reference would deposit here: reported by run_mc ("hill table full")
      atomicMax(reinterpret_cast<unsigned long long*>(mc_dep + M), (unsigned long long)(w + 1));
      overflowed = true;
    }
    if (dep_due && dep < K.mtd_capacity) {
      // Gaussian.update -> _deposite (gaus.py:130-213) at atoms.colv with the VR of the last refresh, then the
      // refresh itself (pe, forces, fixf, biase with the new hill, VR, totale), gaus.py:138-146
      const R h = dep == 0 ? K.mtd_w : K.mtd_w * nds_exp(-vr / K.mtd_kBdT);
      const int rs = K.hill_stride;
      R* rec = hills_rw + w;
#pragma unroll
      for (int q = 0; q < ND; q++)
        if (q < cd) rec[(dep * rs + q) * M] = cur.cv.c[q];
      rec[(dep * rs + K.hill_hidx) * M] = h;
      if (K.mtd_adaptive) {  // gaus.py:162-165,189-192 (colvardim 1)
        R g[ND], row[ND];
#pragma unroll
        for (int d = 0; d < ND; d++) g[d] = d == 0 ? (R)1 : (R)0;
        cv_pullback<R, ND>(K.cv, K.rot, x, cur.cv, g, row);
        R jj = (R)0;
#pragma unroll
        for (int d = 0; d < ND; d++) jj += row[d] * row[d];
        const R s2 = dep == 0 ? jj : jj * K.mtd_sig2[0];
        rec[(dep * rs + 2) * M] = (R)1.0 / s2;
        rec[(dep * rs + 3) * M] = s2;
      }
      dep += 1;
      eval_all<R, ND, true>(K, K.pot, K.cv, K.tcv, cd, bias_mask, x, wb, hills, dep, hstride, cur);
      vr = cur.vm;
      deposited = true;
#pragma unroll
      for (int d = 0; d < ND; d++) { A.f[d * M + w] = cur.fp[d]; A.fixf[d * M + w] = cur.fb[d]; }
      tot = ke + cur.pe + cur.be;  // gaus.py:146
    }
    R u[ND + 1 > 3 ? ND + 1 : 3];  // "single": pick, displacement, acceptance; "all": ND displacements + acceptance
    if (ext) {
      const double* row = A.noise + ((s - A.noise_first) * nu) * M + w;
      for (int j = 0; j < nu; j++) u[j] = (R)row[(long long)j * M];
    } else {
      const u32x4 r0 = philox4x32_10(u32x4{(uint32_t)(2 * s), (uint32_t)((2 * s) >> 32), (uint32_t)gid, (uint32_t)(gid >> 32)}, K.rk_mc);
      u[0] = u01(r0.x, (R)0); u[1] = u01(r0.y, (R)0); u[2] = u01(r0.z, (R)0);
      if (ND + 1 > 3) {
        if (K.mc_all) {
          const u32x4 r1 = philox4x32_10(u32x4{(uint32_t)(2 * s + 1), (uint32_t)((2 * s + 1) >> 32), (uint32_t)gid, (uint32_t)(gid >> 32)}, K.rk_mc);
          const uint32_t words[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < ND + 1; j++) u[j] = u01(words[j], (R)0);
        }
      }
    }
    // propose, mc.py:47-62
    R xn[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) xn[d] = x[d];
    int k = 0;
    if (!K.mc_all) {
      int dim = (int)floor((double)u[k++] * ND);
      if (dim >= ND) dim = ND - 1;
      const R du = (R)2.0 * u[k++] - (R)1.0;
#pragma unroll
      for (int d = 0; d < ND; d++)
        if (d == dim) xn[d] += du * K.mc_bounds[d];
    } else {
#pragma unroll
      for (int d = 0; d < ND; d++) xn[d] += ((R)2.0 * u[k++] - (R)1.0) * K.mc_bounds[d];
    }
    if (K.mc_periodic) {  // numpy %: the result takes the sign of the divisor
#pragma unroll
      for (int d = 0; d < ND; d++) {
        R r = fmod(xn[d], K.mc_gb[d]);
        if (r != (R)0 && ((r < (R)0) != (K.mc_gb[d] < (R)0))) r += K.mc_gb[d];
        xn[d] = r;
      }
    }
    Eval<R, ND> tr;
    eval_all<R, ND, true>(K, K.pot, K.cv, K.tcv, cd, bias_mask, xn, wb, hills, dep, hstride, tr);
    const R pe = cur.pe + cur.be, pe_new = tr.pe + tr.be;  // mc.py:64-76
    R pb = (pe_new > pe) ? nds_exp(-(pe_new - pe) / K.kBT) : (R)1.0;
    int eid1 = 0;
    if (K.method == NDS_METHOD_WL) {  // wang_landau.py:113-126, one basin
      if (pe_new > (R)0.2) pb = (R)-1.0;  // pb = 0: never accepted
      else {
        eid1 = wl_bin<R>(pe_new);
        const R h1 = wl_hE[eid1 * M + w];
        if (h1 == (R)0) pb = (R)1.0;
        else { const R r = wl_hE[wl_bin<R>(pe) * M + w] / h1; pb = r < (R)1.0 ? r : (R)1.0; }
      }
    }
    if (u[k] <= pb) {  // mc.py:83-96
      if (K.method == NDS_METHOD_WL) {  // scaleup (wang_landau.py:51-55), f = 2
        const R h1 = wl_hE[eid1 * M + w];
        wl_hE[eid1 * M + w] = (h1 == (R)0) ? (R)1.0 : h1 * (R)2.0;
      }
#pragma unroll
      for (int d = 0; d < ND; d++) x[d] = xn[d];
      cur = tr;
      acc += 1;
      tw += K.dt64;  // ndrun.py:235-237: time advances on accepted moves only
      ke = (R)0; tot = cur.pe;  // mc.py:93-95
    }
  }
  if (mtd) { mc_time[w] = tw; mc_dep[w] = dep; if (deposited) A.vr[w] = vr; }

#pragma unroll
  for (int d = 0; d < ND; d++) A.x[d * M + w] = x[d];
#pragma unroll
  for (int q = 0; q < ND; q++)
    if (q < cd) A.colv[q * M + w] = cur.cv.c[q];
  A.thermo[1 * M + w] = cur.pe;
  A.thermo[3 * M + w] = cur.be;
  A.thermo[2 * M + w] = ke;
  A.thermo[4 * M + w] = tot;
  if (K.method == NDS_METHOD_WL && acc != acc0) {  // np.copyto(atoms.forces, forces_new), wang_landau.py:152
#pragma unroll
    for (int d = 0; d < ND; d++) A.f[d * M + w] = cur.fp[d];
  }
  if (acc != acc0) accepted[w] = acc;
}

template <typename R>
cudaError_t launch_mc_block(int nd, const Consts<R>& K, const BlockArgs<R>& A, long long* accepted,
                            R* wl_hE, double* mc_time, long long* mc_dep, R* hills_rw, cudaStream_t st);

}  // namespace nds
