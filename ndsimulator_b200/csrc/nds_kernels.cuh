// nds_kernels.cuh -- the fused multi-walker step kernel and its small companions.
//
// md_block<Spec>: one thread per walker, walker state (x, v, total force) register-resident for the
// whole launch; each launch advances every live walker by `nsteps` iterations of the reference loop
//   NDRun.run body (ndrun.py:227-251) -> MD.update (engines/md.py:112-231)
//                                     -> Committor.update (engines/committor.py:82-116)
// including potential / colvar / bias evaluation, the Langevin / NVE / rescale update, Philox noise,
// the committor box test and frame recording.  HBM is touched only at launch entry/exit (SoA state,
// coalesced) and for the hill table / frames.
#pragma once
#include "nds_common.cuh"
#include "nds_philox.cuh"
#include "nds_physics.cuh"

namespace nds {

constexpr int NDS_BLOCK = 128;
constexpr int NDS_BLOCK_WIDE = 512;  // md_block_wide (grid-accumulated bias: one CTA per SM holds the grid)

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk; SASS: UBLKCP + SYNCS), shared by mtd_block and the grid loader
NDS_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

NDS_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
NDS_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NDS_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D TMA: global -> shared bulk copy that completes on an mbarrier (SASS: UBLKCP)
NDS_DEV void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Whole bias grid -> dynamic shared memory of this CTA (mtd_grid): thread 0 issues the bulk copies, everybody waits on
// the mbarrier.  bytes is a multiple of 16 and below the 2^20 - 1 transaction-count limit of one mbarrier phase.
NDS_DEV void grid_to_smem(unsigned char* smem, const void* grid, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    constexpr uint32_t CHUNK = 32768;
    for (uint32_t off = 0; off < bytes; off += CHUNK)
      tma_load_1d(smem + off, (const unsigned char*)grid + off, bytes - off < CHUNK ? bytes - off : CHUNK, bar);
  }
  mbar_wait(bar, 0);
}

// NDS_INT_NVE, NDS_INT_RESCALE and NDS_INT_MINIMIZE draw no noise
__host__ __device__ constexpr int normals_per_step(int integ, int nd) {
  return integ == NDS_INT_LANGEVIN ? nd : (integ == NDS_INT_LANGEVIN2 ? 2 * nd : 0);
}
__host__ __device__ constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
// Steps per noise group and Philox calls per group (one call = NDS_NPC = 6 normals).  G*NN is a multiple
// of 6 except for NN = 5 (5-D Langevin), where one call per step wastes a normal rather than unrolling six
// 5-D steps.  NN: 1 -> G 6, 2 -> 3, 4 -> 3, 5 -> 1, 10 -> 3.
__host__ __device__ constexpr int noise_group(int nn) {
  return nn == 0 ? 1 : (nn == 5 ? 1 : NDS_NPC / gcd_(nn, NDS_NPC));
}
__host__ __device__ constexpr int noise_calls(int nn) {
  return nn == 0 ? 1 : (noise_group(nn) * nn + NDS_NPC - 1) / NDS_NPC;
}

// frame events (io/dump.py:76-102 + the lagged thermo row): a frame is recorded after the
// step that makes NDRun.step == s when s is a dump step, or when s is the last stat sample strictly
// before the next dump step (that sample is what thermo.dat prints beside the next dump row).
__host__ __device__ __forceinline__ bool frame_event(long long s, long long dump_freq, long long stat_freq) {
  if (dump_freq <= 0) return false;
  if (s % dump_freq == 0) return true;
  if (stat_freq > 0 && s % stat_freq == 0) {
    const long long next_dump = (s / dump_freq + 1) * dump_freq;
    return next_dump - s <= stat_freq;
  }
  return false;
}

// Everything evaluated at one configuration.
template <typename R, int ND>
struct Eval {
  CvOut<R, ND> cv;
  R pe, be, vm;
  R fp[ND], fb[ND];
};

// PREFETCH (kernels whose bias mask is the grid-accumulated metadynamics bias at compile time): the four grid nodes
// are requested as soon as the colvar is known, the potential is evaluated while they travel from L2 / shared memory
template <typename R, int ND, bool ENERGY, bool PREFETCH = false>
NDS_DEV void eval_all(const Consts<R>& K, int pot, int cv, int tcv, int cd, int bias_mask,
                      const R (&x)[ND], const WalkerBias<R, ND>& wb, const R* hills, long long n_hills,
                      long long elem_stride, Eval<R, ND>& e, const GridView<R> gv = GridView<R>{nullptr, 0, nullptr}) {
  cv_eval<R, ND>(K, cv, K.rot, x, e.cv);
  if constexpr (PREFETCH) {
    GridFetch<R> gf;
    if (gv.smem) grid_fetch<R, ND, true>(K, cd, gv.grid, e.cv.c, gf, gv.outside);
    else grid_fetch<R, ND, false>(K, cd, gv.grid, e.cv.c, gf, gv.outside);
    if (ENERGY || (bias_mask & BIAS_PE)) pot_eval<R, ND, true>(K, pot, tcv, x, e.pe, e.fp);
    else pot_eval<R, ND, false>(K, pot, tcv, x, e.pe, e.fp);
    bias_eval<R, ND>(K, bias_mask, cv, cd, x, e.cv, wb, hills, n_hills, elem_stride, e.pe, e.fp, e.be, e.fb, e.vm, gv, &gf);
    return;
  }
  // PE bias needs the potential energy even when the caller does not
  if (ENERGY || (bias_mask & BIAS_PE)) pot_eval<R, ND, true>(K, pot, tcv, x, e.pe, e.fp);
  else pot_eval<R, ND, false>(K, pot, tcv, x, e.pe, e.fp);
  if (bias_mask != 0)
    bias_eval<R, ND>(K, bias_mask, cv, cd, x, e.cv, wb, hills, n_hills, elem_stride, e.pe, e.fp, e.be,
                     e.fb, e.vm, gv);
  else {
    e.be = (R)0; e.vm = (R)0;
#pragma unroll
    for (int d = 0; d < ND; d++) e.fb[d] = (R)0;
  }
}

// one frame row of walker w (layout documented at nds_drain_frames in include/ndsb200.h).  Rows are written in the
// engine's own real type: fp32 engines record float rows (half the HBM bytes of fp64 rows, no F2F.F64 per field),
// the packed kernel stores the pair (walker 2w, 2w+1) with one 8-byte store per field; every store of a warp
// covers one contiguous 128 / 256-byte run of the [frame][field][walker] ring.
template <typename R, int ND>
NDS_DEV void write_frame(const Consts<R>& K, const BlockArgs<R>& A, long long w, long long frame,
                         long long s, double time, int cd, int bias_mask, const R (&x)[ND],
                         const R (&v)[ND], const Eval<R, ND>& o, bool begin_mode) {
  if (frame >= A.frame_cap) return;
  const long long st = K.dump_walkers;
  R* row = A.frames + (frame * A.frame_width) * st + w;
  int r = 0;
  row[(r++) * st] = (R)(double)s;
  row[(r++) * st] = (R)time;
#pragma unroll
  for (int d = 0; d < ND; d++) row[(r++) * st] = x[d];
#pragma unroll
  for (int d = 0; d < ND; d++) row[(r++) * st] = v[d];
#pragma unroll
  for (int d = 0; d < ND; d++) row[(r++) * st] = o.fp[d];
#pragma unroll
  for (int k = 0; k < ND; k++)
    if (k < cd) row[(r++) * st] = o.cv.c[k];
  R v2 = (R)0;
#pragma unroll
  for (int d = 0; d < ND; d++) v2 += v[d] * v[d];
  const R ke = v2 * K.m / (R)2.0;
  row[(r++) * st] = ke / (R)ND / K.kB * (R)2.0;
  row[(r++) * st] = o.pe;
  row[(r++) * st] = ke;
  row[(r++) * st] = o.be;
  row[(r++) * st] = begin_mode ? (ke + o.pe) : (o.pe + ke + o.be);
  row[(r++) * st] = (bias_mask & BIAS_MTD) ? A.vr[w] : (R)0;
  if constexpr (pack_of<R>::value == 1) {
    if (A.vr_umb)  // one more field per harmonic restraint: its VR as of the last refresh
      for (int u = 0; u < K.n_umb; u++) row[(r++) * st] = A.vr_umb[u * A.M + w];
  }
}

// in-kernel colvar histogram at dump cadence (io/plot.py:562-572 accumulated over the run): shared-memory bins
NDS_DEV void hist_count(unsigned int* bins, float c0, float c1, float lo0, float inv0, int nx, float lo1, float inv1, int ny) {
  // floor to int, then ONE unsigned compare per axis rejects both sides of the box (negative indices wrap high)
  const int i = __float2int_rd((c0 - lo0) * inv0), j = __float2int_rd((c1 - lo1) * inv1);
  const int bin = ((unsigned)i < (unsigned)nx && (unsigned)j < (unsigned)ny) ? j * nx + i : -1;
  // (one atomic per distinct bin of the warp through __match_any_sync was measured SLOWER: 3.00 vs 2.92 ms)
  if (bin >= 0) atomicAdd(&bins[bin], 1u);
}

// resident CTAs per SM asked of ptxas (register cap): the packed committor kernel is held to 64 registers
template <class S>
constexpr int md_block_min_ctas() {
  return (pack_of<typename S::R>::value == 2 && S::METHOD == NDS_METHOD_COMMITTOR) ? 8 : 1;
}

template <class S>
NDS_DEV void md_body(const Consts<typename S::R>& K, const BlockArgs<typename S::R>& A) {
  using R = typename S::R;
  constexpr int ND = S::ND;
  constexpr int INTEG = S::INTEG;  // always compile-time
  constexpr int PACK = pack_of<R>::value;  // walkers per thread (2 for the packed-FP32 real type f2)
  static_assert(PACK == 1 || ((S::METHOD == NDS_METHOD_MD || S::METHOD == NDS_METHOD_COMMITTOR) &&
                              S::NOISE == NDS_NOISE_PHILOX && S::DUMP >= 0 && INTEG != NDS_INT_MINIMIZE &&
                              (S::DUMP == 0 || S::METHOD == NDS_METHOD_MD) &&
                              S::BIAS >= 0 && (S::BIAS & (BIAS_PE | BIAS_MTD)) == 0),
                "packed kernels: md / committor, Philox noise, frames only for md, no PE / metadynamics bias");
  // S::DUMP: bit 0 = frames (dump rows of the first dump_walkers walkers), bit 1 = colvar histogram of ALL walkers
  // at dump cadence; DYN: both decided by the constant block
  const bool frames_on = S::DUMP >= 0 ? (S::DUMP & 1) != 0 : (K.dump_freq > 0 && K.dump_walkers > 0);
  const bool hist_on = S::DUMP >= 0 ? (S::DUMP & 2) != 0 : (K.hist_nx > 0);
  constexpr int NN = normals_per_step(INTEG, ND);
  constexpr int G = noise_group(NN);
  constexpr int CG = noise_calls(NN);
  // fp32 first-order Langevin: the noise enters only as c2 * xi (md.py:170,196), so the Box-Muller radius is
  // generated already multiplied by c2 and xi[] below holds c2 * xi
  constexpr bool PRESCALED = is_f32<R>::value && INTEG == NDS_INT_LANGEVIN;
  constexpr bool PF = PACK == 1 && S::BIAS >= 0 && (S::BIAS & BIAS_MTD_GRID) != 0;  // grid nodes prefetched in eval_all

  const long long w = A.slot0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;  // slot
  const long long M = A.M;
  // mtd_grid: every CTA first copies the whole bias grid into its shared memory (1-D TMA bulk copies)
  GridView<R> gv{A.grid, 0, A.grid_outside};
  extern __shared__ __align__(128) unsigned char md_smem[];
  unsigned int* hist_bins = nullptr;
  if constexpr (PACK == 1 && (S::BIAS < 0 || (S::BIAS & BIAS_MTD_GRID) != 0)) {
    if (A.grid_smem) {
      __shared__ uint64_t grid_bar;
      grid_to_smem(md_smem, A.grid, (uint32_t)(4u * sizeof(R)) * (uint32_t)(K.grid_nx * K.grid_ny), &grid_bar);
      gv.grid = reinterpret_cast<const R*>(md_smem);
      gv.smem = 1;
    }
  }
  if constexpr (S::DUMP < 0 || (S::DUMP & 2) != 0) {
    if (hist_on) {  // bins behind the grid copy (if any); cleared before any thread leaves
      hist_bins = reinterpret_cast<unsigned int*>(md_smem + A.hist_smem_off);
      for (int b = threadIdx.x; b < K.hist_nx * K.hist_ny; b += blockDim.x) hist_bins[b] = 0u;
      __syncthreads();
    }
  }
  if (w >= (A.n_active > 0 ? A.n_active : M)) return;
  // local walker index of this slot (Philox stream, noise row); a packed thread owns slots 2w and 2w+1
  const long long lw = A.ids ? A.ids[w * PACK] : w * PACK;
  const long long lw_b = PACK == 2 ? (A.ids ? A.ids[w * PACK + 1] : lw + 1) : lw;

  const int pot = S::POT >= 0 ? S::POT : K.pot;
  const int cv = S::CV >= 0 ? S::CV : K.cv;
  const int tcv = S::TCV >= 0 ? S::TCV : K.tcv;
  const int cd = S::CV >= 0 ? cv_dim(S::CV, ND) : K.cd;
  const int bias_mask = S::BIAS >= 0 ? S::BIAS : K.bias_mask;
  const bool committor = S::METHOD >= 0 ? (S::METHOD == NDS_METHOD_COMMITTOR) : (K.method == NDS_METHOD_COMMITTOR);
  const bool ext_noise = S::NOISE >= 0 ? (S::NOISE == NDS_NOISE_EXTERNAL) : (K.noise == NDS_NOISE_EXTERNAL);
  const bool dumping = frames_on && w < K.dump_walkers;

  // committor: lanes that stopped in an earlier launch (packed: a0 / a1 for slots 2w, 2w+1)
  bool a0 = true, a1 = true;
  if constexpr (PACK == 1) {
    if (committor && !A.alive[w]) return;
  } else {
    if (committor) {
      a0 = A.alive[2 * w] != 0; a1 = A.alive[2 * w + 1] != 0;
      if (!a0 && !a1) return;
    }
  }
  const bool was0 = a0, was1 = a1;

  // ---- load walker state (coalesced SoA)
  R x[ND], v[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) { x[d] = A.x[d * M + w]; v[d] = A.v[d * M + w]; }
  WalkerBias<R, ND> wb;
#pragma unroll
  for (int k = 0; k < ND; k++) { wb.k[k] = (R)0; wb.r0[k] = (R)0; }
  if (bias_mask & BIAS_UMB_WALKER) {
#pragma unroll
    for (int k = 0; k < ND; k++)
      if (k < cd) { wb.k[k] = A.umb_k[k * M + w]; wb.r0[k] = A.umb_r0[k * M + w]; }
  }
  const unsigned long long gid = (unsigned long long)(A.walker_offset + lw);
  const unsigned long long gid_b = (unsigned long long)(A.walker_offset + lw_b);
  const long long hstride = K.mtd_shared ? 1 : M;
  const R* hills = K.mtd_shared ? A.hills : (A.hills ? A.hills + w : nullptr);

  if (A.need_init_v) {  // Modify.set_velocity(T=initial_temperature), engines/modify.py:44-54
    static_assert(ND <= NDS_NPC, "one Philox call covers the initial velocities");
    R z[NDS_NPC];
    normals6_pair<R>(K.rk_init, 0ull, gid, gid_b, z);
#pragma unroll
    for (int d = 0; d < ND; d++) v[d] = K.vel_scale * z[d];
  }

  // ---- forces at the current positions: MD.begin (md.py:101-110) / refresh after a deposition
  //      (gaus.py:138-146).  Identical to the values the previous launch ended with.
  Eval<R, ND> e;
  eval_all<R, ND, INTEG == NDS_INT_MINIMIZE>(K, pot, cv, tcv, cd, bias_mask, x, wb, hills, A.n_hills, hstride, e, gv);
  if (A.refresh_vr && (bias_mask & BIAS_MTD)) A.vr[w] = e.vm;  // Gaussian.VR, gaus.py:292-293
  if constexpr (PACK == 1 && (S::BIAS < 0 || ((S::BIAS & BIAS_UMB) != 0 && (S::BIAS & BIAS_MTD) != 0))) {
    // Harmonic.VR: the no-argument fix.compute() of Gaussian.update / MD.begin refreshes it too (umb.py:32-33)
    if (A.refresh_vr && A.vr_umb) {
#pragma unroll
      for (int u = 0; u < NDS_MAX_UMB; u++) {
        if (u >= K.n_umb) break;
        R en = (R)0;
#pragma unroll
        for (int k = 0; k < ND; k++)
          if (k < cd) {
            const bool pw = (bias_mask & BIAS_UMB_WALKER) && u == 0;
            const R kk = pw ? wb.k[k] : K.umb_k[u][k];
            const R dR = e.cv.c[k] - (pw ? wb.r0[k] : K.umb_r0[u][k]);
            en += kk * (dR * dR);
          }
        A.vr_umb[u * M + w] = (R)0.5 * en;
      }
    }
  }
  R f[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) f[d] = bias_mask ? e.fp[d] + e.fb[d] : e.fp[d];  // md.py:133

  // Loop control in 32-bit: `left` steps remain in this launch; NDRun.step is s_end - left (64-bit only
  // where it is really needed: rare paths).  Noise groups are counted separately.
  const long long s_end = A.step0 + A.nsteps;
  long long grp = A.step0 / G;
  int first = (int)(A.step0 - grp * G);
  int left = (int)A.nsteps;
  // committor.py:100: (s - 1) >= diffusion_time with s = s_end - left  <=>  left <= commit_left
  const long long thr_ = s_end - 1 - K.diffusion_time;
  const int commit_left = thr_ < -1 ? -1 : (thr_ > 0x7fffffffLL ? 0x7fffffff : (int)thr_);
  double time = A.time0;
  long long rescale_count = A.rescale_count;
  long long frame = A.frame_base;
  // frame_event(s) evaluated incrementally: rd = s % dump_freq, rs = s % stat_freq are carried along instead
  // of two 64-bit divisions per step
  const bool any_frames = frames_on || hist_on;
  // 32-bit counters (the frequencies are far below 2^31): 64-bit compares and selects cost two issue slots each
  const int dfreq = (int)K.dump_freq, sfreq = (int)K.stat_freq;
  int rd = 0, rs = 0;
  if (any_frames) {
    rd = (int)(A.step0 % K.dump_freq);
    rs = K.stat_freq > 0 ? (int)(A.step0 % K.stat_freq) : 0;
  }
  int to_hist = any_frames ? dfreq - rd : 0x7fffffff;  // executed steps until the next dump step (hist-only kernels)
  bool alive = true;
  int basin = -1;
  long long cv_taken = 0;  // Stat.colv samples written by this launch (BIAS_CVHIST)

  long long exit_s = 0, exit_b = 0;
  int basin_b = -1;
  bool min_stop = false;
  // packed committor: a stopped lane keeps integrating (its results are ignored); its state at the exit
  // step is parked here and restored at launch exit
  R xf[PACK == 2 ? ND : 1], vf[PACK == 2 ? ND : 1];
  if constexpr (PACK == 2) {
#pragma unroll
    for (int d = 0; d < ND; d++) { xf[d] = x[d]; vf[d] = v[d]; }
  }
  while (left > 0) {
    // the step counter stays warp-uniform: a stopped walker idles its lane until the whole warp is done
    if (committor && !__any_sync(0xffffffffu, PACK == 1 ? alive : (a0 || a1))) break;
    R zb[G * NN > 0 ? G * NN : 1];
    // several Philox calls per group (2nd-order Langevin): draw them step by step, right before the step that needs
    // them -- at most NN + NDS_NPC - 1 normals are live instead of G * NN (C4: 90 -> 64 registers, +3 %)
    constexpr bool LAZY = PACK == 1 && NN > NDS_NPC / 2 && G > 1;
    if (NN > 0 && !ext_noise && !LAZY) {
#pragma unroll
      for (int c = 0; c < CG; c++) {
        R z6[NDS_NPC];
        if constexpr (PRESCALED) normals6_pair<R>(K.rk_step, (unsigned long long)(grp * CG + c), gid, gid_b, z6, K.bm_c2);
        else normals6_pair<R>(K.rk_step, (unsigned long long)(grp * CG + c), gid, gid_b, z6);
#pragma unroll
        for (int j = 0; j < NDS_NPC; j++)
          if (NDS_NPC * c + j < G * NN) zb[NDS_NPC * c + j] = z6[j];
      }
    }
#pragma unroll
    for (int i = 0; i < G; i++) {
      if constexpr (LAZY) {
        // the calls that complete the normals of step i: keeps at most NN + NDS_NPC - 1 normals live instead of G * NN
        if (!ext_noise) {
#pragma unroll
          for (int c = (i * NN + NDS_NPC - 1) / NDS_NPC; c < ((i + 1) * NN + NDS_NPC - 1) / NDS_NPC; c++) {
            R z6[NDS_NPC];
            normals6_pair<R>(K.rk_step, (unsigned long long)(grp * CG + c), gid, gid_b, z6);
#pragma unroll
            for (int j = 0; j < NDS_NPC; j++)
              if (NDS_NPC * c + j < G * NN) zb[NDS_NPC * c + j] = z6[j];
          }
        }
      }
      if (i >= first && left > 0) {
        long long s = s_end - left;  // NDRun.step before this update
       if (alive) {
        R xi[ND], eta[ND];
        if (NN > 0) {
          if (PACK == 1 && ext_noise) {  // md.py:120,127-128 with numbers supplied by the caller (parity tests)
            const double* row = A.noise + ((s - A.noise_first) * NN) * M + lw;
#pragma unroll
            for (int d = 0; d < ND; d++) {
              xi[d] = (R)row[(long long)d * M];
              if (PRESCALED) xi[d] = K.c2 * xi[d];
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
        if constexpr (INTEG == NDS_INT_MINIMIZE) {
          // Minimize.update, engines/minimize.py:41-82: x += f/m dt^2 with dt = 10 * config dt; a step
          // that does not lower pe + biase (or starts below gtol) is undone and stops the walker, but
          // pe / forces / colv stay those of the trial point.
          const R pe_prev = e.pe + e.be;
          R n2 = (R)0, xold[ND];
#pragma unroll
          for (int d = 0; d < ND; d++) { n2 += f[d] * f[d]; xold[d] = x[d]; x[d] += f[d] / K.m * K.dt * K.dt; }
          eval_all<R, ND, true>(K, pot, cv, tcv, cd, bias_mask, x, wb, hills, A.n_hills, hstride, e, gv);
#pragma unroll
          for (int d = 0; d < ND; d++) f[d] = bias_mask ? e.fp[d] + e.fb[d] : e.fp[d];
          const R dE = e.pe + e.be - pe_prev;
          bool stop = false;
          if (dE >= (R)0) stop = true;
          if (nds_sqrt(n2) < K.min_gtol) stop = true;
          if (stop) {
#pragma unroll
            for (int d = 0; d < ND; d++) x[d] = xold[d];
          }
          if (s > K.min_maxiter) stop = true;  // Minimize.step counts the updates done so far
          min_stop = stop;
        } else {
        // ---- first half kick + drift, md.py:164-178
#pragma unroll
        for (int d = 0; d < ND; d++) {
          R dx;
          if constexpr (is_f32<R>::value) {
            if (INTEG == NDS_INT_LANGEVIN) {
              v[d] = fmaf(K.hdtm, f[d], v[d]);
              v[d] = fmaf(K.c1, v[d], xi[d]);  // xi = c2 * normal (PRESCALED)
              dx = v[d] * K.dt;
            } else if (INTEG == NDS_INT_LANGEVIN2) {
              // v += c1 f/m - c2 v + c3 xi - c4 eta  ==  a v + (c1/m) f + c3 xi - c4 eta,  a = 1 - c2
              v[d] = fmaf(K.l2_a, v[d], fmaf(K.l2_c1m, f[d], fmaf(K.c3, xi[d], -(K.c4 * eta[d]))));
              dx = fmaf(K.l2_c5dt, eta[d], K.dt * v[d]);
            } else {
              v[d] = fmaf(K.hdtm, f[d], v[d]);
              dx = v[d] * K.dt;
            }
          } else {
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
          }
          x[d] = x[d] + dx;
        }
        // ---- colvar, potential and biases at the new positions, md.py:179-188
        eval_all<R, ND, false, PF>(K, pot, cv, tcv, cd, bias_mask, x, wb, hills, A.n_hills, hstride, e, gv);
#pragma unroll
        for (int d = 0; d < ND; d++) f[d] = bias_mask ? e.fp[d] + e.fb[d] : e.fp[d];
        // ---- second half kick, md.py:190-198
#pragma unroll
        for (int d = 0; d < ND; d++) {
          if constexpr (is_f32<R>::value) {
            if (INTEG == NDS_INT_LANGEVIN) {
              v[d] = fmaf(K.hdtm, f[d], v[d]);
              v[d] = fmaf(K.c1, v[d], xi[d]);
            } else if (INTEG == NDS_INT_LANGEVIN2) {
              v[d] = fmaf(K.l2_a, v[d], fmaf(K.l2_c1m, f[d], fmaf(K.c3, xi[d], -(K.c4 * eta[d]))));
            } else {
              v[d] = fmaf(K.hdtm, f[d], v[d]);
            }
          } else {
            if (INTEG == NDS_INT_LANGEVIN) {
              v[d] += f[d] / (R)2.0 / K.m * K.dt;
              v[d] = K.c1 * v[d] + K.c2 * xi[d];
            } else if (INTEG == NDS_INT_LANGEVIN2) {
              v[d] += K.c1 * f[d] / K.m - K.c2 * v[d] + K.c3 * xi[d] - K.c4 * eta[d];
            } else {
              v[d] += f[d] / K.m * K.dt / (R)2.0;
            }
          }
        }
        }  // !MINIMIZE
        if (INTEG == NDS_INT_RESCALE) {  // md.py:219-224; `time` is the pre-step time (ndrun.py:230)
          const long long count = (long long)floor(time / K.rescale_freq) + 1;
          if (count > rescale_count) {
            rescale_count = count;
            R n2 = (R)0;
#pragma unroll
            for (int d = 0; d < ND; d++) n2 += v[d] * v[d];
            const R scale = K.v_target / nds_sqrt(n2);
#pragma unroll
            for (int d = 0; d < ND; d++) v[d] *= scale;
          }
        }
       }  // alive
        left -= 1;
        s += 1;            // ndrun.py:232
        time += K.dt64;    // ndrun.py:240
        if constexpr (PACK == 1 && S::BIAS < 0) {
          // Stat.append (ndrun.py:250-251, io/stat.py:70-72): atoms.colv joins the history when step % stat.freq == 0
          if ((bias_mask & BIAS_CVHIST) && s % K.stat_freq == 0) {
            A.cvhist[((A.cvhist_n + cv_taken) % A.cvhist_cap) * M + w] = e.cv.c[0];
            cv_taken += 1;
          }
        }
        if constexpr (PACK == 1) {
          if (committor && alive) {
            bool stopped = (INTEG == NDS_INT_MINIMIZE) && min_stop;  // committor.py:95-97
            if (left <= commit_left) {  // committor.py:100-111
              basin = commit_test<R, ND, INTEG == NDS_INT_MINIMIZE, S::NBASIN>(K, cd, e.cv.c, e.pe);
              if (basin >= 0) stopped = true;
            }
            if (stopped) { alive = false; exit_s = s; }
          }
          bool event = false;
          if (any_frames) {
            rd += 1; if (rd == dfreq) rd = 0;
            if (sfreq > 0) { rs += 1; if (rs == sfreq) rs = 0; }
            event = rd == 0 || (sfreq > 0 && rs == 0 && dfreq - rd <= sfreq);
          }
          if (alive || exit_s == s) {
          if (dumping && event) {
            // io/dump.py:76-102: full observables of this walker at NDRun.step == s
            Eval<R, ND> o;
            eval_all<R, ND, true>(K, pot, cv, tcv, cd, bias_mask, x, wb, hills, A.n_hills, hstride, o, gv);
            write_frame<R, ND>(K, A, w, frame, s, time, cd, bias_mask, x, v, o, false);
          }
          if constexpr (S::DUMP < 0 || (S::DUMP & 2) != 0) {
            if (hist_on && any_frames && rd == 0)  // colvars of the step's own evaluation = atoms.colv at this step
              hist_count(hist_bins, (float)e.cv.c[0], cd > 1 ? (float)e.cv.c[ND > 1 ? 1 : 0] : K.hist_lo[1], K.hist_lo[0],
                         K.hist_inv[0], K.hist_nx, K.hist_lo[1], K.hist_inv[1], K.hist_ny);
          }
          }
          if (event) frame += 1;
        } else {
          if constexpr (S::DUMP == 2) {
            // histogram only: one compare per step against the value `left` has at the next dump step
            // warp-uniform countdown, updated by a select (not inside the branch) so that it stays on the uniform datapath
            const bool hist_now = to_hist == 1;
            to_hist = hist_now ? dfreq : to_hist - 1;
            if (hist_now) {
              hist_count(hist_bins, e.cv.c[0].v.x, cd > 1 ? e.cv.c[ND > 1 ? 1 : 0].v.x : K.hist_lo[1], K.hist_lo[0], K.hist_inv[0],
                         K.hist_nx, K.hist_lo[1], K.hist_inv[1], K.hist_ny);
              hist_count(hist_bins, e.cv.c[0].v.y, cd > 1 ? e.cv.c[ND > 1 ? 1 : 0].v.y : K.hist_lo[1], K.hist_lo[0], K.hist_inv[0],
                         K.hist_nx, K.hist_lo[1], K.hist_inv[1], K.hist_ny);
            }
          } else if constexpr (S::DUMP > 0) {  // packed md kernel: frames (+ histogram), no committor here (static_assert)
            rd += 1; if (rd == dfreq) rd = 0;
            if (sfreq > 0) { rs += 1; if (rs == sfreq) rs = 0; }
            const bool event = rd == 0 || (sfreq > 0 && rs == 0 && dfreq - rd <= sfreq);
            if (dumping && event) {
              Eval<R, ND> o;
              eval_all<R, ND, true>(K, pot, cv, tcv, cd, bias_mask, x, wb, hills, A.n_hills, hstride, o, gv);
              write_frame<R, ND>(K, A, w, frame, s, time, cd, bias_mask, x, v, o, false);
            }
            if constexpr ((S::DUMP & 2) != 0) {
              if (rd == 0) {
                hist_count(hist_bins, e.cv.c[0].v.x, cd > 1 ? e.cv.c[ND > 1 ? 1 : 0].v.x : K.hist_lo[1], K.hist_lo[0], K.hist_inv[0],
                           K.hist_nx, K.hist_lo[1], K.hist_inv[1], K.hist_ny);
                hist_count(hist_bins, e.cv.c[0].v.y, cd > 1 ? e.cv.c[ND > 1 ? 1 : 0].v.y : K.hist_lo[1], K.hist_lo[0], K.hist_inv[0],
                           K.hist_nx, K.hist_lo[1], K.hist_inv[1], K.hist_ny);
              }
            }
            if (event) frame += 1;
          }
          if (committor && left <= commit_left) {  // committor.py:100-111, once per lane
            float c0[ND], c1[ND];
#pragma unroll
            for (int k = 0; k < ND; k++) { c0[k] = e.cv.c[k].v.x; c1[k] = e.cv.c[k].v.y; }
            const int b0 = commit_test<float, ND, false, S::NBASIN>(K, cd, c0, 0.f);
            const int b1 = commit_test<float, ND, false, S::NBASIN>(K, cd, c1, 0.f);
            const bool hit0 = a0 && b0 >= 0, hit1 = a1 && b1 >= 0;
            if (hit0 || hit1) {  // rare
              if (hit0) {
                a0 = false; exit_s = s; basin = b0;
#pragma unroll
                for (int d = 0; d < ND; d++) { xf[d].v.x = x[d].v.x; vf[d].v.x = v[d].v.x; }
              }
              if (hit1) {
                a1 = false; exit_b = s; basin_b = b1;
#pragma unroll
                for (int d = 0; d < ND; d++) { xf[d].v.y = x[d].v.y; vf[d].v.y = v[d].v.y; }
              }
            }
          }
        }
      }
    }
    first = 0;
    grp += 1;
  }
  const long long s = s_end - left;

  // ---- launch exit: observables at the final positions (md.py:210-217,226-229) and state write-back
  if constexpr (PACK == 2) {
    if (committor) {  // stopped lanes go back to the state of their exit step (or of launch entry)
#pragma unroll
      for (int d = 0; d < ND; d++) {
        if (!a0) { x[d].v.x = xf[d].v.x; v[d].v.x = vf[d].v.x; }
        if (!a1) { x[d].v.y = xf[d].v.y; v[d].v.y = vf[d].v.y; }
      }
    }
  }
  Eval<R, ND> o;
  if constexpr (INTEG == NDS_INT_MINIMIZE) o = e;  // pe / forces / colv of the last (trial) point, minimize.py:61-67
  else eval_all<R, ND, true>(K, pot, cv, tcv, cd, bias_mask, x, wb, hills, A.n_hills, hstride, o, gv);
  R v2 = (R)0;
#pragma unroll
  for (int d = 0; d < ND; d++) {
    A.x[d * M + w] = x[d];
    A.v[d * M + w] = v[d];
    A.f[d * M + w] = o.fp[d];
    A.fixf[d * M + w] = o.fb[d];
    v2 += v[d] * v[d];
  }
#pragma unroll
  for (int k = 0; k < ND; k++)
    if (k < cd) A.colv[k * M + w] = o.cv.c[k];
  const R ke = v2 * K.m / (R)2.0;  // md.py:227
  A.thermo[1 * M + w] = o.pe;
  A.thermo[3 * M + w] = o.be;
  if (INTEG != NDS_INT_MINIMIZE || A.begin_mode) {  // Minimize never touches atoms.ke / T / totale (minimize.py:41-82)
    A.thermo[0 * M + w] = A.begin_mode ? ke / (R)ND * (R)2.0 / K.kB : ke / (R)ND / K.kB * (R)2.0;  // ndrun.py:195 / md.py:228
    A.thermo[2 * M + w] = ke;
    A.thermo[4 * M + w] = A.begin_mode ? ke + o.pe : o.pe + ke + o.be;  // ndrun.py:196 / md.py:229
  }
  if constexpr (PACK == 1) {
    if (committor) {
      if (!alive) { A.alive[w] = 0; A.exit_step[w] = exit_s; }
      A.basin[w] = basin;
    }
    // frame 0: Dump.begin() writes the initial state (io/dump.py:45-47)
    if (A.begin_mode && dumping) write_frame<R, ND>(K, A, w, A.frame_base, s, time, cd, bias_mask, x, v, o, true);
    if constexpr (S::BIAS < 0) {  // NDRun.begin appends the initial state to Stat (ndrun.py:201)
      if (A.begin_mode && (bias_mask & BIAS_CVHIST)) A.cvhist[(A.cvhist_n % A.cvhist_cap) * M + w] = o.cv.c[0];
    }
  } else {
    if constexpr (S::DUMP > 0) {
      if (A.begin_mode && dumping) write_frame<R, ND>(K, A, w, A.frame_base, s, time, cd, bias_mask, x, v, o, true);
    }
    if (committor) {
      if (was0) {
        if (!a0) { A.alive[2 * w] = 0; A.exit_step[2 * w] = exit_s; }
        A.basin[2 * w] = basin;
      }
      if (was1) {
        if (!a1) { A.alive[2 * w + 1] = 0; A.exit_step[2 * w + 1] = exit_b; }
        A.basin[2 * w + 1] = basin_b;
      }
    }
  }
  if constexpr (S::DUMP < 0 || (S::DUMP & 2) != 0) {
    if (hist_on) {  // one global atomic per non-empty bin per CTA (threads that left early count as arrived)
      __syncthreads();
      for (int b = threadIdx.x; b < K.hist_nx * K.hist_ny; b += blockDim.x)
        if (hist_bins[b]) atomicAdd(&A.hist[b], (unsigned long long)hist_bins[b]);
    }
  }
}

template <class S>
__global__ void __launch_bounds__(NDS_BLOCK, md_block_min_ctas<S>())
md_block(const __grid_constant__ Consts<typename S::R> K, const __grid_constant__ BlockArgs<typename S::R> A) {
  md_body<S>(K, A);
}

// same body, CTAs of up to 512 threads: with mtd_grid the shared-memory copy of the grid leaves room for one CTA
// per SM, so large ensembles need wide CTAs to keep the SM busy
template <class S>
__global__ void __launch_bounds__(NDS_BLOCK_WIDE, 1)
md_block_wide(const __grid_constant__ Consts<typename S::R> K, const __grid_constant__ BlockArgs<typename S::R> A) {
  md_body<S>(K, A);
}

// ---------------------------------------------------------------------------------- eval_at
// potential.compute(x), colvar.compute/jacobian(x), sum of fix.compute(x0, col0) at n configurations.
template <typename R, int ND>
__global__ void __launch_bounds__(NDS_BLOCK)
eval_at_kernel(const __grid_constant__ Consts<R> K, long long n, const R* __restrict__ xin,
               const R* hills, long long n_hills, long long elem_stride, R* V, R* F, R* colv, R* J,
               R* Vb, R* Fb, const R* grid = nullptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  R x[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) x[d] = xin[d * n + i];
  WalkerBias<R, ND> wb;
#pragma unroll
  for (int k = 0; k < ND; k++) { wb.k[k] = (R)0; wb.r0[k] = (R)0; }
  Eval<R, ND> e;
  const int cd = K.cd;
  eval_all<R, ND, true>(K, K.pot, K.cv, K.tcv, cd, K.bias_mask & ~BIAS_UMB_WALKER, x, wb, hills,
                        n_hills, elem_stride, e, GridView<R>{grid, 0, nullptr});
  if (V) V[i] = e.pe;
  if (Vb) Vb[i] = e.be;
#pragma unroll
  for (int d = 0; d < ND; d++) {
    if (F) F[d * n + i] = e.fp[d];
    if (Fb) Fb[d * n + i] = e.fb[d];
  }
  for (int k = 0; k < cd; k++) {
    if (colv) colv[k * n + i] = e.cv.c[k];
    if (J) {  // row k of the Jacobian = pull-back of the k-th unit vector
      R g[ND], row[ND];
#pragma unroll
      for (int d = 0; d < ND; d++) g[d] = (d == k) ? (R)1 : (R)0;
      cv_pullback<R, ND>(K.cv, K.rot, x, e.cv, g, row);
#pragma unroll
      for (int d = 0; d < ND; d++) J[(k * ND + d) * n + i] = row[d];
    }
  }
}

// ---------------------------------------------------------------------------------- deposit
// Gaussian._deposite (bias/gaus.py:148-213), one new hill per live walker at its current colvar with
// the well-tempered height w_init * exp(-VR / (kB (biasf-1) T)) from the STALE VR; the
// first deposition uses w_init.  Shared table: walker w of this rank writes record n_hills +
// walker_offset + w, i.e. straight into its slot of the segment an allgather completes (section 8e).
template <typename R>
__global__ void deposit_kernel(const __grid_constant__ Consts<R> K, long long M, long long walker_offset,
                               const R* __restrict__ colv, const R* __restrict__ vr,
                               const unsigned char* alive, R* hills, long long n_hills, int first) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= M) return;
  const bool live = alive ? alive[w] != 0 : true;
  R h = first ? K.mtd_w : K.mtd_w * nds_exp(-vr[w] / K.mtd_kBdT);
  if (!live) h = (R)0;
  const int cd = K.cd, rs = K.hill_stride, hi = K.hill_hidx;
  if (K.mtd_shared) {
    R* rec = hills + (n_hills + walker_offset + w) * rs;
    for (int k = 0; k < cd; k++) rec[k] = colv[k * M + w];
    for (int k = cd; k < hi; k++) rec[k] = (R)0;  // colvardim 1 in a 4-word record: second centre = 0
    rec[hi] = h;
    // 4-word records: the last word carries log2(height) for the MUFU-bound fp32 loop of mtd_block
    // (-inf for a zero-height record, which then contributes exactly 0)
    if (!K.mtd_adaptive)
      for (int k = hi + 1; k < rs; k++) rec[k] = (R)log2((double)h);
  } else {
    for (int k = 0; k < cd; k++) hills[(n_hills * rs + k) * M + w] = colv[k * M + w];
    hills[(n_hills * rs + hi) * M + w] = h;
  }
}

// adapsig_geo (gaus.py:162-165, 189-192): variance of the new hill of every walker, s = J.J^T at the
// deposition point (first hill of a table) or J.J^T * sigma^2 (later hills); fields 2, 3 of the record
// get 1/s and s.  Runs right after deposit_kernel, before any exchange.  colvardim 1.
template <typename R, int ND>
__global__ void hill_variance_kernel(const __grid_constant__ Consts<R> K, long long M, long long walker_offset,
                                     const R* __restrict__ x, R* hills, long long n_hills, int first) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= M) return;
  R xx[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) xx[d] = x[d * M + w];
  CvOut<R, ND> o;
  cv_eval<R, ND>(K, K.cv, K.rot, xx, o);
  R g[ND], row[ND];
#pragma unroll
  for (int d = 0; d < ND; d++) g[d] = d == 0 ? (R)1 : (R)0;
  cv_pullback<R, ND>(K.cv, K.rot, xx, o, g, row);  // row 0 of colvar.jacobian(x)
  R jj = (R)0;
#pragma unroll
  for (int d = 0; d < ND; d++) jj += row[d] * row[d];
  const R s = first ? jj : jj * K.mtd_sig2[0];
  const R inv = (R)1.0 / s;
  const int rs = K.hill_stride;
  if (K.mtd_shared) {
    R* rec = hills + (n_hills + walker_offset + w) * rs;
    rec[2] = inv; rec[3] = s;
  } else {
    hills[(n_hills * rs + 2) * M + w] = inv;
    hills[(n_hills * rs + 3) * M + w] = s;
  }
}

// adapsig without adapsig_geo (gaus.py:166-183 first hill, :193-212 later hills): the variance of the new hill is the
// exponentially weighted variance of the walker's recent colvar history, Stat.colv, kept by the step kernel in a ring
// (BIAS_CVHIST).  n samples exist; the window is the last tao_d of them -- the FIRST hill of a table takes
// stat.colv[-tao_d:-1] (one sample fewer) when the history is longer than tao_d.  A sample within 1e-10 of the mean
// (in particular a history of one sample) gives variance 1.  colvardim 1.
template <typename R>
__global__ void hill_variance_hist_kernel(const __grid_constant__ Consts<R> K, long long M, long long walker_offset,
                                          const R* __restrict__ ring, int cap, long long n, int tao_d,
                                          R* hills, long long n_hills, int first) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= M) return;
  long long j0, j1;  // samples [j0, j1)
  if (n <= tao_d) { j0 = 0; j1 = n; }
  else if (first) { j0 = n - tao_d; j1 = n - 1; }
  else { j0 = n - tao_d; j1 = n; }
  const int n0 = (int)(j1 - j0);
  R avg = (R)0;
  for (long long j = j0; j < j1; j++) avg += ring[(j % cap) * M + w];
  avg /= (R)n0;  // np.average(colv, axis=0)
  R mind = (R)3e38, acc = (R)0;
  for (int i = 0; i < n0; i++) {
    const R d = ring[((j0 + i) % cap) * M + w] - avg;
    const R ad = d < (R)0 ? -d : d;
    mind = ad < mind ? ad : mind;
    acc += nds_exp(-((R)(n0 - i)) / (R)n0) * d * d;  // exp(-(n0 - arange(n0)) / n0), gaus.py:180,208
  }
  const R s = (n0 == 0 || mind < (R)1e-10) ? (R)1 : acc / (R)n0;
  const R inv = (R)1.0 / s;
  const int rs = K.hill_stride;
  if (K.mtd_shared) {
    R* rec = hills + (n_hills + walker_offset + w) * rs;
    rec[2] = inv; rec[3] = s;
  } else {
    hills[(n_hills * rs + 2) * M + w] = inv;
    hills[(n_hills * rs + 3) * M + w] = s;
  }
}

// Fused deposition + exchange over peer memory (shared table, multi-GPU): the record of every local
// walker is stored straight into the hill table of EVERY rank through NVLink-mapped peer pointers
// (tables[r], r = 0..nranks-1, own table included), at the same slot n_hills + global walker id.  One
// 16-byte (fp32) / 32-byte (fp64) vector store per rank per walker; no allgather, no staging buffer.
template <typename R>
__global__ void deposit_p2p_kernel(const __grid_constant__ Consts<R> K, long long M, long long walker_offset,
                                   const R* __restrict__ colv, const R* __restrict__ vr,
                                   const unsigned char* alive, R* const* __restrict__ tables, int nranks,
                                   long long n_hills, int first) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < M) {
  const bool live = alive ? alive[w] != 0 : true;
  R h = first ? K.mtd_w : K.mtd_w * nds_exp(-vr[w] / K.mtd_kBdT);
  if (!live) h = (R)0;
  const int cd = K.cd, rs = K.hill_stride, hi = K.hill_hidx;
  R rec[8];
  for (int k = 0; k < cd; k++) rec[k] = colv[k * M + w];
  for (int k = cd; k < hi; k++) rec[k] = (R)0;
  rec[hi] = h;
  for (int k = hi + 1; k < rs; k++) rec[k] = (R)log2((double)h);
  const long long slot = (n_hills + walker_offset + w) * rs;
  for (int r = 0; r < nranks; r++) {
    R* dst = tables[r] + slot;
    if (rs == 4) {
      if constexpr (sizeof(R) == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(rec[0], rec[1], rec[2], rec[3]);
      } else {
        reinterpret_cast<double2*>(dst)[0] = make_double2(rec[0], rec[1]);
        reinterpret_cast<double2*>(dst)[1] = make_double2(rec[2], rec[3]);
      }
    } else {
      for (int k = 0; k < rs; k++) dst[k] = rec[k];
    }
  }
  }
  // records visible to the peers before the barrier kernel signals: one system fence per CTA behind the block
  // barrier (fences are cumulative; a fence per thread is several microseconds slower, see grid_exchange_kernel)
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

// ------------------------------------------------------------------------- FES grid, histogram
// Gaussian.projection (bias/gaus.py:365-393, 413-425): sum of hills on a regular grid.
template <typename R>
__global__ void fes_grid_kernel(const __grid_constant__ Consts<R> K, const R* __restrict__ hills,
                                long long n_hills, long long elem_stride, double x0, double dx,
                                long long nx, double y0, double dy, long long ny, double* grid) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nx * ny) return;
  const long long i = p % nx, j = p / nx;
  const double X = x0 + i * dx, Y = y0 + j * dy;
  const int cd = K.cd, rs = K.hill_stride;
  const double s0 = (double)1.0 / ((double)K.mtd_invsig2[0]);
  const double s1 = cd > 1 ? (double)1.0 / ((double)K.mtd_invsig2[1]) : 1.0;
  double acc = 0.0;
  for (long long k = 0; k < n_hills; k++) {
    const R* h = hills + k * rs * elem_stride;
    const double hx = (double)h[0];
    double ex = -((X - hx) * (X - hx)) / s0 / 2.0;
    if (K.mtd_adaptive) ex = -(((X - hx) * (X - hx)) * (double)h[2 * elem_stride] / 2.0);  // gaus.py:428-434
    if (cd > 1) {
      const double hy = (double)h[elem_stride];
      ex += -((Y - hy) * (Y - hy)) / s1 / 2.0;
    }
    acc += (double)h[K.hill_hidx * elem_stride] * exp(ex);
  }
  grid[p] = acc;
}

// histogram of the current colvar values (io/plot.py:562-572): shared-memory privatised bins when they
// fit, one global atomic per non-empty bin per block.
template <typename R>
__global__ void histogram_kernel(long long M, int cd, const R* __restrict__ colv,
                                 const unsigned char* alive, double x0, double inv_dx, int nx,
                                 double y0, double inv_dy, int ny, unsigned long long* counts) {
  extern __shared__ unsigned int bins[];
  const int nb = nx * ny;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) bins[b] = 0u;
  __syncthreads();
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < M;
       w += (long long)gridDim.x * blockDim.x) {
    const double cx = (double)colv[w];
    const long long ix = (long long)floor((cx - x0) * inv_dx);
    long long iy = 0;
    if (cd > 1) iy = (long long)floor(((double)colv[M + w] - y0) * inv_dy);
    if (ix >= 0 && ix < nx && iy >= 0 && iy < ny) atomicAdd(&bins[iy * nx + ix], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += blockDim.x)
    if (bins[b]) atomicAdd(&counts[b], (unsigned long long)bins[b]);
}

// ---------------------------------------------------------------- committor lane compaction
// Stopped walkers idle their lane (57 % of the example shots never commit while others exit
// after a few hundred steps).  Between launches the live walkers are packed into slots [0, n_alive):
// every dead slot in the front ("hole") swaps with a live slot from the back ("mover"), matched by rank
// (exclusive scan of the alive flags), so the next launch covers only live walkers.  ids[] keeps
// slot -> walker so that Philox streams, noise rows and the getters are unaffected.
template <typename R>
__global__ void compact_swap_kernel(long long M, int nd, int cd, const int* holes, const int* movers,
                                    const int* n_holes, R* x, R* v, R* f, R* fixf, R* colv, R* thermo,
                                    R* vr, R* umb_k, R* umb_r0, unsigned char* alive, int* basin,
                                    long long* exit_step, long long* ids) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= *n_holes) return;
  const long long a = holes[r], b = movers[r];
  auto sw = [&](R* arr, int rows) {
    for (int q = 0; q < rows; q++) { const R t = arr[q * M + a]; arr[q * M + a] = arr[q * M + b]; arr[q * M + b] = t; }
  };
  sw(x, nd); sw(v, nd); sw(f, nd); sw(fixf, nd); sw(colv, cd); sw(thermo, 5); sw(vr, 1);
  if (umb_k) { sw(umb_k, cd); sw(umb_r0, cd); }
  { const unsigned char t = alive[a]; alive[a] = alive[b]; alive[b] = t; }
  { const int t = basin[a]; basin[a] = basin[b]; basin[b] = t; }
  { const long long t = exit_step[a]; exit_step[a] = exit_step[b]; exit_step[b] = t; }
  { const long long t = ids[a]; ids[a] = ids[b]; ids[b] = t; }
}

// empty hill records: height 0, log2(height) = -inf
template <typename R>
__global__ void fill_hills_kernel(long long n_records, int rs, int hidx, R* hills) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_records) return;
  for (int k = 0; k < rs; k++) hills[i * rs + k] = (k > hidx) ? (R)(-INFINITY) : (R)0;
}

template <typename A, typename B>
__global__ void cast_kernel(long long n, const A* __restrict__ in, B* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (B)in[i];
}
template <typename T>
__global__ void convert_kernel(long long n, const double* __restrict__ in, T* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (T)in[i];
}
template <typename T>
__global__ void convert_back_kernel(long long n, const T* __restrict__ in, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

}  // namespace nds
