// nds_common.cuh -- shared host/device definitions of the B200 multi-walker engine.
//
// Device-side constants, kernel argument blocks and the compile-time "Spec" that selects a kernel
// variant.  A Spec field equal to DYN means "read the choice from the constant block at run time"
// (generic kernels); a non-negative value bakes the choice into the kernel (hot-path variants).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/ndsb200.h"
#include "nds_f2.cuh"
#include "nds_philox.cuh"

namespace nds {

constexpr int DYN = -1;

// bias bit mask
constexpr int BIAS_UMB = 1;         // >= 1 harmonic restraint (bias/umb.py)
constexpr int BIAS_MTD = 2;         // metadynamics hill sum (bias/gaus.py)
constexpr int BIAS_PE = 4;          // potential-energy bias (bias/pe.py)
constexpr int BIAS_UMB_WALKER = 8;  // restraint 0 takes (k, r0) per walker
constexpr int BIAS_MTD_GRID = 16;   // metadynamics bias interpolated from the accumulated grid (mtd_grid)
constexpr int BIAS_CVHIST = 32;     // the step kernel keeps Stat.colv (io/stat.py:70-72) of every walker in a ring:
                                    // history-based adaptive hill widths (gaus.py:166-183, 193-212)

template <typename R_, int ND_, int POT_ = DYN, int CV_ = DYN, int TCV_ = DYN, int INTEG_ = DYN,
          int BIAS_ = DYN, int METHOD_ = DYN, int NOISE_ = DYN, int DUMP_ = DYN, int NBASIN_ = DYN>
struct Spec {
  using R = R_;
  static constexpr int ND = ND_;
  static constexpr int POT = POT_;
  static constexpr int CV = CV_;
  static constexpr int TCV = TCV_;
  static constexpr int INTEG = INTEG_;
  static constexpr int BIAS = BIAS_;
  static constexpr int METHOD = METHOD_;
  static constexpr int NOISE = NOISE_;
  static constexpr int DUMP = DUMP_;
  static constexpr int NBASIN = NBASIN_;  // committor: number of basin boxes when known at compile time
};

// Uniform constants of a run (one copy per engine, passed by value to every kernel).
template <typename R>
struct Consts {
  int ndim, cd, tcd;
  int pot, cv, tcv, integ, method, noise, bias_mask, n_umb, n_basins;
  int mtd_shared, hill_stride;  // reals per hill record (cd + 1, padded to 4 when cd == 2 or adaptive)
  int mtd_adaptive;             // per-hill variance (adapsig_geo, colvardim 1): record = (c, w, 1/s, s)
  int hill_hidx;                // field of the height: 2 in the 4-word records (c0, c1 | 0, w, log2 w), else cd
  // units / integrator (engines/md.py:81-99, data.py:26-28, constant.py)
  R m, inv_m, dt, kB, kBT;
  R c1, c2, c3, c4, c5, v_target;
  R hdtm;           // dt / (2 m): fused half-kick factor of the fp32 path
  R l2_a, l2_c1m, l2_c5dt;  // fp32 2nd-order Langevin: 1 - c2, c1 / m, c5 * dt
  R vel_scale;      // sqrt(kB * T_init / m)  (engines/modify.py:45)
  double dt64, rescale_freq;
  // four-Gaussian surfaces (Mueller*: potentials/mueller.py:8-16; ThreeHole*: three_hole.py:10-22)
  R gA[4], ga[4], gb[4], gc[4], gx[4], gy[4];
  // fp32 fast form: G = k*grad (k = log2(e)/2), A2 = A / k  (see nds_physics.cuh)
  R fa[4], fb[4], fc[4], fA[4];
  BmScale bm_c2;  // Box-Muller radius constants scaled by c2^2 (fp32 first-order Langevin: noise arrives as c2 * xi)
  R flA[4];  // log2 |fA|: the amplitude rides in the exponent, its (class-constant) sign in the FMA operand
  R th_D, th_E, th_x1, th_y1, th_ref;
  R hq_k[5], hq_q[5], hq_c[5];  // NDS_POT_HARMONIC_QUARTIC (extension)
  R rot;            // 1/sqrt(2) as NumPy computes it (colvars/two.py:53)
  // path colvar (colvars/path.py): reference points [n][path_dim], indices [n] in device memory
  const R* path_ref; const R* path_ind; int n_path, path_dim;
  R path_sig2_invh, path_inv_sig2, path_N;
  // DoubleWell(1d): K1, K2, x1[2], x2[2], invsigma2 (double_well.py:9-53)
  R dw_K1, dw_K2, dw_x1[2], dw_x2[2], dw_invsig2;
  // biases
  R umb_k[NDS_MAX_UMB][NDS_MAX_CV], umb_r0[NDS_MAX_UMB][NDS_MAX_CV];
  R mtd_w, mtd_invsig2[NDS_MAX_CV], mtd_kBdT;  // kB*(biasf-1)*T, gaus.py:102-103
  R mtd_sig2[NDS_MAX_CV];                      // sigma^2 (gaus.py:90)
  double mtd_dep_freq64; long long mtd_capacity;  // in-kernel deposition of the MC engine (per-walker clocks)
  R mtd_pre[NDS_MAX_CV];  // fp32 fast form: sqrt(log2(e) / 2) / sigma
  // grid-accumulated bias (mtd_grid): point (i, j) at lo + (i, j) * h holds (V, h0 dV/dc0, h1 dV/dc1, h0 h1 d2V/dc0dc1)
  int grid_nx, grid_ny;
  R grid_lo[2], grid_h[2], grid_invh[2];
  R pe_thred;
  R crit[NDS_MAX_BASINS][NDS_MAX_CV][2];
  long long diffusion_time;
  // Minimize (engines/minimize.py) and the T = 0 committor criterion pe < emin (committor.py:128-131)
  R min_gtol, emin; long long min_maxiter; int zero_temperature;
  long long dump_freq, stat_freq, dump_walkers;
  int hist_nx, hist_ny;  // in-kernel colvar histogram (0: off)
  float hist_lo[2], hist_inv[2];
  unsigned int key0, key1;  // Philox key = seed
  PhiloxKeys rk_step;       // round keys of the per-step noise stream (domain 0)
  PhiloxKeys rk_init;       // round keys of the initial-velocity stream (domain 1)
  PhiloxKeys rk_mc;         // round keys of the Monte Carlo stream (domain 2)
  // Metropolis MC (engines/mc.py)
  R mc_bounds[NDS_MAX_DIM], mc_gb[NDS_MAX_DIM]; int mc_all, mc_periodic;
};

// Per-launch arguments of the fused step kernel.
template <typename R>
struct BlockArgs {
  long long M;             // local walkers = stride of the SoA arrays
  long long n_active;      // slots [0, n_active) are launched (0 = all M); committor compaction shrinks it
  long long slot0;         // first slot of this launch (walker chunks of the pipelined host path; 0 otherwise)
  const long long* ids;    // slot -> local walker index after compaction (null = identity)
  long long walker_offset; // global id of local walker 0
  long long step0;         // NDRun.step before the first step of this launch
  long long nsteps;        // steps to run in this launch
  double time0;            // NDRun.time before the first step (fp64 accumulation, ndrun.py:240)
  long long rescale_count; // MD.rescale_count at launch start (md.py:77,220-222)
  // state, SoA [d][M]
  R* x; R* v; R* f; R* fixf; R* colv; R* thermo; R* vr;
  const R* umb_k; const R* umb_r0;     // per-walker window [cd][M] or null
  unsigned char* alive; int* basin; long long* exit_step;  // committor
  // hills
  const R* hills; long long n_hills;
  // accumulated bias grid (mtd_grid): 4 reals per point, row-major [ny][nx]; grid_smem = 1: the launch carries
  // dynamic shared memory for a copy of the whole grid
  const R* grid; int grid_smem; unsigned long long* grid_outside;
  int refresh_vr;          // 1: the launch follows a deposition (or begin): store V_mtd into vr[]
  int need_init_v;         // 1: draw Maxwell-Boltzmann velocities first (modify.py:44-54)
  int begin_mode;          // 1: NDRun.begin semantics: totale = ke + pe, bias omitted (ndrun.py:196)
  // external noise table [step - noise_first][j][M] (parity tests)
  const double* noise; long long noise_first;
  // frames: rows in the engine's real type, [frame][field][walker]
  R* frames; long long frame_base; long long frame_cap; int frame_width;
  // colvar histogram accumulated at dump cadence: global counts [ny][nx]; shared-memory bins at this byte offset
  unsigned long long* hist; int hist_smem_off;
  int dyn_smem;            // dynamic shared memory of the launch: grid copy (mtd_grid) + histogram bins
  // Stat.colv ring (BIAS_CVHIST): sample j of walker w lives at cvhist[(j % cvhist_cap) * M + w]; cvhist_n samples
  // exist at launch entry
  R* cvhist; long long cvhist_n; int cvhist_cap;
  // Harmonic.VR of every restraint [n_umb][M], refreshed where Gaussian.update recomputes every fix (gaus.py:138-146)
  // and by begin(); null unless umbrella AND metadynamics biases are present and frames are recorded
  R* vr_umb;
};

}  // namespace nds
