/* ndsb200.h -- C ABI of the B200-native multi-walker engine for NDSimulator.
 *
 * The reference (mir-group/NDSimulator, pure Python) has no FFI boundary; its hot path sits behind a
 * duck-typed plugin protocol.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).  Conventions:
 *   - C linkage, plain pointers and sizes, no exceptions across the boundary;
 *   - every function returns 0 on success, non-zero on error; nds_last_error() gives the message;
 *   - the caller owns all HOST buffers, the engine owns device memory and its CUDA stream;
 *   - one host thread per engine handle; distinct handles on distinct devices are independent;
 *   - all host-side arrays are fp64 ("double") whatever the engine's compute precision, laid out
 *     structure-of-arrays with the walker index fastest:  x[d * M + w].
 *   - there is NO CPU fallback: without a CUDA device nds_create() fails.
 */
#ifndef NDSB200_H
#define NDSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NDS_API __attribute__((visibility("default")))
#else
#define NDS_API
#endif

#define NDS_ABI_VERSION 3
#define NDS_MAX_DIM 5     /* ndim of the built-in potentials is 1, 2 or 5                       */
#define NDS_MAX_CV 5      /* colvardim <= ndim for every built-in colvar                         */
#define NDS_MAX_UMB 4     /* harmonic restraints acting on one walker (umb_n, bias/__init__.py:31) */
#define NDS_MAX_BASINS 8  /* committor basin boxes (engines/committor.py:31-42)                   */
#define NDS_MAX_PATH_REF 256 /* reference points of a path colvar (colvars/path.py)                */

/* compute precision of the kernels ("precision" key, new) */
enum { NDS_F32 = 0, NDS_F64 = 1 };

/* method: engines/__init__.py:11-31: md, committor, mc (Metropolis Monte Carlo, engines/mc.py) */
enum { NDS_METHOD_MD = 0, NDS_METHOD_COMMITTOR = 1, NDS_METHOD_MC = 2,
       NDS_METHOD_WL = 3 /* Wang-Landau MC, engines/wang_landau.py:65-161 (one basin, f = 2) */ };
#define NDS_WL_GRID 101   /* energy bins over [-1.5, 0.2], wang_landau.py:12-19 */

/* integrate: engines/md.py:59; NDS_INT_MINIMIZE = the fixed-step steepest descent of
 * engines/minimize.py:41-82 (the inner engine of a committor run at temperature 0,
 * engines/__init__.py:22-23) */
enum { NDS_INT_LANGEVIN = 0, NDS_INT_LANGEVIN2 = 1, NDS_INT_NVE = 2, NDS_INT_RESCALE = 3,
       NDS_INT_MINIMIZE = 4 };

/* potential: class names of ndsimulator/potentials (potentials/_build.py:12-18) */
enum {
  NDS_POT_MUELLER2D = 0,   /* potentials/mueller.py:6-45     ndim 2 */
  NDS_POT_MUELLER5D = 1,   /* potentials/mueller.py:108-142  ndim 5, needs true_colvar (cd 2) */
  NDS_POT_MUELLER5DSHL = 2,/* potentials/mueller.py:155-157  A/5 */
  NDS_POT_DOUBLEWELL1D = 3,/* potentials/double_well.py:6-25 ndim 1 */
  NDS_POT_DOUBLEWELL = 4,  /* potentials/double_well.py:34-53 ndim 2 */
  NDS_POT_THREEHOLE2D = 5, /* potentials/three_hole.py:6-57  ndim 2, needs true_colvar (cd 2) */
  NDS_POT_THREEHOLE5D = 6, /* potentials/three_hole.py:72-125 ndim 5, needs true_colvar (cd 2) */
  NDS_POT_GAUSSIAN2D = 7,  /* potentials/misc.py:7-26 */
  NDS_POT_GAUSSIAN = 8,    /* potentials/misc.py:32-43 (factory default, _build.py:10) */
  NDS_POT_FLAT1D = 9,      /* potentials/flat.py:17-24 */
  NDS_POT_FLAT2D = 10,     /* potentials/flat.py:5-14 */
  NDS_POT_HARMONIC_QUARTIC = 11, /* EXTENSION, no reference class: separable N-D well
                                    V = sum_d k_d u_d^2 / 2 + q_d u_d^4 / 4, u = x - centre; ndim 1, 2 or 5.
                                    pot_params: k[0..4], q[5..9], centre[10..14].  Pinned analytically
                                    (harmonic velocity-Verlet map), not against the reference.           */
  NDS_POT_COUNT = 12
};

/* colvar / true_colvar: class names of ndsimulator/colvars (colvars/_build.py:10-19) */
enum {
  NDS_CV_ORIGINAL = 0,     /* colvars/misc.py:6-16  identity, colvardim = ndim */
  NDS_CV_EXTRACTX = 1,     /* colvars/two.py:17-25 */
  NDS_CV_EXTRACTY = 2,     /* colvars/two.py:6-14 */
  NDS_CV_XMY = 3,          /* colvars/two.py:28-36 */
  NDS_CV_XPY = 4,          /* colvars/two.py:39-47 */
  NDS_CV_ROTATE2D = 5,     /* colvars/two.py:50-61 */
  NDS_CV_PROJ2DTO1D = 6,   /* colvars/two.py:64-73 */
  NDS_CV_PROJ5DTO1DX0 = 7, /* colvars/five.py:5-15 */
  NDS_CV_PROJ5DTO1DX2 = 8, /* colvars/five.py:18-28 */
  NDS_CV_PROJ5DTO1D = 9,   /* colvars/five.py:31-44 (Jacobian quirk 2e-7) */
  NDS_CV_PROJ5DTO2D = 10,  /* colvars/five.py:47-71 */
  NDS_CV_PROJ5DTO2DV2 = 11,/* colvars/five.py:74-98 */
  NDS_CV_PATH = 12,        /* colvars/path.py:14-102  Branduardi path variable, colvardim 1, any ndim;
                              reference points set with nds_set_path() */
  NDS_CV_PATH5DTO2D = 13,  /* colvars/path.py:105-180 path variable in Proj5dto2d space, ndim 5 */
  NDS_CV_COUNT = 14
};

/* where the per-step Gaussian numbers come from */
enum {
  NDS_NOISE_PHILOX = 0,    /* Philox4x32-10, key = seed, counter = (stream index, global walker id) */
  NDS_NOISE_EXTERNAL = 1   /* read from a caller-provided table (nds_set_noise); parity tests only   */
};

/* POD configuration.  Every numeric field mirrors a YAML / kwargs key of the reference
 * (SURVEY.md section 8b); the Python front-end fills it with the reference's prefix rules. */
typedef struct nds_config {
  int32_t abi_version;        /* must be NDS_ABI_VERSION */
  int32_t device;             /* CUDA device ordinal */
  int32_t precision;          /* NDS_F32 | NDS_F64 */
  int32_t ndim;               /* ndrun.py:79 */
  int32_t method;             /* ndrun.py:86 */
  int32_t integrator;         /* md.py:52 "integrate" */
  int32_t potential;          /* potentials/_build.py:10 */
  int32_t colvar;             /* colvars/_build.py:10 "colvar_name" */
  int32_t true_colvar;        /* "true_colvar_name" */
  int32_t noise_mode;         /* NDS_NOISE_* */
  int32_t steps_per_launch;   /* K: MD steps fused into one kernel launch (new key) */
  int32_t reserved0;

  int64_t n_walkers;          /* M: walkers owned by this engine (new key n_walkers)         */
  int64_t walker_offset;      /* global id of local walker 0 (multi-GPU sharding)            */
  int64_t n_walkers_global;   /* total walkers over all ranks (shared-hill table layout)     */
  uint64_t seed;              /* ndrun.py:88; Philox key */

  double temperature;         /* ndrun.py:80 */
  double initial_temperature; /* ndrun.py:149 */
  double mass;                /* ndrun.py:81; m = au_mass * mass (data.py:27-28) */
  double dt;                  /* md.py:50 */
  double gamma;               /* md.py:51 */
  double rescale_freq;        /* md.py:53 (time units); <= 0 means "None" */

  /* potential parameters.  Mueller2d/5d: [0..3] x_center, [4..7] y_center (mueller.py:15-16).
   * DoubleWell1d: K1, K2, x1, x2, sigma.  DoubleWell: K1, K2, x1[2], x2[2], sigma.            */
  double pot_params[16];

  /* umbrella sampling, bias/umb.py:9-40: V = 1/2 sum k (R-R0)^2 */
  int32_t n_umb;              /* number of harmonic restraints on each walker (umb_n) */
  int32_t umb_per_walker;     /* 1: restraint 0 takes k, r0 per walker from nds_set_walker_bias */
  double umb_k[NDS_MAX_UMB][NDS_MAX_CV];
  double umb_r0[NDS_MAX_UMB][NDS_MAX_CV];

  /* well-tempered metadynamics, bias/gaus.py:48-128 (fixed sigma) */
  int32_t mtd_enabled;
  int32_t mtd_shared;         /* 1: all walkers (all ranks) share one hill table; 0: one table per walker */
  double mtd_w;               /* initial height */
  double mtd_sigma[NDS_MAX_CV];
  double mtd_dep_freq;        /* deposition period in TIME units (gaus.py:134) */
  double mtd_biasf;           /* bias factor, must be > 1 */
  int64_t mtd_capacity;       /* hills the table can hold (per walker when !mtd_shared) */
  int32_t mtd_adapsig_geo;    /* adapsig + adapsig_geo (gaus.py:162-165,189-192,263-288): every hill carries its
                                 own variance J.J^T (first hill) / J.J^T * sigma^2 (later hills) taken at the
                                 deposition point.  colvardim 1 only: for 2-D colvars the reference takes the
                                 element-wise reciprocal of a matrix with zero off-diagonals (inf / NaN bias). */
  int32_t reserved3;

  /* potential-energy bias, bias/pe.py:30-51 */
  int32_t pe_enabled;
  int32_t reserved1;
  double pe_thred;

  /* committor, engines/committor.py:31-42,118-142 */
  int32_t n_basins;
  int32_t reserved2;
  int64_t diffusion_time;
  double criteria[NDS_MAX_BASINS][NDS_MAX_CV][2];

  /* observable sampling, io/stat.py:57-74 + io/dump.py:76-102 */
  int64_t dump_freq;          /* frame every dump_freq steps; 0 = no frames */
  int64_t dump_walkers;       /* frames are recorded for local walkers [0, dump_walkers) */
  int64_t dump_capacity;      /* frames the device ring can hold between two drains */
  int64_t stat_freq;          /* io/stat.py:10; defines which thermo sample a dump row shows (A16) */

  /* Minimize (engines/minimize.py:8-22) and the T = 0 committor energy criterion (committor.py:41,130) */
  double emin;                /* committor.py:41, default -1.2: at T = 0 a basin needs pe < emin */
  double min_gtol;            /* minimize.py:12, default 1e-5 */
  int64_t min_maxiter;        /* minimize.py:11,26-27, default steps * 10 */

  /* Metropolis Monte Carlo, engines/mc.py:7-21 */
  double mc_bounds[NDS_MAX_DIM];        /* proposal half-widths ("mc_bounds") */
  double mc_global_bounds[NDS_MAX_DIM]; /* periodic box ("global_bounds") */
  int32_t mc_propose_all;               /* propose_dim: 0 = "single" (one random coordinate), 1 = "all" */
  int32_t mc_periodic;                  /* bound_cond == "periodic" (mc.py:57-62) */

  /* grid-accumulated metadynamics bias (new key mtd_grid; the reference's own incremental trick for its FES plot,
   * bias/gaus.py:380-393 `_dep_projection`, promoted to the step loop): every deposition adds the footprint of the
   * new hills to a regular grid of (V, dV/dc0, dV/dc1, d2V/dc0dc1) over [lo, hi] per colvar; the step kernel
   * evaluates the bias by bicubic Hermite interpolation from a shared-memory copy of the grid, so a step costs the
   * same whatever the number of hills.  Shared table, fixed sigma, colvardim 1 or 2.  The exact hill sum
   * (mtd_grid = 0) stays the parity path.                                                                        */
  int32_t mtd_grid;                     /* 0: exact hill sum; 1: grid-accumulated bias */
  int32_t mtd_grid_n[2];                /* grid points per colvar (n[1] = 1 for colvardim 1); spacing (hi-lo)/(n-1) */
  int32_t reserved4;
  double mtd_grid_lo[2], mtd_grid_hi[2];

  /* history-based adaptive sigma (gaus.py:166-183, 193-212: `adapsig` without `adapsig_geo`): the variance of a new
   * hill is the exponentially weighted variance of the walker's last tao_d = min(50 mtd_dep_freq, 100) Stat samples
   * of the colvar (io/stat.py:57-74, sampled every stat_freq steps); the step kernel keeps that window in a device
   * ring.  colvardim 1, method md; per-hill variance records like mtd_adapsig_geo.                                */
  int32_t mtd_adapsig_hist;
  int32_t reserved5;

  /* colvar histogram accumulated ON THE DEVICE over the run ("CV histogram on-device"; the
   * reference's io/plot.py:562-572 histogram, but of every walker at every dump step): bins hist_n per colvar over
   * [hist_lo, hist_hi); sampled at dump_freq cadence (dump_freq > 0 required), read with nds_get_histogram().   */
  int32_t hist_n[2];                    /* 0: off; n[1] = 1 for colvardim 1 */
  double hist_lo[2], hist_hi[2];
} nds_config;

typedef struct nds_engine nds_engine;

/* message of the last failure on this handle (or of the last failed nds_create when e == NULL) */
NDS_API const char* nds_last_error(const nds_engine* e);

/* fills cfg with the reference's defaults (ndrun.py:77-97, md.py:48-55, gaus.py:48-57, ...) */
NDS_API void nds_config_defaults(nds_config* cfg);

/* colvardim of a colvar id at a given ndim, or -1 when the pair is invalid */
NDS_API int nds_colvar_dim(int colvar, int ndim);

/* replaces NDRun.__init__ (ndrun.py:77-171): validates cfg (the reference's assertions become error
 * codes), allocates SoA walker state on cfg->device.  All walkers start at x = 0, v unset.       */
NDS_API int nds_create(const nds_config* cfg, nds_engine** out);
NDS_API void nds_destroy(nds_engine* e);

/* replaces Atom.initialize (data.py:69-76) + Modify.set_velocity (modify.py:38-63).
 * x: [ndim][M].  v: [ndim][M] or NULL => Maxwell-Boltzmann at initial_temperature drawn from Philox
 * at nds_begin() (per-walker stream).                                                           */
NDS_API int nds_set_state(nds_engine* e, const double* x, const double* v);

/* per-walker umbrella windows (new; the reference runs one window per process).
 * k, r0: [colvardim][M].  Requires cfg.umb_per_walker.                                          */
NDS_API int nds_set_walker_bias(nds_engine* e, const double* k, const double* r0);

/* reference points of a Path / Path5dto2d colvar (colvars/path.py:32-71: `ref`, `ref_ind`, `sigma`,
 * normally read from `colvar_file_name` with np.loadtxt).  ref: [n_ref][path_dim] row-major,
 * ref_ind: [n_ref]; path_dim = ndim for Path, 2 for Path5dto2d.  Must be called before nds_begin(). */
NDS_API int nds_set_path(nds_engine* e, int64_t n_ref, int32_t path_dim, const double* ref,
                         const double* ref_ind, double sigma);

/* external noise table for parity tests (NDS_NOISE_EXTERNAL): noise[step][j][w], j < normals per
 * step (ndim for langevin: xi; 2*ndim for 2nd-langevin: xi then eta, md.py:120,127-128).
 * Covers steps [first_step, first_step + nsteps).                                               */
NDS_API int nds_set_noise(nds_engine* e, const double* noise, int64_t first_step, int64_t nsteps);

/* replaces NDRun.begin (ndrun.py:182-206) + MD.begin (md.py:101-110): initial potential / bias
 * evaluation, velocities if unset, ke, T, frame 0.                                              */
NDS_API int nds_begin(nds_engine* e);

/* replaces the NDRun.run loop (ndrun.py:214-251) for nsteps more steps: fix.update (metadynamics
 * deposition schedule, gaus.py:130-146) -> engine.update (md.py:112-231 / committor.py:82-116)
 * -> step += 1, time += dt -> frames.  Stream-ordered; returns after enqueueing unless the
 * deposition exchange callback or the frame ring needs the host.                                */
NDS_API int nds_run(nds_engine* e, int64_t nsteps);
NDS_API int nds_sync(nds_engine* e);

/* current state -> host.  Any pointer may be NULL.  x, v, f (potential force, atoms.forces),
 * fixf (bias force, atoms.fixf): [ndim][M];  colv: [colvardim][M];
 * thermo: [5][M] = T, pe, ke, biase, totale (data.py:77-81, md.py:226-229).                     */
NDS_API int nds_get_state(nds_engine* e, double* x, double* v, double* f, double* fixf, double* colv,
                  double* thermo);
/* single-precision host buffers (same layouts): for fp32 engines the copies go straight between the
 * caller's buffers and the SoA state, half the PCIe bytes of the fp64 calls and no conversion pass. */
NDS_API int nds_set_state_f32(nds_engine* e, const float* x, const float* v);
NDS_API int nds_get_state_f32(nds_engine* e, float* x, float* v, float* f, float* fixf, float* colv,
                              float* thermo);
NDS_API int64_t nds_step(const nds_engine* e);   /* NDRun.step  */
NDS_API double nds_time(const nds_engine* e);    /* NDRun.time (accumulated in fp64 like ndrun.py:240) */

/* checkpoint / resume.  Philox noise is indexed by (global walker id, global step), so a run is fully
 * described by x, v (nds_get_state), the hill table + VR (nds_get_hills, nds_get_vr) and this clock.
 * nds_set_clock restores NDRun.step / NDRun.time (ndrun.py:199-200,232,240), the number of depositions
 * done (Gaussian._dep_count, gaus.py:107) and MD.rescale_count (md.py:77) on a begun engine.          */
NDS_API int nds_get_clock(const nds_engine* e, int64_t* step, double* time, int64_t* dep_count,
                          int64_t* rescale_count);
NDS_API int nds_set_clock(nds_engine* e, int64_t step, double time, int64_t dep_count,
                          int64_t rescale_count);
NDS_API int nds_set_vr(nds_engine* e, const double* vr);

/* walkers whose position or velocity is not finite (the reference only logs T > 10000, ndrun.py:253) */
NDS_API int64_t nds_count_nonfinite(nds_engine* e);

/* deterministic parity entry: replaces potential.compute(x) (potentials/potential.py:14),
 * colvar.compute / jacobian (colvars/colvar.py:16-20) and the sum of fix.compute(x0, col0)
 * (md.py:182-188) at n given configurations x[ndim][n], using the engine's current hill table
 * and uniform umbrella parameters.  Outputs (any may be NULL): V[n], F[ndim][n], colv[cd][n],
 * J[cd][ndim][n], Vb[n], Fb[ndim][n].                                                           */
NDS_API int nds_eval_at(nds_engine* e, const double* x, int64_t n, double* V, double* F, double* colv,
                double* J, double* Vb, double* Fb);

/* hill table (bias/gaus.py _history, _w).  Shared table: centers[n][cd], heights[n].
 * Per-walker tables (!mtd_shared): `walker` selects the table.                                  */
NDS_API int nds_get_hills(nds_engine* e, int64_t walker, int64_t* n, double* centers, double* heights,
                  int64_t capacity);
/* per-hill variances of an adaptive table (Gaussian._sigma_history; the sigma^2 of the config for a
 * fixed-sigma table, first colvar).  nds_set_hill_sigma2 follows nds_set_hills when a table is restored. */
NDS_API int nds_get_hill_sigma2(nds_engine* e, int64_t walker, double* sigma2, int64_t cap);
NDS_API int nds_set_hill_sigma2(nds_engine* e, int64_t walker, int64_t n, const double* sigma2);
/* whole hill storage in engine layout, as doubles (checkpoints of per-walker tables): n_hills * hill_stride
 * values for a shared table, n_hills * hill_stride * M otherwise (field j of record i of walker w at
 * (i * hill_stride + j) * M + w).  nds_set_hill_table_raw installs n records (n <= mtd_capacity).           */
NDS_API int64_t nds_hill_table_size(const nds_engine* e, int64_t n);
NDS_API int nds_get_hill_table_raw(nds_engine* e, double* out, int64_t cap);
NDS_API int nds_set_hill_table_raw(nds_engine* e, int64_t n, const double* in);
NDS_API int nds_set_hills(nds_engine* e, int64_t walker, int64_t n, const double* centers,
                  const double* heights);
/* stale bias energy per walker used by the well-tempered height (fix.VR): [M] */
NDS_API int nds_get_vr(nds_engine* e, double* vr);

/* multi-GPU shared-hill exchange.  When set, every deposition writes this rank's new hill records
 * into its slot of the table and then calls  cb(user, dev_ptr_table_segment, bytes_per_rank)  on
 * the engine's stream-ordered host path; the callback performs the allgather IN PLACE over the
 * segment [n_hills, n_hills + n_walkers_global) (e.g. torch.distributed / NCCL).               */
/* Calls made through the callback: (segment, bytes_per_rank > 0, rank_offset_bytes >= 0) = allgather in place;
 * (NULL, 0, 0) = barrier (peer-memory exchange); (buffer, bytes, -1) = all-reduce (sum) in place of
 * bytes / sizeof(real) reals of the engine's precision (grid increment of mtd_grid).                  */
typedef int (*nds_exchange_fn)(void* user, void* dev_segment, int64_t bytes_per_rank,
                               int64_t rank_offset_bytes);
NDS_API int nds_set_exchange(nds_engine* e, nds_exchange_fn cb, void* user);

/* peer-memory variant of the exchange (NVLink / NVSwitch): after nds_open_peer_tables() the deposition
 * kernel stores every new hill record directly into the table of EVERY rank (its own included), so no
 * allgather runs; the exchange callback is then invoked with bytes_per_rank == 0 and only has to act as
 * a barrier between the ranks.  handle: 64 bytes (cudaIpcMemHandle_t) of this engine's hill table;
 * handles: [nranks][64] gathered from all ranks of the node, indexed by rank.
 * nds_set_peer_tables() takes raw device pointers instead (engines living in one process).       */
NDS_API int nds_hills_ipc_handle(nds_engine* e, void* handle);
NDS_API int nds_open_peer_tables(nds_engine* e, const void* handles, int nranks, int rank);
NDS_API int nds_set_peer_tables(nds_engine* e, void* const* tables, int nranks, int rank);

/* device-side barrier of the peer-memory exchange: one warp per rank signals its arrival into every peer's
 * window and waits for theirs (no host round trip, no callback).  ONLY for one rank per GPU: ranks that share
 * a GPU must keep the callback barrier (their waiting kernels are not guaranteed to be co-resident).  A rank
 * that does not arrive within 10 s makes nds_sync() / nds_wait() fail instead of hanging the GPU.
 * With mtd_grid the device barrier is part of ONE exchange kernel per deposition stride: it stores the rank's
 * grid increment into its slot on every peer, signals, waits for the peers and adds the slots in rank order.  */
NDS_API int nds_set_peer_barrier(nds_engine* e, int on_device);

/* native NCCL exchange (SURVEY.md section 8b `nds_set_comm`): a C caller gets the multi-GPU hill exchange
 * without Python.  Either pass a ready ncclComm_t (nccl_comm != NULL; the caller keeps ownership), or the
 * 128-byte ncclUniqueId produced once by nds_nccl_unique_id() on one rank and handed to every rank by the
 * launcher (MPI_Bcast, torch.distributed.broadcast_object_list, a file...): the engine then creates and owns
 * its communicator.  Afterwards each deposition runs ONE in-place ncclAllGather of the new table segment
 * (exact hill sum) or ONE ncclAllReduce of the grid increment (mtd_grid) on the engine's stream; nds_run()
 * no longer synchronises with the host.  NCCL is resolved with dlopen("libnccl.so.2") at the first call
 * (the copy already loaded in the process wins; NDS_NCCL_LIB overrides the name).                     */
NDS_API int nds_nccl_unique_id(void* id128);
NDS_API int nds_set_comm(nds_engine* e, const void* unique_id, void* nccl_comm, int rank, int nranks);

/* device time spent in the exchange steps so far (CUDA events around every allgather / all-reduce /
 * peer-store + barrier) and their number                                                               */
NDS_API int nds_exchange_info(nds_engine* e, double* exchange_ms, int64_t* n_exchanges);
/* position-weighted 64-bit checksum of hill records [first, first + n) of this rank's table, computed on
 * the device: after an exchange every rank must report the same value for the same segment.           */
NDS_API int nds_table_checksum(nds_engine* e, int64_t first, int64_t n, uint64_t* sum);

/* mtd_grid: the accumulated bias grid, [ny][nx][4] doubles = (V, h0 dV/dc0, h1 dV/dc1, h0 h1 d2V/dc0dc1) at
 * the point lo + (i, j) * h (checkpoints, FES output); nds_mtd_grid_outside() = evaluations that fell
 * outside [lo, hi] so far (they see no bias: the box must cover the reachable colvar range).          */
NDS_API int nds_get_mtd_grid(nds_engine* e, double* out, int64_t cap);
NDS_API int nds_set_mtd_grid(nds_engine* e, const double* in, int64_t n);
NDS_API int64_t nds_mtd_grid_outside(nds_engine* e);

/* Monte Carlo / Wang-Landau bookkeeping of EVERY walker, for checkpoints (nds_get_mc / nds_get_wl_histogram serve one
 * quantity or one walker): accepted[M] (MC.accept_rate, mc.py:85); with a metadynamics bias time[M] and dep[M], each
 * walker's own NDRun.time and Gaussian._dep_count (ndrun.py:234-237: the clock advances on accepted moves only);
 * wl[101][M] = WangLandau.hE[0] of every walker (wang_landau.py:17).  Null pointers are skipped.  *max_dep = the
 * longest per-walker table; after the call nds_get_hill_table_raw covers that many records.                    */
NDS_API int nds_get_mc_state(nds_engine* e, int64_t* accepted, double* time, int64_t* dep, double* wl, int64_t* max_dep);
NDS_API int nds_set_mc_state(nds_engine* e, const int64_t* accepted, const double* time, const int64_t* dep,
                             const double* wl);

/* thermo[5][n_walkers] = (T, pe, ke, biase, totale) as nds_get_state returns them.  Under MC they follow the last
 * event, not the state (mc.py:93-95; gaus.py:146): a resumed run writes them back after nds_begin.          */
NDS_API int nds_set_thermo(nds_engine* e, const double* thermo);

/* Harmonic.VR of every restraint as of the last refresh, [n_umb][n_walkers] (umb.py:32-33; refreshed by begin() and by
 * every deposition, gaus.py:138-146).  Kept only when umbrella AND metadynamics biases record frames (it feeds the
 * fix_VR column of thermo.dat); *n = 0 otherwise.  For checkpoints.                                          */
NDS_API int nds_get_umb_vr(nds_engine* e, double* vr, int64_t cap, int64_t* n);
NDS_API int nds_set_umb_vr(nds_engine* e, const double* vr, int64_t n);

/* mtd_adapsig_hist: the colvar history window of every walker (Stat.colv, io/stat.py:70-72), for checkpoints:
 * ring[(j % window) * n_walkers + w] = sample j of walker w, *n_samples = len(stat.colv).  ring may be NULL to
 * query the sizes only.                                                                                    */
NDS_API int nds_get_cv_history(nds_engine* e, double* ring, int64_t cap_elems, int64_t* n_samples, int64_t* window);
NDS_API int nds_set_cv_history(nds_engine* e, const double* ring, int64_t n_samples);

/* asynchronous host path, one block per direction: host_xv = [2 ndim][M] (x rows, then v rows) and host_xvt =
 * [2 ndim + 5][M] (x, v, thermo) in the ENGINE's precision (float for NDS_F32), ideally pinned; either may be
 * NULL.  Enqueues H2D -> nsteps steps -> D2H and returns; nds_wait() completes the pass (the buffers must stay
 * untouched until then).  chunks > 1 pipelines the pass over walker chunks on separate copy / compute streams
 * (independent walkers only: no metadynamics, no committor compaction, nsteps <= steps_per_launch).     */
NDS_API int nds_submit(nds_engine* e, const void* host_xv, int64_t nsteps, void* host_xvt, int32_t chunks);
NDS_API int nds_wait(nds_engine* e);

/* committor outcome (engines/committor.py:103-111; ndrun.py:261-264): basin[M] (-1 = none),
 * exit_step[M] (value of NDRun.step when the walker stopped; == steps run when uncommitted).    */
NDS_API int nds_get_commit(nds_engine* e, int32_t* basin, int64_t* exit_step);
/* checkpoint / resume of a committor ensemble: alive[w] = 1 while walker w is still running.  nds_set_commit
 * follows nds_begin() (which revives every walker) on the restoring engine.                               */
NDS_API int nds_get_alive(nds_engine* e, uint8_t* alive);
NDS_API int nds_set_commit(nds_engine* e, const int32_t* basin, const int64_t* exit_step, const uint8_t* alive);
NDS_API int64_t nds_count_alive(nds_engine* e);

/* Metropolis MC (engines/mc.py:42-101): accepted moves per walker [M] (MC.accept_rate); a walker's
 * NDRun.time is accepted * dt because time only advances on acceptance (ndrun.py:235-238).
 * With NDS_NOISE_EXTERNAL the table of nds_set_noise holds UNIFORMS in [0,1):
 *   propose "single": [step][3][M] = (coordinate pick, displacement, acceptance);
 *   propose "all":    [step][ndim + 1][M] = (displacements..., acceptance).                       */
NDS_API int nds_get_mc(nds_engine* e, int64_t* accepted);
/* Wang-Landau: the energy histogram hE[NDS_WL_GRID] of one walker (WangLandau.hE[0]) */
NDS_API int nds_get_wl_histogram(nds_engine* e, int64_t walker, double* hE);
/* WangLandau.copy_hE (wang_landau.py:24-34, one basin): install a saved histogram; walker < 0: every walker.
 * nds_begin() clears the histograms, so call it afterwards.                                          */
NDS_API int nds_set_wl_histogram(nds_engine* e, int64_t walker, const double* hE);

/* frames recorded since the last drain (io/dump.py:76-102 content for walkers [0, dump_walkers)).
 * Row layout per frame and walker: step, time, x[ndim], v[ndim], f[ndim], colv[cd],
 * T, pe, ke, biase, totale, VR, then the lagged thermo sample (A16): time, T, pe, ke, biase, totale.
 * out: [n_frames][nds_frame_width()][dump_walkers].  out == NULL discards the pending frames (the
 * ring is rewound, nothing is copied).                                                          */
NDS_API int nds_frame_width(const nds_engine* e);
NDS_API int64_t nds_frames_pending(const nds_engine* e);
NDS_API int nds_drain_frames(nds_engine* e, double* out, int64_t max_frames, int64_t* n_frames);

/* free-energy estimate of the reference: sum of hills on a grid (bias/gaus.py:365-393 with the
 * grid of io/plot.py:517-519).  grid[ny][nx], X = x0 + i*dx, Y = y0 + j*dy.                     */
NDS_API int nds_fes_grid(nds_engine* e, int64_t walker, double x0, double dx, int64_t nx, double y0,
                 double dy, int64_t ny, double* grid);

/* the histogram accumulated inside the step kernels since nds_begin() (cfg.hist_n): counts[ny][nx]; clear != 0
 * zeroes it afterwards.  nds_frames_f32() = 1 when the device frame ring holds float rows (fp32 engines;
 * nds_drain_frames always delivers doubles).                                                            */
NDS_API int nds_get_histogram(nds_engine* e, int64_t* counts, int clear);
NDS_API int nds_frames_f32(const nds_engine* e);

/* on-device histogram of the current colvar values of all local walkers (io/plot.py:562-572).
 * counts[ny][nx] (ny = 1 for colvardim 1).                                                       */
NDS_API int nds_colvar_histogram(nds_engine* e, double x0, double dx, int64_t nx, double y0, double dy,
                         int64_t ny, int64_t* counts);

/* raw device pointers for zero-copy plumbing (torch / NCCL): positions, velocities (engine
 * precision, SoA [ndim][M]) and the hill table (records of (colvardim + 1) reals, padded to 4). */
NDS_API int nds_device_pointers(nds_engine* e, void** x, void** v, void** hills, int64_t* hill_stride);
NDS_API void* nds_stream(nds_engine* e);          /* cudaStream_t of the engine */
NDS_API int64_t nds_n_hills(const nds_engine* e, int64_t walker);

/* launches issued since creation (bench.py "gpu_launches") and duration of the last nds_run in
 * milliseconds measured with CUDA events on the engine's stream.                                */
NDS_API int64_t nds_launch_count(const nds_engine* e);
NDS_API double nds_last_run_ms(nds_engine* e);
/* name of the kernel variant selected for the hot loop ("md_block<f32,2,Mueller2d,...>") */
NDS_API const char* nds_kernel_name(const nds_engine* e);

/* Philox4x32-10 known-answer hook (tests): out[4] = philox(ctr[4], key[2]) computed ON DEVICE.  */
NDS_API int nds_philox_kat(int device, const uint32_t* ctr, const uint32_t* key, uint32_t* out);

/* Tail census of the Gaussian stream (tests): draws n_walkers * calls_per_walker * 6 normals of the per-step stream
 * of `seed` through the very function the step kernels call, in the given precision, and returns counts[t] =
 * #{|z| > thresholds[t]} (n_thresholds <= 8) and stats4 = {sum z, sum z^2, sum z^4, max |z|}.               */
NDS_API int nds_noise_tail_counts(int device, int precision, uint64_t seed, int64_t n_walkers,
                                  int64_t calls_per_walker, const double* thresholds, int n_thresholds,
                                  uint64_t* counts, double* stats4);

/* Map50to2d (colvars/misc.py:19-156): values c[2][n] (compute, :140-150) and Jacobians jac[2][50][n] (jacobian,
 * :152-155) of the two 50-dimensional quadratic forms at n configurations x[50][n], fp64, on `device`.  Either
 * output may be NULL.  No potential of the reference accepts ndim = 50 (potential.py:12 with mueller.py:59 /
 * misc.py:53), so this evaluation is everything the reference itself can do with the colvar.                  */
NDS_API int nds_map50to2d(int device, const double* x, int64_t n, double* c, double* jac);

#ifdef __cplusplus
}
#endif
#endif /* NDSB200_H */
