// nds_grid_kernels.cuh -- grid-accumulated metadynamics bias (mtd_grid) and the peer-memory barrier.
//
// The reference keeps its free-energy estimate as a grid that is updated incrementally, hill by hill
// (Gaussian.projection / _dep_projection, bias/gaus.py:365-393).  Here the same idea serves the step loop: at
// every deposition the footprint of the NEW hills is added to a grid of (V, h0 V_0, h1 V_1, h0 h1 V_01); the
// step kernel interpolates (grid_interp, nds_physics.cuh), so the cost of a step no longer grows with the number
// of hills, and the multi-GPU exchange becomes a sum of grid increments instead of a record allgather.
#pragma once
#include "nds_kernels.cuh"

namespace nds {

constexpr int GRID_TX = 16, GRID_TY = 8;  // grid points per CTA tile (one thread per point)
constexpr int NDS_MAX_RANKS = 16;         // ranks that can share a hill table / grid through peer memory

// Footprint of n_new hill records (4-word records (c0, c1 | 0, w, log2 w), n_new consecutive records starting at
// `rec`) on the tile of this CTA.  A record whose centre is further than `cut` sigma from the tile is skipped by
// the whole CTA (uniform branch); exp(-cut^2 / 2) is below the rounding of the accumulator.
//   add != 0: grid[p] += footprint           (single rank, and rebuilding a grid from a hill table)
//   add == 0: dst[r][p] = footprint, r < ndst (the increment slot of this rank on every peer, or the local
//             all-reduce buffer) -- every point of the slot is written, so slots never need clearing.
template <typename R>
__global__ void __launch_bounds__(GRID_TX * GRID_TY)
grid_accumulate_kernel(const __grid_constant__ Consts<R> K, const R* __restrict__ rec, long long n_new,
                       R* const* __restrict__ dst, int ndst, R* grid, int add) {
  using V4 = typename Real4<R>::type;
  __shared__ V4 stage[GRID_TX * GRID_TY];
  const int nx = K.grid_nx, ny = K.grid_ny;
  const int tiles_x = (nx + GRID_TX - 1) / GRID_TX;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int i = tx * GRID_TX + (threadIdx.x % GRID_TX), j = ty * GRID_TY + (threadIdx.x / GRID_TX);
  const bool inside = i < nx && j < ny;
  const R X = K.grid_lo[0] + (R)i * K.grid_h[0];
  const R Y = K.grid_lo[1] + (R)j * K.grid_h[1];
  const R is0 = K.mtd_invsig2[0], is1 = K.cd > 1 ? K.mtd_invsig2[1] : (R)0;
  const R cut = sizeof(R) == 4 ? (R)6.5 : (R)8.5;
  // tile centre and half extent (+ cut-off) per colvar
  const R cx = K.grid_lo[0] + ((R)(tx * GRID_TX) + (R)0.5 * (GRID_TX - 1)) * K.grid_h[0];
  const R cy = K.grid_lo[1] + ((R)(ty * GRID_TY) + (R)0.5 * (GRID_TY - 1)) * K.grid_h[1];
  const R rx = (R)0.5 * (GRID_TX - 1) * K.grid_h[0] + cut * nds_sqrt(K.mtd_sig2[0]);
  const R ry = K.cd > 1 ? (R)0.5 * (GRID_TY - 1) * K.grid_h[1] + cut * nds_sqrt(K.mtd_sig2[1]) : (R)1e30;
  R a0 = (R)0, a1 = (R)0, a2 = (R)0, a3 = (R)0;
  for (long long base = 0; base < n_new; base += GRID_TX * GRID_TY) {
    const long long k = base + threadIdx.x;
    V4 h;
    h.x = (R)0; h.y = (R)0; h.z = (R)0; h.w = (R)0;
    if (k < n_new) {
      const R* r = rec + k * 4;
      h.x = r[0]; h.y = r[1]; h.z = r[2]; h.w = r[3];
    }
    __syncthreads();
    stage[threadIdx.x] = h;
    __syncthreads();
    const int m = n_new - base < GRID_TX * GRID_TY ? (int)(n_new - base) : GRID_TX * GRID_TY;
    for (int q = 0; q < m; q++) {
      const V4 g = stage[q];
      const R ux = g.x - cx, uy = g.y - cy;
      if ((ux < (R)0 ? -ux : ux) > rx || (uy < (R)0 ? -uy : uy) > ry || g.z == (R)0) continue;  // uniform
      const R d0 = X - g.x, d1 = K.cd > 1 ? Y - g.y : (R)0;
      const R s0 = d0 * is0, s1 = d1 * is1;
      const R e = g.z * nds_exp(-(R)0.5 * (d0 * s0 + d1 * s1));  // gaus.py:240-250
      a0 += e;
      a1 -= e * s0;
      a2 -= e * s1;
      a3 += e * (s0 * s1);
    }
  }
  if (!inside) return;
  a1 *= K.grid_h[0]; a2 *= K.grid_h[1]; a3 *= K.grid_h[0] * K.grid_h[1];
  const long long p = ((long long)j * nx + i) * 4;
  if (add) {
    grid[p + 0] += a0; grid[p + 1] += a1; grid[p + 2] += a2; grid[p + 3] += a3;
  } else {
    for (int r = 0; r < ndst; r++) {
      R* o = dst[r] + p;
      if constexpr (sizeof(R) == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(a0, a1, a2, a3);
      } else {
        reinterpret_cast<double2*>(o)[0] = make_double2(a0, a1);
        reinterpret_cast<double2*>(o)[1] = make_double2(a2, a3);
      }
    }
    if (ndst > 1) __threadfence_system();  // peer slots: visible before the barrier kernel signals
  }
}

// grid += sum of nsrc increment buffers, in the given (rank) order: every rank ends with bit-identical grids
template <typename R>
__global__ void grid_apply_kernel(long long n, R* grid, const R* const* __restrict__ src, int nsrc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  R g = grid[i];
  for (int r = 0; r < nsrc; r++) g += src[r][i];
  grid[i] = g;
}

// Bias energy from the grid at the points of a regular output grid (nds_fes_grid in mtd_grid mode)
template <typename R>
__global__ void grid_sample_kernel(const __grid_constant__ Consts<R> K, const R* __restrict__ grid, double x0,
                                   double dx, long long nx, double y0, double dy, long long ny, double* out) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nx * ny) return;
  R c[2] = {(R)(x0 + (p % nx) * dx), (R)(y0 + (p / nx) * dy)};
  R V, g[2];
  grid_interp<R, 2, false>(K, K.cd, grid, c, V, g, nullptr);
  out[p] = (double)V;
}

// ------------------------------------------------------------------------ peer-memory barrier (one GPU per rank)
// Lane r signals "rank `rank` has reached `epoch`" into the flag array of peer r (a store over NVLink, after a
// system-scope fence that orders it behind this rank's earlier peer stores), then waits until peer r's signal has
// arrived in this rank's own flag array.  One warp, no host involvement.  A peer that never arrives makes the wait
// give up after `timeout_ns` and raises *err (checked by nds_sync), so a lost rank cannot hang the GPU.
// NEVER launched for ranks that share one GPU (co-residency of the waiting kernels is not guaranteed there).
__global__ void peer_barrier_kernel(unsigned long long* const* __restrict__ flags, int rank, int nranks,
                                    unsigned long long epoch, unsigned long long timeout_ns,
                                    unsigned long long* err) {
  const int r = threadIdx.x;
  if (r >= nranks) return;
  __threadfence_system();
  unsigned long long* out = flags[r] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(out), "l"(epoch) : "memory");
  const unsigned long long* in = flags[rank] + r;
  unsigned long long t0, t1, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(in) : "memory");
    if (v >= epoch) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) { atomicMax(err, epoch); break; }
    __nanosleep(200);
  }
  __threadfence_system();
}

}  // namespace nds
