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

// Footprint of n_new hill records (4-word records (c0, c1 | 0, w, log2 w), consecutive from `rec`) on the grid.
// CTA (tile, slice): a tile of GRID_TX x GRID_TY points, one thread per point, and one slice of the records.
//   phase 1 (ordered compaction): the threads test their slice's records against the tile box grown by `cut` sigma
//     -- one record per thread per pass, exp(-cut^2 / 2) is below the rounding of the accumulator -- and queue the
//     survivors (warp ballots + warp prefix), prescaled to the units of the exponent, in shared memory;
//   phase 2: every thread sums the queued records at its point: 2 FADD + 2 FFMA + MUFU.EX2 + 5 accumulate ops.
// Partial sums of slice s go to part[s][point][4] (plain stores: the order of every sum is fixed, so results are
// reproducible and identical wherever they are recomputed).  grid_reduce_kernel adds the slices.
constexpr int GRID_QUEUE = 512;  // survivors queued per pass

template <typename R>
__global__ void __launch_bounds__(GRID_TX * GRID_TY)
grid_accumulate_kernel(const __grid_constant__ Consts<R> K, const R* __restrict__ rec, long long n_new,
                       int n_slices, R* __restrict__ part) {
  using V4 = typename Real4<R>::type;
  constexpr int NT = GRID_TX * GRID_TY;
  __shared__ V4 queue[GRID_QUEUE];
  const int nx = K.grid_nx, ny = K.grid_ny;
  const int tiles_x = (nx + GRID_TX - 1) / GRID_TX;
  const int n_tiles = tiles_x * ((ny + GRID_TY - 1) / GRID_TY);
  const int tile = blockIdx.x % n_tiles, slice = blockIdx.x / n_tiles;
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int i = tx * GRID_TX + (threadIdx.x % GRID_TX), j = ty * GRID_TY + (threadIdx.x / GRID_TX);
  const bool two = K.cd > 1;
  // exponent in log2 units: -(d0'^2 + d1'^2) + log2 w with d' = (X - c) * p,  p = sqrt(log2(e) / 2) / sigma
  const R p0 = (R)0.8493218002880191 * nds_sqrt(K.mtd_invsig2[0]);
  const R p1 = two ? (R)0.8493218002880191 * nds_sqrt(K.mtd_invsig2[1]) : (R)0;
  const R X = (K.grid_lo[0] + (R)i * K.grid_h[0]) * p0;
  const R Y = two ? (K.grid_lo[1] + (R)j * K.grid_h[1]) * p1 : (R)0;
  const R cut = sizeof(R) == 4 ? (R)6.5 : (R)8.5;
  // tile centre and half extent (+ cut-off), original units
  const R cx = K.grid_lo[0] + ((R)(tx * GRID_TX) + (R)0.5 * (GRID_TX - 1)) * K.grid_h[0];
  const R cy = K.grid_lo[1] + ((R)(ty * GRID_TY) + (R)0.5 * (GRID_TY - 1)) * K.grid_h[1];
  const R rx = (R)0.5 * (GRID_TX - 1) * K.grid_h[0] + cut * nds_sqrt(K.mtd_sig2[0]);
  const R ry = two ? (R)0.5 * (GRID_TY - 1) * K.grid_h[1] + cut * nds_sqrt(K.mtd_sig2[1]) : (R)1e30;
  const long long per = (n_new + n_slices - 1) / n_slices;
  const long long k0 = slice * per, k1 = k0 + per < n_new ? k0 + per : n_new;
  R a0 = (R)0, a1 = (R)0, a2 = (R)0, a3 = (R)0;
  __shared__ int warp_count[NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = k0; base < k1; base += GRID_QUEUE) {
    // ordered compaction: survivors keep the order of the records, so every sum below has a fixed order
    int m = 0;
    for (long long sub = base; sub < k1 && sub < base + GRID_QUEUE; sub += NT) {
      const long long k = sub + threadIdx.x;
      bool keep = false;
      V4 q;
      q.x = (R)0; q.y = (R)0; q.z = (R)0; q.w = (R)0;
      if (k < k1) {
        const R* r = rec + k * 4;
        const R h0 = r[0], h1 = r[1], w = r[2];
        const R ux = h0 - cx, uy = h1 - cy;
        keep = w != (R)0 && (ux < (R)0 ? -ux : ux) <= rx && (uy < (R)0 ? -uy : uy) <= ry;
        q.x = h0 * p0; q.y = h1 * p1; q.z = w;
        if constexpr (sizeof(R) == 4) q.w = r[3]; else q.w = (R)0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) warp_count[warp] = __popc(bal);
      __syncthreads();
      int off = m;
#pragma unroll
      for (int wv = 0; wv < NT / 32; wv++) {
        const int c = warp_count[wv];
        if (wv < warp) off += c;
        m += c;
      }
      if (keep) queue[off + __popc(bal & ((1u << lane) - 1u))] = q;
      __syncthreads();
    }
    for (int t = 0; t < m; t++) {
      const V4 g = queue[t];
      const R d0 = X - g.x, d1 = Y - g.y;
      R e;
      if constexpr (sizeof(R) == 4) e = ex2_approx(g.w - (d0 * d0 + d1 * d1));
      else e = g.z * exp(-(d0 * d0 + d1 * d1) * (R)0.6931471805599453);
      const R e0 = e * d0;
      a0 += e; a1 += e0; a2 += e * d1; a3 += e0 * d1;
    }
    __syncthreads();
  }
  if (i >= nx || j >= ny) return;
  // back to (V, h0 V_0, h1 V_1, h0 h1 V_01):  dV/dc0 = -(c0 - h0)/sigma0^2 e = -2 ln2 p0 d0' e
  const R f0 = -(R)1.3862943611198906 * p0 * K.grid_h[0], f1 = -(R)1.3862943611198906 * p1 * K.grid_h[1];
  R* o = part + ((long long)slice * nx * ny + (long long)j * nx + i) * 4;
  o[0] = a0; o[1] = a1 * f0; o[2] = a2 * f1; o[3] = a3 * (f0 * f1);
}

// Adds the slices in ascending order.  add != 0: grid[p] += sum (single rank / rebuilding a grid from a table);
// add == 0: dst[r][p] = sum for r < ndst (this rank's increment slot on every peer over NVLink, or the local
// all-reduce buffer) -- every point is written, so slots never need clearing.
template <typename R>
__global__ void grid_reduce_kernel(long long n_pts4, int n_slices, const R* __restrict__ part, R* const* __restrict__ dst,
                                   int ndst, R* grid, int add) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pts4) return;
  R v = (R)0;
  for (int s2 = 0; s2 < n_slices; s2++) v += part[(long long)s2 * n_pts4 + p];
  if (add) { grid[p] += v; return; }
  for (int r = 0; r < ndst; r++) dst[r][p] = v;
  if (ndst > 1) __threadfence_system();  // peer slots: visible before the barrier kernel signals (measured: 2 us of a 109 us stride)
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
