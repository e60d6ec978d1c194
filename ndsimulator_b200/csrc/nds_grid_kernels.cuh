// nds_grid_kernels.cuh -- grid-accumulated metadynamics bias (mtd_grid) and the peer-memory barrier.
//
// The reference keeps its free-energy estimate as a grid that is updated incrementally, hill by hill
// (Gaussian.projection / _dep_projection, bias/gaus.py:365-393).  Here the same idea serves the step loop: at
// every deposition the footprint of the NEW hills is added to a grid of (V, h0 V_0, h1 V_1, h0 h1 V_01); the
// step kernel interpolates (grid_interp, nds_physics.cuh), so the cost of a step no longer grows with the number
// of hills, and the multi-GPU exchange becomes a sum of grid increments instead of a record allgather.
#pragma once
#include <cooperative_groups.h>
#include "nds_kernels.cuh"

namespace nds {

constexpr int GRID_TX = 16, GRID_TY = 16;  // grid points per tile
constexpr int GRID_NT = 128;               // threads per CTA: four warps, each with the WHOLE tile and a quarter of the records
constexpr int NDS_MAX_RANKS = 16;          // ranks that can share a hill table / grid through peer memory
constexpr int GRID_SLICES = 8;             // record slices per tile = CTAs per cluster

// Footprint of n_new hill records (4-word records (c0, c1 | 0, w, log2 w), consecutive from `rec`) on the grid.
// A Gaussian hill is SEPARABLE: w exp(-d0^2 / 2 s0^2) exp(-d1^2 / 2 s1^2), so a tile's share of a record is the
// outer product of a 16-entry column profile and a 16-entry row profile -- 32 exponentials and 4 FMA per point
// instead of one exponential and ~10 operations per point; summed over the records this is a rank-K update
// (Ex | Ex d0)^T (Ey | Ey d1) of the tile's four fields on the FP32 pipe (packed FFMA2).
// Work split: CTA = (tile, slice of the records); warp = a quarter of the slice, with the whole 16 x 16 tile in
// registers (thread = 4 columns x 2 rows = 32 accumulators: 3 shared-memory loads feed 16 FFMA2 -- with one point
// pair per thread the kernel was bound by shared-memory wavefronts, 71 % of the pipe, profiles/r02t_gridacc_*).
// Everything inside a warp is warp-synchronous (no block barrier in the loops):
//   cull: the lanes test 32 records per pass against the tile rectangle grown by `cut` sigma (rounded corners;
//     exp(-cut^2 / 2) is below the rounding of the accumulator) and queue the survivors IN ORDER (ballot + prefix),
//     prescaled to the units of the exponent;
//   per chunk of queued records: lane l writes (E, E d) of column l (l < 16) or row l - 16 of every record of the
//     chunk to shared memory, then every lane adds the chunk to its 8 points.
// The order of every sum is fixed -- records in order inside a warp, warps 0..3 inside a CTA, slices 0..15 across
// the CLUSTER that the CTAs of a tile form (co-scheduled on one GPC): each CTA parks its sums in its own shared
// memory, and after a cluster barrier CTA r adds row r of the tile over the slices through distributed shared
// memory -- so results are reproducible and identical wherever they are recomputed.  Then
//   add != 0: grid[p] += sum (single rank / rebuilding a grid from a table);
//   add == 0: dst[r][p] = sum for r < ndst (the local increment buffer that grid_push_kernel / the all-reduce
//             sends on) -- every point is written, so the buffer never needs clearing.
// No partial grids in global memory, no second kernel.
template <typename R> struct GridAccCfg {
  static constexpr int QN = sizeof(R) == 4 ? 128 : 64;  // records culled per pass and warp (queue capacity)
  static constexpr int CH = sizeof(R) == 4 ? 16 : 8;    // queued records whose profiles are staged together
};

template <typename R>
__global__ void __launch_bounds__(GRID_NT)
grid_accumulate_kernel(const __grid_constant__ Consts<R> K, const R* __restrict__ rec, long long n_new,
                       R* const* __restrict__ dst, int ndst, R* grid, int add) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  using V4 = typename Real4<R>::type;
  constexpr bool F32 = sizeof(R) == 4;
  constexpr int QN = GridAccCfg<R>::QN, CH = GridAccCfg<R>::CH, NW = GRID_NT / 32;
  static_assert(GRID_TX == 16 && GRID_TY == 16 && GRID_NT == 128, "thread <-> point mapping below");
  static_assert(GRID_TY % GRID_SLICES == 0 && (GRID_TY / GRID_SLICES) * GRID_TX * 4 <= GRID_NT, "cluster rank r reduces rows r * RPR .. of the tile");
  __shared__ V4 queue[NW][QN];                       // survivors of the current pass, per warp
  __shared__ __align__(16) R prof[NW][CH][32][2];    // (E, E d): entries 0..15 columns, 16..31 rows
  __shared__ V4 cpart[GRID_TX * GRID_TY];            // sums of this CTA, point (x, y) at [y * GRID_TX + x]
  const int nx = K.grid_nx, ny = K.grid_ny;
  const int tiles_x = (nx + GRID_TX - 1) / GRID_TX;
  const int tile = blockIdx.x / GRID_SLICES, slice = (int)cluster.block_rank();  // cluster = consecutive CTAs
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xg = (lane & 3) * 4, yg = (lane >> 2) * 2;  // this thread's points: columns xg..xg+3, rows yg, yg+1
  const bool two = K.cd > 1;
  // exponent in log2 units: -(d0'^2 + d1'^2) + log2 w with d' = (X - c) * p,  p = sqrt(log2(e) / 2) / sigma
  const R p0 = (R)0.8493218002880191 * nds_sqrt(K.mtd_invsig2[0]);
  const R p1 = two ? (R)0.8493218002880191 * nds_sqrt(K.mtd_invsig2[1]) : (R)0;
  // profile duty of this lane: lanes 0..15 column `lane` of the tile, lanes 16..31 row `lane - 16`
  const bool is_x = lane < GRID_TX;
  const R coord = is_x ? (K.grid_lo[0] + (R)(tx * GRID_TX + lane) * K.grid_h[0]) * p0
                       : (two ? (K.grid_lo[1] + (R)(ty * GRID_TY + lane - GRID_TX) * K.grid_h[1]) * p1 : (R)0);
  const R cut = F32 ? (R)6.5 : (R)8.5;
  // tile centre and half extent (original units); distance to the rectangle in units of sigma
  const R cx = K.grid_lo[0] + ((R)(tx * GRID_TX) + (R)0.5 * (GRID_TX - 1)) * K.grid_h[0];
  const R cy = K.grid_lo[1] + ((R)(ty * GRID_TY) + (R)0.5 * (GRID_TY - 1)) * K.grid_h[1];
  const R hx = (R)0.5 * (GRID_TX - 1) * K.grid_h[0], hy = (R)0.5 * (GRID_TY - 1) * K.grid_h[1];
  const R is0 = nds_sqrt(K.mtd_invsig2[0]), is1 = two ? nds_sqrt(K.mtd_invsig2[1]) : (R)0;
  const R cut2 = cut * cut * (R)1.0001;
  // records of this warp: quarter `warp` of slice `slice`
  const long long per = (n_new + GRID_SLICES - 1) / GRID_SLICES, per4 = (per + NW - 1) / NW;
  const long long s0 = slice * per, s1 = s0 + per < n_new ? s0 + per : n_new;
  const long long k0 = s0 + warp * per4, k1 = k0 + per4 < s1 ? k0 + per4 : s1;
  // acc[c][r][f]: column xg + c, row yg + r, fields (V, V d0 | V d1, V d0 d1)
  R acc[4][2][4];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int f = 0; f < 4; f++) acc[c][r][f] = (R)0;
  V4* const q = queue[warp];
  for (long long base = k0; base < k1; base += QN) {
    // ordered compaction: survivors keep the order of the records
    int m = 0;
#pragma unroll 2
    for (int sub = 0; sub < QN; sub += 32) {
      const long long k = base + sub + lane;
      bool keep = false;
      V4 g;
      g.x = (R)0; g.y = (R)0; g.z = (R)0; g.w = (R)0;
      if (k < k1) {
        const V4 r = *reinterpret_cast<const V4*>(rec + k * 4);  // one 16 / 32-byte record
        const R h0 = r.x, h1 = r.y, w = r.z;
        R ux = h0 - cx, uy = h1 - cy;
        ux = ((ux < (R)0 ? -ux : ux) - hx) * is0;
        uy = ((uy < (R)0 ? -uy : uy) - hy) * is1;
        ux = ux > (R)0 ? ux : (R)0;
        uy = uy > (R)0 ? uy : (R)0;
        keep = w != (R)0 && ux * ux + uy * uy <= cut2;
        g.x = h0 * p0; g.y = h1 * p1; g.z = w;
        if constexpr (F32) g.w = r.w;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) q[m + __popc(bal & ((1u << lane) - 1u))] = g;
      m += __popc(bal);
    }
    __syncwarp();
    for (int cb = 0; cb < m; cb += CH) {
      const int mc = m - cb < CH ? m - cb : CH;
      // (a) profiles: all queue loads first, so that the shared-memory and MUFU latencies of the records overlap
      V4 g[CH];
#pragma unroll
      for (int r = 0; r < CH; r++) g[r] = q[cb + (r < mc ? r : 0)];
#pragma unroll
      for (int r = 0; r < CH; r++) {
        const R d = coord - (is_x ? g[r].x : g[r].y);
        R e;
        if constexpr (F32) e = ex2_approx((is_x ? g[r].w : (R)0) - d * d);
        else e = (is_x ? g[r].z : (R)1) * exp(-(d * d) * (R)0.6931471805599453);
        if (r < mc) {
          if constexpr (F32) *reinterpret_cast<float2*>(&prof[warp][r][lane][0]) = make_float2(e, e * d);
          else *reinterpret_cast<double2*>(&prof[warp][r][lane][0]) = make_double2(e, e * d);
        }
      }
      __syncwarp();
      // (b) rank-mc update of this thread's 4 x 2 points
      if constexpr (F32) {
#pragma unroll 4
        for (int r = 0; r < mc; r++) {
          const float4 Xa = *reinterpret_cast<const float4*>(&prof[warp][r][xg][0]);      // (E, E d0) of columns xg, xg + 1
          const float4 Xb = *reinterpret_cast<const float4*>(&prof[warp][r][xg + 2][0]);  // ... xg + 2, xg + 3
          const float4 Y = *reinterpret_cast<const float4*>(&prof[warp][r][16 + yg][0]);  // (E, E d1) of rows yg, yg + 1
          const float2 X[4] = {make_float2(Xa.x, Xa.y), make_float2(Xa.z, Xa.w), make_float2(Xb.x, Xb.y), make_float2(Xb.z, Xb.w)};
          const float2 Ye[2] = {make_float2(Y.x, Y.x), make_float2(Y.z, Y.z)};
          const float2 Yd[2] = {make_float2(Y.y, Y.y), make_float2(Y.w, Y.w)};
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
              float2 lo = make_float2(acc[c][rr][0], acc[c][rr][1]), hi = make_float2(acc[c][rr][2], acc[c][rr][3]);
              lo = __ffma2_rn(X[c], Ye[rr], lo);
              hi = __ffma2_rn(X[c], Yd[rr], hi);
              acc[c][rr][0] = lo.x; acc[c][rr][1] = lo.y; acc[c][rr][2] = hi.x; acc[c][rr][3] = hi.y;
            }
        }
      } else {
#pragma unroll 2
        for (int r = 0; r < mc; r++) {
          R ey[2], eyd[2];
#pragma unroll
          for (int rr = 0; rr < 2; rr++) { ey[rr] = prof[warp][r][16 + yg + rr][0]; eyd[rr] = prof[warp][r][16 + yg + rr][1]; }
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const R ex = prof[warp][r][xg + c][0], exd = prof[warp][r][xg + c][1];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
              acc[c][rr][0] += ex * ey[rr]; acc[c][rr][1] += exd * ey[rr];
              acc[c][rr][2] += ex * eyd[rr]; acc[c][rr][3] += exd * eyd[rr];
            }
          }
        }
      }
      __syncwarp();
    }
  }
  // back to (V, h0 V_0, h1 V_1, h0 h1 V_01):  dV/dc0 = -(c0 - h0)/sigma0^2 e = -2 ln2 p0 d0' e
  const R f0 = -(R)1.3862943611198906 * p0 * K.grid_h[0], f1 = -(R)1.3862943611198906 * p1 * K.grid_h[1];
  // the four warps add their sums to the CTA's in warp order
  for (int wq = 0; wq < NW; wq++) {
    if (warp == wq) {
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
          V4 v;
          v.x = acc[c][rr][0]; v.y = acc[c][rr][1] * f0; v.z = acc[c][rr][2] * f1; v.w = acc[c][rr][3] * (f0 * f1);
          V4* o = &cpart[(yg + rr) * GRID_TX + xg + c];
          if (wq > 0) { const V4 u = *o; v.x = u.x + v.x; v.y = u.y + v.y; v.z = u.z + v.z; v.w = u.w + v.w; }
          *o = v;
        }
    }
    __syncthreads();
  }
  cluster.sync();
  // CTA `slice` adds rows slice * RPR .. of the tile over all slices (ascending): thread = (row, point, field)
  constexpr int RPR = GRID_TY / GRID_SLICES;
  if (threadIdx.x < RPR * GRID_TX * 4) {
    const int row = slice * RPR + threadIdx.x / (GRID_TX * 4), col = (threadIdx.x >> 2) % GRID_TX, fld = threadIdx.x & 3;
    const int i = tx * GRID_TX + col, j = ty * GRID_TY + row;
    const R* mine = reinterpret_cast<const R*>(&cpart[row * GRID_TX + col]) + fld;
    R v = (R)0;
#pragma unroll
    for (int s2 = 0; s2 < GRID_SLICES; s2++) v += *cluster.map_shared_rank(mine, s2);
    if (i < nx && j < ny) {
      const long long p = ((long long)j * nx + i) * 4 + fld;
      if (add) grid[p] += v;
      else {
        for (int r = 0; r < ndst; r++) dst[r][p] = v;
      }
    }
  }
  cluster.sync();  // nobody leaves while its shared memory may still be read
}

// This rank's increment -> its slot on every peer (stores over NVLink), float4 / double2 per thread.  Kept apart from
// grid_accumulate_kernel: system-scope fences inside the cluster kernel stall whole clusters (2 GPUs: 104 vs 95 us
// per stride, r02z).
template <typename R>
__global__ void grid_push_kernel(long long n16, const R* __restrict__ inc, R* const* __restrict__ dst, int ndst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) {
    const int4 v = reinterpret_cast<const int4*>(inc)[i];
    for (int r = 0; r < ndst; r++) reinterpret_cast<int4*>(dst[r])[i] = v;
  }
  // peer slots visible before the barrier kernel signals: ONE system fence per CTA behind the block barrier (fences
  // are cumulative); a fence per thread costs 13 us here (2 GPUs, NDS_EXCHANGE_TRACE)
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
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

// ------------------------------------------------------------------- fused grid exchange (peer path, device barrier)
// push + barrier + apply of one deposition stride in ONE kernel instead of three launches:
//   (1) every CTA stores its part of this rank's increment into the rank's slot on EVERY peer (NVLink), fences;
//   (2) the LAST CTA to get there (ticket counter) signals `epoch` into every peer's flag array;
//   (3) every CTA waits until all peers' epochs have arrived in THIS rank's flag array, then adds the nranks slots of
//       its part to the grid in rank order (bit-identical grids on every rank).
// All CTAs are co-resident (the host launches this kernel only when the grid has at most one CTA per SM), so the
// waiting CTAs cannot keep a pushing CTA off the machine.  NEVER launched for ranks that share one GPU.
template <typename R>
__global__ void __launch_bounds__(256)
grid_exchange_kernel(long long n16, const R* __restrict__ inc, R* const* __restrict__ dst, const R* const* __restrict__ src,
                     int nranks, R* grid, unsigned long long* const* __restrict__ flags, int rank,
                     unsigned long long epoch, unsigned long long timeout_ns, unsigned long long* err,
                     unsigned int* ticket, unsigned long long* dbg) {
  // dbg (NDS_EXCHANGE_TRACE=1): sums of globaltimer deltas from kernel entry, seen by CTA 0 -- [0] own push fenced,
  // [1] this rank's signal sent (all its CTAs fenced), [2] all peers' signals seen, [3] slots added; [5] launches
  unsigned long long dt0 = 0;
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dt0));
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) {
    const int4 v = reinterpret_cast<const int4*>(inc)[i];
    for (int r = 0; r < nranks; r++) reinterpret_cast<int4*>(dst[r])[i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // ONE system fence per CTA behind the block barrier (fences are cumulative): a fence per thread made this phase
    // 13 us on 2 GPUs (NDS_EXCHANGE_TRACE)
    __threadfence_system();
    unsigned long long dt1 = 0;
    if (dbg && blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dt1));
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {  // every CTA of this rank has pushed and fenced
      *ticket = 0u;            // next launch (nobody reads it again in this one)
      __threadfence_system();  // orders the other CTAs' fenced stores (seen through the ticket) before the signals
      for (int r = 0; r < nranks; r++) {
        unsigned long long* out = flags[r] + rank;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(out), "l"(epoch) : "memory");
      }
      if (dbg) { unsigned long long ts; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts)); dbg[4] = ts; }
    }
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < nranks; r++) {
      const unsigned long long* in = flags[rank] + r;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(in) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { atomicMax(err, epoch); break; }
        __nanosleep(100);
      }
    }
    __threadfence_system();
    if (dbg && blockIdx.x == 0) {
      unsigned long long dt3;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dt3));
      const unsigned long long ts = *reinterpret_cast<volatile unsigned long long*>(dbg + 4);
      atomicAdd(dbg + 0, dt1 - dt0); atomicAdd(dbg + 1, ts > dt0 ? ts - dt0 : 0ull); atomicAdd(dbg + 2, dt3 - dt0);
      atomicAdd(dbg + 5, 1ull);
    }
  }
  __syncthreads();
  if (i < n16) {
    constexpr int NV = 16 / (int)sizeof(R);
    union V { int4 q; R a[NV]; };
    V g;
    g.q = reinterpret_cast<const int4*>(grid)[i];
    for (int r = 0; r < nranks; r++) {
      V sv;
      sv.q = __ldcg(reinterpret_cast<const int4*>(src[r]) + i);  // written by the peers: straight from L2
#pragma unroll
      for (int k = 0; k < NV; k++) g.a[k] += sv.a[k];
    }
    reinterpret_cast<int4*>(grid)[i] = g.q;
  }
  if (dbg && blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long dt4;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dt4));
      atomicAdd(dbg + 3, dt4 - dt0);
    }
  }
}

}  // namespace nds
