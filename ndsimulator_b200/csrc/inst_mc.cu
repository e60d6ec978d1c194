// inst_mc.cu -- mc_block instantiations (fp32 / fp64, ndim 1 / 2 / 5).
#include "nds_mc_kernel.cuh"

namespace nds {
template <typename R>
cudaError_t launch_mc_block(int nd, const Consts<R>& K, const BlockArgs<R>& A, long long* accepted,
                            R* wl_hE, double* mc_time, long long* mc_dep, R* hills_rw, cudaStream_t st) {
  const unsigned grid = (unsigned)((A.M + NDS_BLOCK - 1) / NDS_BLOCK);
  if (nd == 1) mc_block<R, 1><<<grid, NDS_BLOCK, 0, st>>>(K, A, accepted, wl_hE, mc_time, mc_dep, hills_rw);
  else if (nd == 2) mc_block<R, 2><<<grid, NDS_BLOCK, 0, st>>>(K, A, accepted, wl_hE, mc_time, mc_dep, hills_rw);
  else mc_block<R, 5><<<grid, NDS_BLOCK, 0, st>>>(K, A, accepted, wl_hE, mc_time, mc_dep, hills_rw);
  return cudaGetLastError();
}
template cudaError_t launch_mc_block<float>(int, const Consts<float>&, const BlockArgs<float>&, long long*, float*, double*, long long*, float*, cudaStream_t);
template cudaError_t launch_mc_block<double>(int, const Consts<double>&, const BlockArgs<double>&, long long*, double*, double*, long long*, double*, cudaStream_t);
}  // namespace nds
