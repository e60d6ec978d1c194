// nds_variants.h -- registry of compiled md_block variants.
//
// A variant is one instantiation of md_block<Spec>; fields equal to DYN (-1) are resolved at run time
// from the constant block.  The engine picks the matching variant with the most compile-time fields.
#pragma once
#include "nds_kernels.cuh"

namespace nds {

struct VariantKey {
  int prec, nd, pot, cv, tcv, integ, bias, method, noise, dump, nbasin;
};

struct Variant {
  VariantKey key;
  const char* name;
  void* launch;  // BlockLaunchFn<R>
  bool wide;     // md_block_wide: may carry the bias grid in shared memory
};

template <typename R>
using BlockLaunchFn = cudaError_t (*)(const Consts<R>&, const BlockArgs<R>&, cudaStream_t);

template <class S>
cudaError_t launch_md_block(const Consts<typename S::R>& K, const BlockArgs<typename S::R>& A,
                            cudaStream_t st) {
  const long long n = (A.n_active > 0 ? A.n_active : A.M) - A.slot0;
  const unsigned grid = (unsigned)((n + NDS_BLOCK - 1) / NDS_BLOCK);
  md_block<S><<<grid, NDS_BLOCK, (size_t)A.dyn_smem, st>>>(K, A);
  return cudaGetLastError();
}

// wide CTAs (md_block_wide): used by the mtd_grid variants, whose launch carries the shared-memory copy of the grid
template <class S>
cudaError_t launch_md_block_wide(const Consts<typename S::R>& K, const BlockArgs<typename S::R>& A,
                                 cudaStream_t st) {
  const long long n = (A.n_active > 0 ? A.n_active : A.M) - A.slot0;
  const size_t smem = (size_t)A.dyn_smem;
  // few walkers: many small CTAs (latency-bound, one warp per scheduler); many walkers: one wide CTA per SM
  const int threads = n >= 148LL * 512 ? 512 : (n >= 148LL * 256 ? 256 : 128);
  if (smem > 48 * 1024) {  // per device, cheap: set on every launch that needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(md_block_wide<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = (unsigned)((n + threads - 1) / threads);
  md_block_wide<S><<<grid, threads, smem, st>>>(K, A);
  return cudaGetLastError();
}

constexpr int PREC_F32_PACKED = 2;  // internal: fp32 with two walkers per thread (real type f2)
template <typename R> struct prec_of { static constexpr int value = NDS_F64; };
template <> struct prec_of<float> { static constexpr int value = NDS_F32; };
template <> struct prec_of<f2> { static constexpr int value = PREC_F32_PACKED; };

template <class S>
Variant make_variant(const char* name) {
  return Variant{VariantKey{prec_of<typename S::R>::value, S::ND, S::POT, S::CV, S::TCV, S::INTEG,
                            S::BIAS, S::METHOD, S::NOISE, S::DUMP, S::NBASIN},
                 name, (void*)&launch_md_block<S>, false};
}

template <class S>
Variant make_variant_wide(const char* name) {
  return Variant{VariantKey{prec_of<typename S::R>::value, S::ND, S::POT, S::CV, S::TCV, S::INTEG,
                            S::BIAS, S::METHOD, S::NOISE, S::DUMP, S::NBASIN},
                 name, (void*)&launch_md_block_wide<S>, true};
}

// shared-hill metadynamics kernels (nds_mtd_kernel.cuh): W walkers per warp, NW warps per CTA
struct MtdVariant {
  VariantKey key;
  int W, NW;
  const char* name;
  void* launch;
};
const MtdVariant* mtd_variants_f32_nd1(int* n);
const MtdVariant* mtd_variants_f64_nd1(int* n);
const MtdVariant* mtd_variants_f32_nd2(int* n);
const MtdVariant* mtd_variants_f32_nd5(int* n);
const MtdVariant* mtd_variants_f64_nd2(int* n);
const MtdVariant* mtd_variants_f64_nd5(int* n);

// one table per translation unit
const Variant* variants_generic_f32_nd1(int* n);
const Variant* variants_generic_f32_nd2(int* n);
const Variant* variants_generic_f32_nd5(int* n);
const Variant* variants_generic_f64_nd1(int* n);
const Variant* variants_generic_f64_nd2(int* n);
const Variant* variants_generic_f64_nd5(int* n);
const Variant* variants_hot(int* n);
const Variant* variants_grid(int* n);

}  // namespace nds
