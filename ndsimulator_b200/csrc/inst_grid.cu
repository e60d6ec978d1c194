// inst_grid.cu -- md_block_wide variants for the grid-accumulated metadynamics bias (mtd_grid): the bias mask is a
// compile-time constant (metadynamics from the grid, nothing else), CTAs of up to 512 threads share one
// shared-memory copy of the grid.  Other bias combinations (umb + mtd, pe + mtd) run the generic md_block and read
// the grid from global memory.
#include "nds_variants.h"
#include "nds_pair_kernel.cuh"

namespace nds {
namespace {
constexpr int GB = BIAS_MTD | BIAS_MTD_GRID;
constexpr int MD = NDS_METHOD_MD, PHX = NDS_NOISE_PHILOX;
}  // namespace

const Variant* variants_grid(int* n) {
  static const Variant table[] = {
      // C5 with mtd_grid: 5-D Mueller + Proj5dto2d, Langevin, Philox noise
      // (two lanes per walker up to NDS_PAIR_MAX walkers per launch, nds_pair_kernel.cuh; md_block_wide above that)
      make_variant_pair<Spec<float, 5, NDS_POT_MUELLER5D, NDS_CV_PROJ5DTO2D, NDS_CV_PROJ5DTO2D, NDS_INT_LANGEVIN, GB, MD, PHX, 0>>(
          "md_pair_grid|md_block_wide<f32,nd5,Mueller5d,Proj5dto2d,langevin,mtd-grid>"),
      // examples/2d-mtd.yaml with mtd_grid
      make_variant_wide<Spec<float, 2, NDS_POT_MUELLER2D, NDS_CV_ORIGINAL, NDS_CV_ORIGINAL, NDS_INT_LANGEVIN, GB, MD, PHX, 0>>(
          "md_block_wide<f32,nd2,Mueller2d,Original,langevin,mtd-grid>"),
      // generic: any potential / colvar, first-order Langevin, no frames
      make_variant_wide<Spec<float, 1, DYN, DYN, DYN, NDS_INT_LANGEVIN, GB, MD, DYN, 0>>("md_block_wide<f32,nd1,generic,langevin,mtd-grid>"),
      make_variant_wide<Spec<float, 2, DYN, DYN, DYN, NDS_INT_LANGEVIN, GB, MD, DYN, 0>>("md_block_wide<f32,nd2,generic,langevin,mtd-grid>"),
      make_variant_wide<Spec<float, 5, DYN, DYN, DYN, NDS_INT_LANGEVIN, GB, MD, DYN, 0>>("md_block_wide<f32,nd5,generic,langevin,mtd-grid>"),
      make_variant_wide<Spec<double, 1, DYN, DYN, DYN, NDS_INT_LANGEVIN, GB, MD, DYN, 0>>("md_block_wide<f64,nd1,generic,langevin,mtd-grid>"),
      make_variant_wide<Spec<double, 2, DYN, DYN, DYN, NDS_INT_LANGEVIN, GB, MD, DYN, 0>>("md_block_wide<f64,nd2,generic,langevin,mtd-grid>"),
      make_variant_wide<Spec<double, 5, DYN, DYN, DYN, NDS_INT_LANGEVIN, GB, MD, DYN, 0>>("md_block_wide<f64,nd5,generic,langevin,mtd-grid>"),
  };
  *n = (int)(sizeof(table) / sizeof(table[0]));
  return table;
}
}  // namespace nds
