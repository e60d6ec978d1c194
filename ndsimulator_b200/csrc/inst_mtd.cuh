// inst_mtd.cuh -- mtd_block variants of one (precision, ndim).  Included by inst_mtd_<prec>_nd<N>.cu.
#include "nds_variants.h"
#include "nds_mtd_kernel.cuh"

namespace nds {
namespace {
constexpr int HS = 2;   // warps per walker group (hill split)
template <class S, int W, int NW = 8>
MtdVariant mk(const char* name) {
  return MtdVariant{VariantKey{prec_of<typename S::R>::value, S::ND, S::POT, S::CV, S::TCV, S::INTEG,
                               S::BIAS, S::METHOD, S::NOISE, S::DUMP},
                    W, NW, name, (void*)&launch_mtd_block<S, W, NW, HS>};
}
using GenL = Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_LANGEVIN>;
using GenL2 = Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_LANGEVIN2>;
}  // namespace

const MtdVariant* NDS_FN(int* n) {
  static const MtdVariant table[] = {
      mk<GenL, 4>("mtd_block<" NDS_TAG ",generic,langevin,W4>"),
      mk<GenL, 7>("mtd_block<" NDS_TAG ",generic,langevin,W7>"),
      mk<GenL, 8>("mtd_block<" NDS_TAG ",generic,langevin,W8>"),
      mk<GenL, 8, 7>("mtd_block<" NDS_TAG ",generic,langevin,W8,NW7>"),
      mk<GenL2, 4>("mtd_block<" NDS_TAG ",generic,2nd-langevin,W4>"),
      mk<GenL2, 7>("mtd_block<" NDS_TAG ",generic,2nd-langevin,W7>"),
      mk<GenL2, 8>("mtd_block<" NDS_TAG ",generic,2nd-langevin,W8>"),
      mk<GenL2, 8, 7>("mtd_block<" NDS_TAG ",generic,2nd-langevin,W8,NW7>"),
#ifdef NDS_HOT_C5
      // C5: 5-D Mueller + Proj5dto2d, Langevin, shared hills, Philox noise -- everything compile-time
      mk<Spec<float, 5, NDS_POT_MUELLER5D, NDS_CV_PROJ5DTO2D, NDS_CV_PROJ5DTO2D, NDS_INT_LANGEVIN, BIAS_MTD,
              NDS_METHOD_MD, NDS_NOISE_PHILOX, 0>, 4>("mtd_block<f32,nd5,Mueller5d,Proj5dto2d,langevin,W4>"),
      mk<Spec<float, 5, NDS_POT_MUELLER5D, NDS_CV_PROJ5DTO2D, NDS_CV_PROJ5DTO2D, NDS_INT_LANGEVIN, BIAS_MTD,
              NDS_METHOD_MD, NDS_NOISE_PHILOX, 0>, 7>("mtd_block<f32,nd5,Mueller5d,Proj5dto2d,langevin,W7>"),
      mk<Spec<float, 5, NDS_POT_MUELLER5D, NDS_CV_PROJ5DTO2D, NDS_CV_PROJ5DTO2D, NDS_INT_LANGEVIN, BIAS_MTD,
              NDS_METHOD_MD, NDS_NOISE_PHILOX, 0>, 8>("mtd_block<f32,nd5,Mueller5d,Proj5dto2d,langevin,W8>"),
      mk<Spec<float, 5, NDS_POT_MUELLER5D, NDS_CV_PROJ5DTO2D, NDS_CV_PROJ5DTO2D, NDS_INT_LANGEVIN, BIAS_MTD,
              NDS_METHOD_MD, NDS_NOISE_PHILOX, 0>, 8, 7>("mtd_block<f32,nd5,Mueller5d,Proj5dto2d,langevin,W8,NW7>"),
#endif
  };
  *n = (int)(sizeof(table) / sizeof(table[0]));
  return table;
}
}  // namespace nds
