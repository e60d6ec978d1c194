// inst_generic.cuh -- generic md_block variants of one (precision, ndim): the integrator is a
// compile-time choice (it fixes the number of Philox normals per step), everything else is read from
// the constant block at run time.  Included by inst_generic_<prec>_nd<N>.cu with NDS_R / NDS_ND /
// NDS_FN / NDS_TAG defined.
#include "nds_variants.h"

namespace nds {
const Variant* NDS_FN(int* n) {
  static const Variant table[] = {
      make_variant<Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_LANGEVIN>>("md_block<" NDS_TAG ",generic,langevin>"),
      make_variant<Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_LANGEVIN2>>("md_block<" NDS_TAG ",generic,2nd-langevin>"),
      make_variant<Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_NVE>>("md_block<" NDS_TAG ",generic,nve>"),
      make_variant<Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_RESCALE>>("md_block<" NDS_TAG ",generic,rescale>"),
      make_variant<Spec<NDS_R, NDS_ND, DYN, DYN, DYN, NDS_INT_MINIMIZE>>("md_block<" NDS_TAG ",generic,minimize>"),
  };
  *n = (int)(sizeof(table) / sizeof(table[0]));
  return table;
}
}  // namespace nds
