#define NDS_R double
#define NDS_ND 1
#define NDS_FN variants_generic_f64_nd1
#define NDS_TAG "f64,nd1"
#include "inst_generic.cuh"
