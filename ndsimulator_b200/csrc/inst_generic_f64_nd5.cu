#define NDS_R double
#define NDS_ND 5
#define NDS_FN variants_generic_f64_nd5
#define NDS_TAG "f64,nd5"
#include "inst_generic.cuh"
