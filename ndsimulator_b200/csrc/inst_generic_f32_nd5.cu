#define NDS_R float
#define NDS_ND 5
#define NDS_FN variants_generic_f32_nd5
#define NDS_TAG "f32,nd5"
#include "inst_generic.cuh"
