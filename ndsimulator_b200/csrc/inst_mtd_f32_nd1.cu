#define NDS_R float
#define NDS_ND 1
#define NDS_FN mtd_variants_f32_nd1
#define NDS_TAG "f32,nd1"

#include "inst_mtd.cuh"
