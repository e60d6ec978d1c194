#define NDS_R float
#define NDS_ND 2
#define NDS_FN mtd_variants_f32_nd2
#define NDS_TAG "f32,nd2"

#include "inst_mtd.cuh"
