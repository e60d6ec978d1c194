#define NDS_R float
#define NDS_ND 5
#define NDS_FN mtd_variants_f32_nd5
#define NDS_TAG "f32,nd5"
#define NDS_HOT_C5 1
#include "inst_mtd.cuh"
