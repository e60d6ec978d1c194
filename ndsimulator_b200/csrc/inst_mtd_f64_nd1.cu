#define NDS_R double
#define NDS_ND 1
#define NDS_FN mtd_variants_f64_nd1
#define NDS_TAG "f64,nd1"

#include "inst_mtd.cuh"
