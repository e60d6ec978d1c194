#define NDS_R double
#define NDS_ND 5
#define NDS_FN mtd_variants_f64_nd5
#define NDS_TAG "f64,nd5"

#include "inst_mtd.cuh"
