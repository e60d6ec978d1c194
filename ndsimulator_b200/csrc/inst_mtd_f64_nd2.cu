#define NDS_R double
#define NDS_ND 2
#define NDS_FN mtd_variants_f64_nd2
#define NDS_TAG "f64,nd2"

#include "inst_mtd.cuh"
