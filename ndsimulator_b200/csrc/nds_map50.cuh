// nds_map50.cuh -- Map50to2d (colvars/misc.py:19-156): two quadratic forms of a 50-dimensional configuration,
//   c_k = x K_k x^T,   J_k = 2 x K_k          (compute :140-150, jacobian :152-155)
// with K_1 = diag(vec1 / |vec1|), K_1[0][1] = K_1[1][0] = 1 (:128-131) and K_2 = diag(vec2 / |vec2|),
// K_2[3][5] = K_2[5][3] = 0.2, K_2[4][8] = K_2[8][4] = 0.5, K_2[2][4] = K_2[4][2] = 1 (:129,132-137).
// No potential of the reference accepts ndim = 50, so the reference cannot run a trajectory with this
// colvar; what it CAN do -- values and Jacobians at given configurations -- is served here by one thread per
// configuration (x[50][n] SoA, coalesced), in fp64.  The 2 x 50 table values are the reference's own class
// constants (misc.py:21-125), data that has to be identical.
#pragma once
#include <cuda_runtime.h>

namespace nds {

constexpr int MAP50_ND = 50;

struct Map50Table { double v1[MAP50_ND], v2[MAP50_ND]; };  // vec / |vec|

inline Map50Table map50_table() {
  static const double vec1[MAP50_ND] = {
    0.9621493306867017, 0.2611502170144795, 0.0000000000000000, 0.7331647024340944, 0.8708183789325146,
    0.0805017856037096, 0.0000000000000000, 0.5785249542521097, 0.3783625809001937, 0.0000000000000000,
    0.0571747072559683, 0.9724187143800709, 0.0000000000000000, 0.041131319898087, 0.0000000000000000,
    0.2765931554549067, 0.0912801608664523, 0.0000000000000000, 0.0431413377076534, 0.21668041341737354,
    0.04381849659171735, 0.0000000000000000, 0.0120163151455723, 0.00000000000000000, 0.0038527070151484,
    0.729285346078329, 0.09400213558501833, 0.00000000000000000, 0.0345084258996749, 0.0000000000000000,
    0.0602460013004428, 0.0000000000000000, 0.0102889267052817, 0.7013536441873398, 0.00125254636835465,
    0.000000000000000, 0.0006896724138744, 0.0042907976312725, 0.0100253907332234, 0.0005460275194114628,
    0.00567411387523127, 0.0025103010306175, 0.0007127704701611, 0.0001890413255533, 0.0407958220608383,
    0.0061146053157138, 0.00491691153887918, 0.0024312731237804, 0.0761174326669661, 0.00586435669539784,
  };
  static const double vec2[MAP50_ND] = {
    0.7200723641543899, 0.938701822380177, 0.063114103710782, 0.0, 0.5650229051546,
    0.0, 0.0, 0.0, 0.0047707786427023, 0.0,
    0.0, 0.0, 0.0, 0.0254131316679295, 0.0,
    0.0, 0.0, 0.0, 0.0006544091825368, 0.0,
    0.0, 0.7915907738260504, 0.0006566689112959, 0.0308050889004108, 0.0306030445075646,
    0.0683335335391172, 0.0927967809881871, 0.0, 0.0, 0.0668273272540637,
    0.9543323517079718, 0.0, 0.0, 0.0, 0.0705778822165263,
    0.0, 0.0167903655654037, 0.9330919923191575, 0.0, 0.9176521018708412,
    0.0505357766611562, 0.0, 0.0750136062509377, 0.5709571823782142, 0.0591655909593535,
    0.0, 0.0, 0.0113295431324928, 0.0, 0.0444879007946433,
  };
  Map50Table t;
  double n1 = 0, n2 = 0;
  for (int d = 0; d < MAP50_ND; d++) { n1 += vec1[d] * vec1[d]; n2 += vec2[d] * vec2[d]; }
  n1 = sqrt(n1); n2 = sqrt(n2);  // np.linalg.norm
  for (int d = 0; d < MAP50_ND; d++) { t.v1[d] = vec1[d] / n1; t.v2[d] = vec2[d] / n2; }
  return t;
}

// c[2][n], J[2][50][n] (either may be null)
__global__ void map50to2d_kernel(const __grid_constant__ Map50Table T, const double* __restrict__ x, long long n,
                                 double* __restrict__ c, double* __restrict__ J) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // the coordinates that carry off-diagonal couplings
  const double x0 = x[0 * n + i], x1 = x[1 * n + i], x2 = x[2 * n + i], x3 = x[3 * n + i], x4 = x[4 * n + i];
  const double x5 = x[5 * n + i], x8 = x[8 * n + i];
  double c1 = 0.0, c2 = 0.0;
  for (int d = 0; d < MAP50_ND; d++) {
    const double xd = x[d * n + i];
    // row d of x K: the diagonal term plus the (at most two) couplings of column d
    // (explicit roundings: products and sums rounded one by one like the reference's matrix products, and the same
    // arithmetic whether or not the Jacobian is stored)
    double y1 = __dmul_rn(xd, T.v1[d]), y2 = __dmul_rn(xd, T.v2[d]);
    if (d == 0) y1 = __dadd_rn(y1, x1);
    if (d == 1) y1 = __dadd_rn(y1, x0);
    if (d == 2) y2 = __dadd_rn(y2, x4);
    if (d == 3) y2 = __dadd_rn(y2, __dmul_rn(0.2, x5));
    if (d == 4) y2 = __dadd_rn(__dadd_rn(y2, __dmul_rn(0.5, x8)), x2);
    if (d == 5) y2 = __dadd_rn(y2, __dmul_rn(0.2, x3));
    if (d == 8) y2 = __dadd_rn(y2, __dmul_rn(0.5, x4));
    c1 = __dadd_rn(c1, __dmul_rn(y1, xd));
    c2 = __dadd_rn(c2, __dmul_rn(y2, xd));
    if (J) { J[(0 * MAP50_ND + d) * n + i] = 2.0 * y1; J[(1 * MAP50_ND + d) * n + i] = 2.0 * y2; }
  }
  if (c) { c[0 * n + i] = c1; c[1 * n + i] = c2; }
}

}  // namespace nds
