// inst_hot.cu -- fully specialised md_block variants for the BASELINE.json configurations.
// Every choice is a compile-time constant, so the switch statements of nds_physics.cuh fold away and the
// step loop contains only the arithmetic of that configuration.
#include "nds_variants.h"

namespace nds {
namespace {
constexpr int OFF = 0;
constexpr int MD = NDS_METHOD_MD, CM = NDS_METHOD_COMMITTOR, PHX = NDS_NOISE_PHILOX;
constexpr int ORI = NDS_CV_ORIGINAL, P52 = NDS_CV_PROJ5DTO2D;
}  // namespace

const Variant* variants_hot(int* n) {
  static const Variant table[] = {
      // C2: 2-D Mueller-Brown Langevin, no bias, no frames
      make_variant<Spec<float, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, OFF>>(
          "md_block<f32,nd2,Mueller2d,Original,langevin,nobias>"),
      make_variant<Spec<double, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, OFF>>(
          "md_block<f64,nd2,Mueller2d,Original,langevin,nobias>"),
      // C2 with Blackwell packed FP32 (FFMA2): two walkers per thread, same physics templates
      make_variant<Spec<f2, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, OFF>>(
          "md_block<f32x2,nd2,Mueller2d,Original,langevin,nobias>"),
      // C2 physics with device-side frames (dump rows of io/dump.py) for the dump-bandwidth measurement
      make_variant<Spec<float, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, 1>>(
          "md_block<f32,nd2,Mueller2d,Original,langevin,nobias,frames>"),
      // C2 as SURVEY 8d defines it: frames for a walker prefix (packed, float rows) and the in-kernel colvar histogram
      make_variant<Spec<f2, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, 1>>(
          "md_block<f32x2,nd2,Mueller2d,Original,langevin,nobias,frames>"),
      make_variant<Spec<f2, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, 2>>(
          "md_block<f32x2,nd2,Mueller2d,Original,langevin,nobias,hist>"),
      make_variant<Spec<f2, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, 3>>(
          "md_block<f32x2,nd2,Mueller2d,Original,langevin,nobias,frames+hist>"),
      // C3: committor shooting on the same surface
      make_variant<Spec<float, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, CM, PHX, OFF>>(
          "md_block<f32,nd2,Mueller2d,Original,langevin,committor>"),
      make_variant<Spec<float, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, CM, PHX, OFF, 2>>(
          "md_block<f32,nd2,Mueller2d,Original,langevin,committor,2 basins>"),
      // two shots per thread (packed FP32); a stopped lane parks its exit state and rides along
      make_variant<Spec<f2, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, 0, CM, PHX, OFF, 2>>(
          "md_block<f32x2,nd2,Mueller2d,Original,langevin,committor,2 basins>"),
      // C4: 5-D umbrella windows, 2nd-order Langevin, per-walker (k, r0)
      make_variant<Spec<float, 5, NDS_POT_MUELLER5D, P52, P52, NDS_INT_LANGEVIN2, BIAS_UMB | BIAS_UMB_WALKER, MD, PHX, OFF>>(
          "md_block<f32,nd5,Mueller5d,Proj5dto2d,2nd-langevin,umb-windows>"),
      make_variant<Spec<float, 5, NDS_POT_MUELLER5D, P52, P52, NDS_INT_LANGEVIN2, BIAS_UMB, MD, PHX, OFF>>(
          "md_block<f32,nd5,Mueller5d,Proj5dto2d,2nd-langevin,umb>"),
      // (a packed f2 instantiation of C4 was measured 7 % SLOWER than the scalar one: 117 registers,
      //  half the occupancy, and the step is dominated by Philox + Box-Muller, which do not pack)
      // the remaining shipped examples: examples/2d-umb.yaml and examples/1d-md.yaml
      make_variant<Spec<float, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN2, BIAS_UMB, MD, PHX, OFF>>(
          "md_block<f32,nd2,Mueller2d,Original,2nd-langevin,umb>"),
      make_variant<Spec<float, 1, NDS_POT_DOUBLEWELL1D, ORI, ORI, NDS_INT_LANGEVIN, 0, MD, PHX, OFF>>(
          "md_block<f32,nd1,DoubleWell1d,Original,langevin,nobias>"),
      // C5 / 2d-mtd: metadynamics
      make_variant<Spec<float, 5, NDS_POT_MUELLER5D, P52, P52, NDS_INT_LANGEVIN, BIAS_MTD, MD, PHX, OFF>>(
          "md_block<f32,nd5,Mueller5d,Proj5dto2d,langevin,mtd>"),
      make_variant<Spec<float, 2, NDS_POT_MUELLER2D, ORI, ORI, NDS_INT_LANGEVIN, BIAS_MTD, MD, PHX, OFF>>(
          "md_block<f32,nd2,Mueller2d,Original,langevin,mtd>"),
  };
  *n = (int)(sizeof(table) / sizeof(table[0]));
  return table;
}
}  // namespace nds
