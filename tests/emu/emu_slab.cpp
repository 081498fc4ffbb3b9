// CPU execution of the 16-byte slab staging of tc_pslab_vec_kernel (vibravox_b200/csrc/ps_vec_stage.h): the SAME
// function the device kernel calls, driven item by item for the 256 producer threads of one CTA and one row tile.
// TEST INFRASTRUCTURE ONLY (tests/test_emu_slab.py); nothing under vibravox_b200/ loads this.
#include <cstring>
#include "../../include/vbx.h"
#include "../../vibravox_b200/csrc/ps_vec_stage.h"

using namespace vbx;

// geom_out: [npos, groups4, units, a_stage, plane, half, R, ncg, a, q_origin]
extern "C" int emu_ps_vec_stage(const vbx_conv_desc* d, int grp, const float* x, int x_aligned, int tile,
                                unsigned char* slot, long long slot_bytes, int* geom_out) {
  GemmP G;
  std::memset(&G, 0, sizeof(G));
  G.B = d->B; G.Cin = d->Cin; G.Cout = d->Cout; G.Tin = d->Tin; G.Tout = d->Tout; G.K = d->K;
  G.stride = d->stride; G.dil = d->dil; G.pad = d->pad; G.refl = d->refl; G.groups = d->groups;
  G.Cin_g = d->Cin / d->groups; G.Cout_g = d->Cout / d->groups;
  G.X = x;
  if (G.stride != 1) return -1;
  const PsVec V = ps_vec_geom(G);
  const int R = G.Tout + (G.K - 1) * G.dil;                  // plan_slab: virtual rows per batch item at stride 1
  const int ncg = (G.Cin_g + 15) / 16;
  if ((long long)ncg * V.a_stage > slot_bytes) return -2;
  int q_origin, a;
  ps_vec_tile(G, R, tile, q_origin, a);
  for (int tid = 0; tid < 256; ++tid)
    for (int item = tid; ps_vec_stage_item(G, V, R, grp, ncg, q_origin, item, slot, x_aligned != 0); item += 256) {
    }
  const int g[10] = {V.npos, V.groups4, V.units, V.a_stage, V.plane, V.half, R, ncg, a, q_origin};
  std::memcpy(geom_out, g, sizeof(g));
  return 0;
}
