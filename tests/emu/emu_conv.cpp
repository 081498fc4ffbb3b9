// CPU emulation of the implicit-GEMM Conv1d kernels: runs the SAME tile code
// (vibravox_b200/csrc/gemm_conv.cuh) and the SAME launch plan (conv_plan.h) with the
// 256 threads of each block serialised between barriers.  TEST INFRASTRUCTURE ONLY -
// nothing under vibravox_b200/ loads this; there is no GPU in the build container, so
// this is how the index math is checked before a gpurun call.
#include <vector>
#include <cstring>
#include "../../vibravox_b200/csrc/conv_plan.h"

using namespace vbx;

template <class C, int MODE, bool BK>
static void run(const GemmP& P, const Plan& pl) {
  using T = Tile<C, MODE, BK>;
  std::vector<float> As(C::KC * C::LDA), Bs(C::KC * C::LDB);
  std::vector<typename T::TS> st(C::NT);
  for (int bz = 0; bz < pl.grid[2]; ++bz)
    for (int by = 0; by < pl.grid[1]; ++by)
      for (int bx = 0; bx < pl.grid[0]; ++bx) {
        Blk blk{bx, by, bz};
        for (int t = 0; t < C::NT; ++t) T::prologue(P, blk, t, st[t]);
        if (st[0].n_base >= st[0].N) continue;
        if (MODE == DGRAD) {
          for (int img = 0; img < 3; ++img) {
            if (!T::dgrad_need_img(P, blk, img)) continue;
            for (int t = 0; t < C::NT; ++t) T::dgrad_begin_img(P, t, st[t], img);
            int nch = T::num_chunks(st[0]);
            for (int c = 0; c < nch; ++c) {
              for (int t = 0; t < C::NT; ++t) T::load_chunk(P, t, st[t], c, As.data(), Bs.data());
              for (int t = 0; t < C::NT; ++t) T::compute_chunk(t, st[t], As.data(), Bs.data());
            }
          }
        } else {
          int nch = T::num_chunks(st[0]);
          for (int c = 0; c < nch; ++c) {
            for (int t = 0; t < C::NT; ++t) T::load_chunk(P, t, st[t], c, As.data(), Bs.data());
            for (int t = 0; t < C::NT; ++t) T::compute_chunk(t, st[t], As.data(), Bs.data());
          }
        }
        for (int t = 0; t < C::NT; ++t)
          T::epilogue(P, blk, t, st[t], [](float* p, float v) { *p += v; });
      }
}

template <int MODE, bool BK>
static void dispatch(const GemmP& P, const Plan& pl) {
  switch (pl.tm) {
    case 128: run<C128, MODE, BK>(P, pl); break;
    case 64: run<C64, MODE, BK>(P, pl); break;
    case 32: run<C32, MODE, BK>(P, pl); break;
    case 16: run<C16, MODE, BK>(P, pl); break;
    case 8: run<C8, MODE, BK>(P, pl); break;
    default: run<C4, MODE, BK>(P, pl); break;
  }
}

extern "C" int emu_conv(int mode, const vbx_conv_desc* d, const float* a, const float* b,
                        const vbx_epilogue* e, float* out) {
  int code = 0;
  if (check_desc_msg(d, &code)) return code;
  GemmP P; fill(P, d); fill_epi(P, e);
  Plan pl;
  if (mode == FWD) { P.X = a; P.W = b; P.Y = out; pl = plan_conv(FWD, P);
    if (pl.bk) dispatch<FWD, true>(P, pl); else dispatch<FWD, false>(P, pl); }
  else if (mode == DGRAD) { P.X = a; P.W = b; P.Y = out; pl = plan_conv(DGRAD, P); dispatch<DGRAD, false>(P, pl); }
  else if (mode == WGRAD) { P.X = a; P.DY = b; P.Y = out; pl = plan_conv(WGRAD, P); dispatch<WGRAD, true>(P, pl); }
  else { P.X = a; P.W = b; P.Y = out; pl = plan_conv(SCATTER, P); dispatch<SCATTER, false>(P, pl); }
  return 0;
}
