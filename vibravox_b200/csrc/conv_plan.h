// Launch planning for the Conv1d family: tile-config choice, grid shape, GemmP fill.
// Plain C++ (no CUDA) so tests/emu can drive the identical plan on the CPU.
#pragma once
#include "../../include/vbx.h"
#include "gemm_conv.cuh"

namespace vbx {

inline int plan_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// vbx_set_deterministic: one CTA per output tile walks a whole reduction in a fixed order (no split-K atomics)
inline int& deterministic_flag() { static int flag = 0; return flag; }

// Where the gate stage of vbx_epilogue runs, per input-gradient kernel form (bit mask; VBX_GATE_EPILOGUE, default 1|8):
// 1 = gather-form tensor-core kernel, 2 = streaming slab kernel, 4 = persistent slab kernel, 8 = fp32 FMA / direct
// kernels.  A form whose bit is clear runs its kernel without the gate and then vbx_fm_gate_bwd in place on the same
// stream (same bits).  The slab kernels run one tile at a time per CTA and their epilogue is not overlapped with another
// tile's main loop, so the two extra global loads per element are exposed latency there.  Measured on the bs=32 step
// (ms per step; separate aten::add + L1-pair backward + LeakyReLU backward passes = 41.36): all forms in the epilogue
// 43.31, all but the streaming slab 40.87, gather + FMA 40.59, gather only 40.57, none (in-place pass everywhere) 40.67.
int gate_epilogue_forms();
int launch_fm_gate(const float* y, const float* other, const float* coef, float gslope, const float* g, long long n,
                   float* out, void* stream);
// dx = dy * (ref > 0 ? 1 : slope) (dx may alias dy, or be NULL) and dbias[c] += sum of it: misc.cu
int launch_lrelu_bwd_bias(const float* dy, const float* ref, float* dx, float* dbias, int B, int C, int T, float slope,
                          void* stream);

struct GateArgs {
  const float* y; const float* other; const float* coef; float slope; float* dbias;
  // takes the gate stage out of P when `form` does not run it in its epilogue
  GateArgs(GemmP& P, int form) : y(nullptr), other(P.fm_other), coef(P.fm_coef), slope(P.gate_slope), dbias(P.gate_dbias) {
    if (P.gate && !(gate_epilogue_forms() & form)) {
      y = P.gate;
      P.gate = nullptr; P.fm_other = nullptr; P.fm_coef = nullptr;
    }
  }
  // out is (B, C, T)
  int finish(int rc, float* out, int B, int C, int T, void* stream) const {
    if (rc != 0) return rc;
    const long long n = (long long)B * C * T;
    if (y && dbias && !other)                     // gate and bias gradient in ONE in-place pass
      return launch_lrelu_bwd_bias(out, y, out, dbias, B, C, T, slope, stream);
    if (y) rc = launch_fm_gate(y, other, coef, slope, out, n, out, stream);
    if (rc == 0 && dbias) rc = launch_lrelu_bwd_bias(out, nullptr, nullptr, dbias, B, C, T, 1.f, stream);
    return rc;
  }
};

inline int pick_tm(int M) {
  if (M > 64) return 128;
  if (M > 32) return 64;
  if (M > 16) return 32;
  if (M > 8) return 16;
  if (M > 4) return 8;
  return 4;
}
inline int tn_of(int tm) { return tm >= 32 ? 128 : 256; }

// returns NULL when the descriptor is valid, else the complaint
inline const char* check_desc_msg(const vbx_conv_desc* d, int* code) {
  *code = VBX_BAD_POINTER;
  if (!d) return "conv: null descriptor";
  *code = VBX_BAD_SHAPE;
  if (!(d->B > 0 && d->Cin > 0 && d->Cout > 0 && d->Tin > 0 && d->Tout > 0 && d->K > 0 &&
        d->stride > 0 && d->dil > 0 && d->pad >= 0 && d->refl >= 0 && d->groups > 0))
    return "conv: non-positive dimension";
  if (!(d->Cin % d->groups == 0 && d->Cout % d->groups == 0))
    return "conv: channels not divisible by groups";
  if (!(d->refl <= d->pad && d->refl <= d->Tin - 1))
    return "conv: reflect halo larger than pad or than the signal";
  long long span = (long long)d->Tin + 2LL * d->pad - (long long)d->dil * (d->K - 1) - 1;
  if (!(span >= 0 && span / d->stride + 1 == d->Tout))
    return "conv: Tout inconsistent with Tin/pad/K/stride/dil";
  *code = VBX_UNSUPPORTED;
  if (!((long long)d->B * d->Cin * d->Tin < (1LL << 31) &&
        (long long)d->B * d->Cout * d->Tout < (1LL << 31)))
    return "conv: tensor too large for 32-bit tile indexing";
  if (d->stride > 65535) return "conv: stride too large";
  *code = 0;
  return nullptr;
}

inline void fill(GemmP& P, const vbx_conv_desc* d) {
  P.B = d->B; P.Cin = d->Cin; P.Cout = d->Cout; P.Tin = d->Tin; P.Tout = d->Tout; P.K = d->K;
  P.stride = d->stride; P.dil = d->dil; P.pad = d->pad; P.refl = d->refl; P.groups = d->groups;
  P.Cin_g = d->Cin / d->groups; P.Cout_g = d->Cout / d->groups;
  P.mtiles = 1; P.split = 0;
  P.W = nullptr; P.X = nullptr; P.DY = nullptr; P.Y = nullptr;
  P.bias = nullptr; P.res = nullptr; P.mask = nullptr; P.slope = 1.f; P.beta = 0.f;
  P.gate = nullptr; P.fm_other = nullptr; P.fm_coef = nullptr; P.gate_slope = 1.f; P.gate_dbias = nullptr;
}
inline void fill_epi(GemmP& P, const vbx_epilogue* e) {
  if (!e) return;
  P.bias = e->bias; P.res = e->res; P.mask = e->mask; P.slope = e->slope; P.beta = e->beta;
  P.gate = e->gate; P.gate_slope = e->gate_slope;
  P.fm_other = e->gate ? e->fm_other : nullptr; P.fm_coef = P.fm_other ? e->fm_coef : nullptr;
  P.gate_dbias = e->gate ? e->gate_dbias : nullptr;
}

struct Plan { int tm; bool bk; int grid[3]; };

inline Plan plan_conv(int mode, GemmP& P) {
  Plan pl;
  pl.bk = false;
  if (mode == FWD) {
    const int M = P.Cout_g;
    pl.tm = pick_tm(M);
    P.mtiles = plan_cdiv(M, pl.tm);
    pl.grid[0] = plan_cdiv((long long)P.B * P.Tout, tn_of(pl.tm));
    pl.grid[1] = P.mtiles * P.groups; pl.grid[2] = 1;
    pl.bk = P.stride >= 16;
  } else if (mode == DGRAD) {
    const int M = P.Cin_g;
    pl.tm = pick_tm(M);
    P.mtiles = plan_cdiv(M, pl.tm);
    const long long Up = plan_cdiv(P.Tin, P.stride);
    pl.grid[0] = plan_cdiv((long long)P.B * Up, tn_of(pl.tm));
    pl.grid[1] = P.mtiles * P.groups; pl.grid[2] = P.stride;
  } else if (mode == WGRAD) {
    const int M = P.Cout_g;
    pl.tm = pick_tm(M);
    const int tn = tn_of(pl.tm);
    P.mtiles = plan_cdiv(M, pl.tm);
    const int ntiles = plan_cdiv((long long)P.Cin_g * P.K, tn);
    const long long red = (long long)P.B * P.Tout;
    // split the (b,t) reduction so that the grid covers the 148 SMs a few times over
    long long tiles = (long long)ntiles * P.mtiles * P.groups;
    long long want = (148 * 4 + tiles - 1) / tiles;
    long long max_split = (red + 16 * 8 - 1) / (16 * 8);   // >= 8 chunks per slice
    if (want > max_split) want = max_split;
    if (want < 1 || deterministic_flag()) want = 1;
    if (want > 65535) want = 65535;
    long long split = (red + want - 1) / want;
    split = (split + 15) / 16 * 16;
    P.split = (int)split;
    pl.grid[0] = ntiles; pl.grid[1] = P.mtiles * P.groups; pl.grid[2] = plan_cdiv(red, split);
    pl.bk = true;
  } else {  // SCATTER
    const int M = P.Cin_g * P.K;
    pl.tm = pick_tm(M);
    P.mtiles = plan_cdiv(M, pl.tm);
    pl.grid[0] = plan_cdiv((long long)P.B * P.Tout, tn_of(pl.tm));
    pl.grid[1] = P.mtiles * P.groups; pl.grid[2] = 1;
  }
  return pl;
}

using C128 = Cfg<128, 128, 8, 8>;
using C64 = Cfg<64, 128, 8, 4>;
using C32 = Cfg<32, 128, 4, 4>;
using C16 = Cfg<16, 256, 4, 4>;
using C8 = Cfg<8, 256, 2, 4>;
using C4 = Cfg<4, 256, 1, 4>;

}  // namespace vbx
