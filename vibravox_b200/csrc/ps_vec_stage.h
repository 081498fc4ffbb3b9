// 16-byte staging of a stride-1 slab tile: index math and the fp32 -> bf16 hi/lo split, written as plain
// host / device C++ so that tests/emu/emu_slab.cpp executes THIS code on the CPU (no GPU in the build container).
// Used by tc_pslab_vec_kernel (tc_pslab_vec.cuh).  EXPERIMENTAL: built, checked in emulation, not yet enabled.
//
// The scalar staging of tc_pslab_kernel issues one LDG.32 per (position, channel): 32 load instructions per
// 4 positions x 8 channels.  Activations are time-contiguous and every generator length is a multiple of 4, so
// the same block is 8 LDG.128 if the slab starts on a multiple of 4 samples: the slab origin is moved back by
//     a = (tau0 mod 4),  tau0 = first input sample the tile reads,
// i.e. slab unit j holds input sample tau0 - a + j, and every tap's descriptor start address is advanced by `a`
// units.  One work item = 4 consecutive slab units x 8 channels (one half of a 16-channel group):
//     8 x float4 loads  ->  4 x (16-byte hi unit, 16-byte lo unit)  in the [plane][half][unit][8 ch] slab layout.
// Groups that touch a halo, a batch-item boundary, the end of the tensor or an unaligned address take the
// per-sample path (map_pos: reflect / zero), so the result is identical to the scalar staging shifted by `a`.
#pragma once
#include <stdint.h>
#include "gemm_conv.cuh"

namespace vbx {

struct PsVec {
  int npos;       // samples a 128-row tile touches: 127 + (K-1)*dil + 1   (stride 1)
  int groups4;    // 4-sample groups staged per channel: ceil((npos + 3) / 4)
  int units;      // slab units per plane half = 4 * groups4
  int a_stage;    // bytes of one 16-channel group: 2 planes x 2 halves x units x 16
  int plane, half;
};

VBX_HD PsVec ps_vec_geom(const GemmP& G) {
  PsVec v;
  v.npos = 127 + (G.K - 1) * G.dil + 1;
  v.groups4 = (v.npos + 3 + 3) / 4;
  v.units = 4 * v.groups4;
  v.a_stage = 64 * v.units;
  v.plane = v.a_stage / 2;
  v.half = v.plane / 2;
  return v;
}

// first input sample of row tile `tile` (may be negative: left halo) and the alignment shift a in [0, 4)
VBX_HD void ps_vec_tile(const GemmP& G, int R, int tile, int& q_origin, int& a) {
  const int q0 = tile * 128;                       // virtual timeline, period R per batch item (stride 1)
  const int p0 = q0 % R;
  a = mod_pos(p0 - G.pad, 4);
  q_origin = q0 - a;                               // slab unit j <-> virtual position q_origin + j
}

// round-to-nearest-even fp32 -> bf16 (bit pattern), finite inputs; what cvt.rn.bf16.f32 returns
VBX_HD uint32_t ps_bf16_rn(float x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
#else
  union { float f; uint32_t u; } c; c.f = x;
  const uint32_t lsb = (c.u >> 16) & 1u;
  return (c.u + 0x7FFFu + lsb) >> 16;
#endif
}
VBX_HD float ps_bf16_to_float(uint32_t h) {
  union { float f; uint32_t u; } c; c.u = h << 16;
  return c.f;
}

// Stage one item.  item = (cg * 2 + half) * groups4 + jg;  st = base of the slab slot (all channel groups).
// Returns false when the item index is past the tile's work list.
VBX_HD bool ps_vec_stage_item(const GemmP& G, const PsVec& V, int R, int grp, int ncg, int q_origin, int item,
                              unsigned char* st, bool x_aligned) {
  const int jg = item % V.groups4;
  const int rest = item / V.groups4;
  const int hf = rest & 1, cg = rest >> 1;
  if (cg >= ncg) return false;
  const int c0 = cg * 16 + hf * 8;                               // first channel (inside the group) of this item
  const int nch = G.Cin_g - c0 < 8 ? (G.Cin_g - c0 < 0 ? 0 : G.Cin_g - c0) : 8;
  const int q = q_origin + 4 * jg;                               // virtual position of the first of 4 samples
  float v[8][4];
  bool fast = false;
  if (q >= 0 && x_aligned && (G.Tin & 3) == 0) {
    const int b = q / R, p = q % R;
    const int tau = p - G.pad;
    if (b < G.B && p + 3 < R && tau >= 0 && tau + 3 < G.Tin && (tau & 3) == 0) {
      fast = true;
      const float* src = G.X + ((long long)b * G.Cin + (long long)grp * G.Cin_g + c0) * G.Tin + tau;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e < nch) {
#if defined(__CUDA_ARCH__)
          const float4 t = *reinterpret_cast<const float4*>(src + (long long)e * G.Tin);
          v[e][0] = t.x; v[e][1] = t.y; v[e][2] = t.z; v[e][3] = t.w;
#else
          for (int r = 0; r < 4; ++r) v[e][r] = src[(long long)e * G.Tin + r];
#endif
        } else {
          v[e][0] = v[e][1] = v[e][2] = v[e][3] = 0.f;
        }
      }
    }
  }
  if (!fast) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qq = q + r;
      int off = -1;
      if (qq >= 0) {
        const int b = qq / R, p = qq % R;
        const int tau = map_pos(p - G.pad, G.Tin, G.refl);
        if (b < G.B && tau >= 0) off = (b * G.Cin + grp * G.Cin_g + c0) * G.Tin + tau;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e][r] = (off >= 0 && e < nch) ? G.X[off + e * G.Tin] : 0.f;
    }
  }
  unsigned char* d0 = st + (size_t)cg * V.a_stage + (size_t)hf * V.half + (size_t)(4 * jg) * 16;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = v[2 * e][r], x1 = v[2 * e + 1][r];
      const uint32_t h0 = ps_bf16_rn(x0), h1 = ps_bf16_rn(x1);
      const uint32_t l0 = ps_bf16_rn(x0 - ps_bf16_to_float(h0)), l1 = ps_bf16_rn(x1 - ps_bf16_to_float(h1));
      hi[e] = h0 | (h1 << 16);
      lo[e] = l0 | (l1 << 16);
    }
    uint32_t* ph = reinterpret_cast<uint32_t*>(d0 + r * 16);
    uint32_t* pl = reinterpret_cast<uint32_t*>(d0 + r * 16 + V.plane);
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(ph) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(pl) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#else
    for (int e = 0; e < 4; ++e) { ph[e] = hi[e]; pl[e] = lo[e]; }
#endif
  }
  return true;
}

}  // namespace vbx
