// Fused ResidualUnit forward (eben_generator.py:287-316 of the reference):
//     out = x + LeakyReLU_slope( W_pw (1x1) * ( W_dil (k3, dilation d, reflect halo d) * x ) )
// as ONE persistent kernel: x is read from HBM once, `out` is written once; the intermediate `h` never leaves the SM.
//
//   TMA      one thread streams the fp32 tile x[b, 0:C, t0-dA : t0+128+dA] (dA = d rounded up to 4 samples: the box must
//            start on a 16-byte boundary of the row - an unaligned or negative start traps, measured with
//            tools/tma_probe.cu - so the first tile of an item starts at 0) into a shared-memory ring with ONE 3-D
//            tensor-map copy per tile (cp.async.bulk.tensor: columns beyond T arrive as zeros, nothing is gathered by threads);
//   convert  4 warps turn the raw tile into the K-major bf16 hi/lo slab of tc_slab.cuh ([position][8 channels] units; the
//            three taps of the dilated conv are the SAME slab read through descriptors advanced by k*d units) and fold the
//            reflect halo in by reading the mirrored column of the raw tile;
//   MMA      one thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM, three products per operand pair: hi*hi + hi*lo +
//            lo*hi) for the dilated conv into D1, later for the pointwise conv into D2, whichever operand is ready first;
//   mid      4 warps read D1 (tcgen05.ld), split it into bf16 hi/lo and store it as the A operand of the pointwise conv
//            over the slab that the first conv has finished reading;
//   epilogue 4 warps read D2, apply LeakyReLU, add the residual from the RAW fp32 tile still in shared memory (exact, and
//            no second trip to HBM / L2), and write `out` coalesced along time.
// The weights of both convs (packed once per step by ru_pack_kernel) stay resident in shared memory for the life of the
// CTA; tiles flow through rings (raw: NR deep, slab: NA deep, TMEM accumulators: 2 deep) guarded by mbarriers, so the
// load of tile i+2, the conversion of tile i+1 and the epilogue of tile i overlap.
// Training-time extras (optional pointers): h (what the weight gradient of the pointwise conv needs) is written by the
// mid warps, the 1-byte activation mask by the epilogue warps.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace vbx {
namespace ru {
using namespace vbx::tc;

static const int kRows = 128;
static const int kThreads = 448;            // warps: 0 TMA, 1 MMA, 2-5 convert, 6-9 mid, 10-13 epilogue
static const int kSmemLimit = 227 * 1024;

struct RuP {
  int B, C, T, d;
  int Wpos;         // 128 + 2d positions staged per tile
  int dA;           // d rounded up to a multiple of 4: the raw tile starts at max(t0 - dA, 0) (16-byte aligned box start)
  int Wraw;         // 128 + 2 dA columns of the raw tile
  int ncg;          // C / 16
  int tpi;          // 128-row tiles per batch item
  int ntiles;
  int NR, NA;       // ring depths: raw tiles / slabs
  int tmem_cols;
  int dbg;          // VBX_RU_DBG bit mask (bring-up only): 1 no tensormap prefetch, 2 no tensor load, 4 no MMAs, 8 / 16 no TMEM loads in mid / epilogue
  float slope;
  const unsigned char* packed;
  float* out;
  float* h;         // optional
  unsigned char* mask;   // optional
};

__host__ __device__ inline int ru_w_bytes(int C) { return (C / 16) * 4 * C * 64; }       // W1 (3 taps) + W2, hi + lo
__host__ __device__ inline int ru_raw_bytes(const RuP& P) { return P.C * P.Wraw * 4; }
__host__ __device__ inline int ru_ab_bytes(const RuP& P) { return P.ncg * P.Wpos * 64; }
__host__ __device__ inline int ru_nbars(const RuP& P) { return 2 * P.NR + 4 * P.NA + 3; }
static size_t ru_smem_bytes(const RuP& P) {
  return (size_t)ru_w_bytes(P.C) + (size_t)P.NR * ru_raw_bytes(P) + (size_t)P.NA * ru_ab_bytes(P) +
         (size_t)ru_nbars(P) * 8 + 16;
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__global__ void __launch_bounds__(kThreads, 2) ru_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const RuP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, d = P.d, Wpos = P.Wpos, Wraw = P.Wraw, ncg = P.ncg, NR = P.NR, NA = P.NA;
  const int w_bytes = ru_w_bytes(C), raw_bytes = ru_raw_bytes(P), ab_bytes = ru_ab_bytes(P);
  unsigned char* w0 = smem;
  unsigned char* raw0 = smem + w_bytes;
  unsigned char* ab0 = raw0 + (size_t)NR * raw_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ab0 + (size_t)NA * ab_bytes);
  uint64_t* raw_full = bars;                 // [NR] TMA bytes landed
  uint64_t* raw_free = raw_full + NR;        // [NR] 4 epilogue warps done with the raw tile
  uint64_t* a1_full = raw_free + NR;         // [NA] 4 convert warps: slab of the dilated conv staged
  uint64_t* mma1_done = a1_full + NA;        // [NA] D1 complete (and the slab no longer read)
  uint64_t* a2_full = mma1_done + NA;        // [NA] 4 mid warps: operand of the pointwise conv staged
  uint64_t* mma2_done = a2_full + NA;        // [NA] D2 complete (and the slab free again)
  uint64_t* d2_free = mma2_done + NA;        // [2]  4 epilogue warps have drained D2[j]
  uint64_t* w_full = d2_free + 2;            // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int my_tiles = ((int)blockIdx.x < P.ntiles) ? (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < NR; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], 4); }
    for (int i = 0; i < NA; ++i) {
      mbar_init(&a1_full[i], 4); mbar_init(&mma1_done[i], 1); mbar_init(&a2_full[i], 4); mbar_init(&mma2_done[i], 1);
    }
    mbar_init(&d2_free[0], 4); mbar_init(&d2_free[1], 4);
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: D1[j] at j*C, D2[j] at 2C + j*C   (j = tile parity)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && my_tiles > 0) {
      if (!(P.dbg & 1)) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
      mbar_expect_tx(w_full, (uint32_t)w_bytes);
      for (int off = 0; off < w_bytes; off += 32768) {
        const int n = w_bytes - off < 32768 ? w_bytes - off : 32768;
        bulk_copy_g2s(w0 + off, P.packed + off, (uint32_t)n, w_full);
      }
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % NR;
        const uint32_t par = (uint32_t)((i / NR) & 1);
        mbar_wait(&raw_free[s], par ^ 1u);
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
        if (P.dbg & 2) { mbar_arrive(&raw_full[s]); continue; }
        mbar_expect_tx(&raw_full[s], (uint32_t)raw_bytes);
        const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;
        tma_load_3d(raw0 + (size_t)s * raw_bytes, &tmap_x, &raw_full[s], start, 0, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = make_idesc_bf16(C, /*a_mn=*/false, /*b_mn=*/false);
      const uint32_t lbo_a1 = (uint32_t)Wpos * 16, lbo_b = (uint32_t)C * 16;
      const uint32_t plane_a1 = (uint32_t)Wpos * 32, plane_b = (uint32_t)C * 32, tile_b = (uint32_t)C * 64;
      const uint32_t wbase = smem_u32(w0), w2base = wbase + (uint32_t)ncg * 3u * tile_b;
      mbar_wait(w_full, 0);
      int n1 = 0, n2 = 0;                       // next tile of the dilated / pointwise conv
      while (n2 < my_tiles) {
        bool did = false;
        if (n1 < my_tiles && n1 - n2 < 2) {     // (D1 / D2 are two deep)
          const int sa = n1 % NA;
          if (mbar_try_wait(&a1_full[sa], (uint32_t)((n1 / NA) & 1))) {
            tc_fence_after();
            const uint32_t abase = smem_u32(ab0 + (size_t)sa * ab_bytes);
            const uint32_t dcol = tmem_base + (uint32_t)((n1 & 1) * C);
            uint32_t acc = 0;
            for (int cg = 0; cg < ((P.dbg & 4) ? 0 : ncg); ++cg) {
              const uint32_t a_cg = abase + (uint32_t)cg * (uint32_t)Wpos * 64u;
#pragma unroll
              for (int tap = 0; tap < 3; ++tap) {
                const uint32_t a_hi = a_cg + (uint32_t)(tap * d) * 16u;
                const uint32_t b_hi = wbase + (uint32_t)(cg * 3 + tap) * tile_b;
                const uint64_t da_hi = make_desc(a_hi, lbo_a1, 128), da_lo = make_desc(a_hi + plane_a1, lbo_a1, 128);
                const uint64_t db_hi = make_desc(b_hi, lbo_b, 128), db_lo = make_desc(b_hi + plane_b, lbo_b, 128);
                mma_bf16_ss(dcol, da_hi, db_hi, idesc, acc);
                acc = 1;
                mma_bf16_ss(dcol, da_hi, db_lo, idesc, 1);
                mma_bf16_ss(dcol, da_lo, db_hi, idesc, 1);
              }
            }
            mma_commit(&mma1_done[sa]);
            ++n1;
            did = true;
          }
        }
        if (n2 < n1) {
          const int sa = n2 % NA, j = n2 & 1;
          if (mbar_try_wait(&a2_full[sa], (uint32_t)((n2 / NA) & 1))) {
            if (n2 >= 2) mbar_wait(&d2_free[j], (uint32_t)(((n2 >> 1) - 1) & 1));
            tc_fence_after();
            const uint32_t abase = smem_u32(ab0 + (size_t)sa * ab_bytes);
            const uint32_t dcol = tmem_base + (uint32_t)(2 * C + j * C);
            uint32_t acc = 0;
            for (int cg = 0; cg < ((P.dbg & 4) ? 0 : ncg); ++cg) {
              const uint32_t a_hi = abase + (uint32_t)cg * 8192u;
              const uint32_t b_hi = w2base + (uint32_t)cg * tile_b;
              const uint64_t da_hi = make_desc(a_hi, 2048, 128), da_lo = make_desc(a_hi + 4096u, 2048, 128);
              const uint64_t db_hi = make_desc(b_hi, lbo_b, 128), db_lo = make_desc(b_hi + plane_b, lbo_b, 128);
              mma_bf16_ss(dcol, da_hi, db_hi, idesc, acc);
              acc = 1;
              mma_bf16_ss(dcol, da_hi, db_lo, idesc, 1);
              mma_bf16_ss(dcol, da_lo, db_hi, idesc, 1);
            }
            mma_commit(&mma2_done[sa]);
            ++n2;
            did = true;
          }
        }
        if (!did) __nanosleep(32);
      }
    }
  } else if (warp < 6) {
    // ===================== convert: raw fp32 tile -> K-major bf16 hi/lo slab (+ reflect halo) =====================
    const int ct = tid - 64;                    // 0..127
    const int nitems = (C / 8) * Wpos;          // (8-channel unit, position)
    for (int i = 0; i < my_tiles; ++i) {
      const int sr = i % NR, sa = i % NA;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int t0 = (tile % P.tpi) * kRows;
      if (i >= NA) mbar_wait(&mma2_done[sa], (uint32_t)(((i / NA) - 1) & 1));     // slab free again
      mbar_wait(&raw_full[sr], (uint32_t)((i / NR) & 1));
      const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)sr * raw_bytes);
      unsigned char* ab = ab0 + (size_t)sa * ab_bytes;
      const bool edge = t0 - d < 0 || t0 + kRows + d > P.T;
      const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;      // time of raw column 0
      const int col0 = t0 - d - start;                      // raw column of slab position 0 (negative on the first tile)
      int u = ct, c8 = 0;
      while (u >= Wpos) { u -= Wpos; ++c8; }
      for (int it = ct; it < nitems; it += 128) {
        int us = col0 + u;
        if (edge) {                             // mirror (no edge repeat): t -> -t, t -> 2(T-1) - t
          int t = t0 - d + u;
          if (t < 0) t = -t;
          else if (t >= P.T) t = 2 * (P.T - 1) - t;
          us = t - start;
          us = us < 0 ? 0 : (us >= Wraw ? Wraw - 1 : us);   // (rows beyond T + d are never stored)
        }
        const float* src = raw + (size_t)(c8 * 8) * Wraw + us;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[e * Wraw];
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          const float2 hf = __bfloat1622float2(h2);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
          hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        unsigned char* dst = ab + (size_t)(c8 >> 1) * Wpos * 64 + (size_t)(c8 & 1) * Wpos * 16 + (size_t)u * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + (size_t)Wpos * 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        u += 128;
        while (u >= Wpos) { u -= Wpos; ++c8; }
      }
      fence_proxy_async();
      warp_arrive(&a1_full[sa], lane);
    }
  } else if (warp < 10) {
    // ===================== mid: D1 -> bf16 hi/lo A operand of the pointwise conv (+ optional h) =====================
    const int q = warp & 3, m = q * 32 + lane;
    for (int i = 0; i < my_tiles; ++i) {
      const int sa = i % NA;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
      mbar_wait(&mma1_done[sa], (uint32_t)((i / NA) & 1));
      tc_fence_after();
      unsigned char* ab = ab0 + (size_t)sa * ab_bytes;
      const uint32_t dcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * C);
      const bool hv = P.h != nullptr && t0 + m < P.T;
      float* hp = P.h ? P.h + ((size_t)b * C) * P.T + t0 + m : nullptr;
      for (int cg = 0; cg < ncg; ++cg) {
        float v[16];
        if (P.dbg & 8) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        } else {
          tmem_ld16(dcol + (uint32_t)(cg * 16), v);
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          const float2 hf = __bfloat1622float2(h2);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
          hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        unsigned char* dst = ab + (size_t)cg * 8192 + (size_t)m * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(dst + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(dst + 4096 + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        if (hv) {
#pragma unroll
          for (int e = 0; e < 16; ++e) hp[(size_t)(cg * 16 + e) * P.T] = v[e];
        }
      }
      fence_proxy_async();
      tc_fence_before();
      warp_arrive(&a2_full[sa], lane);
    }
  } else {
    // ===================== epilogue: D2 -> LeakyReLU -> + x (raw tile) -> out =====================
    const int q = warp & 3, m = q * 32 + lane;
    const float slope = P.slope;
    for (int i = 0; i < my_tiles; ++i) {
      const int sr = i % NR, sa = i % NA, j = i & 1;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
      mbar_wait(&mma2_done[sa], (uint32_t)((i / NA) & 1));
      tc_fence_after();
      const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;
      const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)sr * raw_bytes) + (t0 - start) + m;
      const uint32_t dcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * C + j * C);
      const bool ev = t0 + m < P.T;
      const size_t o = ((size_t)b * C) * P.T + t0 + m;
      for (int cg = 0; cg < ncg; ++cg) {
        float v[16];
        if (P.dbg & 16) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        } else {
          tmem_ld16(dcol + (uint32_t)(cg * 16), v);
        }
        if (cg == ncg - 1) {                    // D2[j] drained: the pointwise conv of tile i+2 may overwrite it
          tc_fence_before();
          warp_arrive(&d2_free[j], lane);
        }
        float r[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = raw[(size_t)(cg * 16 + e) * Wraw];
        if (ev) {
          if (P.mask) {
            unsigned char* mp = P.mask + o + (size_t)(cg * 16) * P.T;
#pragma unroll
            for (int e = 0; e < 16; ++e) mp[(size_t)e * P.T] = v[e] > 0.f ? 1 : 0;
          }
          float* yp = P.out + o + (size_t)(cg * 16) * P.T;
#pragma unroll
          for (int e = 0; e < 16; ++e) yp[(size_t)e * P.T] = (v[e] > 0.f ? v[e] : v[e] * slope) + r[e];
        }
      }
      warp_arrive(&raw_free[sr], lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// Weight pre-pack: one thread per 16-byte unit (8 consecutive input channels of one output channel and tap).
// Blob = W1 tiles [16-channel group][tap] then W2 tiles [16-channel group]; tile = [hi|lo][half][n][8] (C*64 bytes).
__global__ void ru_pack_kernel(const float* __restrict__ w1, const float* __restrict__ w2, unsigned char* __restrict__ out,
                               int C) {
  const int ncg = C / 16;
  const int units = ncg * 4 * 2 * C;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int n = u % C;
    int r = u / C;
    const int hf = r % 2; r /= 2;
    const int ti = r;                                        // tile index: W1 (cg*3 + tap), then W2 (ncg*3 + cg)
    const bool first = ti < ncg * 3;
    const int cg = first ? ti / 3 : ti - ncg * 3, tap = first ? ti % 3 : 0;
    const int c0 = cg * 16 + hf * 8;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = first ? w1[((size_t)n * C + c0 + e) * 3 + tap] : w2[(size_t)n * C + c0 + e];
      split_bf16(v, hi[e], lo[e]);
    }
    unsigned char* tile = out + (size_t)ti * C * 64;
    const size_t off = ((size_t)hf * C + n) * 16;
    *reinterpret_cast<uint4*>(tile + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(tile + (size_t)C * 32 + off) = *reinterpret_cast<const uint4*>(lo);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

static bool plan(RuP& P, int B, int C, int T, int d) {
  if (C % 16 || C < 16 || C > 64 || T % 4 || d < 1 || d > 16 || T < d + 1 || B < 1) return false;
  P.B = B; P.C = C; P.T = T; P.d = d;
  P.Wpos = kRows + 2 * d;
  P.dA = (d + 3) & ~3;
  P.Wraw = kRows + 2 * P.dA;
  P.ncg = C / 16;
  P.tpi = (T + kRows - 1) / kRows;
  if ((long long)B * P.tpi >= (1ll << 31) || (long long)B * C * T >= (1ll << 31)) return false;
  P.ntiles = B * P.tpi;
  static const int env_nr = getenv("VBX_RU_NR") ? atoi(getenv("VBX_RU_NR")) : 0;
  static const int env_na = getenv("VBX_RU_NA") ? atoi(getenv("VBX_RU_NA")) : 0;
  // deepest rings that still leave two CTAs per SM; else the deepest that fit one
  static const int order[6][2] = {{3, 2}, {2, 2}, {3, 1}, {2, 1}, {1, 1}, {0, 0}};
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; order[i][0]; ++i) {
      P.NR = env_nr ? env_nr : order[i][0];
      P.NA = env_na ? env_na : order[i][1];
      if (P.NA > 2 || P.NR < 1 || P.NA < 1) return false;
      const size_t lim = pass == 0 ? (size_t)(kSmemLimit / 2 - 1024) : (size_t)kSmemLimit;
      if (ru_smem_bytes(P) <= lim) {
        int cols = 32;
        while (cols < 4 * C) cols <<= 1;
        P.tmem_cols = cols;
        return true;
      }
    }
  }
  return false;
}

}  // namespace ru
}  // namespace vbx

using namespace vbx;
using namespace vbx::ru;

extern "C" int vbx_ru_supported(int32_t B, int32_t C, int32_t T, int32_t dil) {
  RuP P;
  return plan(P, B, C, T, dil) ? 1 : 0;
}
extern "C" int64_t vbx_ru_pack_bytes(int32_t C) { return (C % 16 || C < 16 || C > 64) ? -1 : (int64_t)ru_w_bytes(C); }

extern "C" int vbx_ru_pack(int32_t C, const float* w_dil, const float* w_pw, void* packed, void* stream) {
  VBX_REQUIRE(C % 16 == 0 && C >= 16 && C <= 64, VBX_UNSUPPORTED, "ru_pack: C must be 16, 32, 48 or 64");
  VBX_REQUIRE(w_dil && w_pw && packed, VBX_BAD_POINTER, "ru_pack: null tensor");
  VBX_REQUIRE(((uintptr_t)packed & 15) == 0, VBX_BAD_POINTER, "ru_pack: packed buffer must be 16-byte aligned");
  const int units = (C / 16) * 4 * 2 * C;
  ru_pack_kernel<<<(units + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_dil, w_pw, (unsigned char*)packed, C);
  return launched("ru_pack_kernel");
}

extern "C" int vbx_ru_fwd(int32_t B, int32_t C, int32_t T, int32_t dil, float slope, const float* x, const void* packed,
                          float* out, float* h, uint8_t* mask, void* stream) {
  RuP P;
  VBX_REQUIRE(plan(P, B, C, T, dil), VBX_UNSUPPORTED,
              "ru_fwd: unsupported shape (C in {16..64} multiple of 16, T % 4 == 0, 1 <= dil <= 16, T > dil)");
  VBX_REQUIRE(x && packed && out, VBX_BAD_POINTER, "ru_fwd: null tensor");
  VBX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)packed & 15) == 0, VBX_BAD_POINTER,
              "ru_fwd: x and the packed weights must be 16-byte aligned");
  EncodeTiledFn enc = encode_fn();
  VBX_REQUIRE(enc != nullptr, VBX_UNSUPPORTED, "ru_fwd: cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)T * 4, (cuuint64_t)C * T * 4};
  const cuuint32_t box[3] = {(cuuint32_t)P.Wraw, (cuuint32_t)C, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "ru_fwd: cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);
    return VBX_UNSUPPORTED;
  }
  static const int dbg = getenv("VBX_RU_DBG") ? atoi(getenv("VBX_RU_DBG")) : 0;
  P.dbg = dbg;
  P.slope = slope; P.packed = (const unsigned char*)packed; P.out = out; P.h = h; P.mask = mask;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(ru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (ce != cudaSuccess) return fail((int)ce, "ru_fwd: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  const size_t smem = ru_smem_bytes(P);
  const int occ = smem <= (size_t)(kSmemLimit / 2 - 1024) && 2 * P.tmem_cols <= 512 ? 2 : 1;
  int grid = 148 * occ;
  if (grid > P.ntiles) grid = P.ntiles;
  ru_fwd_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tm, P);
  return launched("ru_fwd_kernel");
}
