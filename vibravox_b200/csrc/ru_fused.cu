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
//   MMA      one thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) for the dilated conv into D1, later for the
//            pointwise conv into D2, whichever operand is ready first.  The three products of the split (hi*hi + hi*lo +
//            lo*hi) take TWO instructions per (16-channel group, tap): the weight tile holds [W_hi | W_lo] side by side as
//            2C columns, so x_hi * [W_hi | W_lo] is one N = 2C MMA and x_lo * W_hi one N = C MMA onto the first C
//            columns; the consumer adds the two column halves (fewer instructions for the issuing thread and a third
//            fewer shared-memory reads of the A operand - both measured as the kernel's limiter);
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
// warps of a CTA: 0 TMA, 1 MMA, then CW convert warps, MW mid warps, EW epilogue warps (4 or 8 each: 4 when two CTAs
// share an SM, 8 when the CTA owns it - two warps per TMEM lane quadrant then split the 16-channel groups)
__host__ __device__ constexpr int ru_threads(int CW, int MW, int EW) { return (2 + CW + MW + EW) * 32; }
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
  int wait_ns;      // suspend-time hint of the mbarrier waits (VBX_RU_WAIT_NS, 0 = plain spin)
  float slope;
  int res_global;   // 1: the epilogue re-reads the residual from global memory (L2) and the raw tile is released
                    //    right after the conversion (shallow raw rings, C = 64); 0: residual from the raw tile
  const float* x;
  const unsigned char* packed;
  float* out;
  float* h;         // optional
  unsigned char* mask;   // optional
  long long* prof;  // optional (vbx_ru_set_profile_buffer): CTA 0 records clock64() per tile and pipeline event
};
#define RU_PROF(ev) do { if (P.prof && blockIdx.x == 0 && lane == 0) P.prof[(i) * 16 + (ev)] = clock64(); } while (0)

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

// v = D[cols a .. a+15] + D[cols b .. b+15] of this thread's TMEM lane: both loads are issued before the one wait
__device__ __forceinline__ void tmem_ld16_sum2(uint32_t ta, uint32_t tb, float* v) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(tb)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(q[i]);
}

template <int CW, int MW, int EW, int MINB>
__global__ void __launch_bounds__(ru_threads(CW, MW, EW), MINB) ru_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const RuP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t wns = (uint32_t)P.wait_ns;
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
    for (int i = 0; i < NR; ++i) {             // released by the epilogue warps, or by the convert warps (res_global)
      mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], P.res_global ? CW : EW);
    }
    for (int i = 0; i < NA; ++i) {
      mbar_init(&a1_full[i], CW); mbar_init(&mma1_done[i], 1); mbar_init(&a2_full[i], MW); mbar_init(&mma2_done[i], 1);
    }
    mbar_init(&d2_free[0], EW); mbar_init(&d2_free[1], EW);
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: D1[j] at j*2C, D2[j] at 4C + j*2C   (j = tile parity; each accumulator is [x_hi W_hi + x_lo W_hi | x_hi W_lo])

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && my_tiles > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
      mbar_expect_tx(w_full, (uint32_t)w_bytes);
      for (int off = 0; off < w_bytes; off += 32768) {
        const int n = w_bytes - off < 32768 ? w_bytes - off : 32768;
        bulk_copy_g2s(w0 + off, P.packed + off, (uint32_t)n, w_full);
      }
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % NR;
        const uint32_t par = (uint32_t)((i / NR) & 1);
        mbar_wait_hint(&raw_free[s], par ^ 1u, wns);
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
        mbar_expect_tx(&raw_full[s], (uint32_t)raw_bytes);
        const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;
        RU_PROF(0);
        tma_load_3d(raw0 + (size_t)s * raw_bytes, &tmap_x, &raw_full[s], start, 0, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the (warp-uniform) control flow so that the descriptor arithmetic stays in uniform
    // registers; lane 0 alone issues tcgen05.mma / commit.  A descriptor is a constant 64-bit pattern plus the
    // operand address >> 4 in its low bits (all offsets are multiples of 16 bytes, addresses < 256 KB: no carry into
    // the stride fields), so per MMA it costs a 32-bit add instead of being rebuilt - measured: with the descriptors
    // rebuilt per MMA by one thread, this warp was the bottleneck of the whole CTA.
    if (my_tiles > 0) {
      const uint32_t idesc2 = make_idesc_bf16(2 * C, /*a_mn=*/false, /*b_mn=*/false);   // x_hi * [W_hi | W_lo]
      const uint32_t idesc1 = make_idesc_bf16(C, /*a_mn=*/false, /*b_mn=*/false);       // x_lo * W_hi
      const uint32_t lbo_a1 = (uint32_t)Wpos * 16, lbo_b = (uint32_t)C * 32;
      const uint32_t wbase = smem_u32(w0);
      // low / high words of the descriptors at offset 0 of each operand region
      const uint64_t dA1 = make_desc(0, lbo_a1, 128), dA2 = make_desc(0, 2048, 128), dB = make_desc(0, lbo_b, 128);
      const uint32_t a1_hi32 = (uint32_t)(dA1 >> 32), a1_lo32 = (uint32_t)dA1;
      const uint32_t a2_hi32 = (uint32_t)(dA2 >> 32), a2_lo32 = (uint32_t)dA2;
      const uint32_t b_hi32 = (uint32_t)(dB >> 32), b_lo32 = (uint32_t)dB + (wbase >> 4);
      const uint32_t pa1 = (uint32_t)Wpos * 2, tb = (uint32_t)C * 4;   // lo plane of the slab / weight tile, in 16-byte units
      const uint32_t w2u = (uint32_t)ncg * 3u * tb;
      auto mk = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      // warp-uniform copy of the TMEM base (a value loaded from shared memory lives in a vector register, and a
      // vector-register operand makes ptxas wrap every tcgen05.mma in an ELECT / R2UR / branch "waterfall" of ~10
      // dependent instructions; a warp reduction result is uniform by construction)
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
      mbar_wait_hint(w_full, 0, wns);
      int n1 = 0, n2 = 0;                       // next tile of the dilated / pointwise conv
      while (n2 < my_tiles) {
        bool did = false;
        if (n1 < my_tiles && n1 - n2 < 2) {     // (D1 / D2 are two deep)
          const int sa = n1 % NA;
          if (mbar_test_wait(&a1_full[sa], (uint32_t)((n1 / NA) & 1))) {
            tc_fence_after();
            const uint32_t dcol = tmem_u + (uint32_t)((n1 & 1) * 2 * C);
            uint32_t a_u = a1_lo32 + (smem_u32(ab0 + (size_t)sa * ab_bytes) >> 4);
            uint32_t b_u = b_lo32;
            if (lane == 0) {
              uint32_t acc = 0;
              for (int cg = 0; cg < ncg; ++cg) {
#pragma unroll
                for (int tap = 0; tap < 3; ++tap) {
                  const uint32_t at = a_u + (uint32_t)(tap * d);
                  const uint64_t db = mk(b_hi32, b_u);
                  mma_bf16_ss(dcol, mk(a1_hi32, at), db, idesc2, acc);
                  acc = 1;
                  mma_bf16_ss(dcol, mk(a1_hi32, at + pa1), db, idesc1, 1);
                  b_u += tb;
                }
                a_u += (uint32_t)Wpos * 4;
              }
              mma_commit(&mma1_done[sa]);
              { const int i = n1; RU_PROF(4); }
            }
            __syncwarp();
            ++n1;
            did = true;
          }
        }
        if (n2 < n1) {
          const int sa = n2 % NA, j = n2 & 1;
          if (mbar_test_wait(&a2_full[sa], (uint32_t)((n2 / NA) & 1))) {
            if (n2 >= 2) mbar_wait_hint(&d2_free[j], (uint32_t)(((n2 >> 1) - 1) & 1), wns);
            tc_fence_after();
            const uint32_t dcol = tmem_u + (uint32_t)(4 * C + j * 2 * C);
            uint32_t a_u = a2_lo32 + (smem_u32(ab0 + (size_t)sa * ab_bytes) >> 4);
            uint32_t b_u = b_lo32 + w2u;
            if (lane == 0) {
              uint32_t acc = 0;
              for (int cg = 0; cg < ncg; ++cg) {
                const uint64_t db = mk(b_hi32, b_u);
                mma_bf16_ss(dcol, mk(a2_hi32, a_u), db, idesc2, acc);
                acc = 1;
                mma_bf16_ss(dcol, mk(a2_hi32, a_u + 256u), db, idesc1, 1);
                a_u += 512u;
                b_u += tb;
              }
              mma_commit(&mma2_done[sa]);
              { const int i = n2; RU_PROF(7); }
            }
            __syncwarp();
            ++n2;
            did = true;
          }
        }
        if (!did) __nanosleep(20);
      }
    }
  } else if (warp < 2 + CW) {
    // ===================== convert: raw fp32 tile -> K-major bf16 hi/lo slab (+ reflect halo) =====================
    const int ct = tid - 64;                    // 0 .. CW*32-1
    constexpr int kCT = CW * 32;
    const int nitems = (C / 8) * Wpos;          // (8-channel unit, position)
    for (int i = 0; i < my_tiles; ++i) {
      const int sr = i % NR, sa = i % NA;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int t0 = (tile % P.tpi) * kRows;
      if (i >= NA) mbar_wait_hint(&mma2_done[sa], (uint32_t)(((i / NA) - 1) & 1), wns);     // slab free again
      if (warp == 2) RU_PROF(1);
      mbar_wait_hint(&raw_full[sr], (uint32_t)((i / NR) & 1), wns);
      if (warp == 2) RU_PROF(2);
      const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)sr * raw_bytes);
      unsigned char* ab = ab0 + (size_t)sa * ab_bytes;
      const bool edge = t0 - d < 0 || t0 + kRows + d > P.T;
      const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;      // time of raw column 0
      const int col0 = t0 - d - start;                      // raw column of slab position 0 (negative on the first tile)
      int u = ct, c8 = 0;
      while (u >= Wpos) { u -= Wpos; ++c8; }
      for (int it = ct; it < nitems; it += kCT) {
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
        u += kCT;
        while (u >= Wpos) { u -= Wpos; ++c8; }
      }
      fence_proxy_async();
      warp_arrive(&a1_full[sa], lane);
      if (P.res_global) warp_arrive(&raw_free[sr], lane);     // (the epilogue will not touch the raw tile)
      if (warp == 2) RU_PROF(3);
    }
  } else if (warp < 2 + CW + MW) {
    // ===================== mid: D1 -> bf16 hi/lo A operand of the pointwise conv (+ optional h) =====================
    const int q = warp & 3, m = q * 32 + lane;
    const int mset = (warp - (2 + CW)) >> 2;    // 0, or 0 / 1 when two warps share a lane quadrant
    for (int i = 0; i < my_tiles; ++i) {
      const int sa = i % NA;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
      mbar_wait_hint(&mma1_done[sa], (uint32_t)((i / NA) & 1), wns);
      tc_fence_after();
      if (warp == 2 + CW) RU_PROF(5);
      unsigned char* ab = ab0 + (size_t)sa * ab_bytes;
      const uint32_t dcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * 2 * C);
      const bool hv = P.h != nullptr && t0 + m < P.T;
      float* hp = P.h ? P.h + ((size_t)b * C) * P.T + t0 + m : nullptr;
      for (int cg = mset; cg < ncg; cg += MW / 4) {
        float v[16];
        tmem_ld16_sum2(dcol + (uint32_t)(cg * 16), dcol + (uint32_t)(C + cg * 16), v);
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
      if (warp == 2 + CW) RU_PROF(6);
    }
  } else {
    // ===================== epilogue: D2 -> LeakyReLU -> + x (raw tile) -> out =====================
    const int q = warp & 3, m = q * 32 + lane;
    const int eset = (warp - (2 + CW + MW)) >> 2;
    const float slope = P.slope;
    for (int i = 0; i < my_tiles; ++i) {
      const int sr = i % NR, sa = i % NA, j = i & 1;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
      mbar_wait_hint(&mma2_done[sa], (uint32_t)((i / NA) & 1), wns);
      tc_fence_after();
      if (warp == 2 + CW + MW) RU_PROF(8);
      const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;
      const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)sr * raw_bytes) + (t0 - start) + m;
      const uint32_t dcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(4 * C + j * 2 * C);
      const bool ev = t0 + m < P.T;
      const float* xres = P.x + ((size_t)b * C) * P.T + t0 + m;
      const size_t o = ((size_t)b * C) * P.T + t0 + m;
      const int cg_last = ncg - 1 - ((ncg - 1 - eset) % (EW / 4));     // last group this warp handles
      for (int cg = eset; cg < ncg; cg += EW / 4) {
        float v[16];
        tmem_ld16_sum2(dcol + (uint32_t)(cg * 16), dcol + (uint32_t)(C + cg * 16), v);
        if (cg == cg_last) {                    // D2[j] drained: the pointwise conv of tile i+2 may overwrite it
          tc_fence_before();
          warp_arrive(&d2_free[j], lane);
        }
        float r[16];
        if (P.res_global) {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = ev ? __ldg(xres + (size_t)(cg * 16 + e) * P.T) : 0.f;
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = raw[(size_t)(cg * 16 + e) * Wraw];
        }
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
      if (!P.res_global) warp_arrive(&raw_free[sr], lane);
      if (warp == 2 + CW + MW) RU_PROF(9);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// Weight pre-pack: one thread per (8 consecutive input channels of one output channel and tap) -> two 16-byte units.
// Blob = W1 tiles [16-channel group][tap] then W2 tiles [16-channel group]; tile = [half][2C rows][8] (C*64 bytes) with
// rows 0..C-1 = bf16 hi of output channel n, rows C..2C-1 = bf16 lo of output channel n - C.
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
    const size_t off = ((size_t)hf * 2 * C + n) * 16;
    *reinterpret_cast<uint4*>(tile + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(tile + (size_t)C * 16 + off) = *reinterpret_cast<const uint4*>(lo);
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
        while (cols < 8 * C) cols <<= 1;
        P.tmem_cols = cols;
        static const int env_rg = getenv("VBX_RU_RESG") ? atoi(getenv("VBX_RU_RESG")) : -1;
        P.res_global = env_rg >= 0 ? env_rg : (P.NR < 3 ? 1 : 0);
        return true;
      }
    }
  }
  return false;
}


// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of the residual-unit convs (stride 1, C -> C, K = 3 dilated with reflect halo or K = 1):
//     dW[co, ci, k] = sum_{b,t} dy[b, co, t] * x[b, ci, mirror(t + (k-1)*d)]
// on the same skeleton: TMA tensor-map loads of the fp32 x / dy tiles (128 time positions per tile), conversion to
// bf16 hi/lo slabs [plane][8-channel group][position] (16-byte units, positions 16 bytes apart), which the tensor core
// reads MN-major with the reduction running over time: per tap D_k[co (M = 128 lanes, the first C real), ci] +=
// dy_hi * [x_hi | x_lo] (N = 2C) + dy_lo * x_hi (N = C); the x slab serves every tap through the descriptor start
// address (+ k*d units).  The accumulators stay in TMEM for the whole life of the persistent CTA; at the end each CTA
// writes its partial dW to a workspace and a second tiny kernel sums the partials in a FIXED order (deterministic,
// no atomics).  Two issuing threads (taps 0, 2 / tap 1): one thread cannot issue small-N MMAs faster than one per
// ~76 cycles (profiles/r1_mma_probe.txt) and 48 of them per tile would be the kernel's limiter.
struct WgP {
  int B, C, T, d, K;
  int Wpos, dA, Wraw;   // x slab positions (128 + (K-1)*d), aligned halo, raw x columns
  int xu, yu;           // 16-byte units between consecutive 8-channel groups of the x / dy slab (odd: conflict-free MMA fetch)
  int tpi, ntiles, NR, NA, tmem_cols, wait_ns, nissue;
  float* partial;       // [gridDim.x][C][C][K]
};
static const int kWgThreads = 11 * 32;      // warps: 0 TMA, 1-2 MMA issuers, 3-6 convert x, 7-10 convert dy + final drain

__host__ __device__ inline int wg_x_slab(const WgP& P) { return 2 * (P.C / 8) * P.xu * 16; }
__host__ __device__ inline int wg_y_slab(const WgP& P) { return 2 * (P.C / 8) * P.yu * 16; }
__host__ __device__ inline int wg_raw(const WgP& P) { return P.C * (P.Wraw + kRows) * 4; }
static size_t wg_smem_bytes(const WgP& P) {
  return (size_t)P.NA * (wg_y_slab(P) + wg_x_slab(P)) + (size_t)P.NR * wg_raw(P) + (size_t)(2 * P.NR + 2 * P.NA + 1) * 8 + 16;
}

template <int MINB>
__global__ void __launch_bounds__(kWgThreads, MINB) ru_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                     const __grid_constant__ CUtensorMap tmap_dy, const WgP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t wns = (uint32_t)P.wait_ns;
  const int C = P.C, d = P.d, K = P.K, Wpos = P.Wpos, Wraw = P.Wraw, NR = P.NR, NA = P.NA, nc8 = C / 8;
  const int ysl = wg_y_slab(P), xsl = wg_x_slab(P), rawb = wg_raw(P), rawx = C * Wraw * 4;
  unsigned char* y0 = smem;                                   // dy slabs first: their M = 128 descriptor reads run past
  unsigned char* x0 = y0 + (size_t)NA * ysl;                  // the C real channels into whatever follows
  unsigned char* raw0 = x0 + (size_t)NA * xsl;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw0 + (size_t)NR * rawb);
  uint64_t* raw_full = bars;                 // [NR] TMA bytes (x + dy tile) landed
  uint64_t* raw_free = raw_full + NR;        // [NR] 8 convert warps done
  uint64_t* slab_full = raw_free + NR;       // [NA] 8 convert warps: both slabs staged
  uint64_t* slab_free = slab_full + NA;      // [NA] every issuing thread's MMAs of the tile retired
  uint64_t* acc_done = slab_free + NA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
  const int my_tiles = ((int)blockIdx.x < P.ntiles) ? (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < NR; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], 8); }
    for (int i = 0; i < NA; ++i) { mbar_init(&slab_full[i], 8); mbar_init(&slab_free[i], P.nissue); }
    mbar_init(acc_done, P.nissue);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && my_tiles > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_dy)) : "memory");
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % NR;
        mbar_wait_hint(&raw_free[s], (uint32_t)((i / NR) & 1) ^ 1u, wns);
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = tile / P.tpi, t0 = (tile % P.tpi) * kRows;
        const int start = t0 - P.dA > 0 ? t0 - P.dA : 0;
        mbar_expect_tx(&raw_full[s], (uint32_t)rawb);
        tma_load_3d(raw0 + (size_t)s * rawb, &tmap_x, &raw_full[s], start, 0, b);
        tma_load_3d(raw0 + (size_t)s * rawb + rawx, &tmap_dy, &raw_full[s], t0, 0, b);
      }
    }
  } else if (warp < 3) {
    // ===================== MMA issuers: issuer j takes taps j, j + 2 =====================
    const int j = warp - 1;
    if (my_tiles > 0 && j < P.nissue) {
      const uint32_t idesc2 = make_idesc_bf16(2 * C, /*a_mn=*/true, /*b_mn=*/true);
      const uint32_t idesc1 = make_idesc_bf16(C, /*a_mn=*/true, /*b_mn=*/true);
      const uint64_t dY = make_desc(0, 128, (uint32_t)P.yu * 16), dX = make_desc(0, 128, (uint32_t)P.xu * 16);
      const uint32_t y_hi32 = (uint32_t)(dY >> 32), y_lo32 = (uint32_t)dY, x_hi32 = (uint32_t)(dX >> 32), x_lo32 = (uint32_t)dX;
      const uint32_t ypl = (uint32_t)(nc8 * P.yu);            // lo plane of the dy slab, in 16-byte units
      auto mk = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
      for (int i = 0; i < my_tiles; ++i) {
        const int sa = i % NA;
        mbar_wait_hint(&slab_full[sa], (uint32_t)((i / NA) & 1), wns);
        tc_fence_after();
        const uint32_t yu0 = y_lo32 + (smem_u32(y0 + (size_t)sa * ysl) >> 4);
        const uint32_t xu0 = x_lo32 + (smem_u32(x0 + (size_t)sa * xsl) >> 4);
        if (lane == 0) {
          for (int tap = j; tap < K; tap += 2) {
            const uint32_t dcol = tmem_u + (uint32_t)(tap * 2 * C);
            uint32_t yk = yu0, xk = xu0 + (uint32_t)(tap * d);
#pragma unroll
            for (int ks = 0; ks < kRows / 16; ++ks) {
              const uint64_t db = mk(x_hi32, xk);
              mma_bf16_ss(dcol, mk(y_hi32, yk), db, idesc2, (i > 0 || ks > 0) ? 1u : 0u);
              mma_bf16_ss(dcol, mk(y_hi32, yk + ypl), db, idesc1, 1);
              yk += 16; xk += 16;                             // 16 positions = 16 units further on
            }
          }
          mma_commit(&slab_free[sa]);
          if (i == my_tiles - 1) mma_commit(acc_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== convert: x (warps 3-6, reflect halo folded in) / dy (warps 7-10) =====================
    const bool is_x = warp < 7;
    const int ct = (warp - (is_x ? 3 : 7)) * 32 + lane;         // 0..127
    const int npos = is_x ? Wpos : kRows, gu = is_x ? P.xu : P.yu, rw = is_x ? Wraw : kRows;
    const int nitems = nc8 * npos;
    for (int i = 0; i < my_tiles; ++i) {
      const int sr = i % NR, sa = i % NA;
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int t0 = (tile % P.tpi) * kRows;
      if (i >= NA) mbar_wait_hint(&slab_free[sa], (uint32_t)(((i / NA) - 1) & 1), wns);
      mbar_wait_hint(&raw_full[sr], (uint32_t)((i / NR) & 1), wns);
      const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)sr * rawb + (is_x ? 0 : rawx));
      unsigned char* slab = is_x ? x0 + (size_t)sa * xsl : y0 + (size_t)sa * ysl;
      const int halo = (K - 1) * d / 2;                          // positions of the x slab start at t0 - halo
      const bool edge = is_x && (t0 - halo < 0 || t0 + kRows + halo > P.T);
      const int start = is_x ? (t0 - P.dA > 0 ? t0 - P.dA : 0) : t0;
      const int col0 = is_x ? t0 - halo - start : 0;
      int u = ct, c8 = 0;
      while (u >= npos) { u -= npos; ++c8; }
      for (int it = ct; it < nitems; it += 128) {
        int us = col0 + u;
        if (edge) {
          int t = t0 - halo + u;
          if (t < 0) t = -t;
          else if (t >= P.T) t = 2 * (P.T - 1) - t;
          us = t - start;
          us = us < 0 ? 0 : (us >= rw ? rw - 1 : us);
        }
        const float* src = raw + (size_t)(c8 * 8) * rw + us;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[e * rw];
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          const float2 hf = __bfloat1622float2(h2);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
          hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        unsigned char* dst = slab + ((size_t)c8 * gu + u) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + (size_t)nc8 * gu * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        u += 128;
        while (u >= npos) { u -= npos; ++c8; }
      }
      fence_proxy_async();
      warp_arrive(&slab_full[sa], lane);
      warp_arrive(&raw_free[sr], lane);
    }
    // ===================== final drain (dy warps): this CTA's partial dW -> workspace =====================
    if (!is_x && my_tiles > 0) {
      const int q = warp & 3, co = q * 32 + lane;
      if (q * 32 < C) {                                          // (warp-uniform)
        mbar_wait_hint(acc_done, 0, wns);
        tc_fence_after();
        float* dst = P.partial + (size_t)blockIdx.x * C * C * K + (size_t)co * C * K;
        for (int tap = 0; tap < K; ++tap) {
          for (int blk = 0; blk < C / 16; ++blk) {
            float v[16];
            const uint32_t col = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * 2 * C + blk * 16);
            tmem_ld16_sum2(col, col + (uint32_t)C, v);
            if (co < C) {
#pragma unroll
              for (int e = 0; e < 16; ++e) dst[(size_t)(blk * 16 + e) * K + tap] = v[e];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// dw[i] = beta * dw[i] + sum over CTAs of partial[cta][i] in a FIXED order: 32 outputs x 8 slices per block, slice s sums
// the partials p = s, s + 8, ... (coalesced 128-byte rows), the 8 slice sums are then added in slice order
__global__ void __launch_bounds__(256) ru_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int n,
                                                              int nparts, float beta) {
  __shared__ float sh[8][32];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + o;
  float acc = 0.f;
  if (i < n)
    for (int p = sl; p < nparts; p += 8) acc += partial[(size_t)p * n + i];
  sh[sl][o] = acc;
  __syncthreads();
  if (sl == 0 && i < n) {
    float tot = sh[0][o];
#pragma unroll
    for (int k = 1; k < 8; ++k) tot += sh[k][o];
    dw[i] = beta != 0.f ? beta * dw[i] + tot : tot;
  }
}

static bool plan_wg(WgP& P, int B, int C, int T, int d, int K, int* grid_out, int* occ_out) {
  if (C % 32 || C < 32 || C > 64 || T % 4 || (K != 1 && K != 3) || d < 1 || d > 16 || T < d + 1 || B < 1) return false;
  if (K == 1) d = 1;
  P.B = B; P.C = C; P.T = T; P.d = d; P.K = K;
  const int halo = (K - 1) * d / 2;
  P.Wpos = kRows + 2 * halo;
  P.dA = (halo + 3) & ~3;
  P.Wraw = kRows + 2 * P.dA;
  P.xu = P.Wpos | 1; P.yu = kRows | 1;
  P.tpi = (T + kRows - 1) / kRows;
  if ((long long)B * P.tpi >= (1ll << 31) || (long long)B * C * T >= (1ll << 31)) return false;
  P.ntiles = B * P.tpi;
  P.nissue = K == 1 ? 1 : 2;
  int cols = 32;
  while (cols < 2 * C * K) cols <<= 1;
  P.tmem_cols = cols;
  static const int env_nr = getenv("VBX_RUW_NR") ? atoi(getenv("VBX_RUW_NR")) : 0;
  static const int env_na = getenv("VBX_RUW_NA") ? atoi(getenv("VBX_RUW_NA")) : 0;
  static const int order[5][2] = {{2, 1}, {3, 2}, {2, 2}, {2, 1}, {0, 0}};     // first entry: two CTAs per SM
  for (int i = 0; order[i][0]; ++i) {
    P.NR = env_nr ? env_nr : order[i][0];
    P.NA = env_na ? env_na : order[i][1];
    const bool two = i == 0 && 2 * P.tmem_cols <= 512;
    const size_t lim = two ? (size_t)(kSmemLimit / 2 - 1024) : (size_t)kSmemLimit;
    if (i == 0 && !two) continue;
    if (wg_smem_bytes(P) <= lim) {
      const int occ = two ? 2 : 1;
      int grid = 148 * occ;
      if (grid > P.ntiles) grid = P.ntiles;
      *grid_out = grid; *occ_out = occ;
      return true;
    }
  }
  return false;
}
}  // namespace ru
}  // namespace vbx

using namespace vbx;
using namespace vbx::ru;

static long long* g_ru_prof = nullptr;
extern "C" int vbx_ru_set_profile_buffer(void* buf) { g_ru_prof = (long long*)buf; return 0; }

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
  static const int wait_ns = getenv("VBX_RU_WAIT_NS") ? atoi(getenv("VBX_RU_WAIT_NS")) : 20000;
  P.wait_ns = wait_ns;
  P.prof = g_ru_prof;
  P.x = x;
  P.slope = slope; P.packed = (const unsigned char*)packed; P.out = out; P.h = h; P.mask = mask;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(ru_fwd_kernel<4, 4, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(ru_fwd_kernel<8, 8, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (ce != cudaSuccess) return fail((int)ce, "ru_fwd: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  const size_t smem = ru_smem_bytes(P);
  const int occ = smem <= (size_t)(kSmemLimit / 2 - 1024) && 2 * P.tmem_cols <= 512 ? 2 : 1;
  int grid = 148 * occ;
  if (grid > P.ntiles) grid = P.ntiles;
  static const int wide = getenv("VBX_RU_WIDE") ? atoi(getenv("VBX_RU_WIDE")) : 1;
  if (occ == 1 && wide && P.ncg >= 2)        // the CTA owns the SM: twice the warps per pipeline stage
    ru_fwd_kernel<8, 8, 8, 1><<<grid, ru_threads(8, 8, 8), smem, (cudaStream_t)stream>>>(tm, P);
  else
    ru_fwd_kernel<4, 4, 4, 2><<<grid, ru_threads(4, 4, 4), smem, (cudaStream_t)stream>>>(tm, P);
  return launched("ru_fwd_kernel");
}

extern "C" int64_t vbx_ru_wgrad_workspace(int32_t B, int32_t C, int32_t T, int32_t dil, int32_t K) {
  WgP P; int grid, occ;
  if (!plan_wg(P, B, C, T, dil, K, &grid, &occ)) return -1;
  return (int64_t)grid * C * C * K * 4;
}

extern "C" int vbx_ru_wgrad(int32_t B, int32_t C, int32_t T, int32_t dil, int32_t K, const float* x, const float* dy,
                            float* dw, float beta, void* workspace, void* stream) {
  WgP P; int grid, occ;
  VBX_REQUIRE(plan_wg(P, B, C, T, dil, K, &grid, &occ), VBX_UNSUPPORTED,
              "ru_wgrad: unsupported shape (C in {32, 64}, K in {1, 3}, T % 4 == 0, 1 <= dil <= 16, T > dil)");
  VBX_REQUIRE(x && dy && dw && workspace, VBX_BAD_POINTER, "ru_wgrad: null tensor");
  VBX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)workspace & 15) == 0, VBX_BAD_POINTER,
              "ru_wgrad: x, dy and the workspace must be 16-byte aligned");
  EncodeTiledFn enc = encode_fn();
  VBX_REQUIRE(enc != nullptr, VBX_UNSUPPORTED, "ru_wgrad: cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tmx, tmy;
  const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)T * 4, (cuuint64_t)C * T * 4};
  const cuuint32_t boxx[3] = {(cuuint32_t)P.Wraw, (cuuint32_t)C, 1}, boxy[3] = {(cuuint32_t)kRows, (cuuint32_t)C, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, boxx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr == CUDA_SUCCESS)
    cr = enc(&tmy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(dy), dims, strides, boxy, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "ru_wgrad: cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);
    return VBX_UNSUPPORTED;
  }
  static const int wait_ns = getenv("VBX_RU_WAIT_NS") ? atoi(getenv("VBX_RU_WAIT_NS")) : 20000;
  P.wait_ns = wait_ns;
  P.partial = (float*)workspace;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(ru_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(ru_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (ce != cudaSuccess) return fail((int)ce, "ru_wgrad: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  const size_t smem = wg_smem_bytes(P);
  if (occ == 2) ru_wgrad_kernel<2><<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(tmx, tmy, P);
  else ru_wgrad_kernel<1><<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(tmx, tmy, P);
  if (int r = launched("ru_wgrad_kernel")) return r;
  const int n = C * C * K;
  ru_wgrad_reduce_kernel<<<(n + 31) / 32, 256, 0, (cudaStream_t)stream>>>(P.partial, dw, n, grid, beta);
  return launched("ru_wgrad_reduce_kernel");
}
