// Tensor-core (tcgen05 / TMEM) implicit-GEMM Conv1d for the dense layers of the EBEN step.
//
// Orientation: rows of the MMA (M = 128) are output TIME positions, columns (N <= 256) are output
// channels, the reduction runs over (input channel, tap):
//     D[(b,t), co] = sum_{(ci,k)} A[(b,t), (ci,k)] * W[co, (ci,k)]
//   * A is the im2col view of the fp32 (B,C,T) activations.  It is never materialised in HBM: eight
//     producer warps gather it (lanes walk t => coalesced 128-byte reads, reflect / zero halo folded
//     into the index), split every value into bf16 hi + lo and store it straight into the MN-major
//     canonical shared-memory layout the MMA descriptor expects.
//   * W is pre-packed once per weight update (vbx_tc_pack_fwd) into K-major bf16 hi / lo tiles that
//     are exactly the shared-memory image, so one thread streams them in with a linear bulk-async
//     copy (TMA engine) that completes on an mbarrier.
//   * One thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), three per k-step:
//     hi*hi + hi*lo + lo*hi  ("bf16x3": 16 mantissa bits per operand, fp32 accumulate; measured
//     1e-5-class agreement with the fp32 reference, inside the 1e-4 contract).
//   * Epilogue: TMEM lane == time position, so after tcgen05.ld each warp writes 32 consecutive t of
//     one output channel per store: coalesced into the reference's (B,C,T) layout, with bias /
//     LeakyReLU / residual / mask fused exactly as in the SIMT path.
// Two CTAs per SM (<= 256 TMEM columns and <= ~110 KB smem each) overlap one tile's epilogue with
// the other's main loop.
#include "common.cuh"
#include "conv_plan.h"
#include "tc_common.cuh"
#include <limits.h>
#include <stdlib.h>

namespace vbx {
namespace tc {

template <int V> struct IntC { static constexpr int value = V; };

static const int kRows = 128;               // MMA M
static const int kKC = 32;                  // reduction elements per stage (2 MMA k-steps of 16)
static const int kSboA = 144;               // bytes between consecutive 8-row units of A (padded: conflict-free stores)
static const int kLboA = 16 * kSboA;        // bytes between 8-k groups of A
static const int kPlaneA = (kKC / 8) * kLboA;   // 9216
static const int kProducers = 256;
static const int kThreads = 320;            // 8 producer/epilogue warps + MMA warp + weight-copy warp

struct TcP {
  GemmP g;
  const unsigned char* packed;
  int NT, ntiles_n, nchunks, tmem_cols, stages;
  int nphase;       // DGRAD: packed tiles are indexed [g][nt][phase][chunk], nchunks = chunks per phase
  int cpad;         // reduction channels padded to a multiple of 8 (or to 1/2/4): kk = tap*cpad + channel
  int nsplit;       // bf16 components per operand: 2 = hi+lo, 3 MMAs (16 mantissa bits); 3 = hi+mid+lo, 6 MMAs (24 bits)
  // Merged-phase ("sub-pixel") input gradient of a strided conv (dil = 1, zero halo): all `stride` phases of
  // dx share the same dy window, so they become COLUMNS of one stride-1 forward-style GEMM
  //   D[(b,v), (ph,ci)] = sum_{co,m'} dy[b,co, v + m' - padp] * Wm[(ph,ci), (co,m')],   dx[b,ci, r(ph) + s*v] = D
  // -> dy is gathered once instead of once per phase and the MMA is `stride` times wider.  The kernel then runs
  // in FWD mode on a re-labelled geometry (g) and only the output indexing differs.
  int merged;
  int mg_s, mg_Tx, mg_Cin, mg_Cing;           // stride, dx length, dx channels, dx channels per group
  int mg_r[8];                                // first position u of each phase
  // "slab" form (tc_slab.cuh): no im2col replication; chosen by fill_tc from the geometry alone
  int slab;
  int sl_R;         // virtual output rows per batch item (Tout real ones + the tail the taps overhang)
  int sl_ncg;       // 16-channel groups of the reduction
  int sl_tpb;       // taps per weight stage
  int sl_nbst;      // weight stages per channel group = ceil(K / sl_tpb)
  int sl_SA, sl_SB; // ring depths: slab stages / weight stages
  int sl_rt;        // 128-row tiles per CTA of the streaming slab kernel (1, or 2: two accumulators share every weight tile)
  // Densified groups (slab form only): a grouped conv whose groups are narrower than the 16-channel reduction unit
  // runs as ONE dense conv over all channels against block-diagonal packed weights.  The MMAs are far from the
  // bound on such layers; what counts is that every input position is staged once per tile instead of once per
  // group and that a tile writes all output channels.  dg = original group count (0: not densified).
  int dg, dg_cin, dg_cout;
  // Persistent slab form (tc_pslab.cuh): weights resident in shared memory, CTAs walk row tiles.
  int ps;           // 1: launch tc_pslab_kernel (then sl_tpb = K, sl_nbst = 1, tmem_cols = two accumulator buffers)
  int ps_slots;     // slab ring depth in whole tiles (1 or 2)
  int ps_gx;        // CTAs along the row-tile axis
};

// one reduction segment of a tile: FWD has one; DGRAD has one per mirror image it touches
struct Seg {
  int nchunks, phase, img;
  int k0, kstep, ntaps, tstep, kd0_hi;
};

__host__ __device__ inline int plane_b(int NT) { return NT * 64; }                 // NT rows x 32 bf16
__host__ __device__ inline int stage_bytes(int NT, int nsplit = 2) { return nsplit * (kPlaneA + plane_b(NT)); }

inline int pick_nt(int Cn) {
  int r = (Cn + 15) / 16 * 16;
  if (r <= 256) return r;
  int tiles = (r + 255) / 256;                      // split evenly into tiles of <= 256 columns
  return ((r + tiles - 1) / tiles + 15) / 16 * 16;
}
inline int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }
// Pipeline depth.  A weight tile takes ~1.5k cycles to arrive through the bulk-copy engine and the gathers
// see L2 latency too, while one stage is worth 768 tensor-core cycles at N = 256: long reductions get the
// whole SM (one CTA, up to 4-6 stages); short ones keep two CTAs per SM so that one tile's epilogue
// overlaps the other's main loop.
static const int kSmemMax = 225 * 1024;
inline int pick_stages_for(int stage_sz, int nchunks) {
  // shared-memory budget per CTA: three CTAs per SM (more producer warps in flight) unless a single stage
  // is already > 36 KB (N = 256 tiles), where two CTAs per SM is the most that fits
  static const int small_kb = getenv("VBX_TC_BUDGET_KB") ? atoi(getenv("VBX_TC_BUDGET_KB")) : 72;
  static const int tiny_kb = getenv("VBX_TC_TINY_KB") ? atoi(getenv("VBX_TC_TINY_KB")) : 54;
  const int budget = (2 * stage_sz <= tiny_kb * 1024 ? tiny_kb : 2 * stage_sz > small_kb * 1024 ? 108 : small_kb) * 1024;
  int s = budget / stage_sz;
  const int cap = 4;
  s = s > cap ? cap : s;
  if (s > nchunks) s = nchunks;
  return s < 2 ? 2 : s;
}

__device__ __forceinline__ float finish(const GemmP& P, float v, int ch, long long idx) {
  if (P.bias) v += P.bias[ch];
  if (P.mask) P.mask[idx] = v > 0.f ? 1 : 0;
  if (P.slope != 1.f) v = v > 0.f ? v : v * P.slope;
  if (P.res) v += P.res[idx];
  if (P.gate) {
    const bool fm = P.fm_other != nullptr;
    v = gate_apply(v, P.gate[idx], fm, fm ? P.fm_other[idx] : 0.f, fm ? P.fm_coef[0] : 0.f, fm ? P.fm_coef[1] : 0.f,
                   P.gate_slope);
  }
  if (P.beta != 0.f) v += P.beta * P.Y[idx];
  return v;
}

// gate stage of the full-block epilogue paths: 16 channel rows of one time column, loads issued before use
__device__ __forceinline__ void gate_block16(const GemmP& G, long long o, long long Tlen, float (&acc)[16]) {
  const bool fm = G.fm_other != nullptr;
  const float c1 = fm ? __ldg(G.fm_coef) : 0.f, c2 = fm ? __ldg(G.fm_coef + 1) : 0.f;
#pragma unroll
  for (int h = 0; h < 16; h += 8) {                     // (two halves: 16 more live registers, not 32)
    const float* gp = G.gate + o + (long long)h * Tlen;
    float y[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = gp[(long long)j * Tlen];
    if (fm) {
      const float* op = G.fm_other + o + (long long)h * Tlen;
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = op[(long long)j * Tlen];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[h + j] = gate_apply(acc[h + j], y[j], fm, fm ? b[j] : 0.f, c1, c2, G.gate_slope);
  }
}

// MODE FWD  : rows (b,t),  cols co, reduction (ci,k):  A = x[b,ci,map(t*s + k*d - pad)]
// MODE DGRAD: rows (b,u) of one phase, cols ci, reduction (co,j): A = dy[b,co,(u+pad-k*d)/s]
// MINB = CTAs per SM the register budget is sized for (3: 64 registers, 4: 48 registers for N <= 64 tiles,
// whose short reductions are latency bound and want as many tiles in flight as possible)
template <int MODE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) tc_conv_kernel(const TcP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmP& G = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = P.stages, NT = P.NT, NS = P.nsplit;
  const int stage_sz = stage_bytes(NT, NS);
  unsigned char* stage0 = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_sz);
  uint64_t* full_a = bars;
  uint64_t* full_b = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* acc_full = bars + 3 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);
  Seg* segs = reinterpret_cast<Seg*>(tmem_slot + 2);
  int* nseg_p = reinterpret_cast<int*>(segs + 3);

  const int grp = blockIdx.y / P.ntiles_n, nt = blockIdx.y % P.ntiles_n;
  const int row_base = blockIdx.x * kRows;
  const int Cred = MODE == FWD ? G.Cin_g : G.Cout_g;     // channels walked by the reduction
  const int Ccol = MODE == FWD ? G.Cout_g : G.Cin_g;     // channels along the MMA N dimension
  DgradPhase dp{0, 0};
  int N;
  if (MODE == FWD) {
    N = G.B * G.Tout;
  } else {
    dp = dgrad_phase(G, blockIdx.z);
    N = G.B * dp.Up;
  }
  if (row_base >= N) return;                             // uniform (DGRAD phases differ in length)

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_a[s], kProducers);
      mbar_init(&full_b[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    int ns = 0;
    if (MODE == FWD) {
      segs[0] = Seg{P.nchunks, 0, 0, 0, 1, G.K, 0, 0};
      ns = 1;
    } else {
      for (int img = 0; img < 3; ++img) {
        if (!dgrad_tile_needs_img(G, blockIdx.z, row_base, kRows, img)) continue;
        DgradImg im = dgrad_img(G, dp.r, img);
        const int nch = (P.cpad * im.ntaps + kKC - 1) / kKC;
        if (nch == 0) continue;
        segs[ns++] = Seg{nch, im.phase, img, im.k0, im.kstep, im.ntaps, im.tstep, im.kd0_hi};
      }
    }
    *nseg_p = ns;
  }
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = *nseg_p;

  if (warp < 8) {
    // ===================== A producers: im2col gather -> bf16 hi/lo -> MN-major smem =====================
    // Reduction order is tap-major / channel-minor: kk = tap*cpad + channel (cpad = channel count padded
    // to 8, or to 1/2/4 for tiny layers).  A thread owns a PAIR of adjacent rows (time positions) and the
    // 8 consecutive kk of one k8-group per chunk; when cpad % 8 == 0 those are 8 channels of ONE tap, so the
    // halo / reflect / validity logic runs once per 16 loaded values and the loads walk a constant stride.
    // Two rows are split with one packed convert (cvt.rn.bf16x2) and stored with one 4-byte store per plane.
    const int pair = tid & 63, kq = tid >> 6;
    const int r0 = 2 * pair;
    const uint32_t st_off = (uint32_t)kq * kLboA + (uint32_t)(r0 >> 3) * kSboA + (uint32_t)(r0 & 7) * 2;
    const int cpad = P.cpad;
    const int chan_stride = MODE == FWD ? G.Tin : G.Tout;
    int s = 0;                                           // pipeline position: stage and parity of its use count
    uint32_t par = 0;
    for (int sg = 0; sg < nseg; ++sg) {
      const Seg seg = segs[sg];
      const int Kmod = MODE == FWD ? G.K : seg.ntaps;    // taps walked by the reduction
      const float* srcr[2];
      int baser[2];
      bool validr[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = row_base + r0 + h;
        bool valid = n < N;
        const float* src = G.X;
        int base;
        if (MODE == FWD) {
          const int b = valid ? n / G.Tout : 0, t = valid ? n % G.Tout : 0;
          src += ((long long)b * G.Cin + grp * G.Cin_g) * G.Tin;
          base = t * G.stride - G.pad;
        } else {
          const int b = valid ? n / dp.Up : 0, u = dp.r + (valid ? n % dp.Up : 0) * G.stride;
          int p = u;
          if (seg.img == 1) { p = -u; if (u < 1 || u > G.refl) valid = false; }
          else if (seg.img == 2) { p = 2 * (G.Tin - 1) - u; if (u > G.Tin - 2 || u < G.Tin - 1 - G.refl) valid = false; }
          src += ((long long)b * G.Cout + grp * G.Cout_g) * G.Tout;
          base = (p + G.pad) / G.stride - seg.kd0_hi;    // p + pad >= 0 because refl <= pad
        }
        srcr[h] = src; baser[h] = base; validr[h] = valid;
      }
      // position of tap `tap` for row h, or -1 (zero halo / outside)
      auto tap_pos = [&](int h, int tap) -> int {
        if (!validr[h] || tap >= Kmod) return -1;
        if (MODE == FWD) return map_pos(baser[h] + tap * G.dil, G.Tin, G.refl);
        const int t = baser[h] - tap * seg.tstep;
        return (t >= 0 && t < G.Tout) ? t : -1;
      };
      int tap = 0, c0 = kq * 8;                          // cpad % 8 == 0 walk: (tap, first channel) of this thread's run
      if ((cpad & 7) == 0) { tap = c0 / cpad; c0 %= cpad; }
      for (int cs = 0; cs < seg.nchunks; ++cs) {
        float xa[8], xb[8];
        if ((cpad & 7) == 0) {
          const int pa = tap_pos(0, tap), pb = tap_pos(1, tap);
          const int nvalid = Cred - c0;
          const float* qa = srcr[0] + (long long)c0 * chan_stride + (pa >= 0 ? pa : 0);
          const float* qb = srcr[1] + (long long)c0 * chan_stride + (pb >= 0 ? pb : 0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            xa[i] = (pa >= 0 && i < nvalid) ? qa[i * chan_stride] : 0.f;
            xb[i] = (pb >= 0 && i < nvalid) ? qb[i * chan_stride] : 0.f;
          }
          c0 += kKC;
          while (c0 >= cpad) { c0 -= cpad; ++tap; }
        } else {                                         // cpad in {1,2,4}: a run spans 8/cpad taps
          const int kk0 = cs * kKC + kq * 8;
          auto small = [&](auto CP_) {
            constexpr int CP = decltype(CP_)::value;
            const int tp0 = kk0 / CP;
#pragma unroll
            for (int ti = 0; ti < 8 / CP; ++ti) {
              const int pa = tap_pos(0, tp0 + ti), pb = tap_pos(1, tp0 + ti);
#pragma unroll
              for (int ci = 0; ci < CP; ++ci) {
                const bool cv = ci < Cred;
                const float va = srcr[0][(cv ? ci : 0) * chan_stride + (pa >= 0 ? pa : 0)];
                const float vb = srcr[1][(cv ? ci : 0) * chan_stride + (pb >= 0 ? pb : 0)];
                xa[ti * CP + ci] = (pa >= 0 && cv) ? va : 0.f;
                xb[ti * CP + ci] = (pb >= 0 && cv) ? vb : 0.f;
              }
            }
          };
          if (cpad == 4) small(IntC<4>{});
          else if (cpad == 2) small(IntC<2>{});
          else small(IntC<1>{});
        }
        mbar_wait(&empty[s], par ^ 1u);
        unsigned char* a_hi = stage0 + (size_t)s * stage_sz + st_off;
        unsigned char* a_lo = a_hi + kPlaneA;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 hi2 = __floats2bfloat162_rn(xa[i], xb[i]);      // .x = row r0, .y = row r0+1
          const float2 hf = __bfloat1622float2(hi2);
          const float ra = xa[i] - hf.x, rb = xb[i] - hf.y;
          const __nv_bfloat162 lo2 = __floats2bfloat162_rn(ra, rb);
          *reinterpret_cast<__nv_bfloat162*>(a_hi + i * 16) = hi2;
          *reinterpret_cast<__nv_bfloat162*>(a_lo + i * 16) = lo2;
          if (NS == 3) {
            const float2 mf = __bfloat1622float2(lo2);
            *reinterpret_cast<__nv_bfloat162*>(a_lo + kPlaneA + i * 16) = __floats2bfloat162_rn(ra - mf.x, rb - mf.y);
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_a[s]);
        if (++s == S) { s = 0; par ^= 1u; }
      }
    }
    // ===================== epilogue: TMEM -> registers -> fused output stage -> (B,C,T) =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int en = row_base + q * 32 + lane;
    const bool ev = en < N;
    long long out_base;                                  // index of (b, first channel of the group, position)
    int Tlen, Ctot;
    if (MODE == FWD) {
      const int eb = ev ? en / G.Tout : 0, et = ev ? en % G.Tout : 0;
      Tlen = G.Tout; Ctot = G.Cout;
      out_base = ((long long)eb * Ctot + grp * Ccol) * Tlen + et;
    } else {
      const int eb = ev ? en / dp.Up : 0, eu = dp.r + (ev ? en % dp.Up : 0) * G.stride;
      Tlen = G.Tin; Ctot = G.Cin;
      out_base = ((long long)eb * Ctot + grp * Ccol) * Tlen + eu;
    }
    const int nblk = NT / 16;
    const int blk_lo = half == 0 ? 0 : (nblk + 1) / 2, blk_hi = half == 0 ? (nblk + 1) / 2 : nblk;
    for (int blk = blk_lo; blk < blk_hi; ++blk) {
      float acc[16];
      if (nseg > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 16), acc);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      }
      if (MODE == FWD && P.merged) {
        // rows are (b, v); column = phase * Cin_g + ci  ->  dx[b, ci, r(phase) + s*v]
        const int eb = ev ? en / G.Tout : 0, v = ev ? en % G.Tout : 0;
        int col = nt * NT + blk * 16;
        int ph = col / P.mg_Cing, ci = col % P.mg_Cing;
#pragma unroll
        for (int j = 0; j < 16; ++j, ++col) {
          if (ev && col < Ccol) {
            const int u = P.mg_r[ph] + P.mg_s * v;
            if (u < P.mg_Tx) {
              const int ch = grp * P.mg_Cing + ci;
              const long long idx = ((long long)eb * P.mg_Cin + ch) * P.mg_Tx + u;
              G.Y[idx] = finish(G, acc[j], ch, idx);
            }
          }
          if (++ci == P.mg_Cing) { ci = 0; ++ph; }
        }
        continue;
      }
      const int cb = nt * NT + blk * 16;
      if (!ev) continue;                                 // (the TMEM load above is warp-collective)
      if (cb + 16 <= Ccol && G.beta == 0.f) {
        // full block: uniform option branches outside the column loops, 16 independent loads / stores each
        const long long o = out_base + (long long)cb * Tlen;
        if (G.bias) {
          const float* bp = G.bias + grp * Ccol + cb;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __ldg(bp + j);
        }
        if (G.mask) {
          unsigned char* mp = G.mask + o;
#pragma unroll
          for (int j = 0; j < 16; ++j) mp[(long long)j * Tlen] = acc[j] > 0.f ? 1 : 0;
        }
        if (G.slope != 1.f) {
          const float sl = G.slope;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = acc[j] > 0.f ? acc[j] : acc[j] * sl;
        }
        if (G.res) {
          const float* rp = G.res + o;
          float r[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = rp[(long long)j * Tlen];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += r[j];
        }
        if (G.gate) gate_block16(G, o, Tlen, acc);
        float* yp = G.Y + o;
#pragma unroll
        for (int j = 0; j < 16; ++j) yp[(long long)j * Tlen] = acc[j];
        continue;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {                      // (unrolled: acc[] stays in registers, no local-memory frame)
        const int col = cb + j;
        if (col < Ccol) {
          const long long idx = out_base + (long long)col * Tlen;
          G.Y[idx] = finish(G, acc[j], grp * Ccol + col, idx);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(NT, /*a_mn=*/true, /*b_mn=*/false);
      const uint32_t lbo_b = (uint32_t)NT * 16;
      uint32_t accumulate = 0;
      int s = 0;
      uint32_t par = 0;
      for (int sg = 0; sg < nseg; ++sg) {
        const int nch = segs[sg].nchunks;
        for (int cs = 0; cs < nch; ++cs) {
          mbar_wait(&full_a[s], par);
          mbar_wait(&full_b[s], par);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stage0 + (size_t)s * stage_sz);
          const uint32_t b0 = a0 + NS * kPlaneA;
#pragma unroll
          for (int ks = 0; ks < kKC / 16; ++ks) {
            uint64_t da[3], db[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              da[i] = make_desc(a0 + i * kPlaneA + ks * 2 * kLboA, kLboA, kSboA);
              db[i] = make_desc(b0 + i * plane_b(NT) + ks * 2 * lbo_b, lbo_b, 128);
            }
            // all component products a_i * b_j with i + j < NS (the dropped ones are below 2^-8NS relative)
            mma_bf16_ss(tmem_base, da[0], db[0], idesc, accumulate);
            accumulate = 1;
            mma_bf16_ss(tmem_base, da[0], db[1], idesc, 1);
            mma_bf16_ss(tmem_base, da[1], db[0], idesc, 1);
            if (NS == 3) {
              mma_bf16_ss(tmem_base, da[0], db[2], idesc, 1);
              mma_bf16_ss(tmem_base, da[2], db[0], idesc, 1);
              mma_bf16_ss(tmem_base, da[1], db[1], idesc, 1);
            }
          }
          mma_commit(&empty[s]);       // frees the stage once these MMAs have read it
          if (++s == S) { s = 0; par ^= 1u; }
        }
      }
      if (nseg > 0) mma_commit(acc_full);
      else mbar_arrive(acc_full);
    }
  } else {
    // ===================== weight tiles: linear bulk copies (TMA engine) =====================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)NS * (uint32_t)plane_b(NT);
      int s = 0;
      uint32_t par = 0;
      for (int sg = 0; sg < nseg; ++sg) {
        const Seg seg = segs[sg];
        const unsigned char* src =
            P.packed + (((size_t)(grp * P.ntiles_n + nt) * P.nphase + seg.phase) * P.nchunks) * bytes;
        for (int cs = 0; cs < seg.nchunks; ++cs) {
          mbar_wait(&empty[s], par ^ 1u);
          mbar_expect_tx(&full_b[s], bytes);
          bulk_copy_g2s(stage0 + (size_t)s * stage_sz + NS * kPlaneA, src + (size_t)cs * bytes, bytes, &full_b[s]);
          if (++s == S) { s = 0; par ^= 1u; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// Weight pre-pack: one thread per 16-byte unit (8 consecutive reduction elements of one column).
// FWD  : column = co, reduction kk = (ci,k)          -> W[co][kk]
// DGRAD: column = ci, reduction kk = (co,j) of phase -> W[co][ci][k0 + j*kstep]
template <int MODE>
__global__ void tc_pack_kernel(const float* __restrict__ w, unsigned char* __restrict__ out, const TcP P) {
  const GemmP& G = P.g;
  const int NT = P.NT;
  const int Ccol = MODE == FWD ? G.Cout_g : G.Cin_g;
  const long long units = (long long)G.groups * P.ntiles_n * P.nphase * P.nchunks * 4 * NT;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < units;
       u += (long long)gridDim.x * blockDim.x) {
    int n = (int)(u % NT);
    long long r = u / NT;
    int ku = (int)(r % 4); r /= 4;
    int c = (int)(r % P.nchunks); r /= P.nchunks;
    int ph = (int)(r % P.nphase); r /= P.nphase;
    int nt = (int)(r % P.ntiles_n);
    int g = (int)(r / P.ntiles_n);
    const int col = nt * NT + n;
    const int kk0 = c * kKC + ku * 8;
    DgradImg im = dgrad_taps(G, MODE == FWD ? 0 : ph);
    __align__(16) __nv_bfloat16 hi[8], lo[8], lo3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      const int kk = kk0 + j;
      const int tap = kk / P.cpad, cr = kk % P.cpad;     // tap-major / channel-minor reduction order
      if (col < Ccol) {
        if (MODE == FWD) {
          if (tap < G.K && cr < G.Cin_g)
            v = w[(((long long)g * G.Cout_g + col) * G.Cin_g + cr) * G.K + tap];
        } else if (tap < im.ntaps && cr < G.Cout_g) {
          v = w[(((long long)g * G.Cout_g + cr) * G.Cin_g + col) * G.K + im.k0 + tap * im.kstep];
        }
      }
      split_bf16(v, hi[j], lo[j]);
      lo3[j] = __float2bfloat16_rn((v - __bfloat162float(hi[j])) - __bfloat162float(lo[j]));
    }
    unsigned char* base =
        out + ((((size_t)(g * P.ntiles_n + nt) * P.nphase + ph) * P.nchunks + c)) * P.nsplit * plane_b(NT);
    const size_t off = ((size_t)ku * NT + n) * 16;
    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + plane_b(NT) + off) = *reinterpret_cast<const uint4*>(lo);
    if (P.nsplit == 3) *reinterpret_cast<uint4*>(base + 2 * plane_b(NT) + off) = *reinterpret_cast<const uint4*>(lo3);
  }
}

// per-phase tap bookkeeping of the merged-phase dgrad (host side, stride <= 8)
struct MergedPlan {
  int c[8], k0[8], nt[8], r[8];
  int cmax, J, tstep;
  int dilm, Kp;     // tap lattice of the re-labelled conv: Kp taps spaced dilm (J = (Kp-1)*dilm + 1 unit slots)
};
static MergedPlan merged_plan(const GemmP& G) {
  MergedPlan M;
  M.cmax = -(1 << 30);
  for (int ph = 0; ph < G.stride; ++ph) {
    DgradImg im = dgrad_taps(G, ph);
    M.r[ph] = mod_pos(ph - G.pad, G.stride);
    M.c[ph] = (M.r[ph] + G.pad) / G.stride - im.kd0_hi;
    M.k0[ph] = im.k0; M.nt[ph] = im.ntaps;
    if (im.ntaps > 0 && M.c[ph] > M.cmax) M.cmax = M.c[ph];
  }
  // taps of a phase read dy at v + c(ph) - jj*tstep (tstep = dil / gcd(dil, stride)); the common tap grid has
  // unit spacing and spans all of them: m' = (J-1) - (cmax - c(ph)) - jj*tstep
  M.tstep = dgrad_taps(G, 0).tstep;
  M.J = 1;
  for (int ph = 0; ph < G.stride; ++ph) {
    const int span = (M.nt[ph] - 1) * M.tstep + 1 + M.cmax - M.c[ph];
    if (M.nt[ph] > 0 && span > M.J) M.J = span;
  }
  // if every phase's taps fall on one lattice of spacing tstep (always so for stride 1), keep the dilation
  M.dilm = M.tstep;
  for (int ph = 0; ph < G.stride; ++ph)
    if (M.nt[ph] > 0 && (M.cmax - M.c[ph]) % M.tstep != 0) M.dilm = 1;
  M.Kp = (M.J - 1) / M.dilm + 1;
  return M;
}
struct MergedDev { int c[8], k0[8], nt[8], cmax, J, s, Cin_g, Cout_g, K, tstep, kstep, dilm; };
// index jj of the phase's tap that sits at tap `mp` of the re-labelled conv (unit slot mp*dilm), or -1
__device__ __forceinline__ int merged_tap(const MergedDev& M, int ph, int mp) {
  const int num = (M.J - 1 - mp * M.dilm) - (M.cmax - M.c[ph]);
  if (num < 0 || num % M.tstep != 0) return -1;
  const int jj = num / M.tstep;
  return jj < M.nt[ph] ? jj : -1;
}

// Wm[(ph,ci), (m', co)] = W[co][ci][k0(ph) + jj*s],  jj = (J-1-m') - (cmax - c(ph)),  zero outside the phase's taps
__global__ void tc_pack_merged_kernel(const float* __restrict__ w, unsigned char* __restrict__ out, const TcP P,
                                      const MergedDev M) {
  const int NT = P.NT;
  const int Ccol = M.s * M.Cin_g;
  const long long units = (long long)P.g.groups * P.ntiles_n * P.nchunks * 4 * NT;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < units;
       u += (long long)gridDim.x * blockDim.x) {
    int n = (int)(u % NT);
    long long r = u / NT;
    int ku = (int)(r % 4); r /= 4;
    int c = (int)(r % P.nchunks); r /= P.nchunks;
    int nt = (int)(r % P.ntiles_n);
    int g = (int)(r / P.ntiles_n);
    const int col = nt * NT + n;
    const int ph = col / M.Cin_g, ci = col % M.Cin_g;
    const int kk0 = c * kKC + ku * 8;
    __align__(16) __nv_bfloat16 hi[8], lo[8], lo3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      const int kk = kk0 + j;
      const int mp = kk / P.cpad, cr = kk % P.cpad;
      if (col < Ccol && cr < M.Cout_g && mp * M.dilm < M.J) {
        const int jj = merged_tap(M, ph, mp);
        if (jj >= 0)
          v = w[(((long long)g * M.Cout_g + cr) * M.Cin_g + ci) * M.K + M.k0[ph] + jj * M.kstep];
      }
      split_bf16(v, hi[j], lo[j]);
      lo3[j] = __float2bfloat16_rn((v - __bfloat162float(hi[j])) - __bfloat162float(lo[j]));
    }
    unsigned char* base = out + ((((size_t)(g * P.ntiles_n + nt)) * P.nchunks + c)) * P.nsplit * plane_b(NT);
    const size_t off = ((size_t)ku * NT + n) * 16;
    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + plane_b(NT) + off) = *reinterpret_cast<const uint4*>(lo);
    if (P.nsplit == 3) *reinterpret_cast<uint4*>(base + 2 * plane_b(NT) + off) = *reinterpret_cast<const uint4*>(lo3);
  }
}

#include "tc_slab.cuh"
#include "tc_pslab.cuh"

// Merged-phase input gradient on the gather kernel: worth it only where the common tap grid has no holes
// (dil = 1) and a phase does not already fill a 256-wide tile.  The slab kernel takes the general case.
static bool use_merged(const GemmP& G) {
  static const bool off = getenv("VBX_TC_MERGED") && atoi(getenv("VBX_TC_MERGED")) == 0;
  return !off && G.stride > 1 && G.stride <= 8 && G.dil == 1 && G.refl == 0 && G.Cin_g < 256;
}

// Slab form: forward-style problems (incl. merged-phase input gradients) with >= 2 taps whose reduction has enough
// channels for 16-wide groups, when everything a producer thread indexes fits 32 bits and the rings fit shared memory.
static size_t slab_smem_bytes(const TcP& P) {
  return (size_t)P.sl_SA * slab_a_stage(P.g, P.sl_rt) + (size_t)P.sl_SB * slab_b_stage(P.NT, P.sl_tpb) +
         (2 * P.sl_SA + 2 * P.sl_SB + 1) * sizeof(uint64_t) + 16 + (size_t)P.g.K * sizeof(int) + 16;
}
static size_t pslab_smem_bytes(const TcP& P, int slots) {
  const size_t a_stage = (size_t)slab_a_stage(P.g);
  return (size_t)P.sl_ncg * slab_b_stage(P.NT, P.g.K) + (size_t)slots * P.sl_ncg * a_stage +
         (2 * slots + 5) * sizeof(uint64_t) + 16 + (size_t)P.g.K * sizeof(int) + 16;
}
// Persistent form when the packed weights of one (group, column tile) fit beside at least one tile's slab:
// picks CTAs per SM (shared memory, TMEM columns, 64-register budget) and the slab ring depth.
static void plan_pslab(TcP& P) {
  static const int max_kb = getenv("VBX_TC_PS_MAX_KB") ? atoi(getenv("VBX_TC_PS_MAX_KB")) : 112;   // 0: off
  P.ps = 0;
  const size_t wbytes = (size_t)P.sl_ncg * slab_b_stage(P.NT, P.g.K);
  if (wbytes > (size_t)max_kb * 1024) return;
  const int cols2 = 2 * pow2_cols(P.NT);
  if (cols2 > 512) return;
  const int occ_tmem = 512 / cols2;
  // (CTAs per SM, ring slots).  One CTA per SM is not offered: with nothing to overlap its stage / MMA / drain chain
  // it measured slower than the streaming kernel (192>384 k7 s2 input gradient: 77 vs 60 us).
  static const int order[4][2] = {{3, 2}, {2, 2}, {3, 1}, {2, 1}};
  for (int i = 0; i < 4; ++i) {
    const int occ = order[i][0], slots = order[i][1];
    if (occ > occ_tmem) continue;
    if (pslab_smem_bytes(P, slots) + 1024 > (size_t)(227 * 1024) / occ) continue;
    const long long row_tiles = ((long long)P.g.B * P.sl_R + kRows - 1) / kRows;
    const int gy = P.ntiles_n * P.g.groups;
    long long gx = (148ll * occ) / gy;
    if (gx < 1) gx = 1;
    if (gx > row_tiles) gx = row_tiles;
    P.ps = 1; P.ps_slots = slots; P.ps_gx = (int)gx;
    P.sl_tpb = P.g.K; P.sl_nbst = 1;
    P.tmem_cols = cols2;
    return;
  }
}

static void plan_slab(TcP& P, const vbx_conv_desc* d) {
  static const int min_cin = getenv("VBX_TC_SLAB_MIN_CIN") ? atoi(getenv("VBX_TC_SLAB_MIN_CIN")) : 8;
  static const int min_k = getenv("VBX_TC_SLAB_MIN_K") ? atoi(getenv("VBX_TC_SLAB_MIN_K")) : 2;
  static const bool off = getenv("VBX_TC_SLAB") && atoi(getenv("VBX_TC_SLAB")) == 0;
  const GemmP& G = P.g;
  if (off || P.nsplit != 2 || G.K < min_k || G.K > 128 || G.Cin_g < min_cin || G.stride > 8) return;
  if (slab_npos(G) > kSlabMaxU * kProducers) return;
  const int R = G.Tout - 1 + ((G.K - 1) * G.dil + 1 + G.stride - 1) / G.stride;
  if ((long long)d->B * R * G.stride + 2048 >= (1ll << 31) || (long long)d->B * G.Cin * G.Tin >= (1ll << 31)) return;
  P.sl_R = R;
  P.sl_rt = 1;
  P.sl_ncg = (G.Cin_g + 15) / 16;
  int tpb = 16384 / (P.NT * 64);
  if (tpb < 1) tpb = 1;
  if (tpb > G.K) tpb = G.K;
  P.sl_tpb = tpb;
  P.sl_nbst = (G.K + tpb - 1) / tpb;
  P.sl_SA = P.sl_ncg >= 2 ? 2 : 1;
  const int total_b = P.sl_ncg * P.sl_nbst;
  P.sl_SB = total_b < 4 ? total_b : 4;
  if (slab_smem_bytes(P) > 200 * 1024) {
    P.sl_SB = total_b < 2 ? total_b : 2;
    if (slab_smem_bytes(P) > 200 * 1024) return;
  }
  {
    // A third CTA per SM beats a deeper weight ring where a tile is a latency chain (stage -> K small MMAs -> drain) and
    // the weight tiles are small: ncu on MelGAN stage 1 (N = 64) showed 2 CTAs per SM, 30 % of the warp slots, long-
    // scoreboard stalls, tensor pipe 24 %.  Measured with a 2-deep ring and 3 CTAs: MelGAN stages 1-2 forward 205 -> 155 us,
    // input gradients 244 -> 193 us, PQMF-discriminator stages and C = 128 generator convs -15..-20 %; NOT for N = 256
    // tiles (MelGAN stages 3-4 input gradients 342 -> 396 us: those are bound by the weight stream).
    static const int sb3 = getenv("VBX_TC_SLAB_SB3") ? atoi(getenv("VBX_TC_SLAB_SB3")) : 1;
    if (sb3 > 0 && P.NT <= 128 && P.sl_SB > 2 && slab_smem_bytes(P) + 1024 > (227 * 1024) / 3) {
      TcP Q = P;
      Q.sl_SB = 2;
      if (slab_smem_bytes(Q) + 1024 <= (227 * 1024) / 3) P.sl_SB = 2;
    }
  }
  P.slab = 1;
  plan_pslab(P);
  if (!P.ps && !P.merged) {
    // Two row tiles per CTA where the kernel is bound by streaming weight tiles from L2 (see tc_slab_kernel): forward
    // layers with a long reduction whose single-tile form already owns the SM (> half the shared memory), so nothing
    // is lost in CTAs per SM.  Measured (MelGAN stage 4 forward, alone): 0.52 -> 0.34 ms.  NOT for the merged-phase
    // input gradients (scattered epilogue stores: they need a second CTA per SM to hide the drain; 347 -> 400 us) and
    // not for short-K layers that fit two CTAs per SM (1024 -> 1024 k5: 170 -> 460 us when halved to 96 CTAs).
    static const int rt2_min = getenv("VBX_TC_SLAB_RT2_MIN") ? atoi(getenv("VBX_TC_SLAB_RT2_MIN")) : 96;   // 0: off
    const long long row_tiles = ((long long)d->B * R + kRows - 1) / kRows;
    if (rt2_min > 0 && P.NT >= 128 && 2 * pow2_cols(P.NT) <= 512 && P.sl_ncg * G.K >= rt2_min && row_tiles >= 2 &&
        slab_smem_bytes(P) > 113 * 1024 && slab_npos(G, 2) <= kSlabMaxU * kProducers) {
      TcP Q = P;
      Q.sl_rt = 2;
      if (slab_smem_bytes(Q) > 200 * 1024) { Q.sl_SB = total_b < 2 ? total_b : 2; }
      if (slab_smem_bytes(Q) <= 200 * 1024) { P.sl_rt = 2; P.sl_SB = Q.sl_SB; }
    }
    // Narrow column tiles with many taps (MelGAN stages 1-2: N = 64, k 41): every 128-row tile pulls K x 4 KB of weight
    // tiles through the bulk-copy ring for 123 small MMAs; two row tiles per CTA halve that stream and the per-tile
    // fixed costs.
    static const int n64 = getenv("VBX_TC_SLAB_RT2_N64") ? atoi(getenv("VBX_TC_SLAB_RT2_N64")) : 0;
    if (n64 > 0 && P.sl_rt == 1 && P.NT <= 64 && G.K >= 32 && row_tiles >= 4 * 148 &&
        slab_npos(G, 2) <= kSlabMaxU * kProducers) {
      TcP Q = P;
      Q.sl_rt = 2;
      if (slab_smem_bytes(Q) <= 200 * 1024) P.sl_rt = 2;
    }
  }
}

static void densify(GemmP& g) { g.groups = 1; g.Cin_g = g.Cin; g.Cout_g = g.Cout; }

// all input channels of the layer fit a few 16-channel reduction units: see TcP::dg
static bool dense_candidate(const vbx_conv_desc* d, int nsplit) {
  static const int max_cin = getenv("VBX_TC_DENSE_MAX_CIN") ? atoi(getenv("VBX_TC_DENSE_MAX_CIN")) : 48;
  return nsplit == 2 && d->groups > 1 && d->Cin / d->groups > 1 && d->Cin <= max_cin && d->Cout <= 256;
}

static int fill_tc_geom(TcP& P, const vbx_conv_desc* d, int mode, int nsplit, bool dense);

static int fill_tc(TcP& P, const vbx_conv_desc* d, int mode, int nsplit) {
  int code = 0;
  const char* msg = check_desc_msg(d, &code);
  if (msg) return fail(code, msg);
  if (nsplit != 2 && nsplit != 3) return fail(VBX_UNSUPPORTED, "tc: nsplit must be 2 (bf16x3) or 3 (bf16x6)");
  if (dense_candidate(d, nsplit) && fill_tc_geom(P, d, mode, nsplit, true) == 0 && P.slab) return 0;
  return fill_tc_geom(P, d, mode, nsplit, false);
}

static int fill_tc_geom(TcP& P, const vbx_conv_desc* d, int mode, int nsplit, bool dense) {
  fill(P.g, d);
  P.dg = 0; P.dg_cin = P.g.Cin_g; P.dg_cout = P.g.Cout_g;
  if (dense) { P.dg = P.g.groups; densify(P.g); }
  P.nsplit = nsplit;
  P.merged = 0;
  P.slab = 0;
  P.sl_rt = 1;
  P.ps = 0;
  if (mode == DGRAD && P.g.refl == 0 && P.g.stride <= 8) {
    // Zero-halo input gradient as a stride-1 forward conv over dy (Cout ch, Tout long) -> D (s*Cin ch, V long):
    // every stride phase of dx becomes a block of columns, the taps sit on a unit-spaced grid of J slots
    // (for stride 1 this is just the flipped-tap correlation).  Kept if the slab kernel takes it, or - on the
    // gather kernel - where use_merged() says it pays.
    const GemmP o = P.g;
    const MergedPlan M = merged_plan(o);
    P.merged = 1;
    P.mg_s = o.stride; P.mg_Tx = o.Tin; P.mg_Cin = o.Cin; P.mg_Cing = o.Cin_g;
    for (int ph = 0; ph < 8; ++ph) P.mg_r[ph] = ph < o.stride ? M.r[ph] : 0;
    P.g.Cin = o.Cout; P.g.Cin_g = o.Cout_g;
    P.g.Cout = o.stride * o.Cin; P.g.Cout_g = o.stride * o.Cin_g;
    P.g.Tin = o.Tout; P.g.Tout = (o.Tin + o.stride - 1) / o.stride;
    P.g.K = M.Kp; P.g.stride = 1; P.g.dil = M.dilm; P.g.pad = M.J - 1 - M.cmax; P.g.refl = 0;
    P.NT = pick_nt(P.g.Cout_g);
    P.ntiles_n = (P.g.Cout_g + P.NT - 1) / P.NT;
    // (a unit lattice under dilated taps multiplies the MMA work by Kp*stride / K: only worth it while the layer
    //  is far from MMA-bound, i.e. for modest reductions)
    const bool stuffed = (long long)M.Kp * o.stride * 2 > 3ll * o.K && o.Cout_g >= 128;
    if (M.J >= 1 && P.g.pad >= 0 && !stuffed) plan_slab(P, d);
    if (P.slab || use_merged(o)) {
      mode = FWD;
    } else {
      P.g = o;
      P.merged = 0;
    }
  }
  const int Ccol = mode == FWD ? P.g.Cout_g : P.g.Cin_g;
  P.NT = pick_nt(Ccol);
  P.ntiles_n = (Ccol + P.NT - 1) / P.NT;
  const int Cred = mode == FWD ? P.g.Cin_g : P.g.Cout_g;
  P.cpad = Cred >= 5 ? (Cred + 7) / 8 * 8 : (Cred >= 3 ? 4 : Cred);
  if (mode == FWD) {
    P.nphase = 1;
    P.nchunks = (P.cpad * P.g.K + kKC - 1) / kKC;
  } else {
    P.nphase = P.g.stride;
    int g = gcd_i(P.g.dil, P.g.stride);
    int kstep = P.g.stride / g;
    int max_taps = (P.g.K + kstep - 1) / kstep;
    P.nchunks = (P.cpad * max_taps + kKC - 1) / kKC;
  }
  P.tmem_cols = pow2_cols(P.NT);
  P.stages = pick_stages_for(stage_bytes(P.NT, P.nsplit), P.nchunks);
  if (mode == FWD && !P.slab) plan_slab(P, d);
  if (P.slab && P.ps) plan_pslab(P);          // (re-applies tmem_cols = two accumulator buffers)
  return 0;
}

static size_t smem_bytes(const TcP& P) {
  return (size_t)P.stages * stage_bytes(P.NT, P.nsplit) + (3 * P.stages + 1) * sizeof(uint64_t) + 16 + 3 * sizeof(Seg) + 16;
}

template <int MODE>
static int launch_tc(const TcP& P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_conv_kernel<MODE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(tc_conv_kernel<MODE, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce != cudaSuccess) return fail((int)ce, "tc_conv: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  long long rows = MODE == FWD ? (long long)P.g.B * P.g.Tout
                               : (long long)P.g.B * ((P.g.Tin + P.g.stride - 1) / P.g.stride);
  dim3 grid((unsigned)((rows + kRows - 1) / kRows), (unsigned)(P.ntiles_n * P.g.groups),
            (unsigned)(MODE == FWD ? 1 : P.g.stride));
  if (grid.y > 65535 || grid.z > 65535) return fail(VBX_UNSUPPORTED, "tc_conv: grid too large");
  static const int max_chunks4 = getenv("VBX_TC_MINB4_CHUNKS") ? atoi(getenv("VBX_TC_MINB4_CHUNKS")) : 0;   // 0: the 48-register variant is off (it spills; measured slower on every layer)
  if (smem_bytes(P) <= 56 * 1024 && P.tmem_cols <= 128 && P.nchunks <= max_chunks4)
    tc_conv_kernel<MODE, 4><<<grid, kThreads, smem_bytes(P), st>>>(P);
  else
    tc_conv_kernel<MODE, 3><<<grid, kThreads, smem_bytes(P), st>>>(P);
  return launched("tc_conv_kernel");
}

static int launch_slab(const TcP& P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce != cudaSuccess) return fail((int)ce, "tc_slab: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  const long long rows = (long long)P.g.B * P.sl_R;
  const int per = kRows * P.sl_rt;
  dim3 grid((unsigned)((rows + per - 1) / per), (unsigned)(P.ntiles_n * P.g.groups), 1);
  if (grid.y > 65535) return fail(VBX_UNSUPPORTED, "tc_slab: grid too large");
  TcP Q = P;
  Q.tmem_cols = P.sl_rt * pow2_cols(P.NT);
  tc_slab_kernel<<<grid, kThreads, slab_smem_bytes(P), st>>>(Q);
  return launched("tc_slab_kernel");
}

static int launch_pslab(const TcP& P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_pslab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return fail((int)ce, "tc_pslab: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  dim3 grid((unsigned)P.ps_gx, (unsigned)(P.ntiles_n * P.g.groups), 1);
  if (grid.y > 65535) return fail(VBX_UNSUPPORTED, "tc_pslab: grid too large");
  tc_pslab_kernel<<<grid, kThreads, pslab_smem_bytes(P, P.ps_slots), st>>>(P);
  return launched("tc_pslab_kernel");
}

template <int MODE>
static int pack_tc(const TcP& P, const float* w, void* packed, cudaStream_t st) {
  long long units = (long long)P.g.groups * P.ntiles_n * P.nphase * P.nchunks * 4 * P.NT;
  int blocks = (int)((units + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  tc_pack_kernel<MODE><<<blocks, 256, 0, st>>>(w, (unsigned char*)packed, P);
  return launched("tc_pack_kernel");
}

}  // namespace tc
}  // namespace vbx

namespace vbx {
bool skinny_wgrad_ok(const GemmP& P);
int skinny_wgrad(const GemmP& P, cudaStream_t st);
}
using namespace vbx;
using namespace vbx::tc;

// ---------------------------------------------------------------------------------------------
// WGRAD on tensor cores: dW[co, (ci,k)] += sum_{(b,t)} dy[b,co,t] * x[b,ci,map(t*s + k*d - pad)].
// Rows of the MMA are output channels (M = 128 lanes), columns are (ci,k) (N <= 256), and the
// reduction walks TIME, which is the contiguous axis of both operands: both tiles are K-major and
// both are gathered by the producer warps (lanes walk t: coalesced), nothing is pre-packed.
// The (b,t) range is split over blockIdx.z; partial tiles are added with fp32 reductions.
static const int kLboW = kRows * 16 + 16;     // A: bytes between 8-t units (+16: conflict-free 2-byte stores)

// Reduction split of a weight-gradient grid: `tiles` independent output tiles, `slots` CTAs resident on the GPU at
// once.  Pick the split whose total CTA count fills whole waves best (an extra 0.3 wave costs a full one).
static long long pick_split(long long tiles, long long slots, long long max_split) {
  long long best = 1;
  double best_eff = 0.0;
  for (int w = 1; w <= 6; ++w) {
    long long sp = w * slots / tiles;
    if (sp < 1) sp = 1;
    if (sp > max_split) sp = max_split;
    const double waves = (double)(tiles * sp) / (double)slots;
    const double eff = waves / (double)(long long)(waves + 0.999999);
    if (eff > best_eff + 0.02) { best_eff = eff; best = sp; }
  }
  return best;
}

struct TcW {
  GemmP g;
  int NT, ntiles_n, mtiles, tmem_cols, stages;
  int red_per;      // reduction elements per blockIdx.z (multiple of 32)
};
__host__ __device__ inline int lbo_wb(int NT) { return NT * 16 + 16; }
__host__ __device__ inline int wstage_bytes(int NT) { return 2 * 4 * kLboW + 2 * 4 * lbo_wb(NT); }

// Row loops of the wgrad producers: U rows' loads are issued back to back, then the U converts + stores; the tail
// goes through 4 / 2 / 1-row batches so short tiles also keep several loads in flight.
template <int U, class L, class St>
__device__ __forceinline__ void row_batch(int m, L& load, St& store) {
  float v[U][4];
#pragma unroll
  for (int u = 0; u < U; ++u) load(m + 16 * u, v[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) store(m + 16 * u, v[u]);
}
template <int UNR, class L, class St>
__device__ __forceinline__ void row_loop(int m, int end, L load, St store) {
  for (; m + 16 * (UNR - 1) < end; m += 16 * UNR) row_batch<UNR>(m, load, store);
  if (UNR > 4 && m + 48 < end) { row_batch<4>(m, load, store); m += 64; }
  if (UNR > 2 && m + 16 < end) { row_batch<2>(m, load, store); m += 32; }
  if (m < end) row_batch<1>(m, load, store);
}

// MINB / UNR: CTAs per SM the register budget allows and the row-loop unroll (loads in flight per lane).
// N = 256 tiles own 256 TMEM columns, so only two CTAs fit per SM anyway: they get 96 registers and unroll 8.
template <int MINB, int UNR>
__global__ void __launch_bounds__(kThreads, MINB) tc_wgrad_kernel(const TcW P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmP& G = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = P.stages, NT = P.NT;
  const int stage_sz = wstage_bytes(NT);
  const int plane_a = 4 * kLboW, plane_bw = 4 * lbo_wb(NT);
  unsigned char* stage0 = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_sz);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  int2* rowinfo = reinterpret_cast<int2*>(tmem_slot + 2);          // per column n: (ci*Tin, k*d - pad)
  int* roff = reinterpret_cast<int*>(rowinfo + NT);                // per column n: ci*Tin + k*d - pad
  int* kspan = roff + NT;                                          // [min, max] of k*d - pad over the tile

  const int per_g = P.mtiles * P.ntiles_n;
  const int grp = blockIdx.y / per_g, mt = (blockIdx.y % per_g) / P.ntiles_n, nt = blockIdx.y % P.ntiles_n;
  const int Ncols = G.Cin_g * G.K;
  const int total = G.B * G.Tout;
  const int red_lo = blockIdx.z * P.red_per;
  const int red_hi = min(red_lo + P.red_per, total);
  if (red_lo >= red_hi) return;
  const int nchunks = (red_hi - red_lo + kKC - 1) / kKC;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], kProducers / 2); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < NT; i += kThreads) {
    const int n = nt * NT + i;
    int2 ri = make_int2(0, INT_MIN);
    if (n < Ncols) { const int ci = n / G.K, k = n % G.K; ri = make_int2(ci * G.Tin, k * G.dil - G.pad); }
    rowinfo[i] = ri;
    roff[i] = n < Ncols ? ri.x + ri.y : INT_MIN;
  }
  if (tid == 0) {
    // columns of a tile cover consecutive (ci,k): if they span a whole channel all taps occur
    const int n_lo = nt * NT, n_hi = min(n_lo + NT, Ncols) - 1;
    int klo = n_lo % G.K, khi = n_hi % G.K;
    if (n_hi / G.K != n_lo / G.K) { klo = 0; khi = G.K - 1; }
    kspan[0] = klo * G.dil - G.pad;
    kspan[1] = khi * G.dil - G.pad;
  }
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===================== producers: dy rows and im2col'd x rows, lanes walk t =====================
    // A lane owns FOUR consecutive reduction positions of a row (8 lanes = the 32 positions of a stage), so the
    // address / halo / batch-boundary logic is paid once per four values, the split uses packed converts and a
    // row costs one 8-byte store per plane.  A warp covers 4 rows per pass (rows 4 apart: disjoint banks); the
    // 8 warps form two groups that fill alternate stages, so two stages' loads are in flight per CTA.
    const int tq = lane & 7, rsel = lane >> 3;
    const int grp2 = warp >> 2;
    const int rl = (warp & 3) + 4 * rsel;                            // row within a 16-row pass
    const uint32_t lane_off = (uint32_t)(tq >> 1) * 0 + (uint32_t)(tq & 1) * 8;
    const int ku = tq >> 1;
    const int co_base = mt * kRows;
    const int rows_a = min(kRows, G.Cout_g - co_base);
    const uint32_t lbo_b = (uint32_t)lbo_wb(NT);
    const int kmin = kspan[0], kmax = kspan[1];          // range of k*d - pad over this tile's columns
    // Rows beyond the layer's channel / column count stay zero for the whole kernel: clear the stages once
    // (generic-proxy writes, published by the fence that precedes every arrive) and never touch them again.
    const int ncols_tile = min(NT, Ncols - nt * NT);
    const int rows_a16 = (rows_a + 15) & ~15, rows_b16 = (ncols_tile + 15) & ~15;
    if (rows_a16 < kRows || rows_b16 < NT) {
      uint4* z = reinterpret_cast<uint4*>(stage0);
      const int n16 = S * stage_sz / 16;
      for (int i = tid; i < n16; i += kProducers) z[i] = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("bar.sync 1, 256;" ::: "memory");     // producer warps only
    }
    auto put4 = [](unsigned char* hi_p, unsigned char* lo_p, const float (&v)[4]) {
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
      const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - f01.x, v[1] - f01.y);
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(v[2] - f23.x, v[3] - f23.y);
      uint2 h, l;
      h.x = *reinterpret_cast<const uint32_t*>(&h01); h.y = *reinterpret_cast<const uint32_t*>(&h23);
      l.x = *reinterpret_cast<const uint32_t*>(&l01); l.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(hi_p) = h;
      *reinterpret_cast<uint2*>(lo_p) = l;
    };
    for (int c = grp2; c < nchunks; c += 2) {
      const int s = c % S, use = c / S;
      const int r0 = red_lo + c * kKC + 4 * tq;          // this lane's four reduction elements r0 .. r0+3
      const int nval = max(0, min(4, red_hi - r0));
      const int b0 = nval > 0 ? r0 / G.Tout : 0, t0 = nval > 0 ? r0 % G.Tout : 0;
      const bool fast = nval == 4 && t0 + 3 < G.Tout;    // all four valid and inside one batch item
      const float* dyp = G.DY + ((long long)b0 * G.Cout + grp * G.Cout_g + co_base) * G.Tout + t0;
      const float* xp = G.X + ((long long)b0 * G.Cin + grp * G.Cin_g) * G.Tin;
      const int ts0 = t0 * G.stride;
      const bool interior = fast && ts0 + kmin >= 0 && ts0 + 3 * G.stride + kmax < G.Tin;
      // generic per-element decode (lanes that straddle a batch boundary or the end of the slice)
      int eb[4], et[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int t = t0 + e, b = b0;
        if (t >= G.Tout) { t -= G.Tout; ++b; }           // 4 consecutive positions cross at most one boundary
        eb[e] = b; et[e] = e < nval ? t : -1;
      }
      // The row loops below contain no control flow (out-of-range rows / elements load a safe address and are
      // zeroed by selects), so the unrolled iterations' loads are all issued before the first convert waits.
      // A warp with any boundary lane takes the masked form for all its lanes (cost = max, not sum, of the two).
      const bool wfast = __all_sync(0xffffffffu, fast);
      const bool wint = __all_sync(0xffffffffu, interior);
      mbar_wait(&empty[s], (use & 1) ^ 1);
      unsigned char* a_hi = stage0 + (size_t)s * stage_sz + (uint32_t)ku * kLboW + lane_off;
      unsigned char* a_lo = a_hi + plane_a;
      unsigned char* b_hi = stage0 + (size_t)s * stage_sz + 2 * plane_a + (uint32_t)ku * lbo_b + lane_off;
      unsigned char* b_lo = b_hi + plane_bw;
      auto store_a = [&](int m, const float (&v)[4]) { put4(a_hi + m * 16, a_lo + m * 16, v); };
      auto store_b = [&](int n, const float (&v)[4]) { put4(b_hi + n * 16, b_lo + n * 16, v); };
      if (wfast) {
        row_loop<UNR>(rl, rows_a16, [&](int m, float (&v)[4]) {
          const bool ok = m < rows_a;
          const float* q = dyp + (long long)(ok ? m : 0) * G.Tout;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = q[e];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = ok ? v[e] : 0.f;
        }, store_a);
      } else {
        const float* pa[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          pa[e] = et[e] >= 0 ? G.DY + ((long long)eb[e] * G.Cout + grp * G.Cout_g + co_base) * G.Tout + et[e] : G.DY;
        row_loop<UNR>(rl, rows_a16, [&](int m, float (&v)[4]) {
          const bool ok = m < rows_a;
          const long long mo = ok ? (long long)m * G.Tout : 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = pa[e][et[e] >= 0 ? mo : 0];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (ok && et[e] >= 0) ? v[e] : 0.f;
        }, store_a);
      }
      if (wint) {
        const float* xt = xp + ts0;
        const int st = G.stride;
        const int ro0 = roff[0];                         // column 0 of a tile always exists: the safe address
        row_loop<UNR>(rl, rows_b16, [&](int n, float (&v)[4]) {
          const int ro = roff[n];
          const bool ok = ro != INT_MIN;
          const float* q = xt + (ok ? ro : ro0);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = q[e * st];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = ok ? v[e] : 0.f;
        }, store_b);
      } else {
        const float* xb[4];
        int ets[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xb[e] = G.X + (et[e] >= 0 ? ((long long)eb[e] * G.Cin + grp * G.Cin_g) * G.Tin : 0);
          ets[e] = et[e] * G.stride;
        }
        row_loop<UNR>(rl, rows_b16, [&](int n, float (&v)[4]) {
          const int2 ri = rowinfo[n];
          const bool ok = ri.y != INT_MIN;
          const int ky = ok ? ri.y : 0;
          bool val[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int pp = ets[e] + ky;                  // map_pos as selects (no branches between the loads)
            const int p = pp < 0 ? -pp : (pp >= G.Tin ? 2 * (G.Tin - 1) - pp : pp);
            val[e] = ok & (et[e] >= 0) & (pp >= -G.refl) & (pp < G.Tin + G.refl);
            v[e] = xb[e][val[e] ? ri.x + p : 0];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = val[e] ? v[e] : 0.f;
        }, store_b);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
    // ===================== epilogue: TMEM -> fp32 reductions into dW =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int co = co_base + q * 32 + lane;
    const bool ev = co < G.Cout_g;
    float* dst = G.Y + ((long long)grp * G.Cout_g + co) * Ncols;
    const int nblk = NT / 16;
    const int blk_lo = half == 0 ? 0 : (nblk + 1) / 2, blk_hi = half == 0 ? (nblk + 1) / 2 : nblk;
    for (int blk = blk_lo; blk < blk_hi; ++blk) {
      float acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 16), acc);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = nt * NT + blk * 16 + j;
        if (ev && n < Ncols) atomicAdd(dst + n, acc[j]);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(NT, /*a_mn=*/false, /*b_mn=*/false);
      const uint32_t lbo_b = (uint32_t)lbo_wb(NT);
      uint32_t accumulate = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S, use = c / S;
        mbar_wait(&full[s], use & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(stage0 + (size_t)s * stage_sz), a_lo = a_hi + plane_a;
        const uint32_t b_hi = a_lo + plane_a, b_lo = b_hi + plane_bw;
#pragma unroll
        for (int ks = 0; ks < kKC / 16; ++ks) {
          const uint64_t da_hi = make_desc(a_hi + ks * 2 * kLboW, kLboW, 128);
          const uint64_t da_lo = make_desc(a_lo + ks * 2 * kLboW, kLboW, 128);
          const uint64_t db_hi = make_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
          const uint64_t db_lo = make_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
          mma_bf16_ss(tmem_base, da_hi, db_hi, idesc, accumulate);
          accumulate = 1;
          mma_bf16_ss(tmem_base, da_hi, db_lo, idesc, 1);
          mma_bf16_ss(tmem_base, da_lo, db_hi, idesc, 1);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(acc_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

static int64_t pack_bytes(const TcP& P) {
  if (P.slab)
    return (int64_t)P.g.groups * P.ntiles_n * P.sl_ncg * P.sl_nbst * slab_b_stage(P.NT, P.sl_tpb);
  return (int64_t)P.g.groups * P.ntiles_n * P.nphase * P.nchunks * P.nsplit * plane_b(P.NT);
}

extern "C" int64_t vbx_tc_pack_bytes(const vbx_conv_desc* d, int32_t mode, int32_t nsplit) {
  TcP P;
  if (mode != FWD && mode != DGRAD) return -1;
  if (fill_tc(P, d, mode, nsplit)) return -1;
  return pack_bytes(P);
}

extern "C" int vbx_tc_pack(const vbx_conv_desc* d, int32_t mode, int32_t nsplit, const float* w, void* packed,
                           void* stream) {
  TcP P;
  VBX_REQUIRE(mode == FWD || mode == DGRAD, VBX_UNSUPPORTED, "tc_pack: mode must be 0 (fwd) or 1 (dgrad)");
  if (int r = fill_tc(P, d, mode, nsplit)) return r;
  VBX_REQUIRE(w && packed, VBX_BAD_POINTER, "tc_pack: null tensor");
  VBX_REQUIRE(((uintptr_t)packed & 15) == 0, VBX_BAD_POINTER, "tc_pack: packed buffer must be 16-byte aligned");
  MergedDev D = {};
  if (P.merged) {
    GemmP o; fill(o, d);
    if (P.dg) densify(o);
    const MergedPlan M = merged_plan(o);
    for (int i = 0; i < 8; ++i) { D.c[i] = M.c[i]; D.k0[i] = M.k0[i]; D.nt[i] = i < o.stride ? M.nt[i] : 0; }
    D.cmax = M.cmax; D.J = M.J; D.s = o.stride; D.Cin_g = o.Cin_g; D.Cout_g = o.Cout_g; D.K = o.K;
    D.tstep = M.tstep; D.kstep = dgrad_taps(o, 0).kstep; D.dilm = M.dilm;
  }
  if (P.slab) {
    const long long units = (long long)P.g.groups * P.ntiles_n * P.sl_ncg * P.sl_nbst * P.sl_tpb * 2 * P.NT;
    int blocks = (int)((units + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (P.merged) tc_pack_slab_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (unsigned char*)packed, P, D);
    else tc_pack_slab_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (unsigned char*)packed, P, D);
    return launched("tc_pack_slab_kernel");
  }
  if (P.merged) {
    long long units = (long long)P.g.groups * P.ntiles_n * P.nchunks * 4 * P.NT;
    int blocks = (int)((units + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tc_pack_merged_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (unsigned char*)packed, P, D);
    return launched("tc_pack_merged_kernel");
  }
  return mode == FWD ? pack_tc<FWD>(P, w, packed, (cudaStream_t)stream)
                     : pack_tc<DGRAD>(P, w, packed, (cudaStream_t)stream);
}

extern "C" int vbx_tc_conv1d_fwd(const vbx_conv_desc* d, const float* x, const void* packed,
                                 const vbx_epilogue* e, float* y, int32_t nsplit, void* stream) {
  TcP P;
  if (int r = fill_tc(P, d, FWD, nsplit)) return r;
  VBX_REQUIRE(x && packed && y, VBX_BAD_POINTER, "tc_conv1d_fwd: null tensor");
  VBX_REQUIRE(((uintptr_t)packed & 15) == 0, VBX_BAD_POINTER, "tc_conv1d_fwd: packed weights must be 16-byte aligned");
  fill_epi(P.g, e);
  P.g.X = x; P.g.Y = y;
  P.packed = (const unsigned char*)packed;
  if (P.slab) return P.ps ? launch_pslab(P, (cudaStream_t)stream) : launch_slab(P, (cudaStream_t)stream);
  return launch_tc<FWD>(P, (cudaStream_t)stream);
}

extern "C" int vbx_tc_conv1d_dgrad(const vbx_conv_desc* d, const float* dy, const void* packed,
                                   const vbx_epilogue* e, float* dx, int32_t nsplit, void* stream) {
  TcP P;
  if (int r = fill_tc(P, d, DGRAD, nsplit)) return r;
  VBX_REQUIRE(dy && packed && dx, VBX_BAD_POINTER, "tc_conv1d_dgrad: null tensor");
  VBX_REQUIRE(((uintptr_t)packed & 15) == 0, VBX_BAD_POINTER, "tc_conv1d_dgrad: packed weights must be 16-byte aligned");
  fill_epi(P.g, e);
  P.g.X = dy; P.g.Y = dx;
  P.packed = (const unsigned char*)packed;
  VBX_REQUIRE(!P.g.gate || P.g.beta == 0.f, VBX_UNSUPPORTED, "tc_conv1d_dgrad: gate stage with beta != 0");
  const GateArgs gate(P.g, P.slab ? (P.ps ? 4 : 2) : 1);
  int rc;
  if (P.slab) rc = P.ps ? launch_pslab(P, (cudaStream_t)stream) : launch_slab(P, (cudaStream_t)stream);
  else if (P.merged) rc = launch_tc<FWD>(P, (cudaStream_t)stream);
  else rc = launch_tc<DGRAD>(P, (cudaStream_t)stream);
  return gate.finish(rc, dx, d->B, d->Cin, d->Tin, stream);
}

#include "tc_wslab.cuh"

extern "C" int vbx_tc_conv1d_wgrad(const vbx_conv_desc* d, const float* x, const float* dy, float* dw,
                                   void* stream) {
  int code = 0;
  const char* msg = check_desc_msg(d, &code);
  if (msg) return fail(code, msg);
  VBX_REQUIRE(x && dy && dw, VBX_BAD_POINTER, "tc_conv1d_wgrad: null tensor");
  {
    GemmP S;                       // a handful of outputs per input row: streaming dot products (direct_conv.cu), exact fp32
    fill(S, d);
    if (skinny_wgrad_ok(S)) {
      S.X = x; S.DY = dy; S.Y = dw;
      return skinny_wgrad(S, (cudaStream_t)stream);
    }
  }
  {
    TcWS W;
    if (plan_wslab(W, d)) {
      W.g.X = x; W.g.DY = dy; W.g.Y = dw;
      return launch_wslab(W, (cudaStream_t)stream);
    }
  }
  TcW P;
  fill(P.g, d);
  P.g.X = x; P.g.DY = dy; P.g.Y = dw;
  const int Ncols = P.g.Cin_g * P.g.K;
  P.NT = pick_nt(Ncols);
  P.ntiles_n = (Ncols + P.NT - 1) / P.NT;
  P.mtiles = (P.g.Cout_g + kRows - 1) / kRows;
  P.tmem_cols = pow2_cols(P.NT);
  const long long total = (long long)P.g.B * P.g.Tout;
  const long long tiles = (long long)P.ntiles_n * P.mtiles * P.g.groups;
  long long max_split = (total + kKC * 8 - 1) / (kKC * 8);            // >= 8 chunks per slice
  if (max_split > 65535) max_split = 65535;
  if (max_split < 1) max_split = 1;
  long long want = pick_split(tiles, 148 * (P.tmem_cols > 128 ? 2 : 3), max_split);
  if (deterministic_flag()) want = 1;
  long long per = (total + want - 1) / want;
  per = (per + kKC - 1) / kKC * kKC;
  P.red_per = (int)per;
  P.stages = pick_stages_for(wstage_bytes(P.NT), (int)(per / kKC));
  dim3 grid(1, (unsigned)tiles, (unsigned)((total + per - 1) / per));
  VBX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, VBX_UNSUPPORTED, "tc_conv1d_wgrad: grid too large");
  const size_t smem = (size_t)P.stages * wstage_bytes(P.NT) + (2 * P.stages + 1) * sizeof(uint64_t) + 16 +
                      (size_t)P.NT * (sizeof(int2) + sizeof(int)) + 2 * sizeof(int) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_wgrad_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(tc_wgrad_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce != cudaSuccess) return fail((int)ce, "tc_conv1d_wgrad: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  if (P.tmem_cols > 128) tc_wgrad_kernel<2, 8><<<grid, kThreads, smem, (cudaStream_t)stream>>>(P);
  else tc_wgrad_kernel<3, 4><<<grid, kThreads, smem, (cudaStream_t)stream>>>(P);
  return launched("tc_wgrad_kernel");
}
