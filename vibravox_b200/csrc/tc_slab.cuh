// Tensor-core conv WITHOUT im2col replication ("slab" form), included by tc_conv.cu.
//
// The gather kernel (tc_conv_kernel) rebuilds an im2col'd A tile per 32-element reduction chunk, so every input
// value is loaded, split and stored K/stride times.  Here the producers stage each input value ONCE:
//   * all batch items are laid on one virtual timeline with period P = s*R positions per item
//     (R = Tout - 1 + ceil(((K-1)*d + 1) / s) virtual output rows, of which the last R - Tout are discarded), so
//     output row rv reads positions s*rv + k*d for every tap k - uniformly, across item boundaries;
//   * a stage holds the bf16 hi/lo split of 16 channels x the 127*s + (K-1)*d + 1 positions a 128-row tile touches,
//     stored K-major: one 16-byte unit = 8 channels of one position, units of one stride phase consecutive
//     (position q = s*u + rho lives at [rho][u]);
//   * tap k (offset k*d = s*j + rho) is then simply the same slab read through a descriptor whose start address is
//     advanced by (rho*U + j) units: rows are 16 bytes apart (SBO = 128), the two 8-channel halves one plane apart
//     (LBO), so each tap costs three MMAs and no data movement.
// Weights are pre-packed per (16-channel group, tap) and streamed with bulk copies through their own ring.
// Used for forward convs and for merged-phase input gradients (which are stride-1 forward convs on dy).
#pragma once

static const int kSlabMaxU = 5;             // positions a producer thread stages per channel group

// rt = row tiles (128 rows each) a CTA of the streaming kernel works on at once: their positions are adjacent on the
// virtual timeline, so ONE slab of rt*128 rows (+ the shared tap overhang) serves rt accumulators - see tc_slab_kernel
__host__ __device__ inline int slab_npos(const GemmP& G, int rt = 1) { return (kRows * rt - 1) * G.stride + (G.K - 1) * G.dil + 1; }
__host__ __device__ inline int slab_U(const GemmP& G, int rt = 1) { return kRows * rt + ((G.K - 1) * G.dil) / G.stride + 1; }
__host__ __device__ inline int slab_a_stage(const GemmP& G, int rt = 1) { return 64 * G.stride * slab_U(G, rt); }   // 2 planes x 2 halves
__host__ __device__ inline int slab_b_stage(int NT, int tpb) { return tpb * NT * 64; }

// Epilogue of one 128-row tile: the eight producer warps read the accumulator (warp & 3 = TMEM lane quadrant,
// warp >> 2 = which half of the 16-column blocks) and write bias / LeakyReLU / residual / mask fused, coalesced
// along time.  Shared by the streaming (tc_slab_kernel) and the persistent (tc_pslab_kernel) forms.
__device__ __forceinline__ void slab_epilogue(const TcP& P, uint32_t tmem_base, int rv0, int nt, int grp, int warp,
                                              int lane) {
  const GemmP& G = P.g;
  const int NT = P.NT, R = P.sl_R, Ccol = G.Cout_g;
  const int q = warp & 3, half = warp >> 2;
  const int rv = rv0 + q * 32 + lane;
  const int eb = rv / R, et = rv % R;
  const bool ev = eb < G.B && et < G.Tout;
  const int nblk = NT / 16;
  const int blk_lo = half == 0 ? 0 : (nblk + 1) / 2, blk_hi = half == 0 ? (nblk + 1) / 2 : nblk;
  for (int blk = blk_lo; blk < blk_hi; ++blk) {
    float acc[16];
    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 16), acc);
    if (!ev) continue;
    const int cb = nt * NT + blk * 16;
    if (P.merged) {
      // rows are (b, v); column = phase * Cin_g + ci  ->  dx[b, ci, r(phase) + s*v]
      int col = cb;
      int ph = col / P.mg_Cing, ci = col % P.mg_Cing;
#pragma unroll
      for (int j = 0; j < 16; ++j, ++col) {
        if (col < Ccol) {
          const int u = P.mg_r[ph] + P.mg_s * et;
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
    const long long out_base = ((long long)eb * G.Cout + grp * Ccol) * G.Tout + et;
    const int Tlen = G.Tout;
    if (cb + 16 <= Ccol && G.beta == 0.f) {
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
    for (int j = 0; j < 16; ++j) {                        // (unrolled: acc[] stays in registers, no local-memory frame)
      const int col = cb + j;
      if (col < Ccol) {
        const long long idx = out_base + (long long)col * Tlen;
        G.Y[idx] = finish(G, acc[j], grp * Ccol + col, idx);
      }
    }
  }
}

// Wide tiles (N >= 128 with long reductions) are bound by the weight tiles they stream from L2: 16 KB per tap and
// 16-channel group against 3 x 128 tensor-core cycles = 43 bytes / cycle / SM, i.e. ~12 TB/s over 148 SMs - more
// than L2 delivers (measured: tensor pipe 66 % active).  With sl_rt = 2 a CTA owns TWO adjacent 128-row tiles: one
// slab of 256 rows, two accumulators in TMEM, and every weight tile that arrives feeds 6 MMAs instead of 3 - half
// the L2 -> SM weight traffic per MMA.
__global__ void __launch_bounds__(kThreads, 3) tc_slab_kernel(const TcP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmP& G = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = P.NT, SA = P.sl_SA, SB = P.sl_SB, s = G.stride, rt = P.sl_rt;
  const int U = slab_U(G, rt);
  const int a_stage = slab_a_stage(G, rt), b_stage = slab_b_stage(NT, P.sl_tpb);
  const int plane_a = a_stage / 2, half_a = plane_a / 2;      // hi / lo planes; channels 0-7 / 8-15 inside a plane
  unsigned char* a0 = smem;
  unsigned char* b0 = smem + (size_t)SA * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b0 + (size_t)SB * b_stage);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + SA;
  uint64_t* full_b = bars + 2 * SA;
  uint64_t* empty_b = bars + 2 * SA + SB;
  uint64_t* acc_full = bars + 2 * SA + 2 * SB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  int* tapoff = reinterpret_cast<int*>(tmem_slot + 2);        // per tap: byte offset of its view into a stage

  const int grp = blockIdx.y / P.ntiles_n, nt = blockIdx.y % P.ntiles_n;
  const int R = P.sl_R, Ppos = R * s;
  const int ncg = P.sl_ncg, nbst = P.sl_nbst, tpb = P.sl_tpb;
  const int rv0 = blockIdx.x * kRows * rt;                     // first virtual output row of this CTA's tile(s)
  const uint32_t dcols = (uint32_t)P.tmem_cols / (uint32_t)rt; // TMEM columns per accumulator

  if (tid == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(&full_a[i], kProducers); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  for (int k = tid; k < G.K; k += kThreads) {
    const int off = k * G.dil;
    tapoff[k] = ((off % s) * U + off / s) * 16;
  }
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===================== producers: stage the slab, 16 channels per stage =====================
    const int npos = slab_npos(G, rt);
    int goff[kSlabMaxU], soff[kSlabMaxU];                      // per staged position: x offset (or -1) / smem offset
#pragma unroll
    for (int n = 0; n < kSlabMaxU; ++n) {
      const int i = tid + n * kProducers;
      goff[n] = -1; soff[n] = -1;
      if (i < npos) {
        const unsigned q = (unsigned)rv0 * (unsigned)s + (unsigned)i;
        const int b = (int)(q / (unsigned)Ppos), p = (int)(q % (unsigned)Ppos);
        const int tau = map_pos(p - G.pad, G.Tin, G.refl);
        if (b < G.B && tau >= 0) goff[n] = b * G.Cin * G.Tin + tau;
        soff[n] = ((i % s) * U + i / s) * 16;
      }
    }
    const float* xg = G.X + (long long)grp * G.Cin_g * G.Tin;
    int sa = 0;
    uint32_t para = 0;
    for (int cg = 0; cg < ncg; ++cg) {
      const int nch = min(16, G.Cin_g - cg * 16);              // valid channels of this group
      const float* xc = xg + (long long)cg * 16 * G.Tin;
      mbar_wait(&empty_a[sa], para ^ 1u);
      unsigned char* st = a0 + (size_t)sa * a_stage;
#pragma unroll
      for (int n = 0; n < kSlabMaxU; ++n) {
        if (soff[n] >= 0) {                                    // (uniform per n except in the last pass)
          const bool pv = goff[n] >= 0;
          const float* src = xc + (pv ? goff[n] : 0);
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const bool ok = pv && e < nch;
            v[e] = src[ok ? (long long)e * G.Tin : 0];
            v[e] = ok ? v[e] : 0.f;
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
          unsigned char* d0 = st + soff[n];
          *reinterpret_cast<uint4*>(d0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(d0 + half_a) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<uint4*>(d0 + plane_a) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(d0 + plane_a + half_a) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full_a[sa]);
      if (++sa == SA) { sa = 0; para ^= 1u; }
    }
    // ===================== epilogue =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int r = 0; r < rt; ++r)
      slab_epilogue(P, tmem_base + (uint32_t)r * dcols, rv0 + r * kRows, nt, grp, warp, lane);
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(NT, /*a_mn=*/false, /*b_mn=*/false);
      const uint32_t lbo_a = (uint32_t)half_a, lbo_b = (uint32_t)NT * 16, plane_bt = (uint32_t)NT * 32;
      uint32_t accumulate = 0;
      int sa = 0, sb = 0;
      uint32_t para = 0, parb = 0;
      for (int cg = 0; cg < ncg; ++cg) {
        mbar_wait(&full_a[sa], para);
        const uint32_t abase = smem_u32(a0 + (size_t)sa * a_stage);
        int tap = 0;
        for (int bs = 0; bs < nbst; ++bs) {
          mbar_wait(&full_b[sb], parb);
          tc_fence_after();
          const uint32_t bbase = smem_u32(b0 + (size_t)sb * b_stage);
          for (int tt = 0; tt < tpb && tap < G.K; ++tt, ++tap) {
            const uint32_t a_hi = abase + (uint32_t)tapoff[tap];
            const uint32_t b_hi = bbase + (uint32_t)tt * (uint32_t)NT * 64u;
            const uint64_t db_hi = make_desc(b_hi, lbo_b, 128), db_lo = make_desc(b_hi + plane_bt, lbo_b, 128);
            for (int r = 0; r < rt; ++r) {                     // row tile r: the same slab, 128 rows (units) further on
              const uint32_t ar = a_hi + (uint32_t)r * (uint32_t)(kRows * 16), dr = tmem_base + (uint32_t)r * dcols;
              const uint64_t da_hi = make_desc(ar, lbo_a, 128), da_lo = make_desc(ar + plane_a, lbo_a, 128);
              mma_bf16_ss(dr, da_hi, db_hi, idesc, accumulate);
              mma_bf16_ss(dr, da_hi, db_lo, idesc, 1);
              mma_bf16_ss(dr, da_lo, db_hi, idesc, 1);
            }
            accumulate = 1;
          }
          mma_commit(&empty_b[sb]);
          if (++sb == SB) { sb = 0; parb ^= 1u; }
        }
        mma_commit(&empty_a[sa]);
        if (++sa == SA) { sa = 0; para ^= 1u; }
      }
      mma_commit(acc_full);
    }
  } else {
    // ===================== weight tiles: one bulk copy per B stage =====================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)b_stage;
      const unsigned char* src = P.packed + ((size_t)(grp * P.ntiles_n + nt) * ncg) * nbst * bytes;
      int sb = 0;
      uint32_t parb = 0;
      for (int i = 0; i < ncg * nbst; ++i) {
        mbar_wait(&empty_b[sb], parb ^ 1u);
        mbar_expect_tx(&full_b[sb], bytes);
        bulk_copy_g2s(b0 + (size_t)sb * b_stage, src + (size_t)i * bytes, bytes, &full_b[sb]);
        if (++sb == SB) { sb = 0; parb ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// Weight pre-pack for the slab kernel: one thread per 16-byte unit = 8 consecutive channels of one (tap, column).
// Stage layout [tt][hi|lo][half][n][8];  stages ordered [group][ntile][channel group][b stage].
//   forward : value(col, c, tap) = W[g*Cout_g + col][c][tap]
//   merged  : value((ph,ci), co, m') = W[co][ci][k0(ph) + jj*s],  jj = (J-1-m') - (cmax - c(ph))
template <bool MERGED>
__global__ void tc_pack_slab_kernel(const float* __restrict__ w, unsigned char* __restrict__ out, const TcP P,
                                    const MergedDev M) {
  const GemmP& G = P.g;
  const int NT = P.NT, tpb = P.sl_tpb;
  const long long units = (long long)G.groups * P.ntiles_n * P.sl_ncg * P.sl_nbst * tpb * 2 * NT;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < units;
       u += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(u % NT);
    long long r = u / NT;
    const int hf = (int)(r % 2); r /= 2;
    const int tt = (int)(r % tpb); r /= tpb;
    const int bs = (int)(r % P.sl_nbst); r /= P.sl_nbst;
    const int cg = (int)(r % P.sl_ncg); r /= P.sl_ncg;
    const int nt = (int)(r % P.ntiles_n);
    const int g = (int)(r / P.ntiles_n);
    const int col = nt * NT + n, tap = bs * tpb + tt, c0 = cg * 16 + hf * 8;
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      float v = 0.f;
      if (col < G.Cout_g && c < G.Cin_g && tap < G.K) {
        if (!MERGED) {
          if (!P.dg) {
            v = w[(((long long)g * G.Cout_g + col) * G.Cin_g + c) * G.K + tap];
          } else {                                   // block-diagonal: channel c feeds column col only inside its group
            const int cl = c - (col / P.dg_cout) * P.dg_cin;
            if (cl >= 0 && cl < P.dg_cin) v = w[((long long)col * P.dg_cin + cl) * G.K + tap];
          }
        } else {
          const int ph = col / M.Cin_g, ci = col % M.Cin_g;
          const int jj = merged_tap(M, ph, tap);
          if (jj >= 0) {
            if (!P.dg) {
              v = w[(((long long)g * M.Cout_g + c) * M.Cin_g + ci) * M.K + M.k0[ph] + jj * M.kstep];
            } else {                                 // c = output channel of the conv, ci = its input channel (dense)
              const int cl = ci - (c / P.dg_cout) * P.dg_cin;
              if (cl >= 0 && cl < P.dg_cin)
                v = w[((long long)c * P.dg_cin + cl) * M.K + M.k0[ph] + jj * M.kstep];
            }
          }
        }
      }
      split_bf16(v, hi[e], lo[e]);
    }
    unsigned char* stage = out + ((((size_t)(g * P.ntiles_n + nt) * P.sl_ncg + cg) * P.sl_nbst + bs)) *
                                     (size_t)slab_b_stage(NT, tpb);
    unsigned char* tile = stage + (size_t)tt * NT * 64;
    const size_t off = ((size_t)hf * NT + n) * 16;
    *reinterpret_cast<uint4*>(tile + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(tile + (size_t)NT * 32 + off) = *reinterpret_cast<const uint4*>(lo);
  }
}
