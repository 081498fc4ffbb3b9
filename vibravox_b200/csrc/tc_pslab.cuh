// Persistent slab conv with RESIDENT weights, included by tc_conv.cu after tc_slab.cuh.
//
// On the narrow layers (generator residual units, first discriminator stages, densified groups) the streaming
// slab kernel is bound by what every 128-row tile repeats: a fresh CTA (barrier init, TMEM allocation), and above
// all the whole packed weight set pulled from L2 again - 12 to 170 KB per tile against the 20 to 35 KB of
// activations the tile reads.  Here a CTA stays on its SM and walks row tiles blockIdx.x, blockIdx.x + gridDim.x, ...:
//   * the packed weights of its (group, column tile) are bulk-copied into shared memory ONCE;
//   * the slab of a tile (all 16-channel groups) is one slot of a small ring, staged by the eight producer warps
//     as a flat list of (position, channel group) items so that all 256 threads load whatever the tile shape;
//   * the accumulator is double-buffered in TMEM: the MMAs of tile i+1 run while the same eight warps drain
//     tile i (they stage tile i+1 BEFORE draining tile i, so its loads are in flight under the epilogue stores).
// Same staging layout, descriptors, weight pack format (one stage per channel group holding all K taps) and
// epilogue as tc_slab_kernel.
#pragma once

static const int kPsMaxBulk = 32768;        // bytes per bulk copy of the resident weights

__host__ __device__ inline int pslab_w_stage(const TcP& P) { return slab_b_stage(P.NT, P.g.K); }
__host__ __device__ inline int pslab_slot(const TcP& P) { return P.sl_ncg * slab_a_stage(P.g); }
__host__ __device__ inline int pslab_bufcols(const TcP& P) { return P.tmem_cols / 2; }

__global__ void __launch_bounds__(kThreads, 3) tc_pslab_kernel(const TcP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmP& G = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = P.NT, s = G.stride, ncg = P.sl_ncg, slots = P.ps_slots;
  const int U = slab_U(G);
  const int a_stage = slab_a_stage(G), w_stage = pslab_w_stage(P), slot_bytes = pslab_slot(P);
  const int plane_a = a_stage / 2, half_a = plane_a / 2;
  unsigned char* w0 = smem;                                   // resident weights: [channel group][tap][hi|lo][half][n][8]
  unsigned char* a0 = smem + (size_t)ncg * w_stage;           // slab ring: [slot][channel group]
  uint64_t* bars = reinterpret_cast<uint64_t*>(a0 + (size_t)slots * slot_bytes);
  uint64_t* full_a = bars;                                    // [slots]  256 producer arrivals
  uint64_t* empty_a = bars + slots;                           // [slots]  MMAs of the tile retired
  uint64_t* w_full = bars + 2 * slots;
  uint64_t* acc_full = w_full + 1;                            // [2]      accumulator buffer complete
  uint64_t* acc_empty = acc_full + 2;                         // [2]      256 epilogue arrivals: buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* tapoff = reinterpret_cast<int*>(tmem_slot + 2);

  const int grp = blockIdx.y / P.ntiles_n, nt = blockIdx.y % P.ntiles_n;
  const int R = P.sl_R, Ppos = R * s;
  const int row_tiles = (int)(((long long)G.B * R + kRows - 1) / kRows);
  const int my_tiles = ((int)blockIdx.x < row_tiles) ? (row_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < slots; ++i) { mbar_init(&full_a[i], kProducers); mbar_init(&empty_a[i], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kProducers); }
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
  const uint32_t bufcols = (uint32_t)pslab_bufcols(P);

  if (warp < 8) {
    const int npos = slab_npos(G);
    const float* xg = G.X + (long long)grp * G.Cin_g * G.Tin;
    int slot = 0;
    uint32_t par = 0;
    // stage the slab of row tile `tile` into the next ring slot: items (position i, channel group cg), item = cg*npos + i
    auto stage_tile = [&](int tile) {
      mbar_wait(&empty_a[slot], par ^ 1u);
      unsigned char* st = a0 + (size_t)slot * slot_bytes;
      const unsigned q0 = (unsigned)tile * (unsigned)(kRows * s);
      int i = tid, cg = 0;
      while (i >= npos) { i -= npos; ++cg; }
      while (cg < ncg) {
        const unsigned q = q0 + (unsigned)i;
        const int b = (int)(q / (unsigned)Ppos), p = (int)(q % (unsigned)Ppos);
        const int tau = map_pos(p - G.pad, G.Tin, G.refl);
        const bool pv = b < G.B && tau >= 0;
        const int nch = min(16, G.Cin_g - cg * 16);
        const float* src = xg + (long long)cg * 16 * G.Tin + (pv ? b * G.Cin * G.Tin + tau : 0);
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
        unsigned char* d0 = st + (size_t)cg * a_stage + ((i % s) * U + i / s) * 16;
        *reinterpret_cast<uint4*>(d0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(d0 + half_a) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(d0 + plane_a) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(d0 + plane_a + half_a) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        i += kProducers;
        while (i >= npos) { i -= npos; ++cg; }
      }
      fence_proxy_async();
      mbar_arrive(&full_a[slot]);
      if (++slot == slots) { slot = 0; par ^= 1u; }
    };
    if (my_tiles > 0) stage_tile((int)blockIdx.x);
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      if (it + 1 < my_tiles) stage_tile(tile + (int)gridDim.x);
      mbar_wait(&acc_full[it & 1], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      slab_epilogue(P, tmem_base + (uint32_t)(it & 1) * bufcols, tile * kRows, nt, grp, warp, lane);
      tc_fence_before();
      mbar_arrive(&acc_empty[it & 1]);
    }
  } else if (warp == 8) {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = make_idesc_bf16(NT, /*a_mn=*/false, /*b_mn=*/false);
      const uint32_t lbo_a = (uint32_t)half_a, lbo_b = (uint32_t)NT * 16, plane_bt = (uint32_t)NT * 32;
      mbar_wait(w_full, 0);
      int slot = 0;
      uint32_t par = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(&acc_empty[buf], (uint32_t)(((it >> 1) - 1) & 1));
        mbar_wait(&full_a[slot], par);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * bufcols;
        uint32_t accumulate = 0;
        for (int cg = 0; cg < ncg; ++cg) {
          const uint32_t abase = smem_u32(a0 + (size_t)slot * slot_bytes + (size_t)cg * a_stage);
          const uint32_t bbase = smem_u32(w0 + (size_t)cg * w_stage);
          for (int tap = 0; tap < G.K; ++tap) {
            const uint32_t a_hi = abase + (uint32_t)tapoff[tap];
            const uint32_t b_hi = bbase + (uint32_t)tap * (uint32_t)NT * 64u;
            const uint64_t da_hi = make_desc(a_hi, lbo_a, 128), da_lo = make_desc(a_hi + plane_a, lbo_a, 128);
            const uint64_t db_hi = make_desc(b_hi, lbo_b, 128), db_lo = make_desc(b_hi + plane_bt, lbo_b, 128);
            mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, accumulate);
            accumulate = 1;
            mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1);
            mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1);
          }
        }
        mma_commit(&empty_a[slot]);
        mma_commit(&acc_full[buf]);
        if (++slot == slots) { slot = 0; par ^= 1u; }
      }
    }
  } else {
    if (lane == 0 && my_tiles > 0) {
      const uint32_t total = (uint32_t)ncg * (uint32_t)w_stage;
      const unsigned char* src = P.packed + (size_t)(grp * P.ntiles_n + nt) * total;
      mbar_expect_tx(w_full, total);
      for (uint32_t off = 0; off < total; off += kPsMaxBulk) {
        const uint32_t n = total - off < (uint32_t)kPsMaxBulk ? total - off : (uint32_t)kPsMaxBulk;
        bulk_copy_g2s(w0 + off, src + off, n, w_full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}
