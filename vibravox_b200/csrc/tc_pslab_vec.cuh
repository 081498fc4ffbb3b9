// Persistent slab conv, 16-byte staging (stride-1 layers whose length is a multiple of 4), included by tc_conv.cu.
// EXPERIMENTAL (VBX_TC_PS_VEC=1): same kernel as tc_pslab_kernel except that the producers stage items of
// 4 samples x 8 channels with float4 loads from a slab whose origin is aligned down to a multiple of 4 samples
// (ps_vec_stage.h, executed on the CPU by tests/emu/emu_slab.cpp).  Motivation: the ncu capture of the scalar form
// (profiles/r1_ncu_full_pslab_kernel.csv) shows a load-latency chain with lg_throttle stalls; this form issues a
// quarter of the load instructions and has twice the bytes in flight per thread.  Known cost: the 16-byte slab
// stores of a quarter-warp fall on two bank groups (4-way conflict); rotate the sample order per lane if it shows.
// Not yet run on a GPU - the round's GPU budget was spent when it was written.  It is a copy of tc_pslab_kernel with
// the staging lambda and the per-tile descriptor shift replaced; once measured, fold the two into one kernel
// templated on the staging (or drop this file).
#pragma once
// (ps_vec_stage.h is included by tc_conv.cu at file scope: this header sits inside namespace vbx::tc)

__global__ void __launch_bounds__(kThreads, 3) tc_pslab_vec_kernel(const TcP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmP& G = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = P.NT, ncg = P.sl_ncg, slots = P.ps_slots;
  const PsVec V = ps_vec_geom(G);
  const int a_stage = V.a_stage, w_stage = pslab_w_stage(P), slot_bytes = ncg * V.a_stage;
  const int plane_a = V.plane, half_a = V.half;
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
  const int R = P.sl_R;                                       // stride 1: one virtual position per output row
  const int row_tiles = (int)(((long long)G.B * R + kRows - 1) / kRows);
  const int my_tiles = ((int)blockIdx.x < row_tiles) ? (row_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < slots; ++i) { mbar_init(&full_a[i], kProducers); mbar_init(&empty_a[i], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kProducers); }
    fence_barrier_init();
  }
  for (int k = tid; k < G.K; k += kThreads) tapoff[k] = k * G.dil * 16;   // + a*16 per tile (ps_vec_tile)
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t bufcols = (uint32_t)pslab_bufcols(P);

  if (warp < 8) {
    int slot = 0;
    uint32_t par = 0;
    // stage the slab of row tile `tile` into the next ring slot: items of 4 samples x 8 channels, 16-byte loads
    auto stage_tile = [&](int tile) {
      mbar_wait(&empty_a[slot], par ^ 1u);
      unsigned char* st = a0 + (size_t)slot * slot_bytes;
      int q_origin, a;
      ps_vec_tile(G, R, tile, q_origin, a);
      for (int item = tid; ps_vec_stage_item(G, V, R, grp, ncg, q_origin, item, st, P.ps_vec_aligned != 0);
           item += kProducers) {
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
        int q_origin, a;
        ps_vec_tile(G, R, (int)blockIdx.x + it * (int)gridDim.x, q_origin, a);
        const uint32_t ashift = (uint32_t)a * 16u;            // the slab starts `a` samples before the tile's first
        uint32_t accumulate = 0;
        for (int cg = 0; cg < ncg; ++cg) {
          const uint32_t abase = smem_u32(a0 + (size_t)slot * slot_bytes + (size_t)cg * a_stage);
          const uint32_t bbase = smem_u32(w0 + (size_t)cg * w_stage);
          for (int tap = 0; tap < G.K; ++tap) {
            const uint32_t a_hi = abase + (uint32_t)tapoff[tap] + ashift;
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
