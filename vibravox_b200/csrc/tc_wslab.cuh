// Weight gradient on tensor cores WITHOUT im2col replication ("slab" form), included by tc_conv.cu (global scope,
// after `using namespace vbx::tc`).
//
//   dW[co, ci, k] = sum_{(b,t)} dy[b, co, t] * x[b, ci, map(s*t + k*d - pad)]
//
// One MMA family per tap:  D_k[co (M = 128), ci (N = NCI)] += A[co, t] * B_k[ci, t]  with the reduction over time.
// Both operands are staged MN-major - a 16-byte unit holds 8 consecutive CHANNELS of one time position and
// consecutive positions are 16 bytes apart - so B_k is the SAME staged x slab for every tap, read through a
// descriptor whose start address is advanced by the tap offset (k*d = s*j + rho -> phase plane rho, j units).
// Every x / dy value is loaded, split into bf16 hi/lo and stored once per (co tile, ci tile, tap group) instead of
// once per tap column.  TG taps accumulate side by side in TMEM (TG * NCI <= 512 columns); time runs over the
// virtual timeline of tc_slab.cuh (period R rows per batch item; dy is zero on the rows that are not real) and is
// split over blockIdx.z; partial results are added to dW with fp32 atomics.
#pragma once

static const int kWsTC = 32;                // time positions per stage (two MMA k-steps)
static const int kWsMaxItems = 4;           // x items (position, 8-channel group) a producer thread carries per stage

struct TcWS {
  GemmP g;
  int NCI, TG, ntg, ci_tiles, co_tiles, tmem_cols, stages;
  int R;            // virtual rows per batch item
  int UB;           // x units per stride phase in a stage
  int rows_per;     // virtual rows per blockIdx.z (multiple of kWsTC)
  int ts;           // 1: dy (the A operand) is staged in TENSOR memory (two buffers of 32 columns behind the accumulators)
};
static const int kWsTsCols = 64;            // TMEM columns of the A ring in ts mode: 2 buffers x (hi | lo) x 2 k-steps x 8
__host__ __device__ inline int ws_UB(const GemmP& G, int TG) { return kWsTC + ((TG - 1) * G.dil) / G.stride + 1; }
// bytes between consecutive 8-channel groups of a slab: an odd number of 16-byte units, so that the 16 (or NCI/8)
// units the tensor core fetches for one time position fall in different shared-memory banks
__host__ __device__ inline int ws_sbo_a() { return kWsTC * 16 + 16; }
__host__ __device__ inline int ws_sbo_b(const GemmP& G, int UB) { return (G.stride * UB | 1) * 16; }
// (ts mode: the A part of a stage is the fp32 transpose scratch [128 channels][32 positions, row pitch 36 floats])
__host__ __device__ inline int ws_a_stage(int ts = 0) { return ts ? kRows * 36 * 4 : 2 * (kRows / 8) * ws_sbo_a(); }   // hi + lo
__host__ __device__ inline int ws_b_stage(const GemmP& G, int NCI, int UB) { return 2 * (NCI / 8) * ws_sbo_b(G, UB); }
// positions of x a tile with `ntap` taps stages per time chunk (its own tap span only)
__host__ __device__ inline int ws_npos_taps(const GemmP& G, int ntap) {
  return (kWsTC - 1) * G.stride + (ntap - 1) * G.dil + 1;
}

// PW producer warps (+ one MMA warp): 16 when the tile owns all 512 TMEM columns (one CTA per SM: the extra warps
// are the memory-level parallelism), 8 with two CTAs per SM otherwise.
// NI MMA-issuing warps (one thread each): an N = 64 MMA is 32 tensor-core cycles of work, but ONE thread cannot issue
// them faster than one per ~76 cycles (profiles/r1_mma_probe.txt: 76 per issuer, 48 aggregate with four issuers), so
// with a single issuer the tensor pipe idled at 26-42 %.  The taps of a CTA accumulate into separate TMEM columns, so
// they are dealt round-robin to the issuers and no two threads ever touch the same accumulator.
static const int kWsIssuers = 4;
template <int PW, int MINB, int MAXI>
__global__ void __launch_bounds__(PW * 32 + kWsIssuers * 32, MINB) tc_wslab_kernel(const TcWS P) {
  constexpr int kProd = PW * 32;
  constexpr int NI = kWsIssuers;
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmP& G = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = P.stages, NCI = P.NCI, s = G.stride, UB = P.UB;
  const int a_stage = ws_a_stage(P.ts), b_stage = ws_b_stage(G, NCI, UB), stage_sz = a_stage + b_stage;
  const uint32_t acol0 = (uint32_t)(P.TG * NCI);                  // ts mode: first TMEM column of the A ring
  const int plane_a = a_stage / 2, plane_b = b_stage / 2;
  const int sbo_a = ws_sbo_a(), sbo_b = ws_sbo_b(G, UB);            // bytes between 8-channel groups
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_sz);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  int* tapoff = reinterpret_cast<int*>(tmem_slot + 2);          // per tap of this group: byte offset into the x slab

  // blockIdx.x = tap group * ci_tiles + ci tile;  blockIdx.y = group * co_tiles + co tile;  blockIdx.z = time slice
  const int tg = blockIdx.x / P.ci_tiles, cit = blockIdx.x % P.ci_tiles;
  const int grp = blockIdx.y / P.co_tiles, cot = blockIdx.y % P.co_tiles;
  const int tap0 = tg * P.TG, ntap = min(P.TG, G.K - tap0);
  const int co0 = cot * kRows, ci0 = cit * NCI;
  const int R = P.R, Ppos = R * s;
  const int total = G.B * R;
  const int rv_lo = blockIdx.z * P.rows_per, rv_hi = min(rv_lo + P.rows_per, total);
  if (rv_lo >= rv_hi) return;
  const int nstage = (rv_hi - rv_lo + kWsTC - 1) / kWsTC;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], kProd); mbar_init(&empty[i], NI); }
    mbar_init(acc_full, NI);
    fence_barrier_init();
  }
  const int pos0 = tap0 * G.dil;                                // first x position (relative to s*rv) this tile reads
  if (tid < ntap) {
    const int off = tid * G.dil;                                // relative to pos0
    tapoff[tid] = ((off % s) * UB + off / s) * 16;
  }
  if (warp == PW) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PW) {
    // ===================== producers =====================
    // dy: one position x (16 / PW) 8-channel groups per thread.  x: a flat list of (position, 8-channel group) items,
    // item = group * npos + position, dealt round-robin to the producer threads (lanes walk positions: coalesced rows),
    // so every thread carries the same number of loads whatever the tap span.
    // The loads of stage c + 1 are ISSUED before stage c is converted and stored (two register sets, loop unrolled by
    // two): a stage then costs its conversion + stores, not a global-memory round trip - with the loads issued and
    // consumed inside one stage this kernel sat at 26 % tensor-pipe activity, all 16 warps waiting on long scoreboards.
    constexpr int NA = 16 / PW;
    // MAXI (template): x items per thread and stage (host: nitems <= MAXI * kProd)
    const int npos = ws_npos_taps(G, ntap);
    const int ncg = NCI / 8;
    const int nitems = npos * ncg;
    const int tA = tid & 31, cogA = (tid >> 5) * NA;
    // ts mode: thread = TMEM lane (output channel co0 + 32 * (warp % 4) + lane); the PW / 4 warps of a lane quadrant
    // split the stage's 32 time positions (8 * NA each)
    const int rowT = (warp & 3) * 32 + lane, partT = warp >> 2;
    const float* dyg = G.DY + (long long)(grp * G.Cout_g + co0) * G.Tout;
    const float* xg = G.X + (long long)(grp * G.Cin_g + ci0) * G.Tin;
    const int rows_a = min(kRows, G.Cout_g - co0), cols_b = min(NCI, G.Cin_g - ci0);
    auto put16 = [](unsigned char* hi_p, unsigned char* lo_p, const float (&v)[8]) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __bfloat1622float2(h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      *reinterpret_cast<uint4*>(hi_p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(lo_p) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    };
    // 8 channels (rows `stride` apart) of one position; `nvalid` of them exist, the rest read as zero
    auto load8 = [](const float* p, long long stride, int nvalid, float (&v)[8]) {
      if (nvalid >= 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(p + e * stride);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = e < nvalid ? __ldg(p + e * stride) : 0.f;
      }
    };
    // running (batch item, offset) of this thread's dy row and x items: advanced per stage, never divided again
    int ba = (rv_lo + tA) / R, ta = (rv_lo + tA) % R;
    int bx[MAXI], px[MAXI], sx[MAXI], gx[MAXI];
#pragma unroll
    for (int j = 0; j < MAXI; ++j) {
      const int i = tid + j * kProd;
      const int g = i < nitems ? i / npos : 0, pp = i < nitems ? i - g * npos : 0;
      const unsigned q = (unsigned)rv_lo * (unsigned)s + (unsigned)(pos0 + pp);
      bx[j] = (int)(q / (unsigned)Ppos); px[j] = (int)(q % (unsigned)Ppos);
      gx[j] = g;
      sx[j] = i < nitems ? ((pp % s) * UB + pp / s) * 16 + g * sbo_b : -1;
    }
    // issue the loads of stage c (call in stage order: advances the running counters)
    auto issue = [&](int c, float (&a)[NA][8], float (&v)[MAXI][8]) {
      const bool va = rv_lo + c * kWsTC + tA < rv_hi && ba < G.B && ta < G.Tout;
      const float* pa = dyg + ((long long)ba * G.Cout + (long long)cogA * 8) * G.Tout + ta;
#pragma unroll
      for (int h = 0; h < NA; ++h)
        load8(pa + (long long)h * 8 * G.Tout, G.Tout, va ? rows_a - (cogA + h) * 8 : 0, a[h]);
#pragma unroll
      for (int j = 0; j < MAXI; ++j) {
        if (sx[j] >= 0) {
          const int tau = map_pos(px[j] - G.pad, G.Tin, G.refl);
          const bool vb = bx[j] < G.B && tau >= 0;
          const float* pb = xg + ((long long)bx[j] * G.Cin * G.Tin + (vb ? tau : 0));
          load8(pb + (long long)gx[j] * 8 * G.Tin, G.Tin, vb ? cols_b - gx[j] * 8 : 0, v[j]);
        }
      }
      ta += kWsTC;
      while (ta >= R) { ta -= R; ++ba; }
#pragma unroll
      for (int j = 0; j < MAXI; ++j) {
        px[j] += kWsTC * s;
        while (px[j] >= Ppos) { px[j] -= Ppos; ++bx[j]; }
      }
    };
    // convert + store stage c and hand it to the MMA issuer
    auto commit = [&](int c, const float (&a)[NA][8], const float (&v)[MAXI][8]) {
      const int st = c % S;
      mbar_wait(&empty[st], (uint32_t)((c / S) & 1) ^ 1u);
      unsigned char* sa = smem + (size_t)st * stage_sz;
      unsigned char* sb = sa + a_stage;
      if (P.ts) {
        // The loads are coalesced along time (lane = position, 8 channels each); tensor memory wants lane = channel.
        // Transpose through the stage's fp32 scratch (row pitch 36 floats: conflict-free both ways), then each thread
        // packs 8 * NA consecutive positions of ITS channel into bf16 hi / lo pairs and stores them to its TMEM lane:
        // column = buffer + (hi | lo) * 16 + position / 2.
        float* scr = reinterpret_cast<float*>(sa);
#pragma unroll
        for (int h = 0; h < NA; ++h)
#pragma unroll
          for (int e = 0; e < 8; ++e) scr[((cogA + h) * 8 + e) * 36 + tA] = a[h][e];
        asm volatile("bar.sync 1, %0;" ::"r"(kProd) : "memory");          // producer warps only
        const uint32_t tcol = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acol0 + (uint32_t)((c & 1) * 32) +
                              (uint32_t)(partT * 4 * NA);
        const float4* row = reinterpret_cast<const float4*>(scr + rowT * 36 + partT * 8 * NA);
#pragma unroll
        for (int h = 0; h < NA; ++h) {
          const float4 p0 = row[2 * h], p1 = row[2 * h + 1];
          const float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            const float2 hf = __bfloat1622float2(h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          tmem_st4(tcol + (uint32_t)(h * 4), hi[0], hi[1], hi[2], hi[3]);
          tmem_st4(tcol + 16u + (uint32_t)(h * 4), lo[0], lo[1], lo[2], lo[3]);
        }
        tmem_st_wait();
        tc_fence_before();
      } else {
#pragma unroll
        for (int h = 0; h < NA; ++h) {
          unsigned char* d0 = sa + (cogA + h) * sbo_a + tA * 16;
          put16(d0, d0 + plane_a, a[h]);
        }
      }
#pragma unroll
      for (int j = 0; j < MAXI; ++j)
        if (sx[j] >= 0) put16(sb + sx[j], sb + sx[j] + plane_b, v[j]);
      fence_proxy_async();
      mbar_arrive(&full[st]);
    };
    if (PW == 16) {                       // one CTA per SM: registers to spare for the second set
      float a0[NA][8], v0[MAXI][8], a1[NA][8], v1[MAXI][8];
      issue(0, a0, v0);
      for (int c = 0; c < nstage; c += 2) {
        if (c + 1 < nstage) issue(c + 1, a1, v1);
        commit(c, a0, v0);
        if (c + 1 < nstage) {
          if (c + 2 < nstage) issue(c + 2, a0, v0);
          commit(c + 1, a1, v1);
        }
      }
    } else {                              // two CTAs per SM (80 registers): loads and stores of a stage back to back
      float a0[NA][8], v0[MAXI][8];
      for (int c = 0; c < nstage; ++c) {
        issue(c, a0, v0);
        commit(c, a0, v0);
      }
    }
    // ===================== epilogue: TMEM -> fp32 reductions into dW[co][ci][k] =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;                     // TMEM lane quarter; PW / 4 warps share it
    const int co = co0 + q * 32 + lane;
    const bool ev = co < G.Cout_g;
    float* dst = G.Y + ((long long)(grp * G.Cout_g + co) * G.Cin_g + ci0) * G.K + tap0;
    const int nblk = ntap * NCI / 16;
    for (int blk = part; blk < nblk; blk += PW / 4) {
      float acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 16), acc);
      const int tl = (blk * 16) / NCI, n0 = (blk * 16) % NCI;
      if (!ev) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < cols_b) atomicAdd(dst + (long long)(n0 + j) * G.K + tl, acc[j]);
    }
  } else {
    // ===================== MMA issuers: issuer `iw` takes taps iw, iw + NI, ... =====================
    if (lane == 0) {
      const int iw = warp - PW;
      const uint32_t idesc = make_idesc_bf16(NCI, /*a_mn=*/P.ts == 0, /*b_mn=*/true);
      int st = 0;
      uint32_t par = 0;
      for (int c = 0; c < nstage; ++c) {
        mbar_wait(&full[st], par);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + (size_t)st * stage_sz), a_lo = a_hi + plane_a;
        const uint32_t b_base = a_hi + a_stage;
        for (int tl = iw; tl < ntap; tl += NI) {
          const uint32_t b_hi = b_base + (uint32_t)tapoff[tl], b_lo = b_hi + plane_b;
          const uint32_t d = tmem_base + (uint32_t)(tl * NCI);
#pragma unroll
          for (int ks = 0; ks < kWsTC / 16; ++ks) {
            const uint64_t db_hi = make_desc(b_hi + ks * 256, 128, sbo_b), db_lo = make_desc(b_lo + ks * 256, 128, sbo_b);
            if (P.ts) {
              const uint32_t ta_hi = tmem_base + acol0 + (uint32_t)((c & 1) * 32 + ks * 8), ta_lo = ta_hi + 16u;
              mma_bf16_ts(d, ta_hi, db_hi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
              mma_bf16_ts(d, ta_hi, db_lo, idesc, 1);
              mma_bf16_ts(d, ta_lo, db_hi, idesc, 1);
            } else {
              const uint64_t da_hi = make_desc(a_hi + ks * 256, 128, sbo_a), da_lo = make_desc(a_lo + ks * 256, 128, sbo_a);
              mma_bf16_ss(d, da_hi, db_hi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
              mma_bf16_ss(d, da_hi, db_lo, idesc, 1);
              mma_bf16_ss(d, da_lo, db_hi, idesc, 1);
            }
          }
        }
        mma_commit(&empty[st]);           // (an issuer without taps arrives at once: nothing of its own is in flight)
        if (++st == S) { st = 0; par ^= 1u; }
      }
      mma_commit(acc_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

static size_t ws_smem_bytes(const TcWS& P) {
  return (size_t)P.stages * (ws_a_stage(P.ts) + ws_b_stage(P.g, P.NCI, P.UB)) + (2 * P.stages + 1) * sizeof(uint64_t) + 16 +
         (size_t)P.TG * sizeof(int) + 16;
}

// Returns false when the geometry is left to the gather kernel.
static bool plan_wslab(TcWS& P, const vbx_conv_desc* d) {
  static const bool off = getenv("VBX_TC_WSLAB") && atoi(getenv("VBX_TC_WSLAB")) == 0;
  static const int min_k = getenv("VBX_TC_WSLAB_MIN_K") ? atoi(getenv("VBX_TC_WSLAB_MIN_K")) : 2;
  static const int min_cin = getenv("VBX_TC_WSLAB_MIN_CIN") ? atoi(getenv("VBX_TC_WSLAB_MIN_CIN")) : 64;   // measured: narrower reductions are faster on the gather kernel
  fill(P.g, d);
  const GemmP& G = P.g;
  if (off || G.K < min_k || G.K > 128 || G.Cin_g < min_cin || G.stride > 8) return false;
  const int R = G.Tout - 1 + ((G.K - 1) * G.dil + 1 + G.stride - 1) / G.stride;
  if ((long long)G.B * R * G.stride + 4096 >= (1ll << 31)) return false;
  P.R = R;
  P.NCI = G.Cin_g >= 64 ? 64 : (G.Cin_g + 15) / 16 * 16;
  P.ci_tiles = (G.Cin_g + P.NCI - 1) / P.NCI;
  P.co_tiles = (G.Cout_g + kRows - 1) / kRows;
  // ts mode (VBX_WS_TS=1, off by default): dy - the A operand, shared by every tap - staged in TENSOR memory
  // (tcgen05.st after a shared-memory transpose, tcgen05.mma with A from TMEM): an N = 64 SS MMA reads 4 KB of A + 2 KB
  // of B from shared memory per 32 cycles of tensor work, A from TMEM leaves 2 KB.  Correct (same tests), but measured
  // SLOWER on the B200: MelGAN stage 4 0.97 ms against 0.76 ms - the kernel is bound by the ~26 thread instructions it
  // spends per staged element (ncu: issue slots 43 % busy, tensor pipe 21 %), not by operand bandwidth.
  static const bool ts_on = getenv("VBX_WS_TS") && atoi(getenv("VBX_WS_TS")) == 1;
  P.ts = ts_on && G.K * P.NCI > 256 ? 1 : 0;
  const int cols = G.K * P.NCI <= 256 ? 256 : 512;
  int tgmax = (cols - (P.ts ? kWsTsCols : 0)) / P.NCI;
  P.ntg = (G.K + tgmax - 1) / tgmax;
  P.TG = (G.K + P.ntg - 1) / P.ntg;
  P.tmem_cols = pow2_cols(P.TG * P.NCI + (P.ts ? kWsTsCols : 0));
  P.UB = ws_UB(G, P.TG);
  if (ws_npos_taps(G, P.TG) * (P.NCI / 8) > kWsMaxItems * (P.tmem_cols > 256 ? 16 : 8) * 32) return false;
  const long long total = (long long)G.B * R;
  const long long tiles = (long long)P.ntg * P.ci_tiles * P.co_tiles * G.groups;
  long long max_split = (total + kWsTC * 8 - 1) / (kWsTC * 8);
  if (max_split > 65535) max_split = 65535;
  if (max_split < 1) max_split = 1;
  // (measured: with few tiles, ~3 short slices per SM beat one long one; with many, whole waves win)
  long long want = tiles >= 48 ? pick_split(tiles, 148 * (P.tmem_cols > 256 ? 1 : 2), max_split)
                               : (148 * 3 + tiles - 1) / tiles;
  if (want > max_split) want = max_split;
  if (deterministic_flag()) want = 1;
  long long per = (total + want - 1) / want;
  per = (per + kWsTC - 1) / kWsTC * kWsTC;
  P.rows_per = (int)per;
  const int stage = ws_a_stage(P.ts) + ws_b_stage(G, P.NCI, P.UB);
  int stg = (P.tmem_cols <= 256 ? 100 * 1024 : 200 * 1024) / stage;
  if (stg > 4) stg = 4;
  if (stg < 2) {
    stg = 200 * 1024 / stage;
    if (stg < 2) return false;
    if (stg > 2) stg = 2;
  }
  P.stages = P.ts ? 2 : stg;           // (ts: the two A buffers in TMEM pace the ring)
  return ws_smem_bytes(P) <= (size_t)kSmemMax;
}

static int launch_wslab(const TcWS& P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(tc_wslab_kernel<16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(tc_wslab_kernel<16, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(tc_wslab_kernel<16, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(tc_wslab_kernel<8, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(tc_wslab_kernel<8, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (ce != cudaSuccess) return fail((int)ce, "tc_wslab: cannot raise the dynamic shared memory limit");
    attr_set = true;
  }
  const long long total = (long long)P.g.B * P.R;
  dim3 grid((unsigned)(P.ntg * P.ci_tiles), (unsigned)(P.co_tiles * P.g.groups),
            (unsigned)((total + P.rows_per - 1) / P.rows_per));
  if (grid.y > 65535 || grid.z > 65535) return fail(VBX_UNSUPPORTED, "tc_wslab: grid too large");
  const int prod = (P.tmem_cols > 256 ? 16 : 8) * 32;
  const int need = (ws_npos_taps(P.g, P.TG) * (P.NCI / 8) + prod - 1) / prod;        // x items per producer thread
  if (P.tmem_cols > 256) {
    if (need <= 1) tc_wslab_kernel<16, 1, 1><<<grid, 16 * 32 + kWsIssuers * 32, ws_smem_bytes(P), st>>>(P);
    else if (need <= 3) tc_wslab_kernel<16, 1, 3><<<grid, 16 * 32 + kWsIssuers * 32, ws_smem_bytes(P), st>>>(P);
    else tc_wslab_kernel<16, 1, 4><<<grid, 16 * 32 + kWsIssuers * 32, ws_smem_bytes(P), st>>>(P);
  } else {
    if (need <= 2) tc_wslab_kernel<8, 2, 2><<<grid, 8 * 32 + kWsIssuers * 32, ws_smem_bytes(P), st>>>(P);
    else tc_wslab_kernel<8, 2, 4><<<grid, 8 * 32 + kWsIssuers * 32, ws_smem_bytes(P), st>>>(P);
  }
  return launched("tc_wslab_kernel");
}
