// Direct (non-GEMM) kernels for convolutions with ONE input channel per group: the A-weighting FIR of the
// MR-STFT loss (1 -> 1, k = 101), the first layer of the MelGAN discriminator (1 -> 16, k = 15) and of the
// PQMF-band discriminators (4 -> 24, groups 4, k = 3).  Their reduction is only K long, so an implicit GEMM
// spends all its time building tiles; here a block stages one input window in shared memory, every thread
// keeps a few positions x all output channels of the group in registers, and the kernel runs at the rate the
// output can be written.  Replaces F.conv1d (and autograd's input gradient) at eben_discriminator.py:66-90,
// melgan_discriminator.py:89-100 and auraloss's FIRFilter for these shapes; reached through
// vbx_conv1d_fwd / vbx_conv1d_dgrad (include/vbx.h), which pick it when the geometry qualifies.
#include "common.cuh"
#include "conv_plan.h"

namespace vbx {

static const int kDirThreads = 256;

__device__ __forceinline__ float dir_finish(const GemmP& P, float v, int ch, long long idx) {
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

// y[b, g*Cg + j, t] = sum_k w[g*Cg + j, 0, k] * x[b, g, map(t*s + k*d - pad)]
// block = (b, g, tile of PT*256 output positions); CG = register rows (>= Cout_g)
template <int CG, int PT>
__global__ void __launch_bounds__(kDirThreads) direct_fwd_kernel(const GemmP P, int tiles) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % tiles, bg = blockIdx.x / tiles;
  const int b = bg / P.groups, g = bg % P.groups;
  const int Cg = P.Cout_g, K = P.K, s = P.stride, d = P.dil;
  const int TT = PT * kDirThreads;
  const int t0 = tile * TT;
  const int win = (TT - 1) * s + (K - 1) * d + 1;
  float* xs = sm;
  float* ws = sm + win;
  const float* xrow = P.X + ((long long)b * P.Cin + g) * P.Tin;
  for (int i = tid; i < win; i += kDirThreads) {
    const int q = map_pos(t0 * s - P.pad + i, P.Tin, P.refl);
    xs[i] = q >= 0 ? xrow[q] : 0.f;
  }
  for (int i = tid; i < Cg * K; i += kDirThreads) ws[i] = P.W[(long long)g * Cg * K + i];
  __syncthreads();
  float acc[CG][PT];
#pragma unroll
  for (int j = 0; j < CG; ++j)
#pragma unroll
    for (int i = 0; i < PT; ++i) acc[j][i] = 0.f;
  for (int k = 0; k < K; ++k) {
    float xv[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) xv[i] = xs[(tid + kDirThreads * i) * s + k * d];
#pragma unroll
    for (int j = 0; j < CG; ++j) {
      if (j < Cg) {                                      // uniform
        const float w = ws[j * K + k];
#pragma unroll
        for (int i = 0; i < PT; ++i) acc[j][i] = fmaf(w, xv[i], acc[j][i]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    if (j < Cg) {
      const int ch = g * Cg + j;
#pragma unroll
      for (int i = 0; i < PT; ++i) {
        const int t = t0 + tid + kDirThreads * i;
        if (t < P.Tout) {
          const long long idx = ((long long)b * P.Cout + ch) * P.Tout + t;
          P.Y[idx] = dir_finish(P, acc[j][i], ch, idx);
        }
      }
    }
  }
}

// Input gradient, stride 1:  dxe[p] = sum_j sum_k w[g*Cg + j, 0, k] * dy[b, g*Cg + j, p + pad - k*d]  on the
// padded domain p in [-refl, Tin + refl), folded back through the mirror:  dx[u] = dxe[u] + dxe[-u] (1 <= u <= refl)
// + dxe[2(Tin-1) - u] (Tin-1-refl <= u <= Tin-2).  The mirror terms touch <= 2*refl positions per row and are
// summed straight from global memory.
__device__ float dir_dxe_global(const GemmP& P, const float* dyg, int p) {
  float v = 0.f;
  for (int j = 0; j < P.Cout_g; ++j)
    for (int k = 0; k < P.K; ++k) {
      const int t = p + P.pad - k * P.dil;
      if (t >= 0 && t < P.Tout) v = fmaf(P.W[j * P.K + k], dyg[(long long)j * P.Tout + t], v);
    }
  return v;
}

template <int PT>
__global__ void __launch_bounds__(kDirThreads) direct_dgrad_kernel(const GemmP Pin, int tiles) {
  extern __shared__ float sm[];
  GemmP P = Pin;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % tiles, bg = blockIdx.x / tiles;
  const int b = bg / P.groups, g = bg % P.groups;
  const int Cg = P.Cout_g, K = P.K, d = P.dil;
  const int TT = PT * kDirThreads;
  const int u0 = tile * TT;
  const int halo = (K - 1) * d;
  const int win = TT + halo;
  const int tlo = u0 + P.pad - halo;                     // dy position held at column 0 of the window
  float* dys = sm;                                       // [Cg][win]
  float* ws = sm + (size_t)Cg * win;
  const float* dyg = P.X + ((long long)b * P.Cout + (long long)g * Cg) * P.Tout;
  P.W += (long long)g * Cg * K;
  for (int j = 0; j < Cg; ++j)
    for (int i = tid; i < win; i += kDirThreads) {
      const int t = tlo + i;
      dys[j * win + i] = (t >= 0 && t < P.Tout) ? dyg[(long long)j * P.Tout + t] : 0.f;
    }
  for (int i = tid; i < Cg * K; i += kDirThreads) ws[i] = P.W[i];
  __syncthreads();
  float acc[PT];
#pragma unroll
  for (int i = 0; i < PT; ++i) acc[i] = 0.f;
  for (int j = 0; j < Cg; ++j) {
    const float* row = dys + j * win + tid;
    for (int k = 0; k < K; ++k) {
      const float w = ws[j * K + k];
      const int o = (K - 1 - k) * d;
#pragma unroll
      for (int i = 0; i < PT; ++i) acc[i] = fmaf(w, row[o + kDirThreads * i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < PT; ++i) {
    const int u = u0 + tid + kDirThreads * i;
    if (u < P.Tin) {
      float v = acc[i];
      if (P.refl > 0) {
        if (u >= 1 && u <= P.refl) v += dir_dxe_global(P, dyg, -u);
        if (u <= P.Tin - 2 && u >= P.Tin - 1 - P.refl) v += dir_dxe_global(P, dyg, 2 * (P.Tin - 1) - u);
      }
      const long long idx = ((long long)b * P.Cin + g) * P.Tin + u;
      P.Y[idx] = dir_finish(P, v, g, idx);
    }
  }
}

// Weight gradient of a conv with a handful of outputs per input row - the certainty convs (C -> 1, k3), the generator's
// last conv (32 -> 4, k3), the one-input-channel-per-group first stages (4 -> 24 g4 k3, 1 -> 16 k15) - as what it is: a
// streaming dot product.  dW[co, ci_g, k] += sum_{b,t} dy[b, co, t] * x[b, ci, map(t + k*d - pad)]  (stride 1).
// One block per (block of NCO output channels, input channel of the group) and batch slice: every thread walks quads of
// 4 consecutive t, keeps NCO x K partial sums in registers, and the block reduces them in a fixed order; one writer per
// output when the batch is not sliced (deterministic mode), fp32 atomics across slices otherwise.  x and dy are read
// once from HBM (the K shifted reads of x hit L1); the implicit-GEMM kernels spent a 128-row tile on 1-24 useful rows
// here (7-20x their HBM roofline).
template <int NCO, int K, bool WINDOW>
__global__ void __launch_bounds__(kDirThreads) skinny_wgrad_kernel(const GemmP P, int bper) {
  constexpr int R = 4;
  __shared__ float red[kDirThreads / 32][NCO * K];
  const int cob = blockIdx.x / P.Cin_g, cig = blockIdx.x % P.Cin_g;
  const int co0 = cob * NCO;
  const int ci = (co0 / P.Cout_g) * P.Cin_g + cig;
  const int b0 = blockIdx.y * bper, b1 = min(b0 + bper, P.B);
  const int d = P.dil, pad = P.pad, Tin = P.Tin, Tout = P.Tout;
  float acc[NCO][K];
#pragma unroll
  for (int i = 0; i < NCO; ++i)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[i][k] = 0.f;
  const int nq = (Tout + R - 1) / R;
  {
    // (batch item, quad) pairs as ONE index space: short rows (T = 187 ... 375 on the certainty convs) would otherwise
    // leave most of the block idle in every per-item pass
    for (int idx = threadIdx.x; idx < (b1 - b0) * nq; idx += kDirThreads) {
      const int b = b0 + idx / nq, q = idx % nq;
      const float* __restrict__ xr = P.X + ((long long)b * P.Cin + ci) * Tin;
      const float* __restrict__ dyr = P.DY + ((long long)b * P.Cout + co0) * Tout;
      const int t0 = q * R;
      float dyv[NCO][R];
#pragma unroll
      for (int i = 0; i < NCO; ++i)
#pragma unroll
        for (int r = 0; r < R; ++r) dyv[i][r] = t0 + r < Tout ? dyr[(long long)i * Tout + t0 + r] : 0.f;
      const int p0 = t0 - pad;
      const bool interior = p0 >= 0 && p0 + (R - 1) + (K - 1) * d < Tin;
      if (WINDOW) {                                       // d == 1: one sliding window of R + K - 1 values
        float w[R + K - 1];
#pragma unroll
        for (int j = 0; j < R + K - 1; ++j) {
          if (interior) w[j] = xr[p0 + j];
          else { const int p = map_pos(p0 + j, Tin, P.refl); w[j] = p >= 0 ? xr[p] : 0.f; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < NCO; ++i) acc[i][k] = fmaf(dyv[i][r], w[k + r], acc[i][k]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float xv;
            if (interior) xv = xr[p0 + r + k * d];
            else { const int p = map_pos(p0 + r + k * d, Tin, P.refl); xv = p >= 0 ? xr[p] : 0.f; }
#pragma unroll
            for (int i = 0; i < NCO; ++i) acc[i][k] = fmaf(dyv[i][r], xv, acc[i][k]);
          }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NCO; ++i)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float v = acc[i][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][i * K + k] = v;
    }
  __syncthreads();
  if (threadIdx.x < NCO * K) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDirThreads / 32; ++w) v += red[w][threadIdx.x];
    const int i = threadIdx.x / K, k = threadIdx.x % K;
    float* out = P.Y + ((long long)(co0 + i) * P.Cin_g + cig) * K + k;
    if (gridDim.y == 1) *out += v;
    else atomicAdd(out, v);
  }
}

bool skinny_wgrad_ok(const GemmP& P) {
  static const bool off = getenv("VBX_SKINNY_WGRAD") && atoi(getenv("VBX_SKINNY_WGRAD")) == 0;
  if (off || P.stride != 1 || P.refl > P.pad) return false;
  if (P.Cin_g == 1) return P.K == 3 || (P.K == 15 && P.dil == 1);
  return P.groups == 1 && P.K == 3 && (P.Cout == 1 || P.Cout == 4);
}

int skinny_wgrad(const GemmP& P, cudaStream_t st) {
  const int nco = P.Cin_g == 1 ? 1 : P.Cout;
  const long long gx = (long long)(P.Cout / nco) * P.Cin_g;
  if (gx > 0x7fffffffLL) return fail(VBX_UNSUPPORTED, "skinny_wgrad: grid too large");
  long long split = (148 * 8 + gx - 1) / gx;              // ~8 blocks per SM in total
  if (split > P.B) split = P.B;
  if (split > 65535) split = 65535;
  if (split < 1 || deterministic_flag()) split = 1;
  const int bper = (int)((P.B + split - 1) / split);
  dim3 grid((unsigned)gx, (unsigned)((P.B + bper - 1) / bper));
  if (P.K == 15) skinny_wgrad_kernel<1, 15, true><<<grid, kDirThreads, 0, st>>>(P, bper);
  else if (nco == 1 && P.dil == 1) skinny_wgrad_kernel<1, 3, true><<<grid, kDirThreads, 0, st>>>(P, bper);
  else if (nco == 1) skinny_wgrad_kernel<1, 3, false><<<grid, kDirThreads, 0, st>>>(P, bper);
  else if (P.dil == 1) skinny_wgrad_kernel<4, 3, true><<<grid, kDirThreads, 0, st>>>(P, bper);
  else skinny_wgrad_kernel<4, 3, false><<<grid, kDirThreads, 0, st>>>(P, bper);
  return launched("skinny_wgrad_kernel");
}

// ---- quad kernels: 4 consecutive positions per thread, a register window over the taps (stride 1, dilation 1) --------
// MelGAN stage 0 (1 -> 16, k 15, reflect 7): the generic direct kernels above index their shared-memory window and
// weights with run-time K / Cg and executed 7.8 instructions per FMA (ncu: 89.7 M warp instructions for 11.5 M warp
// FMAs, 168 us for a 98 MB output).  Here K and the channel count are compile-time, the 18-value input window of a quad is
// loaded once for all 16 output channels, the weights come as 16-byte shared-memory broadcasts, and a quad's outputs are
// stored as one float4 per channel.
template <int CG, int K>
__global__ void __launch_bounds__(kDirThreads) direct_fwd4_kernel(const GemmP P, long long nquads, int nq) {
  constexpr int R = 4, KP = (K + 3) / 4 * 4;
  __shared__ __align__(16) float ws[CG * KP];
  const int Cg = P.Cout_g;
  const long long idx = (long long)blockIdx.x * kDirThreads + threadIdx.x;
  // (a block may straddle groups: the weights are staged per thread-group below only when groups == 1)
  for (int i = threadIdx.x; i < CG * KP; i += kDirThreads) {
    const int j = i / KP, k = i % KP;
    ws[i] = (j < Cg && k < K) ? P.W[j * K + k] : 0.f;
  }
  __syncthreads();
  if (idx >= nquads) return;
  const int b = (int)(idx / nq), q = (int)(idx % nq);
  const int t0 = q * R;
  const float* __restrict__ xr = P.X + (long long)b * P.Tin;
  float w[R + K - 1];
  const int p0 = t0 - P.pad;
  if (p0 >= 0 && p0 + R + K - 2 < P.Tin) {
#pragma unroll
    for (int j = 0; j < R + K - 1; ++j) w[j] = xr[p0 + j];
  } else {
#pragma unroll
    for (int j = 0; j < R + K - 1; ++j) {
      const int p = map_pos(p0 + j, P.Tin, P.refl);
      w[j] = p >= 0 ? xr[p] : 0.f;
    }
  }
  const bool vec = (P.Tout & 3) == 0 && t0 + R <= P.Tout && !P.mask && !P.res && P.beta == 0.f && !P.gate &&
                   (reinterpret_cast<uintptr_t>(P.Y) & 15) == 0;
  // (the channel loop stays ROLLED: fully unrolled, the 16 x 15 x 4 FMA body runs once per warp and the kernel is bound by
  //  instruction fetch - ncu: top stall no_instruction at 7.9 cycles per issue)
#pragma unroll 1
  for (int j = 0; j < CG; ++j) {
    if (j < Cg) {
      float wt[KP];
#pragma unroll
      for (int k4 = 0; k4 < KP; k4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(ws + j * KP + k4);
        wt[k4] = v.x; wt[k4 + 1] = v.y; wt[k4 + 2] = v.z; wt[k4 + 3] = v.w;
      }
      float acc[R] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(wt[k], w[k + r], acc[r]);
      const long long o = ((long long)b * P.Cout + j) * P.Tout + t0;
      if (vec) {                                         // bias + LeakyReLU only: nothing per element but the select
        const float bv = P.bias ? __ldg(P.bias + j) : 0.f, sl = P.slope;
        float4 v;
        v.x = acc[0] + bv; v.y = acc[1] + bv; v.z = acc[2] + bv; v.w = acc[3] + bv;
        v.x = v.x > 0.f ? v.x : v.x * sl; v.y = v.y > 0.f ? v.y : v.y * sl;
        v.z = v.z > 0.f ? v.z : v.z * sl; v.w = v.w > 0.f ? v.w : v.w * sl;
        *reinterpret_cast<float4*>(P.Y + o) = v;
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (t0 + r < P.Tout) P.Y[o + r] = dir_finish(P, acc[r], j, o + r);
      }
    }
  }
}

// Input gradient of the same layer:  dx[u] = sum_j sum_k w[j, k] * dy[j, u + pad - k]  (+ the mirror terms of the reflect
// halo on the <= 2*refl edge positions, as in direct_dgrad_kernel).  Per channel the 18-value window of a quad is read as
// six ALIGNED float4s starting at (u0 + pad - (K-1)) rounded down to a multiple of 4; OFF is that rounding, a
// compile-time constant per geometry, so the window stays in registers.
template <int CG, int K, int OFF>
__global__ void __launch_bounds__(kDirThreads) direct_dgrad4_kernel(const GemmP P, long long nquads, int nq) {
  constexpr int R = 4, KP = (K + 3) / 4 * 4, NW = (OFF + R + K - 1 + 3) / 4 * 4;
  __shared__ __align__(16) float ws[CG * KP];
  const int Cg = P.Cout_g;
  for (int i = threadIdx.x; i < CG * KP; i += kDirThreads) {
    const int j = i / KP, k = i % KP;
    ws[i] = (j < Cg && k < K) ? P.W[j * K + (K - 1 - k)] : 0.f;      // flipped: window index = k' + r
  }
  __syncthreads();
  const long long idx = (long long)blockIdx.x * kDirThreads + threadIdx.x;
  const bool live = idx < nquads;
  const int b = live ? (int)(idx / nq) : 0, q = live ? (int)(idx % nq) : 0;
  const int u0 = q * R;
  const int a0 = u0 + P.pad - (K - 1) - OFF;                       // multiple of 4 by construction
  const float* __restrict__ dyb = P.X + (long long)b * P.Cout * P.Tout;
  const bool interior = a0 >= 0 && a0 + NW <= P.Tout;
  float acc[R] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int j = 0; j < CG; ++j) {
    if (live && j < Cg) {
      const float* __restrict__ row = dyb + (long long)j * P.Tout;
      float w[NW];
      if (interior) {
#pragma unroll
        for (int i = 0; i < NW; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(row + a0 + i);
          w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const int t = a0 + i;
          w[i] = (t >= 0 && t < P.Tout) ? row[t] : 0.f;
        }
      }
      float wt[KP];
#pragma unroll
      for (int k4 = 0; k4 < KP; k4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(ws + j * KP + k4);
        wt[k4] = v.x; wt[k4 + 1] = v.y; wt[k4 + 2] = v.z; wt[k4 + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(wt[k], w[OFF + k + r], acc[r]);
    }
  }
  if (P.refl > 0) {
    // mirror terms of the <= 2*refl edge positions per item: the WARP sums the Cg*K products of one such position
    // together (a single thread walking them is a 240-long chain of dependent cache misses that set the kernel's duration)
    const int lane = threadIdx.x & 31;
    const bool edge = live && (u0 <= P.refl || u0 + R - 1 >= P.Tin - 1 - P.refl);
    unsigned need = __ballot_sync(0xffffffffu, edge);
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      const int sb = __shfl_sync(0xffffffffu, b, src), su0 = __shfl_sync(0xffffffffu, u0, src);
      const float* __restrict__ sdy = P.X + (long long)sb * P.Cout * P.Tout;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int u = su0 + r;
        if (u >= P.Tin) continue;
        for (int side = 0; side < 2; ++side) {
          const bool hit = side == 0 ? (u >= 1 && u <= P.refl) : (u <= P.Tin - 2 && u >= P.Tin - 1 - P.refl);
          if (!hit) continue;                                       // (warp-uniform)
          const int p = side == 0 ? -u : 2 * (P.Tin - 1) - u;
          float v = 0.f;
          for (int i = lane; i < Cg * K; i += 32) {
            const int j = i / K, k = i % K;
            const int t = p + P.pad - k * P.dil;
            if (t >= 0 && t < P.Tout) v = fmaf(P.W[j * K + k], sdy[(long long)j * P.Tout + t], v);
          }
#pragma unroll
          for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
          if (lane == src) acc[r] += v;
        }
      }
    }
  }
  if (!live) return;
  const long long o = (long long)b * P.Tin + u0;                    // Cin == groups == 1
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (u0 + r < P.Tin) P.Y[o + r] = dir_finish(P, acc[r], 0, o + r);
}

static bool quad_geometry(const GemmP& P) {
  static const bool off = getenv("VBX_DIRECT_QUAD") && atoi(getenv("VBX_DIRECT_QUAD")) == 0;
  return !off && P.groups == 1 && P.Cin == 1 && P.Cout == 16 && P.K == 15 && P.stride == 1 && P.dil == 1 &&
         P.refl <= P.pad && (long long)P.B * ((P.Tout + 3) / 4) < (1ll << 40);
}

// ---- a handful of OUTPUT channels over many input channels (certainty convs C -> 1 k3, generator last conv 32 -> 4 k3):
// y[b, co, t] = sum_ci sum_k w[co, ci, k] * x[b, ci, map(t + k*d - pad)], one thread per position and input-channel slice
// (coalesced loads: consecutive lanes = consecutive t), the slices of a block summed through shared memory in a fixed
// order.  The implicit-GEMM kernels gave these layers a 128-row x 16-column tile with ONE useful column (10-19x their HBM
// roofline).
template <int NCO>
__global__ void __launch_bounds__(kDirThreads) skinny_fwd_kernel(const GemmP P, int nslice, int tiles) {
  __shared__ float red[kDirThreads * NCO];
  const int ppb = kDirThreads / nslice;
  const int pos = threadIdx.x % ppb, slice = threadIdx.x / ppb;
  const int b = blockIdx.x / tiles, t = (blockIdx.x % tiles) * ppb + pos;
  const int cper = (P.Cin + nslice - 1) / nslice;
  const int c0 = slice * cper, c1 = min(c0 + cper, P.Cin);
  float acc[NCO];
#pragma unroll
  for (int i = 0; i < NCO; ++i) acc[i] = 0.f;
  if (t < P.Tout) {
    int p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = map_pos(t + k * P.dil - P.pad, P.Tin, P.refl);
    const float* __restrict__ xb = P.X + (long long)b * P.Cin * P.Tin;
#pragma unroll 8
    for (int c = c0; c < c1; ++c) {
      const float* __restrict__ xr = xb + (long long)c * P.Tin;
      float xv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) xv[k] = p[k] >= 0 ? xr[p[k]] : 0.f;
#pragma unroll
      for (int i = 0; i < NCO; ++i) {
        const float* __restrict__ wr = P.W + ((long long)i * P.Cin + c) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[i] = fmaf(__ldg(wr + k), xv[k], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCO; ++i) red[(slice * NCO + i) * ppb + pos] = acc[i];
  __syncthreads();
  if (slice == 0 && t < P.Tout) {
#pragma unroll
    for (int i = 0; i < NCO; ++i) {
      float v = 0.f;
      for (int s2 = 0; s2 < nslice; ++s2) v += red[(s2 * NCO + i) * ppb + pos];
      const long long o = ((long long)b * P.Cout + i) * P.Tout + t;
      P.Y[o] = dir_finish(P, v, i, o);
    }
  }
}

bool skinny_fwd_ok(const GemmP& P) {
  static const bool off = getenv("VBX_SKINNY_FWD") && atoi(getenv("VBX_SKINNY_FWD")) == 0;
  return !off && P.groups == 1 && P.K == 3 && P.stride == 1 && (P.Cout == 1 || P.Cout == 4) && P.Cin >= 8 &&
         P.refl <= P.pad;
}

int skinny_fwd(const GemmP& P, cudaStream_t st) {
  // slices of the input channels per block: enough blocks to fill the machine on the short certainty maps
  int nslice = 1;
  while (nslice < 16 && (long long)P.B * P.Tout * nslice < 148LL * 8 * kDirThreads && P.Cin / (2 * nslice) >= 8) nslice *= 2;
  const int ppb = kDirThreads / nslice;
  const int tiles = (P.Tout + ppb - 1) / ppb;
  const long long grid = (long long)P.B * tiles;
  if (grid > 0x7fffffffLL) return fail(VBX_UNSUPPORTED, "skinny_fwd: grid too large");
  if (P.Cout == 1) skinny_fwd_kernel<1><<<(unsigned)grid, kDirThreads, 0, st>>>(P, nslice, tiles);
  else skinny_fwd_kernel<4><<<(unsigned)grid, kDirThreads, 0, st>>>(P, nslice, tiles);
  return launched("skinny_fwd_kernel");
}

bool direct_fwd_ok(const GemmP& P) {
  return P.Cin_g == 1 && P.Cout_g <= 16 && P.K <= 128 && P.stride <= 2 &&
         (long long)P.B * P.groups * ((P.Tout + 1023) / 1024) < (1ll << 31);
}
bool direct_dgrad_ok(const GemmP& P) {
  return P.Cin_g == 1 && P.Cout_g <= 16 && P.K <= 128 && P.stride == 1 && P.refl <= P.pad &&
         (size_t)P.Cout_g * (512 + (P.K - 1) * P.dil) * 4 + (size_t)P.Cout_g * P.K * 4 <= 48 * 1024;
}

int direct_fwd(const GemmP& P, cudaStream_t st) {
  if (quad_geometry(P)) {
    const int nq = (P.Tout + 3) / 4;
    const long long nquads = (long long)P.B * nq;
    direct_fwd4_kernel<16, 15><<<(unsigned)((nquads + kDirThreads - 1) / kDirThreads), kDirThreads, 0, st>>>(P, nquads, nq);
    return launched("direct_fwd4_kernel");
  }
  const int PT = 4, TT = PT * kDirThreads;
  const int tiles = (P.Tout + TT - 1) / TT;
  const int win = (TT - 1) * P.stride + (P.K - 1) * P.dil + 1;
  const size_t smem = ((size_t)win + (size_t)P.Cout_g * P.K) * sizeof(float);
  if (smem > 48 * 1024) return fail(VBX_UNSUPPORTED, "direct_fwd: window too large");
  const unsigned grid = (unsigned)((long long)P.B * P.groups * tiles);
  if (P.Cout_g == 1) direct_fwd_kernel<1, PT><<<grid, kDirThreads, smem, st>>>(P, tiles);
  else if (P.Cout_g <= 8) direct_fwd_kernel<8, PT><<<grid, kDirThreads, smem, st>>>(P, tiles);
  else direct_fwd_kernel<16, PT><<<grid, kDirThreads, smem, st>>>(P, tiles);
  return launched("direct_fwd_kernel");
}

int direct_dgrad(const GemmP& P, cudaStream_t st) {
  if (quad_geometry(P) && (P.Tout & 3) == 0 && (reinterpret_cast<uintptr_t>(P.X) & 15) == 0) {
    const int nq = (P.Tin + 3) / 4;
    const long long nquads = (long long)P.B * nq;
    const unsigned grid = (unsigned)((nquads + kDirThreads - 1) / kDirThreads);
    switch (((P.pad - (P.K - 1)) % 4 + 4) % 4) {                    // distance of a window start above a multiple of 4
      case 0: direct_dgrad4_kernel<16, 15, 0><<<grid, kDirThreads, 0, st>>>(P, nquads, nq); break;
      case 1: direct_dgrad4_kernel<16, 15, 1><<<grid, kDirThreads, 0, st>>>(P, nquads, nq); break;
      case 2: direct_dgrad4_kernel<16, 15, 2><<<grid, kDirThreads, 0, st>>>(P, nquads, nq); break;
      default: direct_dgrad4_kernel<16, 15, 3><<<grid, kDirThreads, 0, st>>>(P, nquads, nq); break;
    }
    return launched("direct_dgrad4_kernel");
  }
  const int PT = 2, TT = PT * kDirThreads;
  const int tiles = (P.Tin + TT - 1) / TT;
  const int win = TT + (P.K - 1) * P.dil;
  const size_t smem = ((size_t)P.Cout_g * win + (size_t)P.Cout_g * P.K) * sizeof(float);
  const unsigned grid = (unsigned)((long long)P.B * P.groups * tiles);
  direct_dgrad_kernel<PT><<<grid, kDirThreads, smem, st>>>(P, tiles);
  return launched("direct_dgrad_kernel");
}

}  // namespace vbx
