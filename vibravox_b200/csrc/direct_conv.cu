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

bool direct_fwd_ok(const GemmP& P) {
  return P.Cin_g == 1 && P.Cout_g <= 16 && P.K <= 128 && P.stride <= 2 &&
         (long long)P.B * P.groups * ((P.Tout + 1023) / 1024) < (1ll << 31);
}
bool direct_dgrad_ok(const GemmP& P) {
  return P.Cin_g == 1 && P.Cout_g <= 16 && P.K <= 128 && P.stride == 1 && P.refl <= P.pad &&
         (size_t)P.Cout_g * (512 + (P.K - 1) * P.dil) * 4 + (size_t)P.Cout_g * P.K * 4 <= 48 * 1024;
}

int direct_fwd(const GemmP& P, cudaStream_t st) {
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
  const int PT = 2, TT = PT * kDirThreads;
  const int tiles = (P.Tin + TT - 1) / TT;
  const int win = TT + (P.K - 1) * P.dil;
  const size_t smem = ((size_t)P.Cout_g * win + (size_t)P.Cout_g * P.K) * sizeof(float);
  const unsigned grid = (unsigned)((long long)P.B * P.groups * tiles);
  direct_dgrad_kernel<PT><<<grid, kDirThreads, smem, st>>>(P, tiles);
  return launched("direct_dgrad_kernel");
}

}  // namespace vbx
