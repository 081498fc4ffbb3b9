// Everything on the EBEN training-step path that is not a Conv1d tile: weight norm,
// the polyphase PQMF kernels, element-wise stages, loss reductions, balancing, Adam.
// All HBM-bound streaming kernels: grid-stride loops sized to a multiple of the 148
// SMs, float4 access where the layout allows, warp-shuffle + one atomic per block for
// reductions (double accumulators so the result does not depend on block order at
// fp32 resolution).  include/vbx.h cites the reference call site of each entry point.
#include "common.cuh"
#include "conv_plan.h"

namespace vbx {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
static int g_tc_mode = 0;

static const int kSMs = 148;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; result valid in thread 0.  blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}
static inline int stream_blocks(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  if (b > kSMs * 8) b = kSMs * 8;
  if (b < 1) b = 1;
  return (int)b;
}
// grid of a reduction that ends in one atomic per block: a single block in deterministic mode (fixed order)
static inline int reduce_blocks(long long n, int per_block) {
  return deterministic_flag() ? 1 : stream_blocks(n, per_block);
}

// ------------------------------------------------------------------ weight norm
__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                       float* __restrict__ w, float* __restrict__ wt,
                                       float* __restrict__ inv_norm, int Cin_g, int K, int Cout_g) {
  __shared__ double sh[32];
  __shared__ float s_scale;
  const int r = blockIdx.x, row = Cin_g * K;
  const float* vr = v + (long long)r * row;
  double ss = 0.0;
  for (int i = threadIdx.x; i < row; i += blockDim.x) { float a = vr[i]; ss += (double)a * a; }
  ss = block_sum(ss, sh);
  if (threadIdx.x == 0) {
    float nrm = sqrtf((float)ss);
    s_scale = g[r] / nrm;
    inv_norm[r] = 1.f / nrm;
  }
  __syncthreads();
  const float sc = s_scale;
  const int grp = r / Cout_g, col = r % Cout_g;
  for (int i = threadIdx.x; i < row; i += blockDim.x) {
    float o = vr[i] * sc;
    w[(long long)r * row + i] = o;
    if (wt) {
      int ci = i / K, k = i % K;
      wt[(((long long)grp * Cin_g + ci) * Cout_g + col) * K + k] = o;
    }
  }
}

__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                       const float* __restrict__ inv_norm, const float* __restrict__ dw,
                                       float* __restrict__ dg, float* __restrict__ dv, int row, float beta) {
  __shared__ double sh[32];
  __shared__ float s_dot;
  const int r = blockIdx.x;
  const float* vr = v + (long long)r * row;
  const float* dr = dw + (long long)r * row;
  double dot = 0.0;
  for (int i = threadIdx.x; i < row; i += blockDim.x) dot += (double)vr[i] * dr[i];
  dot = block_sum(dot, sh);
  const float inv = inv_norm[r], gr = g[r];
  if (threadIdx.x == 0) {
    s_dot = (float)dot;
    float o = (float)dot * inv;
    dg[r] = beta != 0.f ? beta * dg[r] + o : o;
  }
  __syncthreads();
  const float a = gr * inv, c = gr * s_dot * inv * inv * inv;
  float* o = dv + (long long)r * row;
  for (int i = threadIdx.x; i < row; i += blockDim.x) {
    float val = a * dr[i] - c * vr[i];
    o[i] = beta != 0.f ? beta * o[i] + val : val;
  }
}

// ------------------------------------------------------------------ PQMF
// analysis: 256 band-rate outputs per block; the input span is staged in shared
// memory in polyphase order (phase-major) so that thread t reads consecutive words.
// M, N: decimation / taps known at compile time (the reference's bank is 4 x 32: pqmf.py:26-64, eben_generator.py:101)
// so the phase arithmetic (k % m, k / m) folds into the unrolled tap loop; 0 = run-time values (any other bank).
template <bool PER_BAND, int M, int N>
__global__ void __launch_bounds__(256) pqmf_analysis_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     float* __restrict__ y, int L, int T, int m_rt, int n_rt, int bands) {
  extern __shared__ float sm[];
  const int m = M ? M : m_rt, n = N ? N : n_rt;
  const int TT = 256;
  const int QL = TT + (n + m - 1) / m + 1;          // words per phase
  float* ws = sm;                                   // [bands][n]
  float* xs = sm + bands * n;                       // [m][QL]   (PER_BAND: re-staged per band)
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  for (int i = threadIdx.x; i < bands * n; i += blockDim.x) ws[i] = w[i];
  const int span = m * TT + n;
  const long long p0 = (long long)m * t0 - (n - 1);
  const int t = t0 + threadIdx.x;
  const int xch = PER_BAND ? bands : 1;
  for (int c0 = 0; c0 < (PER_BAND ? bands : 1); ++c0) {
    __syncthreads();
    const float* xb = x + ((long long)b * xch + c0) * L;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
      long long p = p0 + i;
      float v = (p >= 0 && p < L) ? xb[p] : 0.f;
      xs[(i % m) * QL + i / m] = v;
    }
    __syncthreads();
    if (t < T) {
      for (int c = (PER_BAND ? c0 : 0); c < (PER_BAND ? c0 + 1 : bands); ++c) {
        float acc = 0.f;
        if (N) {
#pragma unroll
          for (int k = 0; k < (N ? N : 1); ++k) acc = fmaf(ws[c * n + k], xs[(k % m) * QL + threadIdx.x + k / m], acc);
        } else {
          for (int k = 0; k < n; ++k) acc = fmaf(ws[c * n + k], xs[(k % m) * QL + threadIdx.x + k / m], acc);
        }
        y[((long long)b * bands + c) * T + t] = acc;
      }
    }
  }
}

// synthesis: 1024 full-rate outputs per block (4 per thread, stride 256 for coalescing).
template <int M, int N>
__global__ void __launch_bounds__(256) pqmf_synthesis_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      float* __restrict__ y, int T, int L, int m_rt, int n_rt, int bands,
                                      int sum_bands) {
  extern __shared__ float sm[];
  const int m = M ? M : m_rt, n = N ? N : n_rt;
  const int UU = 1024;
  const int J = (n + m - 1) / m;                    // taps per phase (upper bound)
  const int QL = UU / m + J + 2;
  float* ws = sm;                                   // [bands][n]
  float* xs = sm + bands * n;                       // [bands][QL]
  const int b = blockIdx.y, u0 = blockIdx.x * UU;
  for (int i = threadIdx.x; i < bands * n; i += blockDim.x) ws[i] = w[i];
  const int tbase = (u0 + n - 1) / m - (J - 1) - 1;
  for (int i = threadIdx.x; i < bands * QL; i += blockDim.x) {
    int c = i / QL, q = i % QL;
    int t = tbase + q;
    xs[i] = (t >= 0 && t < T) ? x[((long long)b * bands + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = 0; r < UU / 256; ++r) {
    const int u = u0 + r * 256 + threadIdx.x;
    if (u >= L) break;
    const int a = u + n - 1, ph = a % m, tq = a / m - tbase;
    float tot = 0.f;
    for (int c = 0; c < bands; ++c) {
      float acc = 0.f;
      if (M && N) {
#pragma unroll
        for (int j = 0; j < (M && N ? N / (M ? M : 1) : 1); ++j) acc = fmaf(ws[c * n + ph + j * m], xs[c * QL + tq - j], acc);
      } else {
        int j = 0;
        for (int k = ph; k < n; k += m, ++j) acc = fmaf(ws[c * n + k], xs[c * QL + tq - j], acc);
      }
      if (sum_bands) tot += acc;
      else y[((long long)b * bands + c) * L + u] = acc;
    }
    if (sum_bands) y[(long long)b * L + u] = tot;
  }
}

// ------------------------------------------------------------------ element-wise
__global__ void leaky_relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float slope) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (long long i = i0; i < n4; i += stride) {
    float4 v = x4[i];
    v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
    v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
    y4[i] = v;
  }
  for (long long i = (n4 << 2) + i0; i < n; i += stride) { float v = x[i]; y[i] = v > 0.f ? v : v * slope; }
}

// dx = dy * (pre-activation > 0 ? 1 : slope) (+ beta * dx); the sign comes from the byte mask the forward stored or
// from `ref`.  VEC: everything 16-byte aligned (4 elements per thread and access).
template <bool VEC>
__global__ void __launch_bounds__(256) leaky_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ ref,
                                      const unsigned char* __restrict__ mask, float* __restrict__ dx,
                                      long long n, float slope, float beta) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (VEC) {
    const long long n4 = n >> 2;
    for (long long i = i0; i < n4; i += stride) {
      const float4 d = reinterpret_cast<const float4*>(dy)[i];
      bool p0, p1, p2, p3;
      if (mask) {
        const uchar4 m = reinterpret_cast<const uchar4*>(mask)[i];
        p0 = m.x != 0; p1 = m.y != 0; p2 = m.z != 0; p3 = m.w != 0;
      } else {
        const float4 r = reinterpret_cast<const float4*>(ref)[i];
        p0 = r.x > 0.f; p1 = r.y > 0.f; p2 = r.z > 0.f; p3 = r.w > 0.f;
      }
      float4 v = make_float4(d.x * (p0 ? 1.f : slope), d.y * (p1 ? 1.f : slope), d.z * (p2 ? 1.f : slope),
                             d.w * (p3 ? 1.f : slope));
      if (beta != 0.f) {
        const float4 o = reinterpret_cast<const float4*>(dx)[i];
        v.x += beta * o.x; v.y += beta * o.y; v.z += beta * o.z; v.w += beta * o.w;
      }
      reinterpret_cast<float4*>(dx)[i] = v;
    }
    return;
  }
  for (long long i = i0; i < n; i += stride) {
    bool pos = mask ? mask[i] != 0 : ref[i] > 0.f;
    float v = dy[i] * (pos ? 1.f : slope);
    dx[i] = beta != 0.f ? beta * dx[i] + v : v;
  }
}

// per-channel variant that also reduces the bias gradient: grid (C, S); a block walks (batch item, 1024-position
// chunk) pairs of its channel, 4 consecutive positions per thread (one 16-byte access when VEC)
template <bool VEC>
__global__ void __launch_bounds__(256) leaky_relu_bwd_bias_kernel(const float* __restrict__ dy, const float* __restrict__ ref,
                                           const unsigned char* __restrict__ mask, float* __restrict__ dx,
                                           float* __restrict__ dbias, int B, int C, int T, float slope,
                                           float beta) {
  __shared__ double sh[32];
  const int c = blockIdx.x;
  const int nchunk = (T + 1023) / 1024;
  float acc = 0.f;
  for (int sidx = blockIdx.y; sidx < B * nchunk; sidx += gridDim.y) {
    const int b = sidx / nchunk, t = (sidx % nchunk) * 1024 + threadIdx.x * 4;
    if (t >= T) continue;
    const long long i = ((long long)b * C + c) * T + t;
    if (VEC) {                                           // T % 4 == 0: the four positions exist and are aligned
      const float4 d = *reinterpret_cast<const float4*>(dy + i);
      bool p0, p1, p2, p3;
      if (mask) {
        const uchar4 m = *reinterpret_cast<const uchar4*>(mask + i);
        p0 = m.x != 0; p1 = m.y != 0; p2 = m.z != 0; p3 = m.w != 0;
      } else if (ref) {
        const float4 r = *reinterpret_cast<const float4*>(ref + i);
        p0 = r.x > 0.f; p1 = r.y > 0.f; p2 = r.z > 0.f; p3 = r.w > 0.f;
      } else {
        p0 = p1 = p2 = p3 = true;
      }
      float4 v = make_float4(d.x * (p0 ? 1.f : slope), d.y * (p1 ? 1.f : slope), d.z * (p2 ? 1.f : slope),
                             d.w * (p3 ? 1.f : slope));
      acc += (v.x + v.y) + (v.z + v.w);
      if (dx) {
        if (beta != 0.f) {
          const float4 o = *reinterpret_cast<const float4*>(dx + i);
          v.x += beta * o.x; v.y += beta * o.y; v.z += beta * o.z; v.w += beta * o.w;
        }
        *reinterpret_cast<float4*>(dx + i) = v;
      }
    } else {
      for (int e = 0; e < 4 && t + e < T; ++e) {
        const long long j = i + e;
        const bool pos = mask ? mask[j] != 0 : (ref ? ref[j] > 0.f : true);
        const float v = dy[j] * (pos ? 1.f : slope);
        acc += v;
        if (dx) dx[j] = beta != 0.f ? beta * dx[j] + v : v;
      }
    }
  }
  const double tot = block_sum((double)acc, sh);
  if (threadIdx.x == 0) atomicAdd(dbias + c, (float)tot);
}

__global__ void tanh_recompose_fwd_kernel(const float* __restrict__ x, const float* __restrict__ first,
                                          float* __restrict__ y, int B, int m, int p, int T) {
  const long long n = (long long)B * m * T;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    int t = (int)(i % T);
    long long r = i / T;
    int c = (int)(r % m), b = (int)(r / m);
    float v = x[i];
    if (c < p) v += first[((long long)b * p + c) * T + t];
    y[i] = tanhf(v);
  }
}
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float t = y[i];
    dx[i] = dy[i] * (1.f - t * t);
  }
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n) {
  const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long i = i0; i < n4; i += stride) {
    float4 u = reinterpret_cast<const float4*>(a)[i], v = reinterpret_cast<const float4*>(b)[i];
    u.x += v.x; u.y += v.y; u.z += v.z; u.w += v.w;
    reinterpret_cast<float4*>(y)[i] = u;
  }
  for (long long i = (n4 << 2) + i0; i < n; i += stride) y[i] = a[i] + b[i];
}
__global__ void axpby_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float alpha, float beta) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = alpha * x[i];
    y[i] = beta != 0.f ? beta * y[i] + v : v;
  }
}
__global__ void axpby_dev_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                 const float* __restrict__ alpha, float beta) {
  const float a = alpha[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = a * x[i];
    y[i] = beta != 0.f ? beta * y[i] + v : v;
  }
}
__global__ void fill_kernel(float* __restrict__ p, long long n, float value) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) p[i] = value;
}

// ------------------------------------------------------------------ losses
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(256) l1_pair_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                    double* __restrict__ sums) {
  __shared__ double sh[32];
  double s_ab = 0.0, s_a = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float x = a[i], y = b[i];
    s_ab += fabsf(x - y);
    s_a += fabsf(x);
  }
  s_ab = block_sum(s_ab, sh);
  s_a = block_sum(s_a, sh);
  if (threadIdx.x == 0) { atomicAdd(sums, s_ab); atomicAdd(sums + 1, s_a); }
}
__global__ void fm_finalize_kernel(const double* __restrict__ sums, int npairs, float scale, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < npairs; ++i) tot += (float)(sums[2 * i] / sums[2 * i + 1]);
    loss[0] = tot * scale;
  }
}
__global__ void l1_pair_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                   const double* __restrict__ sums, const float* __restrict__ go, float scale,
                                   float* __restrict__ da, float* __restrict__ db) {
  const double s_ab = sums[0], s_a = sums[1];
  const float gsc = go[0] * scale;
  const float c1 = (float)(1.0 / s_a) * gsc, c2 = (float)(s_ab / (s_a * s_a)) * gsc;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float x = a[i], y = b[i];
    float sd = sgn(x - y);
    if (da) da[i] = c1 * sd - c2 * sgn(x);
    if (db) db[i] = -c1 * sd;
  }
}
// ResidualUnit backward through the COMPOSED conv (include/vbx.h: vbx_unit_combine / vbx_unit_split_grads).
// wf[co][ci][k] = sum_m w2[co][m] * w1[m][ci][k]
__global__ void unit_combine_kernel(const float* __restrict__ w1, const float* __restrict__ w2, int C, int K,
                                    float* __restrict__ wf) {
  const int n = C * C * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int co = i / (C * K), r = i % (C * K);
    float v = 0.f;
    for (int m = 0; m < C; ++m) v = fmaf(w2[co * C + m], w1[m * C * K + r], v);
    wf[i] = v;
  }
}
// dw1[m][ci][k] = beta*dw1 + sum_co w2[co][m] * dwf[co][ci][k];   dw2[co][m] = beta*dw2 + sum_{ci,k} dwf[co][ci][k] * w1[m][ci][k]
__global__ void unit_split_grads_kernel(const float* __restrict__ dwf, const float* __restrict__ w1,
                                        const float* __restrict__ w2, int C, int K, float* __restrict__ dw1,
                                        float* __restrict__ dw2, float beta, int nb1) {
  if ((int)blockIdx.x < nb1) {                            // dw1: one thread per output, coalesced over (ci, k)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * C * K) return;
    const int m = i / (C * K), r = i % (C * K);
    float v = 0.f;
#pragma unroll 4
    for (int co = 0; co < C; ++co) v = fmaf(w2[co * C + m], dwf[co * C * K + r], v);
    dw1[i] = beta != 0.f ? fmaf(beta, dw1[i], v) : v;
  } else {                                                // dw2: one warp per output, lanes over the (ci, k) products
    const int j = ((int)blockIdx.x - nb1) * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= C * C) return;
    const int co = j / C, m = j % C;
    const float* a = dwf + (long long)co * C * K;
    const float* b = w1 + (long long)m * C * K;
    float v = 0.f;
    for (int r = lane; r < C * K; r += 32) v = fmaf(a[r], b[r], v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) dw2[j] = beta != 0.f ? fmaf(beta, dw2[j], v) : v;
  }
}
// Mirror terms of a k = 3, stride-1 conv whose reflect halo equals its dilation d (the residual units' dilated conv):
// the input gradient is the ZERO-halo input gradient plus, on the d positions next to each edge,
//   dx[b, ci, u]         += sum_co w[co, ci, 0] * dy[b, co, d - u]              (u = 1 .. d)
//   dx[b, ci, T - 1 - j] += sum_co w[co, ci, 2] * dy[b, co, T - 1 + j - d]      (j = 1 .. d)
// (the other taps of a mirrored position fall outside the signal).  One thread per (b, ci, side, j).
__global__ void reflect_fold_k3_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                       int B, int C, int T, int d) {
  const int n = B * C * 2 * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i % d + 1, side = (i / d) & 1, ci = (i / (2 * d)) % C, b = i / (2 * d * C);
    const int u = side == 0 ? j : T - 1 - j;
    const int t = side == 0 ? d - j : T - 1 + j - d;
    if (u < 0 || u >= T || t < 0 || t >= T) continue;
    const int k = side == 0 ? 0 : 2;
    const float* dyb = dy + (long long)b * C * T + t;
    float v = 0.f;
#pragma unroll 4
    for (int co = 0; co < C; ++co) v = fmaf(w[((long long)co * C + ci) * 3 + k], dyb[(long long)co * T], v);
    dx[((long long)b * C + ci) * T + u] += v;
  }
}
// the two scalars of l1_pair_bwd_kernel per layer, for the conv epilogue's gate stage and fm_gate_bwd_kernel
__global__ void fm_coef_kernel(const double* __restrict__ sums, int n, const float* __restrict__ go, float scale,
                               float* __restrict__ coef) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s_ab = sums[2 * i], s_a = sums[2 * i + 1];
  const float gsc = go[0] * scale;
  coef[2 * i] = (float)(1.0 / s_a) * gsc;
  coef[2 * i + 1] = (float)(s_ab / (s_a * s_a)) * gsc;
}
__global__ void fm_gate_bwd_kernel(const float* __restrict__ y, const float* __restrict__ other,
                                   const float* __restrict__ coef, float gslope, const float* __restrict__ g,
                                   long long n, float* __restrict__ out) {
  const bool fm = other != nullptr;
  const float c1 = fm ? coef[0] : 0.f, c2 = fm ? coef[1] : 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = vbx::gate_apply(g ? g[i] : 0.f, y[i], fm, fm ? other[i] : 0.f, c1, c2, gslope);
}
__global__ void __launch_bounds__(256) hinge_fwd_kernel(const float* __restrict__ c, long long n, float target, float scale,
                                 double* __restrict__ acc) {
  __shared__ double sh[32];
  double s = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = 1.f - target * c[i];
    s += v > 0.f ? v : 0.f;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(acc, s * (double)scale);
}
__global__ void hinge_bwd_kernel(const float* __restrict__ c, long long n, float target, float scale,
                                 const float* __restrict__ go, float* __restrict__ dc) {
  const float g = go[0] * scale;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    dc[i] = (1.f - target * c[i]) > 0.f ? -target * g : 0.f;
}
__global__ void d2f_kernel(const double* __restrict__ src, float* __restrict__ dst, int n, float scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i] * scale;
}

__global__ void __launch_bounds__(256) stft_stats_kernel(const float* __restrict__ X, const float* __restrict__ Y, int B, int bins, int F,
                                  float eps, double* __restrict__ stats) {
  __shared__ double sh[32];
  const long long n = (long long)B * bins * F, per = (long long)bins * F;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    long long b = i / per, r = i % per;
    long long ire = b * 2 * per + r, iim = ire + per;
    float xr = X[ire], xi = X[iim], yr = Y[ire], yi = Y[iim];
    float xm = sqrtf(fmaxf(xr * xr + xi * xi, eps)), ym = sqrtf(fmaxf(yr * yr + yi * yi, eps));
    float d = ym - xm;
    s0 += (double)d * d;
    s1 += (double)ym * ym;
    s2 += fabsf(logf(xm) - logf(ym));
  }
  s0 = block_sum(s0, sh); s1 = block_sum(s1, sh); s2 = block_sum(s2, sh);
  if (threadIdx.x == 0) { atomicAdd(stats, s0); atomicAdd(stats + 1, s1); atomicAdd(stats + 2, s2); }
}
__global__ void stft_finalize_kernel(const double* __restrict__ stats, const double* __restrict__ counts, int nres, float w,
                                     float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float tot = 0.f;
    for (int r = 0; r < nres; ++r) {
      float sc = sqrtf((float)stats[3 * r]) / sqrtf((float)stats[3 * r + 1]);
      float lg = (float)(stats[3 * r + 2] / counts[r]);
      tot += sc + lg;
    }
    loss[0] = tot * w;
  }
}
__global__ void stft_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, int B, int bins, int F, float eps,
                                const double* __restrict__ stats, double count, const float* __restrict__ go, float w,
                                float* __restrict__ dX) {
  const long long n = (long long)B * bins * F, per = (long long)bins * F;
  const float g = go[0] * w;
  const float c_sc = g / (sqrtf((float)stats[0]) * sqrtf((float)stats[1]));
  const float c_lg = g / (float)count;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    long long b = i / per, r = i % per;
    long long ire = b * 2 * per + r, iim = ire + per;
    float xr = X[ire], xi = X[iim], yr = Y[ire], yi = Y[iim];
    float px = xr * xr + xi * xi;
    float xm = sqrtf(fmaxf(px, eps)), ym = sqrtf(fmaxf(yr * yr + yi * yi, eps));
    float dxm = c_sc * (xm - ym) + c_lg * sgn(logf(xm) - logf(ym)) / xm;
    float k = px >= eps ? dxm / xm : 0.f;
    dX[ire] = k * xr;
    dX[iim] = k * xi;
  }
}

__global__ void weighted_sum_kernel(const float* x0, const float* x1, const float* x2, const float* x3, int n,
                                    const float* lam, float* terms, float* total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float* xs[4] = {x0, x1, x2, x3};
  float tot = 0.f;
  for (int i = 0; i < n; ++i) {
    float v = xs[i][0] * (lam ? lam[i] : 1.f);
    if (terms) terms[i] = v;
    tot += v;
  }
  total[0] = tot;
}
__global__ void scalar_mul_kernel(const float* go, const float* lam, float* out, int n) {
  int i = threadIdx.x;
  if (i < n) out[i] = go[0] * (lam ? lam[i] : 1.f);
}
// STFT framing: U[b,k,f] = x[b, map(f*hop + k - pad)] (reflect halo), and its adjoint.
__global__ void unfold_frames_kernel(const float* __restrict__ x, float* __restrict__ U, int L, int K, int F,
                                     int hop, int pad) {
  const int b = blockIdx.z, k = blockIdx.y;
  const float* xb = x + (long long)b * L;
  float* out = U + ((long long)b * K + k) * F;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
    int p = f * hop + k - pad;
    if (p < 0) p = -p;
    else if (p >= L) p = 2 * (L - 1) - p;
    out[f] = (p >= 0 && p < L) ? xb[p] : 0.f;
  }
}
// dx[b,p] (+)= sum over the (f,k) whose (mirrored) position is p of dU[b,k,f]; gather form, no atomics
__global__ void fold_frames_kernel(const float* __restrict__ dU, float* __restrict__ dx, int L, int K, int F,
                                   int hop, int pad, float beta) {
  const int b = blockIdx.y;
  const float* ub = dU + (long long)b * K * F;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < L; p += gridDim.x * blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int img = 0; img < 3; ++img) {
      int q;
      if (img == 0) q = p;
      else if (img == 1) { if (p < 1 || p > pad) continue; q = -p; }
      else { if (p > L - 2 || p < L - 1 - pad) continue; q = 2 * (L - 1) - p; }
      const int a = q + pad;                                  // = f*hop + k
      int f_hi = a / hop;
      if (f_hi > F - 1) f_hi = F - 1;
      int f_lo = (a - K + 1 + hop - 1) / hop;
      if (a - K + 1 <= 0) f_lo = 0;
      for (int f = f_lo; f <= f_hi; ++f) acc += ub[(long long)(a - f * hop) * F + f];
    }
    float* o = dx + (long long)b * L + p;
    *o = beta != 0.f ? beta * *o + acc : acc;
  }
}

// ------------------------------------------------------------------ reductions / optimiser
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc) {
  __shared__ double sh[32];
  double s = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) { float v = x[i]; s += (double)v * v; }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}
__global__ void balance_kernel(const double* __restrict__ sumsq, float* __restrict__ norms_old, int* __restrict__ initialised,
                               float* __restrict__ lambdas, float* __restrict__ norms_out, int n, float beta, int mode) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool first = initialised[0] == 0;
  for (int i = 0; i < n; ++i) {
    float nm = sqrtf((float)sumsq[i]);
    if (norms_out) norms_out[i] = nm;
    float old = (first || mode == 0) ? nm : norms_old[i];
    if (mode == 1) old = beta * old + (1.f - beta) * nm;
    norms_old[i] = old;
    float lam = 1.f / (old + 1e-4f);
    lam = fminf(fmaxf(lam, 0.f), 1e4f);
    lambdas[i] = lam;
  }
  initialised[0] = 1;
}
__global__ void adam_tick_kernel(int* step) { if (threadIdx.x == 0 && blockIdx.x == 0) step[0] += 1; }
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                 long long n, const int* __restrict__ step, float lr, float b1, float b2, float eps,
                                 float grad_scale) {
  const int t = step[0];
  const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
  const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = grad[i] * grad_scale;
    float mi = m[i] + (g - m[i]) * (1.f - b1);       // torch: exp_avg.lerp_(grad, 1-beta1)
    float vi = v[i] * b2 + (1.f - b2) * g * g;       // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}
__global__ void noise_mix_crop_kernel(const float* __restrict__ body, const float* __restrict__ air, const float* __restrict__ noise,
                                      const int* __restrict__ start, const int* __restrict__ off, float* __restrict__ out_body,
                                      float* __restrict__ out_air, int Ls, int Ln, int len) {
  const int b = blockIdx.y;
  const int o = off[b], s = start[b];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
    int src = o + i;
    float bo = 0.f, ai = 0.f;
    if (src < Ls) {
      bo = body[(long long)b * Ls + src];
      ai = air[(long long)b * Ls + src];
      int ni = s + src;
      if (ni >= 0 && ni < Ln) bo += noise[(long long)b * Ln + ni];
    }
    out_body[(long long)b * len + i] = bo;
    out_air[(long long)b * len + i] = ai;
  }
}
}  // namespace vbx

using namespace vbx;
#define ST ((cudaStream_t)stream)

extern "C" int vbx_abi_version(void) { return VBX_ABI_VERSION; }
extern "C" const char* vbx_last_error(void) { return g_err; }
extern "C" uint64_t vbx_launch_count(void) { return g_launches.load(); }
extern "C" int vbx_set_tensor_core_mode(int mode) { int o = g_tc_mode; g_tc_mode = mode; return o; }
extern "C" int vbx_set_deterministic(int on) { int o = deterministic_flag(); deterministic_flag() = on ? 1 : 0; return o; }

namespace vbx {
struct ScalarPtrs { const float* p[8]; };
__global__ void gather_scalars_kernel(ScalarPtrs s, int n, float scale, float* __restrict__ out) {
  const int i = threadIdx.x;
  if (i < n && s.p[i]) out[i] = scale * s.p[i][0];
}
}  // namespace vbx
extern "C" int vbx_gather_scalars(const float* p0, const float* p1, const float* p2, const float* p3, const float* p4,
                                  const float* p5, const float* p6, const float* p7, int32_t n, float scale,
                                  float* out, void* stream) {
  VBX_REQUIRE(out, VBX_BAD_POINTER, "gather_scalars: null output");
  VBX_REQUIRE(n >= 1 && n <= 8, VBX_BAD_SHAPE, "gather_scalars: 1..8 scalars");
  ScalarPtrs s{{p0, p1, p2, p3, p4, p5, p6, p7}};
  gather_scalars_kernel<<<1, 32, 0, ST>>>(s, n, scale, out);
  return launched("gather_scalars_kernel");
}

extern "C" int vbx_weight_norm_fwd(const float* g, const float* v, float* w, float* wt, float* inv_norm,
                                   int32_t R, int32_t Cin_g, int32_t K, int32_t groups, void* stream) {
  VBX_REQUIRE(g && v && w && inv_norm, VBX_BAD_POINTER, "weight_norm_fwd: null tensor");
  VBX_REQUIRE(R > 0 && Cin_g > 0 && K > 0 && groups > 0 && R % groups == 0, VBX_BAD_SHAPE, "weight_norm_fwd: bad shape");
  weight_norm_fwd_kernel<<<R, 256, 0, ST>>>(g, v, w, wt, inv_norm, Cin_g, K, R / groups);
  return launched("weight_norm_fwd_kernel");
}
extern "C" int vbx_weight_norm_bwd(const float* g, const float* v, const float* inv_norm, const float* dw,
                                   float* dg, float* dv, int32_t R, int32_t row, float beta, void* stream) {
  VBX_REQUIRE(g && v && inv_norm && dw && dg && dv, VBX_BAD_POINTER, "weight_norm_bwd: null tensor");
  VBX_REQUIRE(R > 0 && row > 0, VBX_BAD_SHAPE, "weight_norm_bwd: bad shape");
  weight_norm_bwd_kernel<<<R, 256, 0, ST>>>(g, v, inv_norm, dw, dg, dv, row, beta);
  return launched("weight_norm_bwd_kernel");
}

extern "C" int vbx_pqmf_analysis(const float* x, const float* w, float* y, int32_t B, int32_t L, int32_t T,
                                 int32_t m, int32_t n, int32_t bands, int32_t x_per_band, void* stream) {
  VBX_REQUIRE(x && w && y, VBX_BAD_POINTER, "pqmf_analysis: null tensor");
  VBX_REQUIRE(B > 0 && L > 0 && T > 0 && m > 0 && n > 0 && bands > 0 && bands <= m && B <= 65535, VBX_BAD_SHAPE,
              "pqmf_analysis: bad shape");
  const int QL = 256 + (n + m - 1) / m + 1;
  size_t smem = sizeof(float) * ((size_t)bands * n + (size_t)m * QL);
  VBX_REQUIRE(smem <= 48 * 1024, VBX_UNSUPPORTED, "pqmf_analysis: filter bank too large for shared memory");
  dim3 grid(cdiv(T, 256), B);
  if (m == 4 && n == 32) {
    if (x_per_band) pqmf_analysis_kernel<true, 4, 32><<<grid, 256, smem, ST>>>(x, w, y, L, T, m, n, bands);
    else pqmf_analysis_kernel<false, 4, 32><<<grid, 256, smem, ST>>>(x, w, y, L, T, m, n, bands);
  } else if (x_per_band) pqmf_analysis_kernel<true, 0, 0><<<grid, 256, smem, ST>>>(x, w, y, L, T, m, n, bands);
  else pqmf_analysis_kernel<false, 0, 0><<<grid, 256, smem, ST>>>(x, w, y, L, T, m, n, bands);
  return launched("pqmf_analysis_kernel");
}
extern "C" int vbx_pqmf_synthesis(const float* x, const float* w, float* y, int32_t B, int32_t T, int32_t L,
                                  int32_t m, int32_t n, int32_t bands, int32_t sum_bands, void* stream) {
  VBX_REQUIRE(x && w && y, VBX_BAD_POINTER, "pqmf_synthesis: null tensor");
  VBX_REQUIRE(B > 0 && L > 0 && T > 0 && m > 0 && n > 0 && bands > 0 && bands <= m && B <= 65535, VBX_BAD_SHAPE,
              "pqmf_synthesis: bad shape");
  const int J = (n + m - 1) / m, QL = 1024 / m + J + 2;
  size_t smem = sizeof(float) * ((size_t)bands * n + (size_t)bands * QL);
  VBX_REQUIRE(smem <= 48 * 1024, VBX_UNSUPPORTED, "pqmf_synthesis: filter bank too large for shared memory");
  dim3 grid(cdiv(L, 1024), B);
  if (m == 4 && n == 32) pqmf_synthesis_kernel<4, 32><<<grid, 256, smem, ST>>>(x, w, y, T, L, m, n, bands, sum_bands);
  else pqmf_synthesis_kernel<0, 0><<<grid, 256, smem, ST>>>(x, w, y, T, L, m, n, bands, sum_bands);
  return launched("pqmf_synthesis_kernel");
}

extern "C" int vbx_leaky_relu_fwd(const float* x, float* y, int64_t n, float slope, void* stream) {
  VBX_REQUIRE(x && y, VBX_BAD_POINTER, "leaky_relu_fwd: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "leaky_relu_fwd: empty");
  VBX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, VBX_BAD_POINTER, "leaky_relu_fwd: misaligned");
  leaky_relu_fwd_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(x, y, n, slope);
  return launched("leaky_relu_fwd_kernel");
}
namespace vbx {
int launch_lrelu_bwd_bias(const float* dy, const float* ref, float* dx, float* dbias, int B, int C, int T, float slope,
                          void* stream) {
  return vbx_leaky_relu_bwd(dy, ref, nullptr, dx, dbias, B, C, T, slope, 0.f, stream);
}
}  // namespace vbx
extern "C" int vbx_leaky_relu_bwd(const float* dy, const float* ref, const uint8_t* mask, float* dx, float* dbias,
                                  int32_t B, int32_t C, int32_t T, float slope, float beta, void* stream) {
  VBX_REQUIRE(dy && (dx || dbias), VBX_BAD_POINTER, "leaky_relu_bwd: null tensor");
  VBX_REQUIRE(B > 0 && C > 0 && T > 0, VBX_BAD_SHAPE, "leaky_relu_bwd: bad shape");
  long long n = (long long)B * C * T;
  const uintptr_t al = (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)ref;
  if (dbias) {
    VBX_REQUIRE(C <= 65535 * 32, VBX_UNSUPPORTED, "leaky_relu_bwd: too many channels");
    const long long units = (long long)B * ((T + 1023) / 1024);      // (item, chunk) pairs per channel
    int S = (int)(units < 65535 ? units : 65535);
    int cap = (kSMs * 16 + C - 1) / C;
    if (S > cap) S = cap;
    if (S < 1 || deterministic_flag()) S = 1;
    dim3 grid(C, S);
    const bool vec = (T & 3) == 0 && (al & 15) == 0 && ((uintptr_t)mask & 3) == 0;
    if (vec) leaky_relu_bwd_bias_kernel<true><<<grid, 256, 0, ST>>>(dy, ref, mask, dx, dbias, B, C, T, slope, beta);
    else leaky_relu_bwd_bias_kernel<false><<<grid, 256, 0, ST>>>(dy, ref, mask, dx, dbias, B, C, T, slope, beta);
    return launched("leaky_relu_bwd_bias_kernel");
  }
  VBX_REQUIRE(ref || mask, VBX_BAD_POINTER, "leaky_relu_bwd: need ref or mask");
  if ((n & 3) == 0 && (al & 15) == 0 && ((uintptr_t)mask & 3) == 0)
    leaky_relu_bwd_kernel<true><<<stream_blocks(n, 4096), 256, 0, ST>>>(dy, ref, mask, dx, n, slope, beta);
  else
    leaky_relu_bwd_kernel<false><<<stream_blocks(n, 1024), 256, 0, ST>>>(dy, ref, mask, dx, n, slope, beta);
  return launched("leaky_relu_bwd_kernel");
}
extern "C" int vbx_tanh_recompose_fwd(const float* x, const float* first, float* y, int32_t B, int32_t m,
                                      int32_t p, int32_t T, void* stream) {
  VBX_REQUIRE(x && y && (first || p == 0), VBX_BAD_POINTER, "tanh_recompose_fwd: null tensor");
  VBX_REQUIRE(B > 0 && m > 0 && p >= 0 && p <= m && T > 0, VBX_BAD_SHAPE, "tanh_recompose_fwd: bad shape");
  tanh_recompose_fwd_kernel<<<stream_blocks((long long)B * m * T, 1024), 256, 0, ST>>>(x, first, y, B, m, p, T);
  return launched("tanh_recompose_fwd_kernel");
}
extern "C" int vbx_tanh_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  VBX_REQUIRE(dy && y && dx, VBX_BAD_POINTER, "tanh_bwd: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "tanh_bwd: empty");
  tanh_bwd_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(dy, y, dx, n);
  return launched("tanh_bwd_kernel");
}
extern "C" int vbx_add(const float* a, const float* b, float* y, int64_t n, void* stream) {
  VBX_REQUIRE(a && b && y, VBX_BAD_POINTER, "add: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "add: empty");
  VBX_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 15) == 0, VBX_BAD_POINTER, "add: misaligned");
  add_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(a, b, y, n);
  return launched("add_kernel");
}
extern "C" int vbx_axpby(const float* x, float* y, int64_t n, float alpha, float beta, void* stream) {
  VBX_REQUIRE(x && y, VBX_BAD_POINTER, "axpby: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "axpby: empty");
  axpby_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(x, y, n, alpha, beta);
  return launched("axpby_kernel");
}
extern "C" int vbx_axpby_dev(const float* x, float* y, int64_t n, const float* alpha, float beta, void* stream) {
  VBX_REQUIRE(x && y && alpha, VBX_BAD_POINTER, "axpby_dev: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "axpby_dev: empty");
  axpby_dev_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(x, y, n, alpha, beta);
  return launched("axpby_dev_kernel");
}
extern "C" int vbx_fill(float* p, int64_t n, float value, void* stream) {
  VBX_REQUIRE(p, VBX_BAD_POINTER, "fill: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "fill: empty");
  fill_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(p, n, value);
  return launched("fill_kernel");
}

extern "C" int vbx_l1_pair_sums(const float* a, const float* b, int64_t n, double* sums, void* stream) {
  VBX_REQUIRE(a && b && sums, VBX_BAD_POINTER, "l1_pair_sums: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "l1_pair_sums: empty");
  l1_pair_sums_kernel<<<reduce_blocks(n, 2048), 256, 0, ST>>>(a, b, n, sums);
  return launched("l1_pair_sums_kernel");
}
extern "C" int vbx_fm_finalize(const double* sums, int32_t npairs, float scale, float* loss, void* stream) {
  VBX_REQUIRE(sums && loss, VBX_BAD_POINTER, "fm_finalize: null tensor");
  fm_finalize_kernel<<<1, 32, 0, ST>>>(sums, npairs, scale, loss);
  return launched("fm_finalize_kernel");
}
extern "C" int vbx_l1_pair_bwd(const float* a, const float* b, int64_t n, const double* sums, const float* go,
                               float scale, float* da, float* db, void* stream) {
  VBX_REQUIRE(a && b && sums && go && (da || db), VBX_BAD_POINTER, "l1_pair_bwd: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "l1_pair_bwd: empty");
  l1_pair_bwd_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(a, b, n, sums, go, scale, da, db);
  return launched("l1_pair_bwd_kernel");
}
extern "C" int vbx_reflect_fold_k3(const float* dy, const float* w, float* dx, int32_t B, int32_t C, int32_t T,
                                   int32_t dil, void* stream) {
  VBX_REQUIRE(dy && w && dx, VBX_BAD_POINTER, "reflect_fold_k3: null tensor");
  VBX_REQUIRE(B > 0 && C > 0 && T > 1 && dil > 0 && dil <= T - 1 && (long long)B * C * 2 * dil < (1ll << 31), VBX_BAD_SHAPE,
              "reflect_fold_k3: bad shape");
  reflect_fold_k3_kernel<<<cdiv(B * C * 2 * dil, 128), 128, 0, ST>>>(dy, w, dx, B, C, T, dil);
  return launched("reflect_fold_k3_kernel");
}
extern "C" int vbx_unit_combine(const float* w1, const float* w2, int32_t C, int32_t K, float* wf, void* stream) {
  VBX_REQUIRE(w1 && w2 && wf, VBX_BAD_POINTER, "unit_combine: null tensor");
  VBX_REQUIRE(C > 0 && K > 0 && (long long)C * C * K < (1 << 30), VBX_BAD_SHAPE, "unit_combine: bad shape");
  unit_combine_kernel<<<cdiv(C * C * K, 128), 128, 0, ST>>>(w1, w2, C, K, wf);
  return launched("unit_combine_kernel");
}
extern "C" int vbx_unit_split_grads(const float* dwf, const float* w1, const float* w2, int32_t C, int32_t K,
                                    float* dw1, float* dw2, float beta, void* stream) {
  VBX_REQUIRE(dwf && w1 && w2 && (dw1 || dw2), VBX_BAD_POINTER, "unit_split_grads: null tensor");
  VBX_REQUIRE(C > 0 && K > 0 && (long long)C * C * K < (1 << 30), VBX_BAD_SHAPE, "unit_split_grads: bad shape");
  const int nb1 = dw1 ? cdiv(C * C * K, 128) : 0, nb2 = dw2 ? cdiv(C * C, 4) : 0;      // 4 warps per block
  unit_split_grads_kernel<<<nb1 + nb2, 128, 0, ST>>>(dwf, w1, w2, C, K, dw1, dw2, beta, nb1);
  return launched("unit_split_grads_kernel");
}
extern "C" int vbx_fm_coef(const double* sums, int32_t npairs, const float* go, float scale, float* coef, void* stream) {
  VBX_REQUIRE(sums && go && coef, VBX_BAD_POINTER, "fm_coef: null tensor");
  VBX_REQUIRE(npairs > 0, VBX_BAD_SHAPE, "fm_coef: no layers");
  fm_coef_kernel<<<cdiv(npairs, 64), 64, 0, ST>>>(sums, npairs, go, scale, coef);
  return launched("fm_coef_kernel");
}
namespace vbx {
int gate_epilogue_forms() {
  static const int forms = getenv("VBX_GATE_EPILOGUE") ? atoi(getenv("VBX_GATE_EPILOGUE")) : (1 | 8);
  return forms;
}
int launch_fm_gate(const float* y, const float* other, const float* coef, float gslope, const float* g, long long n,
                   float* out, void* stream) {
  fm_gate_bwd_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(y, other, coef, gslope, g, n, out);
  return launched("fm_gate_bwd_kernel");
}
}  // namespace vbx
extern "C" int vbx_fm_gate_bwd(const float* y, const float* other, const float* coef, float gate_slope, const float* g,
                               int64_t n, float* out, void* stream) {
  VBX_REQUIRE(y && out && (!other || coef), VBX_BAD_POINTER, "fm_gate_bwd: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "fm_gate_bwd: empty");
  return launch_fm_gate(y, other, coef, gate_slope, g, n, out, stream);
}
extern "C" int vbx_hinge_fwd(const float* c, int64_t n, float target, float scale, double* acc, void* stream) {
  VBX_REQUIRE(c && acc, VBX_BAD_POINTER, "hinge_fwd: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "hinge_fwd: empty");
  hinge_fwd_kernel<<<reduce_blocks(n, 2048), 256, 0, ST>>>(c, n, target, scale, acc);
  return launched("hinge_fwd_kernel");
}
extern "C" int vbx_hinge_bwd(const float* c, int64_t n, float target, float scale, const float* go, float* dc,
                             void* stream) {
  VBX_REQUIRE(c && go && dc, VBX_BAD_POINTER, "hinge_bwd: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "hinge_bwd: empty");
  hinge_bwd_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(c, n, target, scale, go, dc);
  return launched("hinge_bwd_kernel");
}
extern "C" int vbx_d2f(const double* src, float* dst, int32_t n, float scale, void* stream) {
  VBX_REQUIRE(src && dst, VBX_BAD_POINTER, "d2f: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "d2f: empty");
  d2f_kernel<<<cdiv(n, 128), 128, 0, ST>>>(src, dst, n, scale);
  return launched("d2f_kernel");
}
extern "C" int vbx_stft_stats(const float* X, const float* Y, int32_t B, int32_t bins, int32_t F, float eps,
                              double* stats, void* stream) {
  VBX_REQUIRE(X && Y && stats, VBX_BAD_POINTER, "stft_stats: null tensor");
  VBX_REQUIRE(B > 0 && bins > 0 && F > 0, VBX_BAD_SHAPE, "stft_stats: bad shape");
  stft_stats_kernel<<<reduce_blocks((long long)B * bins * F, 2048), 256, 0, ST>>>(X, Y, B, bins, F, eps, stats);
  return launched("stft_stats_kernel");
}
extern "C" int vbx_stft_finalize(const double* stats, const double* counts, int32_t nres, float w, float* loss,
                                 void* stream) {
  VBX_REQUIRE(stats && counts && loss, VBX_BAD_POINTER, "stft_finalize: null tensor");
  stft_finalize_kernel<<<1, 32, 0, ST>>>(stats, counts, nres, w, loss);
  return launched("stft_finalize_kernel");
}
extern "C" int vbx_stft_bwd(const float* X, const float* Y, int32_t B, int32_t bins, int32_t F, float eps,
                            const double* stats, double count, const float* go, float w, float* dX, void* stream) {
  VBX_REQUIRE(X && Y && stats && go && dX, VBX_BAD_POINTER, "stft_bwd: null tensor");
  VBX_REQUIRE(B > 0 && bins > 0 && F > 0, VBX_BAD_SHAPE, "stft_bwd: bad shape");
  stft_bwd_kernel<<<stream_blocks((long long)B * bins * F, 1024), 256, 0, ST>>>(X, Y, B, bins, F, eps, stats, count, go, w, dX);
  return launched("stft_bwd_kernel");
}
extern "C" int vbx_unfold_frames(const float* x, float* U, int32_t B, int32_t L, int32_t K, int32_t hop,
                                 int32_t pad, void* stream) {
  VBX_REQUIRE(x && U, VBX_BAD_POINTER, "unfold_frames: null tensor");
  VBX_REQUIRE(B > 0 && L > 1 && K > 0 && hop > 0 && pad >= 0 && pad <= L - 1 && K <= 65535 && B <= 65535 &&
                  L + 2 * pad >= K, VBX_BAD_SHAPE, "unfold_frames: bad shape");
  const int F = (L + 2 * pad - K) / hop + 1;
  dim3 grid(cdiv(F, 256), K, B);
  unfold_frames_kernel<<<grid, 256, 0, ST>>>(x, U, L, K, F, hop, pad);
  return launched("unfold_frames_kernel");
}
extern "C" int vbx_fold_frames(const float* dU, float* dx, int32_t B, int32_t L, int32_t K, int32_t hop,
                               int32_t pad, float beta, void* stream) {
  VBX_REQUIRE(dU && dx, VBX_BAD_POINTER, "fold_frames: null tensor");
  VBX_REQUIRE(B > 0 && L > 1 && K > 0 && hop > 0 && pad >= 0 && pad <= L - 1 && B <= 65535 && L + 2 * pad >= K,
              VBX_BAD_SHAPE, "fold_frames: bad shape");
  const int F = (L + 2 * pad - K) / hop + 1;
  dim3 grid(cdiv(L, 256), B);
  fold_frames_kernel<<<grid, 256, 0, ST>>>(dU, dx, L, K, F, hop, pad, beta);
  return launched("fold_frames_kernel");
}
extern "C" int vbx_weighted_sum(const float* x0, const float* x1, const float* x2, const float* x3, int32_t n,
                                const float* lam, float* terms, float* total, void* stream) {
  VBX_REQUIRE(n >= 1 && n <= 4, VBX_BAD_SHAPE, "weighted_sum: n must be 1..4");
  const float* xs[4] = {x0, x1, x2, x3};
  for (int i = 0; i < n; ++i) VBX_REQUIRE(xs[i], VBX_BAD_POINTER, "weighted_sum: null term");
  VBX_REQUIRE(total, VBX_BAD_POINTER, "weighted_sum: null output");
  weighted_sum_kernel<<<1, 32, 0, ST>>>(x0, x1, x2, x3, n, lam, terms, total);
  return launched("weighted_sum_kernel");
}
extern "C" int vbx_scalar_mul(const float* go, const float* lam, float* out, int32_t n, void* stream) {
  VBX_REQUIRE(go && out, VBX_BAD_POINTER, "scalar_mul: null tensor");
  VBX_REQUIRE(n >= 1 && n <= 32, VBX_BAD_SHAPE, "scalar_mul: n must be 1..32");
  scalar_mul_kernel<<<1, 32, 0, ST>>>(go, lam, out, n);
  return launched("scalar_mul_kernel");
}
extern "C" int vbx_sumsq(const float* x, int64_t n, double* acc, void* stream) {
  VBX_REQUIRE(x && acc, VBX_BAD_POINTER, "sumsq: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "sumsq: empty");
  sumsq_kernel<<<reduce_blocks(n, 2048), 256, 0, ST>>>(x, n, acc);
  return launched("sumsq_kernel");
}
extern "C" int vbx_balance(const double* sumsq, float* norms_old, int32_t* initialised, float* lambdas,
                           float* norms_out, int32_t n, float beta_ema, int32_t mode, void* stream) {
  VBX_REQUIRE(sumsq && norms_old && initialised && lambdas, VBX_BAD_POINTER, "balance: null tensor");
  VBX_REQUIRE(n > 0 && (mode == 0 || mode == 1), VBX_BAD_SHAPE, "balance: bad arguments");
  balance_kernel<<<1, 32, 0, ST>>>(sumsq, norms_old, initialised, lambdas, norms_out, n, beta_ema, mode);
  return launched("balance_kernel");
}
extern "C" int vbx_adam_tick(int32_t* step, void* stream) {
  VBX_REQUIRE(step, VBX_BAD_POINTER, "adam_tick: null");
  adam_tick_kernel<<<1, 32, 0, ST>>>(step);
  return launched("adam_tick_kernel");
}
extern "C" int vbx_adam_step(float* p, const float* grad, float* m, float* v, int64_t n, const int32_t* step,
                             float lr, float b1, float b2, float eps, float grad_scale, void* stream) {
  VBX_REQUIRE(p && grad && m && v && step, VBX_BAD_POINTER, "adam_step: null tensor");
  VBX_REQUIRE(n > 0, VBX_BAD_SHAPE, "adam_step: empty");
  adam_step_kernel<<<stream_blocks(n, 1024), 256, 0, ST>>>(p, grad, m, v, n, step, lr, b1, b2, eps, grad_scale);
  return launched("adam_step_kernel");
}
extern "C" int vbx_noise_mix_crop(const float* body, const float* air, const float* noise, const int32_t* start,
                                  const int32_t* off, float* out_body, float* out_air, int32_t B, int32_t Ls,
                                  int32_t Ln, int32_t len, void* stream) {
  VBX_REQUIRE(body && air && noise && start && off && out_body && out_air, VBX_BAD_POINTER, "noise_mix_crop: null tensor");
  VBX_REQUIRE(B > 0 && B <= 65535 && Ls > 0 && Ln > 0 && len > 0, VBX_BAD_SHAPE, "noise_mix_crop: bad shape");
  dim3 grid(cdiv(len, 1024), B);
  noise_mix_crop_kernel<<<grid, 256, 0, ST>>>(body, air, noise, start, off, out_body, out_air, Ls, Ln, len);
  return launched("noise_mix_crop_kernel");
}
