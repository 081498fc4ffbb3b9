// Conv1d family entry points: tile-config dispatch over the implicit-GEMM skeleton
// in gemm_conv.cuh.  See include/vbx.h for the reference call sites replaced.
#include "common.cuh"
#include "conv_plan.h"

namespace vbx {

template <class C, int MODE, bool BK>
__global__ void __launch_bounds__(256) gemm_conv_kernel(const GemmP P) {
  using T = Tile<C, MODE, BK>;
  __shared__ __align__(16) float As[C::KC * C::LDA];
  __shared__ __align__(16) float Bs[C::KC * C::LDB];
  typename T::TS s;
  const Blk blk{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z};
  const int tid = threadIdx.x;
  T::prologue(P, blk, tid, s);
  if (s.n_base >= s.N) return;                       // uniform: DGRAD phases differ in length
  if (MODE == DGRAD) {
    for (int img = 0; img < 3; ++img) {
      if (!T::dgrad_need_img(P, blk, img)) continue;  // uniform
      T::dgrad_begin_img(P, tid, s, img);
      const int nch = T::num_chunks(s);
      for (int c = 0; c < nch; ++c) {
        T::load_chunk(P, tid, s, c, As, Bs);
        __syncthreads();
        T::compute_chunk(tid, s, As, Bs);
        __syncthreads();
      }
    }
  } else {
    const int nch = T::num_chunks(s);
    for (int c = 0; c < nch; ++c) {
      T::load_chunk(P, tid, s, c, As, Bs);
      __syncthreads();
      T::compute_chunk(tid, s, As, Bs);
      __syncthreads();
    }
  }
  T::epilogue(P, blk, tid, s, [](float* p, float v) { atomicAdd(p, v); });
}

template <int MODE, bool BK>
static int launch_cfg(const Plan& pl, const GemmP& P, cudaStream_t st) {
  dim3 grid(pl.grid[0], pl.grid[1], pl.grid[2]);
  if (grid.y > 65535 || grid.z > 65535) return fail(VBX_UNSUPPORTED, "conv: grid too large");
  switch (pl.tm) {
    case 128: gemm_conv_kernel<C128, MODE, BK><<<grid, 256, 0, st>>>(P); break;
    case 64: gemm_conv_kernel<C64, MODE, BK><<<grid, 256, 0, st>>>(P); break;
    case 32: gemm_conv_kernel<C32, MODE, BK><<<grid, 256, 0, st>>>(P); break;
    case 16: gemm_conv_kernel<C16, MODE, BK><<<grid, 256, 0, st>>>(P); break;
    case 8: gemm_conv_kernel<C8, MODE, BK><<<grid, 256, 0, st>>>(P); break;
    default: gemm_conv_kernel<C4, MODE, BK><<<grid, 256, 0, st>>>(P); break;
  }
  return launched("gemm_conv_kernel");
}

// direct_conv.cu: one input channel per group
bool direct_fwd_ok(const GemmP& P);
bool direct_dgrad_ok(const GemmP& P);
int direct_fwd(const GemmP& P, cudaStream_t st);
int direct_dgrad(const GemmP& P, cudaStream_t st);
bool skinny_wgrad_ok(const GemmP& P);
int skinny_wgrad(const GemmP& P, cudaStream_t st);
bool skinny_fwd_ok(const GemmP& P);
int skinny_fwd(const GemmP& P, cudaStream_t st);

static int check_desc(const vbx_conv_desc* d) {
  int code = 0;
  const char* msg = check_desc_msg(d, &code);
  return msg ? fail(code, msg) : 0;
}

}  // namespace vbx

using namespace vbx;

extern "C" int vbx_conv1d_fwd(const vbx_conv_desc* d, const float* x, const float* w,
                              const vbx_epilogue* e, float* y, void* stream) {
  if (int r = check_desc(d)) return r;
  VBX_REQUIRE(x && w && y, VBX_BAD_POINTER, "conv1d_fwd: null tensor");
  GemmP P; fill(P, d); fill_epi(P, e);
  P.W = w; P.X = x; P.Y = y;
  if (direct_fwd_ok(P)) return direct_fwd(P, (cudaStream_t)stream);
  if (skinny_fwd_ok(P)) return skinny_fwd(P, (cudaStream_t)stream);
  Plan pl = plan_conv(FWD, P);
  if (pl.bk) return launch_cfg<FWD, true>(pl, P, (cudaStream_t)stream);
  return launch_cfg<FWD, false>(pl, P, (cudaStream_t)stream);
}

extern "C" int vbx_conv1d_dgrad(const vbx_conv_desc* d, const float* dy, const float* wt,
                                const vbx_epilogue* e, float* dx, void* stream) {
  if (int r = check_desc(d)) return r;
  VBX_REQUIRE(dy && wt && dx, VBX_BAD_POINTER, "conv1d_dgrad: null tensor");
  GemmP P; fill(P, d); fill_epi(P, e);
  P.W = wt; P.X = dy; P.Y = dx;
  VBX_REQUIRE(!P.gate || P.beta == 0.f, VBX_UNSUPPORTED, "conv1d_dgrad: gate stage with beta != 0");
  const GateArgs gate(P, 8);
  int rc;
  if (direct_dgrad_ok(P)) rc = direct_dgrad(P, (cudaStream_t)stream);
  else {
    Plan pl = plan_conv(DGRAD, P);
    rc = launch_cfg<DGRAD, false>(pl, P, (cudaStream_t)stream);
  }
  return gate.finish(rc, dx, d->B, d->Cin, d->Tin, stream);
}

extern "C" int vbx_conv1d_wgrad(const vbx_conv_desc* d, const float* x, const float* dy, float* dw,
                                void* stream) {
  if (int r = check_desc(d)) return r;
  VBX_REQUIRE(x && dy && dw, VBX_BAD_POINTER, "conv1d_wgrad: null tensor");
  GemmP P; fill(P, d);
  P.X = x; P.DY = dy; P.Y = dw;
  if (skinny_wgrad_ok(P)) return skinny_wgrad(P, (cudaStream_t)stream);
  Plan pl = plan_conv(WGRAD, P);
  return launch_cfg<WGRAD, true>(pl, P, (cudaStream_t)stream);
}

extern "C" int vbx_conv1d_dgrad_scatter(const vbx_conv_desc* d, const float* dy, const float* wk,
                                        float* dx, void* stream) {
  if (int r = check_desc(d)) return r;
  VBX_REQUIRE(dy && wk && dx, VBX_BAD_POINTER, "conv1d_dgrad_scatter: null tensor");
  GemmP P; fill(P, d);
  P.W = wk; P.X = dy; P.Y = dx;
  Plan pl = plan_conv(SCATTER, P);
  return launch_cfg<SCATTER, false>(pl, P, (cudaStream_t)stream);
}

namespace vbx {
__global__ void transpose_weight_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout,
                                        int Cin_g, int K, int groups) {
  const long long total = (long long)Cout * Cin_g * K;
  const int Cout_g = Cout / groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long r = i / K;
    int ci = (int)(r % Cin_g);
    int co = (int)(r / Cin_g);
    int g = co / Cout_g, col = co % Cout_g;
    wt[(((long long)g * Cin_g + ci) * Cout_g + col) * K + k] = w[i];
  }
}
}  // namespace vbx

extern "C" int vbx_transpose_weight(const float* w, float* wt, int32_t Cout, int32_t Cin_g, int32_t K,
                                    int32_t groups, void* stream) {
  VBX_REQUIRE(w && wt, VBX_BAD_POINTER, "transpose_weight: null tensor");
  VBX_REQUIRE(Cout > 0 && Cin_g > 0 && K > 0 && groups > 0 && Cout % groups == 0, VBX_BAD_SHAPE,
              "transpose_weight: bad shape");
  long long total = (long long)Cout * Cin_g * K;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  transpose_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, wt, Cout, Cin_g, K, groups);
  return launched("transpose_weight_kernel");
}
