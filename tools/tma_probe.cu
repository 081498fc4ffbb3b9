// Bring-up probe for 3-D tensor-map TMA loads on sm_100a: one tiny kernel, several descriptor / issue variants
// (argv[1]), each checked element-wise on the host incl. out-of-range zero fill.  A device-side trap kills the
// context, so run one variant per process:  for v in 0 1 2 3 4 5 6 7; do ./tools/tma_probe $v; done
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Prm { int W, C, c0, c1, c2, rank, elect, fromglobal; };

__global__ void probe(const __grid_constant__ CUtensorMap tm, const CUtensorMap* gtm, float* out, Prm P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  float* buf = reinterpret_cast<float*>(smem);
  const uint32_t bytes = (uint32_t)(P.W * P.C * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const CUtensorMap* map = P.fromglobal ? gtm : &tm;
  bool issue = threadIdx.x == 0;
  if (P.elect && threadIdx.x < 32) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    issue = pred != 0;
  }
  if (issue) {
    if (P.fromglobal) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    if (P.rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(buf)), "l"(reinterpret_cast<uint64_t>(map)), "r"(s32(bar)), "r"(P.c0), "r"(P.c1), "r"(P.c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(s32(buf)), "l"(reinterpret_cast<uint64_t>(map)), "r"(s32(bar)), "r"(P.c0), "r"(P.c1) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)) : "memory");
  for (int i = threadIdx.x; i < P.W * P.C; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  const int T = 256, C = 32, B = 2;
  Prm P{132, C, -1, 0, 1, 3, 0, 0};
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (v == 1) P.c0 = 0;
  if (v == 2) P.W = 128;
  if (v == 3) { P.rank = 2; P.c1 = C; }            // 2-D view (T, C*B): row C = item 1, channel 0
  if (v == 4) promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  if (v == 5) P.fromglobal = 1;
  if (v == 6) P.elect = 1;
  if (v == 7) P.W = 64;
  if (v == 8) { P.W = 256; P.c0 = 100; }
  std::vector<float> h((size_t)B * C * T);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *x, *out;
  cudaMalloc(&x, h.size() * 4);
  cudaMalloc(&out, 256 * C * 4);
  cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  if (ce != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("v%d: no entry point (%d, %d)\n", v, (int)ce, (int)q); return 2; }
  CUtensorMap tm;
  cuuint64_t dims3[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B}, str3[2] = {(cuuint64_t)T * 4, (cuuint64_t)C * T * 4};
  cuuint64_t dims2[2] = {(cuuint64_t)T, (cuuint64_t)C * B}, str2[1] = {(cuuint64_t)T * 4};
  cuuint32_t box3[3] = {(cuuint32_t)P.W, (cuuint32_t)C, 1}, box2[2] = {(cuuint32_t)P.W, (cuuint32_t)C}, es[3] = {1, 1, 1};
  CUresult cr = P.rank == 3
      ? ((Enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims3, str3, box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
      : ((Enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims2, str2, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("v%d: encode -> %d; map words:", v, (int)cr);
  for (int i = 0; i < 8; ++i) printf(" %016llx", (unsigned long long)tm.opaque[i]);
  printf("\n");
  if (cr != CUDA_SUCCESS) return 3;
  CUtensorMap* gtm;
  cudaMalloc(&gtm, sizeof(tm));
  cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  probe<<<1, 128, 96 * 1024>>>(tm, gtm, out, P);
  ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { printf("v%d: kernel -> %s\n", v, cudaGetErrorString(ce)); return 1; }
  std::vector<float> r((size_t)P.W * C);
  cudaMemcpy(r.data(), out, r.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  const int b = 1;
  for (int c = 0; c < C; ++c)
    for (int w = 0; w < P.W; ++w) {
      const int t = P.c0 + w;
      const float want = (t >= 0 && t < T) ? h[((size_t)b * C + c) * T + t] : 0.f;
      if (r[(size_t)c * P.W + w] != want) ++bad;
    }
  printf("v%d: kernel ok, %ld mismatches of %d\n", v, bad, P.W * C);
  return 0;
}
