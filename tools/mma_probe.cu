// Micro-probe: cycles per tcgen05.mma (kind::f16, M=128) for the shared-memory layouts used by tc_conv.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../vibravox_b200/csrc/tc_common.cuh"
using namespace vbx::tc;

__global__ void probe(int N, int iters, int a_mn, int sbo_a, int lbo_a, int lbo_b, int per_commit, int nacc, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 40000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if ((threadIdx.x & 31) == 0 && warp < nacc) {
    const uint32_t idesc = make_idesc_bf16(N, a_mn != 0, false);
    const uint32_t a = smem_u32(smem), b = a + 20000;
    long long t0 = clock64();
    uint32_t parity = 0;
    for (int it = 0; it < iters; it += per_commit) {
      for (int j = 0; j < per_commit; ++j) {
        uint64_t da = make_desc(a, lbo_a, sbo_a);
        uint64_t db = make_desc(b, lbo_b, 128);
        mma_bf16_ss(tm + (uint32_t)(warp * N), da, db, idesc, 1);
      }
      mma_commit(&bar[warp]);
      mbar_wait(&bar[warp], parity);
      parity ^= 1;
    }
    long long t1 = clock64();
    if (warp == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
  long long* out;
  cudaMalloc(&out, 1024 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct Cfg { int N, a_mn, sbo_a, lbo_a, lbo_b, per, nacc; const char* name; };
  Cfg cfgs[] = {   // nacc = number of issuing threads (one per warp), each with its own accumulator
      {256, 1, 144, 2304, 4096, 48, 1, "N=256 1 issuer"},
      {128, 1, 144, 2304, 2048, 48, 1, "N=128 1 issuer"},
      {128, 1, 144, 2304, 2048, 48, 2, "N=128 2 issuers"},
      {64, 1, 144, 2304, 1024, 48, 1, "N=64 1 issuer"},
      {64, 1, 144, 2304, 1024, 48, 2, "N=64 2 issuers"},
      {64, 1, 144, 2304, 1024, 48, 4, "N=64 4 issuers"},
      {32, 1, 144, 2304, 512, 48, 4, "N=32 4 issuers"},
      {32, 1, 144, 2304, 512, 48, 8, "N=32 8 issuers"},
  };


  for (int grid : {148}) {
    for (auto& c : cfgs) {
      const int iters = 4800;
      probe<<<grid, 256, 48 * 1024>>>(c.N, iters, c.a_mn, c.sbo_a, c.lbo_a, c.lbo_b, c.per, c.nacc, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[296];
      cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
      printf("grid %3d  %-45s : %.1f cycles / MMA per issuer, %.1f aggregate (ideal %d)\n", grid, c.name, avg / iters, avg / iters / c.nacc, c.N / 2);
    }
  }
  return 0;
}
