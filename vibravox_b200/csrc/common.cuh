// Shared host-side plumbing of libvbx_b200: error text, launch counter, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/vbx.h"

namespace vbx {
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}
inline int launched(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", name, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
}  // namespace vbx

#define VBX_REQUIRE(cond, code, msg) \
  do { if (!(cond)) return vbx::fail(code, msg); } while (0)
