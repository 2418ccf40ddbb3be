// libscae_b200: error reporting, device attribute cache and the shared row-reduction kernel.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace scae {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return SCAE_ECUDA;
}

static int cached_attr(cudaDeviceAttr attr, int* cache /* [64] */) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess) return 0;
    cache[dev] = v;
  }
  return cache[dev];
}

int sm_count() {
  static int cache[64] = {0};
  const int n = cached_attr(cudaDevAttrMultiProcessorCount, cache);
  return n > 0 ? n : 148;
}

int max_smem_optin() {
  static int cache[64] = {0};
  const int n = cached_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, cache);
  return n > 0 ? n : 48 * 1024;
}

// out[i] = sum_p partials[p][i] in a fixed order: deterministic.  A CTA owns 32 columns; its 16 warps take the rows
// p = w, w+16, w+32, ... with eight independent accumulators (eight loads in flight per thread: the kernel is pure
// latency), and the per-warp sums are combined in warp order.  Lanes are consecutive columns, so every row read is one
// 128-byte line per warp.
constexpr int kReduceWarps = 16;
__global__ void __launch_bounds__(32 * kReduceWarps) reduce_rows_kernel(const float* __restrict__ partials,
                                                                        float* __restrict__ out, int n_parts, int n) {
  __shared__ float part[kReduceWarps][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (i < n) {
    int p = w;
    for (; p + 7 * kReduceWarps < n_parts; p += 8 * kReduceWarps) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += partials[(size_t)(p + k * kReduceWarps) * n + i];
    }
    for (; p < n_parts; p += kReduceWarps) acc[0] += partials[(size_t)p * n + i];
  }
  part[w][lane] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (w == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kReduceWarps; ++k) t += part[k][lane];
    out[i] = t;
  }
}

int launch_reduce_rows(const float* partials, float* out, int n_parts, int n, cudaStream_t stream) {
  if (n <= 0) return SCAE_OK;
  reduce_rows_kernel<<<(n + 31) / 32, 32 * kReduceWarps, 0, stream>>>(partials, out, n_parts, n);
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

}  // namespace scae

#define SCAE_EXPORT __attribute__((visibility("default")))
extern "C" {

SCAE_EXPORT int scae_abi_version(void) { return SCAE_B200_ABI_VERSION; }

SCAE_EXPORT const char* scae_last_error(void) { return scae::g_error; }

SCAE_EXPORT const char* scae_build_arch(void) { return "sm_100a"; }

}  // extern "C"
