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

// out[i] = sum_p partials[p][i], p ascending: deterministic.  One thread per column, coalesced across columns.
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partials, float* __restrict__ out,
                                                          int n_parts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int p = 0;
  for (; p + 4 <= n_parts; p += 4) {
    acc0 += partials[(size_t)(p + 0) * n + i];
    acc1 += partials[(size_t)(p + 1) * n + i];
    acc2 += partials[(size_t)(p + 2) * n + i];
    acc3 += partials[(size_t)(p + 3) * n + i];
  }
  for (; p < n_parts; ++p) acc0 += partials[(size_t)p * n + i];
  out[i] = (acc0 + acc1) + (acc2 + acc3);
}

int launch_reduce_rows(const float* partials, float* out, int n_parts, int n, cudaStream_t stream) {
  if (n <= 0) return SCAE_OK;
  reduce_rows_kernel<<<(n + 255) / 256, 256, 0, stream>>>(partials, out, n_parts, n);
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
