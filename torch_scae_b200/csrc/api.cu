// libscae_b200: error reporting, device attribute cache and the shared row-reduction kernel.
#include <stdarg.h>

#include <atomic>
#include <stdio.h>

#include "common.cuh"

namespace scae {

static thread_local char g_error[512] = "";
static thread_local unsigned long long g_launches = 0;

static std::atomic<unsigned long long> g_fast_path{0};   // process-wide: backward calls come from autograd's thread

void note_launch() { ++g_launches; }
void note_fast_path() { g_fast_path.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<unsigned long long> g_persistent_path{0};
void note_persistent_path() { g_persistent_path.fetch_add(1, std::memory_order_relaxed); }

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
    for (int p = w; p < n_parts; p += 8 * kReduceWarps) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // predicated, so the last round also keeps all of its loads in flight
        const int q = p + k * kReduceWarps;
        acc[k] += q < n_parts ? partials[(size_t)q * n + i] : 0.0f;
      }
    }
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
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

// ---- column sums of a tall-skinny matrix --------------------------------------------------------------------------
// x [rows, cols] row-major -> partial[split][cols]; a second reduce_rows pass sums the splits (fixed order).
// A CTA of 256 threads owns a tile of W = min(cols, 256) columns and a slab of rows; thread t sits on column t % W and
// walks rows t / W, t / W + 256 / W, ...: for cols <= 256 the CTA reads its slab as one contiguous stream.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ partial, long rows,
                                                     int cols, int W, long rows_per_split) {
  __shared__ float part[256];
  const int t = threadIdx.x;
  const int rpi = 256 / W;                              // rows per iteration
  const int col = blockIdx.x * W + (t % W);
  const long r0 = (long)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long r = r0 + t / W; r < r1; r += 8L * rpi) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long q = r + (long)k * rpi;
      acc[k] += q < r1 ? __ldg(x + q * cols + col) : 0.0f;
    }
  }
  part[t] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (t < W) {
    float s = 0.f;
    for (int k = 0; k < rpi; ++k) s += part[t + k * W];
    partial[(size_t)blockIdx.y * cols + col] = s;
  }
}

static bool colsum_shape_ok(long rows, int cols) {
  return rows > 0 && cols > 0 && ((cols <= 256 && 256 % cols == 0) || cols % 256 == 0);
}
static int colsum_splits(long rows, int cols) {
  const int tiles = cols <= 256 ? 1 : cols / 256;
  const int rpi = cols <= 256 ? 256 / cols : 1;
  long s = (4L * sm_count() + tiles - 1) / tiles;       // about four CTAs per SM
  const long max_s = (rows + 8L * rpi - 1) / (8L * rpi); // at least one full round of loads per CTA
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace scae

#define SCAE_EXPORT __attribute__((visibility("default")))
extern "C" {

SCAE_EXPORT int scae_abi_version(void) { return SCAE_B200_ABI_VERSION; }

SCAE_EXPORT const char* scae_last_error(void) { return scae::g_error; }

SCAE_EXPORT const char* scae_build_arch(void) { return "sm_100a"; }

#ifndef SCAE_BUILD_ID
#define SCAE_BUILD_ID "unknown"
#endif
SCAE_EXPORT const char* scae_build_id(void) { return SCAE_BUILD_ID; }

SCAE_EXPORT unsigned long long scae_launch_count(void) { return scae::g_launches; }

SCAE_EXPORT unsigned long long scae_caps_fast_path_count(void) { return scae::g_fast_path.load(); }
SCAE_EXPORT unsigned long long scae_caps_persistent_path_count(void) { return scae::g_persistent_path.load(); }

SCAE_EXPORT size_t scae_colsum_workspace_bytes(long rows, int cols) {
  if (!scae::colsum_shape_ok(rows, cols)) return 0;
  return (size_t)scae::colsum_splits(rows, cols) * cols * sizeof(float);
}

SCAE_EXPORT int scae_colsum(const float* x, long rows, int cols, float* out, void* workspace, size_t workspace_bytes,
                            scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(x && out, SCAE_EINVAL, "colsum: x and out are required");
  SCAE_REQUIRE(colsum_shape_ok(rows, cols), SCAE_ELIMIT,
               "colsum: cols=%d must divide 256 or be a multiple of 256 (rows=%ld)", cols, rows);
  const int splits = colsum_splits(rows, cols);
  SCAE_REQUIRE(workspace && workspace_bytes >= (size_t)splits * cols * sizeof(float), SCAE_EINVAL,
               "colsum: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int W = cols <= 256 ? cols : 256;
  const long rows_per_split = (rows + splits - 1) / splits;
  dim3 grid(cols <= 256 ? 1 : cols / 256, splits);
  colsum_kernel<<<grid, 256, 0, stream>>>(x, static_cast<float*>(workspace), rows, cols, W, rows_per_split);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return launch_reduce_rows(static_cast<const float*>(workspace), out, splits, cols, stream);
}

}  // extern "C"
