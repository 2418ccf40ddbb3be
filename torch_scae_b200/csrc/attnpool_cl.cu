// Attention pooling on a channels-last map (sm_100a): the pooling half of the part encoder's capsule head when its 1x1
// attention convolution runs as ONE GEMM over the B*S positions (reference part_encoder.py:95-101 att_conv +
// nn_ext.multiple_attention_pooling_2d :76-101).
//
// cuDNN runs the 1x1 convolution 128 -> n*(D+1) on a 5x5 map as B tiny per-image products (0.9 ms forward + backward at
// B = 1024, 21 TFLOP/s fp32); as a GEMM on the (B*S, 128) matrix of positions it is a plain cuBLAS SGEMM.  That GEMM
// leaves y[B, S, n*(D+1)] -- positions major, channels last -- so the pooling kernels below read that layout instead of
// NCHW.  Same math as attnpool_fwd/bwd_kernel in support.cu: per (image, capsule) group a softmax over the S positions
// of the group's last channel pools its other D channels.
//
// One warp per group.  The group's S x (D+1) tile lies in rows of D+1 contiguous floats (96 B at the MNIST config): the
// warp stages it through shared memory with coalesced row reads -- row stride (D+1) | 1 floats, odd, so that the column
// reads of the softmax and t[s] phases are bank-conflict free --, computes from there and, in the backward, writes the
// gradient tile back row by row.  Deterministic, no atomics.
#include "common.cuh"

namespace scae {

constexpr int kPoolWarps = 8;
constexpr int kPoolFloats = 1408;   // shared-memory floats per warp (static: 8 x 1408 x 4 B = 44 KB)

__host__ __device__ inline int pool_row_stride(int G) { return G | 1; }
// floats a warp needs: tile [S][stride] | softmax weights [S] | logit gradients [S] | upstream gradient [D]
__host__ __device__ inline int pool_floats(int D, int S) { return S * pool_row_stride(D + 1) + 2 * S + D; }

// stage the group's tile: tile[s * stride + d] = y[(b * S + s) * Ctot + cap * G + d]; eight loads in flight per lane
__device__ __forceinline__ void pool_stage(const float* __restrict__ base, float* tile, int S, int G, int Ctot,
                                           int stride, int lane) {
  constexpr int kBatch = 8;
  const float inv_G = 1.0f / (float)G;
  const int n = S * G;
  for (int e0 = lane; e0 < n; e0 += 32 * kBatch) {
    float v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = e0 + 32 * u;
      const int s = (int)(((float)e + 0.5f) * inv_G), d = e - s * G;
      v[u] = e < n ? __ldg(base + (long)s * Ctot + d) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = e0 + 32 * u;
      const int s = (int)(((float)e + 0.5f) * inv_G), d = e - s * G;
      if (e < n) tile[s * stride + d] = v[u];
    }
  }
}

// softmax over the positions of the logit channel (column D of the tile); lane owns positions lane and lane + 32
__device__ __forceinline__ void pool_softmax(const float* tile, int S, int D, int stride, int lane, float& w0,
                                             float& w1) {
  const float l0 = lane < S ? tile[lane * stride + D] : -INFINITY;
  const float l1 = lane + 32 < S ? tile[(lane + 32) * stride + D] : -INFINITY;
  float mx = fmaxf(l0, l1);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  const float e0 = lane < S ? expf(l0 - mx) : 0.0f, e1 = lane + 32 < S ? expf(l1 - mx) : 0.0f;
  const float inv = 1.0f / warp_sum(e0 + e1);
  w0 = e0 * inv;
  w1 = e1 * inv;
}

// y[B, S, n * (D + 1)] -> out[B * n, D]
__global__ void __launch_bounds__(32 * kPoolWarps) attnpool_cl_fwd_kernel(const float* __restrict__ y,
                                                                          float* __restrict__ out, long groups, int n,
                                                                          int D, int S) {
  __shared__ float sm[kPoolWarps * kPoolFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = D + 1, Ctot = n * G, stride = pool_row_stride(G);
  float* tile = sm + warp * kPoolFloats;
  float* wts = tile + S * stride;
  for (long grp = (long)blockIdx.x * kPoolWarps + warp; grp < groups; grp += (long)gridDim.x * kPoolWarps) {
    const long b = grp / n;
    const int cap = (int)(grp - b * n);
    pool_stage(y + b * S * (long)Ctot + (long)cap * G, tile, S, G, Ctot, stride, lane);
    __syncwarp();
    float w0, w1;
    pool_softmax(tile, S, D, stride, lane, w0, w1);
    if (lane < S) wts[lane] = w0;
    if (lane + 32 < S) wts[lane + 32] = w1;
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      float acc = 0.0f;
      for (int s = 0; s < S; ++s) acc = fmaf(tile[s * stride + d], wts[s], acc);
      out[grp * D + d] = acc;
    }
    __syncwarp();   // the next group's staging overwrites the tile
  }
}

// gy[b, s, cap, d] = g[d] * w[s]  (d < D);  gy[b, s, cap, D] = w[s] * (t[s] - sum_s' w[s'] t[s']),  t[s] = sum_d g[d] y[s, d]
__global__ void __launch_bounds__(32 * kPoolWarps) attnpool_cl_bwd_kernel(const float* __restrict__ y,
                                                                          const float* __restrict__ g,
                                                                          float* __restrict__ gy, long groups, int n,
                                                                          int D, int S) {
  __shared__ float sm[kPoolWarps * kPoolFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = D + 1, Ctot = n * G, stride = pool_row_stride(G);
  const float inv_G = 1.0f / (float)G;
  float* tile = sm + warp * kPoolFloats;
  float* wts = tile + S * stride;
  float* glog = wts + S;
  float* gd = glog + S;
  for (long grp = (long)blockIdx.x * kPoolWarps + warp; grp < groups; grp += (long)gridDim.x * kPoolWarps) {
    const long b = grp / n;
    const int cap = (int)(grp - b * n);
    const long off = b * S * (long)Ctot + (long)cap * G;
    pool_stage(y + off, tile, S, G, Ctot, stride, lane);
    for (int d = lane; d < D; d += 32) gd[d] = __ldg(g + grp * D + d);
    __syncwarp();
    float w0, w1;
    pool_softmax(tile, S, D, stride, lane, w0, w1);
    float t0 = 0.0f, t1 = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float gv = gd[d];
      if (lane < S) t0 = fmaf(gv, tile[lane * stride + d], t0);
      if (lane + 32 < S) t1 = fmaf(gv, tile[(lane + 32) * stride + d], t1);
    }
    const float dot = warp_sum(w0 * t0 + w1 * t1);
    if (lane < S) {
      wts[lane] = w0;
      glog[lane] = w0 * (t0 - dot);
    }
    if (lane + 32 < S) {
      wts[lane + 32] = w1;
      glog[lane + 32] = w1 * (t1 - dot);
    }
    __syncwarp();
    float* gbase = gy + off;
    for (int e = lane; e < S * G; e += 32) {
      const int s = (int)(((float)e + 0.5f) * inv_G), d = e - s * G;
      gbase[(long)s * Ctot + d] = d < D ? gd[d] * wts[s] : glog[s];
    }
    __syncwarp();
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)
static bool pool_cl_shape_ok(long groups, int n, int D, int S) {
  return groups > 0 && n > 0 && groups % n == 0 && D > 0 && S > 0 && S <= 64 && (long)S * (D + 1) < (1L << 22) &&
         pool_floats(D, S) <= kPoolFloats;
}

static int pool_cl_grid(long groups) {
  long grid = (groups + kPoolWarps - 1) / kPoolWarps;
  const long cap = 8L * sm_count();
  return (int)(grid > cap ? cap : grid);
}

}  // namespace scae

#define SCAE_EXPORT __attribute__((visibility("default")))
extern "C" {

SCAE_EXPORT int scae_attnpool_cl_supported(long groups, int n, int D, int S) {
  return scae::pool_cl_shape_ok(groups, n, D, S) ? 1 : 0;
}

SCAE_EXPORT int scae_attnpool_cl_fwd(const float* y, float* out, long groups, int n, int D, int S,
                                     scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(y && out, SCAE_EINVAL, "attnpool_cl fwd: y and out are required");
  SCAE_REQUIRE(pool_cl_shape_ok(groups, n, D, S), SCAE_ELIMIT,
               "attnpool_cl: groups=%ld n=%d D=%d S=%d (S <= 64, S*((D+1)|1) + 2S + D <= %d floats)", groups, n, D, S,
               kPoolFloats);
  attnpool_cl_fwd_kernel<<<pool_cl_grid(groups), 32 * kPoolWarps, 0, static_cast<cudaStream_t>(stream_)>>>(
      y, out, groups, n, D, S);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_attnpool_cl_bwd(const float* y, const float* g, float* gy, long groups, int n, int D, int S,
                                     scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(y && g && gy, SCAE_EINVAL, "attnpool_cl bwd: a required pointer is NULL");
  SCAE_REQUIRE(pool_cl_shape_ok(groups, n, D, S), SCAE_ELIMIT,
               "attnpool_cl: groups=%ld n=%d D=%d S=%d (S <= 64, S*((D+1)|1) + 2S + D <= %d floats)", groups, n, D, S,
               kPoolFloats);
  attnpool_cl_bwd_kernel<<<pool_cl_grid(groups), 32 * kPoolWarps, 0, static_cast<cudaStream_t>(stream_)>>>(
      y, g, gy, groups, n, D, S);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

}  // extern "C"
