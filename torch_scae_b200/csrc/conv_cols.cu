// im2col / col2im for the part encoder's unpadded 3x3 convolutions (sm_100a), so that their passes can run as plain
// cuBLAS SGEMMs on the matrix of output positions (reference nn_ext.py:34-59 Conv2dStack via part_encoder.py:26-40).
//
// Why: measured on B200 in strict fp32 (tools/conv_gemm_probe.py, B = 1024) cuDNN's data-gradient engine runs these
// layers at 24-25 TFLOP/s and its stride-2 forward / weight gradient at ~30, while cuBLAS SIMT SGEMMs on the same
// shapes reach 45-60.  The GEMM formulation needs the activations as rows = (image, output position), columns =
// (channel, ky, kx) -- the order of weight.view(C_out, C_in * 9) -- which the two kernels below produce / consume
// directly from / to NCHW; ATen's unfold / fold take 8 ms / 3 ms for the same job.
//
//   im2col3x3: x[B, C, H, W] -> cols[B * L, C * 9]      L = Ho * Wo, Ho = (H - 3) / stride + 1
//   col2im3x3: dcols[B * L, C * 9] -> dx[B, C, H, W]    (the adjoint: every input pixel sums its <= 9 taps, fixed order)
//
// A CTA owns one image and a group of 32 channels.  Both directions go through shared memory so that global reads and
// writes are contiguous: NCHW planes of a channel group are one contiguous run of 32 * H * W floats, and the group's
// slice of a cols row is 288 contiguous floats.  No atomics; deterministic.
#include "common.cuh"

namespace scae {

constexpr int kColsThreads = 256;
constexpr int kColsGroup = 32;   // channels per CTA

// e / d for 0 <= e < 2^22 with inv = 1 / (float)d
__device__ __forceinline__ int cols_div(int e, float inv) { return (int)(((float)e + 0.5f) * inv); }

__global__ void __launch_bounds__(kColsThreads) im2col3x3_kernel(const float* __restrict__ x, float* __restrict__ cols,
                                                                 int C, int H, int W, int Ho, int Wo, int stride) {
  SCAE_DYNAMIC_SMEM(sm);   // [cg][H * W]: the channel group's planes of this image
  const int tid = threadIdx.x, T = blockDim.x;
  const int b = blockIdx.y, c0 = blockIdx.x * kColsGroup;
  const int cg = min(kColsGroup, C - c0);
  const int HW = H * W, L = Ho * Wo, rowlen = cg * 9;
  const float* src = x + ((long)b * C + c0) * HW;
  constexpr int kBatch = 8;
  for (int i0 = tid; i0 < cg * HW; i0 += T * kBatch) {
    float v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) v[u] = i0 + u * T < cg * HW ? __ldg(src + i0 + u * T) : 0.0f;
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
      if (i0 + u * T < cg * HW) sm[i0 + u * T] = v[u];
  }
  __syncthreads();
  float* dst = cols + (long)b * L * ((long)C * 9) + (long)c0 * 9;
  const float inv_row = 1.0f / (float)rowlen, inv_wo = 1.0f / (float)Wo;
  for (int e = tid; e < L * rowlen; e += T) {
    const int l = cols_div(e, inv_row), j = e - l * rowlen;
    const int cl = j / 9, k = j - cl * 9;
    const int ky = k / 3, kx = k - ky * 3;
    const int oy = cols_div(l, inv_wo), ox = l - oy * Wo;
    dst[(long)l * ((long)C * 9) + j] = sm[cl * HW + (oy * stride + ky) * W + ox * stride + kx];
  }
}

// kStride and the channel group are compile-time: a run-time stride costs six integer divisions per input pixel (the
// first version spent 1.2 ms per train step there), and a group of 16 halves the tile of the 19x19 layer to 47 KB so
// that four CTAs fit an SM.
template <int kStride, int kGroup>
__global__ void __launch_bounds__(kColsThreads) col2im3x3_kernel(const float* __restrict__ dcols, float* __restrict__ dx,
                                                                 int C, int H, int W, int Ho, int Wo) {
  SCAE_DYNAMIC_SMEM(sm);   // [L][rowpad]: the channel group's slice of this image's rows; odd row stride: the gather
                           // below walks consecutive rows with consecutive lanes
  const int tid = threadIdx.x, T = blockDim.x;
  const int b = blockIdx.y, c0 = blockIdx.x * kGroup;
  const int cg = min(kGroup, C - c0);
  const int HW = H * W, L = Ho * Wo, rowlen = cg * 9, rowpad = rowlen | 1;
  const float* src = dcols + (long)b * L * ((long)C * 9) + (long)c0 * 9;
  const float inv_row = 1.0f / (float)rowlen;
  // eight loads in flight per thread: with one, the staging loop is a chain of ~45 dependent HBM round trips per CTA
  // (0.5 ms for the 19x19 layer at B = 1024, profiles/r01n)
  constexpr int kBatch = 8;
  const int n = L * rowlen;
  for (int e0 = tid; e0 < n; e0 += T * kBatch) {
    float v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = e0 + u * T;
      const int l = cols_div(e, inv_row), j = e - l * rowlen;
      v[u] = e < n ? __ldg(src + (long)l * ((long)C * 9) + j) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = e0 + u * T;
      const int l = cols_div(e, inv_row), j = e - l * rowlen;
      if (e < n) sm[l * rowpad + j] = v[u];
    }
  }
  __syncthreads();
  float* dst = dx + ((long)b * C + c0) * HW;
  const float inv_hw = 1.0f / (float)HW, inv_w = 1.0f / (float)W;
  for (int i = tid; i < cg * HW; i += T) {
    const int cl = cols_div(i, inv_hw), p = i - cl * HW;
    const int iy = cols_div(p, inv_w), ix = p - iy * W;
    const float* tap = sm + cl * 9;
    // branch-free: an invalid tap reads tap[0] and adds 0.  (With `continue`s the nine taps became nine divergent
    // branches -- the lanes of a warp differ in parity and border position -- each re-deriving the shared-memory window
    // address: 0.45 ms for the 19x19 layer at B = 1024, profiles/r01o.)
    float acc = 0.0f;
    if (kStride == 2) {
      // only taps of the pixel's own parity can hit: ky = (iy & 1) + 2 jy, oy = (iy >> 1) - jy -- four candidates, not nine
      const int py = iy & 1, px = ix & 1, ay = iy >> 1, ax = ix >> 1;
#pragma unroll
      for (int jy = 0; jy < 2; ++jy) {
        const int ky = py + 2 * jy, oy = ay - jy;
        const bool vy = ky < 3 && oy >= 0 && oy < Ho;
#pragma unroll
        for (int jx = 0; jx < 2; ++jx) {
          const int kx = px + 2 * jx, ox = ax - jx;
          const bool ok = vy && kx < 3 && ox >= 0 && ox < Wo;
          const float t = tap[ok ? (oy * Wo + ox) * rowpad + ky * 3 + kx : 0];
          acc += ok ? t : 0.0f;
        }
      }
    } else {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int oy = iy - ky;
        const bool vy = oy >= 0 && oy < Ho;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ox = ix - kx;
          const bool ok = vy && ox >= 0 && ox < Wo;
          const float t = tap[ok ? (oy * Wo + ox) * rowpad + ky * 3 + kx : 0];
          acc += ok ? t : 0.0f;
        }
      }
    }
    dst[i] = acc;
  }
}

// Batched 2-D transpose through a padded shared-memory tile: in[batch][R][Cc] -> out[batch][Cc][R].  NCHW <-> rows
// (R = channels, Cc = positions or the other way round) for the GEMM operands and results; ATen's generic strided copy
// does these permutes at about a third of the HBM rate.  Grid (tiles, batch); 256 threads = a 32 x 8 patch.
__global__ void __launch_bounds__(256) transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                int R, int Cc, int tiles_x) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int r0 = tile_y * 32, c0 = tile_x * 32;
  const float* src = in + (long)blockIdx.y * R * Cc;
  float* dst = out + (long)blockIdx.y * R * Cc;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    if (r < R && c < Cc) tile[j][tx] = __ldg(src + (long)r * Cc + c);
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < Cc && r < R) dst[(long)c * R + r] = tile[tx][j];
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)
static size_t im2col_smem(int C, int H, int W) {
  return (size_t)(C < kColsGroup ? C : kColsGroup) * H * W * sizeof(float);
}
static int col2im_group(int C, int H, int W, int stride) {
  const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
  return (size_t)Ho * Wo * ((kColsGroup * 9) | 1) * sizeof(float) > 64 * 1024 || C <= 16 ? 16 : kColsGroup;
}
static size_t col2im_smem(int C, int H, int W, int stride) {
  const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
  const int group = col2im_group(C, H, W, stride);
  return (size_t)Ho * Wo * (((C < group ? C : group) * 9) | 1) * sizeof(float);
}
static bool cols_shape_ok(int B, int C, int H, int W, int stride) {
  if (B <= 0 || B > 65535 || C <= 0 || H < 3 || W < 3 || stride < 1 || stride > 2) return false;
  if ((long)kColsGroup * H * W >= (1L << 22)) return false;                     // cols_div range
  const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
  if ((long)Ho * Wo * kColsGroup * 9 >= (1L << 22)) return false;
  const size_t lim = (size_t)max_smem_optin();
  return im2col_smem(C, H, W) <= lim && col2im_smem(C, H, W, stride) <= lim;
}

}  // namespace scae

#define SCAE_EXPORT __attribute__((visibility("default")))
extern "C" {

SCAE_EXPORT int scae_conv_cols_supported(int B, int C, int H, int W, int stride) {
  return scae::cols_shape_ok(B, C, H, W, stride) ? 1 : 0;
}

SCAE_EXPORT int scae_im2col3x3(const float* x, float* cols, int B, int C, int H, int W, int stride,
                               scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(x && cols, SCAE_EINVAL, "im2col3x3: x and cols are required");
  SCAE_REQUIRE(cols_shape_ok(B, C, H, W, stride), SCAE_ELIMIT, "im2col3x3: unsupported shape B=%d C=%d H=%d W=%d s=%d",
               B, C, H, W, stride);
  const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
  const size_t smem = im2col_smem(C, H, W);
  SCAE_CUDA_TRY(cudaFuncSetAttribute(im2col3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((C + kColsGroup - 1) / kColsGroup, B);
  im2col3x3_kernel<<<grid, kColsThreads, smem, static_cast<cudaStream_t>(stream_)>>>(x, cols, C, H, W, Ho, Wo, stride);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_col2im3x3(const float* dcols, float* dx, int B, int C, int H, int W, int stride,
                               scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(dcols && dx, SCAE_EINVAL, "col2im3x3: dcols and dx are required");
  SCAE_REQUIRE(cols_shape_ok(B, C, H, W, stride), SCAE_ELIMIT, "col2im3x3: unsupported shape B=%d C=%d H=%d W=%d s=%d",
               B, C, H, W, stride);
  const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
  const size_t smem = col2im_smem(C, H, W, stride);
  const int group = col2im_group(C, H, W, stride);
  auto kern = stride == 2 ? (group == 16 ? col2im3x3_kernel<2, 16> : col2im3x3_kernel<2, kColsGroup>)
                          : (group == 16 ? col2im3x3_kernel<1, 16> : col2im3x3_kernel<1, kColsGroup>);
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((C + group - 1) / group, B);
  kern<<<grid, kColsThreads, smem, static_cast<cudaStream_t>(stream_)>>>(dcols, dx, C, H, W, Ho, Wo);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_transpose_batched(const float* in, float* out, int batch, int R, int Cc, scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(in && out, SCAE_EINVAL, "transpose_batched: in and out are required");
  SCAE_REQUIRE(batch > 0 && batch <= 65535 && R > 0 && Cc > 0, SCAE_ELIMIT, "transpose_batched: bad shape %d x %d x %d",
               batch, R, Cc);
  const int tiles_x = (Cc + 31) / 32, tiles_y = (R + 31) / 32;
  dim3 grid(tiles_x * tiles_y, batch);
  transpose_batched_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(in, out, R, Cc, tiles_x);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

}  // extern "C"
