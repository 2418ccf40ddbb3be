// libscae_b200: small fused kernels for the CALLERS of the two hot paths -- the elementwise /
// reduction tail of the part encoder and of the set transformer that stock PyTorch runs as strings of tiny launches:
//
//   scae_layernorm_fwd / _bwd      LayerNorm over a short last dimension (d = 16 in the set transformer, reference
//                                  set_transformer.py:104-116): ATen's kernel spends 60 us on 40960 rows of 16 floats
//                                  (one CTA per row); here a thread owns a row -- 2.6 MB in, 2.6 MB out, ~5 us.
//   scae_bias_act_fwd / _bwd       per-channel bias (+ ReLU) of an NCHW convolution output in one pass, and in the
//                                  backward the ReLU mask fused with the bias gradient (reference nn_ext.py:34-59 via
//                                  nn.Conv2d + nn.ReLU: add_, clamp_min, threshold_backward, sum = four passes).
//   scae_attnpool_fwd / _bwd       multiple_attention_pooling_2d (reference nn_ext.py:76-101): per (image, capsule) a
//                                  softmax over the G*G positions of the group's last channel pools the other D
//                                  channels; one warp per group instead of mul + softmax + sum + their backward.
//
// All deterministic (fixed-order reductions), fp32, no atomics.  Numerics are checked against the plain PyTorch ops in
// tests/test_gpu_plumbing.py.
#include "common.cuh"

namespace scae {

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm over a last dimension D (multiple of 4, <= 64): thread per row
// ---------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, long rows,
                                                            float* __restrict__ y, float* __restrict__ stats) {
  __shared__ float gb[2 * D];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    gb[i] = gamma ? __ldg(gamma + i) : 1.0f;
    gb[D + i] = beta ? __ldg(beta + i) : 0.0f;
  }
  __syncthreads();
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    float v[D];
    const float4* src = reinterpret_cast<const float4*>(x + r * D);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 t = __ldg(src + q);
      v[4 * q] = t.x;
      v[4 * q + 1] = t.y;
      v[4 * q + 2] = t.z;
      v[4 * q + 3] = t.w;
    }
    float mean = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) mean += v[i];
    mean *= 1.0f / D;
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float d = v[i] - mean;
      var = fmaf(d, d, var);
    }
    const float rstd = rsqrtf(var * (1.0f / D) + eps);
    float4* dst = reinterpret_cast<float4*>(y + r * D);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      float4 t;
      t.x = fmaf((v[4 * q] - mean) * rstd, gb[4 * q], gb[D + 4 * q]);
      t.y = fmaf((v[4 * q + 1] - mean) * rstd, gb[4 * q + 1], gb[D + 4 * q + 1]);
      t.z = fmaf((v[4 * q + 2] - mean) * rstd, gb[4 * q + 2], gb[D + 4 * q + 2]);
      t.w = fmaf((v[4 * q + 3] - mean) * rstd, gb[4 * q + 3], gb[D + 4 * q + 3]);
      dst[q] = t;
    }
    if (stats) reinterpret_cast<float2*>(stats)[r] = make_float2(mean, rstd);
  }
}

// gx = rstd * (g*gamma - mean_D(g*gamma) - xhat * mean_D(g*gamma*xhat));  per-CTA partial sums of g*xhat and g
template <int D>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ stats, long rows,
                                                            float* __restrict__ gx, float* __restrict__ partials) {
  __shared__ float gam[D];
  __shared__ float red[8][2 * D + 1];
  for (int i = threadIdx.x; i < D; i += blockDim.x) gam[i] = gamma ? __ldg(gamma + i) : 1.0f;
  __syncthreads();
  float sg[D], sb[D];
#pragma unroll
  for (int i = 0; i < D; ++i) sg[i] = sb[i] = 0.0f;
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + r);
    const float4* gs = reinterpret_cast<const float4*>(g + r * D);
    const float4* xs = reinterpret_cast<const float4*>(x + r * D);
    float gv[D], xh[D];
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 a = __ldg(gs + q), b = __ldg(xs + q);
      gv[4 * q] = a.x;
      gv[4 * q + 1] = a.y;
      gv[4 * q + 2] = a.z;
      gv[4 * q + 3] = a.w;
      xh[4 * q] = (b.x - st.x) * st.y;
      xh[4 * q + 1] = (b.y - st.x) * st.y;
      xh[4 * q + 2] = (b.z - st.x) * st.y;
      xh[4 * q + 3] = (b.w - st.x) * st.y;
    }
    float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      sg[i] = fmaf(gv[i], xh[i], sg[i]);
      sb[i] += gv[i];
      gv[i] *= gam[i];
      m1 += gv[i];
      m2 = fmaf(gv[i], xh[i], m2);
    }
    m1 *= 1.0f / D;
    m2 *= 1.0f / D;
    float4* dst = reinterpret_cast<float4*>(gx + r * D);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      float4 t;
      t.x = st.y * (gv[4 * q] - m1 - xh[4 * q] * m2);
      t.y = st.y * (gv[4 * q + 1] - m1 - xh[4 * q + 1] * m2);
      t.z = st.y * (gv[4 * q + 2] - m1 - xh[4 * q + 2] * m2);
      t.w = st.y * (gv[4 * q + 3] - m1 - xh[4 * q + 3] * m2);
      dst[q] = t;
    }
  }
  // block reduction of the 2*D column sums: shuffle tree per warp, then the 8 warps in order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float a = warp_sum(sg[i]), b = warp_sum(sb[i]);
    if (lane == 0) {
      red[warp][i] = a;
      red[warp][D + i] = b;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][i];
    partials[(size_t)blockIdx.x * 2 * D + i] = t;
  }
}

static bool ln_dim_ok(int d) { return d == 8 || d == 16 || d == 32 || d == 64; }
static int ln_grid(long rows) {
  long g = (rows + 255) / 256;
  const long cap = 2L * sm_count();
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ---------------------------------------------------------------------------------------------------------------------
// per-channel bias (+ ReLU) on an NCHW tensor.  CTA (c, slab): the planes (n, c) of a slab of images
// ---------------------------------------------------------------------------------------------------------------------
// Loads in flight: 8 CTAs per SM x 256 threads x kBiasUnroll x (1 or 2 streams) x 4 B ~ 64-128 KB per SM; with 4 CTAs
// and 4 elements the kernels ran at half of the HBM rate (profiles/r01m).
constexpr int kBiasUnroll = 8;
__global__ void __launch_bounds__(256) bias_act_fwd_kernel(float* __restrict__ y, const float* __restrict__ bias, int N,
                                                           int C, int HW, int relu, int n_per_slab) {
  const int c = blockIdx.x;
  const float b = __ldg(bias + c);
  const int n0 = blockIdx.y * n_per_slab, n1 = min(N, n0 + n_per_slab);
  const long plane = (long)C * HW;
  const int total = (n1 - n0) * HW;
  const float inv_hw = 1.0f / (float)HW;
  // kBiasUnroll independent elements per thread and iteration: enough loads in flight to cover the HBM latency
  for (int e0 = threadIdx.x; e0 < total; e0 += kBiasUnroll * blockDim.x) {
    float* p[kBiasUnroll];
    float v[kBiasUnroll];
#pragma unroll
    for (int u = 0; u < kBiasUnroll; ++u) {
      const int e = e0 + u * blockDim.x;
      const int dn = (int)(((float)e + 0.5f) * inv_hw), s = e - dn * HW;
      p[u] = y + (long)(n0 + dn) * plane + (long)c * HW + s;
      v[u] = e < total ? *p[u] : 0.0f;
    }
    asm volatile("" ::: "memory");   // all loads of the batch are issued before the first store (keeps them in flight)
#pragma unroll
    for (int u = 0; u < kBiasUnroll; ++u) {
      const float t = v[u] + b;
      if (e0 + u * (int)blockDim.x < total) *p[u] = relu ? fmaxf(t, 0.0f) : t;
    }
  }
}

// gx = relu ? g * (y > 0) : g (may alias g);  partial[slab][c] = sum of gx over the slab's planes of channel c
__global__ void __launch_bounds__(256) bias_act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                           float* __restrict__ gx, float* __restrict__ partial, int N,
                                                           int C, int HW, int relu, int n_per_slab) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const int n0 = blockIdx.y * n_per_slab, n1 = min(N, n0 + n_per_slab);
  const long plane = (long)C * HW;
  const int total = (n1 - n0) * HW;
  const float inv_hw = 1.0f / (float)HW;
  float acc = 0.0f;
  for (int e0 = threadIdx.x; e0 < total; e0 += kBiasUnroll * blockDim.x) {
    long off[kBiasUnroll];
    float v[kBiasUnroll], yv[kBiasUnroll];
#pragma unroll
    for (int u = 0; u < kBiasUnroll; ++u) {
      const int e = e0 + u * blockDim.x;
      const int dn = (int)(((float)e + 0.5f) * inv_hw), s = e - dn * HW;
      off[u] = (long)(n0 + dn) * plane + (long)c * HW + s;
      const bool ok = e < total;
      v[u] = ok ? g[off[u]] : 0.0f;
      yv[u] = (ok && relu) ? y[off[u]] : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < kBiasUnroll; ++u) {
      const float t = yv[u] > 0.0f ? v[u] : 0.0f;
      if (relu && e0 + u * (int)blockDim.x < total) gx[off[u]] = t;
      acc += t;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[(size_t)blockIdx.y * C + c] = t;
  }
}

static int bias_slabs(int N, int C, int HW) {
  // about 8 CTAs per SM over (C x slabs), at least ~2048 elements per CTA, HW * n_per_slab < 2^22 (fast division)
  long slabs = (8L * sm_count() + C - 1) / C;
  const long max_by_work = ((long)N * HW + 2047) / 2048;
  if (slabs > max_by_work) slabs = max_by_work;
  if (slabs > N) slabs = N;
  if (slabs < 1) slabs = 1;
  while ((long)((N + slabs - 1) / slabs) * HW >= (1L << 22)) ++slabs;
  return (int)slabs;
}

// ---------------------------------------------------------------------------------------------------------------------
// multiple attention pooling: one warp per (image, capsule) group of (D + 1) x S floats
// ---------------------------------------------------------------------------------------------------------------------
// out[grp, d] = sum_s h[grp, d, s] * softmax_s(h[grp, D, s]);  S <= 64 positions, any D
__global__ void __launch_bounds__(256) attnpool_fwd_kernel(const float* __restrict__ h, float* __restrict__ out,
                                                           long groups, int D, int S) {
  extern __shared__ float sm[];                 // [warps][S] softmax weights
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* wts = sm + warp * S;
  for (long grp = (long)blockIdx.x * nw + warp; grp < groups; grp += (long)gridDim.x * nw) {
    const float* base = h + grp * (long)(D + 1) * S;
    const float* lg = base + (long)D * S;
    const float l0 = lane < S ? __ldg(lg + lane) : -INFINITY;
    const float l1 = lane + 32 < S ? __ldg(lg + lane + 32) : -INFINITY;
    float mx = fmaxf(l0, l1);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    const float e0 = lane < S ? expf(l0 - mx) : 0.0f, e1 = lane + 32 < S ? expf(l1 - mx) : 0.0f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    __syncwarp();
    if (lane < S) wts[lane] = e0 * inv;
    if (lane + 32 < S) wts[lane + 32] = e1 * inv;
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      const float* row = base + (long)d * S;
      float acc = 0.0f;
      for (int s = 0; s < S; ++s) acc = fmaf(__ldg(row + s), wts[s], acc);
      out[grp * D + d] = acc;
    }
  }
}

// gh[grp, d, s] = g[grp, d] * w[s]  (d < D);  gh[grp, D, s] = w[s] * (t[s] - sum_s' w[s'] t[s']),  t[s] = sum_d g[d] h[d, s]
__global__ void __launch_bounds__(256) attnpool_bwd_kernel(const float* __restrict__ h, const float* __restrict__ g,
                                                           float* __restrict__ gh, long groups, int D, int S) {
  extern __shared__ float sm[];                 // [warps][S + D]: softmax weights, upstream gradient
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* wts = sm + warp * (S + D);
  float* gd = wts + S;
  for (long grp = (long)blockIdx.x * nw + warp; grp < groups; grp += (long)gridDim.x * nw) {
    const float* base = h + grp * (long)(D + 1) * S;
    float* gbase = gh + grp * (long)(D + 1) * S;
    const float* lg = base + (long)D * S;
    const float l0 = lane < S ? __ldg(lg + lane) : -INFINITY;
    const float l1 = lane + 32 < S ? __ldg(lg + lane + 32) : -INFINITY;
    float mx = fmaxf(l0, l1);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    const float e0 = lane < S ? expf(l0 - mx) : 0.0f, e1 = lane + 32 < S ? expf(l1 - mx) : 0.0f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    const float w0 = e0 * inv, w1 = e1 * inv;
    __syncwarp();
    if (lane < S) wts[lane] = w0;
    if (lane + 32 < S) wts[lane + 32] = w1;
    for (int d = lane; d < D; d += 32) gd[d] = __ldg(g + grp * D + d);
    __syncwarp();
    // lanes over positions: t[s] and the logit-row gradient
    float t0 = 0.0f, t1 = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float gv = gd[d];
      if (lane < S) t0 = fmaf(gv, __ldg(base + (long)d * S + lane), t0);
      if (lane + 32 < S) t1 = fmaf(gv, __ldg(base + (long)d * S + lane + 32), t1);
    }
    const float dot = warp_sum(w0 * t0 + w1 * t1);
    if (lane < S) gbase[(long)D * S + lane] = w0 * (t0 - dot);
    if (lane + 32 < S) gbase[(long)D * S + lane + 32] = w1 * (t1 - dot);
    // the D pooled channels: flat, coalesced
    const int n = D * S;
    const float inv_s = 1.0f / (float)S;
    for (int e = lane; e < n; e += 32) {
      const int d = (int)(((float)e + 0.5f) * inv_s), s = e - d * S;
      gbase[e] = gd[d] * wts[s];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// cv_ops.geometric_transform (nonlinear=True, as_matrix=False) on rows of 6 pose parameters: thread per row
// ---------------------------------------------------------------------------------------------------------------------
template <bool kSim>
__global__ void __launch_bounds__(256) pose_transform_fwd_kernel(const float* __restrict__ t, float* __restrict__ out,
                                                                 long rows) {
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    float v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = __ldg(t + r * 6 + k);
    PoseAffine o;
    pose_affine_fwd<kSim>(v, o);
#pragma unroll
    for (int k = 0; k < 6; ++k) out[r * 6 + k] = o.a[k];
  }
}
template <bool kSim>
__global__ void __launch_bounds__(256) pose_transform_bwd_kernel(const float* __restrict__ t, const float* __restrict__ g,
                                                                 float* __restrict__ gt, long rows) {
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long)gridDim.x * blockDim.x) {
    float v[6], ga[6], gv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      v[k] = __ldg(t + r * 6 + k);
      ga[k] = __ldg(g + r * 6 + k);
    }
    PoseAffine o;
    pose_affine_fwd<kSim>(v, o);
    pose_affine_bwd<kSim>(ga, o, gv);
#pragma unroll
    for (int k = 0; k < 6; ++k) gt[r * 6 + k] = gv[k];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// RMSprop with momentum over flat parameter / gradient / state buffers: one pass instead of seven foreach launches
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ sq, float* __restrict__ buf, long n, float lr,
                                                      float alpha, float eps, float momentum) {
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 sv = *reinterpret_cast<float4*>(sq + i), pv = *reinterpret_cast<float4*>(p + i);
      float4 bv = buf ? *reinterpret_cast<float4*>(buf + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
      float ss[4] = {sv.x, sv.y, sv.z, sv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (gg[k] != gg[k]) continue;   // NaN = "no gradient this step": parameter and state untouched (torch.optim skips p.grad None)
        ss[k] = fmaf(1.0f - alpha, gg[k] * gg[k], ss[k] * alpha);
        const float step = __fdiv_rn(gg[k], sqrtf(ss[k]) + eps);
        if (buf) {
          bb[k] = fmaf(bb[k], momentum, step);
          pp[k] = fmaf(-lr, bb[k], pp[k]);
        } else {
          pp[k] = fmaf(-lr, step, pp[k]);
        }
      }
      *reinterpret_cast<float4*>(sq + i) = make_float4(ss[0], ss[1], ss[2], ss[3]);
      *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      if (buf) *reinterpret_cast<float4*>(buf + i) = make_float4(bb[0], bb[1], bb[2], bb[3]);
    } else {
      for (long j = i; j < n; ++j) {
        const float gj = g[j];
        if (gj != gj) continue;
        const float s = fmaf(1.0f - alpha, gj * gj, sq[j] * alpha);
        sq[j] = s;
        const float step = __fdiv_rn(gj, sqrtf(s) + eps);
        if (buf) {
          const float b = fmaf(buf[j], momentum, step);
          buf[j] = b;
          p[j] = fmaf(-lr, b, p[j]);
        } else {
          p[j] = fmaf(-lr, step, p[j]);
        }
      }
    }
  }
}

}  // namespace scae

using namespace scae;
#define SCAE_EXPORT __attribute__((visibility("default")))

extern "C" {

SCAE_EXPORT size_t scae_layernorm_bwd_workspace_bytes(long rows, int d) {
  if (!ln_dim_ok(d) || rows <= 0) return 0;
  return (size_t)ln_grid(rows) * 2 * d * sizeof(float);
}

SCAE_EXPORT int scae_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, long rows, int d,
                                   float* y, float* stats, scae_stream_t stream_) {
  SCAE_REQUIRE(x && y, SCAE_EINVAL, "layernorm fwd: x and y are required");
  SCAE_REQUIRE(rows > 0 && ln_dim_ok(d), SCAE_ELIMIT, "layernorm: last dimension %d not supported (8, 16, 32, 64)", d);
  SCAE_REQUIRE(aligned16(x) && aligned16(y), SCAE_EINVAL, "layernorm: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = ln_grid(rows);
  switch (d) {
    case 8: layernorm_fwd_kernel<8><<<grid, 256, 0, stream>>>(x, gamma, beta, eps, rows, y, stats); break;
    case 16: layernorm_fwd_kernel<16><<<grid, 256, 0, stream>>>(x, gamma, beta, eps, rows, y, stats); break;
    case 32: layernorm_fwd_kernel<32><<<grid, 256, 0, stream>>>(x, gamma, beta, eps, rows, y, stats); break;
    default: layernorm_fwd_kernel<64><<<grid, 256, 0, stream>>>(x, gamma, beta, eps, rows, y, stats); break;
  }
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_layernorm_bwd(const float* g, const float* x, const float* gamma, const float* stats, long rows,
                                   int d, float* gx, float* g_gamma_beta, void* workspace, size_t workspace_bytes,
                                   scae_stream_t stream_) {
  SCAE_REQUIRE(g && x && stats && gx && g_gamma_beta, SCAE_EINVAL, "layernorm bwd: a required pointer is NULL");
  SCAE_REQUIRE(rows > 0 && ln_dim_ok(d), SCAE_ELIMIT, "layernorm: last dimension %d not supported (8, 16, 32, 64)", d);
  SCAE_REQUIRE(aligned16(g) && aligned16(x) && aligned16(gx), SCAE_EINVAL, "layernorm: pointers must be 16-byte aligned");
  const int grid = ln_grid(rows);
  SCAE_REQUIRE(workspace && workspace_bytes >= (size_t)grid * 2 * d * sizeof(float), SCAE_EINVAL,
               "layernorm bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* partials = static_cast<float*>(workspace);
  switch (d) {
    case 8: layernorm_bwd_kernel<8><<<grid, 256, 0, stream>>>(g, x, gamma, stats, rows, gx, partials); break;
    case 16: layernorm_bwd_kernel<16><<<grid, 256, 0, stream>>>(g, x, gamma, stats, rows, gx, partials); break;
    case 32: layernorm_bwd_kernel<32><<<grid, 256, 0, stream>>>(g, x, gamma, stats, rows, gx, partials); break;
    default: layernorm_bwd_kernel<64><<<grid, 256, 0, stream>>>(g, x, gamma, stats, rows, gx, partials); break;
  }
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return launch_reduce_rows(partials, g_gamma_beta, grid, 2 * d, stream);   // [g_gamma(d) | g_beta(d)]
}

SCAE_EXPORT int scae_bias_act_fwd(float* y, const float* bias, int N, int C, int HW, int relu, scae_stream_t stream_) {
  SCAE_REQUIRE(y && bias, SCAE_EINVAL, "bias_act fwd: y and bias are required");
  SCAE_REQUIRE(N > 0 && C > 0 && HW > 0 && C <= 65535, SCAE_ELIMIT, "bias_act: bad shape N=%d C=%d HW=%d", N, C, HW);
  const int slabs = bias_slabs(N, C, HW), per = (N + slabs - 1) / slabs;
  bias_act_fwd_kernel<<<dim3(C, slabs), 256, 0, static_cast<cudaStream_t>(stream_)>>>(y, bias, N, C, HW, relu, per);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT size_t scae_bias_act_bwd_workspace_bytes(int N, int C, int HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  return (size_t)bias_slabs(N, C, HW) * C * sizeof(float);
}

SCAE_EXPORT int scae_bias_act_bwd(const float* g, const float* y, float* gx, float* g_bias, int N, int C, int HW,
                                  int relu, void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  SCAE_REQUIRE(g && g_bias && (!relu || (y && gx)), SCAE_EINVAL, "bias_act bwd: a required pointer is NULL");
  SCAE_REQUIRE(N > 0 && C > 0 && HW > 0 && C <= 65535, SCAE_ELIMIT, "bias_act: bad shape N=%d C=%d HW=%d", N, C, HW);
  const int slabs = bias_slabs(N, C, HW), per = (N + slabs - 1) / slabs;
  SCAE_REQUIRE(workspace && workspace_bytes >= (size_t)slabs * C * sizeof(float), SCAE_EINVAL,
               "bias_act bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* partial = static_cast<float*>(workspace);
  bias_act_bwd_kernel<<<dim3(C, slabs), 256, 0, stream>>>(g, y, gx, partial, N, C, HW, relu, per);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return launch_reduce_rows(partial, g_bias, slabs, C, stream);
}

SCAE_EXPORT int scae_pose_transform(const float* t, const float* g, float* out, long rows, int similarity,
                                    scae_stream_t stream_) {
  SCAE_REQUIRE(t && out && rows > 0, SCAE_EINVAL, "pose_transform: t, out and rows > 0 are required");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  long grid = (rows + 255) / 256;
  const long cap = 8L * sm_count();
  if (grid > cap) grid = cap;
  if (g == nullptr) {
    if (similarity) pose_transform_fwd_kernel<true><<<(int)grid, 256, 0, stream>>>(t, out, rows);
    else pose_transform_fwd_kernel<false><<<(int)grid, 256, 0, stream>>>(t, out, rows);
  } else {
    if (similarity) pose_transform_bwd_kernel<true><<<(int)grid, 256, 0, stream>>>(t, g, out, rows);
    else pose_transform_bwd_kernel<false><<<(int)grid, 256, 0, stream>>>(t, g, out, rows);
  }
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_rmsprop_step(float* param, const float* grad, float* square_avg, float* momentum_buf, long n,
                                  float lr, float alpha, float eps, float momentum, scae_stream_t stream_) {
  SCAE_REQUIRE(param && grad && square_avg, SCAE_EINVAL, "rmsprop: param, grad and square_avg are required");
  SCAE_REQUIRE(n > 0, SCAE_EINVAL, "rmsprop: n must be positive");
  SCAE_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(square_avg) && (!momentum_buf || aligned16(momentum_buf)),
               SCAE_EINVAL, "rmsprop: pointers must be 16-byte aligned");
  long grid = (n / 4 + 255) / 256;
  const long cap = 8L * sm_count();
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  rmsprop_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(param, grad, square_avg, momentum_buf, n, lr,
                                                                           alpha, eps, momentum);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_attnpool_fwd(const float* h, float* out, long groups, int D, int S, scae_stream_t stream_) {
  SCAE_REQUIRE(h && out, SCAE_EINVAL, "attnpool fwd: h and out are required");
  SCAE_REQUIRE(groups > 0 && D > 0 && S > 0 && S <= 64 && (long)D * S < (1L << 22), SCAE_ELIMIT,
               "attnpool: S=%d positions (max 64), D=%d", S, D);
  const int warps = 8;
  long grid = (groups + warps - 1) / warps;
  const long cap = 8L * sm_count();
  if (grid > cap) grid = cap;
  attnpool_fwd_kernel<<<(int)grid, 32 * warps, warps * S * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(
      h, out, groups, D, S);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_attnpool_bwd(const float* h, const float* g, float* gh, long groups, int D, int S,
                                  scae_stream_t stream_) {
  SCAE_REQUIRE(h && g && gh, SCAE_EINVAL, "attnpool bwd: a required pointer is NULL");
  SCAE_REQUIRE(groups > 0 && D > 0 && S > 0 && S <= 64 && (long)D * S < (1L << 22) && D <= 1024, SCAE_ELIMIT,
               "attnpool: S=%d positions (max 64), D=%d (max 1024)", S, D);
  const int warps = 8;
  long grid = (groups + warps - 1) / warps;
  const long cap = 8L * sm_count();
  if (grid > cap) grid = cap;
  attnpool_bwd_kernel<<<(int)grid, 32 * warps, warps * (S + D) * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(
      h, g, gh, groups, D, S);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

}  // extern "C"
