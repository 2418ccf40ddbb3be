// Shared device/host helpers for libscae_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/scae_b200.h"

namespace scae {

// ---- host side -------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);   // records the message, returns SCAE_ECUDA
int sm_count();                                   // SM count of the current device (cached per device)
int max_smem_optin();                             // max dynamic shared memory per block (opt-in) of the device
void note_launch();                               // counts one kernel launch of this library (scae_launch_count())
void note_fast_path();                            // counts one hot-path-2 call served by a bulk-copy fast path
void note_persistent_path();                      // ... by the persistent kernels of caps_ll3*.cu in particular

#define SCAE_CUDA_TRY(expr)                                        \
  do {                                                             \
    cudaError_t e__ = (expr);                                      \
    if (e__ != cudaSuccess) return ::scae::cuda_fail(e__, #expr);  \
  } while (0)

#define SCAE_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::scae::set_error(__VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Sums `n_parts` rows of length `n` (partials laid out [n_parts][n]) into out[n], scaled; fixed order => deterministic.
int launch_reduce_rows(const float* partials, float* out, int n_parts, int n, cudaStream_t stream);

// ---- device side -----------------------------------------------------------------------------------------------
// dynamic shared memory of a kernel as a float array (a macro so that tests/emu/simt.h can substitute a host buffer)
#ifndef SCAE_DYNAMIC_SMEM
#define SCAE_DYNAMIC_SMEM(name) extern __shared__ __align__(16) float name[]
#endif

constexpr float kHalfLog2Pi = 0.91893853320467274178f;  // log(sqrt(2*pi))  (torch/distributions/normal.py log_prob)
constexpr float kTwoPi = 6.283185307179586f;             // 2. * math.pi rounded to fp32 (cv_ops.py:45)
constexpr float kLogSafeEps = 1e-16f;                    // math_ops.py:18
constexpr float kLogSafeFloor = -1e8f;                   // math_ops.py:21
constexpr float kDummyLog = -4.605170185988091f;         // fp32(np.log(0.01)) (object_decoder.py:273-274,:281-282)

__device__ __forceinline__ float sigmoid_f(float x) { return __frcp_rn(1.0f + expf(-x)); }

// torch softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float log_safe_f(float p) { return p < kLogSafeEps ? kLogSafeFloor : logf(p); }

// 2x3 affine [a00 a01 a02; a10 a11 a12] from the raw 6-vector (sx, sy, theta, shear, tx, ty); cv_ops.py:36-63.
struct PoseAffine {
  float a[6];
  // intermediates kept for the backward pass
  float sx, sy, sh, tx, ty, c, s;
};

template <bool kSimilarity>
__device__ __forceinline__ void pose_affine_fwd(const float t[6], PoseAffine& o) {
  o.sx = sigmoid_f(t[0]) + 1e-2f;
  o.sy = sigmoid_f(t[1]) + 1e-2f;
  const float theta = t[2] * kTwoPi;
  o.sh = tanhf(t[3] * 5.0f);
  o.tx = tanhf(t[4] * 5.0f);
  o.ty = tanhf(t[5] * 5.0f);
  sincosf(theta, &o.s, &o.c);
  if (kSimilarity) {
    o.a[0] = o.sx * o.c;
    o.a[1] = -o.sx * o.s;
    o.a[3] = o.sx * o.s;
    o.a[4] = o.sx * o.c;
  } else {
    o.a[0] = o.sx * o.c + o.sh * o.sy * o.s;
    o.a[1] = -o.sx * o.s + o.sh * o.sy * o.c;
    o.a[3] = o.sy * o.s;
    o.a[4] = o.sy * o.c;
  }
  o.a[2] = o.tx;
  o.a[5] = o.ty;
}

// gradient w.r.t. the raw 6-vector given the gradient w.r.t. a[6]; oracle/manual_backward.py:_transform_bwd
template <bool kSimilarity>
__device__ __forceinline__ void pose_affine_bwd(const float ga[6], const PoseAffine& o, float gt[6]) {
  float g_sx, g_sy, g_sh, g_c, g_s;
  if (kSimilarity) {
    g_sx = ga[0] * o.c - ga[1] * o.s + ga[3] * o.s + ga[4] * o.c;
    g_sy = 0.0f;
    g_sh = 0.0f;
    g_c = (ga[0] + ga[4]) * o.sx;
    g_s = (ga[3] - ga[1]) * o.sx;
  } else {
    const float mix = ga[0] * o.s + ga[1] * o.c;
    g_sx = ga[0] * o.c - ga[1] * o.s;
    g_sy = mix * o.sh + ga[3] * o.s + ga[4] * o.c;
    g_sh = mix * o.sy;
    g_c = ga[0] * o.sx + ga[1] * o.sh * o.sy + ga[4] * o.sy;
    g_s = ga[0] * o.sh * o.sy - ga[1] * o.sx + ga[3] * o.sy;
  }
  const float g_th = g_s * o.c - g_c * o.s;
  const float e0 = o.sx - 1e-2f, e1 = o.sy - 1e-2f;
  gt[0] = g_sx * e0 * (1.0f - e0);
  gt[1] = g_sy * e1 * (1.0f - e1);
  gt[2] = g_th * kTwoPi;
  gt[3] = g_sh * 5.0f * (1.0f - o.sh * o.sh);
  gt[4] = ga[2] * 5.0f * (1.0f - o.tx * o.tx);
  gt[5] = ga[5] * 5.0f * (1.0f - o.ty * o.ty);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Streaming-logsumexp state: value = m + log(s).  One exp per update (one of the two exponents is always 0).
struct Lse {
  float m, s;
  __device__ __forceinline__ void init(float v) {
    m = v;
    s = 1.0f;
  }
  // returns the weight e^{v - m_new} given to the new element and rescales through `rescale` (= e^{m_old - m_new})
  __device__ __forceinline__ float push(float v, float& rescale) {
    const float e = __expf(-fabsf(v - m));
    const bool up = v > m;
    rescale = up ? e : 1.0f;
    const float wnew = up ? 1.0f : e;
    s = fmaf(s, rescale, wnew);
    m = up ? v : m;
    return wnew;
  }
  __device__ __forceinline__ void push(float v) {
    float r;
    push(v, r);
  }
  __device__ __forceinline__ float value() const { return m + logf(s); }
};

}  // namespace scae
