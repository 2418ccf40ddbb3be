// Hot path 2 (sm_100a): pieces shared by the two generations of capsule-likelihood kernels.
//
//   caps_ll.cu   thread-per-part kernels: any shape, every upstream gradient (the general path)
//   caps_ll2.cu  pair-parallel kernels staged by TMA bulk copies: the fast path for shapes whose per-image working set
//                fits in shared memory and for the upstream-gradient set a training step produces
#pragma once

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace scae {

// everything the forward computes for one (b, o, v) pair
struct CapsPair {
  PoseAffine pa;   // object-part transform (cpr) with its intermediates
  float dyn[6];    // cpr_dynamic as used (zeros when deformations are disabled)
  float vt[6];     // vote
  float lv, pv, vp, u, sc;
};

// ---- MUFU-based elementary functions ----------------------------------------------------------------------------
// The accurate libdevice versions (expf, logf, tanhf, log1pf) cost 10-25 instructions each; a pair needs ~15 of them.
// The forms below are 2-4 instructions.  Error budget: every one is within ~3e-7 absolute of the exact function on
// the ranges that occur here, against the 1e-5 / 1e-4 parity tolerances (tests/test_gpu_capsule.py).
constexpr float kLog2eF = 1.4426950408889634f;
constexpr float kLn2F = 0.6931471805599453f;

// 1 / (1 + e^-x); saturates cleanly: e^-x = inf -> 0, e^-x = 0 -> 1
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-x * kLog2eF)); }

// tanh(5 t) = 1 - 2 / (1 + e^{10 t})
__device__ __forceinline__ float tanh5_fast(float t) {
  return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(t * (10.0f * kLog2eF))), 1.0f);
}

// log(1 + z) for z in [0, 1]: series below 1/16 (where 1 + z would lose z's low bits), lg2 above
__device__ __forceinline__ float log1p_unit(float z) {
  const float series = z * fmaf(z, fmaf(z, fmaf(z, -0.25f, 0.33333334f), -0.5f), 1.0f);
  return z < 0.0625f ? series : lg2_approx(1.0f + z) * kLn2F;
}

// torch softplus: log(1 + e^x) = max(x, 0) + log1p(e^-|x|)   (the x > 20 shortcut of torch is within 2e-9 of this)
__device__ __forceinline__ float softplus_fast(float x) {
  return fmaxf(x, 0.0f) + log1p_unit(ex2_approx(-fabsf(x) * kLog2eF));
}

__device__ __forceinline__ float log_safe_fast(float p) { return p < kLogSafeEps ? kLogSafeFloor : __logf(p); }

// pose_affine_fwd with the fast sigmoid / tanh.  The rotation angle is theta = 2 pi t (cv_ops.py:45), so
// sin / cos (theta) = sinpi / cospi (2 t): sincospif reduces its argument exactly (2 t is exact in fp32), has no slow
// path and no stack frame, and is closer to the fp64 value than sincosf(fl(t * fl(2 pi))) -- the difference to the
// latter is the reference's own rounding of theta, about |theta| * 9e-8.
template <bool kSimilarity>
__device__ __forceinline__ void pose_affine_fast(const float t[6], PoseAffine& o) {
  o.sx = sigmoid_fast(t[0]) + 1e-2f;
  o.sy = sigmoid_fast(t[1]) + 1e-2f;
  o.sh = tanh5_fast(t[3]);
  o.tx = tanh5_fast(t[4]);
  o.ty = tanh5_fast(t[5]);
  sincospif(2.0f * t[2], &o.s, &o.c);
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

// The same with MUFU sine / cosine: theta = 2 pi t, reduced exactly first (t - rint(t) is exact in fp32, |.| <= 1/2), so
// the hardware approximation (absolute error about 5e-7 on [-pi, pi]) never sees a large argument.  7 instructions
// against sincospif's ~25; used by the persistent capsule kernels (caps_ll3.cu).
template <bool kSimilarity>
__device__ __forceinline__ void pose_affine_mufu(const float t[6], PoseAffine& o) {
  o.sx = sigmoid_fast(t[0]) + 1e-2f;
  o.sy = sigmoid_fast(t[1]) + 1e-2f;
  o.sh = tanh5_fast(t[3]);
  o.tx = tanh5_fast(t[4]);
  o.ty = tanh5_fast(t[5]);
  const float th = (t[2] - rintf(t[2])) * kTwoPi;
  o.s = sin_approx(th);
  o.c = cos_approx(th);
  if (kSimilarity) {
    o.a[0] = o.sx * o.c;
    o.a[1] = -o.sx * o.s;
    o.a[3] = o.sx * o.s;
    o.a[4] = o.sx * o.c;
  } else {
    const float shsy = o.sh * o.sy;
    o.a[0] = fmaf(shsy, o.s, o.sx * o.c);
    o.a[1] = fmaf(shsy, o.c, -o.sx * o.s);
    o.a[3] = o.sy * o.s;
    o.a[4] = o.sy * o.c;
  }
  o.a[2] = o.tx;
  o.a[5] = o.ty;
}

// vote = [r00 r01 r02; r10 r11 r12; 0 0 1] . [a00 a01 a02; a10 a11 a12; 0 0 1], rows 0-1 (object_decoder.py:185-191,:413)
__device__ __forceinline__ void compose_vote(const float* r, const float* A_, float* vt) {
  vt[0] = r[0] * A_[0] + r[1] * A_[3];
  vt[1] = r[0] * A_[1] + r[1] * A_[4];
  vt[2] = r[0] * A_[2] + r[1] * A_[5] + r[2];
  vt[3] = r[3] * A_[0] + r[4] * A_[3];
  vt[4] = r[3] * A_[1] + r[4] * A_[4];
  vt[5] = r[3] * A_[2] + r[4] * A_[5] + r[5];
}

// ---- 1-D bulk (TMA) copy runs ---------------------------------------------------------------------------------
// A run of n floats at a 4-byte-aligned global address, staged so that float i lands at base[off + i] with
// off = (address / 4) mod 4: shared and global addresses are then congruent modulo 16 bytes, the 16-byte-aligned
// interior [head, head + body) moves as one bulk copy and at most 3 + 3 edge floats move through registers.
// (all_param rows are 8V+7 floats, so an image's block is 16-byte aligned only for some (b, O).)  Nothing outside
// [0, n) is ever read.  `base` must be 16-byte aligned with room for n + 4 floats.
struct BulkRun {
  int off, head, body, tail;
};
__device__ __forceinline__ BulkRun bulk_run(const void* gptr, int n) {
  BulkRun r;
  r.off = (int)((reinterpret_cast<uintptr_t>(gptr) >> 2) & 3u);
  r.head = min(n, (4 - r.off) & 3);
  r.body = (n - r.head) & ~3;
  r.tail = n - r.head - r.body;
  return r;
}
// edge floats global -> shared by lanes `lane` = 0..5 of the calling group
__device__ __forceinline__ void bulk_run_edges_in(float* base, const float* g, const BulkRun& r, int lane) {
  if (lane < r.head) base[r.off + lane] = __ldg(g + lane);
  const int t = lane - 3;
  if (t >= 0 && t < r.tail) base[r.off + r.head + r.body + t] = __ldg(g + r.head + r.body + t);
}
// edge floats shared -> global
__device__ __forceinline__ void bulk_run_edges_out(float* g, const float* base, const BulkRun& r, int lane) {
  if (lane < r.head) g[lane] = base[r.off + lane];
  const int t = lane - 3;
  if (t >= 0 && t < r.tail) g[r.head + r.body + t] = base[r.off + r.head + r.body + t];
}

// p / V for 0 <= p < 2^22 with inv = 1 / (float)V
__device__ __forceinline__ int fast_div(int p, float inv) { return (int)(((float)p + 0.5f) * inv); }

bool caps_force_v1();
bool caps_force_v2();   // development switch: skip caps_ll3.cu (A/B timing)

// Persistent, warp-specialised fast path (caps_ll3.cu); same contract as caps2_*.
int caps3_fwd(const scae_caps_args* a, const scae_caps_outputs* out, cudaStream_t stream, bool* handled);
int caps3_bwd(const scae_caps_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up, float* g_all_param,
              float* g_shared, float* g_dummy_vote, float* g_x, float* g_presence, void* workspace,
              size_t workspace_bytes, cudaStream_t stream, bool* handled);
size_t caps3_bwd_workspace_bytes(const scae_caps_args* a);

// Fast-path entry points (caps_ll2.cu).  *handled = false means "shape or request not covered, use the general path".
int caps2_fwd(const scae_caps_args* a, const scae_caps_outputs* out, cudaStream_t stream, bool* handled);
int caps2_bwd(const scae_caps_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up, float* g_all_param,
              float* g_shared, float* g_dummy_vote, float* g_x, float* g_presence, void* workspace,
              size_t workspace_bytes, cudaStream_t stream, bool* handled);
size_t caps2_bwd_workspace_bytes(const scae_caps_args* a);

}  // namespace scae
