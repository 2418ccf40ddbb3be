// Hot path 1 (sm_100a): shared pieces of the template-mixture kernels -- launch geometry, the shared-memory template
// atlas, and the bilinear sampler.
//
// Data layout in shared memory ("atlas"): for every template of the current chunk a zero-padded (h+4) x (w+4) grid of
// texels, each texel holding the C colour channels and (alpha mode) the alpha logit interleaved and padded to 1, 2 or 4
// floats, so that one vector LDS fetches everything one bilinear tap needs.  The 2-texel zero border replaces the
// reference sampler's per-tap bounds checks (ATen grid_sampler_2d, padding_mode='zeros'): coordinates are clamped to
// [-1.5, w+0.5] so every tap lands inside the border, and out-of-range taps read zeros exactly like the reference.
#pragma once

#include "common.cuh"

namespace scae {

template <int C, bool kAlpha>
struct TexTraits {
  static constexpr int kCh = C + (kAlpha ? 1 : 0);
  static constexpr int kPad = kCh <= 1 ? 1 : (kCh <= 2 ? 2 : 4);   // floats per texel
  static constexpr int kPixMax = C == 1 ? 8 : 4;                   // pixels per thread held in registers
};

// How one image is mapped on a CTA: threads form `k` row-groups of `tw` columns; a thread owns column j and rows
// r, r+k, r+2k, ... (`ppt` of them) of the current pixel tile.
struct TmplGeom {
  int tw, k, ppt, threads;
  int tiles_x, tiles_y;
  int mc;          // templates per shared-memory chunk
  int pw, ph;      // padded atlas width / height
  int atlas_floats;   // floats of one atlas (mc templates), rounded up to a multiple of 4 to keep 16-byte alignment
  int grid;
  size_t smem_bytes;
};

template <int N>
struct Texel {
  float v[N];
};

template <int N>
__device__ __forceinline__ Texel<N> ld_texel(const float* p);
template <>
__device__ __forceinline__ Texel<1> ld_texel<1>(const float* p) {
  Texel<1> t;
  t.v[0] = *p;
  return t;
}
template <>
__device__ __forceinline__ Texel<2> ld_texel<2>(const float* p) {
  const float2 q = *reinterpret_cast<const float2*>(p);
  Texel<2> t;
  t.v[0] = q.x;
  t.v[1] = q.y;
  return t;
}
template <>
__device__ __forceinline__ Texel<4> ld_texel<4>(const float* p) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  Texel<4> t;
  t.v[0] = q.x;
  t.v[1] = q.y;
  t.v[2] = q.z;
  t.v[3] = q.w;
  return t;
}

// ATen affine_grid base coordinates, align_corners=False: linspace(-1, 1, n) * (n - 1) / n, mirroring ATen's
// symmetric linspace (RangeFactories) and the two-step scaling of AffineGridGenerator.cpp::linspace_from_neg_one.
__device__ __forceinline__ float base_coord(int i, int n) {
  if (n <= 1) return 0.0f;
  const float step = __fdiv_rn(2.0f, (float)(n - 1));
  const float lin = i < n / 2 ? __fadd_rn(-1.0f, __fmul_rn(step, (float)i))
                              : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
  return __fdiv_rn(__fmul_rn(lin, (float)(n - 1)), (float)n);
}

constexpr float kMagic = 8388608.0f;       // 2^23
constexpr int kMagicBits = 0x4B000000;

// Bilinear footprint of one (pixel, template): atlas offset of the north-west tap (in texels) and the four weights.
struct Tap {
  int off;
  float fx, fy;
  float w00, w10, w01, w11;   // nw, ne, sw, se  (ATen grid_sampler_2d naming)
};

// gx, gy: normalised sampling coordinates (affine_grid output).  wf/hf: template size as float.
// Unnormalise as ATen's grid_sampler_unnormalize (align_corners=False): ((g + 1) * size - 1) / 2.
__device__ __forceinline__ void tap_setup(float gx, float gy, float wf, float hf, int pw, Tap& t) {
  float ix = ((gx + 1.0f) * wf - 1.0f) * 0.5f;
  float iy = ((gy + 1.0f) * hf - 1.0f) * 0.5f;
  ix = fminf(fmaxf(ix, -1.5f), wf + 0.5f);
  iy = fminf(fmaxf(iy, -1.5f), hf + 0.5f);
  const float tx = ix + 2.0f, ty = iy + 2.0f;           // atlas coordinates, >= 0.5
  const float ux = __fadd_rd(tx, kMagic), uy = __fadd_rd(ty, kMagic);   // 2^23 + floor(t): exact floor, no F2I
  t.fx = tx - (ux - kMagic);
  t.fy = ty - (uy - kMagic);
  t.off = (__float_as_int(uy) - kMagicBits) * pw + (__float_as_int(ux) - kMagicBits);
  const float gx1 = 1.0f - t.fx, gy1 = 1.0f - t.fy;
  t.w00 = gx1 * gy1;
  t.w10 = t.fx * gy1;
  t.w01 = gx1 * t.fy;
  t.w11 = t.fx * t.fy;
}

// Stages templates [m0, m0+mc) of image b (and the alpha logits) into the atlas interior.  Borders stay zero.
template <int C, bool kAlpha>
__device__ __forceinline__ void stage_atlas(float* atlas, const scae_tmpl_args& a, int b, int m0, int mc, int pw, int ph,
                                            int nthreads) {
  constexpr int kPad = TexTraits<C, kAlpha>::kPad;
  const int hw = a.h * a.w;
  const float inv_w = 1.0f / (float)a.w;
  const float* src = a.templates + ((size_t)b * a.M + m0) * C * hw;
  const int n = mc * C * hw;
  for (int e = threadIdx.x; e < n; e += nthreads) {
    const int plane = e / hw;            // (m, c)
    const int rem = e - plane * hw;
    const int y = (int)(((float)rem + 0.5f) * inv_w);
    const int x = rem - y * a.w;
    const int m = plane / C, c = plane - m * C;
    atlas[(((size_t)m * ph + (y + 2)) * pw + (x + 2)) * kPad + c] = __ldg(src + e);
  }
  if (kAlpha) {
    const float* asrc = a.templates_alpha + (size_t)m0 * hw;
    const int na = mc * hw;
    for (int e = threadIdx.x; e < na; e += nthreads) {
      const int m = e / hw;
      const int rem = e - m * hw;
      const int y = (int)(((float)rem + 0.5f) * inv_w);
      const int x = rem - y * a.w;
      atlas[(((size_t)m * ph + (y + 2)) * pw + (x + 2)) * kPad + C] = __ldg(asrc + e);
    }
  }
}

// Per-image scalars derived from the raw learnt parameters (part_decoder.py:192,:210,:216,:221).
struct TmplScalars {
  float bg_loc;        // sigmoid(bg_value)            (when no bg_image)
  float bg_logit;      // softplus(bg_mixing_logit)    (alpha mode)
  float inv_tau;       // 1 / (softplus(temperature_logit + .5) + 1e-4)   (temperature mode)
  float sigma;         // softplus(scale) + 1e-4 or 1
  float i2s;           // 1 / (2 sigma^2)
  float log_norm;      // log(sigma) + log(sqrt(2 pi))
};

__device__ __forceinline__ TmplScalars tmpl_scalars(const scae_tmpl_args& a) {
  TmplScalars s;
  s.bg_loc = a.bg_value ? sigmoid_f(__ldg(a.bg_value)) : 0.0f;
  s.bg_logit = a.bg_mixing_logit ? softplus_f(__ldg(a.bg_mixing_logit)) : 0.0f;
  s.inv_tau = a.temperature_logit ? __frcp_rn(softplus_f(__ldg(a.temperature_logit) + 0.5f) + 1e-4f) : 1.0f;
  s.sigma = a.scale ? softplus_f(__ldg(a.scale)) + 1e-4f : 1.0f;
  s.i2s = __frcp_rn(2.0f * s.sigma * s.sigma);
  s.log_norm = logf(s.sigma) + kHalfLog2Pi;
  return s;
}

// host side
int tmpl_validate(const scae_tmpl_args* a);
int tmpl_geometry(const scae_tmpl_args* a, int atlas_copies, size_t extra_smem_bytes, int ctas_per_sm_target,
                  TmplGeom* g);

}  // namespace scae
