// Hot path 1 (sm_100a): shared pieces of the template-mixture kernels -- launch geometry, the shared-memory template
// atlas, the per-template coefficient table and the bilinear sampler.
//
// Data layout in shared memory
//   atlas  for every template of the current chunk a zero-padded (h+4) x (w+4) grid of texels, each texel holding the
//          C colour channels and (alpha mode) the alpha logit interleaved and padded to 1, 2 or 4 floats, so that one
//          vector LDS fetches everything one bilinear tap needs.  The 2-texel zero border replaces the reference
//          sampler's per-tap bounds checks (ATen grid_sampler_2d, padding_mode='zeros'): coordinates are clamped so
//          every tap lands inside the border, and out-of-range taps read zeros exactly like the reference.
//   tp     per template 8 floats {Ax, Bx, Cx, Ay, By, Cy, log_safe(presence), presence}: the pose 2x3 matrix folded with
//          ATen's un-normalisation ((g+1)*size-1)/2 and the +2 border offset, so a pixel's atlas coordinate is
//          tx = Ax*X_j + Bx*Y_i + Cx  (one FFMA per axis per pixel once the column term is hoisted).
#pragma once

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace scae {

constexpr int kTmplThreads = 320;   // max threads per CTA of the path-1 kernels (2 CTAs per SM at <= 96 registers)

template <int C, bool kAlpha>
struct TexTraits {
  static constexpr int kCh = C + (kAlpha ? 1 : 0);
  static constexpr int kPad = kCh <= 1 ? 1 : (kCh <= 2 ? 2 : 4);   // floats per texel
  static constexpr int kPixMax = C == 1 ? 5 : 3;                   // pixels per thread held in registers
};

// How one image is mapped on a CTA: threads form `k` row-groups of `tw` columns; a thread owns column j and rows
// r, r+k, r+2k, ... (`ppt` of them) of the current pixel tile.
struct TmplGeom {
  int tw, k, ppt, threads;   // backward (tmpl_bwd_plan): tw = run length L, k = log2(runs per row), ppt = shift of the record skew
  int tiles_x, tiles_y;      // backward: tiles_x = walks per staged band, tiles_y = walks per image
  int mc;             // forward: templates per shared-memory chunk; backward: templates per warp per work unit
  int pw, ph;         // padded atlas width / height
  int atlas_floats;   // floats of one atlas (mc templates), rounded up to a multiple of 4 to keep 16-byte alignment
  int pix_floats;     // backward: floats of the pixel-record tile of one band
  int groups;         // backward: template groups per image (work unit = image x group of `warps per CTA` templates)
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
template <int N>
__device__ __forceinline__ void st_texel(float* p, const Texel<N>& t);
template <>
__device__ __forceinline__ void st_texel<1>(float* p, const Texel<1>& t) {
  *p = t.v[0];
}
template <>
__device__ __forceinline__ void st_texel<2>(float* p, const Texel<2>& t) {
  *reinterpret_cast<float2*>(p) = make_float2(t.v[0], t.v[1]);
}
template <>
__device__ __forceinline__ void st_texel<4>(float* p, const Texel<4>& t) {
  *reinterpret_cast<float4*>(p) = make_float4(t.v[0], t.v[1], t.v[2], t.v[3]);
}

// texel at a 32-bit shared-memory address
template <int N>
__device__ __forceinline__ Texel<N> lds_texel(unsigned addr);
template <>
__device__ __forceinline__ Texel<1> lds_texel<1>(unsigned addr) {
  Texel<1> t;
  t.v[0] = lds_f32(addr);
  return t;
}
template <>
__device__ __forceinline__ Texel<2> lds_texel<2>(unsigned addr) {
  const float2 q = lds_f32x2(addr);
  Texel<2> t;
  t.v[0] = q.x;
  t.v[1] = q.y;
  return t;
}
template <>
__device__ __forceinline__ Texel<4> lds_texel<4>(unsigned addr) {
  const float4 q = lds_f32x4(addr);
  Texel<4> t;
  t.v[0] = q.x;
  t.v[1] = q.y;
  t.v[2] = q.z;
  t.v[3] = q.w;
  return t;
}

// ... reloaded only where `pred` holds
template <int N>
__device__ __forceinline__ void lds_texel_pred(unsigned addr, Texel<N>& t, bool pred) {
  if (N == 1) lds_pred_f32(addr, t.v[0], pred);
  else if (N == 2) lds_pred_f32x2(addr, t.v[0], t.v[N > 1 ? 1 : 0], pred);
  else lds_pred_f32x4(addr, t.v[0], t.v[N > 1 ? 1 : 0], t.v[N > 2 ? 2 : 0], t.v[N > 3 ? 3 : 0], pred);
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
constexpr unsigned kMagicBits = 0x4B000000u;

// exp2 on the MUFU pipe without the denormal-range fix-up __expf/exp2f emit (results below 2^-126 flush to zero,
// which is what a mixture weight that small should do anyway)
__device__ __forceinline__ float ex2_ftz(float x) { return ex2_approx(x); }
constexpr float kLog2e = 1.4426950408889634f;

// Streaming-logsumexp state for the per-pixel mixtures: value = m + log(s), one MUFU per update.
struct PixLse {
  float m, s;
  __device__ __forceinline__ void init(float v) {
    m = v;
    s = 1.0f;
  }
  __device__ __forceinline__ void push(float v) {
    const float e = ex2_ftz(-fabsf(v - m) * kLog2e);
    const bool up = v > m;
    s = fmaf(s, up ? e : 1.0f, up ? 1.0f : e);
    m = up ? v : m;
  }
  __device__ __forceinline__ float value() const { return m + logf(s); }
};

// Bilinear footprint of one (pixel, template): element offset of the north-west tap inside the atlas and the weights.
struct Tap {
  unsigned off;               // in floats, relative to the atlas base (template offset included)
  bool interior;              // the cell touches at least one interior (non-border) texel
  float tx, ty;               // clamped atlas coordinates
  float fx, fy;
  float w00, w10, w01, w11;   // nw, ne, sw, se  (ATen grid_sampler_2d naming)
};

// tx, ty: atlas coordinates (texel coordinate + 2).  lim_x = w + 2.5, lim_y = h + 2.5.  row = pw * kPad,
// base = template offset - kMagicBits * (row + kPad) (unsigned wrap-around arithmetic).
// kUnit = 4: `row` and `base` in bytes, t.off a shared-memory byte address (the backward kernel's explicit LDS / STS).
template <int kPad, int kUnit = 1>
__device__ __forceinline__ void tap_setup(float tx, float ty, float lim_x, float lim_y, unsigned row, unsigned base,
                                          Tap& t) {
  tx = fminf(fmaxf(tx, 0.5f), lim_x);
  ty = fminf(fmaxf(ty, 0.5f), lim_y);
  const float ux = __fadd_rd(tx, kMagic), uy = __fadd_rd(ty, kMagic);   // 2^23 + floor(t): exact floor, no F2I
  // interior texels sit at 2 .. w + 1, a cell (cx, cy) touches texels cx, cx + 1: interior iff 1 <= cx <= w + 1
  t.interior = tx >= 1.0f && tx < lim_x - 0.5f && ty >= 1.0f && ty < lim_y - 0.5f;
  t.tx = tx;
  t.ty = ty;
  t.fx = tx - (ux - kMagic);
  t.fy = ty - (uy - kMagic);
  t.off = __float_as_uint(uy) * row + __float_as_uint(ux) * (unsigned)(kPad * kUnit) + base;
  const float gx1 = 1.0f - t.fx, gy1 = 1.0f - t.fy;
  t.w00 = gx1 * gy1;
  t.w10 = t.fx * gy1;
  t.w01 = gx1 * t.fy;
  t.w11 = t.fx * t.fy;
}

template <int kPad>
__device__ __forceinline__ float bilerp(const Texel<kPad>& t00, const Texel<kPad>& t10, const Texel<kPad>& t01,
                                        const Texel<kPad>& t11, const Tap& t, int c) {
  return fmaf(t11.v[c], t.w11, fmaf(t01.v[c], t.w01, fmaf(t10.v[c], t.w10, t00.v[c] * t.w00)));
}

// Stages templates [m0, m0+mc) of image b (and the alpha logits) into the atlas interior.  Borders stay zero.
template <int C, bool kAlpha>
__device__ __forceinline__ void stage_atlas(float* atlas, const scae_tmpl_args& a, int b, int m0, int mc, int pw, int ph) {
  constexpr int kPad = TexTraits<C, kAlpha>::kPad;
  const int hw = a.h * a.w;
  const float inv_w = 1.0f / (float)a.w, inv_hw = 1.0f / (float)hw;
  const float* col = a.template_color ? a.template_color + ((size_t)b * a.M + m0) * C : nullptr;   // [mc][C]
  const float* src = a.templates + ((col ? (size_t)0 : (size_t)b * a.M) + m0) * C * hw;
  const int n = mc * C * hw;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int plane = (int)(((float)e + 0.5f) * inv_hw);            // (m, c)
    const int rem = e - plane * hw;
    const int y = (int)(((float)rem + 0.5f) * inv_w);
    const int x = rem - y * a.w;
    const int m = plane / C, c = plane - m * C;
    const float v = __ldg(src + e);
    atlas[(((size_t)m * ph + (y + 2)) * pw + (x + 2)) * kPad + c] = col ? v * __ldg(col + plane) : v;
  }
  if (kAlpha) {
    const float* asrc = a.templates_alpha + (size_t)m0 * hw;
    const int na = mc * hw;
    for (int e = threadIdx.x; e < na; e += blockDim.x) {
      const int m = (int)(((float)e + 0.5f) * inv_hw);
      const int rem = e - m * hw;
      const int y = (int)(((float)rem + 0.5f) * inv_w);
      const int x = rem - y * a.w;
      atlas[(((size_t)m * ph + (y + 2)) * pw + (x + 2)) * kPad + C] = __ldg(asrc + e);
    }
  }
}

// Table-driven staging (forward / render kernels).  tab[(c, y, x)] = offset of texel (y, x), channel c, inside ONE
// template's padded block; built once per CTA, it replaces the per-element index decomposition of stage_atlas (65
// instructions per element, 13 % of the forward kernel in profiles/r01f) by one table lookup.
__device__ __forceinline__ void build_stage_table(int* tab, int C, int h, int w, int pw, int kPad) {
  const int hw = h * w;
  for (int e = threadIdx.x; e < C * hw; e += blockDim.x) {
    const int c = e / hw, r2 = e - c * hw, y = r2 / w, x = r2 - y * w;
    tab[e] = ((y + 2) * pw + (x + 2)) * kPad + c;
  }
}
// colour planes of templates [m0, m0+mc) of image b
template <int C>
__device__ __forceinline__ void stage_templates_tab(float* atlas, const int* tab, const scae_tmpl_args& a, int b, int m0,
                                                    int mc, int tex_stride) {
  const int hw = a.h * a.w, chw = C * hw;
  const float inv_chw = 1.0f / (float)chw, inv_hw = 1.0f / (float)hw;
  const float* col = a.template_color ? a.template_color + ((size_t)b * a.M + m0) * C : nullptr;   // [mc][C]
  const float* src = a.templates + ((col ? (size_t)0 : (size_t)b * a.M) + m0) * chw;
  const int n = mc * chw;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int m = (int)(((float)e + 0.5f) * inv_chw);
    float v = __ldg(src + e);
    if (col) v *= __ldg(col + (C == 1 ? m : (int)(((float)e + 0.5f) * inv_hw)));   // (m, c) plane index = e / hw
    atlas[m * tex_stride + tab[e - m * chw]] = v;
  }
}
// the batch-shared alpha logits of templates [m0, m0+mc): channel slot C of every texel
template <int C>
__device__ __forceinline__ void stage_alpha_tab(float* atlas, const int* tab, const scae_tmpl_args& a, int m0, int mc,
                                                int tex_stride) {
  const int hw = a.h * a.w;
  const float inv_hw = 1.0f / (float)hw;
  const float* src = a.templates_alpha + (size_t)m0 * hw;
  const int n = mc * hw;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int m = (int)(((float)e + 0.5f) * inv_hw);
    atlas[m * tex_stride + tab[e - m * hw] + C] = __ldg(src + e);   // tab[0 .. hw) is channel 0
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

// ---- shared-memory carve-up common to all path-1 kernels ----------------------------------------------------------
struct TmplSmem {
  float* atlas;
  float* extra;    // backward: gradient atlas (atomic variant) or per-pixel gradient buffer (gather variant)
  float* tp;       // [M][8]
  float* xs;       // [W]
  float* ys;       // [H]
  float* red;      // [64] block-reduction scratch, followed by kernel-specific scratch
};

__device__ __forceinline__ TmplSmem tmpl_carve(float* smem, const scae_tmpl_args& a, const TmplGeom& g, int extra_floats) {
  TmplSmem s;
  s.atlas = smem;
  s.extra = smem + g.atlas_floats;
  s.tp = s.extra + extra_floats;
  s.xs = s.tp + (size_t)a.M * 8;
  s.ys = s.xs + a.W;
  s.red = s.ys + a.H;
  return s;
}

__device__ __forceinline__ void tmpl_prologue(const TmplSmem& s, const scae_tmpl_args& a, const TmplGeom& g, int extra_floats) {
  const int n = g.atlas_floats + extra_floats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s.atlas[i] = 0.0f;
  for (int i = threadIdx.x; i < a.W; i += blockDim.x) s.xs[i] = base_coord(i, a.W);
  for (int i = threadIdx.x; i < a.H; i += blockDim.x) s.ys[i] = base_coord(i, a.H);
}

// pose (used directly as affine_grid theta, part_decoder.py:176) folded with grid_sampler_unnormalize and the border
__device__ __forceinline__ void load_pose_table(const TmplSmem& s, const scae_tmpl_args& a, int b) {
  const float hw_x = 0.5f * (float)a.w, hw_y = 0.5f * (float)a.h;
  for (int i = threadIdx.x; i < a.M; i += blockDim.x) {
    const float* p = a.pose + ((size_t)b * a.M + i) * 6;
    float* t = s.tp + i * 8;
    t[0] = __ldg(p + 0) * hw_x;
    t[1] = __ldg(p + 1) * hw_x;
    t[2] = (__ldg(p + 2) + 1.0f) * hw_x + 1.5f;      // ((g+1)*w - 1)/2 + 2
    t[3] = __ldg(p + 3) * hw_y;
    t[4] = __ldg(p + 4) * hw_y;
    t[5] = (__ldg(p + 5) + 1.0f) * hw_y + 1.5f;
    const float pr = a.presence ? __ldg(a.presence + (size_t)b * a.M + i) : 1.0f;
    t[6] = a.presence ? log_safe_f(pr) : 0.0f;
    t[7] = pr;
  }
}

// deterministic block sum; result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];
  }
  return t;
}

// ---- host side ------------------------------------------------------------------------------------------------------
int tmpl_validate(const scae_tmpl_args* a);
int tmpl_texel_floats(const scae_tmpl_args* a);
// Picks the thread/pixel mapping and the template chunk.  per_tmpl_extra_bytes: extra shared memory per chunk template
// (gradient atlas / gradient buffer); fixed_extra_bytes: kernel scratch independent of the chunk size.
int tmpl_geometry(const scae_tmpl_args* a, size_t per_tmpl_extra_bytes, size_t fixed_extra_bytes, size_t smem_budget,
                  TmplGeom* g);

template <typename F>
static int tmpl_prepare_kernel(F kern, size_t smem_bytes) {
  if (smem_bytes > 48 * 1024)
    SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return SCAE_OK;
}

#define SCAE_TMPL_DISPATCH(C_, ALPHA_, ...)                                                       \
  do {                                                                                            \
    if ((C_) == 1 && (ALPHA_)) { constexpr int kC = 1; constexpr bool kA = true; __VA_ARGS__; }   \
    else if ((C_) == 1) { constexpr int kC = 1; constexpr bool kA = false; __VA_ARGS__; }         \
    else if ((C_) == 2 && (ALPHA_)) { constexpr int kC = 2; constexpr bool kA = true; __VA_ARGS__; } \
    else if ((C_) == 2) { constexpr int kC = 2; constexpr bool kA = false; __VA_ARGS__; }         \
    else if ((ALPHA_)) { constexpr int kC = 3; constexpr bool kA = true; __VA_ARGS__; }           \
    else { constexpr int kC = 3; constexpr bool kA = false; __VA_ARGS__; }                        \
  } while (0)

}  // namespace scae
