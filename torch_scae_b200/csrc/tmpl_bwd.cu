// Hot path 1 (sm_100a): backward of the template-mixture log-likelihood.
//
// Math: oracle/manual_backward.py::template_forward_backward.  Per (pixel, template) the kernel recomputes the bilinear
// sample, forms the mixture responsibilities from the two logsumexp terms cached by the forward kernel, and produces
//   g_loc   = G * pN * (x - loc) / sigma^2  (+ logit path in temperature mode)      -> template gradient
//   g_logit = G * (pN - pD)                                                         -> alpha / presence gradient
//   g_tx, g_ty (gradient w.r.t. the sampling coordinates)                           -> pose gradient
//
// The template / alpha gradient is a *transposed* bilinear interpolation.  Written as a scatter it needs 8 float
// atomics on shared memory per (pixel, template); on sm_100a those are ATOMS.CAST spin loops and neighbouring pixels hit
// the same texel (templates are magnified), which made the first version of this kernel 18x slower than the forward
// pass (profiles/r01a).  The shipped formulation is atomics-free and bit-reproducible:
//
//   one warp owns one (image, template) pair and walks the image row-major, 32 pixels per pass.  Along a row the
//   sampling coordinates move on a straight line, so the bilinear cell index is monotone: pixels that fall into the same
//   cell are CONTIGUOUS lanes.  A segmented warp scan (shuffles) pre-reduces the four corner contributions per cell, and
//   only the last lane of each segment does a plain read-modify-write on the warp's private gradient atlas in shared
//   memory -- corner by corner and row by row, so no two lanes ever touch the same address in the same step and the
//   summation order is fixed.  No CTA barriers inside a template, no gradient buffer, any image size.
//
// (A texel-parallel gather formulation -- every texel inverts the affine map and walks its own footprint -- was
// measured at 1.5 ms against this kernel's 1.0 ms at the MNIST config and removed; profiles/r01b.)
#include <stdlib.h>
#include <string.h>

#include "tmpl_common.cuh"

namespace scae {

struct TmplBwdOut {
  float* g_templates;       // [B,M,C,h,w]; unused with template_color (see raw_partials)
  float* g_color;           // [B,M,C]           (template_color only)
  float* raw_partials;      // [grid][M*C*h*w]   (template_color only): per-CTA sums of colour * texel gradient
  float* g_pose;
  float* g_presence;
  float* g_bg_image;
  float* alpha_partials;    // [grid][M*h*w]   (alpha mode)
  float* scalar_partials;   // [grid][4]
};

// whole-kernel accumulators of the batch-reduced scalar gradients
struct ScalarAcc {
  float bgval = 0.f, bglogit = 0.f, tau = 0.f, sig = 0.f, g = 0.f;
};

// background component of one pixel (evaluated once per pixel, with the first template group)
template <int C, bool kAlpha, bool kMode>
__device__ __forceinline__ void bwd_background(const scae_tmpl_args& a, const TmplScalars& sc, const float* xv,
                                               const float* G, const float* Nc, const float* Dc, size_t px0, int HW,
                                               float* g_bg_image, ScalarAcc& acc) {
  const float two_i2s = 2.0f * sc.i2s;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const size_t px = px0 + (size_t)c * HW;
    if (kMode) {   // backward of pdf.mode(): the pixel's gradient goes to the component it was taken from (Nc = its index)
      const float gl = Nc[c] == (float)a.M ? G[c] : 0.0f;
      if (a.bg_image) {
        if (g_bg_image) g_bg_image[px] = gl;
      } else {
        acc.bgval += gl;
      }
      continue;
    }
    const float bg = a.bg_image ? __ldg(a.bg_image + px) : sc.bg_loc;
    const float bl = kAlpha ? sc.bg_logit : bg * sc.inv_tau;
    const float d = xv[c] - bg;
    const float pN = expf(fmaf(d * d, -sc.i2s, bl) - Nc[c]);
    const float pD = expf(bl - Dc[c]);
    const float glog = G[c] * (pN - pD);
    float gl = G[c] * pN * d * two_i2s;
    if (!kAlpha) {
      gl = fmaf(glog, sc.inv_tau, gl);
      acc.tau = fmaf(glog, bg, acc.tau);
    } else {
      acc.bglogit += glog;
    }
    acc.sig = fmaf(G[c] * pN, d * d, acc.sig);
    acc.g += G[c];
    if (a.bg_image) {
      if (g_bg_image) g_bg_image[px] = gl;
    } else {
      acc.bgval += gl;
    }
  }
}

// per-(pixel, template) gradients.  Returns g_loc[c] (c < C) and, in v[C], the summed logit gradient (alpha mode).
template <int C, bool kAlpha, int kPad, bool kMode>
__device__ __forceinline__ Texel<kPad> bwd_pixel(const TmplScalars& sc, const Texel<kPad>& t00, const Texel<kPad>& t10,
                                                 const Texel<kPad>& t01, const Texel<kPad>& t11, const Tap& t,
                                                 float lpres, const float* xv, const float* G, const float* Nc,
                                                 const float* Dc, ScalarAcc& acc, float& glp, float& gtx, float& gty,
                                                 float m_index) {
  const float two_i2s = 2.0f * sc.i2s;
  const float gx1 = 1.0f - t.fx, gy1 = 1.0f - t.fy;
  Texel<kPad> out;
#pragma unroll
  for (int c = 0; c < kPad; ++c) out.v[c] = 0.0f;
  if (kMode) {
    // backward of pdf.mode() (distributions.py:50-77 without the straight-through estimator): d mode / d loc = 1 for the
    // component the arg-max picked (its index sits in Nc), nothing flows to the mixing logits
    glp = gtx = gty = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float gl = Nc[c] == m_index ? G[c] : 0.0f;
      out.v[c] = gl;
      gtx = fmaf(gl, fmaf(t11.v[c] - t01.v[c], t.fy, (t10.v[c] - t00.v[c]) * gy1), gtx);
      gty = fmaf(gl, fmaf(t11.v[c] - t10.v[c], t.fx, (t01.v[c] - t00.v[c]) * gx1), gty);
    }
    return out;
  }
  float al = 0.0f, pD_shared = 0.0f;
  if (kAlpha) {
    al = bilerp<kPad>(t00, t10, t01, t11, t, C) + lpres;
    pD_shared = ex2_ftz((al - Dc[0]) * kLog2e);
  }
  glp = 0.0f;
  gtx = 0.0f;
  gty = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float loc = bilerp<kPad>(t00, t10, t01, t11, t, c);
    const float d = xv[c] - loc;
    const float logit = kAlpha ? al : fmaf(loc, sc.inv_tau, lpres);
    const float pN = ex2_ftz((fmaf(d * d, -sc.i2s, logit) - Nc[c]) * kLog2e);
    const float pD = kAlpha ? pD_shared : ex2_ftz((logit - Dc[c]) * kLog2e);
    const float glog = G[c] * (pN - pD);
    float gl = G[c] * pN * d * two_i2s;
    if (!kAlpha) {
      gl = fmaf(glog, sc.inv_tau, gl);
      acc.tau = fmaf(glog, loc, acc.tau);
    }
    acc.sig = fmaf(G[c] * pN, d * d, acc.sig);
    glp += glog;
    out.v[c] = gl;
    // d loc / d tx = (ne - nw)(1 - fy) + (se - sw) fy ;  d loc / d ty = (sw - nw)(1 - fx) + (se - ne) fx
    gtx = fmaf(gl, fmaf(t11.v[c] - t01.v[c], t.fy, (t10.v[c] - t00.v[c]) * gy1), gtx);
    gty = fmaf(gl, fmaf(t11.v[c] - t10.v[c], t.fx, (t01.v[c] - t00.v[c]) * gx1), gty);
  }
  if (kAlpha) {
    out.v[C] = glp;
    gtx = fmaf(glp, fmaf(t11.v[C] - t01.v[C], t.fy, (t10.v[C] - t00.v[C]) * gy1), gtx);
    gty = fmaf(glp, fmaf(t11.v[C] - t10.v[C], t.fx, (t01.v[C] - t00.v[C]) * gx1), gty);
  }
  return out;
}

__device__ __forceinline__ void write_scalar_partials(const scae_tmpl_args& a, const TmplScalars& sc, bool alpha,
                                                      const ScalarAcc& acc, float* red, float* sp) {
  const float t_bgval = block_sum(acc.bgval, red);
  const float t_bglogit = block_sum(acc.bglogit, red);
  const float t_tau = block_sum(acc.tau, red);
  const float t_sig = block_sum(acc.sig, red);
  const float t_g = block_sum(acc.g, red);
  if (threadIdx.x == 0) {
    // already mapped to the raw parameters (sigmoid / softplus chain rule); the per-CTA rows are summed afterwards
    sp[0] = a.bg_value && !a.bg_image ? t_bgval * sc.bg_loc * (1.0f - sc.bg_loc) : 0.0f;
    sp[1] = alpha ? t_bglogit * sigmoid_f(__ldg(a.bg_mixing_logit)) : 0.0f;
    sp[2] = alpha ? 0.0f : -t_tau * sc.inv_tau * sc.inv_tau * sigmoid_f(__ldg(a.temperature_logit) + 0.5f);
    const float inv_sigma = __frcp_rn(sc.sigma);
    sp[3] = a.scale ? (t_sig * inv_sigma * inv_sigma * inv_sigma - t_g * inv_sigma) * sigmoid_f(__ldg(a.scale)) : 0.0f;
  }
}

// ================================================================================================================
// scan kernel: warp per (image, template), segmented-scan scatter
// ================================================================================================================
constexpr int kScanThreads = 256;

// texel += v over the NCH live channels, as one vector read-modify-write of the padded texel
template <int kPad, int NCH>
__device__ __forceinline__ void texel_add(float* dst, const float* v) {
  Texel<kPad> t = ld_texel<kPad>(dst);
#pragma unroll
  for (int c = 0; c < NCH; ++c) t.v[c] += v[c];
  st_texel<kPad>(dst, t);
}

// Work unit = (image b, template group grp): the CTA's warps take the templates grp * nwarps + warp.  Units are dealt
// round-robin to the persistent CTAs, so a batch of 1024 images is 5120 units over 592 CTAs (8.6 each) instead of
// 1.7 whole images each -- the tail of the last wave shrinks from 14 % to 4 % of the kernel.
// (Image too large for the shared-memory pixel records: one unit = one whole image, so that the pixel data the warps
// re-read from global memory stays in the CTA's L1.)
//
// Occupancy beats instruction count here (profiles/r01l): 4 CTAs/SM at 64 registers is 5-7 % faster than 3 CTAs/SM at
// 80 registers, also with the base-grid coordinates in a shared-memory table.
// kMode: the same scatter for the backward of pdf.mode() -- `gout` is the gradient w.r.t. the mode image and `cache`'s
// first C planes hold the index of the component each pixel took its value from (scae_tmpl_mode_bwd)
template <int C, bool kAlpha, bool kMode>
__global__ void __launch_bounds__(kScanThreads, C == 1 ? 4 : 3) tmpl_ll_bwd_scan_kernel(const scae_tmpl_args a,
                                                                                    const float* __restrict__ x,
                                                                                    const float* __restrict__ gout,
                                                                                    const float* __restrict__ cache,
                                                                                    const TmplBwdOut out, const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, NCH = TT::kCh;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int atlas_floats = g.atlas_floats;               // one padded template, multiple of 4 floats
  float* atlas = smem + (size_t)warp * 2 * atlas_floats;  // this warp's value atlas
  float* gatlas = atlas + atlas_floats;                   // ... and gradient atlas
  float* red = smem + (size_t)nwarps * 2 * atlas_floats;  // [64] block-reduction scratch
  float* xs = red + 64;                                   // [W] affine_grid base coordinates
  float* ys = xs + a.W;                                   // [H]
  const int HW = a.H * a.W, hw = a.h * a.w, pw = g.pw, ph = g.ph, W = a.W;
  // [C][H*W] records {x, upstream gradient, cached numerator lse, cached denominator lse} of the current image: every
  // warp of the CTA works on the same image, so the four global loads (and their 64-bit address arithmetic) that each
  // (template, pass) would repeat are paid once per unit
  const bool staged = g.pix_floats > 0;
  float4* PIX = reinterpret_cast<float4*>(smem + (((size_t)nwarps * 2 * atlas_floats + 64 + a.W + a.H + 3) & ~(size_t)3));
  // When H*W is not a multiple of 32 the last pass has dead lanes; every plane is followed by 32 pad records
  // {0, 0, 1e30, 1e30}: upstream gradient 0 and both responsibilities exp(-huge) = 0, so a dead lane computes
  // exact zeros without any select in the hot loop.
  const bool all_valid = (HW & 31) == 0;
  const int HWp = all_valid ? HW : HW + 32;
  if (staged && !all_valid) {
    for (int e = threadIdx.x; e < C * 32; e += blockDim.x)
      PIX[(e >> 5) * HWp + HW + (e & 31)] = make_float4(0.f, 0.f, 1e30f, 1e30f);
  }
  for (int e = lane; e < 2 * atlas_floats; e += 32) atlas[e] = 0.0f;
  for (int e = threadIdx.x; e < a.W; e += blockDim.x) xs[e] = base_coord(e, a.W);
  for (int e = threadIdx.x; e < a.H; e += blockDim.x) ys[e] = base_coord(e, a.H);
  const TmplScalars sc = tmpl_scalars(a);
  const float lim_x = keep((float)a.w + 2.5f), lim_y = keep((float)a.h + 2.5f);
  const unsigned row = keep((unsigned)(pw * kPad));
  // tap offsets are relative to the start of shared memory: the warp's atlas offset is folded into the constant
  const unsigned atlas_off = (unsigned)(warp * 2 * atlas_floats);
  const unsigned base0 = keep(atlas_off - kMagicBits * (row + (unsigned)kPad));
  const unsigned gat = keep((unsigned)atlas_floats);
  const float hw_x = 0.5f * (float)a.w, hw_y = 0.5f * (float)a.h;
  const float inv_pw = 1.0f / (float)pw;
  float* my_alpha_partial = (kAlpha && out.alpha_partials) ? out.alpha_partials + (size_t)blockIdx.x * a.M * hw : nullptr;
  if (my_alpha_partial)
    for (int e = threadIdx.x; e < a.M * hw; e += blockDim.x) my_alpha_partial[e] = 0.0f;
  // fused colourisation: the batch-shared raw templates get a per-CTA partial row like the alpha logits
  const bool colored = a.template_color != nullptr;
  float* my_raw_partial = colored ? out.raw_partials + (size_t)blockIdx.x * a.M * C * hw : nullptr;
  if (my_raw_partial)
    for (int e = threadIdx.x; e < a.M * C * hw; e += blockDim.x) my_raw_partial[e] = 0.0f;
  ScalarAcc acc;
  __syncthreads();   // partial row zeroed before any warp accumulates into it; xs / ys complete
  // row / column of this lane's pixel in the first pass, and its advance per pass
  const int i0 = lane / W, j0 = lane - i0 * W;
  const int di = 32 / W, dj = 32 - di * W;

  const int groups = g.groups, n_units = a.B * groups;
  for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int b = u / groups, grp = u - b * groups;
    const bool first = grp == 0;   // the background component and g_bg_image belong to the image, not to a template
    // ---- pixel records of the image (whole CTA) and, once per image, the background component ----------------------
    if (staged) __syncthreads();                       // the previous unit's records are no longer read
    if (staged || (first && warp == 0)) {
      const int p_first = staged ? (int)threadIdx.x : lane, p_step = staged ? (int)blockDim.x : 32;
      for (int p = p_first; p < HW; p += p_step) {
        float xv[C], G[C], Nc[C], Dc[C];
        const size_t px0 = (size_t)b * C * HW + p;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const size_t cx = (size_t)b * 2 * C * HW + (size_t)c * HW + p;
          xv[c] = __ldg(x + px0 + (size_t)c * HW);
          G[c] = __ldg(gout + px0 + (size_t)c * HW);
          Nc[c] = __ldg(cache + cx);
          Dc[c] = __ldg(cache + cx + (size_t)C * HW);
          if (staged) PIX[c * HWp + p] = make_float4(xv[c], G[c], Nc[c], Dc[c]);
        }
        if (first) bwd_background<C, kAlpha, kMode>(a, sc, xv, G, Nc, Dc, px0, HW, out.g_bg_image, acc);
      }
    }
    if (staged) __syncthreads();
    for (int tt = 0; tt < g.mc; ++tt) {
      const int m = (grp * g.mc + tt) * nwarps + warp;
      if (m >= a.M) break;
      // ---- per-template setup (all lanes compute the same coefficients) -----------------------------------------
      const float* pp = a.pose + ((size_t)b * a.M + m) * 6;
      const float Ax = __ldg(pp + 0) * hw_x, Bx = __ldg(pp + 1) * hw_x, Cx = (__ldg(pp + 2) + 1.0f) * hw_x + 1.5f;
      const float Ay = __ldg(pp + 3) * hw_y, By = __ldg(pp + 4) * hw_y, Cy = (__ldg(pp + 5) + 1.0f) * hw_y + 1.5f;
      const float pres = a.presence ? __ldg(a.presence + (size_t)b * a.M + m) : 1.0f;
      const float lpres = a.presence ? log_safe_f(pres) : 0.0f;
      const float* src = a.templates + ((colored ? (size_t)0 : (size_t)b * a.M) + m) * C * hw;
      float colv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) colv[c] = colored ? __ldg(a.template_color + ((size_t)b * a.M + m) * C + c) : 1.0f;
      {
        const float inv_w = 1.0f / (float)a.w;
        for (int e = lane; e < hw; e += 32) {
          const int y = (int)(((float)e + 0.5f) * inv_w), xx = e - y * a.w;
          float* q = atlas + ((size_t)(y + 2) * pw + (xx + 2)) * kPad;
#pragma unroll
          for (int c = 0; c < C; ++c) q[c] = __ldg(src + (size_t)c * hw + e) * colv[c];
          if (kAlpha) q[C] = __ldg(a.templates_alpha + (size_t)m * hw + e);
        }
      }
      __syncwarp();
      float sgx = 0.f, sgxX = 0.f, sgxY = 0.f, sgy = 0.f, sgyX = 0.f, sgyY = 0.f, spres = 0.f;

      // ---- passes of 32 consecutive pixels (row-major) -----------------------------------------------------------
      int i = i0, j = j0;
      const float4* pix_ptr = PIX + lane;     // loop-carried pointer: one IADD per pass instead of a re-derived address
      for (int p0 = 0; p0 < HW; p0 += 32, pix_ptr += 32) {
        const int p = p0 + lane;
        const bool valid = p < HW;
        float xv[C], G[C], Nc[C], Dc[C];
        if (staged) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 r4 = pix_ptr[c * HWp];     // a dead lane reads a pad record
            xv[c] = r4.x;
            G[c] = r4.y;
            Nc[c] = r4.z;
            Dc[c] = r4.w;
          }
        } else {
          const int pc = valid ? p : 0;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const size_t px = (size_t)b * C * HW + (size_t)c * HW + pc;
            const size_t cx = (size_t)b * 2 * C * HW + (size_t)c * HW + pc;
            xv[c] = __ldg(x + px);
            G[c] = valid ? __ldg(gout + px) : 0.0f;
            Nc[c] = __ldg(cache + cx);
            Dc[c] = __ldg(cache + cx + (size_t)C * HW);
          }
        }
        const float X = xs[valid ? j : 0], Y = ys[valid ? i : 0];
        i += di;                         // advance to the next pass
        j += dj;
        if (j >= W) {
          j -= W;
          ++i;
        }
        Tap t;
        tap_setup<kPad>(fmaf(Y, Bx, fmaf(X, Ax, Cx)), fmaf(Y, By, fmaf(X, Ay, Cy)), lim_x, lim_y, row, base0, t);
        const float* q = smem + t.off;
        const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
        const Texel<kPad> t01 = ld_texel<kPad>(q + row), t11 = ld_texel<kPad>(q + row + kPad);
        float glp, gtx, gty;
        const Texel<kPad> gv =
            bwd_pixel<C, kAlpha, kPad, kMode>(sc, t00, t10, t01, t11, t, lpres, xv, G, Nc, Dc, acc, glp, gtx, gty, (float)m);
        sgx += gtx;
        sgxX = fmaf(gtx, X, sgxX);
        sgxY = fmaf(gtx, Y, sgxY);
        sgy += gty;
        sgyX = fmaf(gty, X, sgyX);
        sgyY = fmaf(gty, Y, sgyY);
        spres += glp;

        // Cells that touch no interior texel land in the zero border, whose gradient is discarded.  Parts cover a
        // fraction of the image, so whole passes miss the template: they keep their logit / presence gradient (above) and
        // skip the scatter.
        if (!__any_sync(0xffffffffu, valid && t.interior)) continue;
        // ---- segmented scan over lanes that share (row, cell) -------------------------------------------------------
        // the base-grid Y is strictly increasing with the row, so "same row" is "same Y"
        float v[4][NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          v[0][c] = gv.v[c] * t.w00;
          v[1][c] = gv.v[c] * t.w10;
          v[2][c] = gv.v[c] * t.w01;
          v[3][c] = gv.v[c] * t.w11;
        }
        const unsigned key = valid ? t.off : 0xFFFFFFFFu;
        const unsigned key_prev = __shfl_up_sync(0xffffffffu, key, 1);
        const float y_prev = __shfl_up_sync(0xffffffffu, Y, 1);
        const bool head = lane == 0 || key != key_prev || Y != y_prev;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int seg_start = 31 - __clz(heads & (0xFFFFFFFFu >> (31 - lane)));
        const int dist = lane - seg_start;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          if (__ballot_sync(0xffffffffu, dist >= d) == 0u) break;
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              const float up = __shfl_up_sync(0xffffffffu, v[k][c], d);
              if (dist >= d) v[k][c] += up;
            }
        }
        const bool tail = valid && (lane == 31 || ((heads >> (lane + 1)) & 1u));
        // ---- segment tails update the warp's gradient atlas: corner by corner, row by row => no address collisions ---
        const unsigned valid_mask = __ballot_sync(0xffffffffu, valid);
        const float y_first = __shfl_sync(0xffffffffu, Y, 0);
        const float y_last = __shfl_sync(0xffffffffu, Y, 31 - __clz(valid_mask));
        float* gq = smem + (t.off + gat);
        // A pass of 32 consecutive pixels can straddle image rows.  Within one row the cells of different segments are
        // distinct; across rows they could coincide (extreme magnification), so check once with MATCH and only then
        // fall back to updating row by row.
        bool by_row = false;
        if (y_first != y_last) {
          const unsigned tails = __ballot_sync(0xffffffffu, tail);
          const unsigned peers = __match_any_sync(0xffffffffu, tail ? key : (0xFFFFFF00u | (unsigned)lane));
          by_row = __any_sync(0xffffffffu, tail && (peers & tails & ~(1u << lane)) != 0u);
        }
        if (!by_row) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (tail) texel_add<kPad, NCH>(gq + (k & 1 ? kPad : 0) + (k & 2 ? row : 0), v[k]);
            __syncwarp();
          }
        } else {
          float y_cur = y_first;
          for (;;) {
            const bool mine = tail && Y == y_cur;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (mine) texel_add<kPad, NCH>(gq + (k & 1 ? kPad : 0) + (k & 2 ? row : 0), v[k]);
              __syncwarp();
            }
            const unsigned later = __ballot_sync(0xffffffffu, valid && Y > y_cur);
            if (later == 0u) break;
            y_cur = __shfl_sync(0xffffffffu, Y, __ffs(later) - 1);
          }
        }
      }

      // ---- pose / presence gradients of (b, m) -------------------------------------------------------------------
      {
        float v7[7] = {sgxX, sgxY, sgx, sgyX, sgyY, sgy, spres};
#pragma unroll
        for (int q7 = 0; q7 < 7; ++q7) v7[q7] = warp_sum(v7[q7]);
        if (lane == 0) {
          float* gp = out.g_pose + ((size_t)b * a.M + m) * 6;
          gp[0] = v7[0] * hw_x;
          gp[1] = v7[1] * hw_x;
          gp[2] = v7[2] * hw_x;
          gp[3] = v7[3] * hw_y;
          gp[4] = v7[4] * hw_y;
          gp[5] = v7[5] * hw_y;
          if (out.g_presence) out.g_presence[(size_t)b * a.M + m] = pres < kLogSafeEps ? 0.0f : v7[6] / pres;
        }
      }
      __syncwarp();
      // ---- flush + clear the gradient atlas ----------------------------------------------------------------------
      {
        float* dst = colored ? my_raw_partial + (size_t)m * C * hw : out.g_templates + ((size_t)b * a.M + m) * C * hw;
        float gcol[C];
#pragma unroll
        for (int c = 0; c < C; ++c) gcol[c] = 0.0f;
        for (int e = lane; e < pw * ph; e += 32) {
          const int yy = (int)(((float)e + 0.5f) * inv_pw), xx = e - yy * pw;
          float* q = gatlas + (size_t)e * kPad;
          if (yy >= 2 && yy < ph - 2 && xx >= 2 && xx < pw - 2) {
            const int te = (yy - 2) * a.w + (xx - 2);
            if (colored) {
              // template = raw * colour: d/d raw = colour * g (summed over the batch), d/d colour = sum_texels raw * g
#pragma unroll
              for (int c = 0; c < C; ++c) {
                dst[(size_t)c * hw + te] += colv[c] * q[c];
                gcol[c] = fmaf(__ldg(src + (size_t)c * hw + te), q[c], gcol[c]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < C; ++c) dst[(size_t)c * hw + te] = q[c];
            }
            if (kAlpha && my_alpha_partial) my_alpha_partial[(size_t)m * hw + te] += q[C];
          }
#pragma unroll
          for (int c = 0; c < kPad; ++c) q[c] = 0.0f;
        }
        if (colored) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float t = warp_sum(gcol[c]);
            if (lane == 0) out.g_color[((size_t)b * a.M + m) * C + c] = t;
          }
        }
      }
      __syncwarp();
    }
  }
  write_scalar_partials(a, sc, kAlpha, acc, red, out.scalar_partials + (size_t)blockIdx.x * 4);
}

// ================================================================================================================
// host
// ================================================================================================================
static int tmpl_bwd_plan(const scae_tmpl_args* a, TmplGeom* gp) {
  const int kpad = tmpl_texel_floats(a);
  const size_t limit = (size_t)max_smem_optin();
  // one padded value atlas + one gradient atlas per warp, nothing that scales with the image
  TmplGeom& g = *gp;
  memset(&g, 0, sizeof(g));
  g.pw = a->w + 4;
  g.ph = a->h + 4;
  g.atlas_floats = (int)(((size_t)g.pw * g.ph * kpad + 3) / 4 * 4);
  int threads = kScanThreads;
  size_t smem;
  for (;;) {
    smem = ((size_t)(threads / 32) * 2 * g.atlas_floats + 64 + a->W + a->H + 8) * sizeof(float);
    if (smem <= limit || threads == 32) break;
    threads /= 2;                                   // very large templates: fewer warps per CTA
  }
  SCAE_REQUIRE(smem <= limit, SCAE_ELIMIT, "tmpl bwd: a %dx%d template does not fit in shared memory", a->h, a->w);
  const int warps = threads / 32;
  // per-image pixel records in shared memory when at least two CTAs per SM still fit (each CTA also pays 1 KB reserved)
  {
    const size_t hw = (size_t)a->H * a->W, hwp = (hw & 31) == 0 ? hw : hw + 32;   // + pad records for dead lanes
    const size_t pix_bytes = 4 * a->C * hwp * sizeof(float) + 16;
    const char* e = getenv("SCAE_TMPL_BWD_STAGE");
    const bool allow = e == nullptr || strcmp(e, "0") != 0;
    if (allow && 2 * (smem + pix_bytes + 1024) <= limit + 1024) {
      g.pix_floats = (int)(4 * a->C * hwp);
      smem += pix_bytes;
    }
  }
  // work units: (image, group of warps x mc templates).  With staged pixel records a unit is one template per warp;
  // without them a unit is the whole image, so that the pixel data re-read from global memory stays in the CTA's L1.
  g.mc = g.pix_floats > 0 ? 1 : (a->M + warps - 1) / warps;
  g.groups = (a->M + warps * g.mc - 1) / (warps * g.mc);
  g.threads = threads;
  g.smem_bytes = smem;
  int per_sm = (int)((limit + 1024) / (smem + 1024));
  const int by_threads = 2048 / threads;
  if (per_sm > by_threads) per_sm = by_threads;
  const int cap = a->C == 1 ? 4 : 3;               // compiled with __launch_bounds__(256, C == 1 ? 4 : 3)
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  const long slots = (long)sm_count() * per_sm;
  const long units = (long)a->B * g.groups;
  g.grid = units < slots ? (int)units : (int)slots;
  return SCAE_OK;
}

static size_t tmpl_ws_alpha_floats(const scae_tmpl_args* a, int grid) {
  return a->mode == SCAE_TMPL_MODE_ALPHA ? (size_t)grid * a->M * a->h * a->w : 0;
}
static size_t tmpl_ws_raw_floats(const scae_tmpl_args* a, int grid) {
  return a->template_color ? (size_t)grid * a->M * a->C * a->h * a->w : 0;
}

}  // namespace scae

using namespace scae;

extern "C" __attribute__((visibility("default"))) size_t scae_tmpl_ll_bwd_workspace_bytes(const scae_tmpl_args* a) {
  if (tmpl_validate(a) != SCAE_OK) return 0;
  TmplGeom g;
  if (tmpl_bwd_plan(a, &g) != SCAE_OK) return 0;
  return (tmpl_ws_alpha_floats(a, g.grid) + tmpl_ws_raw_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int scae_tmpl_ll_bwd(
    const scae_tmpl_args* a, const float* x, const float* grad_log_prob, const float* cache, float* g_templates,
    float* g_color, float* g_pose, float* g_presence, float* g_bg_image, float* g_alpha, float* g_scalars,
    void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(x && grad_log_prob && cache && g_templates && g_pose && g_scalars, SCAE_EINVAL,
               "tmpl bwd: a required pointer is NULL");
  SCAE_REQUIRE(!a->template_color || g_color, SCAE_EINVAL, "tmpl bwd: g_color is required with template_color");
  TmplGeom g;
  rc = tmpl_bwd_plan(a, &g);
  if (rc != SCAE_OK) return rc;
  const size_t need =
      (tmpl_ws_alpha_floats(a, g.grid) + tmpl_ws_raw_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
  SCAE_REQUIRE(workspace && workspace_bytes >= need, SCAE_EINVAL, "tmpl bwd: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  float* alpha_partials = static_cast<float*>(workspace);
  float* raw_partials = alpha_partials + tmpl_ws_alpha_floats(a, g.grid);
  float* scalar_partials = raw_partials + tmpl_ws_raw_floats(a, g.grid);
  TmplBwdOut out{g_templates, g_color, a->template_color ? raw_partials : nullptr, g_pose, g_presence, g_bg_image,
                 (alpha && g_alpha) ? alpha_partials : nullptr, scalar_partials};
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_ll_bwd_scan_kernel<kC, kA, false>;
    rc = tmpl_prepare_kernel(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, grad_log_prob, cache, out, g);
    note_launch();
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  if (alpha && g_alpha) {
    rc = launch_reduce_rows(alpha_partials, g_alpha, g.grid, a->M * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  if (a->template_color) {
    rc = launch_reduce_rows(raw_partials, g_templates, g.grid, a->M * a->C * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  return launch_reduce_rows(scalar_partials, g_scalars, g.grid, 4, stream);
}

// Backward of pdf.mode() (distributions.py:50-77, straight_through_gradient=False; what SCAE.loss differentiates when
// recon_mse_weight > 0, stacked_capsule_auto_encoder.py:226-230): the gradient w.r.t. the mode image flows, pixel by pixel,
// into the warp of the component the arg-max picked.  `component_cache[B,2,C,H,W]`: planes [b,0,c] hold that component's
// index as a float (M = background; scae_tmpl_render's mode_component, repeated over the channels in alpha mode), planes
// [b,1,c] are ignored.  Same scatter, workspace and outputs as scae_tmpl_ll_bwd; nothing flows to the mixing logits, so
// g_presence / g_alpha are zero and not produced.  g_scalars[4]: only d / d bg_value can be non-zero.
extern "C" __attribute__((visibility("default"))) int scae_tmpl_mode_bwd(
    const scae_tmpl_args* a, const float* grad_mode, const float* component_cache, float* g_templates, float* g_color,
    float* g_pose, float* g_bg_image, float* g_scalars, void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(grad_mode && component_cache && g_templates && g_pose && g_scalars, SCAE_EINVAL,
               "tmpl mode bwd: a required pointer is NULL");
  SCAE_REQUIRE(!a->template_color || g_color, SCAE_EINVAL, "tmpl mode bwd: g_color is required with template_color");
  TmplGeom g;
  rc = tmpl_bwd_plan(a, &g);
  if (rc != SCAE_OK) return rc;
  const size_t need =
      (tmpl_ws_alpha_floats(a, g.grid) + tmpl_ws_raw_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
  SCAE_REQUIRE(workspace && workspace_bytes >= need, SCAE_EINVAL, "tmpl mode bwd: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  float* alpha_partials = static_cast<float*>(workspace);
  float* raw_partials = alpha_partials + tmpl_ws_alpha_floats(a, g.grid);
  float* scalar_partials = raw_partials + tmpl_ws_raw_floats(a, g.grid);
  TmplBwdOut out{g_templates, g_color, a->template_color ? raw_partials : nullptr, g_pose, nullptr, g_bg_image, nullptr,
                 scalar_partials};
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_ll_bwd_scan_kernel<kC, kA, true>;
    rc = tmpl_prepare_kernel(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, grad_mode, grad_mode, component_cache, out, g);
    note_launch();
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  if (a->template_color) {
    rc = launch_reduce_rows(raw_partials, g_templates, g.grid, a->M * a->C * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  return launch_reduce_rows(scalar_partials, g_scalars, g.grid, 4, stream);
}
