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
// pass (profiles/r01a).  Two atomics-free, bit-reproducible formulations are implemented:
//
//  * SCAN (default): one warp owns one (image, template) pair and walks the image row-major, 32 pixels per pass.
//    Along a row the sampling coordinates move on a straight line, so the bilinear cell index is monotone: pixels that
//    fall into the same cell are CONTIGUOUS lanes.  A segmented warp scan (shuffles) pre-reduces the four corner
//    contributions per cell, and only the last lane of each segment does a plain read-modify-write on the warp's private
//    gradient atlas in shared memory -- corner by corner and row by row, so no two lanes ever touch the same address in
//    the same step and the summation order is fixed.  No CTA barriers, no gradient buffer, any image size.
//  * GATHER (SCAE_TMPL_BWD=gather): phase A (pixel-parallel) parks g_loc / g_logit of a template group in shared
//    memory; phase B (texel-parallel) lets every texel invert the affine map, walk its own footprint in the image and
//    accumulate hat(tx - x) * hat(ty - y) * g in registers.  Kept for A/B testing (1.5 ms vs the scan's time at the
//    MNIST config: the footprints of neighbouring texels differ too much for good SIMT efficiency).
#include <stdlib.h>
#include <string.h>

#include "tmpl_common.cuh"

namespace scae {

struct TmplBwdOut {
  float* g_templates;
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
template <int C, bool kAlpha>
__device__ __forceinline__ void bwd_background(const scae_tmpl_args& a, const TmplScalars& sc, const float* xv,
                                               const float* G, const float* Nc, const float* Dc, size_t px0, int HW,
                                               float* g_bg_image, ScalarAcc& acc) {
  const float two_i2s = 2.0f * sc.i2s;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const size_t px = px0 + (size_t)c * HW;
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
template <int C, bool kAlpha, int kPad>
__device__ __forceinline__ Texel<kPad> bwd_pixel(const TmplScalars& sc, const Texel<kPad>& t00, const Texel<kPad>& t10,
                                                 const Texel<kPad>& t01, const Texel<kPad>& t11, const Tap& t,
                                                 float lpres, const float* xv, const float* G, const float* Nc,
                                                 const float* Dc, ScalarAcc& acc, float& glp, float& gtx, float& gty) {
  const float two_i2s = 2.0f * sc.i2s;
  const float gx1 = 1.0f - t.fx, gy1 = 1.0f - t.fy;
  Texel<kPad> out;
#pragma unroll
  for (int c = 0; c < kPad; ++c) out.v[c] = 0.0f;
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

// pose / presence gradients of one template chunk: sum the per-warp partials and write them out
__device__ __forceinline__ void flush_pose(const TmplSmem& s, const scae_tmpl_args& a, const TmplBwdOut& out,
                                           const float* wpart, int nwarps, int mc_stride, int b, int m0, int mc) {
  const float sx = 0.5f * (float)a.w, sy = 0.5f * (float)a.h;
  for (int e = threadIdx.x; e < mc * 7; e += blockDim.x) {
    const int mm = e / 7, q7 = e - mm * 7;
    float t = 0.0f;
    for (int wi = 0; wi < nwarps; ++wi) t += wpart[((size_t)wi * mc_stride + mm) * 8 + q7];
    const int m = m0 + mm;
    if (q7 < 6) {
      out.g_pose[((size_t)b * a.M + m) * 6 + q7] = t * (q7 < 3 ? sx : sy);
    } else if (out.g_presence) {
      const float pr = s.tp[m * 8 + 7];
      out.g_presence[(size_t)b * a.M + m] = pr < kLogSafeEps ? 0.0f : t / pr;
    }
  }
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
// gather variant (default)
// ================================================================================================================
template <int C, bool kAlpha>
__global__ void __launch_bounds__(kTmplThreads, 2) tmpl_ll_bwd_gather_kernel(const scae_tmpl_args a,
                                                                             const float* __restrict__ x,
                                                                             const float* __restrict__ gout,
                                                                             const float* __restrict__ cache,
                                                                             const TmplBwdOut out, const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, PIX = TT::kPixMax;
  extern __shared__ __align__(16) float smem[];
  const TmplSmem s = tmpl_carve(smem, a, g, g.gbuf_floats);
  float* gbuf = s.extra;                           // [mc][H*W][kPad]
  const int nwarps = (blockDim.x + 31) >> 5;
  float* wpart = s.red + 64;                       // [nwarps][mc][8] per-warp partial sums of pose/presence gradients
  tmpl_prologue(s, a, g, g.gbuf_floats);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W, hw = a.h * a.w;
  const int col = threadIdx.x % g.tw, rg = threadIdx.x / g.tw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool thread_ok = (int)threadIdx.x < g.tw * g.k;
  const float lim_x = (float)a.w + 2.5f, lim_y = (float)a.h + 2.5f;
  const unsigned row = (unsigned)(g.pw * kPad), tex_stride = (unsigned)(g.ph * g.pw * kPad);
  const unsigned base0 = 0u - kMagicBits * (row + (unsigned)kPad);
  const float inv_w = 1.0f / (float)a.w, inv_hw = 1.0f / (float)hw;
  const float fH = (float)a.H, fW = (float)a.W;
  const int R = g.split, rshift = 31 - __clz(R);
  float* my_alpha_partial = (kAlpha && out.alpha_partials) ? out.alpha_partials + (size_t)blockIdx.x * a.M * hw : nullptr;
  if (my_alpha_partial)
    for (int e = threadIdx.x; e < a.M * hw; e += blockDim.x) my_alpha_partial[e] = 0.0f;
  ScalarAcc acc;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    load_pose_table(s, a, b);
    for (int m0 = 0; m0 < a.M; m0 += g.mc) {
      const int mc = min(g.mc, a.M - m0);
      __syncthreads();                              // previous group's phase B is done with atlas / gbuf / wpart
      stage_atlas<C, kAlpha>(s.atlas, a, b, m0, mc, g.pw, g.ph);
      for (int e = threadIdx.x; e < nwarps * g.mc * 8; e += blockDim.x) wpart[e] = 0.0f;
      __syncthreads();

      // ---------------- phase A: pixel-parallel ------------------------------------------------------------------
      for (int ty = 0; ty < g.tiles_y; ++ty) {
        for (int tx = 0; tx < g.tiles_x; ++tx) {
          const int j = tx * g.tw + col;
          const bool col_ok = thread_ok && j < a.W;
          const int row0 = ty * g.k * g.ppt + rg;
          const float X = col_ok ? s.xs[j] : 0.0f;
          float Y[PIX], xv[PIX][C], G[PIX][C], Nc[PIX][C], Dc[PIX][C];
          bool ok[PIX];
#pragma unroll
          for (int u = 0; u < PIX; ++u) {
            const int i = row0 + u * g.k;
            ok[u] = col_ok && u < g.ppt && i < a.H;
            Y[u] = ok[u] ? s.ys[i] : 0.0f;
            const size_t px0 = (size_t)b * C * HW + (size_t)i * a.W + j;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const size_t px = px0 + (size_t)c * HW;
              const size_t cx = (size_t)b * 2 * C * HW + (size_t)c * HW + (size_t)i * a.W + j;
              xv[u][c] = ok[u] ? __ldg(x + px) : 0.0f;
              G[u][c] = ok[u] ? __ldg(gout + px) : 0.0f;     // G = 0 switches every contribution of a dead pixel off
              Nc[u][c] = ok[u] ? __ldg(cache + cx) : 0.0f;
              Dc[u][c] = ok[u] ? __ldg(cache + cx + (size_t)C * HW) : 0.0f;
            }
            if (m0 == 0 && ok[u]) bwd_background<C, kAlpha>(a, sc, xv[u], G[u], Nc[u], Dc[u], px0, HW, out.g_bg_image, acc);
          }
          for (int mm = 0; mm < mc; ++mm) {
            const float* t8 = s.tp + (size_t)(m0 + mm) * 8;
            const float4 pa = *reinterpret_cast<const float4*>(t8);
            const float4 pb = *reinterpret_cast<const float4*>(t8 + 4);
            const float cx = fmaf(X, pa.x, pa.z), cy = fmaf(X, pa.w, pb.y);
            const unsigned base = base0 + (unsigned)mm * tex_stride;
            float* gslot = gbuf + (size_t)mm * HW * kPad;
            float sgx = 0.f, sgxy = 0.f, sgy = 0.f, sgyy = 0.f, spres = 0.f;
#pragma unroll
            for (int u = 0; u < PIX; ++u) {
              if (u < g.ppt) {
                Tap t;
                tap_setup<kPad>(fmaf(Y[u], pa.y, cx), fmaf(Y[u], pb.x, cy), lim_x, lim_y, row, base, t);
                const float* q = s.atlas + t.off;
                const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
                const Texel<kPad> t01 = ld_texel<kPad>(q + row), t11 = ld_texel<kPad>(q + row + kPad);
                float glp, gtx, gty;
                const Texel<kPad> gv = bwd_pixel<C, kAlpha, kPad>(sc, t00, t10, t01, t11, t, pb.z, xv[u], G[u], Nc[u],
                                                                  Dc[u], acc, glp, gtx, gty);
                if (ok[u]) st_texel<kPad>(gslot + ((size_t)(row0 + u * g.k) * a.W + j) * kPad, gv);
                sgx += gtx;
                sgxy = fmaf(gtx, Y[u], sgxy);
                sgy += gty;
                sgyy = fmaf(gty, Y[u], sgyy);
                spres += glp;
              }
            }
            // tx = Ax X + Bx Y + Cx with (Ax, Bx, Cx) = (p0, p1, p2 + 1) * w/2 (+ const): the w/2, h/2 factors are
            // applied when the partials are flushed
            float v7[7] = {X * sgx, sgxy, sgx, X * sgy, sgyy, sgy, spres};
#pragma unroll
            for (int q7 = 0; q7 < 7; ++q7) v7[q7] = warp_sum(v7[q7]);
            if (lane == 0) {
              float* wp = wpart + ((size_t)warp * g.mc + mm) * 8;
#pragma unroll
              for (int q7 = 0; q7 < 7; ++q7) wp[q7] += v7[q7];
            }
          }
        }
      }
      __syncthreads();

      // ---------------- phase B: texel-parallel gather -----------------------------------------------------------
      const int items = mc * hw * R;
      for (int base_i = 0; base_i < items; base_i += blockDim.x) {
        const int idx = base_i + threadIdx.x;
        const bool active = idx < items;
        float av[kPad];
#pragma unroll
        for (int c = 0; c < kPad; ++c) av[c] = 0.0f;
        int mm = 0, texel = 0, sub = 0;
        if (active) {
          sub = idx & (R - 1);
          const int tq = idx >> rshift;                              // (mm, texel)
          mm = (int)(((float)tq + 0.5f) * inv_hw);
          texel = tq - mm * hw;
          const int tyi = (int)(((float)texel + 0.5f) * inv_w), txi = texel - tyi * a.w;
          const float* t8 = s.tp + (size_t)(m0 + mm) * 8;
          const float Ax = t8[0], Bx = t8[1], Cx = t8[2], Ay = t8[3], By = t8[4], Cy = t8[5];
          const float txa = (float)(txi + 2), tya = (float)(tyi + 2);
          // Footprint of this texel in the image: invert [tx; ty] = [Ax Bx; Ay By][X; Y] + [Cx; Cy] around the texel
          // (half-width 1 texel per axis) to get a bounding box in normalised image coordinates.  Texels whose box
          // misses the image are skipped outright; the others get their row range from it.
          int i_lo = 0, i_hi = a.H - 1;
          bool visible = true;
          const float det = Ax * By - Bx * Ay;
          if (fabsf(det) > 1e-20f) {
            const float inv = 1.0f / det, ainv = fabsf(inv);
            const float dx = txa - Cx, dy = tya - Cy;
            const float Xc = (By * dx - Bx * dy) * inv, Yc = (Ax * dy - Ay * dx) * inv;
            const float hx = (fabsf(By) + fabsf(Bx)) * ainv, hy = (fabsf(Ay) + fabsf(Ax)) * ainv;
            if (fabsf(Xc) <= 1e6f && fabsf(Yc) <= 1e6f && hx <= 1e6f && hy <= 1e6f) {   // also false for NaN
              const float mx = 2.0f / fW, my = 2.0f / fH;                                // one pixel of slack
              visible = (Xc - hx <= 1.0f + mx) && (Xc + hx >= -1.0f - mx) && (Yc - hy <= 1.0f + my) &&
                        (Yc + hy >= -1.0f - my);
              // row index of Y: i = ((Y + 1) H - 1) / 2
              const float flo = ((Yc - hy + 1.0f) * fH - 1.0f) * 0.5f - 1.0f;
              const float fhi = ((Yc + hy + 1.0f) * fH - 1.0f) * 0.5f + 1.0f;
              i_lo = max(0, (int)floorf(fminf(fmaxf(flo, -2.0f), fH + 1.0f)));
              i_hi = min(a.H - 1, (int)ceilf(fminf(fmaxf(fhi, -2.0f), fH + 1.0f)));
            }
          }
          if (visible) {
            // per row the admissible X form an interval: centre(Y) +- half-width from each axis that depends on X
            const bool has_x = fabsf(Ax) > 1e-12f, has_y = fabsf(Ay) > 1e-12f;
            const float rAx = has_x ? 1.0f / Ax : 0.0f, rAy = has_y ? 1.0f / Ay : 0.0f;
            const float cx0 = (txa - Cx) * rAx, cxs = -Bx * rAx, hwx = has_x ? fabsf(rAx) : 4.0f;
            const float cy0 = (tya - Cy) * rAy, cys = -By * rAy, hwy = has_y ? fabsf(rAy) : 4.0f;
            const float jscale = 0.5f * fW, joff = 0.5f * fW - 0.5f;                     // j = X * W/2 + (W/2 - 1/2)
            const float* gslot = gbuf + (size_t)mm * HW * kPad;
            for (int i = i_lo + sub; i <= i_hi; i += R) {
              const float Yi = s.ys[i];
              const float cxr = fmaf(Yi, Bx, Cx), cyr = fmaf(Yi, By, Cy);
              if (!has_x && !(fabsf(cxr - txa) < 1.0f)) continue;                        // row cannot see the texel
              if (!has_y && !(fabsf(cyr - tya) < 1.0f)) continue;
              const float mx_c = fmaf(Yi, cxs, cx0), my_c = fmaf(Yi, cys, cy0);
              const float lo = fmaxf(fmaxf(mx_c - hwx, my_c - hwy), -2.0f);
              const float hi = fminf(fminf(mx_c + hwx, my_c + hwy), 2.0f);
              if (!(hi >= lo)) continue;
              const int j_lo = max(0, (int)floorf(fmaf(lo, jscale, joff)) - 1);
              const int j_hi = min(a.W - 1, (int)ceilf(fmaf(hi, jscale, joff)) + 1);
              const float* grow = gslot + (size_t)i * a.W * kPad;
              for (int j = j_lo; j <= j_hi; ++j) {
                const float Xj = s.xs[j];
                const float wx = 1.0f - fabsf(fmaf(Xj, Ax, cxr) - txa);
                const float wy = 1.0f - fabsf(fmaf(Xj, Ay, cyr) - tya);
                const float wgt = fmaxf(wx, 0.0f) * fmaxf(wy, 0.0f);
                const Texel<kPad> gv = ld_texel<kPad>(grow + (size_t)j * kPad);
#pragma unroll
                for (int c = 0; c < kPad; ++c) av[c] = fmaf(wgt, gv.v[c], av[c]);
              }
            }
          }
        }
        // combine the R row-interleaved partial sums of a texel (adjacent lanes)
        for (int d = 1; d < R; d <<= 1) {
#pragma unroll
          for (int c = 0; c < kPad; ++c) av[c] += __shfl_xor_sync(0xffffffffu, av[c], d);
        }
        if (active && sub == 0) {
          const int m = m0 + mm;
#pragma unroll
          for (int c = 0; c < C; ++c) out.g_templates[(((size_t)b * a.M + m) * C + c) * hw + texel] = av[c];
          if (kAlpha && my_alpha_partial) my_alpha_partial[(size_t)m * hw + texel] += av[C];
        }
      }
      flush_pose(s, a, out, wpart, nwarps, g.mc, b, m0, mc);
    }
  }
  write_scalar_partials(a, sc, kAlpha, acc, s.red, out.scalar_partials + (size_t)blockIdx.x * 4);
}

// ================================================================================================================
// scan variant (default): warp per (image, template), segmented-scan scatter
// ================================================================================================================
constexpr int kScanThreads = 256;

template <int C, bool kAlpha>
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
  // [C][H*W] records {x, upstream gradient, cached numerator lse, cached denominator lse} of the current image: every
  // warp of the CTA works on the same image, so the four global loads (and their 64-bit address arithmetic) that each
  // (template, pass) used to repeat are paid once per image
  const bool staged = g.pix_floats > 0;
  float4* PIX = reinterpret_cast<float4*>(smem + (((size_t)nwarps * 2 * atlas_floats + 64 + a.W + a.H + 3) & ~(size_t)3));
  for (int e = lane; e < 2 * atlas_floats; e += 32) atlas[e] = 0.0f;
  for (int e = threadIdx.x; e < a.W; e += blockDim.x) xs[e] = base_coord(e, a.W);
  for (int e = threadIdx.x; e < a.H; e += blockDim.x) ys[e] = base_coord(e, a.H);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W, hw = a.h * a.w, pw = g.pw, ph = g.ph;
  const float lim_x = (float)a.w + 2.5f, lim_y = (float)a.h + 2.5f;
  const unsigned row = (unsigned)(pw * kPad);
  const unsigned base0 = 0u - kMagicBits * (row + (unsigned)kPad);
  const float hw_x = 0.5f * (float)a.w, hw_y = 0.5f * (float)a.h;
  const float inv_pw = 1.0f / (float)pw;
  float* my_alpha_partial = (kAlpha && out.alpha_partials) ? out.alpha_partials + (size_t)blockIdx.x * a.M * hw : nullptr;
  if (my_alpha_partial)
    for (int e = threadIdx.x; e < a.M * hw; e += blockDim.x) my_alpha_partial[e] = 0.0f;
  ScalarAcc acc;
  __syncthreads();   // partial row zeroed before any warp accumulates into it

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    // background component, once per pixel -- spread over the whole CTA -- and the pixel records
    if (staged) __syncthreads();                       // the previous image's records are no longer read
    if (staged || warp == 0) {
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
          if (staged) PIX[c * HW + p] = make_float4(xv[c], G[c], Nc[c], Dc[c]);
        }
        bwd_background<C, kAlpha>(a, sc, xv, G, Nc, Dc, px0, HW, out.g_bg_image, acc);
      }
    }
    if (staged) __syncthreads();
    for (int m = warp; m < a.M; m += nwarps) {
      // ---- per-template setup (all lanes compute the same coefficients) -----------------------------------------
      const float* pp = a.pose + ((size_t)b * a.M + m) * 6;
      const float Ax = __ldg(pp + 0) * hw_x, Bx = __ldg(pp + 1) * hw_x, Cx = (__ldg(pp + 2) + 1.0f) * hw_x + 1.5f;
      const float Ay = __ldg(pp + 3) * hw_y, By = __ldg(pp + 4) * hw_y, Cy = (__ldg(pp + 5) + 1.0f) * hw_y + 1.5f;
      const float pres = a.presence ? __ldg(a.presence + (size_t)b * a.M + m) : 1.0f;
      const float lpres = a.presence ? log_safe_f(pres) : 0.0f;
      {
        const float* src = a.templates + ((size_t)b * a.M + m) * C * hw;
        const float inv_w = 1.0f / (float)a.w;
        for (int e = lane; e < hw; e += 32) {
          const int y = (int)(((float)e + 0.5f) * inv_w), xx = e - y * a.w;
          float* q = atlas + ((size_t)(y + 2) * pw + (xx + 2)) * kPad;
#pragma unroll
          for (int c = 0; c < C; ++c) q[c] = __ldg(src + (size_t)c * hw + e);
          if (kAlpha) q[C] = __ldg(a.templates_alpha + (size_t)m * hw + e);
        }
      }
      __syncwarp();
      float sgx = 0.f, sgxX = 0.f, sgxY = 0.f, sgy = 0.f, sgyX = 0.f, sgyY = 0.f, spres = 0.f;

      // ---- passes of 32 consecutive pixels (row-major) -----------------------------------------------------------
      int i = lane / a.W, j = lane - i * a.W;
      const int di = 32 / a.W, dj = 32 - di * a.W;
      for (int p0 = 0; p0 < HW; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < HW;
        float xv[C], G[C], Nc[C], Dc[C];
        if (staged) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 r4 = valid ? PIX[c * HW + p] : make_float4(0.f, 0.f, 0.f, 0.f);   // G = 0 switches a dead lane off
            xv[c] = r4.x;
            G[c] = r4.y;
            Nc[c] = r4.z;
            Dc[c] = r4.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const size_t px = (size_t)b * C * HW + (size_t)c * HW + p;
            const size_t cx = (size_t)b * 2 * C * HW + (size_t)c * HW + p;
            xv[c] = valid ? __ldg(x + px) : 0.0f;
            G[c] = valid ? __ldg(gout + px) : 0.0f;
            Nc[c] = valid ? __ldg(cache + cx) : 0.0f;
            Dc[c] = valid ? __ldg(cache + cx + (size_t)C * HW) : 0.0f;
          }
        }
        const float X = xs[valid ? j : 0], Y = ys[valid ? i : 0];
        Tap t;
        tap_setup<kPad>(fmaf(Y, Bx, fmaf(X, Ax, Cx)), fmaf(Y, By, fmaf(X, Ay, Cy)), lim_x, lim_y, row, base0, t);
        const float* q = atlas + t.off;
        const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
        const Texel<kPad> t01 = ld_texel<kPad>(q + row), t11 = ld_texel<kPad>(q + row + kPad);
        float glp, gtx, gty;
        const Texel<kPad> gv = bwd_pixel<C, kAlpha, kPad>(sc, t00, t10, t01, t11, t, lpres, xv, G, Nc, Dc, acc, glp, gtx, gty);
        sgx += gtx;
        sgxX = fmaf(gtx, X, sgxX);
        sgxY = fmaf(gtx, Y, sgxY);
        sgy += gty;
        sgyX = fmaf(gty, X, sgyX);
        sgyY = fmaf(gty, Y, sgyY);
        spres += glp;

        // ---- segmented scan over lanes that share (row, cell) -------------------------------------------------------
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
        const int i_prev = __shfl_up_sync(0xffffffffu, i, 1);
        const bool head = lane == 0 || key != key_prev || i != i_prev;
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
        const int i_first = __shfl_sync(0xffffffffu, i, 0);
        const int i_last = __shfl_sync(0xffffffffu, valid ? i : -1, 31 - __clz(__ballot_sync(0xffffffffu, valid)));
        float* gq = gatlas + t.off;
        // A pass of 32 consecutive pixels can straddle image rows.  Within one row the cells of different segments are
        // distinct; across rows they could coincide (extreme magnification), so check once with MATCH and only then
        // fall back to updating row by row.
        bool by_row = false;
        if (i_first != i_last) {
          const unsigned tails = __ballot_sync(0xffffffffu, tail);
          const unsigned peers = __match_any_sync(0xffffffffu, tail ? key : (0xFFFFFF00u | (unsigned)lane));
          by_row = __any_sync(0xffffffffu, tail && (peers & tails & ~(1u << lane)) != 0u);
        }
        if (!by_row) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (tail) {
              float* dst = gq + (k & 1 ? kPad : 0) + (k & 2 ? row : 0);
#pragma unroll
              for (int c = 0; c < NCH; ++c) dst[c] += v[k][c];
            }
            __syncwarp();
          }
        } else {
          for (int r = i_first; r <= i_last; ++r) {
            const bool mine = tail && i == r;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (mine) {
                float* dst = gq + (k & 1 ? kPad : 0) + (k & 2 ? row : 0);
#pragma unroll
                for (int c = 0; c < NCH; ++c) dst[c] += v[k][c];
              }
              __syncwarp();
            }
          }
        }
        // advance to the next pass
        i += di;
        j += dj;
        if (j >= a.W) {
          j -= a.W;
          ++i;
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
        float* dst = out.g_templates + ((size_t)b * a.M + m) * C * hw;
        for (int e = lane; e < pw * ph; e += 32) {
          const int yy = (int)(((float)e + 0.5f) * inv_pw), xx = e - yy * pw;
          float* q = gatlas + (size_t)e * kPad;
          if (yy >= 2 && yy < ph - 2 && xx >= 2 && xx < pw - 2) {
            const int te = (yy - 2) * a.w + (xx - 2);
#pragma unroll
            for (int c = 0; c < C; ++c) dst[(size_t)c * hw + te] = q[c];
            if (kAlpha && my_alpha_partial) my_alpha_partial[(size_t)m * hw + te] += q[C];
          }
#pragma unroll
          for (int c = 0; c < kPad; ++c) q[c] = 0.0f;
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
static const size_t kBwdSmemBudget = 100 * 1024;   // gather variant: two CTAs per SM

struct BwdPlan {
  TmplGeom g;
  bool gather;
};

static bool want_gather() {
  const char* e = getenv("SCAE_TMPL_BWD");
  return e != nullptr && strcmp(e, "gather") == 0;
}

static int tmpl_bwd_plan(const scae_tmpl_args* a, BwdPlan* p) {
  const int kpad = tmpl_texel_floats(a);
  const size_t limit = (size_t)max_smem_optin();
  p->gather = false;
  if (want_gather()) {
    const size_t warps = kTmplThreads / 32;
    const size_t wpart_bytes = warps * a->M * 8 * sizeof(float);        // budgeted for the worst case mc = M
    const size_t gbuf_tmpl = (size_t)a->H * a->W * kpad * sizeof(float);
    const size_t atlas_tmpl = (size_t)(a->w + 4) * (a->h + 4) * kpad * sizeof(float);
    const size_t fixed = wpart_bytes + ((size_t)a->M * 8 + a->W + a->H + 64 + 8) * sizeof(float) + 64;
    if (fixed + gbuf_tmpl + atlas_tmpl <= limit) {
      int rc = tmpl_geometry(a, gbuf_tmpl, wpart_bytes, kBwdSmemBudget, &p->g);
      if (rc != SCAE_OK) return rc;
      p->g.gbuf_floats = (int)(((size_t)p->g.mc * a->H * a->W * kpad + 3) / 4 * 4);
      p->g.smem_bytes += 16;
      p->g.split = 4;
      p->gather = true;
      return SCAE_OK;
    }
  }
  // scan variant: one padded value atlas + one gradient atlas per warp, nothing that scales with the image
  TmplGeom& g = p->g;
  memset(&g, 0, sizeof(g));
  g.pw = a->w + 4;
  g.ph = a->h + 4;
  g.mc = 1;
  g.atlas_floats = (int)(((size_t)g.pw * g.ph * kpad + 3) / 4 * 4);
  int threads = kScanThreads;
  size_t smem;
  for (;;) {
    smem = ((size_t)(threads / 32) * 2 * g.atlas_floats + 64 + a->W + a->H + 8) * sizeof(float);
    if (smem <= limit || threads == 32) break;
    threads /= 2;                                   // very large templates: fewer warps per CTA
  }
  SCAE_REQUIRE(smem <= limit, SCAE_ELIMIT, "tmpl bwd: a %dx%d template does not fit in shared memory", a->h, a->w);
  // per-image pixel records in shared memory when at least two CTAs per SM still fit (each CTA also pays 1 KB reserved)
  {
    const size_t pix_bytes = (size_t)4 * a->C * a->H * a->W * sizeof(float) + 16;
    const char* e = getenv("SCAE_TMPL_BWD_STAGE");
    const bool allow = e == nullptr || strcmp(e, "0") != 0;
    if (allow && 2 * (smem + pix_bytes + 1024) <= limit + 1024) {
      g.pix_floats = 4 * a->C * a->H * a->W;
      smem += pix_bytes;
    }
  }
  g.threads = threads;
  g.smem_bytes = smem;
  int per_sm = (int)(limit / smem);
  const int by_threads = 2048 / threads;
  if (per_sm > by_threads) per_sm = by_threads;
  const int cap = a->C == 1 ? 4 : 3;               // compiled with __launch_bounds__(256, C == 1 ? 4 : 3)
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  const long slots = (long)sm_count() * per_sm;
  g.grid = a->B < slots ? a->B : (int)slots;
  return SCAE_OK;
}

static size_t tmpl_ws_alpha_floats(const scae_tmpl_args* a, int grid) {
  return a->mode == SCAE_TMPL_MODE_ALPHA ? (size_t)grid * a->M * a->h * a->w : 0;
}

}  // namespace scae

using namespace scae;

extern "C" __attribute__((visibility("default"))) size_t scae_tmpl_ll_bwd_workspace_bytes(const scae_tmpl_args* a) {
  if (tmpl_validate(a) != SCAE_OK) return 0;
  BwdPlan p;
  if (tmpl_bwd_plan(a, &p) != SCAE_OK) return 0;
  return (tmpl_ws_alpha_floats(a, p.g.grid) + (size_t)p.g.grid * 4) * sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int scae_tmpl_ll_bwd(
    const scae_tmpl_args* a, const float* x, const float* grad_log_prob, const float* cache, float* g_templates,
    float* g_pose, float* g_presence, float* g_bg_image, float* g_alpha, float* g_scalars, void* workspace,
    size_t workspace_bytes, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(x && grad_log_prob && cache && g_templates && g_pose && g_scalars, SCAE_EINVAL,
               "tmpl bwd: a required pointer is NULL");
  BwdPlan p;
  rc = tmpl_bwd_plan(a, &p);
  if (rc != SCAE_OK) return rc;
  const TmplGeom& g = p.g;
  const size_t need = (tmpl_ws_alpha_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
  SCAE_REQUIRE(workspace && workspace_bytes >= need, SCAE_EINVAL, "tmpl bwd: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  float* alpha_partials = static_cast<float*>(workspace);
  float* scalar_partials = alpha_partials + tmpl_ws_alpha_floats(a, g.grid);
  TmplBwdOut out{g_templates, g_pose, g_presence, g_bg_image, (alpha && g_alpha) ? alpha_partials : nullptr,
                 scalar_partials};
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    if (p.gather) {
      auto kern = tmpl_ll_bwd_gather_kernel<kC, kA>;
      rc = tmpl_prepare_kernel(kern, g.smem_bytes);
      if (rc != SCAE_OK) return rc;
      kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, grad_log_prob, cache, out, g);
      note_launch();
    } else {
      auto kern = tmpl_ll_bwd_scan_kernel<kC, kA>;
      rc = tmpl_prepare_kernel(kern, g.smem_bytes);
      if (rc != SCAE_OK) return rc;
      kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, grad_log_prob, cache, out, g);
      note_launch();
    }
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  if (alpha && g_alpha) {
    rc = launch_reduce_rows(alpha_partials, g_alpha, g.grid, a->M * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  return launch_reduce_rows(scalar_partials, g_scalars, g.grid, 4, stream);
}
