// Hot path 1 (sm_100a): template warp + per-pixel template-mixture Gaussian log-likelihood -- forward, render and
// backward kernels.
//
// Replaces TemplateBasedImageDecoder.forward (reference part_decoder.py:152-243) fused with GaussianMixture.log_prob
// (distributions.py:41-48).  Math: oracle/template_likelihood.py::closed_form_log_prob (forward) and
// oracle/manual_backward.py::template_forward_backward (backward), which this file transcribes.
//
// One persistent CTA per image slot: the M templates of the image (and the shared alpha logits) are staged into a
// zero-bordered shared-memory atlas, every thread owns a few pixels of one image column and walks the templates,
// folding each component into streaming-logsumexp state held in registers.  The B x (M+1) x C x H x W warped-template
// and mixing-logit tensors of the reference are never formed; HBM sees templates, poses, the target image and one
// log-prob per pixel.  Roofline discussion (this path is issue/LDS bound, not HBM bound): DESIGN.md section 5.
#include "tmpl_common.cuh"

namespace scae {

constexpr int kMaxThreads = 512;

// ================================================================================================================
// host: validation + geometry
// ================================================================================================================
int tmpl_validate(const scae_tmpl_args* a) {
  SCAE_REQUIRE(a != nullptr, SCAE_EINVAL, "tmpl: args is NULL");
  SCAE_REQUIRE(a->B > 0 && a->M > 0 && a->C > 0 && a->h > 0 && a->w > 0 && a->H > 0 && a->W > 0, SCAE_EINVAL,
               "tmpl: all of B,M,C,h,w,H,W must be positive");
  SCAE_REQUIRE(a->C <= 3, SCAE_ELIMIT, "tmpl: C=%d channels not supported (max 3)", a->C);
  SCAE_REQUIRE(a->mode == SCAE_TMPL_MODE_ALPHA || a->mode == SCAE_TMPL_MODE_TEMPERATURE, SCAE_EINVAL,
               "tmpl: unknown mode %d", a->mode);
  SCAE_REQUIRE(a->templates && a->pose, SCAE_EINVAL, "tmpl: templates and pose are required");
  if (a->mode == SCAE_TMPL_MODE_ALPHA) {
    SCAE_REQUIRE(a->templates_alpha && a->bg_mixing_logit, SCAE_EINVAL,
                 "tmpl: alpha mode needs templates_alpha and bg_mixing_logit");
  } else {
    SCAE_REQUIRE(a->temperature_logit, SCAE_EINVAL, "tmpl: temperature mode needs temperature_logit");
  }
  // part_decoder.py:192 dereferences self.bg_value when no bg_image is given
  SCAE_REQUIRE(a->bg_image || a->bg_value, SCAE_EINVAL, "tmpl: need bg_image or bg_value");
  SCAE_REQUIRE(a->h <= 1024 && a->w <= 1024 && a->H <= 65535 && a->W <= 65535, SCAE_ELIMIT, "tmpl: sizes too large");
  return SCAE_OK;
}

static int texel_floats(const scae_tmpl_args* a) {
  const int ch = a->C + (a->mode == SCAE_TMPL_MODE_ALPHA ? 1 : 0);
  return ch <= 1 ? 1 : (ch <= 2 ? 2 : 4);
}

int tmpl_geometry(const scae_tmpl_args* a, int atlas_copies, size_t extra_smem_bytes, int ctas_per_sm_target,
                  TmplGeom* g) {
  const int pixmax = a->C == 1 ? 8 : 4;
  const int tw = a->W < 64 ? a->W : 64;
  int best_k = 1;
  double best_score = -1.0;
  for (int k = 1; k <= a->H && tw * k <= kMaxThreads; ++k) {
    const int threads = ((tw * k + 31) / 32) * 32;
    int ppt = (a->H + k - 1) / k, tiles_y = 1;
    if (ppt > pixmax) {
      ppt = pixmax;
      tiles_y = (a->H + k * pixmax - 1) / (k * pixmax);
    }
    const double eff = (double)a->H * tw / ((double)threads * ppt * tiles_y);
    // prefer 192..384-thread CTAs (several per SM) and few pixel tiles
    double score = eff - 0.03 * (tiles_y - 1);
    if (threads < 128) score -= 0.10;
    if (threads > 384) score -= 0.02;
    if (score > best_score + 1e-9) {
      best_score = score;
      best_k = k;
    }
  }
  g->tw = tw;
  g->k = best_k;
  g->threads = ((tw * best_k + 31) / 32) * 32;
  g->ppt = (a->H + best_k - 1) / best_k;
  g->tiles_y = 1;
  if (g->ppt > pixmax) {
    g->ppt = pixmax;
    g->tiles_y = (a->H + best_k * pixmax - 1) / (best_k * pixmax);
  }
  g->tiles_x = (a->W + tw - 1) / tw;
  g->pw = a->w + 4;
  g->ph = a->h + 4;

  const size_t per_tmpl = (size_t)g->pw * g->ph * texel_floats(a) * sizeof(float) * atlas_copies;
  const size_t fixed = extra_smem_bytes + ((size_t)a->M * 8 + a->W + a->H + 64) * sizeof(float);
  const size_t limit = (size_t)max_smem_optin();
  size_t budget = limit / (ctas_per_sm_target > 0 ? ctas_per_sm_target : 1);
  if (budget > limit) budget = limit;
  SCAE_REQUIRE(fixed + per_tmpl <= limit, SCAE_ELIMIT,
               "tmpl: one %dx%d template (+%zu fixed bytes) does not fit in %zu bytes of shared memory", a->h, a->w,
               fixed, limit);
  long mc = budget > fixed ? (long)((budget - fixed) / per_tmpl) : 0;
  if (mc < 1) mc = 1;                       // fall back to one CTA per SM with as many templates as fit
  if (mc > a->M) mc = a->M;
  // balance the chunks
  const int nchunks = (a->M + (int)mc - 1) / (int)mc;
  mc = (a->M + nchunks - 1) / nchunks;
  g->mc = (int)mc;
  g->atlas_floats = (int)((((size_t)g->pw * g->ph * texel_floats(a) * g->mc) + 3) / 4 * 4);
  g->smem_bytes = fixed + (size_t)g->atlas_floats * sizeof(float) * atlas_copies;
  int per_sm = (int)(limit / g->smem_bytes);
  const int by_threads = 2048 / g->threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  const long slots = (long)sm_count() * per_sm;
  g->grid = a->B < slots ? a->B : (int)slots;
  return SCAE_OK;
}

// ================================================================================================================
// device: shared prologue
// ================================================================================================================
struct TmplSmem {
  float* atlas;
  float* gatlas;   // backward only
  float* tp;       // [M][8]: pose[6], log_safe(presence), presence
  float* xs;       // [W]
  float* ys;       // [H]
  float* red;      // [64]
};

__device__ __forceinline__ TmplSmem tmpl_carve(float* smem, const scae_tmpl_args& a, const TmplGeom& g, int kPad,
                                               bool with_grad) {
  TmplSmem s;
  const size_t atlas_floats = (size_t)g.atlas_floats;
  s.atlas = smem;
  s.gatlas = with_grad ? smem + atlas_floats : nullptr;
  s.tp = smem + atlas_floats * (with_grad ? 2 : 1);
  s.xs = s.tp + (size_t)a.M * 8;
  s.ys = s.xs + a.W;
  s.red = s.ys + a.H;
  return s;
}

__device__ __forceinline__ void tmpl_prologue(const TmplSmem& s, const scae_tmpl_args& a, const TmplGeom& g, int kPad,
                                              bool with_grad) {
  const size_t atlas_floats = (size_t)g.atlas_floats * (with_grad ? 2 : 1);
  for (size_t i = threadIdx.x; i < atlas_floats; i += blockDim.x) s.atlas[i] = 0.0f;
  for (int i = threadIdx.x; i < a.W; i += blockDim.x) s.xs[i] = base_coord(i, a.W);
  for (int i = threadIdx.x; i < a.H; i += blockDim.x) s.ys[i] = base_coord(i, a.H);
}

__device__ __forceinline__ void load_pose_table(const TmplSmem& s, const scae_tmpl_args& a, int b) {
  for (int i = threadIdx.x; i < a.M; i += blockDim.x) {
    const float* p = a.pose + ((size_t)b * a.M + i) * 6;
#pragma unroll
    for (int q = 0; q < 6; ++q) s.tp[i * 8 + q] = __ldg(p + q);
    const float pr = a.presence ? __ldg(a.presence + (size_t)b * a.M + i) : 1.0f;
    s.tp[i * 8 + 6] = a.presence ? log_safe_f(pr) : 0.0f;
    s.tp[i * 8 + 7] = pr;
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

template <int kPad>
__device__ __forceinline__ float bilerp(const Texel<kPad>& t00, const Texel<kPad>& t10, const Texel<kPad>& t01,
                                        const Texel<kPad>& t11, const Tap& t, int c) {
  return fmaf(t11.v[c], t.w11, fmaf(t01.v[c], t.w01, fmaf(t10.v[c], t.w10, t00.v[c] * t.w00)));
}

// ================================================================================================================
// forward
// ================================================================================================================
template <int C, bool kAlpha>
__global__ void __launch_bounds__(kMaxThreads) tmpl_ll_fwd_kernel(const scae_tmpl_args a, const float* __restrict__ x,
                                                                  float* __restrict__ logp, float* __restrict__ ll,
                                                                  float* __restrict__ cache, const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, PIX = TT::kPixMax, ND = kAlpha ? 1 : C;
  extern __shared__ __align__(16) float smem[];
  const TmplSmem s = tmpl_carve(smem, a, g, kPad, false);
  tmpl_prologue(s, a, g, kPad, false);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W;
  const int col = threadIdx.x % g.tw, rg = threadIdx.x / g.tw;
  const bool thread_ok = (int)threadIdx.x < g.tw * g.k;
  const float wf = (float)a.w, hf = (float)a.h;
  const int tex_stride = g.ph * g.pw * kPad, row_stride = g.pw * kPad;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    load_pose_table(s, a, b);
    float ll_acc = 0.0f;
    for (int ty = 0; ty < g.tiles_y; ++ty) {
      for (int tx = 0; tx < g.tiles_x; ++tx) {
        const int j = tx * g.tw + col;
        const bool col_ok = thread_ok && j < a.W;
        const int row0 = ty * g.k * g.ppt + rg;
        __syncthreads();   // xs/ys/tp visible
        const float X = col_ok ? s.xs[j] : 0.0f;
        float Y[PIX], xv[PIX][C];
        Lse N[PIX][C], D[PIX][ND];
        bool ok[PIX];
#pragma unroll
        for (int u = 0; u < PIX; ++u) {
          const int i = row0 + u * g.k;
          ok[u] = col_ok && u < g.ppt && i < a.H;
          Y[u] = ok[u] ? s.ys[i] : 0.0f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const size_t px = ((size_t)b * C + c) * HW + (size_t)i * a.W + j;
            xv[u][c] = ok[u] ? __ldg(x + px) : 0.0f;
            const float bg = (ok[u] && a.bg_image) ? __ldg(a.bg_image + px) : sc.bg_loc;
            const float bl = kAlpha ? sc.bg_logit : bg * sc.inv_tau;   // bg presence is 1: log_safe adds 0
            const float d = xv[u][c] - bg;
            N[u][c].init(fmaf(d * d, -sc.i2s, bl));
            if (!kAlpha) D[u][c].init(bl);
          }
          if (kAlpha) D[u][0].init(sc.bg_logit);
        }
        for (int m0 = 0; m0 < a.M; m0 += g.mc) {
          const int mc = min(g.mc, a.M - m0);
          __syncthreads();
          stage_atlas<C, kAlpha>(s.atlas, a, b, m0, mc, g.pw, g.ph, blockDim.x);
          __syncthreads();
          for (int mm = 0; mm < mc; ++mm) {
            const float* t8 = s.tp + (size_t)(m0 + mm) * 8;
            const float4 pa = *reinterpret_cast<const float4*>(t8);
            const float4 pb = *reinterpret_cast<const float4*>(t8 + 4);
            const float cx = fmaf(X, pa.x, pa.z);     // p0*X + p2
            const float cy = fmaf(X, pa.w, pb.y);     // p3*X + p5
            const float lpres = pb.z;
            const float* tex = s.atlas + (size_t)mm * tex_stride;
#pragma unroll
            for (int u = 0; u < PIX; ++u) {
              if (u < g.ppt) {
                Tap t;
                tap_setup(fmaf(Y[u], pa.y, cx), fmaf(Y[u], pb.x, cy), wf, hf, g.pw, t);
                const float* q = tex + t.off * kPad;
                const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
                const Texel<kPad> t01 = ld_texel<kPad>(q + row_stride), t11 = ld_texel<kPad>(q + row_stride + kPad);
                float al = 0.0f;
                if (kAlpha) {
                  al = bilerp<kPad>(t00, t10, t01, t11, t, C) + lpres;
                  D[u][0].push(al);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  const float loc = bilerp<kPad>(t00, t10, t01, t11, t, c);
                  const float d = xv[u][c] - loc;
                  const float logit = kAlpha ? al : fmaf(loc, sc.inv_tau, lpres);
                  N[u][c].push(fmaf(d * d, -sc.i2s, logit));
                  if (!kAlpha) D[u][c].push(logit);
                }
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < PIX; ++u) {
          if (ok[u]) {
            const int i = row0 + u * g.k;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float nv = N[u][c].value(), dv = D[u][kAlpha ? 0 : c].value();
              const float lp = nv - dv - sc.log_norm;
              const size_t px = ((size_t)b * C + c) * HW + (size_t)i * a.W + j;
              logp[px] = lp;
              if (cache) {
                cache[((size_t)b * 2 * C + c) * HW + (size_t)i * a.W + j] = nv;
                cache[((size_t)b * 2 * C + C + c) * HW + (size_t)i * a.W + j] = dv;
              }
              ll_acc += lp;
            }
          }
        }
      }
    }
    if (ll) {
      const float t = block_sum(ll_acc, s.red);
      if (threadIdx.x == 0) ll[b] = t;
    }
  }
}

// ================================================================================================================
// render: transformed_templates, mixing_logits, mode, mean (no gradients)
// ================================================================================================================
template <int C, bool kAlpha>
__global__ void __launch_bounds__(kMaxThreads) tmpl_render_kernel(const scae_tmpl_args a, float* __restrict__ tt,
                                                                  float* __restrict__ ml, float* __restrict__ mode,
                                                                  float* __restrict__ mean, const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, ND = kAlpha ? 1 : C;
  extern __shared__ __align__(16) float smem[];
  const TmplSmem s = tmpl_carve(smem, a, g, kPad, false);
  tmpl_prologue(s, a, g, kPad, false);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W, K = a.M + 1, CL = kAlpha ? 1 : C;
  const int col = threadIdx.x % g.tw, rg = threadIdx.x / g.tw;
  const bool thread_ok = (int)threadIdx.x < g.tw * g.k;
  const float wf = (float)a.w, hf = (float)a.h;
  const int tex_stride = g.ph * g.pw * kPad, row_stride = g.pw * kPad;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    load_pose_table(s, a, b);
    for (int ty = 0; ty < g.tiles_y; ++ty)
      for (int tx = 0; tx < g.tiles_x; ++tx)
        for (int u = 0; u < g.ppt; ++u) {
          const int j = tx * g.tw + col;
          const int i = ty * g.k * g.ppt + rg + u * g.k;
          const bool ok = thread_ok && j < a.W && i < a.H;
          __syncthreads();
          const float X = ok ? s.xs[j] : 0.0f, Y = ok ? s.ys[i] : 0.0f;
          const size_t pix = (size_t)i * a.W + j;
          // streaming state seeded with the background component (index M)
          float bgv[C], best_logit[ND], best_loc[C], acc[C];
          Lse L[ND];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            bgv[c] = (ok && a.bg_image) ? __ldg(a.bg_image + ((size_t)b * C + c) * HW + pix) : sc.bg_loc;
            best_loc[c] = bgv[c];
            acc[c] = bgv[c];
          }
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            best_logit[d] = kAlpha ? sc.bg_logit : bgv[d] * sc.inv_tau;
            L[d].init(best_logit[d]);
          }
          // argmax ties go to the lowest component index (torch.argmax); bg has the highest index, so templates win ties
          bool bg_best[ND];
#pragma unroll
          for (int d = 0; d < ND; ++d) bg_best[d] = true;
          for (int m0 = 0; m0 < a.M; m0 += g.mc) {
            const int mc = min(g.mc, a.M - m0);
            __syncthreads();
            stage_atlas<C, kAlpha>(s.atlas, a, b, m0, mc, g.pw, g.ph, blockDim.x);
            __syncthreads();
            for (int mm = 0; mm < mc; ++mm) {
              const int m = m0 + mm;
              const float* t8 = s.tp + (size_t)m * 8;
              Tap t;
              tap_setup(fmaf(Y, t8[1], fmaf(X, t8[0], t8[2])), fmaf(Y, t8[4], fmaf(X, t8[3], t8[5])), wf, hf, g.pw, t);
              const float* q = s.atlas + (size_t)mm * tex_stride + t.off * kPad;
              const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
              const Texel<kPad> t01 = ld_texel<kPad>(q + row_stride), t11 = ld_texel<kPad>(q + row_stride + kPad);
              const float lpres = t8[6];
              float loc[C], logit[ND];
#pragma unroll
              for (int c = 0; c < C; ++c) loc[c] = bilerp<kPad>(t00, t10, t01, t11, t, c);
              if (kAlpha) {
                logit[0] = bilerp<kPad>(t00, t10, t01, t11, t, C) + lpres;
              } else {
#pragma unroll
                for (int c = 0; c < C; ++c) logit[c] = fmaf(loc[c], sc.inv_tau, lpres);
              }
              if (ok) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                  if (tt) tt[(((size_t)b * K + m) * C + c) * HW + pix] = loc[c];
#pragma unroll
                for (int d = 0; d < ND; ++d)
                  if (ml) ml[(((size_t)b * K + m) * CL + d) * HW + pix] = logit[d];
              }
#pragma unroll
              for (int d = 0; d < ND; ++d) {
                float resc;
                const float wn = L[d].push(logit[d], resc);
                const bool better = bg_best[d] ? (logit[d] >= best_logit[d]) : (logit[d] > best_logit[d]);
                if (kAlpha) {
#pragma unroll
                  for (int c = 0; c < C; ++c) {
                    acc[c] = fmaf(acc[c], resc, wn * loc[c]);
                    if (better) best_loc[c] = loc[c];
                  }
                } else {
                  acc[d] = fmaf(acc[d], resc, wn * loc[d]);
                  if (better) best_loc[d] = loc[d];
                }
                if (better) {
                  best_logit[d] = logit[d];
                  bg_best[d] = false;
                }
              }
            }
          }
          if (ok) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
              if (tt) tt[(((size_t)b * K + a.M) * C + c) * HW + pix] = bgv[c];
              if (mode) mode[((size_t)b * C + c) * HW + pix] = best_loc[c];
              if (mean) mean[((size_t)b * C + c) * HW + pix] = acc[c] * __frcp_rn(L[kAlpha ? 0 : c].s);
            }
#pragma unroll
            for (int d = 0; d < ND; ++d)
              if (ml) ml[(((size_t)b * K + a.M) * CL + d) * HW + pix] = kAlpha ? sc.bg_logit : bgv[d] * sc.inv_tau;
          }
        }
  }
}

// ================================================================================================================
// backward
// ================================================================================================================
struct TmplBwdOut {
  float* g_templates;
  float* g_pose;
  float* g_presence;
  float* g_bg_image;
  float* alpha_partials;    // [grid][M*h*w]   (alpha mode)
  float* scalar_partials;   // [grid][4]
};

template <int C, bool kAlpha>
__global__ void __launch_bounds__(kMaxThreads) tmpl_ll_bwd_kernel(const scae_tmpl_args a, const float* __restrict__ x,
                                                                  const float* __restrict__ gout,
                                                                  const float* __restrict__ cache, const TmplBwdOut out,
                                                                  const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, PIX = TT::kPixMax;
  extern __shared__ __align__(16) float smem[];
  const TmplSmem s = tmpl_carve(smem, a, g, kPad, true);
  const int nwarps = (blockDim.x + 31) >> 5;
  float* wpart = s.red + 64;                       // [nwarps][mc][8] per-warp partial sums of pose/presence gradients
  tmpl_prologue(s, a, g, kPad, true);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W, hw = a.h * a.w;
  const int col = threadIdx.x % g.tw, rg = threadIdx.x / g.tw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool thread_ok = (int)threadIdx.x < g.tw * g.k;
  const float wf = (float)a.w, hf = (float)a.h;
  const int tex_stride = g.ph * g.pw * kPad, row_stride = g.pw * kPad;
  const float two_i2s = 2.0f * sc.i2s;
  const float inv_w = 1.0f / (float)a.w;
  float* my_alpha_partial = (kAlpha && out.alpha_partials) ? out.alpha_partials + (size_t)blockIdx.x * a.M * hw : nullptr;
  if (my_alpha_partial)
    for (int e = threadIdx.x; e < a.M * hw; e += blockDim.x) my_alpha_partial[e] = 0.0f;
  // whole-kernel accumulators of the batch-reduced scalar gradients
  float acc_bgval = 0.0f, acc_bglogit = 0.0f, acc_tau = 0.0f, acc_sig = 0.0f, acc_g = 0.0f;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    load_pose_table(s, a, b);
    for (int m0 = 0; m0 < a.M; m0 += g.mc) {
      const int mc = min(g.mc, a.M - m0);
      __syncthreads();
      stage_atlas<C, kAlpha>(s.atlas, a, b, m0, mc, g.pw, g.ph, blockDim.x);
      for (int e = threadIdx.x; e < nwarps * g.mc * 8; e += blockDim.x) wpart[e] = 0.0f;
      __syncthreads();
      for (int ty = 0; ty < g.tiles_y; ++ty) {
        for (int tx = 0; tx < g.tiles_x; ++tx) {
          const int j = tx * g.tw + col;
          const bool col_ok = thread_ok && j < a.W;
          const int row0 = ty * g.k * g.ppt + rg;
          const float X = col_ok ? s.xs[j] : 0.0f;
          float Y[PIX], xv[PIX][C], G[PIX][C], Nc[PIX][C], Dc[PIX][C];
#pragma unroll
          for (int u = 0; u < PIX; ++u) {
            const int i = row0 + u * g.k;
            const bool ok = col_ok && u < g.ppt && i < a.H;
            Y[u] = ok ? s.ys[i] : 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const size_t px = ((size_t)b * C + c) * HW + (size_t)i * a.W + j;
              xv[u][c] = ok ? __ldg(x + px) : 0.0f;
              G[u][c] = ok ? __ldg(gout + px) : 0.0f;     // G = 0 switches every contribution of a dead pixel off
              Nc[u][c] = ok ? __ldg(cache + ((size_t)b * 2 * C + c) * HW + (size_t)i * a.W + j) : 0.0f;
              Dc[u][c] = ok ? __ldg(cache + ((size_t)b * 2 * C + C + c) * HW + (size_t)i * a.W + j) : 0.0f;
            }
            if (m0 == 0 && ok) {
              // background component, once per pixel
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const size_t px = ((size_t)b * C + c) * HW + (size_t)i * a.W + j;
                const float bg = a.bg_image ? __ldg(a.bg_image + px) : sc.bg_loc;
                const float bl = kAlpha ? sc.bg_logit : bg * sc.inv_tau;
                const float d = xv[u][c] - bg;
                const float pN = expf(fmaf(d * d, -sc.i2s, bl) - Nc[u][c]);
                const float pD = expf(bl - Dc[u][c]);
                const float glog = G[u][c] * (pN - pD);
                float gl = G[u][c] * pN * d * two_i2s;
                if (!kAlpha) {
                  gl = fmaf(glog, sc.inv_tau, gl);
                  acc_tau = fmaf(glog, bg, acc_tau);
                } else {
                  acc_bglogit += glog;
                }
                acc_sig = fmaf(G[u][c] * pN, d * d, acc_sig);
                acc_g += G[u][c];
                if (a.bg_image) {
                  if (out.g_bg_image) out.g_bg_image[px] = gl;
                } else {
                  acc_bgval += gl;
                }
              }
            }
          }
          for (int mm = 0; mm < mc; ++mm) {
            const float* t8 = s.tp + (size_t)(m0 + mm) * 8;
            const float4 pa = *reinterpret_cast<const float4*>(t8);
            const float4 pb = *reinterpret_cast<const float4*>(t8 + 4);
            const float cx = fmaf(X, pa.x, pa.z), cy = fmaf(X, pa.w, pb.y);
            const float lpres = pb.z;
            const float* tex = s.atlas + (size_t)mm * tex_stride;
            float* gtex = s.gatlas + (size_t)mm * tex_stride;
            float sgx = 0.f, sgxy = 0.f, sgy = 0.f, sgyy = 0.f, spres = 0.f;
#pragma unroll
            for (int u = 0; u < PIX; ++u) {
              if (u < g.ppt) {
                Tap t;
                tap_setup(fmaf(Y[u], pa.y, cx), fmaf(Y[u], pb.x, cy), wf, hf, g.pw, t);
                const int o00 = t.off * kPad;
                const float* q = tex + o00;
                const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
                const Texel<kPad> t01 = ld_texel<kPad>(q + row_stride), t11 = ld_texel<kPad>(q + row_stride + kPad);
                const float gx1 = 1.0f - t.fx, gy1 = 1.0f - t.fy;
                float al = 0.0f, pD_shared = 0.0f;
                if (kAlpha) {
                  al = bilerp<kPad>(t00, t10, t01, t11, t, C) + lpres;
                  pD_shared = expf(al - Dc[u][0]);
                }
                float glp = 0.0f, gix = 0.0f, giy = 0.0f;
                float* gq = gtex + o00;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  const float loc = bilerp<kPad>(t00, t10, t01, t11, t, c);
                  const float d = xv[u][c] - loc;
                  const float logit = kAlpha ? al : fmaf(loc, sc.inv_tau, lpres);
                  const float pN = expf(fmaf(d * d, -sc.i2s, logit) - Nc[u][c]);
                  const float pD = kAlpha ? pD_shared : expf(logit - Dc[u][c]);
                  const float glog = G[u][c] * (pN - pD);
                  float gl = G[u][c] * pN * d * two_i2s;
                  if (!kAlpha) {
                    gl = fmaf(glog, sc.inv_tau, gl);
                    acc_tau = fmaf(glog, loc, acc_tau);
                  }
                  acc_sig = fmaf(G[u][c] * pN, d * d, acc_sig);
                  glp += glog;
                  if (gl != 0.0f) {
                    atomicAdd(gq + c, gl * t.w00);
                    atomicAdd(gq + kPad + c, gl * t.w10);
                    atomicAdd(gq + row_stride + c, gl * t.w01);
                    atomicAdd(gq + row_stride + kPad + c, gl * t.w11);
                  }
                  gix = fmaf(gl, fmaf(t11.v[c] - t01.v[c], t.fy, (t10.v[c] - t00.v[c]) * gy1), gix);
                  giy = fmaf(gl, fmaf(t11.v[c] - t10.v[c], t.fx, (t01.v[c] - t00.v[c]) * gx1), giy);
                }
                if (kAlpha) {
                  if (glp != 0.0f) {
                    atomicAdd(gq + C, glp * t.w00);
                    atomicAdd(gq + kPad + C, glp * t.w10);
                    atomicAdd(gq + row_stride + C, glp * t.w01);
                    atomicAdd(gq + row_stride + kPad + C, glp * t.w11);
                  }
                  gix = fmaf(glp, fmaf(t11.v[C] - t01.v[C], t.fy, (t10.v[C] - t00.v[C]) * gy1), gix);
                  giy = fmaf(glp, fmaf(t11.v[C] - t10.v[C], t.fx, (t01.v[C] - t00.v[C]) * gx1), giy);
                }
                sgx += gix;
                sgxy = fmaf(gix, Y[u], sgxy);
                sgy += giy;
                sgyy = fmaf(giy, Y[u], sgyy);
                spres += glp;
              }
            }
            // d ix / d gx = w/2, d iy / d gy = h/2;  gx = p0 X + p1 Y + p2,  gy = p3 X + p4 Y + p5
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
      // flush this chunk: template gradients, alpha partials, pose / presence gradients; then clear the gradient atlas
      {
        const int n = mc * C * hw;
        float* dst = out.g_templates + ((size_t)b * a.M + m0) * C * hw;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
          const int plane = e / hw, rem = e - plane * hw;
          const int y = (int)(((float)rem + 0.5f) * inv_w), xx = rem - y * a.w;
          const int m = plane / C, c = plane - m * C;
          dst[e] = s.gatlas[(((size_t)m * g.ph + (y + 2)) * g.pw + (xx + 2)) * kPad + c];
        }
        if (kAlpha && my_alpha_partial) {
          const int na = mc * hw;
          for (int e = threadIdx.x; e < na; e += blockDim.x) {
            const int m = e / hw, rem = e - m * hw;
            const int y = (int)(((float)rem + 0.5f) * inv_w), xx = rem - y * a.w;
            my_alpha_partial[(size_t)m0 * hw + e] += s.gatlas[(((size_t)m * g.ph + (y + 2)) * g.pw + (xx + 2)) * kPad + C];
          }
        }
        for (int e = threadIdx.x; e < mc * 7; e += blockDim.x) {
          const int mm = e / 7, q7 = e - mm * 7;
          float t = 0.0f;
          for (int wi = 0; wi < nwarps; ++wi) t += wpart[((size_t)wi * g.mc + mm) * 8 + q7];
          const int m = m0 + mm;
          if (q7 < 6) {
            out.g_pose[((size_t)b * a.M + m) * 6 + q7] = t * (q7 < 3 ? 0.5f * wf : 0.5f * hf);
          } else if (out.g_presence) {
            const float pr = s.tp[m * 8 + 7];
            out.g_presence[(size_t)b * a.M + m] = pr < kLogSafeEps ? 0.0f : t / pr;
          }
        }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < g.atlas_floats; e += blockDim.x) s.gatlas[e] = 0.0f;
    }
  }

  // batch-reduced scalar gradients -> one partial row per CTA, already mapped to the raw parameters
  const float t_bgval = block_sum(acc_bgval, s.red);
  const float t_bglogit = block_sum(acc_bglogit, s.red);
  const float t_tau = block_sum(acc_tau, s.red);
  const float t_sig = block_sum(acc_sig, s.red);
  const float t_g = block_sum(acc_g, s.red);
  if (threadIdx.x == 0) {
    float* sp = out.scalar_partials + (size_t)blockIdx.x * 4;
    sp[0] = a.bg_value && !a.bg_image ? t_bgval * sc.bg_loc * (1.0f - sc.bg_loc) : 0.0f;
    sp[1] = kAlpha ? t_bglogit * sigmoid_f(__ldg(a.bg_mixing_logit)) : 0.0f;
    sp[2] = kAlpha ? 0.0f : -t_tau * sc.inv_tau * sc.inv_tau * sigmoid_f(__ldg(a.temperature_logit) + 0.5f);
    const float inv_sigma = __frcp_rn(sc.sigma);
    sp[3] = a.scale ? (t_sig * inv_sigma * inv_sigma * inv_sigma - t_g * inv_sigma) * sigmoid_f(__ldg(a.scale)) : 0.0f;
  }
}

// ================================================================================================================
// dispatch
// ================================================================================================================
template <typename F>
static int set_smem(F kern, size_t bytes) {
  if (bytes > 48 * 1024)
    SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SCAE_OK;
}

#define SCAE_TMPL_DISPATCH(C_, ALPHA_, ...)                   \
  do {                                                        \
    if ((C_) == 1 && (ALPHA_)) { constexpr int kC = 1; constexpr bool kA = true; __VA_ARGS__; }        \
    else if ((C_) == 1) { constexpr int kC = 1; constexpr bool kA = false; __VA_ARGS__; }              \
    else if ((C_) == 2 && (ALPHA_)) { constexpr int kC = 2; constexpr bool kA = true; __VA_ARGS__; }   \
    else if ((C_) == 2) { constexpr int kC = 2; constexpr bool kA = false; __VA_ARGS__; }              \
    else if ((ALPHA_)) { constexpr int kC = 3; constexpr bool kA = true; __VA_ARGS__; }                \
    else { constexpr int kC = 3; constexpr bool kA = false; __VA_ARGS__; }                             \
  } while (0)

static size_t tmpl_ws_alpha_floats(const scae_tmpl_args* a, int grid) {
  return a->mode == SCAE_TMPL_MODE_ALPHA ? (size_t)grid * a->M * a->h * a->w : 0;
}

}  // namespace scae

using namespace scae;

extern "C" __attribute__((visibility("default"))) int scae_tmpl_ll_fwd(const scae_tmpl_args* a, const float* x, float* log_prob, float* ll, float* cache,
                                scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(x && log_prob, SCAE_EINVAL, "tmpl fwd: x and log_prob are required");
  TmplGeom g;
  rc = tmpl_geometry(a, 1, 0, 2, &g);
  if (rc != SCAE_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_ll_fwd_kernel<kC, kA>;
    rc = set_smem(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, log_prob, ll, cache, g);
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

extern "C" __attribute__((visibility("default"))) int scae_tmpl_render(const scae_tmpl_args* a, float* transformed_templates, float* mixing_logits, float* mode,
                                float* mean, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  TmplGeom g;
  rc = tmpl_geometry(a, 1, 0, 2, &g);
  if (rc != SCAE_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_render_kernel<kC, kA>;
    rc = set_smem(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, transformed_templates, mixing_logits, mode, mean, g);
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

static int tmpl_bwd_geometry(const scae_tmpl_args* a, TmplGeom* g) {
  // two atlases (values + gradients) and the per-warp pose partials [16 warps][mc][8]; mc is not known yet, so the
  // partials are budgeted for the worst case mc = M
  const size_t extra = (size_t)(kMaxThreads / 32) * a->M * 8 * sizeof(float);
  return tmpl_geometry(a, 2, extra, 2, g);
}

extern "C" __attribute__((visibility("default"))) size_t scae_tmpl_ll_bwd_workspace_bytes(const scae_tmpl_args* a) {
  if (tmpl_validate(a) != SCAE_OK) return 0;
  TmplGeom g;
  if (tmpl_bwd_geometry(a, &g) != SCAE_OK) return 0;
  return (tmpl_ws_alpha_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int scae_tmpl_ll_bwd(const scae_tmpl_args* a, const float* x, const float* grad_log_prob, const float* cache,
                                float* g_templates, float* g_pose, float* g_presence, float* g_bg_image, float* g_alpha,
                                float* g_scalars, void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(x && grad_log_prob && cache && g_templates && g_pose && g_scalars, SCAE_EINVAL,
               "tmpl bwd: a required pointer is NULL");
  TmplGeom g;
  rc = tmpl_bwd_geometry(a, &g);
  if (rc != SCAE_OK) return rc;
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
    auto kern = tmpl_ll_bwd_kernel<kC, kA>;
    rc = set_smem(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, grad_log_prob, cache, out, g);
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  if (alpha && g_alpha) {
    rc = launch_reduce_rows(alpha_partials, g_alpha, g.grid, a->M * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  return launch_reduce_rows(scalar_partials, g_scalars, g.grid, 4, stream);
}
