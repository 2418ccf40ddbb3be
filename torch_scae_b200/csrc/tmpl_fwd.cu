// Hot path 1 (sm_100a): template warp + per-pixel template-mixture Gaussian log-likelihood -- forward and render.
//
// Replaces TemplateBasedImageDecoder.forward (reference part_decoder.py:152-243) fused with GaussianMixture.log_prob
// (distributions.py:41-48).  Math: oracle/template_likelihood.py::closed_form_log_prob.
//
// One persistent CTA per image slot: the M templates of the image (and the shared alpha logits) are staged into a
// zero-bordered shared-memory atlas; every thread owns a few pixels of one image column and walks the templates,
// folding each component into streaming-logsumexp state held in registers.  The B x (M+1) x C x H x W warped-template
// and mixing-logit tensors of the reference are never formed; HBM sees templates, poses, the target image and one
// log-prob per pixel.  Roofline discussion (this path is issue/LDS bound, not HBM bound): DESIGN.md section 5.
#include "tmpl_common.cuh"

namespace scae {

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
  SCAE_REQUIRE((long)a->h * a->w <= 16384 && a->H <= 16384 && a->W <= 16384 && a->M <= 4096, SCAE_ELIMIT,
               "tmpl: sizes too large (template <= 128x128, image side <= 16384, M <= 4096)");
  return SCAE_OK;
}

int tmpl_texel_floats(const scae_tmpl_args* a) {
  const int ch = a->C + (a->mode == SCAE_TMPL_MODE_ALPHA ? 1 : 0);
  return ch <= 1 ? 1 : (ch <= 2 ? 2 : 4);
}

int tmpl_geometry(const scae_tmpl_args* a, size_t per_tmpl_extra_bytes, size_t fixed_extra_bytes, size_t smem_budget,
                  TmplGeom* g) {
  const int pixmax = a->C == 1 ? 5 : 3;
  const int tw = a->W < 64 ? a->W : 64;
  int best_k = 1;
  double best_score = -1e9;
  for (int k = 1; k <= a->H && tw * k <= kTmplThreads; ++k) {
    const int threads = ((tw * k + 31) / 32) * 32;
    int ppt = (a->H + k - 1) / k, tiles_y = 1;
    if (ppt > pixmax) {
      ppt = pixmax;
      tiles_y = (a->H + k * pixmax - 1) / (k * pixmax);
    }
    const double eff = (double)a->H * tw / ((double)threads * ppt * tiles_y);
    double score = eff - 0.02 * (tiles_y - 1);
    if (threads < 128) score -= 0.10;
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
  g->pix_floats = 0;
  g->groups = 1;

  const size_t atlas_tmpl = (size_t)g->pw * g->ph * tmpl_texel_floats(a) * sizeof(float);
  const size_t per_tmpl = atlas_tmpl + per_tmpl_extra_bytes;
  const size_t fixed = fixed_extra_bytes + ((size_t)a->M * 8 + a->W + a->H + 64 + 8) * sizeof(float);
  const size_t limit = (size_t)max_smem_optin();
  if (smem_budget > limit) smem_budget = limit;
  SCAE_REQUIRE(fixed + per_tmpl <= limit, SCAE_ELIMIT,
               "tmpl: one %dx%d template (+%zu bytes of scratch) does not fit in %zu bytes of shared memory", a->h, a->w,
               fixed + per_tmpl_extra_bytes, limit);
  long mc = smem_budget > fixed ? (long)((smem_budget - fixed) / per_tmpl) : 0;
  if (mc < 1) mc = 1;                       // one template per chunk still fits (checked above)
  if (mc > a->M) mc = a->M;
  const int nchunks = (a->M + (int)mc - 1) / (int)mc;      // balance the chunks
  mc = (a->M + nchunks - 1) / nchunks;
  g->mc = (int)mc;
  g->atlas_floats = (int)((((size_t)g->pw * g->ph * tmpl_texel_floats(a) * g->mc) + 3) / 4 * 4);
  g->smem_bytes = fixed + (size_t)g->atlas_floats * sizeof(float) + per_tmpl_extra_bytes * g->mc + 16;
  int per_sm = (int)(limit / g->smem_bytes);
  const int by_threads = 2048 / g->threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > 2) per_sm = 2;               // the kernels are compiled for 2 CTAs per SM (register budget)
  if (per_sm < 1) per_sm = 1;
  const long slots = (long)sm_count() * per_sm;
  g->grid = a->B < slots ? a->B : (int)slots;
  return SCAE_OK;
}

// ================================================================================================================
// forward
// ================================================================================================================
template <int C, bool kAlpha>
__global__ void __launch_bounds__(kTmplThreads, 2) tmpl_ll_fwd_kernel(const scae_tmpl_args a, const float* __restrict__ x,
                                                                      float* __restrict__ logp, float* __restrict__ ll,
                                                                      float* __restrict__ cache, const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, PIX = TT::kPixMax, ND = kAlpha ? 1 : C;
  extern __shared__ __align__(16) float smem[];
  const TmplSmem s = tmpl_carve(smem, a, g, 0);
  int* tab = reinterpret_cast<int*>(s.red + 64);   // [C*h*w] staging table
  tmpl_prologue(s, a, g, 0);
  build_stage_table(tab, C, a.h, a.w, g.pw, kPad);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W;
  const int col = threadIdx.x % g.tw, rg = threadIdx.x / g.tw;
  const bool thread_ok = (int)threadIdx.x < g.tw * g.k;
  const float lim_x = (float)a.w + 2.5f, lim_y = (float)a.h + 2.5f;
  const unsigned row = (unsigned)(g.pw * kPad), tex_stride = (unsigned)(g.ph * g.pw * kPad);
  const unsigned base0 = 0u - kMagicBits * (row + (unsigned)kPad);
  // all templates in one chunk: the alpha logits (shared by the whole batch) are staged once per CTA, not per image
  const bool one_chunk = g.mc >= a.M;
  __syncthreads();                                  // atlas zeroed, table built
  if (kAlpha && one_chunk) stage_alpha_tab<C>(s.atlas, tab, a, 0, a.M, (int)tex_stride);

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
        PixLse N[PIX][C], D[PIX][ND];
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
          stage_templates_tab<C>(s.atlas, tab, a, b, m0, mc, (int)tex_stride);
          if (kAlpha && !one_chunk) stage_alpha_tab<C>(s.atlas, tab, a, m0, mc, (int)tex_stride);
          __syncthreads();
          for (int mm = 0; mm < mc; ++mm) {
            const float* t8 = s.tp + (size_t)(m0 + mm) * 8;
            const float4 pa = *reinterpret_cast<const float4*>(t8);       // Ax Bx Cx Ay
            const float4 pb = *reinterpret_cast<const float4*>(t8 + 4);   // By Cy lpres pres
            const float cx = fmaf(X, pa.x, pa.z);
            const float cy = fmaf(X, pa.w, pb.y);
            const float lpres = pb.z;
            const unsigned base = base0 + (unsigned)mm * tex_stride;
#pragma unroll
            for (int u = 0; u < PIX; ++u) {
              if (u < g.ppt) {
                Tap t;
                tap_setup<kPad>(fmaf(Y[u], pa.y, cx), fmaf(Y[u], pb.x, cy), lim_x, lim_y, row, base, t);
                const float* q = s.atlas + t.off;
                const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
                const Texel<kPad> t01 = ld_texel<kPad>(q + row), t11 = ld_texel<kPad>(q + row + kPad);
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
__global__ void __launch_bounds__(kTmplThreads, 2) tmpl_render_kernel(const scae_tmpl_args a, float* __restrict__ tt,
                                                                      float* __restrict__ ml, float* __restrict__ mode,
                                                                      float* __restrict__ mean, float* __restrict__ comp,
                                                                      const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, ND = kAlpha ? 1 : C;
  extern __shared__ __align__(16) float smem[];
  const TmplSmem s = tmpl_carve(smem, a, g, 0);
  tmpl_prologue(s, a, g, 0);
  const TmplScalars sc = tmpl_scalars(a);
  const int HW = a.H * a.W, K = a.M + 1, CL = kAlpha ? 1 : C;
  const int col = threadIdx.x % g.tw, rg = threadIdx.x / g.tw;
  const bool thread_ok = (int)threadIdx.x < g.tw * g.k;
  const float lim_x = (float)a.w + 2.5f, lim_y = (float)a.h + 2.5f;
  const unsigned row = (unsigned)(g.pw * kPad), tex_stride = (unsigned)(g.ph * g.pw * kPad);
  const unsigned base0 = 0u - kMagicBits * (row + (unsigned)kPad);

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
          int best_m[ND];   // the component the mode takes its value from (M = background)
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            bg_best[d] = true;
            best_m[d] = a.M;
          }
          for (int m0 = 0; m0 < a.M; m0 += g.mc) {
            const int mc = min(g.mc, a.M - m0);
            __syncthreads();
            stage_atlas<C, kAlpha>(s.atlas, a, b, m0, mc, g.pw, g.ph);
            __syncthreads();
            for (int mm = 0; mm < mc; ++mm) {
              const int m = m0 + mm;
              const float* t8 = s.tp + (size_t)m * 8;
              Tap t;
              tap_setup<kPad>(fmaf(Y, t8[1], fmaf(X, t8[0], t8[2])), fmaf(Y, t8[4], fmaf(X, t8[3], t8[5])), lim_x, lim_y,
                              row, base0 + (unsigned)mm * tex_stride, t);
              const float* q = s.atlas + t.off;
              const Texel<kPad> t00 = ld_texel<kPad>(q), t10 = ld_texel<kPad>(q + kPad);
              const Texel<kPad> t01 = ld_texel<kPad>(q + row), t11 = ld_texel<kPad>(q + row + kPad);
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
                  best_m[d] = m;
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
            for (int d = 0; d < ND; ++d) {
              if (ml) ml[(((size_t)b * K + a.M) * CL + d) * HW + pix] = kAlpha ? sc.bg_logit : bgv[d] * sc.inv_tau;
              if (comp) comp[((size_t)b * CL + d) * HW + pix] = (float)best_m[d];
            }
          }
        }
  }
}

}  // namespace scae

using namespace scae;

static const size_t kFwdSmemBudget = 100 * 1024;   // two CTAs per SM

extern "C" __attribute__((visibility("default"))) int scae_tmpl_ll_fwd(const scae_tmpl_args* a, const float* x,
                                                                       float* log_prob, float* ll, float* cache,
                                                                       scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(x && log_prob, SCAE_EINVAL, "tmpl fwd: x and log_prob are required");
  TmplGeom g;
  rc = tmpl_geometry(a, 0, (size_t)a->C * a->h * a->w * sizeof(int), kFwdSmemBudget, &g);   // + the staging table
  if (rc != SCAE_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_ll_fwd_kernel<kC, kA>;
    rc = tmpl_prepare_kernel(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, log_prob, ll, cache, g);
    note_launch();
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

extern "C" __attribute__((visibility("default"))) int scae_tmpl_render(const scae_tmpl_args* a,
                                                                       float* transformed_templates,
                                                                       float* mixing_logits, float* mode, float* mean,
                                                                       float* mode_component, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  TmplGeom g;
  rc = tmpl_geometry(a, 0, 0, kFwdSmemBudget, &g);
  if (rc != SCAE_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_render_kernel<kC, kA>;
    rc = tmpl_prepare_kernel(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, transformed_templates, mixing_logits, mode, mean, mode_component, g);
    note_launch();
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}
