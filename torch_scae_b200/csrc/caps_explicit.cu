// Hot path 2 on EXPLICIT vote tensors (sm_100a): the reference's standalone CapsuleLikelihood(vote, scale, vote_presence,
// dummy_vote)(x, presence) (object_decoder.py:243-372) for callers that build their own votes instead of going through
// CapsuleLayer's parameter head.  Same mixture arithmetic as caps_ll.cu (oracle/capsule_likelihood.py), minus the vote
// composition: one thread per (image, part) walks the O object capsules with streaming-softmax state in registers; lanes
// of a warp are consecutive parts, so every (B,O,V)-shaped tensor is a unit-stride access.  The backward makes two
// passes over the objects (S[v] = sum_j posterior_j h_j, then the gradients); the dummy vote's gradient is summed over the
// batch in a fixed order.  Deterministic.
#include "caps_common.cuh"

namespace scae {

constexpr int kExplicitThreads = 128;

// log-density of x under Normal(vote, scale) summed over the 6 pose dimensions, and q = |x - vote|^2
__device__ __forceinline__ float explicit_log_density(const float x[6], const float vt[6], float sc, float& q) {
  q = 0.0f;
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    const float d = x[p] - vt[p];
    q = fmaf(d, d, q);
  }
  return -q * __frcp_rn(2.0f * sc * sc) - 6.0f * logf(sc) - 6.0f * kHalfLog2Pi;
}

__global__ void __launch_bounds__(kExplicitThreads) caps_explicit_fwd_kernel(const scae_caps_explicit_args a,
                                                                             const scae_caps_outputs o) {
  const int O = a.O, V = a.V;
  const long i = (long)blockIdx.x * kExplicitThreads + threadIdx.x;
  if (i >= (long)a.B * V) return;
  const int b = (int)(i / V), v = (int)(i - (long)b * V);
  const size_t bv = (size_t)i;
  float x[6];
#pragma unroll
  for (int p = 0; p < 6; ++p) x[p] = __ldg(a.x + bv * 6 + p);
  const float pres = a.presence ? __ldg(a.presence + bv) : 1.0f;
  Lse post, mix;
  post.init(kDummyLog + kDummyLog);     // the dummy component: logit log 0.01, log-density log 0.01 (:273-292)
  mix.init(kDummyLog);
  float sw[6], wv[6], swp = 0.0f, best = -INFINITY, wvp = 0.0f;
  int widx = 0;
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    sw[p] = __ldg(a.dummy_vote + (size_t)v * 6 + p);
    wv[p] = 0.0f;
  }
  for (int oo = 0; oo < O; ++oo) {
    const size_t bov = ((size_t)b * O + oo) * V + v;
    const size_t bov1 = ((size_t)b * (O + 1) + oo) * V + v;
    float vt[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) vt[p] = __ldg(a.vote + bov * 6 + p);
    const float sc = __ldg(a.scale + bov), vp = __ldg(a.vote_presence + bov);
    float q;
    const float lp = explicit_log_density(x, vt, sc, q);
    const float ml = log_safe_f(vp);
    const float pl = ml + lp;
    float resc;
    const float wn = post.push(pl, resc);
#pragma unroll
    for (int p = 0; p < 6; ++p) sw[p] = fmaf(sw[p], resc, wn * vt[p]);
    swp = fmaf(swp, resc, wn * vp);
    mix.push(ml);
    if (pl > best) {                      // lowest object index on ties (torch.argmax)
      best = pl;
      widx = oo;
      wvp = vp;
#pragma unroll
      for (int p = 0; p < 6; ++p) wv[p] = vt[p];
    }
    if (o.vote_presence_binary) o.vote_presence_binary[bov] = ml > kDummyLog ? 1.0f : 0.0f;
    if (o.posterior_mixing_prob) o.posterior_mixing_prob[bov] = pl;   // normalised below
    if (o.mixing_logit) o.mixing_logit[bov1] = ml;
    if (o.mixing_log_prob) o.mixing_log_prob[bov1] = ml;             // normalised below
  }
  const float lse = post.value();
  const float inv_s = __frcp_rn(post.s);
  if (o.log_prob_per_point) o.log_prob_per_point[bv] = lse;
  if (o.soft_winner) {
#pragma unroll
    for (int p = 0; p < 6; ++p) o.soft_winner[bv * 6 + p] = sw[p] * inv_s;
  }
  if (o.soft_winner_presence) o.soft_winner_presence[bv] = swp * inv_s;
  if (o.winner) {
#pragma unroll
    for (int p = 0; p < 6; ++p) o.winner[bv * 6 + p] = wv[p];
  }
  if (o.winner_presence) o.winner_presence[bv] = wvp;
  if (o.winner_idx) o.winner_idx[bv] = widx;
  if (o.is_from_capsule) o.is_from_capsule[bv] = widx / V;   // sic (object_decoder.py:334)
  if (o.posterior_mixing_prob) {
    for (int oo = 0; oo < O; ++oo) {
      const size_t bov = ((size_t)b * O + oo) * V + v;
      o.posterior_mixing_prob[bov] = expf(o.posterior_mixing_prob[bov] - lse);
    }
  }
  const size_t dummy_row = ((size_t)b * (O + 1) + O) * V + v;
  if (o.mixing_logit) o.mixing_logit[dummy_row] = kDummyLog;
  if (o.mixing_log_prob) {
    const float mlse = mix.value();
    for (int oo = 0; oo < O; ++oo) o.mixing_log_prob[((size_t)b * (O + 1) + oo) * V + v] -= mlse;
    o.mixing_log_prob[dummy_row] = kDummyLog - mlse;
  }
  // per-point term of ll_per_example; summed per example by the second kernel in a fixed order
  if (o.ll_per_example) a.point_ll[bv] = lse * pres;
}

__global__ void caps_explicit_sum_kernel(const float* __restrict__ point_ll, float* __restrict__ ll_per_example, int B,
                                         int V) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.0f;
  for (int v = 0; v < V; ++v) s += point_ll[(size_t)b * V + v];
  ll_per_example[b] = s;
}

struct ExplicitBwdOut {
  float *g_vote, *g_scale, *g_vote_presence, *g_x, *g_presence, *dummy_rows;
};

__global__ void __launch_bounds__(kExplicitThreads) caps_explicit_bwd_kernel(const scae_caps_explicit_args a,
                                                                             const scae_caps_saved sv,
                                                                             const scae_caps_upstream up,
                                                                             const ExplicitBwdOut out) {
  const int O = a.O, V = a.V;
  const long i = (long)blockIdx.x * kExplicitThreads + threadIdx.x;
  if (i >= (long)a.B * V) return;
  const int b = (int)(i / V), v = (int)(i - (long)b * V);
  const size_t bv = (size_t)i;
  float x[6], gsw[6], gw[6], dum[6];
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    x[p] = __ldg(a.x + bv * 6 + p);
    gsw[p] = up.g_soft_winner ? __ldg(up.g_soft_winner + bv * 6 + p) : 0.0f;
    gw[p] = up.g_winner ? __ldg(up.g_winner + bv * 6 + p) : 0.0f;
    dum[p] = __ldg(a.dummy_vote + (size_t)v * 6 + p);
  }
  const float pres = a.presence ? __ldg(a.presence + bv) : 1.0f;
  const float gswp = up.g_soft_winner_presence ? __ldg(up.g_soft_winner_presence + bv) : 0.0f;
  const float gwp = up.g_winner_presence ? __ldg(up.g_winner_presence + bv) : 0.0f;
  const int widx = (up.g_winner || up.g_winner_presence) ? (int)sv.winner_idx[bv] : -1;
  const float lse = __ldg(sv.log_prob_per_point + bv);
  const float gll = up.g_ll_per_example ? __ldg(up.g_ll_per_example + b) : 0.0f;
  const float g_lse = gll * pres;
  const float post_dummy = expf(2.0f * kDummyLog - lse);
  // ---- pass 1: S = sum_j posterior_j h_j over the O + 1 components, and the mixing-logit normaliser ----------------------
  float hd = 0.0f;
#pragma unroll
  for (int p = 0; p < 6; ++p) hd = fmaf(gsw[p], dum[p], hd);
  float S = post_dummy * hd;
  Lse mix;
  mix.init(kDummyLog);
  float G = 0.0f;   // sum_j of the upstream gradient of mixing_log_prob (its softmax backward)
  for (int oo = 0; oo < O; ++oo) {
    const size_t bov = ((size_t)b * O + oo) * V + v;
    const float vp = __ldg(a.vote_presence + bov);
    float h = up.g_posterior_mixing_prob ? __ldg(up.g_posterior_mixing_prob + bov) : 0.0f;
#pragma unroll
    for (int p = 0; p < 6; ++p) h = fmaf(gsw[p], __ldg(a.vote + bov * 6 + p), h);
    h = fmaf(gswp, vp, h);
    S = fmaf(__ldg(sv.posterior_mixing_prob + bov), h, S);
    if (up.g_mixing_log_prob) {
      mix.push(log_safe_f(vp));
      G += __ldg(up.g_mixing_log_prob + ((size_t)b * (O + 1) + oo) * V + v);
    }
  }
  float mlse = 0.0f;
  if (up.g_mixing_log_prob) {
    G += __ldg(up.g_mixing_log_prob + ((size_t)b * (O + 1) + O) * V + v);
    mlse = mix.value();
  }
  // ---- pass 2: gradients ------------------------------------------------------------------------------------------------
  float gx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int oo = 0; oo < O; ++oo) {
    const size_t bov = ((size_t)b * O + oo) * V + v;
    const size_t bov1 = ((size_t)b * (O + 1) + oo) * V + v;
    float vt[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) vt[p] = __ldg(a.vote + bov * 6 + p);
    const float sc = __ldg(a.scale + bov), vp = __ldg(a.vote_presence + bov);
    const float pst = __ldg(sv.posterior_mixing_prob + bov);
    float h = up.g_posterior_mixing_prob ? __ldg(up.g_posterior_mixing_prob + bov) : 0.0f;
#pragma unroll
    for (int p = 0; p < 6; ++p) h = fmaf(gsw[p], vt[p], h);
    h = fmaf(gswp, vp, h);
    const float g_pl = pst * (h - S) + g_lse * pst;
    // mixing logit: through the posterior logit, directly, and through the normalised mixing_log_prob
    float g_ml = g_pl;
    if (up.g_mixing_logit) g_ml += __ldg(up.g_mixing_logit + bov1);
    if (up.g_mixing_log_prob) g_ml += __ldg(up.g_mixing_log_prob + bov1) - expf(log_safe_f(vp) - mlse) * G;
    const bool is_win = oo == widx;
    float g_vp = gswp * pst + (is_win ? gwp : 0.0f);
    if (!(vp < kLogSafeEps)) g_vp = fmaf(g_ml, __frcp_rn(vp), g_vp);
    const float inv_sc = __frcp_rn(sc), inv2 = inv_sc * inv_sc;
    const float coef = g_pl * inv2;
    float q = 0.0f;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const float d = x[p] - vt[p];
      q = fmaf(d, d, q);
      gx[p] = fmaf(-coef, d, gx[p]);
      out.g_vote[bov * 6 + p] = fmaf(coef, d, fmaf(gsw[p], pst, is_win ? gw[p] : 0.0f));
    }
    out.g_scale[bov] = g_pl * inv_sc * fmaf(q, inv2, -6.0f);
    out.g_vote_presence[bov] = g_vp;
  }
  if (out.g_x) {
#pragma unroll
    for (int p = 0; p < 6; ++p) out.g_x[bv * 6 + p] = gx[p];
  }
  if (out.g_presence) out.g_presence[bv] = gll * lse;
  if (out.dummy_rows) {
#pragma unroll
    for (int p = 0; p < 6; ++p) out.dummy_rows[bv * 6 + p] = gsw[p] * post_dummy;
  }
}

static int explicit_validate(const scae_caps_explicit_args* a) {
  SCAE_REQUIRE(a != nullptr, SCAE_EINVAL, "caps explicit: args is NULL");
  SCAE_REQUIRE(a->B > 0 && a->O > 0 && a->V > 0, SCAE_EINVAL, "caps explicit: B, O, V must be positive (%d, %d, %d)", a->B,
               a->O, a->V);
  SCAE_REQUIRE(a->vote && a->scale && a->vote_presence && a->dummy_vote && a->x, SCAE_EINVAL,
               "caps explicit: a required pointer is NULL");
  return SCAE_OK;
}

}  // namespace scae

using namespace scae;

extern "C" __attribute__((visibility("default"))) int scae_caps_explicit_fwd(const scae_caps_explicit_args* a,
                                                                              const scae_caps_outputs* out,
                                                                              scae_stream_t stream_) {
  int rc = explicit_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(out != nullptr, SCAE_EINVAL, "caps explicit fwd: outputs is NULL");
  SCAE_REQUIRE(!out->ll_per_example || a->point_ll, SCAE_EINVAL, "caps explicit fwd: point_ll scratch is required with ll_per_example");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long n = (long)a->B * a->V;
  const int grid = (int)((n + kExplicitThreads - 1) / kExplicitThreads);
  caps_explicit_fwd_kernel<<<grid, kExplicitThreads, 0, stream>>>(*a, *out);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  if (out->ll_per_example) {
    caps_explicit_sum_kernel<<<(a->B + 127) / 128, 128, 0, stream>>>(a->point_ll, out->ll_per_example, a->B, a->V);
    note_launch();
    SCAE_CUDA_TRY(cudaGetLastError());
  }
  return SCAE_OK;
}

extern "C" __attribute__((visibility("default"))) size_t scae_caps_explicit_bwd_workspace_bytes(
    const scae_caps_explicit_args* a) {
  if (!a || a->B <= 0 || a->V <= 0) return 0;
  return (size_t)a->B * a->V * 6 * sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int scae_caps_explicit_bwd(
    const scae_caps_explicit_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up, float* g_vote,
    float* g_scale, float* g_vote_presence, float* g_dummy_vote, float* g_x, float* g_presence, void* workspace,
    size_t workspace_bytes, scae_stream_t stream_) {
  int rc = explicit_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(saved && up && g_vote && g_scale && g_vote_presence, SCAE_EINVAL,
               "caps explicit bwd: a required pointer is NULL");
  SCAE_REQUIRE(saved->posterior_mixing_prob && saved->log_prob_per_point, SCAE_EINVAL,
               "caps explicit bwd: the saved posterior and per-point log-probabilities are required");
  SCAE_REQUIRE(!(up->g_winner || up->g_winner_presence) || saved->winner_idx, SCAE_EINVAL,
               "caps explicit bwd: winner_idx is required with g_winner / g_winner_presence");
  SCAE_REQUIRE(!g_dummy_vote || (workspace && workspace_bytes >= scae_caps_explicit_bwd_workspace_bytes(a)), SCAE_EINVAL,
               "caps explicit bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long n = (long)a->B * a->V;
  const int grid = (int)((n + kExplicitThreads - 1) / kExplicitThreads);
  ExplicitBwdOut out{g_vote, g_scale, g_vote_presence, g_x, g_presence,
                     g_dummy_vote ? static_cast<float*>(workspace) : nullptr};
  caps_explicit_bwd_kernel<<<grid, kExplicitThreads, 0, stream>>>(*a, *saved, *up, out);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  if (g_dummy_vote) return launch_reduce_rows(out.dummy_rows, g_dummy_vote, a->B, a->V * 6, stream);
  return SCAE_OK;
}
