// Hot path 2 (sm_100a): object->part vote composition + part-pose mixture likelihood, forward and backward.
//
// Replaces the post-MLP half of CapsuleLayer.forward (reference object_decoder.py:160-236), the vote-side use of
// cv_ops.geometric_transform (cv_ops.py:20-76), CapsuleObjectDecoder.forward's glue (:413-415) and
// CapsuleLikelihood.__call__ (:257-372).  Math: oracle/capsule_likelihood.py (forward) and
// oracle/manual_backward.py::capsule_forward_backward (backward), which this file transcribes.
//
// Work decomposition: one thread per (image b, part v); the thread walks the O object capsules serially and keeps the
// mixture over objects as streaming softmax state in registers (running max / sum / soft-winner accumulators), so the
// logsumexp, argmax, posterior and soft-winner reductions over O need no cross-thread traffic.  Lanes of a warp are
// consecutive parts, which makes every (B,O,V)-shaped input and output a unit-stride access.  The per-(b,o)
// capsule-level transform (cvr -> 2x3 matrix, capsule presence) is computed once per CTA into shared memory.
// A CTA covers floor(128 / V) whole images.
//
// Per-image algorithmic HBM traffic and the roofline are in DESIGN.md section 5.
#include "caps_common.cuh"

namespace scae {

constexpr int kCapsThreads = 128;
constexpr int kCapsIlp = 4;      // objects processed per loop iteration as independent dependency chains

__host__ __device__ inline int caps_imgs_per_cta(int V) { return V >= kCapsThreads ? 1 : kCapsThreads / V; }

struct CapsSmem {
  float* R;       // [imgs][O][8]: r00 r01 r02 r10 r11 r12, capsule presence pc, capsule presence logit
  float* tile;    // fwd: vote_presence [imgs][O][Vp]; bwd: accumulated (g_r[6], g_pc) [imgs][O][8]
  float* red;     // fwd: [2][imgs*V] per-point partials; bwd: scratch [7][kCapsThreads+1]
};

// One staging buffer holds, for the current group of kCapsIlp objects, every per-pair input of the CTA's images:
//   prm   [imgs][kCapsIlp][A]   all_param rows              bv / bs [kCapsIlp][V]   bias_vote / bias_scale rows
//   stc   [kCapsIlp][V*6]       cpr_static rows             nz      [imgs][kCapsIlp][V]  noise_vote (if given)
//   post / gpost [imgs][kCapsIlp][V]  saved posterior and its upstream gradient (backward only)
// Two buffers (double buffering).
struct CapsStage {
  const float *prm, *stc, *bv, *bs, *nz, *post, *gpost;
};
__host__ __device__ inline size_t caps_stage_floats(int imgs, int V, bool bwd) {
  return (size_t)imgs * kCapsIlp * (8 * V + 7) + (size_t)kCapsIlp * V * 6 + 2 * (size_t)kCapsIlp * V +
         (size_t)imgs * kCapsIlp * V * (bwd ? 3 : 1);
}
__device__ __forceinline__ CapsStage caps_stage_view(const float* stage, int imgs, int V) {
  const int A = 8 * V + 7;
  CapsStage s;
  s.prm = stage;
  s.stc = s.prm + (size_t)imgs * kCapsIlp * A;
  s.bv = s.stc + (size_t)kCapsIlp * V * 6;
  s.bs = s.bv + (size_t)kCapsIlp * V;
  s.nz = s.bs + (size_t)kCapsIlp * V;
  s.post = s.nz + (size_t)imgs * kCapsIlp * V;
  s.gpost = s.post + (size_t)imgs * kCapsIlp * V;
  return s;
}
__host__ __device__ inline size_t caps_fwd_smem_floats(int imgs, int O, int V) {
  return (size_t)imgs * O * 8 + (size_t)imgs * O * (V | 1) + 2 * (size_t)imgs * V + 2 * caps_stage_floats(imgs, V, false);
}
__host__ __device__ inline size_t caps_bwd_smem_floats(int imgs, int O, int V) {
  return (size_t)imgs * O * 8 * 2 + (size_t)kCapsIlp * 7 * (kCapsThreads + 1) + 2 * caps_stage_floats(imgs, V, true);
}

// Issues the asynchronous, fully coalesced copy of one object group into a staging buffer.  The rows of objects
// o0 .. o0+nobj-1 of one image are contiguous in every source tensor, so each (tensor, image) is one run (4-byte
// cp.async: all_param rows are only 4-byte aligned because A = 8V+7 is odd).  Replaces per-thread global loads whose
// 24-byte lane stride touched 24 sectors per warp-load and whose latency sat on the critical path of every pair
// (profiles/r01e: long_scoreboard 6.4 stall cycles per issue).
__device__ __forceinline__ void caps_copy_run(float* dst, const float* src, int n) {
  for (int e = threadIdx.x; e < n; e += kCapsThreads) cp_async4(dst + e, src + e);
}
__device__ __forceinline__ void caps_stage_issue(const scae_caps_args& a, const float* post, const float* gpost, int b0,
                                                 int nimg, int imgs, int o0, float* stage) {
  const int O = a.O, V = a.V, A = 8 * V + 7;
  const int nobj = min(kCapsIlp, O - o0);
  const CapsStage s = caps_stage_view(stage, imgs, V);
  for (int bi = 0; bi < nimg; ++bi) {
    const size_t bo = (size_t)(b0 + bi) * O + o0;
    caps_copy_run(const_cast<float*>(s.prm) + (size_t)bi * kCapsIlp * A, a.all_param + bo * A, nobj * A);
    if (a.noise_vote) caps_copy_run(const_cast<float*>(s.nz) + (size_t)bi * kCapsIlp * V, a.noise_vote + bo * V, nobj * V);
    if (post) caps_copy_run(const_cast<float*>(s.post) + (size_t)bi * kCapsIlp * V, post + bo * V, nobj * V);
    if (gpost) caps_copy_run(const_cast<float*>(s.gpost) + (size_t)bi * kCapsIlp * V, gpost + bo * V, nobj * V);
  }
  caps_copy_run(const_cast<float*>(s.stc), a.cpr_static + (size_t)o0 * V * 6, nobj * V * 6);
  caps_copy_run(const_cast<float*>(s.bv), a.bias_vote + (size_t)o0 * V, nobj * V);
  caps_copy_run(const_cast<float*>(s.bs), a.bias_scale + (size_t)o0 * V, nobj * V);
  cp_async_commit();
}

// ---- phase 0 (both directions): capsule-level quantities per (image, object) -------------------------------------
template <bool kSim>
__device__ __forceinline__ void caps_phase0(const scae_caps_args& a, int b0, int nimg, float* R,
                                            float* logit_out /* nullable, [B,O] */) {
  const int O = a.O, V = a.V, A = 8 * V + 7;
  for (int i = threadIdx.x; i < nimg * O; i += kCapsThreads) {
    const int bi = i / O, oo = i - bi * O, b = b0 + bi;
    const float* row = a.all_param + ((size_t)b * O + oo) * A;
    float t[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) t[p] = __ldg(row + 6 * V + p) + __ldg(a.bias_cvr + oo * 6 + p);
    PoseAffine r;
    pose_affine_fwd<kSim>(t, r);
    float lc = __ldg(row + 6 * V + 6) + __ldg(a.bias_caps + oo);
    if (a.noise_caps) lc += __ldg(a.noise_caps + (size_t)b * O + oo);
    float* dst = R + (size_t)i * 8;
#pragma unroll
    for (int p = 0; p < 6; ++p) dst[p] = r.a[p];
    dst[6] = sigmoid_f(lc);
    dst[7] = lc;
    if (logit_out) logit_out[(size_t)b * O + oo] = lc;
  }
}

// Per-pair inputs, each pointing at the object's row (shared-memory staging buffer, or global memory):
//   row [A] all_param, srow [V*6] cpr_static, bvote / bscale [V] biases, nvote [V] noise (nullable)
template <bool kSim>
__device__ __forceinline__ void caps_pair_fwd(const float* row, const float* srow, const float* bvote, const float* bscale,
                                              const float* nvote, const float* r, int V, int v, bool deform, bool learn,
                                              CapsPair& c) {
  float t[6];
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    c.dyn[p] = deform ? row[6 * v + p] : 0.0f;
    t[p] = c.dyn[p] + srow[6 * v + p];
  }
  pose_affine_fwd<kSim>(t, c.pa);
  const float* A_ = c.pa.a;
  c.vt[0] = r[0] * A_[0] + r[1] * A_[3];
  c.vt[1] = r[0] * A_[1] + r[1] * A_[4];
  c.vt[2] = r[0] * A_[2] + r[1] * A_[5] + r[2];
  c.vt[3] = r[3] * A_[0] + r[4] * A_[3];
  c.vt[4] = r[3] * A_[1] + r[4] * A_[4];
  c.vt[5] = r[3] * A_[2] + r[4] * A_[5] + r[5];
  c.lv = row[6 * V + 7 + v] + bvote[v];
  if (nvote) c.lv += nvote[v];
  c.pv = sigmoid_f(c.lv);
  c.vp = r[6] * c.pv;
  c.u = row[7 * V + 7 + v] + bscale[v];
  c.sc = learn ? softplus_f(c.u + 0.5f) + 1e-2f : 1.0f;
}

// ================================================================================================================
// forward
// ================================================================================================================
template <bool kSim>
__global__ void __launch_bounds__(kCapsThreads) caps_ll_fwd_kernel(const scae_caps_args a, const scae_caps_outputs o,
                                                                   const int imgs_per_cta) {
  extern __shared__ float smem[];
  const int O = a.O, V = a.V, A = 8 * V + 7, Vp = V | 1;
  const int b0 = blockIdx.x * imgs_per_cta;
  const int nimg = min(imgs_per_cta, a.B - b0);
  float* R = smem;
  float* vp_tile = R + (size_t)imgs_per_cta * O * 8;
  float* red = vp_tile + (size_t)imgs_per_cta * O * Vp;
  float* stage0 = red + 2 * (size_t)imgs_per_cta * V;                     // two staging buffers (see caps_stage_issue)
  const size_t stage_floats = caps_stage_floats(imgs_per_cta, V, false);  // must match caps_fwd_smem_floats
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;

  caps_phase0<kSim>(a, b0, nimg, R, o.presence_logit_per_caps);
  __syncthreads();

  const int n_items = nimg * V;
  for (int base = 0; base < n_items; base += kCapsThreads) {
    const int i = base + threadIdx.x;
    const bool active = i < n_items;
    const int bi = active ? i / V : 0, v = active ? i - bi * V : 0, b = b0 + bi;
    const size_t bv = (size_t)b * V + v;
    float x[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) x[p] = active ? __ldg(a.x + bv * 6 + p) : 0.0f;
    const float pres = (active && a.presence) ? __ldg(a.presence + bv) : 1.0f;
    // kCapsIlp independent streaming-softmax chains (objects oo, oo+1, ... of one iteration): at ~9 resident warps per
    // SM the kernel is latency bound, and the chains give the scheduler independent work; they are merged below.
    Lse post_c[kCapsIlp], mix_c[kCapsIlp];
    float sw_c[kCapsIlp][6], wv_c[kCapsIlp][6], swp_c[kCapsIlp], best_c[kCapsIlp], wvp_c[kCapsIlp];
    int widx_c[kCapsIlp];
    float regsum = 0.0f;
#pragma unroll
    for (int u = 0; u < kCapsIlp; ++u) {
      post_c[u].m = -INFINITY;
      post_c[u].s = 0.0f;
      mix_c[u].m = -INFINITY;
      mix_c[u].s = 0.0f;
      swp_c[u] = 0.0f;
      best_c[u] = -INFINITY;
      wvp_c[u] = 0.0f;
      widx_c[u] = 0;
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        sw_c[u][p] = 0.0f;
        wv_c[u][p] = 0.0f;
      }
    }
    // the dummy component enters chain 0 first, weight 1: dummy logit + dummy log-density (object_decoder.py:273-292)
    post_c[0].init(kDummyLog + kDummyLog);
    mix_c[0].init(kDummyLog);
#pragma unroll
    for (int p = 0; p < 6; ++p) sw_c[0][p] = __ldg(a.dummy_vote + (size_t)v * 6 + p);

    // object groups stream through two shared-memory staging buffers filled by coalesced cp.async copies
    caps_stage_issue(a, nullptr, nullptr, b0, nimg, imgs_per_cta, 0, stage0);
    for (int o0 = 0, g = 0; o0 < O; o0 += kCapsIlp, ++g) {
      float* stage = stage0 + (size_t)(g & 1) * stage_floats;
      if (o0 + kCapsIlp < O) {
        caps_stage_issue(a, nullptr, nullptr, b0, nimg, imgs_per_cta, o0 + kCapsIlp,
                         stage0 + (size_t)((g + 1) & 1) * stage_floats);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();                                   // this group's rows have landed for every thread
      const CapsStage st = caps_stage_view(stage, imgs_per_cta, V);
#pragma unroll
      for (int u = 0; u < kCapsIlp; ++u) {
        const int oo = o0 + u;
        if (active && oo < O) {
          const float* r = R + ((size_t)bi * O + oo) * 8;
          const size_t bov = ((size_t)b * O + oo) * V + v;
          const size_t bov1 = ((size_t)b * (O + 1) + oo) * V + v;
          CapsPair c;
          caps_pair_fwd<kSim>(st.prm + ((size_t)bi * kCapsIlp + u) * A, st.stc + (size_t)u * V * 6, st.bv + (size_t)u * V,
                              st.bs + (size_t)u * V, a.noise_vote ? st.nz + ((size_t)bi * kCapsIlp + u) * V : nullptr, r, V, v,
                              deform, learn, c);
          float q = 0.0f;
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const float d = x[p] - c.vt[p];
            q = fmaf(d, d, q);
            regsum = fmaf(c.dyn[p], c.dyn[p], regsum);
          }
          // sum over the 6 pose dims of Normal(vote, sc).log_prob(x)
          const float lp = -q * __frcp_rn(2.0f * c.sc * c.sc) - 6.0f * logf(c.sc) - 6.0f * kHalfLog2Pi;
          const float ml = log_safe_f(c.vp);
          const float pl = ml + lp;
          float resc;
          const float wn = post_c[u].push(pl, resc);
#pragma unroll
          for (int p = 0; p < 6; ++p) sw_c[u][p] = fmaf(sw_c[u][p], resc, wn * c.vt[p]);
          swp_c[u] = fmaf(swp_c[u], resc, wn * c.vp);
          mix_c[u].push(ml);
          if (pl > best_c[u]) {
            best_c[u] = pl;
            widx_c[u] = oo;
            wvp_c[u] = c.vp;
#pragma unroll
            for (int p = 0; p < 6; ++p) wv_c[u][p] = c.vt[p];
          }
          vp_tile[((size_t)bi * O + oo) * Vp + v] = c.vp;
          if (o.vote) {
#pragma unroll
            for (int p = 0; p < 6; ++p) o.vote[bov * 6 + p] = c.vt[p];
          }
          if (o.scale) o.scale[bov] = c.sc;
          if (o.vote_presence) o.vote_presence[bov] = c.vp;
          if (o.presence_logit_per_vote) o.presence_logit_per_vote[bov] = c.lv;
          if (o.vote_presence_binary) o.vote_presence_binary[bov] = ml > kDummyLog ? 1.0f : 0.0f;
          if (o.posterior_mixing_prob) o.posterior_mixing_prob[bov] = pl;   // normalised below
          if (o.mixing_logit) o.mixing_logit[bov1] = ml;
          if (o.mixing_log_prob) o.mixing_log_prob[bov1] = ml;             // normalised below
        }
      }
      __syncthreads();                                   // everyone is done with this buffer before it is refilled
    }
    // merge the chains into chain 0 (which is never empty: it holds the dummy component)
    Lse post = post_c[0], mix = mix_c[0];
    float sw[6], wv[6], swp = swp_c[0], best = best_c[0], wvp = wvp_c[0];
    int widx = widx_c[0];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      sw[p] = sw_c[0][p];
      wv[p] = wv_c[0][p];
    }
#pragma unroll
    for (int u = 1; u < kCapsIlp; ++u) {
      const float M = fmaxf(post.m, post_c[u].m);
      const float ea = expf(post.m - M), eb = expf(post_c[u].m - M);      // exp(-inf) = 0 for an empty chain
      post.s = post.s * ea + post_c[u].s * eb;
      post.m = M;
#pragma unroll
      for (int p = 0; p < 6; ++p) sw[p] = sw[p] * ea + sw_c[u][p] * eb;
      swp = swp * ea + swp_c[u] * eb;
      const float Mm = fmaxf(mix.m, mix_c[u].m);
      mix.s = mix.s * expf(mix.m - Mm) + mix_c[u].s * expf(mix_c[u].m - Mm);
      mix.m = Mm;
      // hard winner: largest posterior logit, lowest object index on ties (torch.argmax)
      if (best_c[u] > best || (best_c[u] == best && best_c[u] > -INFINITY && widx_c[u] < widx)) {
        best = best_c[u];
        widx = widx_c[u];
        wvp = wvp_c[u];
#pragma unroll
        for (int p = 0; p < 6; ++p) wv[p] = wv_c[u][p];
      }
    }

    if (active) {
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
        for (int oo = 0; oo < O; ++oo) {
          const size_t bov1 = ((size_t)b * (O + 1) + oo) * V + v;
          o.mixing_log_prob[bov1] -= mlse;
        }
        o.mixing_log_prob[dummy_row] = kDummyLog - mlse;
      }
      red[i] = lse * pres;
      red[n_items + i] = 0.5f * regsum;
    }
  }
  __syncthreads();

  // capsule presence = max over parts of vote_presence (object_decoder.py:415); lowest index wins ties
  for (int i = threadIdx.x; i < nimg * O; i += kCapsThreads) {
    const float* rowp = vp_tile + (size_t)i * Vp;
    float best = rowp[0];
    int arg = 0;
    for (int v = 1; v < V; ++v) {
      const float t = rowp[v];
      if (t > best) {
        best = t;
        arg = v;
      }
    }
    const size_t bo = (size_t)b0 * O + i;
    if (o.caps_presence) o.caps_presence[bo] = best;
    if (o.caps_presence_arg) o.caps_presence_arg[bo] = arg;
  }
  // per-example sums, fixed order
  for (int bi = threadIdx.x; bi < nimg; bi += kCapsThreads) {
    float ll = 0.0f, reg = 0.0f;
    for (int v = 0; v < V; ++v) {
      ll += red[bi * V + v];
      reg += red[n_items + bi * V + v];
    }
    if (o.ll_per_example) o.ll_per_example[b0 + bi] = ll;
    if (o.reg_per_example) o.reg_per_example[b0 + bi] = reg;
  }
}

// ================================================================================================================
// backward
// ================================================================================================================
struct CapsBwdOut {
  float* g_all_param;   // [B,O,A]; receives the gradient w.r.t. the pre-activation sums, finalised by the next kernel
  float* g_x;           // [B,V,6] nullable
  float* g_presence;    // [B,V]   nullable
  float* dummy_rows;    // [B,V*6] nullable: g_soft_winner * posterior(dummy), reduced over B afterwards
};

template <bool kSim>
__global__ void __launch_bounds__(kCapsThreads) caps_ll_bwd_kernel(const scae_caps_args a, const scae_caps_saved sv,
                                                                   const scae_caps_upstream up, const CapsBwdOut out,
                                                                   const int imgs_per_cta) {
  extern __shared__ float smem[];
  const int O = a.O, V = a.V, A = 8 * V + 7;
  const int b0 = blockIdx.x * imgs_per_cta;
  const int nimg = min(imgs_per_cta, a.B - b0);
  float* R = smem;
  float* acc = R + (size_t)imgs_per_cta * O * 8;     // [imgs][O][8]: sum over v of (g_r[6], g_vp*pv, unused)
  float* scr = acc + (size_t)imgs_per_cta * O * 8;   // [kCapsIlp][7][kCapsThreads+1]
  constexpr int kScr = kCapsThreads + 1;
  float* stage0 = scr + (size_t)kCapsIlp * 7 * kScr;                       // two staging buffers (see caps_stage_issue)
  const size_t stage_floats = caps_stage_floats(imgs_per_cta, V, true);
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;
  const bool soft = up.g_soft_winner != nullptr || up.g_soft_winner_presence != nullptr;
  const bool need_mix = up.g_mixing_log_prob != nullptr;

  caps_phase0<kSim>(a, b0, nimg, R, nullptr);
  for (int i = threadIdx.x; i < nimg * O * 8; i += kCapsThreads) acc[i] = 0.0f;
  __syncthreads();

  const int n_items = nimg * V;
  for (int base = 0; base < n_items; base += kCapsThreads) {
    const int i = base + threadIdx.x;
    const bool active = i < n_items;
    const int bi = active ? i / V : 0, v = active ? i - bi * V : 0, b = b0 + bi;
    const size_t bv = (size_t)b * V + v;
    const float* rowb = a.all_param + (size_t)b * O * A;
    float x[6], gsw[6], gw[6], gx[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      x[p] = active ? __ldg(a.x + bv * 6 + p) : 0.0f;
      gsw[p] = (active && up.g_soft_winner) ? __ldg(up.g_soft_winner + bv * 6 + p) : 0.0f;
      gw[p] = (active && up.g_winner) ? __ldg(up.g_winner + bv * 6 + p) : 0.0f;
      gx[p] = 0.0f;
    }
    const float pres = (active && a.presence) ? __ldg(a.presence + bv) : 1.0f;
    const float gll = (active && up.g_ll_per_example) ? __ldg(up.g_ll_per_example + b) : 0.0f;
    const float greg = (active && up.g_reg_per_example) ? __ldg(up.g_reg_per_example + b) : 0.0f;
    const float gswp = (active && up.g_soft_winner_presence) ? __ldg(up.g_soft_winner_presence + bv) : 0.0f;
    const float gwp = (active && up.g_winner_presence) ? __ldg(up.g_winner_presence + bv) : 0.0f;
    const int widx = (active && sv.winner_idx) ? (int)sv.winner_idx[bv] : -1;
    const float lse = active ? __ldg(sv.log_prob_per_point + bv) : 0.0f;
    const float post_dummy = expf(kDummyLog + kDummyLog - lse);

    // ---- pass A: S = sum_k posterior_k * h_k (k over objects and the dummy), and the mixing-logit softmax terms ----
    float S = 0.0f, sum_gmlp = 0.0f, mix_lse = 0.0f;
    if (active) {
      if (soft || need_mix) {
        Lse mix;
        mix.init(kDummyLog);
        for (int oo = 0; oo < O; ++oo) {
          const size_t bov = ((size_t)b * O + oo) * V + v;
          CapsPair c;
          caps_pair_fwd<kSim>(rowb + (size_t)oo * A, a.cpr_static + (size_t)oo * V * 6, a.bias_vote + (size_t)oo * V,
                              a.bias_scale + (size_t)oo * V,
                              a.noise_vote ? a.noise_vote + ((size_t)b * O + oo) * V : nullptr,
                              R + ((size_t)bi * O + oo) * 8, V, v, deform, learn, c);
          float h = up.g_posterior_mixing_prob ? __ldg(up.g_posterior_mixing_prob + bov) : 0.0f;
#pragma unroll
          for (int p = 0; p < 6; ++p) h = fmaf(gsw[p], c.vt[p], h);
          h = fmaf(gswp, c.vp, h);
          S = fmaf(__ldg(sv.posterior_mixing_prob + bov), h, S);
          if (need_mix) {
            mix.push(log_safe_f(c.vp));
            sum_gmlp += __ldg(up.g_mixing_log_prob + ((size_t)b * (O + 1) + oo) * V + v);
          }
        }
        float hd = 0.0f;
#pragma unroll
        for (int p = 0; p < 6; ++p) hd = fmaf(gsw[p], __ldg(a.dummy_vote + (size_t)v * 6 + p), hd);
        S = fmaf(post_dummy, hd, S);
        if (need_mix) {
          sum_gmlp += __ldg(up.g_mixing_log_prob + ((size_t)b * (O + 1) + O) * V + v);
          mix_lse = mix.value();
        }
      } else if (up.g_posterior_mixing_prob) {
        for (int oo = 0; oo < O; ++oo) {
          const size_t bov = ((size_t)b * O + oo) * V + v;
          S = fmaf(__ldg(sv.posterior_mixing_prob + bov), __ldg(up.g_posterior_mixing_prob + bov), S);
        }
      }
    }

    // ---- pass B: per-pair gradients -------------------------------------------------------------------------------
    caps_stage_issue(a, sv.posterior_mixing_prob, up.g_posterior_mixing_prob, b0, nimg, imgs_per_cta, 0, stage0);
    for (int o0 = 0, g = 0; o0 < O; o0 += kCapsIlp, ++g) {
      // kCapsIlp objects per iteration as independent dependency chains (latency hiding at low occupancy); their rows
      // arrive through the double-buffered staging area and their per-(image, object) partial sums go through the
      // shared-memory transpose together, one barrier pair per group
      float* stage = stage0 + (size_t)(g & 1) * stage_floats;
      if (o0 + kCapsIlp < O) {
        caps_stage_issue(a, sv.posterior_mixing_prob, up.g_posterior_mixing_prob, b0, nimg, imgs_per_cta, o0 + kCapsIlp,
                         stage0 + (size_t)((g + 1) & 1) * stage_floats);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const CapsStage st = caps_stage_view(stage, imgs_per_cta, V);
      float red7[kCapsIlp][7];
#pragma unroll
      for (int u = 0; u < kCapsIlp; ++u)
#pragma unroll
        for (int k = 0; k < 7; ++k) red7[u][k] = 0.0f;
      if (active) {
#pragma unroll
        for (int u = 0; u < kCapsIlp; ++u) {
          const int oo = o0 + u;
          if (oo < O) {
            const float* r = R + ((size_t)bi * O + oo) * 8;
            const size_t bov = ((size_t)b * O + oo) * V + v;
            const size_t bov1 = ((size_t)b * (O + 1) + oo) * V + v;
            const size_t bo = (size_t)b * O + oo;
            CapsPair c;
            const size_t su = ((size_t)bi * kCapsIlp + u) * V;
            caps_pair_fwd<kSim>(st.prm + ((size_t)bi * kCapsIlp + u) * A, st.stc + (size_t)u * V * 6, st.bv + (size_t)u * V,
                                st.bs + (size_t)u * V, a.noise_vote ? st.nz + su : nullptr, r, V, v, deform, learn, c);
            const float post = st.post[su + v];
            float h = up.g_posterior_mixing_prob ? st.gpost[su + v] : 0.0f;
            float diff[6], q = 0.0f;
    #pragma unroll
            for (int p = 0; p < 6; ++p) {
              h = fmaf(gsw[p], c.vt[p], h);
              diff[p] = x[p] - c.vt[p];
              q = fmaf(diff[p], diff[p], q);
            }
            h = fmaf(gswp, c.vp, h);
            const float g_pl = post * (h - S) + gll * pres * post;
            const bool is_win = oo == widx;
            float g_vp = gswp * post;
            if (up.g_vote_presence) g_vp += __ldg(up.g_vote_presence + bov);
            if (is_win) g_vp += gwp;
            if (up.g_caps_presence && sv.caps_presence_arg[bo] == v) g_vp += __ldg(up.g_caps_presence + bo);
            float g_ml = g_pl;
            if (up.g_mixing_logit) g_ml += __ldg(up.g_mixing_logit + bov1);
            if (need_mix) g_ml += __ldg(up.g_mixing_log_prob + bov1) - expf(log_safe_f(c.vp) - mix_lse) * sum_gmlp;
            if (!(c.vp < kLogSafeEps)) g_vp += g_ml / c.vp;
            const float inv_sc = __frcp_rn(c.sc);
            const float inv2 = inv_sc * inv_sc;
            const float coef = g_pl * inv2;
            float gv[6];
    #pragma unroll
            for (int p = 0; p < 6; ++p) {
              float g = gsw[p] * post + coef * diff[p];
              if (up.g_vote) g += __ldg(up.g_vote + bov * 6 + p);
              if (is_win) g += gw[p];
              gv[p] = g;
              gx[p] = fmaf(-coef, diff[p], gx[p]);
            }
            float g_sc = g_pl * (q * inv2 * inv_sc - 6.0f * inv_sc);
            if (up.g_scale) g_sc += __ldg(up.g_scale + bov);
            const float g_u = learn ? g_sc * sigmoid_f(c.u + 0.5f) : 0.0f;
            float g_lv = g_vp * r[6] * c.pv * (1.0f - c.pv);
            if (up.g_presence_logit_per_vote) g_lv += __ldg(up.g_presence_logit_per_vote + bov);
            // vote = R . A
            const float* A_ = c.pa.a;
            float ga[6];
            ga[0] = r[0] * gv[0] + r[3] * gv[3];
            ga[1] = r[0] * gv[1] + r[3] * gv[4];
            ga[2] = r[0] * gv[2] + r[3] * gv[5];
            ga[3] = r[1] * gv[0] + r[4] * gv[3];
            ga[4] = r[1] * gv[1] + r[4] * gv[4];
            ga[5] = r[1] * gv[2] + r[4] * gv[5];
            red7[u][0] = gv[0] * A_[0] + gv[1] * A_[1] + gv[2] * A_[2];
            red7[u][1] = gv[0] * A_[3] + gv[1] * A_[4] + gv[2] * A_[5];
            red7[u][2] = gv[2];
            red7[u][3] = gv[3] * A_[0] + gv[4] * A_[1] + gv[5] * A_[2];
            red7[u][4] = gv[3] * A_[3] + gv[4] * A_[4] + gv[5] * A_[5];
            red7[u][5] = gv[5];
            red7[u][6] = g_vp * c.pv;
            float gt[6];
            pose_affine_bwd<kSim>(ga, c.pa, gt);
            float* grow = out.g_all_param + ((size_t)b * O + oo) * A;
    #pragma unroll
            for (int p = 0; p < 6; ++p) grow[6 * v + p] = gt[p];
            grow[6 * V + 7 + v] = g_lv;
            grow[7 * V + 7 + v] = g_u;
          }
        }
      }
      // sum over the parts of each image: transpose through shared memory, one reducer thread per (image, object, slot)
#pragma unroll
      for (int u = 0; u < kCapsIlp; ++u)
#pragma unroll
        for (int k = 0; k < 7; ++k) scr[(u * 7 + k) * kScr + threadIdx.x] = red7[u][k];
      __syncthreads();
      {
        const int img_lo = base / V;
        const int img_hi = min(nimg - 1, (base + kCapsThreads - 1) / V);
        const int n_red = (img_hi - img_lo + 1) * 7 * kCapsIlp;
        for (int t = threadIdx.x; t < n_red; t += kCapsThreads) {
          const int rb = img_lo + t / (7 * kCapsIlp), rem = t % (7 * kCapsIlp);
          const int u = rem / 7, k = rem - u * 7;
          if (o0 + u < O) {
            const int j0 = max(0, rb * V - base), j1 = min(kCapsThreads, (rb + 1) * V - base);
            float sum = 0.0f;
            for (int j = j0; j < j1; ++j) sum += scr[(u * 7 + k) * kScr + j];
            acc[((size_t)rb * O + o0 + u) * 8 + k] += sum;
          }
        }
      }
      __syncthreads();
    }

    if (active) {
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
  }
  __syncthreads();

  // ---- capsule-level slots: cvr (6) and the capsule presence logit ---------------------------------------------------
  for (int i = threadIdx.x; i < nimg * O; i += kCapsThreads) {
    const int bi = i / O, oo = i - bi * O, b = b0 + bi;
    const float* row = a.all_param + ((size_t)b * O + oo) * A;
    float t[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) t[p] = __ldg(row + 6 * V + p) + __ldg(a.bias_cvr + oo * 6 + p);
    PoseAffine r;
    pose_affine_fwd<kSim>(t, r);
    const float* g = acc + (size_t)i * 8;
    float gr[6] = {g[0], g[1], g[2], g[3], g[4], g[5]}, gt[6];
    pose_affine_bwd<kSim>(gr, r, gt);
    float* grow = out.g_all_param + ((size_t)b * O + oo) * A;
#pragma unroll
    for (int p = 0; p < 6; ++p) grow[6 * V + p] = gt[p];
    const float pc = R[(size_t)i * 8 + 6];
    float g_lc = g[6] * pc * (1.0f - pc);
    if (up.g_presence_logit_per_caps) g_lc += __ldg(up.g_presence_logit_per_caps + (size_t)b * O + oo);
    grow[6 * V + 6] = g_lc;
  }
}

// Column sums over the batch of the pre-activation gradient (-> g_shared partials), then the in-place finalisation of
// g_all_param: + g_reg * cpr_dynamic on the deformation slots (or zero when deformations are off) and the optional
// ReLU mask of the producing MLP.  One thread per column of the [B, O*A] matrix, unit stride across threads.
__global__ void __launch_bounds__(256) caps_bwd_finalize_kernel(float* __restrict__ g_all_param,
                                                                const float* __restrict__ all_param,
                                                                const float* __restrict__ g_reg_per_example,
                                                                float* __restrict__ partials, int B, int O, int V,
                                                                unsigned flags, int rows_per_split) {
  const int A = 8 * V + 7;
  const int n = O * A;
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= n) return;
  const bool is_dyn = (col % A) < 6 * V;
  const bool deform = (flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool relu = (flags & SCAE_CAPS_RELU_GRAD) != 0;
  const bool rewrite = relu || (is_dyn && (!deform || g_reg_per_example != nullptr));
  const int r0 = blockIdx.y * rows_per_split, r1 = min(B, r0 + rows_per_split);
  float s = 0.0f;
  for (int b = r0; b < r1; ++b) {
    const size_t idx = (size_t)b * n + col;
    const float g = g_all_param[idx];
    s += g;
    if (rewrite) {
      const float ap = __ldg(all_param + idx);
      float o = g;
      if (is_dyn) o = deform ? (g_reg_per_example ? fmaf(__ldg(g_reg_per_example + b), ap, g) : g) : 0.0f;
      if (relu && !(ap > 0.0f)) o = 0.0f;
      g_all_param[idx] = o;
    }
  }
  partials[(size_t)blockIdx.y * n + col] = s;
}

static int caps_validate(const scae_caps_args* a) {
  SCAE_REQUIRE(a != nullptr, SCAE_EINVAL, "caps: args is NULL");
  SCAE_REQUIRE(a->B > 0 && a->O > 0 && a->V > 0, SCAE_EINVAL, "caps: B, O, V must be positive (got %d, %d, %d)", a->B,
               a->O, a->V);
  SCAE_REQUIRE(a->all_param && a->cpr_static && a->bias_cvr && a->bias_caps && a->bias_vote && a->bias_scale && a->x &&
                   a->dummy_vote,
               SCAE_EINVAL, "caps: a required input pointer is NULL");
  SCAE_REQUIRE((a->flags & ~0xFu) == 0, SCAE_EINVAL, "caps: unknown flag bits 0x%x", a->flags);
  return SCAE_OK;
}

static int caps_split(int B) { return B < 32 ? B : 32; }

}  // namespace scae

using namespace scae;

extern "C" __attribute__((visibility("default"))) int scae_caps_ll_fwd(const scae_caps_args* a, const scae_caps_outputs* out, scae_stream_t stream_) {
  int rc = caps_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(out != nullptr, SCAE_EINVAL, "caps fwd: outputs is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!caps_force_v1()) {   // fast paths when the image's working set fits in shared memory: caps_ll3.cu, then caps_ll2.cu
    bool handled = false;
    if (!caps_force_v2()) {
      rc = caps3_fwd(a, out, stream, &handled);
      if (handled) note_fast_path(), note_persistent_path();
      if (rc != SCAE_OK || handled) return rc;
    }
    rc = caps2_fwd(a, out, stream, &handled);
    if (handled) note_fast_path();
    if (rc != SCAE_OK || handled) return rc;
  }
  const int imgs = caps_imgs_per_cta(a->V);
  const size_t smem = caps_fwd_smem_floats(imgs, a->O, a->V) * sizeof(float);
  SCAE_REQUIRE(smem <= (size_t)max_smem_optin(), SCAE_ELIMIT, "caps fwd: O=%d, V=%d need %zu bytes of shared memory",
               a->O, a->V, smem);
  const int grid = (a->B + imgs - 1) / imgs;
  const bool sim = (a->flags & SCAE_CAPS_SIMILARITY) != 0;
  auto kern = sim ? caps_ll_fwd_kernel<true> : caps_ll_fwd_kernel<false>;
  if (smem > 48 * 1024) SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kCapsThreads, smem, stream>>>(*a, *out, imgs);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

extern "C" __attribute__((visibility("default"))) size_t scae_caps_ll_bwd_workspace_bytes(const scae_caps_args* a) {
  if (a == nullptr || a->B <= 0 || a->O <= 0 || a->V <= 0) return 0;
  const size_t n = (size_t)a->O * (8 * a->V + 7);
  const size_t general = (caps_split(a->B) * n + (size_t)a->B * a->V * 6) * sizeof(float);
  const size_t fast = caps2_bwd_workspace_bytes(a) > caps3_bwd_workspace_bytes(a) ? caps2_bwd_workspace_bytes(a)
                                                                                 : caps3_bwd_workspace_bytes(a);
  return general > fast ? general : fast;
}

extern "C" __attribute__((visibility("default"))) int scae_caps_ll_bwd(const scae_caps_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up,
                                float* g_all_param, float* g_shared, float* g_dummy_vote, float* g_x,
                                float* g_presence, void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  int rc = caps_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(saved && up && g_all_param && g_shared, SCAE_EINVAL, "caps bwd: a required pointer is NULL");
  SCAE_REQUIRE(saved->posterior_mixing_prob && saved->log_prob_per_point, SCAE_EINVAL,
               "caps bwd: saved posterior_mixing_prob and log_prob_per_point are required");
  SCAE_REQUIRE(!up->g_caps_presence || saved->caps_presence_arg, SCAE_EINVAL,
               "caps bwd: g_caps_presence needs saved caps_presence_arg");
  SCAE_REQUIRE(!(up->g_winner || up->g_winner_presence) || saved->winner_idx, SCAE_EINVAL,
               "caps bwd: g_winner / g_winner_presence need saved winner_idx");
  SCAE_REQUIRE(workspace && workspace_bytes >= scae_caps_ll_bwd_workspace_bytes(a), SCAE_EINVAL,
               "caps bwd: workspace too small (%zu < %zu)", workspace_bytes, scae_caps_ll_bwd_workspace_bytes(a));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!caps_force_v1()) {   // fast path (caps_ll2.cu): training-step upstream set, working set fits in shared memory
    bool handled = false;
    if (!caps_force_v2()) {
      rc = caps3_bwd(a, saved, up, g_all_param, g_shared, g_dummy_vote, g_x, g_presence, workspace, workspace_bytes, stream,
                     &handled);
      if (handled) note_fast_path(), note_persistent_path();
      if (rc != SCAE_OK || handled) return rc;
    }
    rc = caps2_bwd(a, saved, up, g_all_param, g_shared, g_dummy_vote, g_x, g_presence, workspace, workspace_bytes, stream,
                   &handled);
    if (handled) note_fast_path();
    if (rc != SCAE_OK || handled) return rc;
  }
  const int B = a->B, O = a->O, V = a->V, A = 8 * V + 7, n = O * A;
  const int nsplit = caps_split(B);
  float* partials = static_cast<float*>(workspace);
  float* dummy_rows = partials + (size_t)nsplit * n;
  const bool want_dummy = g_dummy_vote != nullptr;
  const bool have_sw = up->g_soft_winner != nullptr;

  const int imgs = caps_imgs_per_cta(V);
  const size_t smem = caps_bwd_smem_floats(imgs, O, V) * sizeof(float);
  SCAE_REQUIRE(smem <= (size_t)max_smem_optin(), SCAE_ELIMIT, "caps bwd: O=%d needs %zu bytes of shared memory", O, smem);
  const bool sim = (a->flags & SCAE_CAPS_SIMILARITY) != 0;
  auto kern = sim ? caps_ll_bwd_kernel<true> : caps_ll_bwd_kernel<false>;
  if (smem > 48 * 1024) SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CapsBwdOut out{g_all_param, g_x, g_presence, (want_dummy && have_sw) ? dummy_rows : nullptr};
  kern<<<(B + imgs - 1) / imgs, kCapsThreads, smem, stream>>>(*a, *saved, *up, out, imgs);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());

  const int rows_per_split = (B + nsplit - 1) / nsplit;
  dim3 grid((n + 255) / 256, nsplit);
  caps_bwd_finalize_kernel<<<grid, 256, 0, stream>>>(g_all_param, a->all_param, up->g_reg_per_example, partials, B, O, V,
                                                     a->flags, rows_per_split);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  rc = launch_reduce_rows(partials, g_shared, nsplit, n, stream);
  if (rc != SCAE_OK) return rc;
  if (want_dummy) {
    if (have_sw) {
      rc = launch_reduce_rows(dummy_rows, g_dummy_vote, B, V * 6, stream);
      if (rc != SCAE_OK) return rc;
    } else {
      SCAE_CUDA_TRY(cudaMemsetAsync(g_dummy_vote, 0, (size_t)V * 6 * sizeof(float), stream));
    }
  }
  return SCAE_OK;
}
