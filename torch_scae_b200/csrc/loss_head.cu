// Loss head (sm_100a): the (B,O)-sized tail of SCAE.loss as one kernel pair per direction.
//
// Reference code replaced: capsule_l2_loss / capsule_entropy_loss / neg_capsule_kl object_decoder.py:431-493 as called
// from SCAE.loss stacked_capsule_auto_encoder.py:243-271 (prior term on caps_presence, posterior term on
// posterior_mixing_prob.sum(-1) / n_points), the two classifier heads :203-213 (softmax(linear(x)) of the DETACHED
// capsule presences / posterior masses, both through prior_classifier -- sic) and their cross-entropies :279-285
// (F.cross_entropy applied to softmax OUTPUTS -- sic).  In stock PyTorch that is ~60 launches forward and ~70 backward
// on tensors of a few KB.  Formulas: oracle/manual_backward.py::loss_head_forward_backward (checked against autograd).
//
// Mapping: one warp per image row, lane = object capsule (two slots per lane: O <= 64).  The batch-global column sums
// that the between-example terms need are accumulated per lane across the rows of a warp, combined per CTA in warp
// order, written as one partial row per CTA and summed by a single finalising CTA in CTA order: deterministic.  The
// finaliser also leaves d(between term)/d(column sum) behind, so the backward is a single pass over the rows.
#include "common.cuh"

namespace scae {

constexpr int kHeadSlots = 2;                    // object capsules per lane
constexpr int kHeadMaxO = 32 * kHeadSlots;
constexpr int kHeadMaxK = 16;                    // classes
constexpr int kHeadWarps = 8;
constexpr int kHeadStat = 2 * kHeadMaxO + 4;     // partial row: column sums of caps_presence | of mass / V | 4 scalars
constexpr int kHeadClsMax = kHeadMaxK * kHeadMaxO + kHeadMaxK;

// caps_presence and posterior mass of row b for this lane's capsules (0 beyond O)
__device__ __forceinline__ void head_load_row(const scae_loss_head_args& a, int b, int lane, bool vec,
                                              float cp[kHeadSlots], float mass[kHeadSlots]) {
#pragma unroll
  for (int s = 0; s < kHeadSlots; ++s) {
    const int o = lane + 32 * s;
    cp[s] = 0.0f;
    mass[s] = 0.0f;
    if (o < a.O) {
      cp[s] = __ldg(a.caps_presence + (size_t)b * a.O + o);
      const float* pr = a.posterior + ((size_t)b * a.O + o) * a.V;
      float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
      if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(pr);
        for (int i = 0; i < (a.V >> 2); ++i) {
          const float4 q = __ldg(p4 + i);
          s0 += q.x;
          s1 += q.y;
          s2 += q.z;
          s3 += q.w;
        }
      } else {
        int v = 0;
        for (; v + 4 <= a.V; v += 4) {
          s0 += __ldg(pr + v);
          s1 += __ldg(pr + v + 1);
          s2 += __ldg(pr + v + 2);
          s3 += __ldg(pr + v + 3);
        }
        for (; v < a.V; ++v) s0 += __ldg(pr + v);
      }
      mass[s] = (s0 + s1) + (s2 + s3);
    }
  }
}

// Within-example term of one row (all 32 lanes call this; x[s] is capsule lane + 32 s, 0 beyond O).
//   l2:           (sum_o x - c)^2,                              d/dx_j = 2 (sum_o x - c)
//   entropy / kl: p = x / (sum_o x + 1e-8), -sum_o p log_safe(k p),  d/dx_j = (gp_j - sum_o gp_o p_o) / (sum_o x + 1e-8)
//                 with gp = -(log_safe(k p) + [k p >= 1e-16])   (log_safe passes no gradient below its threshold)
template <bool kGrad>
__device__ __forceinline__ float head_within(int type, const float x[kHeadSlots], int O, int lane, float c, float k,
                                             float grad[kHeadSlots]) {
  float part = 0.0f;
#pragma unroll
  for (int s = 0; s < kHeadSlots; ++s) part += x[s];
  const float rs = warp_sum(part);
  if (type == SCAE_LOSS_L2) {
    const float d = rs - c;
    if (kGrad) {
#pragma unroll
      for (int s = 0; s < kHeadSlots; ++s) grad[s] = 2.0f * d;
    }
    return d * d;
  }
  const float den = rs + 1e-8f;
  float gp[kHeadSlots];
  float acc = 0.0f, dacc = 0.0f;
#pragma unroll
  for (int s = 0; s < kHeadSlots; ++s) {
    const float p = x[s] / den;
    const float kp = p * k;
    const bool tiny = kp < kLogSafeEps;
    const float L = tiny ? kLogSafeFloor : logf(kp);
    gp[s] = -(L + (tiny ? 0.0f : 1.0f));
    if (lane + 32 * s < O) {
      acc = fmaf(p, L, acc);
      dacc = fmaf(gp[s], p, dacc);
    }
  }
  const float t = -warp_sum(acc);
  if (kGrad) {
    const float dot = warp_sum(dacc);
#pragma unroll
    for (int s = 0; s < kHeadSlots; ++s) grad[s] = (gp[s] - dot) / den;
  }
  return t;
}

// Between-example term from the column sums col[O] (serial, one thread) and its gradient w.r.t. them.
__device__ float head_between(int type, const float* col, int O, float c, float* dcol) {
  if (type == SCAE_LOSS_L2) {
    float acc = 0.0f;
    for (int o = 0; o < O; ++o) {
      const float d = col[o] - c;
      acc = fmaf(d, d, acc);
      dcol[o] = 2.0f * d / (float)O;
    }
    return acc / (float)O;
  }
  const float k = type == SCAE_LOSS_KL ? (float)O : 1.0f;
  float S = 0.0f;
  for (int o = 0; o < O; ++o) S += col[o];
  const float den = S + 1e-8f;
  float acc = 0.0f, dacc = 0.0f;
  for (int o = 0; o < O; ++o) {
    const float p = col[o] / den;
    const float kp = p * k;
    const bool tiny = kp < kLogSafeEps;
    const float L = tiny ? kLogSafeFloor : logf(kp);
    const float gp = -(L + (tiny ? 0.0f : 1.0f));
    acc = fmaf(p, L, acc);
    dacc = fmaf(gp, p, dacc);
    dcol[o] = gp;
  }
  // the entropy -acc is to be INCREASED: the loss term is its negation (object_decoder.py:469)
  for (int o = 0; o < O; ++o) dcol[o] = -(dcol[o] - dacc) / den;
  return acc;
}

// One classifier head on one row: z = W x + bias, p = softmax(z), xe = -log_softmax(p)[y] (sic: on the probabilities).
// Every lane ends up with all of p (and, with kGrad, dz = d xe / d z = p * (r - p.r), r = softmax(p) - onehot(y)).
template <bool kGrad>
__device__ __forceinline__ float head_classifier(const float w[kHeadMaxK][kHeadSlots], const float* __restrict__ bias,
                                                 int K, const float x[kHeadSlots], int y, float p[kHeadMaxK],
                                                 float dz[kHeadMaxK]) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < kHeadMaxK; ++k) {
    p[k] = 0.0f;
    if (k < K) {
      float part = 0.0f;
#pragma unroll
      for (int s = 0; s < kHeadSlots; ++s) part = fmaf(w[k][s], x[s], part);
      p[k] = warp_sum(part) + __ldg(bias + k);
      m = fmaxf(m, p[k]);
    }
  }
  float Z = 0.0f;
#pragma unroll
  for (int k = 0; k < kHeadMaxK; ++k) {
    if (k < K) {
      p[k] = expf(p[k] - m);
      Z += p[k];
    }
  }
  float m2 = -INFINITY, py = 0.0f;
#pragma unroll
  for (int k = 0; k < kHeadMaxK; ++k) {
    if (k < K) {
      p[k] = p[k] / Z;
      m2 = fmaxf(m2, p[k]);
      if (k == y) py = p[k];
    }
  }
  float e2[kHeadMaxK];
  float Z2 = 0.0f;
#pragma unroll
  for (int k = 0; k < kHeadMaxK; ++k) {
    e2[k] = 0.0f;
    if (k < K) {
      e2[k] = expf(p[k] - m2);
      Z2 += e2[k];
    }
  }
  if (kGrad) {
    float dot = 0.0f;
#pragma unroll
    for (int k = 0; k < kHeadMaxK; ++k) {
      dz[k] = 0.0f;
      if (k < K) {
        dz[k] = e2[k] / Z2 - (k == y ? 1.0f : 0.0f);   // r_k for now
        dot = fmaf(p[k], dz[k], dot);
      }
    }
#pragma unroll
    for (int k = 0; k < kHeadMaxK; ++k) dz[k] = p[k] * (dz[k] - dot);
  }
  return (m2 + logf(Z2)) - py;
}

__device__ __forceinline__ void head_load_weights(const scae_loss_head_args& a, int lane,
                                                  float w[kHeadMaxK][kHeadSlots]) {
#pragma unroll
  for (int k = 0; k < kHeadMaxK; ++k) {
#pragma unroll
    for (int s = 0; s < kHeadSlots; ++s) {
      const int o = lane + 32 * s;
      w[k][s] = (a.label != nullptr && k < a.K && o < a.O) ? __ldg(a.cls_weight + k * a.O + o) : 0.0f;
    }
  }
}

__device__ __forceinline__ float head_type_k(int type, int O) { return type == SCAE_LOSS_KL ? (float)O : 1.0f; }

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kHeadWarps) loss_head_fwd_kernel(const scae_loss_head_args a,
                                                                        float* __restrict__ cls_prob,
                                                                        float* __restrict__ partials, int vec) {
  __shared__ float red[kHeadWarps][kHeadStat];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kHeadWarps + warp, n_warps = gridDim.x * kHeadWarps;
  const bool cls = a.label != nullptr;
  float w[kHeadMaxK][kHeadSlots];
  head_load_weights(a, lane, w);
  float col_cp[kHeadSlots], col_pm[kHeadSlots];
#pragma unroll
  for (int s = 0; s < kHeadSlots; ++s) col_cp[s] = col_pm[s] = 0.0f;
  float sum_pw = 0.0f, sum_qw = 0.0f, sum_x1 = 0.0f, sum_x2 = 0.0f;
  for (int b = gw; b < a.B; b += n_warps) {
    float cp[kHeadSlots], mass[kHeadSlots], pm[kHeadSlots];
    head_load_row(a, b, lane, vec != 0, cp, mass);
#pragma unroll
    for (int s = 0; s < kHeadSlots; ++s) pm[s] = mass[s] / (float)a.V;
    if (a.sparsity) {
      sum_pw += head_within<false>(a.prior_type, cp, a.O, lane, a.prior_within_constant, head_type_k(a.prior_type, a.O),
                                   nullptr);
      sum_qw += head_within<false>(a.posterior_type, pm, a.O, lane, a.posterior_within_constant,
                                   head_type_k(a.posterior_type, a.O), nullptr);
#pragma unroll
      for (int s = 0; s < kHeadSlots; ++s) {
        col_cp[s] += cp[s];
        col_pm[s] += pm[s];
      }
    }
    if (cls) {
      const int y = (int)a.label[b];
      float p[kHeadMaxK];
      sum_x1 += head_classifier<false>(w, a.cls_bias, a.K, cp, y, p, nullptr);
      float mine = 0.0f;
#pragma unroll
      for (int k = 0; k < kHeadMaxK; ++k)
        if (lane == k) mine = p[k];
      if (cls_prob && lane < a.K) cls_prob[(size_t)b * a.K + lane] = mine;
      sum_x2 += head_classifier<false>(w, a.cls_bias, a.K, mass, y, p, nullptr);
#pragma unroll
      for (int k = 0; k < kHeadMaxK; ++k)
        if (lane == k) mine = p[k];
      if (cls_prob && lane < a.K) cls_prob[((size_t)a.B + b) * a.K + lane] = mine;
    }
  }
#pragma unroll
  for (int s = 0; s < kHeadSlots; ++s) {
    red[warp][lane + 32 * s] = col_cp[s];
    red[warp][kHeadMaxO + lane + 32 * s] = col_pm[s];
  }
  if (lane < 4) {
    const float v = lane == 0 ? sum_pw : lane == 1 ? sum_qw : lane == 2 ? sum_x1 : sum_x2;
    red[warp][2 * kHeadMaxO + lane] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < kHeadStat; j += blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < kHeadWarps; ++q) t += red[q][j];
    partials[(size_t)blockIdx.x * kHeadStat + j] = t;
  }
}

// terms[8] = {prior within, prior between, posterior within, posterior between, prior xe, posterior xe, weighted total, 0}
// stats[2 * kHeadMaxO] = d(prior between)/d colsum(caps_presence) | d(posterior between)/d colsum(mass / V)
__global__ void __launch_bounds__(256) loss_head_finalize_kernel(const scae_loss_head_args a,
                                                                 const float* __restrict__ partials, int n_parts,
                                                                 float* __restrict__ terms, float* __restrict__ stats) {
  __shared__ float tot[kHeadStat];
  __shared__ float dcol[2 * kHeadMaxO];
  for (int j = threadIdx.x; j < kHeadStat; j += blockDim.x) {
    float t = 0.0f;
    for (int c = 0; c < n_parts; ++c) t += partials[(size_t)c * kHeadStat + j];
    tot[j] = t;
  }
  for (int j = threadIdx.x; j < 2 * kHeadMaxO; j += blockDim.x) dcol[j] = 0.0f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[8];
    for (int i = 0; i < 8; ++i) t[i] = 0.0f;
    const float Bf = (float)a.B;
    if (a.sparsity) {
      t[0] = tot[2 * kHeadMaxO + 0] / Bf;
      t[1] = head_between(a.prior_type, tot, a.O, a.between_constant, dcol);
      t[2] = tot[2 * kHeadMaxO + 1] / Bf;
      t[3] = head_between(a.posterior_type, tot + kHeadMaxO, a.O, a.between_constant, dcol + kHeadMaxO);
    }
    if (a.label != nullptr) {
      t[4] = tot[2 * kHeadMaxO + 2] / Bf;
      t[5] = tot[2 * kHeadMaxO + 3] / Bf;
    }
    t[6] = a.prior_within_weight * t[0] + a.prior_between_weight * t[1] + a.posterior_within_weight * t[2] +
           a.posterior_between_weight * t[3] + t[4] + t[5];
    for (int i = 0; i < 8; ++i) terms[i] = t[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * kHeadMaxO; j += blockDim.x) stats[j] = dcol[j];
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kHeadWarps) loss_head_bwd_kernel(const scae_loss_head_args a,
                                                                        const float* __restrict__ stats,
                                                                        const float* __restrict__ g_total,
                                                                        float* __restrict__ g_cp,
                                                                        float* __restrict__ g_post,
                                                                        float* __restrict__ partials, int vec) {
  __shared__ float acc[kHeadClsMax];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kHeadWarps + warp, n_warps = gridDim.x * kHeadWarps;
  const bool cls = a.label != nullptr;
  const int n_cls = cls ? a.K * a.O + a.K : 0;
  const float G = __ldg(g_total);
  const float inv_B = 1.0f / (float)a.B;
  float w[kHeadMaxK][kHeadSlots];
  head_load_weights(a, lane, w);
  float d_cp[kHeadSlots], d_pm[kHeadSlots];
#pragma unroll
  for (int s = 0; s < kHeadSlots; ++s) {
    const int o = lane + 32 * s;
    d_cp[s] = (a.sparsity && o < a.O) ? __ldg(stats + o) : 0.0f;
    d_pm[s] = (a.sparsity && o < a.O) ? __ldg(stats + kHeadMaxO + o) : 0.0f;
  }
  float gW[kHeadMaxK][kHeadSlots], gb[kHeadMaxK];
#pragma unroll
  for (int k = 0; k < kHeadMaxK; ++k) {
    gb[k] = 0.0f;
#pragma unroll
    for (int s = 0; s < kHeadSlots; ++s) gW[k][s] = 0.0f;
  }
  for (int i = threadIdx.x; i < n_cls; i += blockDim.x) acc[i] = 0.0f;

  for (int b = gw; b < a.B; b += n_warps) {
    float cp[kHeadSlots], mass[kHeadSlots], pm[kHeadSlots];
    head_load_row(a, b, lane, vec != 0, cp, mass);
#pragma unroll
    for (int s = 0; s < kHeadSlots; ++s) pm[s] = mass[s] / (float)a.V;
    if (a.sparsity) {
      float g1[kHeadSlots], g2[kHeadSlots];
      head_within<true>(a.prior_type, cp, a.O, lane, a.prior_within_constant, head_type_k(a.prior_type, a.O), g1);
      head_within<true>(a.posterior_type, pm, a.O, lane, a.posterior_within_constant,
                        head_type_k(a.posterior_type, a.O), g2);
#pragma unroll
      for (int s = 0; s < kHeadSlots; ++s) {
        const int o = lane + 32 * s;
        if (o < a.O) {
          if (g_cp)
            g_cp[(size_t)b * a.O + o] = G * (a.prior_within_weight * g1[s] * inv_B + a.prior_between_weight * d_cp[s]);
          if (g_post) {
            // d / d mass = (d / d (mass / V)) / V, the same for every part v of the capsule
            const float gm =
                G * (a.posterior_within_weight * g2[s] * inv_B + a.posterior_between_weight * d_pm[s]) / (float)a.V;
            float* dst = g_post + ((size_t)b * a.O + o) * a.V;
            if (vec) {
              const float4 q = make_float4(gm, gm, gm, gm);
              float4* d4 = reinterpret_cast<float4*>(dst);
              for (int i = 0; i < (a.V >> 2); ++i) d4[i] = q;
            } else {
              for (int v = 0; v < a.V; ++v) dst[v] = gm;
            }
          }
        }
      }
    }
    if (cls) {
      const int y = (int)a.label[b];
      float p[kHeadMaxK], dz1[kHeadMaxK], dz2[kHeadMaxK];
      head_classifier<true>(w, a.cls_bias, a.K, cp, y, p, dz1);
      head_classifier<true>(w, a.cls_bias, a.K, mass, y, p, dz2);
      const float scale = G * inv_B;
#pragma unroll
      for (int k = 0; k < kHeadMaxK; ++k) {
        if (k < a.K) {
          const float z1 = scale * dz1[k], z2 = scale * dz2[k];
          gb[k] += z1 + z2;
#pragma unroll
          for (int s = 0; s < kHeadSlots; ++s) gW[k][s] = fmaf(z1, cp[s], fmaf(z2, mass[s], gW[k][s]));
        }
      }
    }
  }
  if (cls) {
    // per-CTA sums of the classifier gradients, warp after warp in a fixed order
    __syncthreads();
    for (int q = 0; q < kHeadWarps; ++q) {
      if (warp == q) {
#pragma unroll
        for (int k = 0; k < kHeadMaxK; ++k) {
          if (k < a.K) {
#pragma unroll
            for (int s = 0; s < kHeadSlots; ++s) {
              const int o = lane + 32 * s;
              if (o < a.O) acc[k * a.O + o] += gW[k][s];
            }
            if (lane == 0) acc[a.K * a.O + k] += gb[k];
          }
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < n_cls; i += blockDim.x) partials[(size_t)blockIdx.x * n_cls + i] = acc[i];
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)
static int head_grid(int B) {
  const int want = (B + kHeadWarps - 1) / kHeadWarps;
  const int cap = 2 * sm_count();
  return want < cap ? want : cap;
}

static bool head_type_ok(int t) { return t == SCAE_LOSS_L2 || t == SCAE_LOSS_ENTROPY || t == SCAE_LOSS_KL; }

static bool head_shape_ok(const scae_loss_head_args* a) {
  if (a == nullptr || a->B <= 0 || a->O <= 0 || a->O > kHeadMaxO || a->V <= 0) return false;
  if (a->label != nullptr && (a->K <= 0 || a->K > kHeadMaxK)) return false;
  if (a->sparsity && !(head_type_ok(a->prior_type) && head_type_ok(a->posterior_type))) return false;
  return true;
}

static size_t head_workspace_floats(const scae_loss_head_args* a) {
  const size_t per_cta = a->label != nullptr && a->K * a->O + a->K > kHeadStat ? (size_t)(a->K * a->O + a->K) : kHeadStat;
  return (size_t)head_grid(a->B) * per_cta;
}

}  // namespace scae

#define SCAE_EXPORT __attribute__((visibility("default")))
extern "C" {

SCAE_EXPORT size_t scae_loss_head_workspace_bytes(const scae_loss_head_args* a) {
  if (!scae::head_shape_ok(a)) return 0;
  return scae::head_workspace_floats(a) * sizeof(float);
}

SCAE_EXPORT int scae_loss_head_fwd(const scae_loss_head_args* a, float* terms, float* cls_prob, float* stats,
                                   void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(head_shape_ok(a), SCAE_ELIMIT, "loss_head: needs B, V > 0, 0 < O <= %d, K <= %d and known loss types",
               kHeadMaxO, kHeadMaxK);
  SCAE_REQUIRE(a->caps_presence && a->posterior && terms && stats, SCAE_EINVAL,
               "loss_head: caps_presence, posterior, terms and stats are required");
  SCAE_REQUIRE(a->label == nullptr || (a->cls_weight && a->cls_bias), SCAE_EINVAL,
               "loss_head: labels need the classifier weight and bias");
  SCAE_REQUIRE(workspace && workspace_bytes >= head_workspace_floats(a) * sizeof(float), SCAE_EINVAL,
               "loss_head: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = head_grid(a->B);
  const int vec = (a->V % 4 == 0) && aligned16(a->posterior);
  float* partials = static_cast<float*>(workspace);
  loss_head_fwd_kernel<<<grid, 32 * kHeadWarps, 0, stream>>>(*a, cls_prob, partials, vec);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  loss_head_finalize_kernel<<<1, 256, 0, stream>>>(*a, partials, grid, terms, stats);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

// The forward in two halves, so that a data-parallel caller can sum the batch-global column statistics over its ranks in
// between (the two O-float all-reduces of the between-example sparsity terms):
//   rows   : per-row terms + this shard's column sums -> colsums[2 * 64 + 4] = [sum_b caps_presence | sum_b mass / V |
//            sum of prior within, posterior within, prior xe, posterior xe over the rows]
//   finish : terms / stats from (possibly all-reduced) colsums; a->between_constant is the GLOBAL batch / n_classes
SCAE_EXPORT int scae_loss_head_fwd_rows(const scae_loss_head_args* a, float* cls_prob, float* colsums, void* workspace,
                                        size_t workspace_bytes, scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(head_shape_ok(a), SCAE_ELIMIT, "loss_head: needs B, V > 0, 0 < O <= %d, K <= %d and known loss types",
               kHeadMaxO, kHeadMaxK);
  SCAE_REQUIRE(a->caps_presence && a->posterior && colsums, SCAE_EINVAL,
               "loss_head: caps_presence, posterior and colsums are required");
  SCAE_REQUIRE(a->label == nullptr || (a->cls_weight && a->cls_bias), SCAE_EINVAL,
               "loss_head: labels need the classifier weight and bias");
  SCAE_REQUIRE(workspace && workspace_bytes >= head_workspace_floats(a) * sizeof(float), SCAE_EINVAL,
               "loss_head: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = head_grid(a->B);
  const int vec = (a->V % 4 == 0) && aligned16(a->posterior);
  float* partials = static_cast<float*>(workspace);
  loss_head_fwd_kernel<<<grid, 32 * kHeadWarps, 0, stream>>>(*a, cls_prob, partials, vec);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return launch_reduce_rows(partials, colsums, grid, kHeadStat, stream);
}

SCAE_EXPORT int scae_loss_head_fwd_finish(const scae_loss_head_args* a, const float* colsums, float* terms, float* stats,
                                          scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(head_shape_ok(a), SCAE_ELIMIT, "loss_head: needs B, V > 0, 0 < O <= %d, K <= %d and known loss types",
               kHeadMaxO, kHeadMaxK);
  SCAE_REQUIRE(colsums && terms && stats, SCAE_EINVAL, "loss_head: colsums, terms and stats are required");
  loss_head_finalize_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(*a, colsums, 1, terms, stats);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT int scae_loss_head_bwd(const scae_loss_head_args* a, const float* stats, const float* g_total,
                                   float* g_caps_presence, float* g_posterior, float* g_cls, void* workspace,
                                   size_t workspace_bytes, scae_stream_t stream_) {
  using namespace scae;
  SCAE_REQUIRE(head_shape_ok(a), SCAE_ELIMIT, "loss_head: needs B, V > 0, 0 < O <= %d, K <= %d and known loss types",
               kHeadMaxO, kHeadMaxK);
  SCAE_REQUIRE(a->caps_presence && a->posterior && stats && g_total, SCAE_EINVAL,
               "loss_head: caps_presence, posterior, stats and g_total are required");
  SCAE_REQUIRE(a->label == nullptr || (a->cls_weight && a->cls_bias && g_cls), SCAE_EINVAL,
               "loss_head: labels need the classifier weight, bias and g_cls");
  SCAE_REQUIRE(workspace && workspace_bytes >= head_workspace_floats(a) * sizeof(float), SCAE_EINVAL,
               "loss_head: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = head_grid(a->B);
  const int vec = (a->V % 4 == 0) && aligned16(a->posterior) && (g_posterior == nullptr || aligned16(g_posterior));
  float* partials = static_cast<float*>(workspace);
  loss_head_bwd_kernel<<<grid, 32 * kHeadWarps, 0, stream>>>(*a, stats, g_total, g_caps_presence, g_posterior, partials,
                                                             vec);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  if (a->label != nullptr) return launch_reduce_rows(partials, g_cls, grid, a->K * a->O + a->K, stream);
  return SCAE_OK;
}

}  // extern "C"
