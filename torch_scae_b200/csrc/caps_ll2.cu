// Hot path 2 (sm_100a), fast path: pair-parallel capsule-likelihood kernels staged by TMA bulk copies.
//
// Same math as caps_ll.cu (reference object_decoder.py:160-236, :257-372, :413-415; oracle/capsule_likelihood.py and
// oracle/manual_backward.py::capsule_forward_backward), different decomposition.  caps_ll.cu gives one thread a whole
// (image, part) column of O object capsules; at a training batch of 1024 that is 9 warps per SM walking 32 objects
// serially -- latency bound at a quarter of the HBM roofline (profiles/r01f).  Here a CTA owns one IMAGE at a time:
//
//   * the image's all_param block (O rows of 8V+7 floats, contiguous) and its noise / saved-posterior rows arrive by
//     cp.async.bulk (UBLKCP) on an mbarrier -- one elected thread issues, nobody spends LSU slots on staging;
//   * one thread per (object, part) PAIR, pairs flattened o*V+v so every (B,O,V) tensor is read and written unit-stride;
//   * the reductions over objects (logsumexp, argmax winner, soft winner) and over parts (capsule presence; in the
//     backward the capsule-level gradient sums) run on shared-memory tiles with 8 lanes per output row;
//   * forward: the (O,V,6) vote block leaves through one bulk store; backward: the gradient rows are produced IN PLACE
//     in the staged all_param buffer and leave through one bulk store, with the MLP-ReLU mask and the deformation
//     regulariser already applied, so g_all_param is written exactly once and never re-read;
//   * backward CTAs are persistent (one per SM, double-buffered prefetch of the next image) so the batch sums that
//     become the gradients of cpr_static and the four biases accumulate in shared memory in a fixed image order:
//     deterministic, and the separate finalisation pass of the general path disappears.
//
// The logsumexp over objects needs no running maximum: the dummy component bounds it below (2 log 0.01) and
// scale >= 0.01 bounds every posterior logit above (about 22), so sum_o exp(logit) is safe in fp32.
#include <stdlib.h>
#include <string.h>

#include "caps_common.cuh"

namespace scae {

constexpr int kF2MaxThreads = 512;   // forward: 2 CTAs per SM (<= 64 registers per thread)
constexpr int kB2MaxThreads = 640;   // backward: 1 persistent CTA per SM (<= 102 registers per thread)
constexpr int kIntMax = 0x7fffffff;

__host__ __device__ inline int round_up4(int n) { return (n + 3) & ~3; }

// ================================================================================================================
// forward
// ================================================================================================================
struct Caps2FwdLayout {   // offsets in floats from the start of dynamic shared memory; every region 16-byte aligned
  int prm, nz, E, PL, VP, VOTE, R, XS, ACC6, SV, INVS, SWP, WIDX, MLSE, RED, SCR, total;
};

static Caps2FwdLayout caps2_fwd_layout(int O, int V, bool noise) {
  const int A = 8 * V + 7, P = O * V, Vp = V | 1;
  Caps2FwdLayout L;
  int at = 4;                                   // [0, 2): the mbarrier
  auto take = [&](int n) {
    const int here = at;
    at += round_up4(n);
    return here;
  };
  L.prm = take(O * A + 4);
  L.E = take(O * Vp);
  L.PL = take(O * Vp);
  L.VP = take(O * Vp);
  L.VOTE = take(P * 6);
  L.R = take(O * 8);
  L.XS = take(V * 8);
  // the staged noise rows are dead after phase A; the per-part results of phases B / C reuse their space
  const int late0 = at;
  L.ACC6 = take(V * 6);
  L.SV = take(V);
  L.INVS = take(V);
  L.SWP = take(V);
  L.WIDX = take(V);
  L.MLSE = take(V);
  L.RED = take(V);
  L.SCR = take(32);
  const int late1 = at;
  L.nz = late0;
  if (noise && late0 + round_up4(P + 4) > late1) at = late0 + round_up4(P + 4);
  L.total = at;
  return L;
}

// deterministic block sum (fixed shuffle tree, fixed warp order); result valid in thread 0
__device__ __forceinline__ float block_sum_t0(float v, float* scr) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) scr[warp] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += scr[i];
  }
  return t;
}

// The batch-shared per-pair parameters come from global memory (L1 / L2 resident).  They are fetched one pass ahead so
// that their latency hides behind the previous pair's arithmetic.
struct PairConst {
  float2 s0, s1, s2;   // cpr_static[o][v][0..5]
  float bv, bs;        // bias_vote[o][v], bias_scale[o][v]
};
__device__ __forceinline__ PairConst load_pair_const(const scae_caps_args& a, int p) {
  PairConst c;
  const float2* sp = reinterpret_cast<const float2*>(a.cpr_static + (size_t)p * 6);
  c.s0 = __ldg(sp);
  c.s1 = __ldg(sp + 1);
  c.s2 = __ldg(sp + 2);
  c.bv = __ldg(a.bias_vote + p);
  c.bs = __ldg(a.bias_scale + p);
  return c;
}

template <bool kSim>
__global__ void __launch_bounds__(kF2MaxThreads, 2) caps2_fwd_kernel(const scae_caps_args a, const scae_caps_outputs o,
                                                                     const Caps2FwdLayout L) {
  SCAE_DYNAMIC_SMEM(smem);
  const int tid = threadIdx.x, T = blockDim.x, b = blockIdx.x;
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V, Vp = V | 1;
  const float inv_V = 1.0f / (float)V;
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;
  const unsigned bar = smem_u32(smem);
  float* prm_base = smem + L.prm;
  float* nz_base = smem + L.nz;
  float* Es = smem + L.E;
  float* PLs = smem + L.PL;
  float* VPs = smem + L.VP;
  float* VOTE = smem + L.VOTE;
  float* R = smem + L.R;
  float* XS = smem + L.XS;
  float* ACC6 = smem + L.ACC6;
  float* SV = smem + L.SV;
  float* INVS = smem + L.INVS;
  float* SWP = smem + L.SWP;
  int* WIDX = reinterpret_cast<int*>(smem + L.WIDX);
  float* MLSE = smem + L.MLSE;
  float* RED = smem + L.RED;
  float* SCR = smem + L.SCR;

  // ---- stage this image's rows: bulk copies for the 16-byte-aligned interiors, registers for the edge floats ----------
  const float* gprm = a.all_param + (size_t)b * O * A;
  const float* gnz = a.noise_vote ? a.noise_vote + (size_t)b * P : nullptr;
  const BulkRun rp = bulk_run(gprm, O * A);
  BulkRun rn = {0, 0, 0, 0};
  if (gnz) rn = bulk_run(gnz, P);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, 4u * (unsigned)(rp.body + rn.body));
    if (rp.body) bulk_g2s(prm_base + rp.off + rp.head, gprm + rp.head, 4u * (unsigned)rp.body, bar);
    if (rn.body) bulk_g2s(nz_base + rn.off + rn.head, gnz + rn.head, 4u * (unsigned)rn.body, bar);
  }
  if (tid < 32) bulk_run_edges_in(prm_base, gprm, rp, tid);
  if (gnz && tid >= 32 && tid < 64) bulk_run_edges_in(nz_base, gnz, rn, tid - 32);
  // part poses and presences of the image: XS[v] = {x[0..5], presence, -}
  for (int i = tid; i < V * 8; i += T) {
    const int v = i >> 3, c = i & 7;
    const size_t bv = (size_t)b * V + v;
    float t = 0.0f;
    if (c < 6) t = __ldg(a.x + bv * 6 + c);
    else if (c == 6) t = a.presence ? __ldg(a.presence + bv) : 1.0f;
    XS[i] = t;
  }
  // first pass of the pair loop: its batch-shared parameters travel while the bulk copy lands
  PairConst cur = {};
  if (tid < P) cur = load_pair_const(a, tid);
  mbar_wait(bar, 0);
  __syncthreads();
  const float* prm = prm_base + rp.off;
  const float* nz = gnz ? nz_base + rn.off : nullptr;

  // ---- phase 0: capsule-level transform and presence, one thread per object ---------------------------------------------
  for (int oo = tid; oo < O; oo += T) {
    const float* row = prm + oo * A;
    float t[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) t[p] = row[6 * V + p] + __ldg(a.bias_cvr + oo * 6 + p);
    PoseAffine r;
    pose_affine_fast<kSim>(t, r);
    float lc = row[6 * V + 6] + __ldg(a.bias_caps + oo);
    if (a.noise_caps) lc += __ldg(a.noise_caps + (size_t)b * O + oo);
    float* dst = R + oo * 8;
#pragma unroll
    for (int p = 0; p < 6; ++p) dst[p] = r.a[p];
    dst[6] = sigmoid_fast(lc);
    dst[7] = lc;
    if (o.presence_logit_per_caps) o.presence_logit_per_caps[(size_t)b * O + oo] = lc;
  }
  __syncthreads();

  // ---- phase A: one thread per (object, part) pair ------------------------------------------------------------------------
  float regsum = 0.0f;
  const size_t bP = (size_t)b * P, bP1 = (size_t)b * (O + 1) * V;
  const int step_o = fast_div(T, inv_V), step_v = T - step_o * V;   // (o, v) advance of one pass of T pairs
  {
    int oo = fast_div(tid, inv_V), v = tid - oo * V;
    for (int p = tid; p < P; p += T) {
      PairConst nxt = cur;
      if (p + T < P) nxt = load_pair_const(a, p + T);
      const float* row = prm + oo * A;
      const float* r = R + oo * 8;
      const float st[6] = {cur.s0.x, cur.s0.y, cur.s1.x, cur.s1.y, cur.s2.x, cur.s2.y};
      float t[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float d = deform ? row[6 * v + c] : 0.0f;
        regsum = fmaf(d, d, regsum);
        t[c] = d + st[c];
      }
      PoseAffine pa;
      pose_affine_fast<kSim>(t, pa);
      float vt[6];
      compose_vote(r, pa.a, vt);
      float lv = row[6 * V + 7 + v] + cur.bv;
      if (nz) lv += nz[p];
      const float vp = r[6] * sigmoid_fast(lv);
      const float u = row[7 * V + 7 + v] + cur.bs;
      const float sc = learn ? softplus_fast(u + 0.5f) + 1e-2f : 1.0f;
      const float* xv = XS + v * 8;
      float q = 0.0f;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float d = xv[c] - vt[c];
        q = fmaf(d, d, q);
      }
      // sum over the 6 pose dims of Normal(vote, sc).log_prob(x)
      const float lp = -q * rcp_approx(2.0f * sc * sc) - 6.0f * logf(sc) - 6.0f * kHalfLog2Pi;
      const float ml = log_safe_fast(vp);
      const float pl = ml + lp;
      const int ov = oo * Vp + v;
      Es[ov] = expf(pl);
      PLs[ov] = pl;
      VPs[ov] = vp;
      float2* vdst = reinterpret_cast<float2*>(VOTE + p * 6);
      vdst[0] = make_float2(vt[0], vt[1]);
      vdst[1] = make_float2(vt[2], vt[3]);
      vdst[2] = make_float2(vt[4], vt[5]);
      if (o.scale) o.scale[bP + p] = sc;
      if (o.vote_presence) o.vote_presence[bP + p] = vp;
      if (o.presence_logit_per_vote) o.presence_logit_per_vote[bP + p] = lv;
      if (o.vote_presence_binary) o.vote_presence_binary[bP + p] = ml > kDummyLog ? 1.0f : 0.0f;
      if (o.mixing_logit) o.mixing_logit[bP1 + p] = ml;
      cur = nxt;
      oo += step_o;
      v += step_v;
      if (v >= V) {
        v -= V;
        ++oo;
      }
    }
  }
  fence_proxy_async();   // the vote tile is read by the bulk store below
  __syncthreads();

  // the (O,V,6) vote block of this image: one bulk store when it is 16-byte shaped, a coalesced copy loop otherwise
  bool bulk_vote = false;
  if (o.vote) {
    float* gv = o.vote + bP * 6;
    bulk_vote = (P & 1) == 0 && (reinterpret_cast<uintptr_t>(gv) & 15u) == 0;
    if (bulk_vote) {
      if (tid == 0) {
        bulk_s2g(gv, VOTE, (unsigned)P * 24u);
        bulk_commit();
      }
    } else {
      for (int i = tid; i < P * 6; i += T) gv[i] = VOTE[i];
    }
  }

  // ---- phase B: reductions over objects and over parts ------------------------------------------------------------------------
  const float e_dummy = expf(kDummyLog + kDummyLog);   // dummy mixing logit + dummy log-density (object_decoder.py:273-292)
  const float p_dummy = expf(kDummyLog);
  // (a) unnormalised soft winner: one thread per (part, pose dim), sum_o E * vote
  {
    const int vstride = V * 6;
    for (int idx = tid; idx < vstride; idx += T) {
      const int v = idx / 6;
      const float* pe = Es + v;
      const float* pv = VOTE + idx;
      float a0 = 0.0f, a1 = 0.0f;
      int oo = 0;
      for (; oo + 2 <= O; oo += 2) {
        a0 = fmaf(pe[0], pv[0], a0);
        a1 = fmaf(pe[Vp], pv[vstride], a1);
        pe += 2 * Vp;
        pv += 2 * vstride;
      }
      if (oo < O) a0 = fmaf(pe[0], pv[0], a0);
      ACC6[idx] = fmaf(e_dummy, __ldg(a.dummy_vote + idx), a0 + a1);
    }
  }
  // (b) 8 lanes per part, each over a chunk of objects: sum E, sum E * presence, sum presence, argmax of the logit
  {
    const int n_items = (V * 8 + 31) & ~31;
    const int chunk = (O + 7) >> 3;
    for (int idx = tid; idx < n_items; idx += T) {
      const int v = idx >> 3, k = idx & 7;
      const bool ok = v < V;
      float sE = 0.0f, sEvp = 0.0f, svp = 0.0f, best = -INFINITY;
      int bidx = kIntMax;
      if (ok) {
        const int o1 = min(O, (k + 1) * chunk);
        for (int oo = k * chunk; oo < o1; ++oo) {
          const int i = oo * Vp + v;
          const float e = Es[i], vp = VPs[i], pl = PLs[i];
          sE += e;
          sEvp = fmaf(e, vp, sEvp);
          svp += vp;
          if (pl > best) {   // lowest o on ties (torch.argmax)
            best = pl;
            bidx = oo;
          }
        }
      }
#pragma unroll
      for (int d = 1; d < 8; d <<= 1) {
        sE += __shfl_xor_sync(0xffffffffu, sE, d);
        sEvp += __shfl_xor_sync(0xffffffffu, sEvp, d);
        svp += __shfl_xor_sync(0xffffffffu, svp, d);
        const float ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, d);
        if (ob > best || (ob == best && oi < bidx)) {
          best = ob;
          bidx = oi;
        }
      }
      if (ok && k == 0) {
        const float S = sE + e_dummy;
        SV[v] = S;
        INVS[v] = __frcp_rn(S);
        SWP[v] = sEvp;                       // the dummy component has presence 0
        WIDX[v] = bidx == kIntMax ? 0 : bidx;
        MLSE[v] = logf(svp + p_dummy);       // logsumexp_o of the mixing logits: exp(log_safe(vp)) = vp
      }
    }
  }
  // (c) capsule presence = max over parts of vote_presence (object_decoder.py:415); lowest part index wins ties
  {
    const int n_items = (O * 8 + 31) & ~31;
    const int chunk = (V + 7) >> 3;
    for (int idx = tid; idx < n_items; idx += T) {
      const int oo = idx >> 3, k = idx & 7;
      const bool ok = oo < O;
      float best = -INFINITY;
      int arg = kIntMax;
      if (ok) {
        const int v1 = min(V, (k + 1) * chunk);
        for (int v = k * chunk; v < v1; ++v) {
          const float t = VPs[oo * Vp + v];
          if (t > best) {
            best = t;
            arg = v;
          }
        }
      }
#pragma unroll
      for (int d = 1; d < 8; d <<= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, d);
        if (ob > best || (ob == best && oi < arg)) {
          best = ob;
          arg = oi;
        }
      }
      if (ok && k == 0) {
        if (arg == kIntMax) {   // every vote presence is NaN: behave like the general path (first element)
          arg = 0;
          best = VPs[oo * Vp];
        }
        if (o.caps_presence) o.caps_presence[(size_t)b * O + oo] = best;
        if (o.caps_presence_arg) o.caps_presence_arg[(size_t)b * O + oo] = arg;
      }
    }
  }
  __syncthreads();

  // ---- phase C: normalised per-part and per-pair results -----------------------------------------------------------------------
  {
    const size_t bV6 = (size_t)b * V * 6;
    for (int idx = tid; idx < V * 6; idx += T) {
      const int v = idx / 6;
      if (o.soft_winner) o.soft_winner[bV6 + idx] = ACC6[idx] * INVS[v];
      if (o.winner) o.winner[bV6 + idx] = VOTE[WIDX[v] * V * 6 + idx];
    }
  }
  for (int v = tid; v < V; v += T) {
    const size_t bv = (size_t)b * V + v;
    const int widx = WIDX[v];
    const float lse = logf(SV[v]);
    if (o.soft_winner_presence) o.soft_winner_presence[bv] = SWP[v] * INVS[v];
    if (o.winner_presence) o.winner_presence[bv] = VPs[widx * Vp + v];
    if (o.log_prob_per_point) o.log_prob_per_point[bv] = lse;
    if (o.winner_idx) o.winner_idx[bv] = widx;
    if (o.is_from_capsule) o.is_from_capsule[bv] = fast_div(widx, inv_V);   // sic (object_decoder.py:334)
    const size_t dummy_row = bP1 + (size_t)O * V + v;
    if (o.mixing_logit) o.mixing_logit[dummy_row] = kDummyLog;
    if (o.mixing_log_prob) o.mixing_log_prob[dummy_row] = kDummyLog - MLSE[v];
    RED[v] = lse * XS[v * 8 + 6];
  }
  if (o.posterior_mixing_prob || o.mixing_log_prob) {
    int oo = fast_div(tid, inv_V), v = tid - oo * V;
    for (int p = tid; p < P; p += T) {
      const int ov = oo * Vp + v;
      if (o.posterior_mixing_prob) o.posterior_mixing_prob[bP + p] = Es[ov] * INVS[v];
      if (o.mixing_log_prob) o.mixing_log_prob[bP1 + p] = log_safe_fast(VPs[ov]) - MLSE[v];
      oo += step_o;
      v += step_v;
      if (v >= V) {
        v -= V;
        ++oo;
      }
    }
  }
  // per-example sums, fixed order
  const float reg = block_sum_t0(regsum, SCR);   // contains a __syncthreads: RED is complete afterwards
  if (tid == 0) {
    float ll = 0.0f;
    for (int v = 0; v < V; ++v) ll += RED[v];
    if (o.ll_per_example) o.ll_per_example[b] = ll;
    if (o.reg_per_example) o.reg_per_example[b] = 0.5f * reg;
    if (bulk_vote) bulk_wait_read_all();   // shared memory must outlive the bulk store's reads
  }
}

// ================================================================================================================
// backward
// ================================================================================================================
struct Caps2BwdLayout {
  int prm, post, gpost, nz, XS, OS;   // stage 0; stage 1 (when present) sits stage_stride floats further
  int stage_stride;
  int COLSUM, RED7, R, SPART, total;
  int stages;
};

static Caps2BwdLayout caps2_bwd_layout(int O, int V, bool noise, bool have_gpost, int stages) {
  const int A = 8 * V + 7, P = O * V;
  Caps2BwdLayout L;
  L.stages = stages;
  int at = 4;                                   // [0, 4): two mbarriers
  auto take = [&](int n) {
    const int here = at;
    at += round_up4(n);
    return here;
  };
  const int stage0 = at;
  L.prm = take(O * A + 4);
  L.post = take(P + 4);
  L.gpost = take(have_gpost ? P + 4 : 0);
  L.nz = take(noise ? P + 4 : 0);
  L.XS = take(V * 8);
  L.OS = take(O * 4);
  L.stage_stride = stages == 2 ? at - stage0 : 0;
  at += L.stage_stride;
  L.COLSUM = take(O * A);
  L.RED7 = take(7 * (P + 1));
  L.R = take(O * 8);
  L.SPART = take(8 * V);
  L.total = at;
  return L;
}

struct Caps2BwdOut {
  float* g_all_param;   // [B,O,A] final: ReLU mask and regulariser applied
  float* g_presence;    // [B,V] nullable
  float* partials;      // [grid][O*A] per-CTA batch sums of the pre-activation gradient
};

// everything of image `b` that the kernel wants in shared memory; called by warp 0 only (lane 0 issues the bulk copies)
__device__ __forceinline__ void caps2_bwd_stage_issue(const scae_caps_args& a, const scae_caps_saved& sv,
                                                      const scae_caps_upstream& up, const Caps2BwdLayout& L, float* smem,
                                                      int s, int b, unsigned bar, int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V;
  const float* g0 = a.all_param + (size_t)b * O * A;
  const float* g1 = sv.posterior_mixing_prob + (size_t)b * P;
  const float* g2 = up.g_posterior_mixing_prob ? up.g_posterior_mixing_prob + (size_t)b * P : nullptr;
  const float* g3 = a.noise_vote ? a.noise_vote + (size_t)b * P : nullptr;
  const BulkRun r0 = bulk_run(g0, O * A), r1 = bulk_run(g1, P);
  BulkRun r2 = {0, 0, 0, 0}, r3 = {0, 0, 0, 0};
  if (g2) r2 = bulk_run(g2, P);
  if (g3) r3 = bulk_run(g3, P);
  if (lane == 0) {
    mbar_expect_tx(bar, 4u * (unsigned)(r0.body + r1.body + r2.body + r3.body));
    if (r0.body) bulk_g2s(smem + (L.prm + s * L.stage_stride) + r0.off + r0.head, g0 + r0.head, 4u * (unsigned)r0.body, bar);
    if (r1.body) bulk_g2s(smem + (L.post + s * L.stage_stride) + r1.off + r1.head, g1 + r1.head, 4u * (unsigned)r1.body, bar);
    if (r2.body) bulk_g2s(smem + (L.gpost + s * L.stage_stride) + r2.off + r2.head, g2 + r2.head, 4u * (unsigned)r2.body, bar);
    if (r3.body) bulk_g2s(smem + (L.nz + s * L.stage_stride) + r3.off + r3.head, g3 + r3.head, 4u * (unsigned)r3.body, bar);
  }
  const int grp = lane >> 3, sub = lane & 7;   // four groups of 8 lanes, one per run
  if (grp == 0) bulk_run_edges_in(smem + (L.prm + s * L.stage_stride), g0, r0, sub);
  if (grp == 1) bulk_run_edges_in(smem + (L.post + s * L.stage_stride), g1, r1, sub);
  if (grp == 2 && g2) bulk_run_edges_in(smem + (L.gpost + s * L.stage_stride), g2, r2, sub);
  if (grp == 3 && g3) bulk_run_edges_in(smem + (L.nz + s * L.stage_stride), g3, r3, sub);
}

// XS[v] = {x[0..5], presence, saved logsumexp};  OS[o] = {noise_caps, g_caps_presence, argmax part (int bits), g_logit}
__device__ __forceinline__ float caps2_bwd_small_load(const scae_caps_args& a, const scae_caps_saved& sv,
                                                      const scae_caps_upstream& up, int b, int i) {
  const int V = a.V, O = a.O;
  if (i < V * 8) {
    const int v = i >> 3, c = i & 7;
    const size_t bv = (size_t)b * V + v;
    if (c < 6) return __ldg(a.x + bv * 6 + c);
    if (c == 6) return a.presence ? __ldg(a.presence + bv) : 1.0f;
    return __ldg(sv.log_prob_per_point + bv);
  }
  const int j = i - V * 8, oo = j >> 2, c = j & 3;
  const size_t bo = (size_t)b * O + oo;
  if (c == 0) return a.noise_caps ? __ldg(a.noise_caps + bo) : 0.0f;
  if (c == 1) return up.g_caps_presence ? __ldg(up.g_caps_presence + bo) : 0.0f;
  if (c == 2) return __int_as_float(up.g_caps_presence ? sv.caps_presence_arg[bo] : -1);
  return up.g_presence_logit_per_caps ? __ldg(up.g_presence_logit_per_caps + bo) : 0.0f;
}

constexpr int kSmallMax = 3;   // small per-image inputs per thread: V*8 + O*4 <= kSmallMax * blockDim

template <bool kSim>
__global__ void __launch_bounds__(kB2MaxThreads, 1) caps2_bwd_kernel(const scae_caps_args a, const scae_caps_saved sv,
                                                                     const scae_caps_upstream up, const Caps2BwdOut out,
                                                                     const Caps2BwdLayout L) {
  SCAE_DYNAMIC_SMEM(smem);
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V, P1 = P + 1;
  const float inv_V = 1.0f / (float)V;
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;
  const bool relu = (a.flags & SCAE_CAPS_RELU_GRAD) != 0;
  const int n_small = V * 8 + O * 4;
  float* COLSUM = smem + L.COLSUM;
  float* RED7 = smem + L.RED7;
  float* R = smem + L.R;
  float* SPART = smem + L.SPART;
  const unsigned bar0 = smem_u32(smem);

  for (int i = tid; i < O * A; i += T) COLSUM[i] = 0.0f;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int n_mine = ((int)blockIdx.x < a.B) ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // prologue: first image into stage 0
  if (n_mine > 0) {
    if (warp == 0) caps2_bwd_stage_issue(a, sv, up, L, smem, 0, blockIdx.x, bar0, lane);
    for (int i = tid; i < n_small; i += T) {
      float* dst = i < V * 8 ? smem + L.XS + i : smem + L.OS + (i - V * 8);
      *dst = caps2_bwd_small_load(a, sv, up, blockIdx.x, i);
    }
  }

  for (int it = 0; it < n_mine; ++it) {
    const int b = blockIdx.x + it * gridDim.x;
    const int s = L.stages == 2 ? (it & 1) : 0;
    const int sn = L.stages == 2 ? (s ^ 1) : 0;
    const unsigned parity = L.stages == 2 ? ((it >> 1) & 1) : (it & 1);
    const bool has_next = it + 1 < n_mine;
    const int b_next = b + gridDim.x;
    // prefetch the next image into the other stage; its previous tenant's gradient block must have left first
    if (L.stages == 2 && has_next && warp == 0) {
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
      caps2_bwd_stage_issue(a, sv, up, L, smem, sn, b_next, bar0 + 8 * sn, lane);
    }
    mbar_wait(bar0 + 8 * s, parity);
    __syncthreads();

    const BulkRun r0 = bulk_run(a.all_param + (size_t)b * O * A, O * A);
    float* prm_base = smem + (L.prm + s * L.stage_stride);
    float* prm = prm_base + r0.off;                      // inputs now, gradient rows after phase 2 / 3 (in place)
    const float* post = smem + (L.post + s * L.stage_stride) + bulk_run(sv.posterior_mixing_prob + (size_t)b * P, P).off;
    const float* gpost = up.g_posterior_mixing_prob
                             ? smem + (L.gpost + s * L.stage_stride) + bulk_run(up.g_posterior_mixing_prob + (size_t)b * P, P).off
                             : nullptr;
    const float* nz = a.noise_vote ? smem + (L.nz + s * L.stage_stride) + bulk_run(a.noise_vote + (size_t)b * P, P).off : nullptr;
    const float* XS = smem + (L.XS + s * L.stage_stride);
    const float* OS = smem + (L.OS + s * L.stage_stride);
    const float gll = up.g_ll_per_example ? __ldg(up.g_ll_per_example + b) : 0.0f;
    const float greg = up.g_reg_per_example ? __ldg(up.g_reg_per_example + b) : 0.0f;

    // ---- phase 0: capsule-level transform; S partials: sum over an object chunk of posterior * upstream ---------------
    for (int oo = tid; oo < O; oo += T) {
      const float* row = prm + oo * A;
      float t[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) t[p] = row[6 * V + p] + __ldg(a.bias_cvr + oo * 6 + p);
      PoseAffine r;
      pose_affine_fast<kSim>(t, r);
      const float lc = row[6 * V + 6] + __ldg(a.bias_caps + oo) + OS[oo * 4 + 0];
      float* dst = R + oo * 8;
#pragma unroll
      for (int p = 0; p < 6; ++p) dst[p] = r.a[p];
      dst[6] = sigmoid_fast(lc);
      dst[7] = lc;
    }
    {
      const int chunk = (O + 7) >> 3;
      for (int idx = tid; idx < 8 * V; idx += T) {
        const int k = fast_div(idx, inv_V), v = idx - k * V;
        float sacc = 0.0f;
        if (gpost) {
          const int o1 = min(O, (k + 1) * chunk);
          for (int oo = k * chunk; oo < o1; ++oo) sacc = fmaf(post[oo * V + v], gpost[oo * V + v], sacc);
        }
        SPART[idx] = sacc;   // S[v] = sum_k SPART[k][v]: the dummy component has no upstream gradient on this path
        if (k == 0 && out.g_presence) out.g_presence[(size_t)b * V + v] = gll * XS[v * 8 + 7];
      }
    }
    __syncthreads();

    // the next image's small inputs: issue the loads now, park them after the pair loop (latency hidden by phase 2)
    float small_next[kSmallMax];
#pragma unroll
    for (int j = 0; j < kSmallMax; ++j) {
      const int i = tid + j * T;
      small_next[j] = (has_next && i < n_small) ? caps2_bwd_small_load(a, sv, up, b_next, i) : 0.0f;
    }

    // ---- phase 2: per-pair gradients --------------------------------------------------------------------------------------
    const size_t bP = (size_t)b * P, bP1 = (size_t)b * (O + 1) * V;
    const int step_o = fast_div(T, inv_V), step_v = T - step_o * V;   // (o, v) advance of one pass of T pairs
    int oo = fast_div(tid, inv_V), v = tid - oo * V;
    PairConst cur = {};
    if (tid < P) cur = load_pair_const(a, tid);
    for (int p = tid; p < P; p += T) {
      PairConst nxt = cur;
      if (p + T < P) nxt = load_pair_const(a, p + T);
      float* row = prm + oo * A;
      const float* r = R + oo * 8;
      const float st[6] = {cur.s0.x, cur.s0.y, cur.s1.x, cur.s1.y, cur.s2.x, cur.s2.y};
      float raw[6], t[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        raw[c] = row[6 * v + c];
        t[c] = (deform ? raw[c] : 0.0f) + st[c];
      }
      PoseAffine pa;
      pose_affine_fast<kSim>(t, pa);
      float vt[6];
      compose_vote(r, pa.a, vt);
      const float raw_lv = row[6 * V + 7 + v], raw_u = row[7 * V + 7 + v];
      float lv = raw_lv + cur.bv;
      if (nz) lv += nz[p];
      const float pv = sigmoid_fast(lv);
      const float vp = r[6] * pv;
      const float u = raw_u + cur.bs;
      const float sc = learn ? softplus_fast(u + 0.5f) + 1e-2f : 1.0f;
      const float* xv = XS + v * 8;
      float diff[6], q = 0.0f;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        diff[c] = xv[c] - vt[c];
        q = fmaf(diff[c], diff[c], q);
      }
      const float pst = post[p];
      const float h = gpost ? gpost[p] : 0.0f;
      float S = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) S += SPART[k * V + v];
      const float g_pl = pst * (h - S) + gll * xv[6] * pst;
      float g_vp = 0.0f;
      if (up.g_vote_presence) g_vp += __ldg(up.g_vote_presence + bP + p);
      if (__float_as_int(OS[oo * 4 + 2]) == v) g_vp += OS[oo * 4 + 1];
      float g_ml = g_pl;
      if (up.g_mixing_logit) g_ml += __ldg(up.g_mixing_logit + bP1 + p);
      if (!(vp < kLogSafeEps)) g_vp += g_ml * rcp_approx(vp);
      const float inv_sc = rcp_approx(sc);
      const float inv2 = inv_sc * inv_sc;
      const float coef = g_pl * inv2;
      float gv[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        gv[c] = coef * diff[c];
        if (up.g_vote) gv[c] += __ldg(up.g_vote + (bP + p) * 6 + c);
      }
      float g_sc = g_pl * (q * inv2 * inv_sc - 6.0f * inv_sc);
      if (up.g_scale) g_sc += __ldg(up.g_scale + bP + p);
      const float g_u = learn ? g_sc * sigmoid_fast(u + 0.5f) : 0.0f;
      float g_lv = g_vp * r[6] * pv * (1.0f - pv);
      if (up.g_presence_logit_per_vote) g_lv += __ldg(up.g_presence_logit_per_vote + bP + p);
      // vote = R . A: gradient w.r.t. A (-> this pair's cpr parameters) and w.r.t. R (-> summed over the parts below)
      const float* A_ = pa.a;
      float ga[6];
      ga[0] = r[0] * gv[0] + r[3] * gv[3];
      ga[1] = r[0] * gv[1] + r[3] * gv[4];
      ga[2] = r[0] * gv[2] + r[3] * gv[5];
      ga[3] = r[1] * gv[0] + r[4] * gv[3];
      ga[4] = r[1] * gv[1] + r[4] * gv[4];
      ga[5] = r[1] * gv[2] + r[4] * gv[5];
      RED7[0 * P1 + p] = gv[0] * A_[0] + gv[1] * A_[1] + gv[2] * A_[2];
      RED7[1 * P1 + p] = gv[0] * A_[3] + gv[1] * A_[4] + gv[2] * A_[5];
      RED7[2 * P1 + p] = gv[2];
      RED7[3 * P1 + p] = gv[3] * A_[0] + gv[4] * A_[1] + gv[5] * A_[2];
      RED7[4 * P1 + p] = gv[3] * A_[3] + gv[4] * A_[4] + gv[5] * A_[5];
      RED7[5 * P1 + p] = gv[5];
      RED7[6 * P1 + p] = g_vp * pv;
      float gt[6];
      pose_affine_bwd<kSim>(ga, pa, gt);
      // gradient rows in place; the batch sums (-> cpr_static and bias gradients) take the pre-activation gradient
      float* cs = COLSUM + oo * A;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        cs[6 * v + c] += gt[c];
        float g = deform ? fmaf(greg, raw[c], gt[c]) : 0.0f;
        if (relu && !(raw[c] > 0.0f)) g = 0.0f;
        row[6 * v + c] = g;
      }
      cs[6 * V + 7 + v] += g_lv;
      cs[7 * V + 7 + v] += g_u;
      row[6 * V + 7 + v] = (relu && !(raw_lv > 0.0f)) ? 0.0f : g_lv;
      row[7 * V + 7 + v] = (relu && !(raw_u > 0.0f)) ? 0.0f : g_u;
      cur = nxt;
      oo += step_o;
      v += step_v;
      if (v >= V) {
        v -= V;
        ++oo;
      }
    }
    if (L.stages == 2 && has_next) {
#pragma unroll
      for (int j = 0; j < kSmallMax; ++j) {
        const int i = tid + j * T;
        if (i < n_small) {
          float* dst = i < V * 8 ? smem + (L.XS + sn * L.stage_stride) + i : smem + (L.OS + sn * L.stage_stride) + (i - V * 8);
          *dst = small_next[j];
        }
      }
    }
    __syncthreads();

    // ---- phase 3: capsule-level slots -- sum the parts (8 lanes per object), chain through the cvr transform ----------------
    {
      const int n_items = (O * 8 + 31) & ~31;
      for (int idx = tid; idx < n_items; idx += T) {
        const int oo = idx >> 3, c = idx & 7;
        const bool ok = oo < O;
        float sum = 0.0f;
        if (ok && c < 7) {
          const float* src = RED7 + c * P1 + oo * V;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int v = 0;
          for (; v + 4 <= V; v += 4) {
            s0 += src[v];
            s1 += src[v + 1];
            s2 += src[v + 2];
            s3 += src[v + 3];
          }
          for (; v < V; ++v) s0 += src[v];
          sum = (s0 + s1) + (s2 + s3);
        }
        // gather the object's seven sums into its first lane
        const int base_lane = lane & ~7;
        float g7[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) g7[k] = __shfl_sync(0xffffffffu, sum, base_lane + k);
        if (ok && c == 0) {
          float* row = prm + oo * A;
          float* cs = COLSUM + oo * A;
          float raw[7], t[6];
#pragma unroll
          for (int p = 0; p < 7; ++p) raw[p] = row[6 * V + p];
#pragma unroll
          for (int p = 0; p < 6; ++p) t[p] = raw[p] + __ldg(a.bias_cvr + oo * 6 + p);
          PoseAffine rr;
          pose_affine_fast<kSim>(t, rr);
          float gt[6];
          pose_affine_bwd<kSim>(g7, rr, gt);
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            cs[6 * V + p] += gt[p];
            row[6 * V + p] = (relu && !(raw[p] > 0.0f)) ? 0.0f : gt[p];
          }
          const float pc = R[oo * 8 + 6];
          const float g_lc = g7[6] * pc * (1.0f - pc) + OS[oo * 4 + 3];
          cs[6 * V + 6] += g_lc;
          row[6 * V + 6] = (relu && !(raw[6] > 0.0f)) ? 0.0f : g_lc;
        }
      }
    }
    fence_proxy_async();   // the finished gradient block is read by the bulk store
    __syncthreads();

    // ---- phase 4: the image's gradient block leaves in one bulk store (edge floats through registers) -------------------
    {
      float* gdst = out.g_all_param + (size_t)b * O * A;   // congruent to the source modulo 16 bytes (checked on the host)
      if (tid == 0 && r0.body) {
        bulk_s2g(gdst + r0.head, prm + r0.head, 4u * (unsigned)r0.body);
        bulk_commit();
      }
      // warp 0 also refills this buffer (next iteration's prefetch), so program order + __syncwarp keep the edge
      // floats of this image from being overwritten before they are written out
      if (warp == 0) bulk_run_edges_out(gdst, prm_base, r0, lane);
      if (L.stages == 1) {
        // single stage: the buffer is refilled right away, so drain the store first, then fetch the next image
        if (warp == 0) {
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
        }
        __syncthreads();   // edge floats written out, small inputs of this image no longer needed
        if (has_next) {
          if (warp == 0) caps2_bwd_stage_issue(a, sv, up, L, smem, 0, b_next, bar0, lane);
#pragma unroll
          for (int j = 0; j < kSmallMax; ++j) {
            const int i = tid + j * T;
            if (i < n_small) {
              float* dst = i < V * 8 ? smem + L.XS + i : smem + L.OS + (i - V * 8);
              *dst = small_next[j];
            }
          }
        }
      }
    }
  }

  // per-CTA batch sums -> partial row; summed over the CTAs in fixed order afterwards
  __syncthreads();
  {
    float* dst = out.partials + (size_t)blockIdx.x * O * A;
    for (int i = tid; i < O * A; i += T) dst[i] = COLSUM[i];
  }
  if (tid == 0) bulk_wait_read_all();
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)
bool caps_force_v1() {
  const char* e = getenv("SCAE_CAPS_IMPL");
  return e != nullptr && strcmp(e, "v1") == 0;
}

bool caps_force_v2() {
  const char* e = getenv("SCAE_CAPS_IMPL");
  return e != nullptr && strcmp(e, "v2") == 0;
}

// Shared-memory carve-out hint: just enough for `ctas` resident CTAs (each also pays 1 KB of system-reserved shared
// memory), so that what is left of the 256 KB unified array serves as L1 for the batch-shared parameters.
static int carveout_percent(size_t smem_bytes, int ctas) {
  const double want = (double)ctas * ((double)smem_bytes + 1024.0) + 2048.0;
  int pct = (int)(want / (228.0 * 1024.0) * 100.0) + 1;
  return pct > 100 ? 100 : pct;
}

// Threads per CTA for a loop of n pairs: the largest multiple of 32 in [lo, hi] whose thread-slot efficiency
// n / (ceil(n / T) * T) is within 6 % of the best one -- these kernels are latency bound, so warps count for more than
// the last few percent of lane utilisation.
static int pick_threads(int n, int lo, int hi) {
  double best_eff = -1.0;
  for (int t = hi; t >= lo; t -= 32) {
    const double eff = (double)n / ((double)((n + t - 1) / t) * t);
    if (eff > best_eff) best_eff = eff;
  }
  for (int t = hi; t >= lo; t -= 32) {
    const double eff = (double)n / ((double)((n + t - 1) / t) * t);
    if (eff >= 0.94 * best_eff) return t;
  }
  return hi;
}

static bool caps2_shape_ok(const scae_caps_args* a) {
  // 16-byte aligned base pointers (promised by the header) are what makes the cpr_static float2 loads and the bulk
  // copies legal; P < 2^22 keeps fast_div exact
  return (long)a->O * a->V < (1L << 22) && aligned16(a->all_param) && aligned16(a->cpr_static) &&
         (!a->noise_vote || aligned16(a->noise_vote));
}

int caps2_fwd(const scae_caps_args* a, const scae_caps_outputs* out, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!caps2_shape_ok(a)) return SCAE_OK;
  const Caps2FwdLayout L = caps2_fwd_layout(a->O, a->V, a->noise_vote != nullptr);
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > (size_t)max_smem_optin()) return SCAE_OK;
  const int threads = pick_threads(a->O * a->V, 320, kF2MaxThreads);
  const bool sim = (a->flags & SCAE_CAPS_SIMILARITY) != 0;
  auto kern = sim ? caps2_fwd_kernel<true> : caps2_fwd_kernel<false>;
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent(smem, 2)));
  kern<<<a->B, threads, smem, stream>>>(*a, *out, L);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  *handled = true;
  return SCAE_OK;
}

static int caps2_bwd_grid(const scae_caps_args* a) {
  const int sms = sm_count();
  return a->B < sms ? a->B : sms;
}

size_t caps2_bwd_workspace_bytes(const scae_caps_args* a) {
  return (size_t)caps2_bwd_grid(a) * a->O * (8 * a->V + 7) * sizeof(float);
}

int caps2_bwd(const scae_caps_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up, float* g_all_param,
              float* g_shared, float* g_dummy_vote, float* g_x, float* g_presence, void* workspace,
              size_t workspace_bytes, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!caps2_shape_ok(a)) return SCAE_OK;
  // upstream gradients that need the vote-weighted sums over objects, and part-side input gradients, stay on the
  // general path (caps_ll.cu); a training step with vote_type = presence_type = 'enc' produces none of them
  if (up->g_soft_winner || up->g_soft_winner_presence || up->g_winner || up->g_winner_presence ||
      up->g_mixing_log_prob || g_x)
    return SCAE_OK;
  if (!aligned16(g_all_param) || !aligned16(saved->posterior_mixing_prob) ||
      (up->g_posterior_mixing_prob && !aligned16(up->g_posterior_mixing_prob)))
    return SCAE_OK;
  const int O = a->O, V = a->V, A = 8 * V + 7, n = O * A;
  const int threads = pick_threads(O * V, 384, kB2MaxThreads);   // e.g. 640 for 1280 pairs: two full passes
  if (V * 8 + O * 4 > kSmallMax * threads) return SCAE_OK;
  const bool noise = a->noise_vote != nullptr, have_gpost = up->g_posterior_mixing_prob != nullptr;
  Caps2BwdLayout L = caps2_bwd_layout(O, V, noise, have_gpost, 2);
  if ((size_t)L.total * sizeof(float) > (size_t)max_smem_optin()) L = caps2_bwd_layout(O, V, noise, have_gpost, 1);
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > (size_t)max_smem_optin()) return SCAE_OK;
  const int grid = caps2_bwd_grid(a);
  if (workspace_bytes < (size_t)grid * n * sizeof(float)) return SCAE_OK;
  const bool sim = (a->flags & SCAE_CAPS_SIMILARITY) != 0;
  auto kern = sim ? caps2_bwd_kernel<true> : caps2_bwd_kernel<false>;
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent(smem, 1)));
  float* partials = static_cast<float*>(workspace);
  Caps2BwdOut out{g_all_param, g_presence, partials};
  kern<<<grid, threads, smem, stream>>>(*a, *saved, *up, out, L);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  int rc = launch_reduce_rows(partials, g_shared, grid, n, stream);
  if (rc != SCAE_OK) return rc;
  if (g_dummy_vote) SCAE_CUDA_TRY(cudaMemsetAsync(g_dummy_vote, 0, (size_t)V * 6 * sizeof(float), stream));
  *handled = true;
  return SCAE_OK;
}

}  // namespace scae
