// Hot path 2 (sm_100a), persistent warp-specialised kernels: the capsule part-pose mixture likelihood at the HBM rate.
//
// Same math as caps_ll.cu / caps_ll2.cu (reference object_decoder.py:160-236, :257-372, :413-415; oracle/
// capsule_likelihood.py, oracle/manual_backward.py::capsule_forward_backward).  The first fast path (caps_ll2.cu) spent
// 650-720 lane-instructions per (object, part) pair and stopped at 50 % / 28 % of the HBM roofline with its SMs
// issue- and barrier-bound (profiles/r01_final_kernels.md).  This generation is built around the instruction count:
//
//   * CTAs are PERSISTENT and every consumer thread keeps the SAME (object group k, part v) for every image it sees:
//     thread t = k V + v handles the pairs (o, v), o = k, k + G, k + 2G, ... (NP of them), which are the flat pair
//     indices t, t + T, t + 2T, ... -- every (B,O,V) tensor is still read and written unit-stride.  Everything that
//     does not depend on the image lives in REGISTERS for the whole kernel: the pair's cpr_static row and its two
//     biases (no global re-reads per image, no address arithmetic), the row offsets, the lane masks of the reductions.
//     The part's pose x[v] is loaded once per image instead of once per pair.
//   * a dedicated PRODUCER WARP feeds a ring of S stages with cp.async.bulk (TMA, one mbarrier per stage), and while it
//     waits it does the per-OBJECT work of the next image (capsule transform + presence, formerly "phase 0" with its
//     own CTA barrier) so that the consumers find it ready;
//   * the reductions over objects (logsumexp, arg-max winner, soft winner) are register accumulators over the thread's
//     own NP pairs plus ONE exchange of 12 partial values per thread through shared memory; the reduction over parts
//     (capsule presence = max_v, with its arg-max) is two REDUX instructions over the lanes that share an object;
//   * one named barrier per image among the consumers, none shared with the producer; the stage is handed back through
//     an "empty" mbarrier that doubles as the split barrier protecting the exchange tile;
//   * the votes are written IN PLACE over the deformation parameters they were computed from and leave through a
//     coalesced copy; the posterior weight and the mixing logit of a pair wait in its own two parameter slots.
//
// MUFU forms for every elementary function (sigmoid, tanh, softplus, sin/cos after an exact range reduction, exp, log):
// about 200 lane-instructions per pair forward.
#include <stdlib.h>
#include <string.h>

#include "caps_common.cuh"

namespace scae {

constexpr int kC3MaxStages = 4;
constexpr int kC3Items = 12;   // partial values per thread: sum E, sum E vp, sum vp, best logit / object / vp, 6 x sum E vote
constexpr unsigned kC3BarConsumers = 1;   // named barrier id
constexpr int kNoWinner = 0x7fffffff;

__host__ __device__ inline int c3_round4(int n) { return (n + 3) & ~3; }

struct Caps3FwdLayout {
  int S;                        // stages of the input ring
  int G, NP, T, Tpad;           // object groups, pairs per thread, consumer threads G*V, padded to whole warps
  int slots;                    // capsule-presence partials per object (warps an object's V lanes can span)
  int stage0, stage_stride;     // in floats
  int prm, nz, R, XS;           // offsets inside a stage
  int PV, CP, REGP, total;      // CTA-wide tiles
};

static Caps3FwdLayout caps3_fwd_layout(int O, int V, bool noise, int G, int NP, int S) {
  const int A = 8 * V + 7, P = O * V;
  Caps3FwdLayout L;
  L.S = S, L.G = G, L.NP = NP, L.T = G * V, L.Tpad = (L.T + 31) & ~31;
  L.slots = (V + 30) / 32 + 1;
  int at = 32;                                   // [0, 32): 3 S mbarriers (full, ready, empty)
  auto take = [&](int n) {
    const int here = at;
    at += c3_round4(n);
    return here;
  };
  L.stage0 = at;
  L.prm = take(O * A + 4) - L.stage0;
  L.nz = take(noise ? P + 4 : 0) - L.stage0;
  L.R = take(O * 8) - L.stage0;
  L.XS = take(V * 8) - L.stage0;
  L.stage_stride = at - L.stage0;
  at = L.stage0 + S * L.stage_stride;
  L.PV = take(kC3Items * L.T);
  L.CP = take(2 * O * L.slots * 2);
  L.REGP = take(2 * 32);
  L.total = at;
  return L;
}

__device__ __forceinline__ unsigned c3_full(unsigned bar0, int s) { return bar0 + 8u * (unsigned)s; }
__device__ __forceinline__ unsigned c3_ready(unsigned bar0, int s) { return bar0 + 8u * (unsigned)(kC3MaxStages + s); }
__device__ __forceinline__ unsigned c3_empty(unsigned bar0, int s) { return bar0 + 8u * (unsigned)(2 * kC3MaxStages + s); }

// ---- producer warp ---------------------------------------------------------------------------------------------------
// stage s <- image b: all_param block and noise rows by bulk copy (edge floats through registers), the part poses as a
// [7][V] tile (x0..x5, presence)
__device__ __forceinline__ void caps3_issue(const scae_caps_args& a, const Caps3FwdLayout& L, float* smem, unsigned bar0,
                                            int s, int b, int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V;
  float* st = smem + L.stage0 + s * L.stage_stride;
  const float* gprm = a.all_param + (size_t)b * O * A;
  const float* gnz = a.noise_vote ? a.noise_vote + (size_t)b * P : nullptr;
  const BulkRun rp = bulk_run(gprm, O * A);
  BulkRun rn = {0, 0, 0, 0};
  if (gnz) rn = bulk_run(gnz, P);
  if (lane == 0) {
    mbar_expect_tx(c3_full(bar0, s), 4u * (unsigned)(rp.body + rn.body));
    if (rp.body) bulk_g2s(st + L.prm + rp.off + rp.head, gprm + rp.head, 4u * (unsigned)rp.body, c3_full(bar0, s));
    if (rn.body) bulk_g2s(st + L.nz + rn.off + rn.head, gnz + rn.head, 4u * (unsigned)rn.body, c3_full(bar0, s));
  }
  if (lane < 8) bulk_run_edges_in(st + L.prm, gprm, rp, lane);
  else if (gnz && lane < 16) bulk_run_edges_in(st + L.nz, gnz, rn, lane - 8);
  float* XS = st + L.XS;
  const float* gx = a.x + (size_t)b * V * 6;
  for (int idx = lane; idx < V * 6; idx += 32) {
    const int v = idx / 6, c = idx - 6 * v;
    XS[c * V + v] = __ldg(gx + idx);
  }
  for (int v = lane; v < V; v += 32) XS[6 * V + v] = a.presence ? __ldg(a.presence + (size_t)b * V + v) : 1.0f;
}

template <bool kSim>
__device__ __forceinline__ void caps3_fwd_producer(const scae_caps_args& a, const scae_caps_outputs& o,
                                                   const Caps3FwdLayout& L, float* smem, unsigned bar0, int n_mine,
                                                   int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7;
  const int S = L.S;
  int issued = 0;   // images whose loads have been issued; image j may be issued once image j - S has been released
  auto issue_next = [&]() {
    if (issued >= S) mbar_wait(c3_empty(bar0, issued % S), (unsigned)(((issued - S) / S) & 1));
    caps3_issue(a, L, smem, bar0, issued % S, blockIdx.x + issued * gridDim.x, lane);
    ++issued;
  };
  while (issued < n_mine && issued < S) issue_next();
  for (int i = 0; i < n_mine; ++i) {
    const int s = i % S, b = blockIdx.x + i * gridDim.x;
    const unsigned parity = (unsigned)((i / S) & 1);
    float* st = smem + L.stage0 + s * L.stage_stride;
    while (issued <= i) issue_next();   // (single-stage ring only)
    mbar_wait(c3_full(bar0, s), parity);
    __syncwarp();   // the edge floats were stored by other lanes of this warp
    // per-object work of image i: R[o] = {capsule -> viewer affine (6), capsule presence, its logit}
    const float* prm = st + L.prm + bulk_run(a.all_param + (size_t)b * O * A, O * A).off;
    float* R = st + L.R;
    for (int oo = lane; oo < O; oo += 32) {
      const float* row = prm + oo * A + 6 * V;
      float t[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) t[p] = row[p] + __ldg(a.bias_cvr + oo * 6 + p);
      PoseAffine r;
      pose_affine_mufu<kSim>(t, r);
      float lc = row[6] + __ldg(a.bias_caps + oo);
      if (a.noise_caps) lc += __ldg(a.noise_caps + (size_t)b * O + oo);
      float4* dst = reinterpret_cast<float4*>(R + oo * 8);
      dst[0] = make_float4(r.a[0], r.a[1], r.a[2], r.a[3]);
      dst[1] = make_float4(r.a[4], r.a[5], sigmoid_fast(lc), lc);
      if (o.presence_logit_per_caps) o.presence_logit_per_caps[(size_t)b * O + oo] = lc;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(c3_ready(bar0, s));
    // refill the stage image i-1 used, once its consumers have let go of it (they are on image i now, which is ready)
    if (issued < n_mine && issued - S <= i - 1) issue_next();
  }
}

// the batch-shared parameters of one pair, kept in registers for the whole kernel
struct C3Const {
  float s[6];   // cpr_static[o][v][0..5]
  float bv;     // bias_vote[o][v]
  float bs;     // bias_scale[o][v] + 0.5 (object_decoder.py:224)
};

// NP pairs per thread; kMaxT / kMinB: launch bounds (consumer + producer threads, resident CTAs per SM)
template <bool kSim, int NP, int kMaxT, int kMinB>
__global__ void __launch_bounds__(kMaxT, kMinB) caps3_fwd_kernel(const scae_caps_args a, const scae_caps_outputs o,
                                                                 const Caps3FwdLayout L) {
  SCAE_DYNAMIC_SMEM(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V;
  const int S = L.S, G = L.G, T = L.T, Tpad = L.Tpad;
  const unsigned bar0 = smem_u32(smem);
  const int n_mine = ((int)blockIdx.x < a.B) ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(c3_full(bar0, s), 1);
      mbar_init(c3_ready(bar0, s), 1);
      mbar_init(c3_empty(bar0, s), (unsigned)Tpad);
    }
    fence_mbar_init();
  }
  __syncthreads();   // the only CTA-wide barrier: producer and consumers part ways here

  if (tid >= Tpad) {
    caps3_fwd_producer<kSim>(a, o, L, smem, bar0, n_mine, lane);
    return;
  }

  // ---- consumer set-up: everything that does not depend on the image ------------------------------------------------
  const bool active = tid < T;
  const float inv_V = 1.0f / (float)V;
  const int k = active ? fast_div(tid, inv_V) : -1;
  const int v = active ? tid - k * V : 0;
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;
  C3Const cst[NP];
  bool valid[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int oj = k + G * j;
    valid[j] = active && oj < O;
    if (valid[j]) {
      const int p = oj * V + v;
      const float2* sp = reinterpret_cast<const float2*>(a.cpr_static + (size_t)p * 6);
      const float2 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
      cst[j].s[0] = s0.x, cst[j].s[1] = s0.y, cst[j].s[2] = s1.x, cst[j].s[3] = s1.y, cst[j].s[4] = s2.x, cst[j].s[5] = s2.y;
      cst[j].bv = __ldg(a.bias_vote + p);
      cst[j].bs = __ldg(a.bias_scale + p) + 0.5f;
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) cst[j].s[c] = 0.0f;
      cst[j].bv = cst[j].bs = 0.0f;
    }
  }
  // lanes of this warp that work on the same objects (same k), their first lane, and which of the object's partial
  // slots this warp fills (an object's V lanes span up to `slots` warps)
  const unsigned segmask = __match_any_sync(0xffffffffu, (unsigned)k);
  const bool seglead = lane == __ffs((int)segmask) - 1;
  const int slot = active ? warp - ((k * V) >> 5) : 0;
  const int n_slots = active ? ((k * V + V - 1) >> 5) - ((k * V) >> 5) + 1 : 0;
  const int n_warps = Tpad >> 5;
  // offsets of the thread's first pair inside an image's parameter block, and the stride to its next pair
  const int rel_dyn = active ? k * A + 6 * v : 0, rel_lv = active ? k * A + 6 * V + 7 + v : 0, stepA = G * A;
  const int cp_rel = active ? (k * L.slots + slot) * 2 : 0, cp_step = G * L.slots * 2;
  const float e_dummy = expf(kDummyLog + kDummyLog);   // dummy mixing logit + dummy log-density (object_decoder.py:273-292)
  const float p_dummy = expf(kDummyLog);
  float* PV = smem + L.PV;

  for (int i = 0; i < n_mine; ++i) {
    const int s = i % S, b = blockIdx.x + i * gridDim.x, par = i & 1;
    const unsigned parity = (unsigned)((i / S) & 1);
    float* st = smem + L.stage0 + s * L.stage_stride;
    float* prm = st + L.prm + bulk_run(a.all_param + (size_t)b * O * A, O * A).off;   // inputs, then results in place
    const float* nz = a.noise_vote ? st + L.nz + bulk_run(a.noise_vote + (size_t)b * P, P).off : nullptr;
    const float4* R4 = reinterpret_cast<const float4*>(st + L.R);
    const float* XS = st + L.XS;
    unsigned* CP = reinterpret_cast<unsigned*>(smem + L.CP) + par * O * L.slots * 2;
    float* REGP = smem + L.REGP + par * 32;
    const size_t bP = (size_t)b * P, bP1 = (size_t)b * (O + 1) * V;

    mbar_wait(c3_full(bar0, s), parity);    // bulk-copied bytes visible to this thread
    mbar_wait(c3_ready(bar0, s), parity);   // edge floats, part poses and the per-object tile written by the producer

    // ---- phase A: the thread's pairs -------------------------------------------------------------------------------
    float xv[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) xv[c] = XS[c * V + v];
    float regsum = 0.0f, sE = 0.0f, sEvp = 0.0f, svp = 0.0f, best = -INFINITY, bvp = 0.0f;
    int bo = kNoWinner;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* dyn0 = prm + rel_dyn;   // the thread's first pair: its 6 deformation parameters, its vote logit slot
    float* lvs0 = prm + rel_lv;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float vp = 0.0f;
      if (valid[j]) {
        const int oj = k + G * j, p = tid + j * T;
        float* dyn = dyn0 + j * stepA;    // row[6 v + c]
        float* lvs = lvs0 + j * stepA;    // row[6 V + 7 + v]; the scale slot sits V floats further
        float t[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float d = deform ? dyn[c] : 0.0f;
          regsum = fmaf(d, d, regsum);
          t[c] = d + cst[j].s[c];
        }
        PoseAffine pa;
        pose_affine_mufu<kSim>(t, pa);
        const float4 r0 = R4[oj * 2], r1 = R4[oj * 2 + 1];
        const float r[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
        float vt[6];
        compose_vote(r, pa.a, vt);
#pragma unroll
        for (int c = 0; c < 6; ++c) dyn[c] = vt[c];   // in place: leaves with the coalesced copy of phase C
        float lv = lvs[0] + cst[j].bv;
        if (nz) lv += nz[p];
        vp = r1.z * sigmoid_fast(lv);
        float sc = 1.0f, lp;
        float q = 0.0f;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float d = xv[c] - vt[c];
          q = fmaf(d, d, q);
        }
        // sum over the 6 pose dims of Normal(vote, sc).log_prob(x)
        if (learn) {
          sc = softplus_fast(lvs[V] + cst[j].bs) + 1e-2f;
          const float inv = rcp_approx(sc);
          lp = fmaf(-0.5f * q, inv * inv, fmaf(-6.0f * kLn2F, lg2_approx(sc), -6.0f * kHalfLog2Pi));
        } else {
          lp = fmaf(-0.5f, q, -6.0f * kHalfLog2Pi);
        }
        const float ml = vp < kLogSafeEps ? kLogSafeFloor : lg2_approx(vp) * kLn2F;   // log_safe (math_ops.py:18-21)
        const float pl = ml + lp;
        const float E = ex2_approx(pl * kLog2eF);
        lvs[0] = E;    // the pair's own two slots carry its posterior weight and mixing logit to phase C
        lvs[V] = ml;
        sE += E;
        sEvp = fmaf(E, vp, sEvp);
        svp += vp;
        if (pl > best) {           // objects ascend with j: lowest o on ties (torch.argmax)
          best = pl;
          bo = oj;
          bvp = vp;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] = fmaf(E, vt[c], acc[c]);
        if (o.scale) o.scale[bP + p] = sc;
        if (o.vote_presence) o.vote_presence[bP + p] = vp;
        if (o.presence_logit_per_vote) o.presence_logit_per_vote[bP + p] = lv;
        if (o.vote_presence_binary) o.vote_presence_binary[bP + p] = ml > kDummyLog ? 1.0f : 0.0f;
        if (o.mixing_logit) o.mixing_logit[bP1 + p] = ml;
      }
      // capsule presence = max over parts (object_decoder.py:415); lowest part index wins ties.  vp >= 0, so the
      // unsigned order of its bits is the order of the floats.  Every lane of the warp takes part (lanes without a pair
      // in this pass share their mask with lanes that have none either).
      {
        const unsigned bits = __float_as_uint(vp);
        const unsigned m = redux_max_u32(segmask, bits);
        const unsigned arg = redux_min_u32(segmask, bits == m ? (unsigned)v : 0xffffffffu);
        if (seglead && valid[j]) {
          CP[cp_rel + j * cp_step] = m;
          CP[cp_rel + j * cp_step + 1] = arg;
        }
      }
    }
    // the exchange tile is single-buffered: the previous image's readers must be done (they arrived on its `empty`)
    if (i > 0) mbar_wait(c3_empty(bar0, (i - 1) % S), (unsigned)(((i - 1) / S) & 1));
    if (active) {
      float* pv = PV + tid;
      pv[0 * T] = sE;
      pv[1 * T] = sEvp;
      pv[2 * T] = svp;
      pv[3 * T] = best;
      pv[4 * T] = __int_as_float(bo);
      pv[5 * T] = bvp;
#pragma unroll
      for (int c = 0; c < 6; ++c) pv[(6 + c) * T] = acc[c];
    }
    regsum = warp_sum(regsum);
    if (lane == 0) REGP[warp] = regsum;
    named_bar_sync(kC3BarConsumers, (unsigned)Tpad);

    // ---- phase C ---------------------------------------------------------------------------------------------------------
    if (active) {
      // every thread of part v: the normalisers of its column (same order in every thread)
      float Ssum = e_dummy, Svp = p_dummy;
      for (int kk = 0; kk < G; ++kk) {
        Ssum += PV[0 * T + kk * V + v];
        Svp += PV[2 * T + kk * V + v];
      }
      const float invS = __frcp_rn(Ssum);
      const float mlse = logf(Svp);   // logsumexp_o of the mixing logits: exp(log_safe(vp)) = vp
      // one job per object group: k = 0 the hard winner, k = 1..6 a pose dimension of the soft winner (wrapping when G < 7)
      for (int job = k; job < 7; job += G) {
        if (job == 0) {
          float bb = -INFINITY, bbvp = 0.0f;
          int bbo = kNoWinner;
          for (int kk = 0; kk < G; ++kk) {
            const float cb = PV[3 * T + kk * V + v];
            const int co = __float_as_int(PV[4 * T + kk * V + v]);
            if (cb > bb || (cb == bb && co < bbo)) {
              bb = cb;
              bbo = co;
              bbvp = PV[5 * T + kk * V + v];
            }
          }
          if (bbo == kNoWinner) {   // every logit is NaN: the first object, like the other paths
            bbo = 0;
            bbvp = 0.0f;
          }
          const size_t bv = (size_t)b * V + v;
          if (o.winner) {
            const float* wr = prm + bbo * A + 6 * v;
#pragma unroll
            for (int c = 0; c < 6; ++c) o.winner[bv * 6 + c] = wr[c];
          }
          if (o.winner_presence) o.winner_presence[bv] = bbvp;
          if (o.winner_idx) o.winner_idx[bv] = bbo;
          if (o.is_from_capsule) o.is_from_capsule[bv] = fast_div(bbo, inv_V);   // sic (object_decoder.py:334)
        } else if (o.soft_winner) {
          const int c = job - 1;
          float sw = e_dummy * __ldg(a.dummy_vote + v * 6 + c);
          for (int kk = 0; kk < G; ++kk) sw += PV[(6 + c) * T + kk * V + v];
          o.soft_winner[((size_t)b * V + v) * 6 + c] = sw * invS;
        }
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (!valid[j]) continue;
        const int oj = k + G * j, p = tid + j * T;
        const float* lvs = prm + rel_lv + j * stepA;
        const float* row = prm + oj * A;
        if (o.posterior_mixing_prob) o.posterior_mixing_prob[bP + p] = lvs[0] * invS;
        if (o.mixing_log_prob) o.mixing_log_prob[bP1 + p] = lvs[V] - mlse;
        if (o.vote) {
          float* gv = o.vote + (bP + (size_t)oj * V) * 6;
#pragma unroll
          for (int m = 0; m < 6; ++m) gv[v + m * V] = row[v + m * V];
        }
        if (v == 0) {   // capsule presence of object oj: combine the warps' partials, parts ascending
          unsigned m = CP[(oj * L.slots) * 2], arg = CP[(oj * L.slots) * 2 + 1];
          for (int q = 1; q < n_slots; ++q) {
            const unsigned cm = CP[(oj * L.slots + q) * 2];
            if (cm > m) {
              m = cm;
              arg = CP[(oj * L.slots + q) * 2 + 1];
            }
          }
          if (arg == 0xffffffffu) arg = 0;
          if (o.caps_presence) o.caps_presence[(size_t)b * O + oj] = __uint_as_float(m);
          if (o.caps_presence_arg) o.caps_presence_arg[(size_t)b * O + oj] = (int)arg;
        }
      }
    }
    if (warp == 0) {   // per-part and per-example results
      float part = 0.0f;
      for (int vv = lane; vv < V; vv += 32) {
        float Ssum = e_dummy, Svp = p_dummy, SEvp = 0.0f;
        for (int kk = 0; kk < G; ++kk) {
          Ssum += PV[0 * T + kk * V + vv];
          SEvp += PV[1 * T + kk * V + vv];
          Svp += PV[2 * T + kk * V + vv];
        }
        const float lse = logf(Ssum);
        const size_t bv = (size_t)b * V + vv;
        if (o.log_prob_per_point) o.log_prob_per_point[bv] = lse;
        if (o.soft_winner_presence) o.soft_winner_presence[bv] = SEvp * __frcp_rn(Ssum);   // the dummy has presence 0
        const size_t dummy_row = bP1 + (size_t)O * V + vv;
        if (o.mixing_logit) o.mixing_logit[dummy_row] = kDummyLog;
        if (o.mixing_log_prob) o.mixing_log_prob[dummy_row] = kDummyLog - logf(Svp);
        part = fmaf(lse, XS[6 * V + vv], part);
      }
      part = warp_sum(part);
      const float reg = warp_sum(lane < n_warps ? REGP[lane] : 0.0f);
      if (lane == 0) {
        if (o.ll_per_example) o.ll_per_example[b] = part;
        if (o.reg_per_example) o.reg_per_example[b] = 0.5f * reg;
      }
    }
    fence_proxy_async();   // this thread's in-place writes are ordered before the bulk copy that refills the stage
    mbar_arrive(c3_empty(bar0, s));
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)

struct Caps3Plan {
  int NP, max_threads, min_blocks;   // the compiled variant
  int threads;                       // consumers (padded to whole warps) + one producer warp
  Caps3FwdLayout L;
  size_t smem;
};

static int c3_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Work split: G object groups x V parts consumer threads, NP = ceil(O / G) pairs each.  Candidates are the compiled
// variants {NP, launch bound, CTAs per SM}; the best lane utilisation O / (G NP) wins, then the most resident consumer
// threads per SM, then the variant listed first.
static bool caps3_plan_fwd(const scae_caps_args* a, Caps3Plan* plan) {
  const int O = a->O, V = a->V;
  const bool noise = a->noise_vote != nullptr;
  const int budget = max_smem_optin();
  struct Cand {
    int NP, max_threads, min_blocks;
  };
  const Cand cands[] = {{2, 672, 1}, {4, 352, 2}, {1, 448, 2}, {1, 672, 1}, {4, 544, 1}};
  const int force_np = c3_env_int("SCAE_CAPS3_NP", 0), force_mb = c3_env_int("SCAE_CAPS3_MINB", 0);
  const int force_s = c3_env_int("SCAE_CAPS3_STAGES", 0);
  double best_eff = 0.0;
  int best_resident = 0;
  bool found = false;
  for (const Cand& c : cands) {
    if (force_np && c.NP != force_np) continue;
    if (force_mb && c.min_blocks != force_mb) continue;
    const int G = (O + c.NP - 1) / c.NP;
    const int T = G * V, threads = ((T + 31) & ~31) + 32;
    if (threads > c.max_threads) continue;
    const double eff = (double)O / ((double)G * c.NP);
    // stages: as many as fit (at most kC3MaxStages); a two-CTA variant must leave room for its twin
    const int room = c.min_blocks == 2 ? (228 * 1024) / 2 - 1024 : budget;
    int S = force_s >= 1 && force_s <= kC3MaxStages ? force_s : kC3MaxStages;
    Caps3FwdLayout L = caps3_fwd_layout(O, V, noise, G, c.NP, S);
    while (S > 1 && (size_t)L.total * sizeof(float) > (size_t)room) L = caps3_fwd_layout(O, V, noise, G, c.NP, --S);
    if ((size_t)L.total * sizeof(float) > (size_t)room) continue;
    const size_t smem = (size_t)L.total * sizeof(float);
    const int regs = 65536 / (c.max_threads * c.min_blocks) & ~7;           // per-thread cap of the variant
    int ctas = 65536 / (regs * threads);
    if (ctas > 2048 / threads) ctas = 2048 / threads;
    if (ctas > (int)((228u * 1024u) / (smem + 1024u))) ctas = (int)((228u * 1024u) / (smem + 1024u));
    if (ctas < 1) ctas = 1;
    const int resident = ctas * T;
    if (found && (eff < best_eff - 1e-9 || (eff < best_eff + 1e-9 && resident <= best_resident))) continue;
    plan->NP = c.NP, plan->max_threads = c.max_threads, plan->min_blocks = c.min_blocks, plan->threads = threads;
    plan->L = L, plan->smem = smem;
    best_eff = eff, best_resident = resident;
    found = true;
  }
  return found;
}

static bool caps3_shape_ok(const scae_caps_args* a) {
  return (long)a->O * a->V < (1L << 22) && aligned16(a->all_param) && aligned16(a->cpr_static) &&
         (!a->noise_vote || aligned16(a->noise_vote));
}

template <bool kSim>
static int caps3_fwd_launch(const scae_caps_args* a, const scae_caps_outputs* out, const Caps3Plan& plan,
                            cudaStream_t stream) {
  void (*kern)(const scae_caps_args, const scae_caps_outputs, const Caps3FwdLayout) = nullptr;
  if (plan.NP == 1 && plan.min_blocks == 2) kern = caps3_fwd_kernel<kSim, 1, 448, 2>;
  else if (plan.NP == 1) kern = caps3_fwd_kernel<kSim, 1, 672, 1>;
  else if (plan.NP == 2) kern = caps3_fwd_kernel<kSim, 2, 672, 1>;
  else if (plan.min_blocks == 2) kern = caps3_fwd_kernel<kSim, 4, 352, 2>;
  else kern = caps3_fwd_kernel<kSim, 4, 544, 1>;
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  // persistent grid: as many CTAs as the device holds at once (registers, shared memory, threads)
  int per_sm = 0;
  SCAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.threads, plan.smem));
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > a->B) grid = a->B;
  kern<<<grid, plan.threads, plan.smem, stream>>>(*a, *out, plan.L);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

int caps3_fwd(const scae_caps_args* a, const scae_caps_outputs* out, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!caps3_shape_ok(a)) return SCAE_OK;
  Caps3Plan plan;
  if (!caps3_plan_fwd(a, &plan)) return SCAE_OK;
  const bool sim = (a->flags & SCAE_CAPS_SIMILARITY) != 0;
  const int rc = sim ? caps3_fwd_launch<true>(a, out, plan, stream) : caps3_fwd_launch<false>(a, out, plan, stream);
  if (rc != SCAE_OK) return rc;
  *handled = true;
  return SCAE_OK;
}

}  // namespace scae
