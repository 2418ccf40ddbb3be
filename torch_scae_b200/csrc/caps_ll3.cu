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
//   * every per-image input (all_param block, noise rows, part poses, presences) arrives by cp.async.bulk (TMA) in a
//     ring of S stages, one "full" mbarrier per stage; the LAST WARP issues the copies for the stage that was just
//     released and does the per-OBJECT work of the NEXT image (capsule transform + presence, formerly "phase 0" with
//     its own CTA barrier) before the current image's barrier, so that nobody waits for either.  (A dedicated producer
//     warp was measured first: with 21 warps one SM sub-partition holds 6 of them and the register cap drops from 96 to
//     80 per thread, which spilled; 20 warps keep 96.);
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
  int prm, nz, xs, ps, nc, R;   // offsets inside a stage: all_param block, noise_vote, x, presence, noise_caps, object tile
  int PV, BIAS, CP, REGP, INVS, MLSE, LLV, total;   // CTA-wide tiles (CP .. LLV double-buffered by image parity)
};

static Caps3FwdLayout caps3_fwd_layout(const scae_caps_args* a, int G, int NP, int S) {
  const int O = a->O, V = a->V, A = 8 * V + 7, P = O * V;
  Caps3FwdLayout L;
  L.S = S, L.G = G, L.NP = NP, L.T = G * V, L.Tpad = (L.T + 31) & ~31;
  L.slots = (V + 30) / 32 + 1;
  int at = 32;                                   // [0, 32): 2 S mbarriers (full, empty)
  auto take = [&](int n) {
    const int here = at;
    at += c3_round4(n);
    return here;
  };
  L.stage0 = at;
  L.prm = take(O * A + 4) - L.stage0;
  L.nz = take(a->noise_vote ? P + 4 : 0) - L.stage0;
  L.xs = take(V * 6 + 4) - L.stage0;
  L.ps = take(a->presence ? V + 4 : 0) - L.stage0;
  L.nc = take(a->noise_caps ? O + 4 : 0) - L.stage0;
  L.R = take(O * 8) - L.stage0;
  L.stage_stride = at - L.stage0;
  at = L.stage0 + S * L.stage_stride;
  L.PV = take(kC3Items * L.T);
  L.BIAS = take(O * 8);
  L.CP = take(2 * O * L.slots * 2);
  L.REGP = take(2 * 32);
  L.INVS = take(2 * V);
  L.MLSE = take(2 * V);
  L.LLV = take(2 * V);
  L.total = at;
  return L;
}

__device__ __forceinline__ unsigned c3_full(unsigned bar0, int s) { return bar0 + 8u * (unsigned)s; }
__device__ __forceinline__ unsigned c3_empty(unsigned bar0, int s) { return bar0 + 8u * (unsigned)(kC3MaxStages + s); }

// ---- staging (one warp) ---------------------------------------------------------------------------------------------------
// One contiguous per-image input: n floats at g -> base[off ..] (bulk_run: interior by bulk copy, <= 3 + 3 edge floats
// through registers).
struct C3Run {
  const float* g;
  float* base;
  int n;
};

// stage s <- image b: everything the image contributes (all_param block, noise rows, part poses and presences).  Called by
// all lanes of one warp; lanes 0..4 issue one run each, lane groups 0..2 load the edge floats.
__device__ __forceinline__ void caps3_issue(const scae_caps_args& a, const Caps3FwdLayout& L, float* smem, unsigned bar0,
                                            int s, int b, int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V;
  float* st = smem + L.stage0 + s * L.stage_stride;
  C3Run run = {nullptr, nullptr, 0};
  const int r = lane & 7;
  if (r == 0) run = {a.all_param + (size_t)b * O * A, st + L.prm, O * A};
  else if (r == 1 && a.noise_vote) run = {a.noise_vote + (size_t)b * P, st + L.nz, P};
  else if (r == 2) run = {a.x + (size_t)b * V * 6, st + L.xs, V * 6};
  else if (r == 3 && a.presence) run = {a.presence + (size_t)b * V, st + L.ps, V};
  else if (r == 4 && a.noise_caps) run = {a.noise_caps + (size_t)b * O, st + L.nc, O};
  BulkRun br = {0, 0, 0, 0};
  if (run.n) br = bulk_run(run.g, run.n);
  unsigned bytes = lane < 5 ? 4u * (unsigned)br.body : 0u;   // total interior bytes of the five runs
#pragma unroll
  for (int d = 4; d > 0; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
  if (lane == 0) mbar_expect_tx(c3_full(bar0, s), bytes);
  __syncwarp();
  if (lane < 5 && br.body)
    bulk_g2s(run.base + br.off + br.head, run.g + br.head, 4u * (unsigned)br.body, c3_full(bar0, s));
  if (run.n) {   // lane group q takes head float q and tail float q of every run
    const int q = lane >> 3;
    if (q < 3) {
      if (q < br.head) run.base[br.off + q] = __ldg(run.g + q);
      if (q < br.tail) run.base[br.off + br.head + br.body + q] = __ldg(run.g + br.head + br.body + q);
    }
  }
}

// per-object work of image b in stage s (one warp, lane = object): R[o] = {capsule -> viewer affine (6), capsule
// presence, its logit}; BIAS[o] = {bias_cvr[6], bias_caps, -}
template <bool kSim>
__device__ __forceinline__ void caps3_object_tile(const scae_caps_args& a, const scae_caps_outputs& o,
                                                  const Caps3FwdLayout& L, float* smem, int s, int b, int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7;
  float* st = smem + L.stage0 + s * L.stage_stride;
  const float* prm = st + L.prm + bulk_run(a.all_param + (size_t)b * O * A, O * A).off;
  const float* nc = a.noise_caps ? st + L.nc + bulk_run(a.noise_caps + (size_t)b * O, O).off : nullptr;
  const float* BIAS = smem + L.BIAS;
  float* R = st + L.R;
  for (int oo = lane; oo < O; oo += 32) {
    const float* row = prm + oo * A + 6 * V;
    const float4 b0 = *reinterpret_cast<const float4*>(BIAS + oo * 8), b1 = *reinterpret_cast<const float4*>(BIAS + oo * 8 + 4);
    float t[6] = {row[0] + b0.x, row[1] + b0.y, row[2] + b0.z, row[3] + b0.w, row[4] + b1.x, row[5] + b1.y};
    PoseAffine r;
    pose_affine_mufu<kSim>(t, r);
    float lc = row[6] + b1.z;
    if (nc) lc += nc[oo];
    float4* dst = reinterpret_cast<float4*>(R + oo * 8);
    dst[0] = make_float4(r.a[0], r.a[1], r.a[2], r.a[3]);
    dst[1] = make_float4(r.a[4], r.a[5], sigmoid_fast(lc), lc);
    if (o.presence_logit_per_caps) o.presence_logit_per_caps[(size_t)b * O + oo] = lc;
  }
}

// the batch-shared parameters of one pair, kept in registers for the whole kernel
struct C3Const {
  float s[6];   // cpr_static[o][v][0..5]
  float bv;     // bias_vote[o][v]
  float bs;     // bias_scale[o][v] + 0.5 (object_decoder.py:224)
};

// NP pairs per thread; kMaxT / kMinB: launch bounds (threads, resident CTAs per SM); kFull: every
// output tensor is requested (what the Python binding does), so no pointer is tested in the loops
template <bool kSim, int NP, int kMaxT, int kMinB, bool kFull>
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
      mbar_init(c3_empty(bar0, s), (unsigned)Tpad);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < O * 8; i += Tpad) {   // BIAS[o] = {bias_cvr[o][0..5], bias_caps[o], 0}
    const int oo = i >> 3, c = i & 7;
    smem[L.BIAS + i] = c < 6 ? __ldg(a.bias_cvr + oo * 6 + c) : c == 6 ? __ldg(a.bias_caps + oo) : 0.0f;
  }
  __syncthreads();
  // The last warp stages: the first S images now, then one image per iteration into the stage that was just released;
  // it also prepares the object tile of the next image one iteration ahead.
  const bool stager = warp == (Tpad >> 5) - 1;
  if (stager && n_mine > 0) {
    for (int i = 0; i < S && i < n_mine; ++i) caps3_issue(a, L, smem, bar0, i, blockIdx.x + i * gridDim.x, lane);
    mbar_wait(c3_full(bar0, 0), 0);
    __syncwarp();   // edge floats were stored by other lanes of this warp
    caps3_object_tile<kSim>(a, o, L, smem, 0, blockIdx.x, lane);
  }
  __syncthreads();

  // ---- consumer set-up: everything that does not depend on the image, pinned in registers ---------------------------------
  const bool active = tid < T;
  const float inv_V = 1.0f / (float)V;
  const int k = keep(active ? fast_div(tid, inv_V) : -1);
  const int v = keep(active ? tid - k * V : 0);
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;
  C3Const cst[NP];
  unsigned vmask = 0;   // bit j: the thread's pair j exists
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int oj = k + G * j;
    const bool ok = active && oj < O;
    if (ok) {
      vmask |= 1u << j;
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
  vmask = keep(vmask);
  // the thread's first reduction job (phase B) is job k; soft-winner jobs (4..9) start from the dummy vote's term
  const float dv0 = (active && k >= 4 && k < 10) ? __ldg(a.dummy_vote + v * 6 + (k - 4)) : 0.0f;
  // lanes of this warp that work on the same objects (same k), their first lane, and which of the object's partial
  // slots this warp fills (an object's V lanes span up to `slots` warps)
  const unsigned segmask = keep(__match_any_sync(0xffffffffu, (unsigned)k));
  const bool seglead = lane == __ffs((int)segmask) - 1;
  const int n_warps = Tpad >> 5;
  const float e_dummy = expf(kDummyLog + kDummyLog);   // dummy mixing logit + dummy log-density (object_decoder.py:273-292)
  const float p_dummy = expf(kDummyLog);
  // shared-memory byte addresses.  Per thread, relative to an image's parameter block: its first pair's deformation
  // slots and vote-logit slot (the scale slot sits 4 V bytes further); the next pair of the thread is strideA further.
  const unsigned d_off = keep((unsigned)(4 * (active ? k * A + 6 * v : 0)));
  const unsigned l_off = keep((unsigned)(4 * (active ? k * A + 6 * V + 7 + v : 0)));
  const unsigned strideA = keep((unsigned)(4 * G * A)), V4 = keep((unsigned)(4 * V)), T4 = keep((unsigned)(4 * T));
  const unsigned strideR = keep((unsigned)(32 * G));
  const int Tk = keep(T);
  const unsigned r_off = keep((unsigned)(4 * (L.R + (active ? k * 8 : 0))));
  const unsigned x_off = keep((unsigned)(4 * (L.xs + v * 6)));
  const unsigned p_off = keep((unsigned)(4 * (L.ps + v)));
  const unsigned cp_rel = keep(4u * (unsigned)(active ? (k * L.slots + warp - ((k * V) >> 5)) * 2 : 0));
  const unsigned cp_step = keep((unsigned)(4 * G * L.slots * 2));
  const unsigned cp_bytes = (unsigned)(4 * O * L.slots * 2);
  const unsigned pv_addr = keep(bar0 + 4u * (unsigned)(L.PV + tid));
  const unsigned pv_col = keep(bar0 + 4u * (unsigned)(L.PV + v));   // partials of part v: item * T4 + kk * V4 further
  const unsigned stage_bytes = (unsigned)(4 * L.stage_stride);
  const unsigned OAmod = (unsigned)(O * A) & 3u, Pmod = (unsigned)P & 3u, V6mod = (unsigned)(V * 6) & 3u, Vmod = (unsigned)V & 3u;

  // normalised results of an image leave one iteration later (after the next image's barrier), so that one named barrier
  // per image is enough: the pair's posterior weight and mixing logit wait in registers
  float Eprev[NP], mlprev[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) Eprev[j] = mlprev[j] = 0.0f;
  int bprev = 0;
  auto late_outputs = [&](int par) {   // of the previous image; INVS / MLSE / LLV [par] are complete and visible
    if (active) {
      const float invS = smem[L.INVS + par * V + v], mlse = smem[L.MLSE + par * V + v];
      const size_t e0prev = (size_t)bprev * P + tid, e1prev = (size_t)bprev * (P + V) + tid;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (!(vmask >> j & 1u)) continue;
        if (kFull || o.posterior_mixing_prob) o.posterior_mixing_prob[e0prev + (size_t)(j * Tk)] = Eprev[j] * invS;
        if (kFull || o.mixing_log_prob) o.mixing_log_prob[e1prev + (size_t)(j * Tk)] = mlprev[j] - mlse;
      }
    }
    if (warp == 0) {
      float part = 0.0f;
      for (int vv = lane; vv < V; vv += 32) part += smem[L.LLV + par * V + vv];
      part = warp_sum(part);
      if (lane == 0 && (kFull || o.ll_per_example)) o.ll_per_example[bprev] = part;
    }
  };

  int s = 0;
  unsigned parity = 0;
  for (int i = 0; i < n_mine; ++i) {
    const int b = keep((int)(blockIdx.x + i * gridDim.x));
    const int par = i & 1;
    const unsigned st = keep(bar0 + 4u * (unsigned)L.stage0 + (unsigned)s * stage_bytes);
    // every input tensor is 16-byte aligned (checked by the host), so an image's block starts at its element offset
    // modulo 4 inside its staging buffer (bulk_run)
    const unsigned prm = keep(st + 4u * (unsigned)L.prm + 4u * (((unsigned)b * OAmod) & 3u));
    const unsigned nza = keep(st + 4u * (unsigned)L.nz + 4u * (((unsigned)b * Pmod) & 3u) + 4u * (unsigned)tid);
    const size_t e0 = keep((size_t)b * P + tid);             // element of the thread's first pair in a (B,O,V) tensor
    const size_t e1 = keep((size_t)b * (P + V) + tid);       // ... in a (B,O+1,V) tensor
    float* INVS = smem + L.INVS + par * V;
    float* MLSE = smem + L.MLSE + par * V;
    float* LLV = smem + L.LLV + par * V;
    float* REGP = smem + L.REGP + par * 32;
    const unsigned cpb = bar0 + 4u * (unsigned)L.CP + (unsigned)par * cp_bytes;

    mbar_wait(c3_full(bar0, s), parity);    // bulk-copied bytes visible to this thread

    // ---- phase A: the thread's pairs -------------------------------------------------------------------------------
    float xv[6];
    {
      const unsigned xa = st + x_off + 4u * (((unsigned)b * V6mod) & 3u);
#pragma unroll
      for (int c = 0; c < 6; ++c) xv[c] = lds_f32(xa + 4 * c);
    }
    const float pres = a.presence ? lds_f32(st + p_off + 4u * (((unsigned)b * Vmod) & 3u)) : 1.0f;
    float regsum = 0.0f, sE = 0.0f, sEvp = 0.0f, svp = 0.0f, best = -INFINITY, bvp = 0.0f;
    int bo = kNoWinner;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float Ecur[NP], mlcur[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float vp = 0.0f;
      Ecur[j] = mlcur[j] = 0.0f;
      if (vmask >> j & 1u) {
        const unsigned da = prm + d_off + (unsigned)j * strideA;   // row[6 v + c]
        const unsigned la = prm + l_off + (unsigned)j * strideA;   // row[6 V + 7 + v]
        float t[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float d = deform ? lds_f32(da + 4 * c) : 0.0f;
          regsum = fmaf(d, d, regsum);
          t[c] = d + cst[j].s[c];
        }
        PoseAffine pa;
        pose_affine_mufu<kSim>(t, pa);
        const float4 r0 = lds_f32x4(st + r_off + (unsigned)j * strideR), r1 = lds_f32x4(st + r_off + (unsigned)j * strideR + 16);
        const float r[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
        float vt[6];
        compose_vote(r, pa.a, vt);
#pragma unroll
        for (int c = 0; c < 6; ++c) sts_f32(da + 4 * c, vt[c]);   // in place: leaves with the coalesced copy below
        float lv = lds_f32(la) + cst[j].bv;
        if (a.noise_vote) lv += lds_f32(nza + (unsigned)j * T4);
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
          sc = softplus_fast(lds_f32(la + V4) + cst[j].bs) + 1e-2f;
          const float inv = rcp_approx(sc);
          lp = fmaf(-0.5f * q, inv * inv, fmaf(-6.0f * kLn2F, lg2_approx(sc), -6.0f * kHalfLog2Pi));
        } else {
          lp = fmaf(-0.5f, q, -6.0f * kHalfLog2Pi);
        }
        const float ml = vp < kLogSafeEps ? kLogSafeFloor : lg2_approx(vp) * kLn2F;   // log_safe (math_ops.py:18-21)
        const float pl = ml + lp;
        const float E = ex2_approx(pl * kLog2eF);
        Ecur[j] = E;
        mlcur[j] = ml;
        sE += E;
        sEvp = fmaf(E, vp, sEvp);
        svp += vp;
        if (pl > best) {        // objects ascend with j: lowest o on ties (torch.argmax)
          best = pl;
          bo = k + G * j;
          bvp = vp;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] = fmaf(E, vt[c], acc[c]);
        const size_t e = e0 + (size_t)(j * Tk);
        if (kFull || o.scale) o.scale[e] = sc;
        if (kFull || o.vote_presence) o.vote_presence[e] = vp;
        if (kFull || o.presence_logit_per_vote) o.presence_logit_per_vote[e] = lv;
        if (kFull || o.vote_presence_binary) o.vote_presence_binary[e] = ml > kDummyLog ? 1.0f : 0.0f;
        if (kFull || o.mixing_logit) o.mixing_logit[e1 + (size_t)(j * Tk)] = ml;
      }
      // capsule presence = max over parts (object_decoder.py:415); lowest part index wins ties.  vp >= 0, so the
      // unsigned order of its bits is the order of the floats.  Every lane of the warp takes part (lanes without a pair
      // in this pass share their mask with lanes that have none either).
      {
        const unsigned bits = __float_as_uint(vp);
        const unsigned m = redux_max_u32(segmask, bits);
        const unsigned arg = redux_min_u32(segmask, bits == m ? (unsigned)v : 0xffffffffu);
        if (seglead && (vmask >> j & 1u)) {
          sts_u32(cpb + cp_rel + (unsigned)j * cp_step, m);
          sts_u32(cpb + cp_rel + (unsigned)j * cp_step + 4, arg);
        }
      }
    }
    // The exchange tile is single-buffered: every consumer must have finished the previous image's jobs, which it has
    // once it arrived on that image's `empty` barrier.  (Rarely waits: a whole phase A lies in between.)
    if (i > 0) mbar_wait(c3_empty(bar0, s == 0 ? S - 1 : s - 1), s == 0 ? parity ^ 1u : parity);
    if (active) {   // the thread's partial sums over its objects, for the exchange between the object groups
      sts_f32(pv_addr + 0 * T4, sE);
      sts_f32(pv_addr + 1 * T4, sEvp);
      sts_f32(pv_addr + 2 * T4, svp);
      sts_f32(pv_addr + 3 * T4, best);
      sts_u32(pv_addr + 4 * T4, (unsigned)bo);
      sts_f32(pv_addr + 5 * T4, bvp);
#pragma unroll
      for (int c = 0; c < 6; ++c) sts_f32(pv_addr + (unsigned)(6 + c) * T4, acc[c]);
    }
    regsum = warp_sum(regsum);
    if (lane == 0) REGP[warp] = regsum;
    if (stager) {
      // refill the stage the previous image used (every thread has let go of it: the wait above) ...
      const int nxt = i - 1 + S;
      if (i >= 1 && nxt < n_mine) caps3_issue(a, L, smem, bar0, s == 0 ? S - 1 : s - 1, blockIdx.x + nxt * gridDim.x, lane);
      // ... and prepare the next image's object tile: its loads were issued at least one iteration ago
      if (i + 1 < n_mine) {
        const int sn = s + 1 == S ? 0 : s + 1;
        mbar_wait(c3_full(bar0, sn), s + 1 == S ? parity ^ 1u : parity);
        __syncwarp();
        caps3_object_tile<kSim>(a, o, L, smem, sn, blockIdx.x + (i + 1) * gridDim.x, lane);
      }
    }
    named_bar_sync(kC3BarConsumers, (unsigned)Tpad);   // the image's only barrier

    // ---- phase B: one reduction job per object group, over the G partials of part v -----------------------------------------
    //   0: sum E -> 1 / S, log S (the per-part log-likelihood)     1: soft winner presence     2: logsumexp of the mixing
    //   logits, dummy rows     3: hard winner     4..9: a pose dimension of the soft winner
    if (active) {
      for (int job = k; job < 10; job += G) {
        const size_t bv = (size_t)b * V + v;
        if (job == 3) {
          float bb = -INFINITY, bbvp = 0.0f;
          int bbo = kNoWinner;
          unsigned pa = pv_col + 3 * T4;
          for (int kk = 0; kk < G; ++kk, pa += V4) {
            const float cb = lds_f32(pa);
            const int co = (int)lds_u32(pa + T4);
            if (cb > bb || (cb == bb && co < bbo)) {
              bb = cb;
              bbo = co;
              bbvp = lds_f32(pa + 2 * T4);
            }
          }
          if (bbo == kNoWinner) {   // every logit is NaN: the first object, like the other paths
            bbo = 0;
            bbvp = 0.0f;
          }
          if (kFull || o.winner) {
            const unsigned wa = prm + (unsigned)(4 * (bbo * A + 6 * v));
#pragma unroll
            for (int c = 0; c < 6; ++c) o.winner[bv * 6 + c] = lds_f32(wa + 4 * c);
          }
          if (kFull || o.winner_presence) o.winner_presence[bv] = bbvp;
          if (kFull || o.winner_idx) o.winner_idx[bv] = bbo;
          if (kFull || o.is_from_capsule) o.is_from_capsule[bv] = fast_div(bbo, inv_V);   // sic (object_decoder.py:334)
        } else if (job == 2) {
          float Svp = p_dummy;
          unsigned pa = pv_col + 2 * T4;
          for (int kk = 0; kk < G; ++kk, pa += V4) Svp += lds_f32(pa);
          const float mlse = logf(Svp);   // logsumexp_o of the mixing logits: exp(log_safe(vp)) = vp
          MLSE[v] = mlse;
          const size_t dummy_row = (size_t)b * (P + V) + P + v;
          if (kFull || o.mixing_logit) o.mixing_logit[dummy_row] = kDummyLog;
          if (kFull || o.mixing_log_prob) o.mixing_log_prob[dummy_row] = kDummyLog - mlse;
        } else {
          // the others need S = sum_o E + the dummy component; job 0 publishes it, 1 and 4..9 normalise one more sum with it
          const unsigned item = job == 0 ? 0u : job == 1 ? 1u : (unsigned)(job + 2);
          float Ssum = e_dummy;
          float other = job < 4 ? 0.0f : e_dummy * (job == k ? dv0 : __ldg(a.dummy_vote + v * 6 + (job - 4)));
          unsigned pa = pv_col;
          for (int kk = 0; kk < G; ++kk, pa += V4) {
            Ssum += lds_f32(pa);
            other += lds_f32(pa + item * T4);
          }
          const float invS = __frcp_rn(Ssum);
          if (job == 0) {
            const float lse = logf(Ssum);
            INVS[v] = invS;
            LLV[v] = lse * pres;
            if (kFull || o.log_prob_per_point) o.log_prob_per_point[bv] = lse;
          } else if (job == 1) {
            if (kFull || o.soft_winner_presence) o.soft_winner_presence[bv] = other * invS;   // the dummy has presence 0
          } else {
            if (kFull || o.soft_winner) o.soft_winner[bv * 6 + (job - 4)] = other * invS;
          }
        }
      }
      // the votes: the object's 6 V floats are contiguous in its row; lane v copies elements v, v + V, ...
      if (kFull || o.vote) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          if (!(vmask >> j & 1u)) continue;
          const unsigned va = prm + (unsigned)(4 * ((k + G * j) * A + v));
          float* gv = o.vote + ((size_t)b * P + (size_t)(k + G * j) * V) * 6 + v;
#pragma unroll
          for (int m = 0; m < 6; ++m) gv[m * V] = lds_f32(va + (unsigned)m * V4);
        }
      }
    }
    if (warp == 0) {
      // capsule presence of object `oo`: combine the partials of the warps its V lanes span, parts ascending
      for (int oo = lane; oo < O; oo += 32) {
        const int ko = oo % G, w0 = (ko * V) >> 5, n = ((ko * V + V - 1) >> 5) - w0 + 1;
        const unsigned ca = cpb + 4u * (unsigned)(oo * L.slots * 2);
        unsigned m = lds_u32(ca), arg = lds_u32(ca + 4);
        for (int q = 1; q < n; ++q) {
          const unsigned cm = lds_u32(ca + 8 * q);
          if (cm > m) {
            m = cm;
            arg = lds_u32(ca + 8 * q + 4);
          }
        }
        if (arg == 0xffffffffu) arg = 0;
        if (kFull || o.caps_presence) o.caps_presence[(size_t)b * O + oo] = __uint_as_float(m);
        if (kFull || o.caps_presence_arg) o.caps_presence_arg[(size_t)b * O + oo] = (int)arg;
      }
      const float reg_total = warp_sum(lane < n_warps ? REGP[lane] : 0.0f);
      if (lane == 0 && (kFull || o.reg_per_example)) o.reg_per_example[b] = 0.5f * reg_total;
    }
    fence_proxy_async();   // this thread's in-place writes are ordered before the bulk copy that refills the stage
    mbar_arrive(c3_empty(bar0, s));

    // ---- the previous image's normalised results (its normalisers were published before this image's barrier) --------
    if (i > 0) late_outputs(par ^ 1);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      Eprev[j] = Ecur[j];
      mlprev[j] = mlcur[j];
    }
    bprev = b;
    if (++s == S) {
      s = 0;
      parity ^= 1u;
    }
  }
  if (n_mine > 0) {
    named_bar_sync(kC3BarConsumers, (unsigned)Tpad);
    late_outputs((n_mine - 1) & 1);
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)

struct Caps3Plan {
  int NP, max_threads, min_blocks;   // the compiled variant
  int threads;                       // G * V, padded to whole warps
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
  const int budget = max_smem_optin();
  struct Cand {
    int NP, max_threads, min_blocks;
  };
  const Cand cands[] = {{2, 640, 1}, {4, 320, 2}, {1, 416, 2}, {1, 640, 1}, {4, 512, 1}};
  const int force_np = c3_env_int("SCAE_CAPS3_NP", 0), force_mb = c3_env_int("SCAE_CAPS3_MINB", 0);
  const int force_s = c3_env_int("SCAE_CAPS3_STAGES", 0);
  double best_eff = 0.0;
  int best_resident = 0;
  bool found = false;
  for (const Cand& c : cands) {
    if (force_np && c.NP != force_np) continue;
    if (force_mb && c.min_blocks != force_mb) continue;
    const int G = (O + c.NP - 1) / c.NP;
    const int T = G * V, threads = (T + 31) & ~31;
    if (threads > c.max_threads) continue;
    const double eff = (double)O / ((double)G * c.NP);
    // stages: as many as fit (at most kC3MaxStages); a two-CTA variant must leave room for its twin
    const int room = c.min_blocks == 2 ? (228 * 1024) / 2 - 1024 : budget;
    int S = force_s >= 2 && force_s <= kC3MaxStages ? force_s : kC3MaxStages;
    Caps3FwdLayout L = caps3_fwd_layout(a, G, c.NP, S);
    while (S > 2 && (size_t)L.total * sizeof(float) > (size_t)room) L = caps3_fwd_layout(a, G, c.NP, --S);   // (>= 2: prefetch)
    if ((size_t)L.total * sizeof(float) > (size_t)room) continue;
    const size_t smem = (size_t)L.total * sizeof(float);
    // per-thread register cap of the variant: the fullest SM sub-partition holds ceil(warps / 4) warps of 16384 / 32 lanes
    const int regs = 512 / ((c.max_threads / 32 * c.min_blocks + 3) / 4) & ~7;
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
  return (long)a->O * a->V < (1L << 22) && aligned16(a->all_param) && aligned16(a->cpr_static) && aligned16(a->x) &&
         (!a->noise_vote || aligned16(a->noise_vote)) && (!a->noise_caps || aligned16(a->noise_caps)) &&
         (!a->presence || aligned16(a->presence));
}

template <bool kSim, bool kFull>
static int caps3_fwd_launch(const scae_caps_args* a, const scae_caps_outputs* out, const Caps3Plan& plan,
                            cudaStream_t stream) {
  void (*kern)(const scae_caps_args, const scae_caps_outputs, const Caps3FwdLayout) = nullptr;
  if (plan.NP == 1 && plan.min_blocks == 2) kern = caps3_fwd_kernel<kSim, 1, 416, 2, kFull>;
  else if (plan.NP == 1) kern = caps3_fwd_kernel<kSim, 1, 640, 1, kFull>;
  else if (plan.NP == 2) kern = caps3_fwd_kernel<kSim, 2, 640, 1, kFull>;
  else if (plan.min_blocks == 2) kern = caps3_fwd_kernel<kSim, 4, 320, 2, kFull>;
  else kern = caps3_fwd_kernel<kSim, 4, 512, 1, kFull>;
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
  const bool full = out->vote && out->scale && out->vote_presence && out->presence_logit_per_caps &&
                    out->presence_logit_per_vote && out->caps_presence && out->caps_presence_arg &&
                    out->log_prob_per_point && out->ll_per_example && out->reg_per_example && out->vote_presence_binary &&
                    out->winner && out->winner_presence && out->winner_idx && out->is_from_capsule && out->soft_winner &&
                    out->soft_winner_presence && out->posterior_mixing_prob && out->mixing_log_prob && out->mixing_logit;
  const int rc = sim ? (full ? caps3_fwd_launch<true, true>(a, out, plan, stream) : caps3_fwd_launch<true, false>(a, out, plan, stream))
                     : (full ? caps3_fwd_launch<false, true>(a, out, plan, stream) : caps3_fwd_launch<false, false>(a, out, plan, stream));
  if (rc != SCAE_OK) return rc;
  *handled = true;
  return SCAE_OK;
}

}  // namespace scae
