// Hot path 2 (sm_100a), persistent backward kernel of the capsule part-pose mixture likelihood (see caps_ll3.cu for the
// forward and the design notes; reference object_decoder.py:160-236, :257-372; formulas: oracle/manual_backward.py::
// capsule_forward_backward, transcribed from caps_ll2.cu).
//
// One persistent CTA per SM, 20 warps (96 registers), every thread keeps the same (object group k, part v) for all its
// images: thread t = k V + v owns the pairs t, t + T, ... (NP of them).  Per image:
//
//   main pass    per pair: recompute the forward (MUFU forms), the pair's gradients; the gradient row is written IN PLACE
//                over the staged all_param block (MLP-ReLU mask and deformation regulariser applied) and leaves through
//                ONE bulk store; the 7 per-object sums over the parts go to a [7][P] tile; the batch sums that become the
//                gradients of cpr_static / bias_vote / bias_scale accumulate in REGISTERS (the thread sees the same
//                pairs in every image);
//   pre-pass     of the NEXT image (its stage has landed): the thread's share of S[v] = sum_o posterior * upstream, so
//                that the image needs only one barrier;
//   barrier      the image's only one (named, all threads);
//   phase B      two threads per (object, slot) sum the [7][P] tile; warp 0 turns the PREVIOUS image's sums into the 7
//                capsule-level gradients (transform backward; one lane per object) and writes them straight to global
//                memory after that image's bulk store has completed; lane 0 issues this image's bulk store.
//
// The last warp stages: bulk copies (TMA) of every per-image input into a ring of S stages and the per-object tile of
// the next image (capsule transform with its intermediates for the backward, capsule presence).
//
// Deterministic: fixed summation orders everywhere; the per-CTA batch sums are reduced by launch_reduce_rows.
#include <stdlib.h>
#include <string.h>

#include "caps_common.cuh"

namespace scae {

constexpr int kB3MaxStages = 4;
constexpr int kB3Save = 12;    // floats per object kept for the capsule-level backward
constexpr int kB3SaveBufs = 4; // image j's tile is written in iteration j-1 and read in iteration j+1
constexpr unsigned kB3Bar = 1;

__host__ __device__ inline int b3_round4(int n) { return (n + 3) & ~3; }

constexpr int kB3PG = 3;        // slots of the posterior ring
constexpr int kB3Tin = 12;      // floats per object of the object-tile inputs: 7 raw parameters, noise, 3 upstream values

struct Caps3BwdLayout {
  int S, G, NP, T, Tpad, Vp;
  int stage0, stage_stride;                      // floats; the in-place ring: S stages
  int prm, nz, xs, ps, lse, R, OS;               // offsets inside a stage
  int gsw, gswp, gw, gwp, widx;                  // ... winner-gradient variant only: upstream gradients of the soft /
                                                 // hard winner (V x 6, V each) and the saved winner index (int64)
  int pg0, pg_stride, gpost;                     // the posterior ring: kB3PG slots of {posterior, its upstream gradient}
  int RED7, SPART, OBJ7, OSAVE, TIN, SC, BIAS, OBJSUM, CST, GX, DSUM, total;   // CTA-wide tiles
  // byte strides / offsets the hot loop uses straight from the constant bank
  unsigned strideA4, V4, T4, strideR, strideO, nz4, gpost4, red_plane, red_step;
};

static Caps3BwdLayout caps3_bwd_layout(const scae_caps_args* a, const scae_caps_upstream* up, int G, int NP, int S,
                                       bool soft = false, bool want_gx = false) {
  const int O = a->O, V = a->V, A = 8 * V + 7, P = O * V;
  Caps3BwdLayout L;
  L.S = S, L.G = G, L.NP = NP, L.T = G * V, L.Tpad = (L.T + 31) & ~31;
  int at = 32;   // [0, 32) floats: mbarriers full[4], bdone[2], pgfull[3], ready[2]
  auto take = [&](int n) {
    const int here = at;
    at += b3_round4(n);
    return here;
  };
  L.stage0 = at;
  L.prm = take(O * A + 4) - L.stage0;
  L.nz = take(a->noise_vote ? P + 4 : 0) - L.stage0;
  L.xs = take(V * 6 + 4) - L.stage0;
  L.ps = take(a->presence ? V + 4 : 0) - L.stage0;
  L.lse = take(V + 4) - L.stage0;
  L.R = take(O * 8) - L.stage0;
  L.OS = take(O * 4) - L.stage0;
  L.gsw = take(soft && up->g_soft_winner ? V * 6 + 4 : 0) - L.stage0;
  L.gswp = take(soft && up->g_soft_winner_presence ? V + 4 : 0) - L.stage0;
  L.gw = take(soft && up->g_winner ? V * 6 + 4 : 0) - L.stage0;
  L.gwp = take(soft && up->g_winner_presence ? V + 4 : 0) - L.stage0;
  L.widx = take(soft && (up->g_winner || up->g_winner_presence) ? 2 * V + 4 : 0) - L.stage0;
  L.stage_stride = at - L.stage0;
  at = L.stage0 + S * L.stage_stride;
  L.pg0 = at;
  take(P + 4);
  L.gpost = take(up->g_posterior_mixing_prob ? P + 4 : 0) - L.pg0;
  L.pg_stride = at - L.pg0;
  at = L.pg0 + kB3PG * L.pg_stride;
  L.Vp = V | 1;                 // odd row pitch of the [7][O][Vp] tile: the row sums of phase B are conflict-free
  L.RED7 = take(7 * O * L.Vp);
  L.SPART = take(2 * L.T);
  L.OBJ7 = take(2 * O * 8);
  L.OSAVE = take(kB3SaveBufs * O * kB3Save);
  L.TIN = take(O * kB3Tin + 4);   // + the image's two upstream scalars
  L.SC = take(2 * 4);             // {g_ll, g_reg} of the image, by image parity
  L.BIAS = take(O * 8);
  L.OBJSUM = take(O * 8);
  // [NP][8][T]: cpr_static (6), bias_vote, bias_scale + 0.5 of the thread's pairs (the winner-gradient variant needs the
  // room for its extra tiles and reads them from global memory / L2 instead)
  L.CST = take(soft ? 0 : 8 * L.T * NP);
  L.GX = take(want_gx ? 6 * L.T : 0);          // the threads' shares of g_x
  L.DSUM = take(soft ? V * 6 : 0);             // batch sum of the dummy vote's gradient
  L.total = at;
  L.strideA4 = 4u * (unsigned)(G * A), L.V4 = 4u * (unsigned)V, L.T4 = 4u * (unsigned)L.T;
  L.strideR = 32u * (unsigned)G, L.strideO = 16u * (unsigned)G;
  L.red_plane = 4u * (unsigned)(O * L.Vp), L.red_step = 4u * (unsigned)(G * L.Vp);
  L.nz4 = 4u * (unsigned)L.nz, L.gpost4 = 4u * (unsigned)L.gpost;
  return L;
}

__device__ __forceinline__ unsigned b3_full(unsigned bar0, int s) { return bar0 + 8u * (unsigned)s; }
__device__ __forceinline__ unsigned b3_bdone(unsigned bar0, int p) { return bar0 + 8u * (unsigned)(kB3MaxStages + p); }
__device__ __forceinline__ unsigned b3_pgfull(unsigned bar0, int p) { return bar0 + 8u * (unsigned)(kB3MaxStages + 2 + p); }
__device__ __forceinline__ unsigned b3_ready(unsigned bar0, int p) { return bar0 + 8u * (unsigned)(kB3MaxStages + 2 + kB3PG + p); }

struct Caps3BwdOut {
  float* g_all_param;   // [B,O,A] final: ReLU mask and regulariser applied
  float* g_presence;    // [B,V] nullable
  float* partials;      // [grid][O*A] per-CTA batch sums of the pre-activation gradient
  float* g_x;           // [B,V,6] nullable (winner-gradient variant)
  float* dummy_partials;   // [grid][V*6] nullable: per-CTA batch sums of the dummy vote's gradient
};

// ---- staging (one warp): each lane owns one contiguous per-image input ----------------------------------------------------
struct B3Run {
  const float* g;
  float* base;
  int n;
};

// the lane's run (n = 0: none) -> shared memory, completion on mbarrier `bar`: interior by bulk copy, edges through registers
__device__ __forceinline__ void caps3_bwd_issue_runs(const B3Run& run, unsigned bar, int lane) {
  BulkRun br = {0, 0, 0, 0};
  if (run.n) br = bulk_run(run.g, run.n);
  unsigned bytes = 4u * (unsigned)br.body;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
  if (lane == 0) mbar_expect_tx(bar, bytes);
  __syncwarp();
  if (br.body) bulk_g2s(run.base + br.off + br.head, run.g + br.head, 4u * (unsigned)br.body, bar);
  if (run.n) {   // the run's (at most 3 + 3) edge floats
    for (int q = 0; q < br.head; ++q) run.base[br.off + q] = __ldg(run.g + q);
    for (int q = 0; q < br.tail; ++q) run.base[br.off + br.head + br.body + q] = __ldg(run.g + br.head + br.body + q);
  }
}

// in-place stage s <- image b: all_param block (becomes the gradient block), noise rows, part poses / presences / lse
__device__ __forceinline__ void caps3_bwd_issue(const scae_caps_args& a, const scae_caps_saved& sv,
                                                const scae_caps_upstream& up, bool soft, const Caps3BwdLayout& L,
                                                float* smem, unsigned bar0, int s, int b, int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V;
  float* st = smem + L.stage0 + s * L.stage_stride;
  B3Run run = {nullptr, nullptr, 0};
  switch (lane) {
    case 0: run = {a.all_param + (size_t)b * O * A, st + L.prm, O * A}; break;
    case 1: if (a.noise_vote) run = {a.noise_vote + (size_t)b * P, st + L.nz, P}; break;
    case 2: run = {a.x + (size_t)b * V * 6, st + L.xs, V * 6}; break;
    case 3: if (a.presence) run = {a.presence + (size_t)b * V, st + L.ps, V}; break;
    case 4: run = {sv.log_prob_per_point + (size_t)b * V, st + L.lse, V}; break;
    case 5: if (soft && up.g_soft_winner) run = {up.g_soft_winner + (size_t)b * V * 6, st + L.gsw, V * 6}; break;
    case 6: if (soft && up.g_soft_winner_presence) run = {up.g_soft_winner_presence + (size_t)b * V, st + L.gswp, V}; break;
    case 7: if (soft && up.g_winner) run = {up.g_winner + (size_t)b * V * 6, st + L.gw, V * 6}; break;
    case 8: if (soft && up.g_winner_presence) run = {up.g_winner_presence + (size_t)b * V, st + L.gwp, V}; break;
    case 9:
      if (soft && (up.g_winner || up.g_winner_presence))
        run = {reinterpret_cast<const float*>(sv.winner_idx) + (size_t)b * V * 2, st + L.widx, V * 2};
      break;
    default: break;
  }
  caps3_bwd_issue_runs(run, b3_full(bar0, s), lane);
}

// posterior ring slot <- image b: saved posterior and its upstream gradient (what the pre-pass needs one image ahead)
__device__ __forceinline__ void caps3_bwd_issue_pg(const scae_caps_saved& sv, const scae_caps_upstream& up,
                                                   const Caps3BwdLayout& L, float* smem, unsigned bar0, int slot, int b,
                                                   int P, int lane) {
  float* pg = smem + L.pg0 + slot * L.pg_stride;
  B3Run run = {nullptr, nullptr, 0};
  if (lane == 0) run = {sv.posterior_mixing_prob + (size_t)b * P, pg, P};
  else if (lane == 1 && up.g_posterior_mixing_prob) run = {up.g_posterior_mixing_prob + (size_t)b * P, pg + L.gpost, P};
  caps3_bwd_issue_runs(run, b3_pgfull(bar0, slot), lane);
}

// object-tile inputs of image b -> TIN[o] = {7 raw capsule-level parameters, noise, upstream gradient of caps_presence, its
// arg-max part, upstream gradient of the presence logit} by 4-byte asynchronous copies (no registers held meanwhile);
// lane = object, so the lane that copies a row is the one that reads it
__device__ __forceinline__ void caps3_bwd_tile_inputs(const scae_caps_args& a, const scae_caps_saved& sv,
                                                      const scae_caps_upstream& up, const Caps3BwdLayout& L, float* smem,
                                                      int b, int lane) {
  const int O = a.O, V = a.V, A = 8 * V + 7;
  for (int oo = lane; oo < O; oo += 32) {
    float* dst = smem + L.TIN + oo * kB3Tin;
    const float* row = a.all_param + ((size_t)b * O + oo) * A + 6 * V;
#pragma unroll
    for (int p = 0; p < 7; ++p) cp_async4(dst + p, row + p);
    const size_t bo = (size_t)b * O + oo;
    if (a.noise_caps) cp_async4(dst + 7, a.noise_caps + bo);
    if (up.g_caps_presence) {
      cp_async4(dst + 8, up.g_caps_presence + bo);
      cp_async4(dst + 9, reinterpret_cast<const float*>(sv.caps_presence_arg) + bo);
    }
    if (up.g_presence_logit_per_caps) cp_async4(dst + 10, up.g_presence_logit_per_caps + bo);
  }
  if (lane == 0 && up.g_ll_per_example) cp_async4(smem + L.TIN + O * kB3Tin, up.g_ll_per_example + b);
  if (lane == 1 && up.g_reg_per_example) cp_async4(smem + L.TIN + O * kB3Tin + 1, up.g_reg_per_example + b);
  cp_async_commit();
}

// per-object work of image b in stage s (one warp, lane = object):
//   R[o]     = {capsule -> viewer affine (6), capsule presence probability, -}
//   OS[o]    = {upstream gradient of caps_presence, its arg-max part (int bits), -, -}
//   SAVE[o]  = {sx, sy, sh, tx, ty, cos, sin, presence probability, upstream gradient of the presence logit,
//               bit c: raw parameter 6V + c is positive (the MLP's ReLU was active), -, -}
template <bool kSim>
__device__ __forceinline__ void caps3_bwd_object_tile(const scae_caps_args& a, const scae_caps_upstream& up,
                                                      const Caps3BwdLayout& L, float* smem, int s, int img, int lane) {
  const int O = a.O;
  float* st = smem + L.stage0 + s * L.stage_stride;
  const float* BIAS = smem + L.BIAS;
  float* R = st + L.R;
  float* OS = st + L.OS;
  float* SAVE = smem + L.OSAVE + (img % kB3SaveBufs) * O * kB3Save;
  for (int oo = lane; oo < O; oo += 32) {
    const float* in = smem + L.TIN + oo * kB3Tin;
    const float4 i0 = *reinterpret_cast<const float4*>(in), i1 = *reinterpret_cast<const float4*>(in + 4),
                 i2 = *reinterpret_cast<const float4*>(in + 8);
    const float4 b0 = *reinterpret_cast<const float4*>(BIAS + oo * 8), b1 = *reinterpret_cast<const float4*>(BIAS + oo * 8 + 4);
    const float raw[7] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z};
    const float t[6] = {raw[0] + b0.x, raw[1] + b0.y, raw[2] + b0.z, raw[3] + b0.w, raw[4] + b1.x, raw[5] + b1.y};
    PoseAffine r;
    pose_affine_mufu<kSim>(t, r);
    float lc = raw[6] + b1.z;
    if (a.noise_caps) lc += i1.w;
    const float pc = sigmoid_fast(lc);
    float4* dst = reinterpret_cast<float4*>(R + oo * 8);
    dst[0] = make_float4(r.a[0], r.a[1], r.a[2], r.a[3]);
    dst[1] = make_float4(r.a[4], r.a[5], pc, 0.0f);
    const bool gc = up.g_caps_presence != nullptr;
    *reinterpret_cast<float4*>(OS + oo * 4) = make_float4(gc ? i2.x : 0.0f, gc ? i2.y : __int_as_float(-1), 0.0f, 0.0f);
    unsigned mask = 0;
#pragma unroll
    for (int p = 0; p < 7; ++p) mask |= raw[p] > 0.0f ? 1u << p : 0u;
    float4* sd = reinterpret_cast<float4*>(SAVE + oo * kB3Save);
    sd[0] = make_float4(r.sx, r.sy, r.sh, r.tx);
    sd[1] = make_float4(r.ty, r.c, r.s, pc);
    sd[2] = make_float4(up.g_presence_logit_per_caps ? i2.z : 0.0f, __uint_as_float(mask), 0.0f, 0.0f);
  }
  if (lane == 0) smem[L.SC + (img & 1) * 4] = up.g_ll_per_example ? smem[L.TIN + O * kB3Tin] : 0.0f;
  if (lane == 1) smem[L.SC + (img & 1) * 4 + 1] = up.g_reg_per_example ? smem[L.TIN + O * kB3Tin + 1] : 0.0f;
}

// the 8 batch-shared constants of a pair: from the thread's shared-memory slots, or (kGlobal) straight from global memory
template <bool kGlobal>
struct C3PairConst {
  unsigned ca, T4;
  const float *stc, *bv, *bs;   // cpr_static + 6 p, bias_vote + p, bias_scale + p
  __device__ __forceinline__ float get(int c) const {
    if (!kGlobal) return lds_f32(ca + (unsigned)c * T4);
    return c < 6 ? __ldg(stc + c) : c == 6 ? __ldg(bv) : __ldg(bs) + 0.5f;
  }
};

// forward of one pair from its staged parameters: the vote and its presence (what the winner-gradient pre-pass needs)
template <bool kSim, class Const>
__device__ __forceinline__ void caps3_pair_vote(unsigned da, unsigned la, const Const& cc, unsigned ra, unsigned nza,
                                                bool deform, bool noise, float vt[6], float& vp) {
  float t[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) t[c] = (deform ? lds_f32(da + 4 * c) : 0.0f) + cc.get(c);
  PoseAffine pa;
  pose_affine_mufu<kSim>(t, pa);
  const float4 r0 = lds_f32x4(ra), r1 = lds_f32x4(ra + 16);
  const float r[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
  compose_vote(r, pa.a, vt);
  float lv = lds_f32(la) + cc.get(6);
  if (noise) lv += lds_f32(nza);
  vp = r1.z * sigmoid_fast(lv);
}

// kExtras: some per-pair upstream gradient beyond the training set (vote_presence, vote, scale, presence_logit_per_vote,
// mixing_logit) is given; without it no pointer is tested in the pair loop
// kSoft: the upstream gradients of the soft / hard winner and the part-side input gradient g_x are handled (the class
// default vote_type = 'soft'): S[v] then needs every object's vote, so the pre-pass recomputes the forward of the
// image it is about to process and the image takes two barriers
template <bool kSim, int NP, int kMaxT, int kMinB, bool kExtras, bool kSoft>
__global__ void __launch_bounds__(kMaxT, kMinB) caps3_bwd_kernel(const scae_caps_args a, const scae_caps_saved sv,
                                                                 const scae_caps_upstream up, const Caps3BwdOut out,
                                                                 const Caps3BwdLayout L) {
  SCAE_DYNAMIC_SMEM(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int O = a.O, V = a.V, A = 8 * V + 7, P = O * V;
  const int S = L.S, G = L.G, T = L.T, Tpad = L.Tpad;
  const unsigned bar0 = smem_u32(smem);
  const int n_mine = ((int)blockIdx.x < a.B) ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const bool deform = (a.flags & SCAE_CAPS_ALLOW_DEFORM) != 0;
  const bool learn = (a.flags & SCAE_CAPS_LEARN_VOTE_SCALE) != 0;
  const bool relu = (a.flags & SCAE_CAPS_RELU_GRAD) != 0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(b3_full(bar0, s), 1);
    }
    mbar_init(b3_bdone(bar0, 0), (unsigned)Tpad);
    mbar_init(b3_bdone(bar0, 1), (unsigned)Tpad);
    for (int p = 0; p < kB3PG; ++p) mbar_init(b3_pgfull(bar0, p), 1);
    mbar_init(b3_ready(bar0, 0), 1);
    mbar_init(b3_ready(bar0, 1), 1);
    fence_mbar_init();
  }
  for (int i = tid; i < O * 8; i += Tpad) {
    const int oo = i >> 3, c = i & 7;
    smem[L.BIAS + i] = c < 6 ? __ldg(a.bias_cvr + oo * 6 + c) : c == 6 ? __ldg(a.bias_caps + oo) : 0.0f;
    smem[L.OBJSUM + i] = 0.0f;
  }
  if (kSoft)
    for (int i = tid; i < V * 6; i += Tpad) smem[L.DSUM + i] = 0.0f;
  __syncthreads();
  const bool stager = warp == (Tpad >> 5) - 1;
  const int chain_warp = Tpad > 32 ? 1 : 0;
  if (stager && n_mine > 0) {
    caps3_bwd_tile_inputs(a, sv, up, L, smem, blockIdx.x, lane);
    for (int i = 0; i < kB3PG && i < n_mine; ++i)
      caps3_bwd_issue_pg(sv, up, L, smem, bar0, i, blockIdx.x + i * gridDim.x, P, lane);
    for (int i = 0; i < S && i < n_mine; ++i) caps3_bwd_issue(a, sv, up, kSoft, L, smem, bar0, i, blockIdx.x + i * gridDim.x, lane);
    cp_async_wait<0>();
    caps3_bwd_object_tile<kSim>(a, up, L, smem, 0, 0, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(b3_ready(bar0, 0));
  }

  // ---- per-thread constants ----------------------------------------------------------------------------------------------
  const bool active = tid < T;
  const float inv_V = 1.0f / (float)V;
  const int k = keep(active ? fast_div(tid, inv_V) : -1);
  const int v = keep(active ? tid - k * V : 0);
  float acc[NP][8];   // batch sums of the pre-activation gradient: 6 cpr slots, vote logit, scale
  unsigned vmask = 0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int oj = k + G * j;
    const bool ok = active && oj < O;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.0f;
    if (ok) {
      vmask |= 1u << j;
      // the pair's batch-shared parameters: thread-private shared-memory slots [j][c][tid] (conflict-free)
      const int p = oj * V + v;
      if (!kSoft) {
        float* cs = smem + L.CST + j * 8 * T + tid;
#pragma unroll
        for (int c = 0; c < 6; ++c) cs[c * T] = __ldg(a.cpr_static + (size_t)p * 6 + c);
        cs[6 * T] = __ldg(a.bias_vote + p);
        cs[7 * T] = __ldg(a.bias_scale + p) + 0.5f;
      }
    }
  }
  vmask = keep(vmask);
  const unsigned d_off = keep((unsigned)(4 * (active ? k * A + 6 * v : 0)));
  const unsigned l_off = keep((unsigned)(4 * (active ? k * A + 6 * V + 7 + v : 0)));
  const unsigned strideA = L.strideA4, V4 = L.V4, T4 = L.T4, strideR = L.strideR, strideO = L.strideO;
  const unsigned pg_bytes = (unsigned)(4 * L.pg_stride), pg_base = bar0 + 4u * (unsigned)L.pg0;
  const unsigned cst_addr = keep(bar0 + 4u * (unsigned)(L.CST + tid));
  const unsigned r_off = keep((unsigned)(4 * (L.R + (active ? k * 8 : 0))));
  const unsigned os_off = keep((unsigned)(4 * (L.OS + (active ? k * 4 : 0))));
  const unsigned red_addr = keep(bar0 + 4u * (unsigned)(L.RED7 + (active ? k * L.Vp + v : 0)));
  const unsigned stage_bytes = (unsigned)(4 * L.stage_stride);
  const unsigned OAmod = (unsigned)(O * A) & 3u, Pmod = (unsigned)P & 3u, V6mod = (unsigned)(V * 6) & 3u, Vmod = (unsigned)V & 3u;
  const bool have_gpost = up.g_posterior_mixing_prob != nullptr;
  __syncthreads();   // the first image's object tile (written by the stager above) is visible

  // the thread's share of S[v] = sum_o posterior * upstream for image `img` in stage `sn` -> SPART[img & 1]
  auto pre_pass = [&](int img) {
    const int bn = blockIdx.x + img * gridDim.x, slot = img % kB3PG;
    mbar_wait(b3_pgfull(bar0, slot), (unsigned)((img / kB3PG) & 1));
    float part = 0.0f;
    if (have_gpost) {
      const unsigned po = pg_base + (unsigned)slot * pg_bytes + 4u * (((unsigned)bn * Pmod) & 3u) + 4u * (unsigned)tid;
#pragma unroll
      for (int j = 0; j < NP; ++j)
        if (vmask >> j & 1u) part = fmaf(lds_f32(po + (unsigned)j * T4), lds_f32(po + L.gpost4 + (unsigned)j * T4), part);
    }
    if (active) smem[L.SPART + (img & 1) * T + tid] = part;
  };
  if (!kSoft && n_mine > 0) pre_pass(0);
  __syncthreads();

  // the capsule-level gradients of image `img` (warp 0, one lane per object) from the sums phase B left in OBJ7
  auto object_chain = [&](int img) {
    const int bi = blockIdx.x + img * gridDim.x;
    const float* G7 = smem + L.OBJ7 + (img & 1) * O * 8;
    const float* SAVE = smem + L.OSAVE + (img % kB3SaveBufs) * O * kB3Save;
    float* OBJSUM = smem + L.OBJSUM;
    for (int oo = lane; oo < O; oo += 32) {
      const float4 g0 = *reinterpret_cast<const float4*>(G7 + oo * 8), g1 = *reinterpret_cast<const float4*>(G7 + oo * 8 + 4);
      const float4 s0 = *reinterpret_cast<const float4*>(SAVE + oo * kB3Save);
      const float4 s1 = *reinterpret_cast<const float4*>(SAVE + oo * kB3Save + 4);
      const float4 s2 = *reinterpret_cast<const float4*>(SAVE + oo * kB3Save + 8);
      PoseAffine r;
      r.sx = s0.x, r.sy = s0.y, r.sh = s0.z, r.tx = s0.w, r.ty = s1.x, r.c = s1.y, r.s = s1.z;
      const float pc = s1.w;
      const float g7[6] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y};
      float gt[7];
      pose_affine_bwd<kSim>(g7, r, gt);
      gt[6] = g1.z * pc * (1.0f - pc) + s2.x;
      const unsigned mask = __float_as_uint(s2.y);
      float* grow = out.g_all_param + ((size_t)bi * O + oo) * A + 6 * V;
#pragma unroll
      for (int p = 0; p < 7; ++p) {
        OBJSUM[oo * 8 + p] += gt[p];
        grow[p] = (relu && !(mask >> p & 1u)) ? 0.0f : gt[p];
      }
    }
  };

  int s = 0;
  unsigned parity = 0;
  for (int i = 0; i < n_mine; ++i) {
    const int b = keep((int)(blockIdx.x + i * gridDim.x));
    const int par = i & 1;
    const unsigned st = keep(bar0 + 4u * (unsigned)L.stage0 + (unsigned)s * stage_bytes);
    const unsigned prm = keep(st + 4u * (unsigned)L.prm + 4u * (((unsigned)b * OAmod) & 3u));
    const unsigned pbase = keep(4u * (((unsigned)b * Pmod) & 3u) + 4u * (unsigned)tid);   // the thread's first pair in a [P] run
    const int sn = s + 1 == S ? 0 : s + 1;
    const unsigned pgp = pg_base + (unsigned)(i % kB3PG) * pg_bytes + pbase;   // the thread's first pair in the posterior ring
    if (stager && i + 1 < n_mine) caps3_bwd_tile_inputs(a, sv, up, L, smem, blockIdx.x + (i + 1) * gridDim.x, lane);
    mbar_wait(b3_full(bar0, s), parity);   // the image's in-place stage has landed

    // ---- winner-gradient variant: the pre-pass of THIS image (it needs every object's vote), then a barrier -------------
    float gsw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gswp = 0.0f, gwp = 0.0f, post_dummy = 0.0f;
    int widx = -1;
    if (kSoft) {
      const unsigned vmod6 = 4u * (((unsigned)b * V6mod) & 3u), vmod1 = 4u * (((unsigned)b * Vmod) & 3u);
      if (up.g_soft_winner) {
#pragma unroll
        for (int c = 0; c < 6; ++c) gsw[c] = lds_f32(st + 4u * (unsigned)L.gsw + vmod6 + 24u * (unsigned)v + 4 * c);
      }
      if (up.g_soft_winner_presence) gswp = lds_f32(st + 4u * (unsigned)L.gswp + vmod1 + 4u * (unsigned)v);
      if (up.g_winner_presence) gwp = lds_f32(st + 4u * (unsigned)L.gwp + vmod1 + 4u * (unsigned)v);
      if (up.g_winner || up.g_winner_presence)   // int64 indices: the low words
        widx = (int)lds_u32(st + 4u * (unsigned)L.widx + 4u * (((unsigned)b * ((unsigned)(2 * V) & 3u)) & 3u) + 8u * (unsigned)v);
      post_dummy = ex2_approx((2.0f * kDummyLog - lds_f32(st + 4u * (unsigned)L.lse + vmod1 + 4u * (unsigned)v)) * kLog2eF);
      mbar_wait(b3_ready(bar0, par), (unsigned)((i >> 1) & 1));
      mbar_wait(b3_pgfull(bar0, i % kB3PG), (unsigned)((i / kB3PG) & 1));
      float part = 0.0f;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (!(vmask >> j & 1u)) continue;
        float vt[6], vp;
        const int pj = tid + j * T;
        const C3PairConst<kSoft> cc = {cst_addr + (unsigned)j * 8u * T4, T4, a.cpr_static + (size_t)pj * 6, a.bias_vote + pj,
                                       a.bias_scale + pj};
        caps3_pair_vote<kSim>(prm + d_off + (unsigned)j * strideA, prm + l_off + (unsigned)j * strideA, cc,
                              st + r_off + (unsigned)j * strideR, st + L.nz4 + pbase + (unsigned)j * T4, deform,
                              a.noise_vote != nullptr, vt, vp);
        float h = have_gpost ? lds_f32(pgp + L.gpost4 + (unsigned)j * T4) : 0.0f;
#pragma unroll
        for (int c = 0; c < 6; ++c) h = fmaf(gsw[c], vt[c], h);
        h = fmaf(gswp, vp, h);
        part = fmaf(lds_f32(pgp + (unsigned)j * T4), h, part);
      }
      if (active) smem[L.SPART + par * T + tid] = part;
      named_bar_sync(kB3Bar, (unsigned)Tpad);
    }
    // ---- per-part values -------------------------------------------------------------------------------------------------
    float xv[6];
    {
      const unsigned xa = st + 4u * (unsigned)L.xs + 4u * (((unsigned)b * V6mod) & 3u) + 24u * (unsigned)v;
#pragma unroll
      for (int c = 0; c < 6; ++c) xv[c] = lds_f32(xa + 4 * c);
    }
    const float pres = a.presence ? lds_f32(st + 4u * (unsigned)L.ps + 4u * (((unsigned)b * Vmod) & 3u) + 4u * (unsigned)v) : 1.0f;
    float gllp = 0.0f, greg = 0.0f;   // upstream scalars of the image: read once the object tile is `ready`
    float Sv = 0.0f;   // S[v]: the G partials the pre-pass left (same order in every thread of the part)
    {
      const float* sp = smem + L.SPART + par * T + v;
      for (int kk = 0; kk < G; ++kk) Sv += sp[kk * V];
    }
    if (kSoft && up.g_soft_winner) {   // the dummy component's share: posterior(dummy) * <g_soft_winner, dummy_vote>
      float hd = 0.0f;
#pragma unroll
      for (int c = 0; c < 6; ++c) hd = fmaf(gsw[c], __ldg(a.dummy_vote + v * 6 + c), hd);
      Sv = fmaf(post_dummy, hd, Sv);
    }
    float gx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // the [7][P] tile is single-buffered: every thread must have finished the previous image's phase B
    if (i > 0) mbar_wait(b3_bdone(bar0, par ^ 1), (unsigned)(((i - 1) >> 1) & 1));

    // ---- main pass: the thread's pairs -------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (!(vmask >> j & 1u)) continue;
      const unsigned da = prm + d_off + (unsigned)j * strideA;   // row[6 v + c]
      const unsigned la = prm + l_off + (unsigned)j * strideA;   // row[6 V + 7 + v]; the scale slot 4 V bytes further
      const size_t e = (size_t)b * P + tid + (size_t)j * T;      // the pair in a (B,O,V) tensor
      const int pj = tid + j * T;
      const C3PairConst<kSoft> cc = {cst_addr + (unsigned)j * 8u * T4, T4, a.cpr_static + (size_t)pj * 6, a.bias_vote + pj,
                                     a.bias_scale + pj};
      float raw[6], t[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        raw[c] = lds_f32(da + 4 * c);
        t[c] = (deform ? raw[c] : 0.0f) + cc.get(c);
      }
      PoseAffine pa;
      pose_affine_mufu<kSim>(t, pa);
      if (j == 0) {
        mbar_wait(b3_ready(bar0, par), (unsigned)((i >> 1) & 1));   // the image's object tile and scalars are there
        const float gll = smem[L.SC + par * 4];
        greg = smem[L.SC + par * 4 + 1];
        gllp = gll * pres;
        if (k == 0 && out.g_presence)
          out.g_presence[(size_t)b * V + v] = gll * lds_f32(st + 4u * (unsigned)L.lse + 4u * (((unsigned)b * Vmod) & 3u) + 4u * (unsigned)v);
      }
      const float4 r0 = lds_f32x4(st + r_off + (unsigned)j * strideR), r1 = lds_f32x4(st + r_off + (unsigned)j * strideR + 16);
      const float4 os = lds_f32x4(st + os_off + (unsigned)j * strideO);
      const float r[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
      const float pc = r1.z;
      float vt[6];
      compose_vote(r, pa.a, vt);
      const float raw_lv = lds_f32(la), raw_u = lds_f32(la + V4);
      float lv = raw_lv + cc.get(6);
      if (a.noise_vote) lv += lds_f32(st + L.nz4 + pbase + (unsigned)j * T4);
      const float pv = sigmoid_fast(lv);
      const float vp = pc * pv;
      const float u05 = raw_u + cc.get(7);
      float sc = 1.0f, dsc = 0.0f;   // scale and d scale / d u
      if (learn) {
        const float z = ex2_approx(-fabsf(u05) * kLog2eF);   // softplus(x) = max(x, 0) + log1p(e^-|x|); its slope = sigmoid(x)
        sc = fmaxf(u05, 0.0f) + log1p_unit(z) + 1e-2f;
        const float rz = rcp_approx(1.0f + z);
        dsc = u05 >= 0.0f ? rz : z * rz;
      }
      float diff[6], q = 0.0f;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        diff[c] = xv[c] - vt[c];
        q = fmaf(diff[c], diff[c], q);
      }
      const float pst = lds_f32(pgp + (unsigned)j * T4);
      float h = have_gpost ? lds_f32(pgp + L.gpost4 + (unsigned)j * T4) : 0.0f;
      const bool is_win = kSoft && (k + G * j) == widx;
      if (kSoft) {
#pragma unroll
        for (int c = 0; c < 6; ++c) h = fmaf(gsw[c], vt[c], h);
        h = fmaf(gswp, vp, h);
      }
      const float g_pl = pst * (h - Sv) + gllp * pst;
      float g_vp = kSoft ? gswp * pst + (is_win ? gwp : 0.0f) : 0.0f;
      if (kExtras && up.g_vote_presence) g_vp += __ldg(up.g_vote_presence + e);
      if (__float_as_int(os.y) == v) g_vp += os.x;
      float g_ml = g_pl;
      if (kExtras && up.g_mixing_logit) g_ml += __ldg(up.g_mixing_logit + (size_t)b * (P + V) + tid + (size_t)j * T);
      if (!(vp < kLogSafeEps)) g_vp = fmaf(g_ml, rcp_approx(vp), g_vp);
      const float inv_sc = rcp_approx(sc);
      const float inv2 = inv_sc * inv_sc;
      const float coef = g_pl * inv2;
      float gv[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        gv[c] = coef * diff[c];
        if (kSoft) {
          gx[c] = fmaf(-coef, diff[c], gx[c]);
          gv[c] = fmaf(gsw[c], pst, gv[c]);
          if (is_win && up.g_winner) gv[c] += lds_f32(st + 4u * (unsigned)L.gw + 4u * (((unsigned)b * V6mod) & 3u) + 24u * (unsigned)v + 4 * c);
        }
        if (kExtras && up.g_vote) gv[c] += __ldg(up.g_vote + e * 6 + c);
      }
      float g_sc = g_pl * inv_sc * fmaf(q, inv2, -6.0f);
      if (kExtras && up.g_scale) g_sc += __ldg(up.g_scale + e);
      const float g_u = g_sc * dsc;
      float g_lv = g_vp * vp * (1.0f - pv);
      if (kExtras && up.g_presence_logit_per_vote) g_lv += __ldg(up.g_presence_logit_per_vote + e);
      // vote = R . A: gradient w.r.t. A (-> this pair's cpr parameters) and w.r.t. R (-> summed over the parts in phase B)
      const float* A_ = pa.a;
      float ga[6];
      ga[0] = r[0] * gv[0] + r[3] * gv[3];
      ga[1] = r[0] * gv[1] + r[3] * gv[4];
      ga[2] = r[0] * gv[2] + r[3] * gv[5];
      ga[3] = r[1] * gv[0] + r[4] * gv[3];
      ga[4] = r[1] * gv[1] + r[4] * gv[4];
      ga[5] = r[1] * gv[2] + r[4] * gv[5];
      const unsigned ra = red_addr + (unsigned)j * L.red_step, RP = L.red_plane;
      sts_f32(ra + 0 * RP, gv[0] * A_[0] + gv[1] * A_[1] + gv[2] * A_[2]);
      sts_f32(ra + 1 * RP, gv[0] * A_[3] + gv[1] * A_[4] + gv[2] * A_[5]);
      sts_f32(ra + 2 * RP, gv[2]);
      sts_f32(ra + 3 * RP, gv[3] * A_[0] + gv[4] * A_[1] + gv[5] * A_[2]);
      sts_f32(ra + 4 * RP, gv[3] * A_[3] + gv[4] * A_[4] + gv[5] * A_[5]);
      sts_f32(ra + 5 * RP, gv[5]);
      sts_f32(ra + 6 * RP, g_vp * pv);
      float gt[6];
      pose_affine_bwd<kSim>(ga, pa, gt);
      // gradient rows in place; the batch sums (-> cpr_static and bias gradients) take the pre-activation gradient
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        acc[j][c] += gt[c];
        float g = deform ? fmaf(greg, raw[c], gt[c]) : 0.0f;
        if (relu && !(raw[c] > 0.0f)) g = 0.0f;
        sts_f32(da + 4 * c, g);
      }
      acc[j][6] += g_lv;
      acc[j][7] += g_u;
      sts_f32(la, (relu && !(raw_lv > 0.0f)) ? 0.0f : g_lv);
      sts_f32(la + V4, (relu && !(raw_u > 0.0f)) ? 0.0f : g_u);
    }
    if (kSoft && out.g_x && active) {
#pragma unroll
      for (int c = 0; c < 6; ++c) smem[L.GX + c * T + tid] = gx[c];
    }
    fence_proxy_async();   // the gradient block is read by the bulk store issued after the barrier

    // ---- pre-pass of the next image ---------------------------------------------------------------------------------------
    if (!kSoft && i + 1 < n_mine) pre_pass(i + 1);
    if (tid == 0) bulk_wait_all();   // the previous image's bulk store (issued a whole pass ago) has completed
    named_bar_sync(kB3Bar, (unsigned)Tpad);   // the image's only barrier

    // ---- phase B ------------------------------------------------------------------------------------------------------------
    if (warp == 0) {   // this image's gradient block leaves
      const BulkRun r0 = bulk_run(a.all_param + (size_t)b * O * A, O * A);
      float* gdst = out.g_all_param + (size_t)b * O * A;   // congruent to the source modulo 16 bytes (checked on the host)
      float* gsrc = smem + L.stage0 + s * L.stage_stride + L.prm;
      if (lane == 0 && r0.body) {
        bulk_s2g(gdst + r0.head, gsrc + r0.off + r0.head, 4u * (unsigned)r0.body);
        bulk_commit();
      }
      bulk_run_edges_out(gdst, gsrc, r0, lane);
      // the posterior ring slot of this image is free (every thread finished its main pass): it takes image i + kB3PG
      if (i + kB3PG < n_mine)
        caps3_bwd_issue_pg(sv, up, L, smem, bar0, i % kB3PG, blockIdx.x + (i + kB3PG) * gridDim.x, P, lane);
    }
    // two threads per (object, slot), taken from the top warps down (warps 0 and 1 are busy above): each sums half of the
    // object's V entries of RED7[slot]; uniform trip count: the shuffle needs every lane
    for (int base = 0; base < 7 * O; base += Tpad >> 1) {
      const int it = base + ((Tpad - 1 - tid) >> 1), half = tid & 1;
      const bool on = it < 7 * O;
      const int c = on ? it / O : 0, oo = on ? it - c * O : 0;
      const int Vh = (V + 1) >> 1;
      const int v0 = half * Vh, v1 = min(V, v0 + Vh);
      float s0 = 0.0f, s1 = 0.0f;
      if (on) {
        const float* src = smem + L.RED7 + (c * O + oo) * L.Vp;
        int vv = v0;
        for (; vv + 2 <= v1; vv += 2) {
          s0 += src[vv];
          s1 += src[vv + 1];
        }
        if (vv < v1) s0 += src[vv];
      }
      float sum = s0 + s1;
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (on && half == 0) smem[L.OBJ7 + par * O * 8 + oo * 8 + c] = sum;
    }
    if (kSoft) {
      // g_x[v][c] = the object groups' shares summed; the dummy vote's gradient accumulates over the CTA's images
      const int skip = Tpad > 64 ? 64 : 0;   // warps 0 and 1 are busy with the store and the capsule-level chain
      for (int idx = tid - skip; idx >= 0 && idx < 6 * V; idx += Tpad - skip) {
        const int c = idx / V, vv = idx - c * V;
        if (out.g_x) {
          float sum = 0.0f;
          for (int kk = 0; kk < G; ++kk) sum += smem[L.GX + c * T + kk * V + vv];
          out.g_x[((size_t)b * V + vv) * 6 + c] = sum;
        }
        if (up.g_soft_winner) {
          const float lse_v = smem[L.stage0 + s * L.stage_stride + L.lse + (int)(((unsigned)b * Vmod) & 3u) + vv];
          const float g = smem[L.stage0 + s * L.stage_stride + L.gsw + (int)(((unsigned)b * V6mod) & 3u) + vv * 6 + c];
          smem[L.DSUM + vv * 6 + c] += g * ex2_approx((2.0f * kDummyLog - lse_v) * kLog2eF);
        }
      }
    }
    mbar_arrive(b3_bdone(bar0, par));   // this thread is done with the [7][O][Vp] tile
    if (warp == 0 && i + S < n_mine) {   // once the block has been read out of the stage, the stage takes image i + S
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
      caps3_bwd_issue(a, sv, up, kSoft, L, smem, bar0, s, blockIdx.x + (i + S) * gridDim.x, lane);
    }
    if (stager && i + 1 < n_mine) {
      // the next image's object tile, from the inputs copied during this pass; the other warps pick it up through `ready`
      // (this dependent chain runs beside the capsule-level chain of warp 1 instead of delaying the barrier)
      cp_async_wait<0>();
      caps3_bwd_object_tile<kSim>(a, up, L, smem, sn, i + 1, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(b3_ready(bar0, par ^ 1));
    }
    if (warp == chain_warp) {
      // the previous image's bulk store completed before this image's barrier (lane 0 of warp 0 waited for it), so its
      // 7 capsule-level slots per row can be overwritten now
      if (i > 0) object_chain(i - 1);
    }
    if (++s == S) {
      s = 0;
      parity ^= 1u;
    }
  }

  // ---- epilogue: the last image's capsule-level gradients, then the per-CTA batch sums ------------------------------------
  if (n_mine > 0) {
    if (tid == 0) bulk_wait_all();
    named_bar_sync(kB3Bar, (unsigned)Tpad);
    if (warp == chain_warp) object_chain(n_mine - 1);
  }
  __syncthreads();
  {
    float* dst = out.partials + (size_t)blockIdx.x * O * A;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (!(vmask >> j & 1u)) continue;
      float* row = dst + (size_t)(k + G * j) * A;
#pragma unroll
      for (int c = 0; c < 6; ++c) row[6 * v + c] = acc[j][c];
      row[6 * V + 7 + v] = acc[j][6];
      row[7 * V + 7 + v] = acc[j][7];
    }
    for (int idx = tid; idx < O * 7; idx += Tpad) {
      const int oo = idx / 7, c = idx - oo * 7;
      dst[(size_t)oo * A + 6 * V + c] = smem[L.OBJSUM + oo * 8 + c];
    }
    if (kSoft && out.dummy_partials)
      for (int idx = tid; idx < V * 6; idx += Tpad) out.dummy_partials[(size_t)blockIdx.x * V * 6 + idx] = smem[L.DSUM + idx];
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
// (tests/emu runs everything ABOVE this line on the CPU under a SIMT emulation: keep device code above, launches below)

static int b3_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

struct Caps3BwdPlan {
  int NP, threads;
  Caps3BwdLayout L;
  size_t smem;
};

// G object groups x V parts threads, NP = ceil(O / G) pairs each; 20 warps at most (96 registers per thread)
static bool caps3_plan_bwd(const scae_caps_args* a, const scae_caps_upstream* up, bool soft, bool want_gx,
                           Caps3BwdPlan* plan) {
  const int O = a->O, V = a->V;
  const int budget = max_smem_optin();
  const int force_np = b3_env_int("SCAE_CAPS3_BWD_NP", 0), force_s = b3_env_int("SCAE_CAPS3_BWD_STAGES", 0);
  double best_eff = 0.0;
  int best_T = 0;
  bool found = false;
  for (int NP : {2, 1, 4}) {
    if (force_np && NP != force_np) continue;
    const int G = (O + NP - 1) / NP;
    const int T = G * V, threads = (T + 31) & ~31;
    if (threads > 640) continue;
    const double eff = (double)O / ((double)G * NP);
    int S = force_s >= 2 && force_s <= kB3MaxStages ? force_s : 3;
    Caps3BwdLayout L = caps3_bwd_layout(a, up, G, NP, S, soft, want_gx);
    while (S > 2 && (size_t)L.total * sizeof(float) > (size_t)budget) L = caps3_bwd_layout(a, up, G, NP, --S, soft, want_gx);
    if ((size_t)L.total * sizeof(float) > (size_t)budget) continue;
    if (found && (eff < best_eff - 1e-9 || (eff < best_eff + 1e-9 && T <= best_T))) continue;
    plan->NP = NP, plan->threads = threads, plan->L = L, plan->smem = (size_t)L.total * sizeof(float);
    best_eff = eff, best_T = T;
    found = true;
  }
  return found;
}

static int caps3_bwd_grid(const scae_caps_args* a) {   // an upper bound: two CTAs per SM for the small shapes
  const int sms = 2 * sm_count();
  return a->B < sms ? a->B : sms;
}

size_t caps3_bwd_workspace_bytes(const scae_caps_args* a) {   // per-CTA partial rows: [O*A] and, behind them, [V*6]
  return (size_t)caps3_bwd_grid(a) * ((size_t)a->O * (8 * a->V + 7) + (size_t)a->V * 6) * sizeof(float);
}

int caps3_bwd(const scae_caps_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up, float* g_all_param,
              float* g_shared, float* g_dummy_vote, float* g_x, float* g_presence, void* workspace,
              size_t workspace_bytes, cudaStream_t stream, bool* handled) {
  *handled = false;
  if ((long)a->O * a->V >= (1L << 22)) return SCAE_OK;
  // the gradient of mixing_log_prob (never differentiated by SCAE) stays on the general path (caps_ll.cu)
  if (up->g_mixing_log_prob) return SCAE_OK;
  // upstream gradients of the soft / hard winner (vote_type / presence_type 'soft', 'hard') and the part-side input
  // gradient (stop_grad_caps_target = False): the winner-gradient variant of the kernel
  const bool soft = up->g_soft_winner || up->g_soft_winner_presence || up->g_winner || up->g_winner_presence || g_x;
  if ((up->g_winner || up->g_winner_presence) && !saved->winner_idx) return SCAE_OK;
  const void* need16[] = {a->all_param, a->cpr_static, a->x, g_all_param, saved->posterior_mixing_prob,
                          saved->log_prob_per_point};
  for (const void* p : need16)
    if (!aligned16(p)) return SCAE_OK;
  const void* opt16[] = {a->noise_vote, a->noise_caps, a->presence, up->g_posterior_mixing_prob, up->g_caps_presence,
                         up->g_presence_logit_per_caps, up->g_caps_presence ? saved->caps_presence_arg : nullptr,
                         up->g_soft_winner, up->g_soft_winner_presence, up->g_winner, up->g_winner_presence,
                         (up->g_winner || up->g_winner_presence) ? saved->winner_idx : nullptr};
  for (const void* p : opt16)
    if (p && !aligned16(p)) return SCAE_OK;
  Caps3BwdPlan plan;
  const bool extras = up->g_vote_presence || up->g_mixing_logit || up->g_vote || up->g_scale || up->g_presence_logit_per_vote;
  const bool general = soft || extras;   // the compiled variant with every upstream gradient: its tiles must exist
  if (!caps3_plan_bwd(a, up, general, g_x != nullptr, &plan)) return SCAE_OK;
  const int O = a->O, V = a->V, A = 8 * V + 7, n = O * A;
  if (workspace_bytes < caps3_bwd_workspace_bytes(a)) return SCAE_OK;
  const bool sim = (a->flags & SCAE_CAPS_SIMILARITY) != 0;
  void (*kern)(const scae_caps_args, const scae_caps_saved, const scae_caps_upstream, const Caps3BwdOut,
               const Caps3BwdLayout) = nullptr;
  // compiled variants: the training set of upstream gradients (fast), and everything else (extras + winner gradients)
#define B3_PICK(NP_, MAXT_, MINB_)                                                                        \
  (sim ? (extras || soft ? caps3_bwd_kernel<true, NP_, MAXT_, MINB_, true, true>                          \
                         : caps3_bwd_kernel<true, NP_, MAXT_, MINB_, false, false>)                       \
       : (extras || soft ? caps3_bwd_kernel<false, NP_, MAXT_, MINB_, true, true>                         \
                         : caps3_bwd_kernel<false, NP_, MAXT_, MINB_, false, false>))
  const bool two = !extras && !soft && plan.NP == 1 && plan.threads <= 416 &&
                   2 * (plan.smem + 1024) <= 228u * 1024u;   // two CTAs per SM
  if (two) kern = caps3_bwd_kernel<false, 1, 416, 2, false, false>;
  if (two && sim) kern = caps3_bwd_kernel<true, 1, 416, 2, false, false>;
  if (!two) {
    if (plan.NP == 1) kern = B3_PICK(1, 640, 1);
    else if (plan.NP == 2) kern = B3_PICK(2, 640, 1);
    else kern = B3_PICK(4, 512, 1);
  }
#undef B3_PICK
  if (plan.NP == 4 && plan.threads > 512) return SCAE_OK;
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int grid = sm_count() * (two ? 2 : 1);
  if (grid > a->B) grid = a->B;
  float* partials = static_cast<float*>(workspace);
  float* dummy_partials = (soft && up->g_soft_winner && g_dummy_vote) ? partials + (size_t)grid * n : nullptr;
  Caps3BwdOut out{g_all_param, g_presence, partials, g_x, dummy_partials};
  kern<<<grid, plan.threads, plan.smem, stream>>>(*a, *saved, *up, out, plan.L);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  int rc = launch_reduce_rows(partials, g_shared, grid, n, stream);
  if (rc != SCAE_OK) return rc;
  if (dummy_partials) {
    rc = launch_reduce_rows(dummy_partials, g_dummy_vote, grid, V * 6, stream);
    if (rc != SCAE_OK) return rc;
  } else if (g_dummy_vote) {
    SCAE_CUDA_TRY(cudaMemsetAsync(g_dummy_vote, 0, (size_t)V * 6 * sizeof(float), stream));
  }
  *handled = true;
  return SCAE_OK;
}

}  // namespace scae
