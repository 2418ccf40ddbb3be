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
// pass (profiles/r01a).  The shipped formulation (tmpl_ll_bwd_run_kernel below) is atomics-free and bit-reproducible: a
// warp owns one (image, template) pair, every lane walks runs of consecutive pixels and keeps the sums of the bilinear
// cell it is in in registers; finished cells go through a per-warp queue to the warp's private gradient atlas.
//
// History (profiles/): texel-parallel gather, every texel inverting the affine map and walking its own footprint --
// 1.5 ms at the MNIST config, divergent (r01b); segmented warp scan of the eight corner contributions on every 32-pixel
// pass -- 0.77 ms, 313 warp instructions per pass (r01, r02 first half); this kernel -- 0.53 ms, ~190 (r02).
#include <stdlib.h>
#include <string.h>

#include "tmpl_common.cuh"

namespace scae {

struct TmplBwdOut {
  float* g_templates;       // [B,M,C,h,w]; unused with template_color (see raw_partials)
  float* g_color;           // [B,M,C]           (template_color only)
  float* raw_partials;      // [grid][M*C*h*w]   (template_color only): per-CTA sums of colour * texel gradient
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
template <int C, bool kAlpha, bool kMode>
__device__ __forceinline__ void bwd_background(const scae_tmpl_args& a, const TmplScalars& sc, const float* xv,
                                               const float* G, const float* Nc, const float* Dc, size_t px0, int HW,
                                               float* g_bg_image, ScalarAcc& acc) {
  const float two_i2s = 2.0f * sc.i2s;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const size_t px = px0 + (size_t)c * HW;
    if (kMode) {   // backward of pdf.mode(): the pixel's gradient goes to the component it was taken from (Nc = its index)
      const float gl = Nc[c] == (float)a.M ? G[c] : 0.0f;
      if (a.bg_image) {
        if (g_bg_image) g_bg_image[px] = gl;
      } else {
        acc.bgval += gl;
      }
      continue;
    }
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

// One channel of the bilinear sample in difference form: value and its derivatives w.r.t. the two atlas coordinates
//   loc = t00 + fx DX + fy (DY + fx DXY),  d loc / d tx = DX + fy DXY,  d loc / d ty = DY + fx DXY
// with DX = t10 - t00, DY = t01 - t00, DXY = t11 - t10 - t01 + t00 (8 instructions instead of 12 for the weighted form
// plus its two difference quotients)
__device__ __forceinline__ float bilerp_d(float t00, float t10, float t01, float t11, float fx, float fy, float& dtx,
                                          float& dty) {
  const float DX = t10 - t00, DY = t01 - t00, DXY = (t11 - t01) - DX;
  dty = fmaf(fx, DXY, DY);
  dtx = fmaf(fy, DXY, DX);
  return fmaf(fy, dty, fmaf(fx, DX, t00));
}

// per-(pixel, template) gradients.  Returns g_loc[c] (c < C) and, in v[C], the summed logit gradient (alpha mode).
template <int C, bool kAlpha, int kPad, bool kMode>
__device__ __forceinline__ Texel<kPad> bwd_pixel(const TmplScalars& sc, const Texel<kPad>& t00, const Texel<kPad>& t10,
                                                 const Texel<kPad>& t01, const Texel<kPad>& t11, float fx, float fy,
                                                 float lpres, const float* xv, const float* G, const float* Nc,
                                                 const float* Dc, ScalarAcc& acc, float& glp, float& gtx, float& gty,
                                                 float m_index) {
  const float two_i2s = 2.0f * sc.i2s;
  Texel<kPad> out;
#pragma unroll
  for (int c = 0; c < kPad; ++c) out.v[c] = 0.0f;
  if (kMode) {
    // backward of pdf.mode() (distributions.py:50-77 without the straight-through estimator): d mode / d loc = 1 for the
    // component the arg-max picked (its index sits in Nc), nothing flows to the mixing logits
    glp = gtx = gty = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float gl = Nc[c] == m_index ? G[c] : 0.0f;
      out.v[c] = gl;
      float dtx, dty;
      bilerp_d(t00.v[c], t10.v[c], t01.v[c], t11.v[c], fx, fy, dtx, dty);
      gtx = fmaf(gl, dtx, gtx);
      gty = fmaf(gl, dty, gty);
    }
    return out;
  }
  float al = 0.0f, pD_shared = 0.0f, atx = 0.0f, aty = 0.0f;
  if (kAlpha) {
    al = bilerp_d(t00.v[C], t10.v[C], t01.v[C], t11.v[C], fx, fy, atx, aty) + lpres;
    pD_shared = ex2_ftz((al - Dc[0]) * kLog2e);
  }
  glp = 0.0f;
  gtx = 0.0f;
  gty = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float dtx, dty;
    const float loc = bilerp_d(t00.v[c], t10.v[c], t01.v[c], t11.v[c], fx, fy, dtx, dty);
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
    gtx = fmaf(gl, dtx, gtx);
    gty = fmaf(gl, dty, gty);
  }
  if (kAlpha) {
    out.v[C] = glp;
    gtx = fmaf(glp, atx, gtx);
    gty = fmaf(glp, aty, gty);
  }
  return out;
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
// run kernel: warp per (image, template); a lane walks runs of consecutive pixels and keeps the sums of its current
// bilinear cell in registers
// ================================================================================================================
//
// Along an image row the sampling coordinates move on a straight line, so the pixels that fall into one bilinear cell are
// CONSECUTIVE.  Every lane walks its own run of `L` consecutive pixels of one row, one pixel per step, and accumulates
// the cell's contributions in registers, in the moment basis {sum g, sum g fx, sum g fy, sum g fx fy} (4 x NCH
// registers; the four corner sums are linear combinations of these).  The cell's four texels stay in registers too: they
// are re-read only when the cell changes -- every `magnification` pixels, 7 at the MNIST configuration -- and the sample
// and its two coordinate derivatives come from the difference form (bilerp_d).
//
// When a lane's cell changes it appends {cell address, moments} to its warp's queue in shared memory (cells that touch
// only the zero border -- pixels that miss the template, most of the image for a small part -- are dropped: their
// gradient would be discarded).  A few lanes change cells in nearly every step, so writing to the gradient atlas right
// away would run the read-modify-write sequence at 10-20 % lane utilisation; the queue is drained 32 entries at a time,
// one entry per lane: entries of the SAME cell are found with one MATCH.ANY and take turns in queue order; within a turn
// all cells are distinct and the corners are updated one after the other with predicated vector read-modify-writes, so
// no two lanes touch an address in the same instruction and the summation order is fixed by the program:
// bit-reproducible, no atomics, no cross-lane scan.
//
// Mapping.  A row is cut into k runs (k a power of two, L = ceil(W / k)); a "walk" is 32 runs in lock-step: lane ->
// (row slot q = lane / k, segment lane % k).  Walk w takes the rows q * walks + w, so the rows a warp works on at the same
// time are `walks` pixels apart and rarely share a cell, and a lane's successive runs are vertically adjacent; odd
// walks run right to left (boustrophedon), so the next run starts next to the pixel the last one ended on -- usually in
// the same cell.  Pixel records {x, upstream gradient, the two cached lse terms} are staged per band of walks in the
// order the lanes read them ([walk][lane][step], skewed so that one LDS.128 per step is bank-conflict-free); ragged ends
// and dead runs hold pad records {0, 0, 1e30, 1e30} that turn every contribution into an exact zero without a select in
// the loop.  Images whose records do not fit next to the atlases are processed band by band (two CTA barriers per band),
// the cell state and pose sums carried in registers across bands.  tmpl_bwd_plan searches run geometry, warps per CTA
// and band size for the best occupancy x lane-efficiency product.
constexpr int kRunThreads = 256;

// [addr] += the NCH live channels of one corner where `pred` holds: one predicated vector read-modify-write of the padded
// texel (values, not a pointer: the sums must stay in registers across the asm statements)
template <int kPad>
__device__ __forceinline__ void texel_add_pred(unsigned addr, float v0, float v1, float v2, float v3, bool pred) {
  if (kPad == 1) smem_add_pred_f32(addr, v0, pred);
  else if (kPad == 2) smem_add_pred_f32x2(addr, v0, v1, pred);
  else smem_add_pred_f32x4(addr, v0, v1, v2, v3, pred);
}

// ---- per-warp queue of finished cells ----------------------------------------------------------------------------------
// A lane that leaves a cell appends {cell address, the cell's four moment sums} to its warp's queue in shared memory (a ring of
// kQueueCap entries: keys [cap] u32, sums [cap][kPad] float4).  Cells change for a few lanes in nearly every step, so
// handing them to the gradient atlas right away would run the read-modify-write sequence at 10-20 % lane utilisation;
// the queue is drained 32 entries at a time instead, one entry per lane.
constexpr int kQueueCap = 64;          // >= 2 x 32: an append of up to 32 entries always fits after a drain
constexpr int kQueueDrainAt = kQueueCap - 32;

struct CellQueue {
  unsigned keys, sums;    // shared-memory byte addresses of the two arrays
  unsigned head, tail;    // monotonic entry counters (warp-uniform); slot = counter & (kQueueCap - 1)
};

// lanes with `on` set append (cell, sums); `lane_lt` = mask of the lower lanes
template <int kPad, int NCH>
__device__ __forceinline__ void queue_append(CellQueue& q, unsigned on_mask, bool on, unsigned lane_lt, unsigned cell,
                                             const float (&acc)[4][NCH]) {
  const unsigned slot = (q.tail + (unsigned)__popc(on_mask & lane_lt)) & (unsigned)(kQueueCap - 1);
  sts_pred_u32(q.keys + slot * 4u, cell, on);
  const unsigned va = q.sums + slot * (unsigned)(16 * kPad);
  if (kPad == 1) {
    sts_pred_f32x4(va, acc[0][0], acc[1][0], acc[2][0], acc[3][0], on);
  } else if (kPad == 2) {
    sts_pred_f32x4(va, acc[0][0], acc[0][NCH > 1 ? 1 : 0], acc[1][0], acc[1][NCH > 1 ? 1 : 0], on);
    sts_pred_f32x4(va + 16u, acc[2][0], acc[2][NCH > 1 ? 1 : 0], acc[3][0], acc[3][NCH > 1 ? 1 : 0], on);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      sts_pred_f32x4(va + 16u * k, acc[k][0], acc[k][NCH > 1 ? 1 : 0], acc[k][NCH > 2 ? 2 : 0],
                     NCH > 3 ? acc[k][NCH > 3 ? 3 : 0] : 0.0f, on);
  }
  q.tail += (unsigned)__popc(on_mask);
}

// The (up to) 32 oldest entries, one per lane, go to the warp's gradient atlas (shared address cell + gat).  Entries of
// the same cell take turns in queue order; within a turn all cells are distinct and the corners are updated one after
// the other, so no two lanes touch an address in the same instruction and the summation order is fixed.
template <int kPad>
__device__ __forceinline__ void queue_drain(CellQueue& q, unsigned gat, unsigned row, int lane, unsigned lane_lt) {
  __syncwarp();                                   // the appended entries are visible to the whole warp
  const unsigned cnt = q.tail - q.head, n = cnt < 32u ? cnt : 32u;
  const bool have = (unsigned)lane < n;
  const unsigned slot = (q.head + (unsigned)lane) & (unsigned)(kQueueCap - 1);
  const unsigned cell = have ? lds_u32(q.keys + slot * 4u) : (0xFFFFFF00u | (unsigned)lane);
  float e[4 * kPad];
#pragma unroll
  for (int p4 = 0; p4 < kPad; ++p4) {
    const float4 v = lds_f32x4(q.sums + slot * (unsigned)(16 * kPad) + 16u * p4);
    e[4 * p4 + 0] = v.x, e[4 * p4 + 1] = v.y, e[4 * p4 + 2] = v.z, e[4 * p4 + 3] = v.w;
  }
  // {S, Sx, Sy, Sxy} -> corner sums: se = Sxy, ne = Sx - Sxy, sw = Sy - Sxy, nw = (S - Sx) - sw   (weights (1-fx)(1-fy),
  // fx (1-fy), (1-fx) fy, fx fy summed over the cell's pixels)
#pragma unroll
  for (int c = 0; c < kPad; ++c) {
    const float S = e[c], Sx = e[kPad + c], Sy = e[2 * kPad + c], Sxy = e[3 * kPad + c];
    const float sw = Sy - Sxy;
    e[c] = (S - Sx) - sw;
    e[kPad + c] = Sx - Sxy;
    e[2 * kPad + c] = sw;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, cell);
  const unsigned rank = (unsigned)__popc(peers & lane_lt);
  const unsigned turns = redux_max_u32(0xffffffffu, have ? rank + 1u : 0u);
  const unsigned ga = cell + gat, gb = ga + row;
#pragma unroll 1
  for (unsigned r = 0; r < turns; ++r) {
    const bool mine = have && rank == r;
    texel_add_pred<kPad>(ga, e[0], e[kPad > 1 ? 1 : 0], e[kPad > 2 ? 2 : 0], e[kPad > 2 ? 3 : 0], mine);
    __syncwarp();
    texel_add_pred<kPad>(ga + kPad * 4, e[kPad], e[kPad + (kPad > 1 ? 1 : 0)], e[kPad + (kPad > 2 ? 2 : 0)],
                         e[kPad + (kPad > 2 ? 3 : 0)], mine);
    __syncwarp();
    texel_add_pred<kPad>(gb, e[2 * kPad], e[2 * kPad + (kPad > 1 ? 1 : 0)], e[2 * kPad + (kPad > 2 ? 2 : 0)],
                         e[2 * kPad + (kPad > 2 ? 3 : 0)], mine);
    __syncwarp();
    texel_add_pred<kPad>(gb + kPad * 4, e[3 * kPad], e[3 * kPad + (kPad > 1 ? 1 : 0)], e[3 * kPad + (kPad > 2 ? 2 : 0)],
                         e[3 * kPad + (kPad > 2 ? 3 : 0)], mine);
    __syncwarp();
  }
  q.head += n;
}

// Work unit = (image b, template group grp): the CTA's warps take the templates grp * nwarps + warp.  A batch of 1024
// images is 5120 units; the persistent CTAs take contiguous chunks of them (11.5 each over 444 CTAs: chunk boundaries
// fall inside images, so the last wave's tail stays a few per cent of the kernel instead of the 14 % whole images gave).
// kMode: the same scatter for the backward of pdf.mode() -- `gout` is the gradient w.r.t. the mode image and `cache`'s
// first C planes hold the index of the component each pixel took its value from (scae_tmpl_mode_bwd)
template <int C, bool kAlpha, bool kMode, int kOcc>
__global__ void __launch_bounds__(kRunThreads, kOcc) tmpl_ll_bwd_run_kernel(const scae_tmpl_args a,
                                                                                   const float* __restrict__ x,
                                                                                   const float* __restrict__ gout,
                                                                                   const float* __restrict__ cache,
                                                                                   const TmplBwdOut out, const TmplGeom g) {
  using TT = TexTraits<C, kAlpha>;
  constexpr int kPad = TT::kPad, NCH = TT::kCh;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int atlas_floats = g.atlas_floats;               // one padded template, multiple of 4 floats
  constexpr int kQueueFloats = kQueueCap * (1 + 4 * kPad);   // keys + sums of one warp's queue
  const int warp_floats = 2 * atlas_floats + kQueueFloats;
  float* atlas = smem + (size_t)warp * warp_floats;       // this warp's value atlas
  float* gatlas = atlas + atlas_floats;                   // ... its gradient atlas
  float* red = smem + (size_t)nwarps * warp_floats;       // [64] block-reduction scratch
  // run geometry (tmpl_bwd_plan): k runs of L pixels per row, `walks` walks per image, `bw` walks per staged band
  const int L = g.tw, skew_shift = g.ppt, kshift = g.k, kruns = 1 << kshift, walks = g.tiles_y, bw = g.tiles_x;
  float* xs = red + 64;                                   // [kruns * L] affine_grid base coordinates (0 past W)
  float* ys = xs + kruns * L;                             // [H]
  const int HW = a.H * a.W, hw = a.h * a.w, pw = g.pw, W = a.W, H = a.H;
  // records {x, upstream gradient, cached numerator lse, cached denominator lse} of the current band, [C][bw][32 L + 8]:
  // run `ln` of a walk starts at ln * L + (ln >> skew_shift); the skew spreads the eight lanes of a quarter-warp over the
  // eight 16-byte bank groups whatever the parity of L (one LDS.128 per step, conflict-free)
  float4* PIX =
      reinterpret_cast<float4*>(smem + (((size_t)nwarps * warp_floats + 64 + kruns * L + H + 3) & ~(size_t)3));
  const int walk_recs = 32 * L + (31 >> skew_shift) + 1, plane = bw * walk_recs;
  #pragma unroll 1
  for (int e = lane; e < 2 * atlas_floats; e += 32) atlas[e] = 0.0f;
  #pragma unroll 1
  for (int e = threadIdx.x; e < kruns * L; e += blockDim.x) xs[e] = e < W ? base_coord(e, W) : 0.0f;
  #pragma unroll 1
  for (int e = threadIdx.x; e < H; e += blockDim.x) ys[e] = base_coord(e, H);
  const TmplScalars sc = tmpl_scalars(a);
  const float lim_x = keep((float)a.w + 2.5f), lim_y = keep((float)a.h + 2.5f);
  const float in_x = keep((float)a.w + 2.0f), in_y = keep((float)a.h + 2.0f);   // interior cells: 1 <= t < size + 2
  // the hot loop addresses shared memory by 32-bit byte address: tap offsets carry the warp's atlas address, so a tap is
  // one LDS [reg + imm] and a cell's identity (the flush key) is the address of its north-west texel
  const unsigned sbase = smem_u32(smem);
  const unsigned row = keep((unsigned)(pw * kPad * 4));
  const unsigned atlas_off = sbase + (unsigned)(warp * warp_floats) * 4u;
  CellQueue queue;                                       // ... and its cell queue: sums (16-byte aligned), then keys
  queue.sums = keep(atlas_off + (unsigned)(2 * atlas_floats) * 4u);
  queue.keys = queue.sums + (unsigned)(kQueueCap * 16 * kPad);
  queue.head = queue.tail = 0u;
  const unsigned lane_lt = keep((1u << lane) - 1u);
  const unsigned base0 = keep(atlas_off - kMagicBits * (row + (unsigned)(kPad * 4)));
  const unsigned gat = keep((unsigned)atlas_floats * 4u);
  const unsigned pix_addr = smem_u32(PIX), xs_addr = smem_u32(xs);
  const float hw_x = 0.5f * (float)a.w, hw_y = 0.5f * (float)a.h;
  const float inv_L = 1.0f / (float)L;
  float* my_alpha_partial = (kAlpha && out.alpha_partials) ? out.alpha_partials + (size_t)blockIdx.x * a.M * hw : nullptr;
  if (my_alpha_partial)
    #pragma unroll 1
    for (int e = threadIdx.x; e < a.M * hw; e += blockDim.x) my_alpha_partial[e] = 0.0f;
  // fused colourisation: the batch-shared raw templates get a per-CTA partial row like the alpha logits
  const bool colored = a.template_color != nullptr;
  float* my_raw_partial = colored ? out.raw_partials + (size_t)blockIdx.x * a.M * C * hw : nullptr;
  if (my_raw_partial)
    #pragma unroll 1
    for (int e = threadIdx.x; e < a.M * C * hw; e += blockDim.x) my_raw_partial[e] = 0.0f;
  ScalarAcc acc;
  __syncthreads();   // partial row zeroed before any warp accumulates into it; xs / ys complete
  const int q_lane = lane >> kshift, col0 = (lane & (kruns - 1)) * L;
  // per-lane constants of the walks: first row, record offset inside a walk, X addresses of a forward / backward run
  const int row0 = q_lane * walks;
  const unsigned rec_lane = keep((unsigned)(lane * L + (lane >> skew_shift)) * 16u);
  const unsigned walk_bytes = (unsigned)walk_recs * 16u, ys_addr = smem_u32(ys);
  const unsigned xa_fwd = keep(xs_addr + (unsigned)col0 * 4u), xa_bwd = xa_fwd + (unsigned)(L - 1) * 4u;
  const int nbands = (walks + bw - 1) / bw;

  // Every CTA takes a CONTIGUOUS chunk of the units: consecutive units are the template groups of one image, so the
  // image's pixel records are staged once per image (not once per group) and the groups in between need no CTA
  // barrier at all -- a warp that finishes its template early starts the next group's while the others catch up.
  const int groups = g.groups, n_units = a.B * groups;
  const int u_begin = (int)((long)blockIdx.x * n_units / gridDim.x), u_end = (int)((long)(blockIdx.x + 1) * n_units / gridDim.x);
  int b_staged = -1;               // image whose records are in shared memory (whole-image staging only)
  #pragma unroll 1
  for (int u = u_begin; u < u_end; ++u) {
    const int b = u / groups, grp = u - b * groups;
    const bool first = grp == 0;   // the background component and g_bg_image belong to the image, not to a template
    const bool restage = nbands > 1 || b != b_staged;
    b_staged = b;
    const int m = grp * nwarps + warp;
    const bool has_m = m < a.M;    // (no early exit: every warp takes part in the band barriers)
    // ---- per-template setup (all lanes compute the same coefficients) -------------------------------------------
    const int mm = has_m ? m : 0;
    const float* pp = a.pose + ((size_t)b * a.M + mm) * 6;
    const float Ax = __ldg(pp + 0) * hw_x, Bx = __ldg(pp + 1) * hw_x, Cx = (__ldg(pp + 2) + 1.0f) * hw_x + 1.5f;
    const float Ay = __ldg(pp + 3) * hw_y, By = __ldg(pp + 4) * hw_y, Cy = (__ldg(pp + 5) + 1.0f) * hw_y + 1.5f;
    const float pres = a.presence ? __ldg(a.presence + (size_t)b * a.M + mm) : 1.0f;
    const float lpres = a.presence ? log_safe_f(pres) : 0.0f;
    const float* src = a.templates + ((colored ? (size_t)0 : (size_t)b * a.M) + mm) * C * hw;
    float colv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) colv[c] = colored ? __ldg(a.template_color + ((size_t)b * a.M + mm) * C + c) : 1.0f;
    if (has_m) {
      const float inv_w = 1.0f / (float)a.w;
      #pragma unroll 4
      for (int e = lane; e < hw; e += 32) {
        const int y = (int)(((float)e + 0.5f) * inv_w), xx = e - y * a.w;
        float* q = atlas + ((size_t)(y + 2) * pw + (xx + 2)) * kPad;
#pragma unroll
        for (int c = 0; c < C; ++c) q[c] = __ldg(src + (size_t)c * hw + e) * colv[c];
        if (kAlpha) q[C] = __ldg(a.templates_alpha + (size_t)mm * hw + e);
      }
    }
    float sgx = 0.f, sgxX = 0.f, sgxY = 0.f, sgy = 0.f, sgyX = 0.f, sgyY = 0.f, spres = 0.f;
    // the lane's current cell (starts in the border corner: zero texels, gradient discarded): its four texels and its
    // four moment sums
    unsigned cur_off = atlas_off;
    unsigned cur_in = 0u;      // the current cell touches an interior texel
    Texel<kPad> t00, t10, t01, t11;
#pragma unroll
    for (int c = 0; c < kPad; ++c) t00.v[c] = t10.v[c] = t01.v[c] = t11.v[c] = 0.0f;
    float cs[4][NCH];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < NCH; ++c) cs[k][c] = 0.0f;
    queue.head = queue.tail = 0u;

    const float* x_img = x + (size_t)b * C * HW;
    const float* g_img = gout + (size_t)b * C * HW;
    const float* c_img_ptr = cache + (size_t)b * 2 * C * HW;
    for (int band = 0; band < nbands; ++band) {
      // ---- pixel records of the band (whole CTA) and, once per image, the background component -------------------
      const int w0 = band * bw;
      if (restage) {
      __syncthreads();                       // the previous band's / image's records are no longer read
      #pragma unroll 2
      for (int e = threadIdx.x; e < bw * 32 * L; e += blockDim.x) {
        const int slot = (int)(((float)e + 0.5f) * inv_L), s = e - slot * L;
        const int ln = slot & 31, wk = slot >> 5;
        // walk w takes the rows q * walks + w; odd walks run right to left, so a lane's next run starts next to the pixel
        // its last one ended on (usually the same cell: no cell change at the turn)
        const int r_img = (ln >> kshift) * walks + w0 + wk;
        const int c_img = (ln & (kruns - 1)) * L + (((w0 + wk) & 1) ? L - 1 - s : s);
        const bool ok = r_img < H && c_img < W && w0 + wk < walks;
        const int p = r_img * W + c_img;
        float xv[C], G[C], Nc[C], Dc[C];
        const size_t px0 = (size_t)b * C * HW + p;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int pc = c * HW + p;          // 32-bit offsets from the image's base pointers: one IMAD.WIDE per load
          xv[c] = ok ? __ldg(x_img + pc) : 0.0f;
          G[c] = ok ? __ldg(g_img + pc) : 0.0f;
          Nc[c] = ok ? __ldg(c_img_ptr + pc) : 1e30f;
          Dc[c] = ok ? __ldg(c_img_ptr + C * HW + pc) : 1e30f;
          PIX[c * plane + wk * walk_recs + ln * L + (ln >> skew_shift) + s] = make_float4(xv[c], G[c], Nc[c], Dc[c]);
        }
        if (first && ok) bwd_background<C, kAlpha, kMode>(a, sc, xv, G, Nc, Dc, px0, HW, out.g_bg_image, acc);
      }
      __syncthreads();
      }
      if (!has_m) continue;
      // ---- walks of the band: 32 runs in lock-step, one pixel per lane per step -----------------------------------
      #pragma unroll 1
      for (int wk = 0; wk < bw && w0 + wk < walks; ++wk) {
        const int w = w0 + wk, r_img = row0 + w;
        const float Y = lds_f32(ys_addr + 4u * (unsigned)(r_img < H ? r_img : H - 1));   // (a dead run reads pad records)
        const float yx = fmaf(Y, Bx, Cx), yy = fmaf(Y, By, Cy);
        const bool backward = (w & 1) != 0;
        const unsigned xstep = keep(backward ? 0u - 4u : 4u);
        unsigned xa = keep(backward ? xa_bwd : xa_fwd);                     // loop-carried addresses: X of the lane's
        unsigned ra = keep(pix_addr + (unsigned)wk * walk_bytes + rec_lane);   // column, its pixel records
        float wgx = 0.f, wgy = 0.f;
        #pragma unroll 1
        for (int s = 0; s < L; ++s, xa += xstep, ra += 16u) {
          float xv[C], G[C], Nc[C], Dc[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float4 r4 = lds_f32x4(ra + (unsigned)(c * plane) * 16u);
            xv[c] = r4.x;
            G[c] = r4.y;
            Nc[c] = r4.z;
            Dc[c] = r4.w;
          }
          const float X = lds_f32(xa);
          Tap t;
          tap_setup<kPad, 4>(fmaf(X, Ax, yx), fmaf(X, Ay, yy), lim_x, lim_y, row, base0, t);
          // ---- the cell changed: queue the finished sums, fetch the new cell's texels; drain the queue when 32 entries
          //      have gathered -----------------------------------------------------------------------------------------
          const bool changed = t.off != cur_off;
          const unsigned chm = __ballot_sync(0xffffffffu, changed);
          if (chm != 0u) {
            // cells that touch no interior texel (the zero border: pixels that miss the template) are not queued --
            // their gradient would be discarded; parts cover a fraction of the image, so this is most cells
            const bool app = changed && cur_in != 0u;
            const unsigned am = __ballot_sync(0xffffffffu, app);
            if (am != 0u) queue_append<kPad, NCH>(queue, am, app, lane_lt, cur_off, cs);
            if (changed) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int c = 0; c < NCH; ++c) cs[k][c] = 0.0f;
              cur_off = t.off;
              cur_in = (t.tx >= 1.0f && t.tx < in_x && t.ty >= 1.0f && t.ty < in_y) ? 1u : 0u;
            }
            lds_texel_pred<kPad>(t.off, t00, changed);
            lds_texel_pred<kPad>(t.off + kPad * 4, t10, changed);
            lds_texel_pred<kPad>(t.off + row, t01, changed);
            lds_texel_pred<kPad>(t.off + row + kPad * 4, t11, changed);
            if (queue.tail - queue.head >= (unsigned)kQueueDrainAt) queue_drain<kPad>(queue, gat, row, lane, lane_lt);
          }
          float glp, gtx, gty;
          const Texel<kPad> gv =
              bwd_pixel<C, kAlpha, kPad, kMode>(sc, t00, t10, t01, t11, t.fx, t.fy, lpres, xv, G, Nc, Dc, acc, glp, gtx, gty, (float)m);
          wgx += gtx;
          sgxX = fmaf(gtx, X, sgxX);
          wgy += gty;
          sgyX = fmaf(gty, X, sgyX);
          spres += glp;
          // the cell's sums in the basis {1, fx, fy, fx fy}; the drain turns them into the four corner sums
          const float fxy = t.fx * t.fy;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            cs[0][c] += gv.v[c];
            cs[1][c] = fmaf(gv.v[c], t.fx, cs[1][c]);
            cs[2][c] = fmaf(gv.v[c], t.fy, cs[2][c]);
            cs[3][c] = fmaf(gv.v[c], fxy, cs[3][c]);
          }
        }
        sgx += wgx;                  // Y is constant along a run
        sgxY = fmaf(wgx, Y, sgxY);
        sgy += wgy;
        sgyY = fmaf(wgy, Y, sgyY);
      }
    }
    if (!has_m) continue;
    {
      const unsigned am = __ballot_sync(0xffffffffu, cur_in != 0u);
      if (am != 0u) queue_append<kPad, NCH>(queue, am, cur_in != 0u, lane_lt, cur_off, cs);
    }
    while (queue.tail != queue.head) queue_drain<kPad>(queue, gat, row, lane, lane_lt);
    __syncwarp();

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
      float* dst = colored ? my_raw_partial + (size_t)m * C * hw : out.g_templates + ((size_t)b * a.M + m) * C * hw;
      float gcol[C];
#pragma unroll
      for (int c = 0; c < C; ++c) gcol[c] = 0.0f;
      // interior texels only: the border of the gradient atlas collects what the cells on the template's edge spill
      // over it, is never read and therefore never cleared
      const float inv_w = 1.0f / (float)a.w;
#pragma unroll 4
      for (int te = lane; te < hw; te += 32) {
        const int yy = (int)(((float)te + 0.5f) * inv_w), xx = te - yy * a.w;
        float* q = gatlas + ((size_t)(yy + 2) * pw + (xx + 2)) * kPad;
        if (colored) {
          // template = raw * colour: d/d raw = colour * g (summed over the batch), d/d colour = sum_texels raw * g
#pragma unroll
          for (int c = 0; c < C; ++c) {
            dst[(size_t)c * hw + te] += colv[c] * q[c];
            gcol[c] = fmaf(__ldg(src + (size_t)c * hw + te), q[c], gcol[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) dst[(size_t)c * hw + te] = q[c];
        }
        if (kAlpha && my_alpha_partial) my_alpha_partial[(size_t)m * hw + te] += q[C];
#pragma unroll
        for (int c = 0; c < kPad; ++c) q[c] = 0.0f;
      }
      if (colored) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = warp_sum(gcol[c]);
          if (lane == 0) out.g_color[((size_t)b * a.M + m) * C + c] = t;
        }
      }
    }
    __syncwarp();
  }
  write_scalar_partials(a, sc, kAlpha, acc, red, out.scalar_partials + (size_t)blockIdx.x * 4);
}

// ================================================================================================================
// host
// ================================================================================================================
// CTAs per SM the kernel is compiled for (register budget 64 / 80 per thread)
// CTAs of 256 threads per SM the kernel is compiled for: __launch_bounds__(256, 3) = 80 registers for one-channel
// images, (256, 2) = 128 registers for colour (12-16 corner sums + 16 cached texel values per lane)
static int tmpl_bwd_occ(int C) { return C == 1 ? 3 : 2; }

// Picks the run geometry (k = 2^kshift runs of L pixels per row, 32 / k rows per walk), the warps per CTA and the band of
// walks whose pixel records are staged at a time.  Every candidate is scored by
//   occupancy (warps per SM, saturating at 24)  x  lane efficiency of the walks  x  L / (L + 1) (every run ends with a
//   cell change)  x  a mild penalty per extra band (two CTA barriers each) and per halving of the warps (the records
//   are staged once per group of `warps` templates)
static int tmpl_bwd_plan(const scae_tmpl_args* a, TmplGeom* gp) {
  const int kpad = tmpl_texel_floats(a);
  const size_t limit = (size_t)max_smem_optin();
  const size_t per_sm = limit + 1024;               // shared memory of one SM; every resident CTA also pays 1 KB
  TmplGeom& g = *gp;
  memset(&g, 0, sizeof(g));
  g.pw = a->w + 4;
  g.ph = a->h + 4;
  g.atlas_floats = (int)(((size_t)g.pw * g.ph * kpad + 3) / 4 * 4);
  const int cap = tmpl_bwd_occ(a->C);
  const int regs = cap == 3 ? 80 : 128;
  double best = -1.0;
  int best_per = 1;
  for (int threads = kRunThreads; threads >= 32; threads /= 2) {
    const int warps = threads / 32;
    int by_regs = 65536 / (regs * threads), by_threads = 2048 / threads;
    const int occ_max = by_regs < by_threads ? by_regs : by_threads;
    for (int ks = 0; ks <= 5; ++ks) {
      const int k = 1 << ks, L = (a->W + k - 1) / k, rows = 32 / k, walks = (a->H + rows - 1) / rows;
      if (ks > 0 && L < 2) break;
      // one padded value atlas + one gradient atlas + one cell queue per warp, block scratch, coordinate tables
      const size_t fixed =
          ((size_t)warps * (2 * g.atlas_floats + kQueueCap * (1 + 4 * kpad)) + 64 + (size_t)k * L + a->H + 8) * sizeof(float);
      // run ln of a walk starts at record ln * L + (ln >> sh): with 2^tz | L, sh = 3 - tz makes the eight lanes of a
      // quarter-warp hit eight different 16-byte bank groups (odd L: no skew needed)
      int tz = 0;
      while (tz < 3 && ((L >> tz) & 1) == 0) ++tz;
      const int sh = tz == 0 ? 5 : 3 - tz;
      const size_t walk_bytes = (size_t)16 * a->C * (32 * L + (31 >> sh) + 1);
      for (int occ = occ_max; occ >= 1; --occ) {
        size_t budget = per_sm / occ - 1024;
        if (budget > limit) budget = limit;
        if (budget < fixed + 16 + walk_bytes) continue;
        int bw = (int)((budget - fixed - 16) / walk_bytes);
        if (bw > walks) bw = walks;
        const int nbands = (walks + bw - 1) / bw;
        bw = (walks + nbands - 1) / nbands;          // even bands
        const double w = (double)occ * warps;
        const double f_occ = w >= 24.0 ? 1.0 : sqrt(w / 24.0);
        const double eff = (double)a->H * a->W / ((double)walks * 32 * L);
        const double score =
            f_occ * eff * (L / (L + 1.0)) / (1.0 + 0.02 * (nbands - 1)) / (1.0 + 0.5 / warps);
        if (score > best * 1.0001) {
          best = score;
          best_per = occ;
          g.k = ks;
          g.tw = L;
          g.tiles_y = walks;
          g.ppt = sh;
          g.tiles_x = bw;
          g.pix_floats = (int)(walk_bytes / 4 * bw);
          g.threads = threads;
          g.smem_bytes = fixed + 16 + walk_bytes * bw;
        }
        break;                                       // lower occupancies of the same geometry only score worse
      }
    }
  }
  SCAE_REQUIRE(g.threads != 0, SCAE_ELIMIT, "tmpl bwd: a %dx%d template with a %dx%d image does not fit in shared memory",
               a->h, a->w, a->H, a->W);
  const int warps = g.threads / 32;
  // work units: (image, group of `warps` templates), one template per warp
  g.mc = 1;
  g.groups = (a->M + warps - 1) / warps;
  const long slots = (long)sm_count() * best_per;
  const long units = (long)a->B * g.groups;
  g.grid = units < slots ? (int)units : (int)slots;
  return SCAE_OK;
}

static size_t tmpl_ws_alpha_floats(const scae_tmpl_args* a, int grid) {
  return a->mode == SCAE_TMPL_MODE_ALPHA ? (size_t)grid * a->M * a->h * a->w : 0;
}
static size_t tmpl_ws_raw_floats(const scae_tmpl_args* a, int grid) {
  return a->template_color ? (size_t)grid * a->M * a->C * a->h * a->w : 0;
}

}  // namespace scae

using namespace scae;

extern "C" __attribute__((visibility("default"))) size_t scae_tmpl_ll_bwd_workspace_bytes(const scae_tmpl_args* a) {
  if (tmpl_validate(a) != SCAE_OK) return 0;
  TmplGeom g;
  if (tmpl_bwd_plan(a, &g) != SCAE_OK) return 0;
  return (tmpl_ws_alpha_floats(a, g.grid) + tmpl_ws_raw_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int scae_tmpl_ll_bwd(
    const scae_tmpl_args* a, const float* x, const float* grad_log_prob, const float* cache, float* g_templates,
    float* g_color, float* g_pose, float* g_presence, float* g_bg_image, float* g_alpha, float* g_scalars,
    void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(x && grad_log_prob && cache && g_templates && g_pose && g_scalars, SCAE_EINVAL,
               "tmpl bwd: a required pointer is NULL");
  SCAE_REQUIRE(!a->template_color || g_color, SCAE_EINVAL, "tmpl bwd: g_color is required with template_color");
  TmplGeom g;
  rc = tmpl_bwd_plan(a, &g);
  if (rc != SCAE_OK) return rc;
  const size_t need =
      (tmpl_ws_alpha_floats(a, g.grid) + tmpl_ws_raw_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
  SCAE_REQUIRE(workspace && workspace_bytes >= need, SCAE_EINVAL, "tmpl bwd: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  float* alpha_partials = static_cast<float*>(workspace);
  float* raw_partials = alpha_partials + tmpl_ws_alpha_floats(a, g.grid);
  float* scalar_partials = raw_partials + tmpl_ws_raw_floats(a, g.grid);
  TmplBwdOut out{g_templates, g_color, a->template_color ? raw_partials : nullptr, g_pose, g_presence, g_bg_image,
                 (alpha && g_alpha) ? alpha_partials : nullptr, scalar_partials};
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_ll_bwd_run_kernel<kC, kA, false, (kC == 1 ? 3 : 2)>;
    rc = tmpl_prepare_kernel(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, x, grad_log_prob, cache, out, g);
    note_launch();
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  if (alpha && g_alpha) {
    rc = launch_reduce_rows(alpha_partials, g_alpha, g.grid, a->M * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  if (a->template_color) {
    rc = launch_reduce_rows(raw_partials, g_templates, g.grid, a->M * a->C * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  return launch_reduce_rows(scalar_partials, g_scalars, g.grid, 4, stream);
}

// Backward of pdf.mode() (distributions.py:50-77, straight_through_gradient=False; what SCAE.loss differentiates when
// recon_mse_weight > 0, stacked_capsule_auto_encoder.py:226-230): the gradient w.r.t. the mode image flows, pixel by pixel,
// into the warp of the component the arg-max picked.  `component_cache[B,2,C,H,W]`: planes [b,0,c] hold that component's
// index as a float (M = background; scae_tmpl_render's mode_component, repeated over the channels in alpha mode), planes
// [b,1,c] are ignored.  Same scatter, workspace and outputs as scae_tmpl_ll_bwd; nothing flows to the mixing logits, so
// g_presence / g_alpha are zero and not produced.  g_scalars[4]: only d / d bg_value can be non-zero.
extern "C" __attribute__((visibility("default"))) int scae_tmpl_mode_bwd(
    const scae_tmpl_args* a, const float* grad_mode, const float* component_cache, float* g_templates, float* g_color,
    float* g_pose, float* g_bg_image, float* g_scalars, void* workspace, size_t workspace_bytes, scae_stream_t stream_) {
  int rc = tmpl_validate(a);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(grad_mode && component_cache && g_templates && g_pose && g_scalars, SCAE_EINVAL,
               "tmpl mode bwd: a required pointer is NULL");
  SCAE_REQUIRE(!a->template_color || g_color, SCAE_EINVAL, "tmpl mode bwd: g_color is required with template_color");
  TmplGeom g;
  rc = tmpl_bwd_plan(a, &g);
  if (rc != SCAE_OK) return rc;
  const size_t need =
      (tmpl_ws_alpha_floats(a, g.grid) + tmpl_ws_raw_floats(a, g.grid) + (size_t)g.grid * 4) * sizeof(float);
  SCAE_REQUIRE(workspace && workspace_bytes >= need, SCAE_EINVAL, "tmpl mode bwd: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool alpha = a->mode == SCAE_TMPL_MODE_ALPHA;
  float* alpha_partials = static_cast<float*>(workspace);
  float* raw_partials = alpha_partials + tmpl_ws_alpha_floats(a, g.grid);
  float* scalar_partials = raw_partials + tmpl_ws_raw_floats(a, g.grid);
  TmplBwdOut out{g_templates, g_color, a->template_color ? raw_partials : nullptr, g_pose, nullptr, g_bg_image, nullptr,
                 scalar_partials};
  SCAE_TMPL_DISPATCH(a->C, alpha, {
    auto kern = tmpl_ll_bwd_run_kernel<kC, kA, true, (kC == 1 ? 3 : 2)>;
    rc = tmpl_prepare_kernel(kern, g.smem_bytes);
    if (rc != SCAE_OK) return rc;
    kern<<<g.grid, g.threads, g.smem_bytes, stream>>>(*a, grad_mode, grad_mode, component_cache, out, g);
    note_launch();
  });
  SCAE_CUDA_TRY(cudaGetLastError());
  if (a->template_color) {
    rc = launch_reduce_rows(raw_partials, g_templates, g.grid, a->M * a->C * a->h * a->w, stream);
    if (rc != SCAE_OK) return rc;
  }
  return launch_reduce_rows(scalar_partials, g_scalars, g.grid, 4, stream);
}
