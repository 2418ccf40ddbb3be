// libscae_b200: one set-attention block (SAB) of the object encoder as ONE kernel per direction.
//
// The set transformer that produces the object encodings for hot path 2 (reference set_transformer.py:74-153: MAB with
// single-head QKV attention, residual, presence mask, LayerNorm, feed-forward, LayerNorm) works on (B, M, 16)
// activations: a dozen 16x16 linears, three (M x M) attention products and a trail of elementwise ops per block, each a
// separate launch on 2.6 MB of data in stock PyTorch -- ~45 launches forward and ~90 backward per block, 1.4 ms of a
// 10 ms train step for three blocks.  Here a CTA owns one image at a time, keeps the whole block in shared memory and
// gives every token a quad of threads (4 features each):
//
//   Q,K,V = X W^T + b                      quad-local, weights broadcast from shared memory
//   A = softmax((Q K^T - (1-p_j) 1e32)/4)  a quad holds one row of A, 10 columns per thread, shuffles within the quad
//   O = A V ; H1 = (O Wo^T + bo + X) p_i ; H2 = LN0(H1) ; H3 = H2 + relu(H2 Wf^T + bf) ; Y = LN1(H3)    quad-local
//
// Only Q/K/V -> attention needs a CTA barrier.  The backward kernel recomputes the forward from X (cheaper than saving
// nine (B,M,16) tensors), walks the chain back with the same mapping, and accumulates the 14 parameter gradients of the
// block in registers across the CTA's images; per-CTA partial rows are summed in a fixed order afterwards
// (deterministic).  The presence mask reproduces the reference exactly, including its fp32 absorption quirk: for a
// presence p_j < 1 the logit becomes -(1 - p_j) 1e32 (set_transformer.py:41-43).
#include "common.cuh"

namespace scae {

constexpr int SD = 16;             // feature width (the reference's dim_hidden)
constexpr int SP = 20;             // padded row stride in floats: quads of 8 consecutive tokens hit distinct banks
constexpr int kSabThreads = 256;   // maximum; a launch uses max(128, 4 N rounded up to a warp) threads
constexpr int kSabMaxN = 64;       // tokens per image (4 threads per token)
constexpr int kSabParamFloats = 5 * SD * SD + 9 * SD;   // wq wk wv wo wf | bq bk bv bo bf g0 b0 g1 b1

struct SabW {          // shared-memory copy of the block's parameters
  float w[5][SD * SP]; // wq, wk, wv, wo, wf: [out][in], rows padded to SP
  float v[9][SD];      // bq, bk, bv, bo, bf, ln0 gamma, ln0 beta, ln1 gamma, ln1 beta
};

__device__ __forceinline__ void sab_load_params(SabW& W, const scae_sab_params& p) {
  const float* wsrc[5] = {p.wq, p.wk, p.wv, p.wo, p.wf};
  const float* vsrc[9] = {p.bq, p.bk, p.bv, p.bo, p.bf, p.ln0_w, p.ln0_b, p.ln1_w, p.ln1_b};
  for (int e = threadIdx.x; e < 5 * SD * SD; e += blockDim.x) {
    const int k = e >> 8, r = e & 255;
    W.w[k][(r >> 4) * SP + (r & 15)] = __ldg(wsrc[k] + r);
  }
  for (int e = threadIdx.x; e < 9 * SD; e += blockDim.x) W.v[e >> 4][e & 15] = __ldg(vsrc[e >> 4] + (e & 15));
}

__device__ __forceinline__ void ld_row(const float* row, float* v) {   // 16 floats
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(row + 4 * q);
    v[4 * q] = t.x;
    v[4 * q + 1] = t.y;
    v[4 * q + 2] = t.z;
    v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// out[r] = sum_c in[c] * w[(4 q + r)][c] + b[4 q + r],  r = 0..3   (nn.Linear: y = x W^T + b)
__device__ __forceinline__ void lin4(const float* in, const float* w, const float* b, int q, float* out) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* wr = w + (4 * q + r) * SP;
    float acc = b ? b[4 * q + r] : 0.0f;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const float4 t = ld4(wr + 4 * c4);
      acc = fmaf(in[4 * c4], t.x, acc);
      acc = fmaf(in[4 * c4 + 1], t.y, acc);
      acc = fmaf(in[4 * c4 + 2], t.z, acc);
      acc = fmaf(in[4 * c4 + 3], t.w, acc);
    }
    out[r] = acc;
  }
}
// out[r] = sum_o g[o] * w[o][4 q + r]   (the input gradient of the same layer: g W)
__device__ __forceinline__ void lin4_t(const float* g, const float* w, int q, float* out) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
#pragma unroll
  for (int o = 0; o < SD; ++o) {
    const float4 t = ld4(w + o * SP + 4 * q);
    out[0] = fmaf(g[o], t.x, out[0]);
    out[1] = fmaf(g[o], t.y, out[1]);
    out[2] = fmaf(g[o], t.z, out[2]);
    out[3] = fmaf(g[o], t.w, out[3]);
  }
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// LayerNorm of a 16-vector spread over a quad (4 values per thread): returns xhat, writes y = xhat * gamma + beta
__device__ __forceinline__ float ln4(const float* x, const float* gamma, const float* beta, float eps, int q, float* xhat,
                                     float* y) {
  const float mean = quad_sum((x[0] + x[1]) + (x[2] + x[3])) * (1.0f / SD);
  float var = 0.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float d = x[r] - mean;
    var = fmaf(d, d, var);
  }
  const float rstd = 1.0f / sqrtf(quad_sum(var) * (1.0f / SD) + eps);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    xhat[r] = (x[r] - mean) * rstd;
    y[r] = fmaf(xhat[r], gamma[4 * q + r], beta[4 * q + r]);
  }
  return rstd;
}
// gx = rstd * (g gamma - mean(g gamma) - xhat mean(g gamma xhat))
__device__ __forceinline__ void ln4_bwd(const float* g, const float* xhat, const float* gamma, float rstd, int q,
                                        float* gx) {
  float gg[4], m1 = 0.0f, m2 = 0.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    gg[r] = g[r] * gamma[4 * q + r];
    m1 += gg[r];
    m2 = fmaf(gg[r], xhat[r], m2);
  }
  m1 = quad_sum(m1) * (1.0f / SD);
  m2 = quad_sum(m2) * (1.0f / SD);
#pragma unroll
  for (int r = 0; r < 4; ++r) gx[r] = rstd * (gg[r] - m1 - xhat[r] * m2);
}

// Shared-memory matrices of one image, [N][SP] each
struct SabFwdBuf {
  float *X, *Q, *K, *V, *O, *H2;   // H2 = LN0 output (input of the feed-forward layer)
  float* A;                        // [N][NA] attention weights
  float* P;                        // [N] presence (1 when absent)
};

// Values a token thread keeps in registers between the forward recomputation and the backward pass
struct SabTok {
  float xh1[4], rstd0;   // LN0: normalised input, 1/std
  float xh3[4], rstd1;   // LN1
  float fpos[4];         // relu mask of the feed-forward layer (1 / 0)
};

// Forward of one image.  Thread (i, q) = (tid >> 2, tid & 3) owns features 4q..4q+3 of token i.  y4 receives the block
// output for that slot.  Contains CTA barriers: call from all threads.
template <bool kMask>
__device__ __forceinline__ void sab_forward(const SabW& W, const SabFwdBuf& s, int N, int NA, float eps0, float eps1,
                                            float* y4, SabTok& tk) {
  const int tid = threadIdx.x, i = min(tid >> 2, N - 1), q = tid & 3;
  const bool tok = tid < 4 * N;
  float xrow[SD];
  ld_row(s.X + i * SP, xrow);
  {
    float o4[4];
    lin4(xrow, W.w[0], W.v[0], q, o4);
    if (tok) st4(s.Q + i * SP + 4 * q, o4);
    lin4(xrow, W.w[1], W.v[1], q, o4);
    if (tok) st4(s.K + i * SP + 4 * q, o4);
    lin4(xrow, W.w[2], W.v[2], q, o4);
    if (tok) st4(s.V + i * SP + 4 * q, o4);
  }
  __syncthreads();
  // ---- attention row i: this thread holds columns j = q, q + 4, ... ---------------------------------------------------
  float qrow[SD];
  ld_row(s.Q + i * SP, qrow);
  float lg[kSabMaxN / 4];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < kSabMaxN / 4; ++t) {
    const int j = q + 4 * t;
    lg[t] = -INFINITY;
    if (j < N) {
      float krow[SD];
      ld_row(s.K + j * SP, krow);
      float acc = 0.0f;
#pragma unroll
      for (int c = 0; c < SD; ++c) acc = fmaf(qrow[c], krow[c], acc);
      if (kMask) acc = acc - (1.0f - s.P[j]) * 1e32f;   // set_transformer.py:41-43, before the 1/sqrt(d) scaling
      lg[t] = acc * 0.25f;                              // / sqrt(16)
      mx = fmaxf(mx, lg[t]);
    }
  }
  mx = quad_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int t = 0; t < kSabMaxN / 4; ++t) {
    const int j = q + 4 * t;
    lg[t] = j < N ? expf(lg[t] - mx) : 0.0f;
    sum += lg[t];
  }
  const float inv = 1.0f / quad_sum(sum);
#pragma unroll
  for (int t = 0; t < kSabMaxN / 4; ++t) {
    const int j = q + 4 * t;
    if (tok && j < N) s.A[i * NA + j] = lg[t] * inv;
  }
  __syncwarp();
  // ---- O = A V ----------------------------------------------------------------------------------------------------
  float o4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < N; ++j) {
    const float a = s.A[i * NA + j];
    const float4 v = ld4(s.V + j * SP + 4 * q);
    o4[0] = fmaf(a, v.x, o4[0]);
    o4[1] = fmaf(a, v.y, o4[1]);
    o4[2] = fmaf(a, v.z, o4[2]);
    o4[3] = fmaf(a, v.w, o4[3]);
  }
  if (tok) st4(s.O + i * SP + 4 * q, o4);
  __syncwarp();
  // ---- H1 = (O Wo^T + bo + X) p_i ; H2 = LN0(H1) --------------------------------------------------------------------
  float orow[SD], h1[4], h2[4];
  ld_row(s.O + i * SP, orow);
  lin4(orow, W.w[3], W.v[3], q, h1);
  const float pi = kMask ? s.P[i] : 1.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) h1[r] = (h1[r] + xrow[4 * q + r]) * pi;
  tk.rstd0 = ln4(h1, W.v[5], W.v[6], eps0, q, tk.xh1, h2);
  if (tok) st4(s.H2 + i * SP + 4 * q, h2);
  __syncwarp();
  // ---- H3 = H2 + relu(H2 Wf^T + bf) ; Y = LN1(H3) ---------------------------------------------------------------------
  float h2row[SD], f4[4], h3[4];
  ld_row(s.H2 + i * SP, h2row);
  lin4(h2row, W.w[4], W.v[4], q, f4);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    tk.fpos[r] = f4[r] > 0.0f ? 1.0f : 0.0f;
    h3[r] = h2[r] + fmaxf(f4[r], 0.0f);
  }
  tk.rstd1 = ln4(h3, W.v[7], W.v[8], eps1, q, tk.xh3, y4);
}

__device__ __forceinline__ void sab_load_image(float* X, float* P, const float* x, const float* presence, int b, int N) {
  for (int e = threadIdx.x; e < N * 4; e += blockDim.x) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x + ((size_t)b * N) * SD) + e);
    *reinterpret_cast<float4*>(X + (e >> 2) * SP + 4 * (e & 3)) = t;
  }
  for (int e = threadIdx.x; e < N; e += blockDim.x) P[e] = presence ? __ldg(presence + (size_t)b * N + e) : 1.0f;
}

static inline int sab_na(int N) { return N | 1; }
static inline size_t sab_fwd_smem(int N) {
  return sizeof(SabW) + ((size_t)6 * N * SP + (size_t)N * sab_na(N) + kSabMaxN) * sizeof(float);
}
static inline size_t sab_bwd_smem(int N) {
  return sizeof(SabW) + ((size_t)16 * N * SP + (size_t)2 * N * sab_na(N) + kSabMaxN) * sizeof(float);
}

template <bool kMask>
__global__ void __launch_bounds__(kSabThreads, 2) sab_fwd_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ presence,
                                                              const scae_sab_params p, int B, int N,
                                                              float* __restrict__ y) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SabW& W = *reinterpret_cast<SabW*>(smem_raw);
  float* f = reinterpret_cast<float*>(smem_raw + sizeof(SabW));
  const int NA = N | 1, mat = N * SP;
  SabFwdBuf s{f, f + mat, f + 2 * mat, f + 3 * mat, f + 4 * mat, f + 5 * mat, f + 6 * mat, f + 6 * mat + N * NA};
  sab_load_params(W, p);
  const int tid = threadIdx.x, i = tid >> 2, q = tid & 3;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();                       // previous image fully consumed (and parameters loaded)
    sab_load_image(s.X, s.P, x, presence, b, N);
    __syncthreads();
    float y4[4];
    SabTok tk;
    sab_forward<kMask>(W, s, N, NA, p.eps0, p.eps1, y4, tk);
    if (tid < 4 * N) st4(y + ((size_t)b * N + i) * SD + 4 * q, y4);
  }
}

template <bool kMask>
__global__ void __launch_bounds__(kSabThreads, 2) sab_bwd_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ presence,
                                                              const scae_sab_params p, const float* __restrict__ gy,
                                                              int B, int N, float* __restrict__ gx,
                                                              float* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SabW& W = *reinterpret_cast<SabW*>(smem_raw);
  float* f = reinterpret_cast<float*>(smem_raw + sizeof(SabW));
  const int NA = N | 1, mat = N * SP;
  SabFwdBuf s{f, f + mat, f + 2 * mat, f + 3 * mat, f + 4 * mat, f + 5 * mat, f + 16 * mat,
              f + 16 * mat + 2 * N * NA};
  // gradient-side matrices
  float* GY = f + 6 * mat;     // dY               -> d ln1 beta
  float* P1 = f + 7 * mat;     // dY * xhat3        -> d ln1 gamma
  float* GF = f + 8 * mat;     // d pre-activation of the feed-forward layer -> d bf, d Wf (with H2)
  float* G2 = f + 9 * mat;     // d H2 (LN0 output)  -> d ln0 beta
  float* P0 = f + 10 * mat;    // d H2 * xhat1       -> d ln0 gamma
  float* G0 = f + 11 * mat;    // d H0               -> d bo, d Wo (with O)
  float* GO = f + 12 * mat;    // d O
  float* GQ = f + 13 * mat;    // d Q -> d bq, d Wq (with X)
  float* GK = f + 14 * mat;    // d K
  float* GV = f + 15 * mat;    // d V
  float* DL = s.A + N * NA;    // [N][NA] d logits (already scaled by 1/4)
  sab_load_params(W, p);
  const int tid = threadIdx.x, i = min(tid >> 2, N - 1), q = tid & 3;
  const bool tok = tid < 4 * N;
  // parameter-gradient accumulators: entries e = tid, tid + blockDim of the 256-entry weight matrices (e = 16 o + c)
  // and of the 144-entry vector block; blockDim >= 128, so two slots per thread are enough
  float acc_w[5][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, acc_v[2] = {0.f, 0.f};

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    sab_load_image(s.X, s.P, x, presence, b, N);
    __syncthreads();
    float y4[4];
    SabTok tk;
    sab_forward<kMask>(W, s, N, NA, p.eps0, p.eps1, y4, tk);

    // ---- LN1 backward, feed-forward backward, LN0 backward (quad-local) -------------------------------------------------
    float g4[4], gh3[4], t4[4];
    {
      const float4 t = __ldg(reinterpret_cast<const float4*>(gy + ((size_t)b * N + i) * SD + 4 * q));
      g4[0] = t.x;
      g4[1] = t.y;
      g4[2] = t.z;
      g4[3] = t.w;
    }
    if (tok) {
      st4(GY + i * SP + 4 * q, g4);
#pragma unroll
      for (int r = 0; r < 4; ++r) t4[r] = g4[r] * tk.xh3[r];
      st4(P1 + i * SP + 4 * q, t4);
    }
    ln4_bwd(g4, tk.xh3, W.v[7], tk.rstd1, q, gh3);
    float gf[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) gf[r] = gh3[r] * tk.fpos[r];
    if (tok) st4(GF + i * SP + 4 * q, gf);
    __syncwarp();
    float gfrow[SD], g2[4];
    ld_row(GF + i * SP, gfrow);
    lin4_t(gfrow, W.w[4], q, g2);
#pragma unroll
    for (int r = 0; r < 4; ++r) g2[r] += gh3[r];
    if (tok) {
      st4(G2 + i * SP + 4 * q, g2);
#pragma unroll
      for (int r = 0; r < 4; ++r) t4[r] = g2[r] * tk.xh1[r];
      st4(P0 + i * SP + 4 * q, t4);
    }
    float g1[4], g0[4];
    ln4_bwd(g2, tk.xh1, W.v[5], tk.rstd0, q, g1);
    const float pi = kMask ? s.P[i] : 1.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r) g0[r] = g1[r] * pi;
    if (tok) st4(G0 + i * SP + 4 * q, g0);
    __syncwarp();
    // ---- d O = d H0 Wo --------------------------------------------------------------------------------------------------
    float g0row[SD], go[4];
    ld_row(G0 + i * SP, g0row);
    lin4_t(g0row, W.w[3], q, go);
    if (tok) st4(GO + i * SP + 4 * q, go);
    __syncwarp();
    // ---- attention backward, row i: dA = dO V^T, dS = A (dA - sum_j A dA), dL = dS / 4 ------------------------------------
    float gorow[SD];
    ld_row(GO + i * SP, gorow);
    float da[kSabMaxN / 4], rs = 0.0f;
#pragma unroll
    for (int t = 0; t < kSabMaxN / 4; ++t) {
      const int j = q + 4 * t;
      da[t] = 0.0f;
      if (j < N) {
        float vrow[SD];
        ld_row(s.V + j * SP, vrow);
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < SD; ++c) acc = fmaf(gorow[c], vrow[c], acc);
        da[t] = acc;
        rs = fmaf(s.A[i * NA + j], acc, rs);
      }
    }
    rs = quad_sum(rs);
#pragma unroll
    for (int t = 0; t < kSabMaxN / 4; ++t) {
      const int j = q + 4 * t;
      if (tok && j < N) DL[i * NA + j] = s.A[i * NA + j] * (da[t] - rs) * 0.25f;
    }
    __syncthreads();    // d logits and d O of every token are in shared memory
    // ---- d Q (row-wise), d K and d V (column-wise) ------------------------------------------------------------------------
    float gq[4] = {0.f, 0.f, 0.f, 0.f}, gk[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < N; ++j) {
      const float dl_ij = DL[i * NA + j];       // row i, column j   -> d Q_i += dL_ij K_j
      const float dl_ji = DL[j * NA + i];       // row j, column i   -> d K_i += dL_ji Q_j
      const float a_ji = s.A[j * NA + i];       //                   -> d V_i += A_ji dO_j
      const float4 kj = ld4(s.K + j * SP + 4 * q), qj = ld4(s.Q + j * SP + 4 * q), oj = ld4(GO + j * SP + 4 * q);
      gq[0] = fmaf(dl_ij, kj.x, gq[0]);
      gq[1] = fmaf(dl_ij, kj.y, gq[1]);
      gq[2] = fmaf(dl_ij, kj.z, gq[2]);
      gq[3] = fmaf(dl_ij, kj.w, gq[3]);
      gk[0] = fmaf(dl_ji, qj.x, gk[0]);
      gk[1] = fmaf(dl_ji, qj.y, gk[1]);
      gk[2] = fmaf(dl_ji, qj.z, gk[2]);
      gk[3] = fmaf(dl_ji, qj.w, gk[3]);
      gv[0] = fmaf(a_ji, oj.x, gv[0]);
      gv[1] = fmaf(a_ji, oj.y, gv[1]);
      gv[2] = fmaf(a_ji, oj.z, gv[2]);
      gv[3] = fmaf(a_ji, oj.w, gv[3]);
    }
    if (tok) {
      st4(GQ + i * SP + 4 * q, gq);
      st4(GK + i * SP + 4 * q, gk);
      st4(GV + i * SP + 4 * q, gv);
    }
    __syncwarp();
    // ---- d X = d H0 (residual) + d Q Wq + d K Wk + d V Wv ---------------------------------------------------------------
    {
      float row[SD], part[4], gx4[4] = {g0[0], g0[1], g0[2], g0[3]};
      ld_row(GQ + i * SP, row);
      lin4_t(row, W.w[0], q, part);
#pragma unroll
      for (int r = 0; r < 4; ++r) gx4[r] += part[r];
      ld_row(GK + i * SP, row);
      lin4_t(row, W.w[1], q, part);
#pragma unroll
      for (int r = 0; r < 4; ++r) gx4[r] += part[r];
      ld_row(GV + i * SP, row);
      lin4_t(row, W.w[2], q, part);
#pragma unroll
      for (int r = 0; r < 4; ++r) gx4[r] += part[r];
      if (tok) st4(gx + ((size_t)b * N + i) * SD + 4 * q, gx4);
    }
    __syncthreads();    // every gradient matrix of the image is complete
    // ---- parameter gradients: d W[o][c] += sum_i G[i][o] IN[i][c] ; vectors: column sums --------------------------------
    {
      const float* G[5] = {GQ, GK, GV, G0, GF};
      const float* IN[5] = {s.X, s.X, s.X, s.O, s.H2};
      const float* V9[9] = {GQ, GK, GV, G0, GF, P0, G2, P1, GY};   // bq bk bv bo bf | ln0 gamma, beta | ln1 gamma, beta
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + u * (int)blockDim.x;
        if (e < SD * SD) {
          const int o = e >> 4, c = e & 15;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            float a = acc_w[k][u];
            for (int t = 0; t < N; ++t) a = fmaf(G[k][t * SP + o], IN[k][t * SP + c], a);
            acc_w[k][u] = a;
          }
        }
        if (e < 9 * SD) {
          const float* src = V9[e >> 4] + (e & 15);
          float a = acc_v[u];
          for (int t = 0; t < N; ++t) a += src[t * SP];
          acc_v[u] = a;
        }
      }
    }
  }
  float* row = partials + (size_t)blockIdx.x * kSabParamFloats;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int e = tid + u * (int)blockDim.x;
    if (e < SD * SD) {
#pragma unroll
      for (int k = 0; k < 5; ++k) row[k * SD * SD + e] = acc_w[k][u];
    }
    if (e < 9 * SD) row[5 * SD * SD + e] = acc_v[u];
  }
}

static int sab_threads(int N) {
  const int t = (4 * N + 31) / 32 * 32;
  return t < 128 ? 128 : t;
}

static int sab_grid(int B, int N, size_t smem, int regs_per_thread) {
  const int threads = sab_threads(N);
  int per_sm = (int)(((size_t)max_smem_optin() + 1024) / (smem + 1024));
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm > 65536 / (threads * regs_per_thread)) per_sm = 65536 / (threads * regs_per_thread);
  if (per_sm < 1) per_sm = 1;
  const long slots = (long)sm_count() * per_sm;
  return (int)(B < slots ? B : slots);
}

static int sab_check(const float* x, const scae_sab_params* p, int B, int N) {
  SCAE_REQUIRE(x && p, SCAE_EINVAL, "sab: x and params are required");
  SCAE_REQUIRE(p->wq && p->bq && p->wk && p->bk && p->wv && p->bv && p->wo && p->bo && p->wf && p->bf && p->ln0_w &&
                   p->ln0_b && p->ln1_w && p->ln1_b,
               SCAE_EINVAL, "sab: every parameter pointer is required");
  SCAE_REQUIRE(B > 0 && N > 0 && N <= kSabMaxN, SCAE_ELIMIT, "sab: N=%d tokens (max %d), B=%d", N, kSabMaxN, B);
  SCAE_REQUIRE(aligned16(x), SCAE_EINVAL, "sab: x must be 16-byte aligned");
  return SCAE_OK;
}

}  // namespace scae

using namespace scae;
#define SCAE_EXPORT __attribute__((visibility("default")))

extern "C" {

SCAE_EXPORT int scae_sab_fwd(const float* x, const float* presence, const scae_sab_params* p, int B, int N, float* y,
                             scae_stream_t stream_) {
  int rc = sab_check(x, p, B, N);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(y && aligned16(y), SCAE_EINVAL, "sab fwd: y is required (16-byte aligned)");
  const size_t smem = sab_fwd_smem(N);
  SCAE_REQUIRE(smem <= (size_t)max_smem_optin(), SCAE_ELIMIT, "sab fwd: %zu bytes of shared memory needed", smem);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  auto kern = presence ? sab_fwd_kernel<true> : sab_fwd_kernel<false>;
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<sab_grid(B, N, smem, 104), sab_threads(N), smem, stream>>>(x, presence, *p, B, N, y);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return SCAE_OK;
}

SCAE_EXPORT size_t scae_sab_bwd_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0 || N > kSabMaxN) return 0;
  return (size_t)sab_grid(B, N, sab_bwd_smem(N), 128) * kSabParamFloats * sizeof(float);
}

SCAE_EXPORT int scae_sab_bwd(const float* x, const float* presence, const scae_sab_params* p, const float* gy, int B,
                             int N, float* gx, float* g_params, void* workspace, size_t workspace_bytes,
                             scae_stream_t stream_) {
  int rc = sab_check(x, p, B, N);
  if (rc != SCAE_OK) return rc;
  SCAE_REQUIRE(gy && gx && g_params && aligned16(gy) && aligned16(gx), SCAE_EINVAL,
               "sab bwd: gy, gx (16-byte aligned) and g_params are required");
  const size_t smem = sab_bwd_smem(N);
  SCAE_REQUIRE(smem <= (size_t)max_smem_optin(), SCAE_ELIMIT, "sab bwd: %zu bytes of shared memory needed", smem);
  const int grid = sab_grid(B, N, smem, 128);
  SCAE_REQUIRE(workspace && workspace_bytes >= (size_t)grid * kSabParamFloats * sizeof(float), SCAE_EINVAL,
               "sab bwd: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  auto kern = presence ? sab_bwd_kernel<true> : sab_bwd_kernel<false>;
  SCAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float* partials = static_cast<float*>(workspace);
  kern<<<grid, sab_threads(N), smem, stream>>>(x, presence, *p, gy, B, N, gx, partials);
  note_launch();
  SCAE_CUDA_TRY(cudaGetLastError());
  return launch_reduce_rows(partials, g_params, grid, kSabParamFloats, stream);
}

}  // extern "C"
