/*
 * scae_b200.h -- C ABI of libscae_b200.so: the two SCAE likelihood hot paths as sm_100a CUDA kernels.
 *
 * The reference (bdsaglam/torch-scae, pure Python) has no FFI; its boundary for these paths is the Python module
 * API.  Each entry point below names the reference code it replaces (paths relative to torch_scae/ in the
 * reference).  The Python mirror of that API (torch_scae_b200/part_decoder.py, object_decoder.py) binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (int64 where stated), at least 16-byte aligned;
 *     nothing is allocated, freed or retained by the library; outputs are caller-allocated;
 *   - "nullable" pointers may be NULL: a NULL input means "absent" with the reference's semantics for None,
 *     a NULL output is simply not written;
 *   - all work is enqueued on `stream` (a cudaStream_t) and is asynchronous with respect to the host;
 *   - return value 0 = success, otherwise a negative SCAE_E* code; scae_last_error() returns a thread-local
 *     human-readable message for the last failing call on this thread;
 *   - re-entrant; no global state besides that message and per-device cached attributes.
 *
 * Shapes use the reference's symbols: B batch, M templates (= part capsules), C channels, h x w template,
 * H x W image, O object capsules, V votes per object (= M), A = 8V+7 the per-capsule MLP output width.
 */
#ifndef SCAE_B200_H_
#define SCAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCAE_B200_ABI_VERSION 5

#define SCAE_OK 0
#define SCAE_EINVAL (-1)   /* bad shape / flag / NULL where a pointer is required / misaligned pointer */
#define SCAE_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define SCAE_ELIMIT (-3)   /* shape exceeds what the kernels support (see scae_*_limits in DESIGN.md) */

typedef void* scae_stream_t; /* cudaStream_t */

int scae_abi_version(void);
const char* scae_last_error(void);
/* Compiled-for architecture string, e.g. "sm_100a". */
const char* scae_build_arch(void);
/* Hash of the sources this library was compiled from (torch_scae_b200/build.py::source_id); the Python binding refuses a
 * library whose id differs from the checked-out sources. */
const char* scae_build_id(void);
/* Number of CUDA kernels this library has launched from the calling thread so far (monotonic; bench.py's launch
 * accounting reads it before and after each entry point). */
unsigned long long scae_launch_count(void);
/* Number of scae_caps_ll_fwd / _bwd calls (process-wide) that the TMA-staged fast path (csrc/caps_ll2.cu) served; the
 * others ran the general kernels (unsupported shape, misaligned pointers, extra upstream gradients). */
unsigned long long scae_caps_fast_path_count(void);
/* ... of which the persistent warp-specialised kernels (csrc/caps_ll3.cu, caps_ll3_bwd.cu) served this many. */
unsigned long long scae_caps_persistent_path_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Hot path 1: template warp + per-pixel template-mixture Gaussian log-likelihood
 * replaces TemplateBasedImageDecoder.forward (part_decoder.py:152-243: F.affine_grid + F.grid_sample, background
 * component, alpha/temperature mixing logits, log_safe(presence)) fused with GaussianMixture.log_prob
 * (distributions.py:41-48) as evaluated by SCAE.loss (stacked_capsule_auto_encoder.py:220).
 * ------------------------------------------------------------------------------------------------------------ */

#define SCAE_TMPL_MODE_ALPHA 0       /* use_alpha_channel=True : logits = warp(templates_alpha)           */
#define SCAE_TMPL_MODE_TEMPERATURE 1 /* use_alpha_channel=False: logits = loc / (softplus(t+.5)+1e-4)     */

/* The learnt scalars are passed RAW (as stored in the reference's state_dict) as device pointers so that no host
 * synchronisation is needed; the kernels apply sigmoid / softplus themselves (part_decoder.py:192,:210,:216,:221). */
typedef struct scae_tmpl_args {
  const float* templates;         /* [B,M,C,h,w]; or, with template_color, the batch-shared raw templates [M,C,h,w] */
  const float* templates_alpha;   /* [M,h,w]      alpha mode; NULL in temperature mode                    */
  const float* pose;              /* [B,M,6]      used directly as the 2x3 affine_grid theta              */
  const float* presence;          /* [B,M]        nullable                                                */
  const float* bg_image;          /* [B,C,H,W]    nullable; when NULL bg_value must be given              */
  const float* bg_value;          /* [1] raw      nullable iff bg_image given                             */
  const float* bg_mixing_logit;   /* [1] raw      alpha mode                                              */
  const float* temperature_logit; /* [1] raw      temperature mode                                        */
  const float* scale;             /* [1] raw learnt output scale; NULL => sigma = 1                       */
  const float* template_color;    /* [B,M,C] nullable: fused colourisation (TemplateGenerator.forward,
                                     part_decoder.py:90-105): template[b,m,c] = templates[m,c] * template_color[b,m,c],
                                     so the (B,M,C,h,w) tensor of coloured templates is never formed              */
  int B, M, C, h, w, H, W;
  int mode;                       /* SCAE_TMPL_MODE_*                                                     */
} scae_tmpl_args;

/* log_prob[B,C,H,W] = pdf.log_prob(x);  ll[B] (nullable) = sum over (C,H,W);  cache[B,2,C,H,W] (nullable) holds
 * the two per-pixel logsumexp terms the backward kernel needs. */
int scae_tmpl_ll_fwd(const scae_tmpl_args* a, const float* x, float* log_prob, float* ll, float* cache,
                     scae_stream_t stream);

/* Bytes of scratch scae_tmpl_ll_bwd needs for its per-CTA partial sums of the batch-reduced gradients. */
size_t scae_tmpl_ll_bwd_workspace_bytes(const scae_tmpl_args* a);

/* Gradients of  sum(grad_log_prob * log_prob)  w.r.t. every differentiable input:
 *   g_templates[B,M,C,h,w], g_pose[B,M,6] (required); g_presence[B,M], g_bg_image[B,C,H,W], g_alpha[M,h,w]
 *   (nullable); g_scalars[4] = d/d raw {bg_value, bg_mixing_logit, temperature_logit, scale} (required; entries
 *   for absent parameters are written as 0).  With template_color: g_templates is the batch-reduced [M,C,h,w]
 *   gradient of the raw templates and g_color[B,M,C] (required then, ignored otherwise) the colour gradient.
 *   Deterministic: batch-reduced gradients are summed in a fixed order. */
int scae_tmpl_ll_bwd(const scae_tmpl_args* a, const float* x, const float* grad_log_prob, const float* cache,
                     float* g_templates, float* g_color, float* g_pose, float* g_presence, float* g_bg_image,
                     float* g_alpha, float* g_scalars, void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* Materialises what the reference's decoder returns eagerly (part_decoder.py:239-243) and the mixture's point
 * estimates (distributions.py:37-39, :50-77 with straight_through_gradient=False); every output nullable:
 *   transformed_templates[B,M+1,C,H,W], mixing_logits[B,M+1,(alpha?1:C),H,W] (presence already added),
 *   mode[B,C,H,W], mean[B,C,H,W], mode_component[B,(alpha?1:C),H,W] = index (as a float) of the component the mode takes
 *   each pixel from, M = background.  No gradients here; scae_tmpl_mode_bwd is the backward of `mode`. */
int scae_tmpl_render(const scae_tmpl_args* a, float* transformed_templates, float* mixing_logits, float* mode,
                     float* mean, float* mode_component, scae_stream_t stream);

/* Backward of pdf.mode() (distributions.py:50-77 with straight_through_gradient=False -- what SCAE.loss differentiates
 * when recon_mse_weight > 0, stacked_capsule_auto_encoder.py:226-230): the gradient w.r.t. the mode image flows, pixel by
 * pixel, into the warp of the component the arg-max picked.  component_cache[B,2,C,H,W]: planes [b,0,c] hold that
 * component's index as a float (scae_tmpl_render's mode_component[B,(alpha?1:C),H,W], M = background; repeated over the
 * channels in alpha mode), planes [b,1,c] are ignored.  Outputs and workspace as scae_tmpl_ll_bwd (same scatter kernel);
 * the mixing logits get no gradient (g_presence = g_alpha = 0, not produced); of g_scalars[4] only d/d bg_value can be
 * non-zero. */
int scae_tmpl_mode_bwd(const scae_tmpl_args* a, const float* grad_mode, const float* component_cache, float* g_templates,
                       float* g_color, float* g_pose, float* g_bg_image, float* g_scalars, void* workspace,
                       size_t workspace_bytes, scae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Hot path 2: object->part vote composition + part-pose mixture likelihood
 * replaces the post-MLP half of CapsuleLayer.forward (object_decoder.py:160-236), cv_ops.geometric_transform
 * (cv_ops.py:20-76) for votes, CapsuleObjectDecoder.forward's glue (:413-415) and CapsuleLikelihood.__call__
 * (:257-372).
 * ------------------------------------------------------------------------------------------------------------ */

#define SCAE_CAPS_SIMILARITY 1u        /* similarity_transform=True  (cv_ops.py:51-54)                      */
#define SCAE_CAPS_LEARN_VOTE_SCALE 2u  /* learn_vote_scale=True      (object_decoder.py:223-227)            */
#define SCAE_CAPS_ALLOW_DEFORM 4u      /* allow_deformations=True    (object_decoder.py:168-169)            */
#define SCAE_CAPS_RELU_GRAD 8u         /* bwd only: multiply g_all_param by (all_param > 0), i.e. fuse the
                                          backward of the MLP's final ReLU (nn_ext.py:19-31)                */

typedef struct scae_caps_args {
  const float* all_param;   /* [B,O,A] per-capsule MLP outputs, split as [6V | 6 | 1 | V | V] (:91-99)      */
  const float* cpr_static;  /* [O,V,6]                                                                    */
  const float* bias_cvr;    /* [O,6]   caps_bias_list[0]                                                   */
  const float* bias_caps;   /* [O]     caps_bias_list[1]                                                   */
  const float* bias_vote;   /* [O,V]   caps_bias_list[2]                                                   */
  const float* bias_scale;  /* [O,V]   caps_bias_list[3]                                                   */
  const float* noise_caps;  /* [B,O]   nullable; already scaled: (rand-.5)*noise_scale (:201)              */
  const float* noise_vote;  /* [B,O,V] nullable                                                            */
  const float* x;           /* [B,V,6] part poses                                                          */
  const float* presence;    /* [B,V]   nullable                                                            */
  const float* dummy_vote;  /* [V,6]                                                                       */
  int B, O, V;
  unsigned flags;           /* SCAE_CAPS_*                                                                 */
} scae_caps_args;

/* Every tensor of the AttrDict CapsuleObjectDecoder.forward returns (object_decoder.py:229-236,:361-372,:413-415),
 * plus per-example partial sums; all nullable except `posterior` when a backward pass will follow. */
typedef struct scae_caps_outputs {
  float* vote;                    /* [B,O,V,6]                                                             */
  float* scale;                   /* [B,O,V]                                                               */
  float* vote_presence;           /* [B,O,V]                                                               */
  float* presence_logit_per_caps; /* [B,O]  (the reference's [B,O,1])                                      */
  float* presence_logit_per_vote; /* [B,O,V]                                                               */
  float* caps_presence;           /* [B,O]   max over V                                                    */
  int32_t* caps_presence_arg;     /* [B,O]   argmax over V (lowest index on ties); needed by backward      */
  float* log_prob_per_point;      /* [B,V]   logsumexp_o of the posterior logits, before presence weighting */
  float* ll_per_example;          /* [B]     sum_v presence * logsumexp_o; log_prob = mean over B          */
  float* reg_per_example;         /* [B]     sum cpr_dynamic^2 / 2; cpr_dynamic_reg_loss = sum / B         */
  float* vote_presence_binary;    /* [B,O,V] 0/1                                                           */
  float* winner;                  /* [B,V,6]                                                               */
  float* winner_presence;         /* [B,V]                                                                 */
  int64_t* winner_idx;            /* [B,V]   argmax over O (lowest index on ties); needed by backward      */
  int64_t* is_from_capsule;       /* [B,V]   winner_idx / V (sic, object_decoder.py:334)                   */
  float* soft_winner;             /* [B,V,6]                                                               */
  float* soft_winner_presence;    /* [B,V]                                                                 */
  float* posterior_mixing_prob;   /* [B,O,V]                                                               */
  float* mixing_log_prob;         /* [B,O+1,V]                                                             */
  float* mixing_logit;            /* [B,O+1,V]                                                             */
} scae_caps_outputs;

int scae_caps_ll_fwd(const scae_caps_args* a, const scae_caps_outputs* out, scae_stream_t stream);

/* Upstream gradients, one per differentiable output; all nullable (NULL = zero). */
typedef struct scae_caps_upstream {
  const float* g_ll_per_example;          /* [B]                                                          */
  const float* g_reg_per_example;         /* [B]                                                          */
  const float* g_posterior_mixing_prob;   /* [B,O,V]                                                      */
  const float* g_caps_presence;           /* [B,O]                                                        */
  const float* g_vote_presence;           /* [B,O,V]                                                      */
  const float* g_soft_winner;             /* [B,V,6]                                                      */
  const float* g_soft_winner_presence;    /* [B,V]                                                        */
  const float* g_winner;                  /* [B,V,6]                                                      */
  const float* g_winner_presence;         /* [B,V]                                                        */
  const float* g_vote;                    /* [B,O,V,6]                                                    */
  const float* g_scale;                   /* [B,O,V]                                                      */
  const float* g_presence_logit_per_caps; /* [B,O]                                                        */
  const float* g_presence_logit_per_vote; /* [B,O,V]                                                      */
  const float* g_mixing_logit;            /* [B,O+1,V]                                                    */
  const float* g_mixing_log_prob;         /* [B,O+1,V]                                                    */
} scae_caps_upstream;

/* Saved forward results the backward kernel reads instead of recomputing. */
typedef struct scae_caps_saved {
  const float* posterior_mixing_prob; /* [B,O,V]  required                                                */
  const float* log_prob_per_point;    /* [B,V]    required                                                */
  const int32_t* caps_presence_arg;   /* [B,O]    required iff g_caps_presence given                      */
  const int64_t* winner_idx;          /* [B,V]    required iff g_winner / g_winner_presence given         */
} scae_caps_saved;

size_t scae_caps_ll_bwd_workspace_bytes(const scae_caps_args* a);

/* g_all_param[B,O,A] (required); g_shared[O,A] (required) = sum over B of the gradient w.r.t. the pre-activation
 * sums, laid out like one all_param row, i.e. [g_cpr_static(6V) | g_bias_cvr(6) | g_bias_caps(1) | g_bias_vote(V) |
 * g_bias_scale(V)] per capsule; g_dummy_vote[V,6], g_x[B,V,6], g_presence[B,V] nullable.  Deterministic. */
int scae_caps_ll_bwd(const scae_caps_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up,
                     float* g_all_param, float* g_shared, float* g_dummy_vote, float* g_x, float* g_presence,
                     void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Hot path 2 on EXPLICIT vote tensors (csrc/caps_explicit.cu): the reference's standalone
 *   CapsuleLikelihood(vote, scale, vote_presence, dummy_vote)(x, presence)          object_decoder.py:243-372
 * for callers that build their own votes instead of going through CapsuleLayer (its own test does,
 * tests/test_object_decoder.py:62-112).  Same outputs as scae_caps_ll_fwd minus the ones that come from the parameter
 * head: of scae_caps_outputs, vote / scale / vote_presence / presence_logit_* / caps_presence(_arg) / reg_per_example
 * are ignored; log_prob = mean over B of ll_per_example.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct scae_caps_explicit_args {
  const float* vote;          /* [B,O,V,6]  rows 0-1 of the 3x3 vote matrices (the reference's [B,O,V,6] `vote`)     */
  const float* scale;         /* [B,O,V]    > 0                                                                      */
  const float* vote_presence; /* [B,O,V]                                                                             */
  const float* dummy_vote;    /* [V,6]                                                                               */
  const float* x;             /* [B,V,6]                                                                             */
  const float* presence;      /* [B,V] nullable (= ones)                                                             */
  float* point_ll;            /* [B,V] scratch: presence * logsumexp_o, summed per example in a fixed order          */
  int B, O, V;
} scae_caps_explicit_args;

int scae_caps_explicit_fwd(const scae_caps_explicit_args* a, const scae_caps_outputs* out, scae_stream_t stream);
size_t scae_caps_explicit_bwd_workspace_bytes(const scae_caps_explicit_args* a);
/* Of scae_caps_upstream: g_ll_per_example, g_posterior_mixing_prob, g_soft_winner(_presence), g_winner(_presence),
 * g_mixing_logit, g_mixing_log_prob are honoured.  g_vote[B,O,V,6], g_scale[B,O,V], g_vote_presence[B,O,V] required;
 * g_dummy_vote[V,6] (summed over the batch in a fixed order; needs the workspace), g_x[B,V,6], g_presence[B,V]
 * nullable.  Deterministic. */
int scae_caps_explicit_bwd(const scae_caps_explicit_args* a, const scae_caps_saved* saved, const scae_caps_upstream* up,
                           float* g_vote, float* g_scale, float* g_vote_presence, float* g_dummy_vote, float* g_x,
                           float* g_presence, void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Plumbing for the callers of hot path 2: column sums of a tall-skinny matrix.
 * out[cols] = sum over rows of x[rows, cols] (row-major), deterministic (two fixed-order stages).  This is the bias
 * gradient of the set transformer's 16- and 256-wide linear layers applied to B*M rows (reference
 * set_transformer.py:24-223; autograd computes it with a generic reduction there), which a generic reduction
 * kernel runs 5-8x below the HBM rate for such shapes.  cols must divide 256 or be a multiple of 256.
 * ------------------------------------------------------------------------------------------------------------ */
size_t scae_colsum_workspace_bytes(long rows, int cols); /* 0 when the shape is not supported */
int scae_colsum(const float* x, long rows, int cols, float* out, void* workspace, size_t workspace_bytes,
                scae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Plumbing for the callers of the hot paths: fused elementwise / reduction kernels (csrc/support.cu).  None of these
 * is on a likelihood path; each replaces a string of small stock-PyTorch launches around it.  Deterministic.
 * ------------------------------------------------------------------------------------------------------------ */

/* LayerNorm over a short last dimension d in {8, 16, 32, 64} (nn.LayerNorm in the set transformer's MAB blocks,
 * reference set_transformer.py:104-116).  y = (x - mean) * rstd * gamma + beta; gamma / beta nullable (1 / 0);
 * stats[rows, 2] = {mean, rstd} (nullable, needed by the backward). */
int scae_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, long rows, int d, float* y,
                       float* stats, scae_stream_t stream);
size_t scae_layernorm_bwd_workspace_bytes(long rows, int d); /* 0 when d is not supported */
/* gx[rows, d]; g_gamma_beta[2 d] = [sum_rows g * xhat | sum_rows g]. */
int scae_layernorm_bwd(const float* g, const float* x, const float* gamma, const float* stats, long rows, int d,
                       float* gx, float* g_gamma_beta, void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* Per-channel bias and optional ReLU on an NCHW tensor y[N, C, HW], in place (nn.Conv2d bias + nn.ReLU of the part
 * encoder's Conv2dStack, reference nn_ext.py:34-59; part_encoder.py:95 att_conv bias with relu = 0). */
int scae_bias_act_fwd(float* y, const float* bias, int N, int C, int HW, int relu, scae_stream_t stream);
size_t scae_bias_act_bwd_workspace_bytes(int N, int C, int HW);
/* relu != 0: gx = g * (y > 0) (gx may alias g), g_bias[C] = sum over (N, HW) of gx.  relu == 0: only g_bias is
 * produced (the input gradient is g itself; y and gx may be NULL). */
int scae_bias_act_bwd(const float* g, const float* y, float* gx, float* g_bias, int N, int C, int HW, int relu,
                      void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* multiple_attention_pooling_2d (reference nn_ext.py:76-101): h[groups, D + 1, S] -> out[groups, D], the softmax over
 * the S <= 64 positions of every group's last channel pools the group's other D channels. */
int scae_attnpool_fwd(const float* h, float* out, long groups, int D, int S, scae_stream_t stream);
/* gh[groups, D + 1, S] from g[groups, D] (recomputes the softmax from h). */
int scae_attnpool_bwd(const float* h, const float* g, float* gh, long groups, int D, int S, scae_stream_t stream);

/* The same pooling on a CHANNELS-LAST map y[B, S, n * (D + 1)] (csrc/attnpool_cl.cu): what the part encoder's 1x1
 * attention convolution (reference part_encoder.py:95) leaves when it runs as one GEMM over the B*S positions instead of
 * B per-image products.  groups = B * n; out[groups, D].  scae_attnpool_cl_supported: 1 when the shape fits the
 * kernels' per-warp shared-memory tile (S <= 64 and S * ((D+1)|1) + 2 S + D <= 1408 floats). */
int scae_attnpool_cl_supported(long groups, int n, int D, int S);
int scae_attnpool_cl_fwd(const float* y, float* out, long groups, int n, int D, int S, scae_stream_t stream);
/* gy[B, S, n * (D + 1)] from g[groups, D] (recomputes the softmax from y). */
int scae_attnpool_cl_bwd(const float* y, const float* g, float* gy, long groups, int n, int D, int S,
                         scae_stream_t stream);

/* im2col / col2im for the part encoder's unpadded 3x3 convolutions (reference nn_ext.py:34-59; csrc/conv_cols.cu), so
 * that their passes run as cuBLAS SGEMMs on rows = (image, output position), columns = (channel, ky, kx) -- the order
 * of weight.view(C_out, C_in * 9).  stride 1 or 2; L = Ho * Wo, Ho = (H - 3) / stride + 1.  Deterministic.
 * scae_conv_cols_supported: 1 when a channel group's tile fits shared memory (e.g. 19x19 / 9x9 / 7x7 maps; not 31x31). */
int scae_conv_cols_supported(int B, int C, int H, int W, int stride);
/* x[B, C, H, W] -> cols[B * L, C * 9] */
int scae_im2col3x3(const float* x, float* cols, int B, int C, int H, int W, int stride, scae_stream_t stream);
/* dcols[B * L, C * 9] -> dx[B, C, H, W]: the adjoint of im2col (the data gradient's scatter, written as a gather) */
int scae_col2im3x3(const float* dcols, float* dx, int B, int C, int H, int W, int stride, scae_stream_t stream);
/* in[batch, R, Cc] -> out[batch, Cc, R]: NCHW <-> rows (channels-last) for the operands / results of those GEMMs. */
int scae_transpose_batched(const float* in, float* out, int batch, int R, int Cc, scae_stream_t stream);

/* One set-attention block of the object encoder (reference set_transformer.py:74-153: MAB(x, x) with single-head QKV
 * attention, residual, presence mask, LayerNorm, feed-forward + residual, LayerNorm) on x[B, N, 16], N <= 64, as one
 * kernel per direction (csrc/sab.cu).  Weights in nn.Linear layout [out, in] = [16, 16], vectors [16]. */
typedef struct scae_sab_params {
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo; /* mqkv.{q,k,v,o}_projector                              */
  const float *wf, *bf;                               /* fc                                                     */
  const float *ln0_w, *ln0_b, *ln1_w, *ln1_b;         /* ln0, ln1                                               */
  float eps0, eps1;
} scae_sab_params;
/* y[B, N, 16]; presence[B, N] nullable (no mask, no row scaling). */
int scae_sab_fwd(const float* x, const float* presence, const scae_sab_params* p, int B, int N, float* y,
                 scae_stream_t stream);
size_t scae_sab_bwd_workspace_bytes(int B, int N);
/* gx[B, N, 16]; g_params[1424] = [g_wq | g_wk | g_wv | g_wo | g_wf (256 each) | g_bq g_bk g_bv g_bo g_bf g_ln0_w
 * g_ln0_b g_ln1_w g_ln1_b (16 each)], summed over the batch in a fixed order.  presence gets no gradient. */
int scae_sab_bwd(const float* x, const float* presence, const scae_sab_params* p, const float* gy, int B, int N,
                 float* gx, float* g_params, void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* cv_ops.geometric_transform(pose, similarity, nonlinear=True, as_matrix=False) (reference cv_ops.py:20-76, as called by
 * the part encoder, part_encoder.py:110) on rows of 6 raw pose parameters.  g == NULL: out[rows, 6] = the affine
 * parameters; g != NULL (the gradient w.r.t. them): out[rows, 6] = the gradient w.r.t. t. */
int scae_pose_transform(const float* t, const float* g, float* out, long rows, int similarity, scae_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss head (csrc/loss_head.cu): the (B,O)-sized tail of SCAE.loss on the outputs of hot
 * path 2, one forward and one backward call instead of ~130 stock launches.  Replaces, with identical results,
 *   - capsule_l2_loss / capsule_entropy_loss / neg_capsule_kl (reference object_decoder.py:431-493) as called by
 *     SCAE.loss (stacked_capsule_auto_encoder.py:243-271): the prior term on caps_presence[B,O] and the posterior term on
 *     posterior_mixing_prob[B,O,V].sum(-1) / V;
 *   - the classifier heads softmax(linear(x)) on the DETACHED caps_presence and posterior mass, both through
 *     prior_classifier (sic, :203-213), and F.cross_entropy applied to their softmax OUTPUTS (sic, :279-285).
 * O <= 64, K <= 16.  Deterministic (fixed-order batch sums).  Batch statistics are those of the B rows passed in (the
 * reference's semantics under Lightning DDP); global-batch statistics: scae_loss_head_fwd_rows / _finish below.
 * ------------------------------------------------------------------------------------------------------------ */
#define SCAE_LOSS_L2 0      /* capsule_l2_loss                                   */
#define SCAE_LOSS_ENTROPY 1 /* capsule_entropy_loss, k = 1                       */
#define SCAE_LOSS_KL 2      /* neg_capsule_kl = capsule_entropy_loss with k = O  */

typedef struct scae_loss_head_args {
  const float* caps_presence; /* [B,O]                                                                          */
  const float* posterior;     /* [B,O,V] posterior_mixing_prob                                                  */
  const long long* label;     /* [B] int64 class labels; NULL: no classifier terms                              */
  const float* cls_weight;    /* [K,O] prior_classifier[0].weight (required with label)                         */
  const float* cls_bias;      /* [K]   prior_classifier[0].bias                                                 */
  int B, O, V, K;
  int sparsity;               /* 0: no sparsity terms (both prior weights are 0, stacked_capsule_auto_encoder.py:243) */
  int prior_type, posterior_type; /* SCAE_LOSS_*                                                                */
  float prior_within_weight, prior_between_weight, posterior_within_weight, posterior_between_weight;
  float prior_within_constant, posterior_within_constant; /* l2 only: O / n_classes unless overridden           */
  float between_constant;     /* l2 only: B / n_classes                                                         */
} scae_loss_head_args;

size_t scae_loss_head_workspace_bytes(const scae_loss_head_args* a); /* 0 when the shape is not supported */
/* terms[8] = {prior within, prior between, posterior within, posterior between, prior xe, posterior xe, TOTAL, 0}
 * with TOTAL = sum of weight * sparsity term + both cross-entropies (what SCAE.loss adds to its loss);
 * cls_prob[2,B,K] nullable = prior_cls_prob | posterior_cls_prob; stats[128] is what the backward needs. */
int scae_loss_head_fwd(const scae_loss_head_args* a, float* terms, float* cls_prob, float* stats, void* workspace,
                       size_t workspace_bytes, scae_stream_t stream);
/* The forward in two halves for data-parallel callers that want the between-example statistics of the GLOBAL batch
 * (SURVEY.md section 8e): rows -> colsums[132] = [sum_b caps_presence (64) | sum_b mass / V (64) | 4 row sums]; the caller
 * all-reduces colsums[0..127] over its ranks (one collective for the two O-float sums), sets between_constant to the
 * global batch / n_classes, and finishes.  scae_loss_head_fwd == rows + finish on one rank. */
int scae_loss_head_fwd_rows(const scae_loss_head_args* a, float* cls_prob, float* colsums, void* workspace,
                            size_t workspace_bytes, scae_stream_t stream);
int scae_loss_head_fwd_finish(const scae_loss_head_args* a, const float* colsums, float* terms, float* stats,
                              scae_stream_t stream);
/* g_total: DEVICE scalar, the gradient w.r.t. TOTAL.  g_caps_presence[B,O] and g_posterior[B,O,V] (nullable; written
 * only when sparsity != 0); g_cls[K*O + K] = gradient of cls_weight | cls_bias (required with label). */
int scae_loss_head_bwd(const scae_loss_head_args* a, const float* stats, const float* g_total, float* g_caps_presence,
                       float* g_posterior, float* g_cls, void* workspace, size_t workspace_bytes, scae_stream_t stream);

/* torch.optim.RMSprop (centered = False, weight_decay = 0; the reference's optimizer, base_experiment.py:47-53) over
 * FLAT buffers of n floats, one pass: square_avg = alpha square_avg + (1 - alpha) g^2; step = g / (sqrt(square_avg) +
 * eps); with momentum_buf: buf = momentum buf + step, param -= lr buf; without (NULL): param -= lr step. */
int scae_rmsprop_step(float* param, const float* grad, float* square_avg, float* momentum_buf, long n, float lr,
                      float alpha, float eps, float momentum, scae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCAE_B200_H_ */
