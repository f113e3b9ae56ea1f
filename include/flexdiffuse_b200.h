/*
 * flexdiffuse_b200 -- C ABI of the B200 (sm_100a) hot path of tim-speed/flexdiffuse.
 *
 * The reference is pure Python and has no FFI of its own, so
 * every entry point below cites the reference *Python* call site whose arithmetic it
 * replaces.  The host-side mirror of the reference API (flexdiffuse_b200/guidance.py,
 * pipeline/guide.py, pipeline/flex.py) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every `*_dev` pointer is DEVICE memory owned by the caller, outputs are
 *     caller-allocated;  pointers without the suffix are HOST memory;
 *   - all launches are asynchronous on the caller-supplied `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - re-entrant, no global mutable state besides a thread-local error string;
 *   - return value 0 = success, negative = error (see fd_last_error_string()).
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute
 *     entry point returns FD_ERR_ARCH.
 */
#ifndef FLEXDIFFUSE_B200_H
#define FLEXDIFFUSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FD_ABI_VERSION 4

/* error codes */
#define FD_OK            0
#define FD_ERR_ARG      -1   /* bad argument (shape, alignment, null pointer)       */
#define FD_ERR_ARCH     -2   /* no CUDA device / device is not sm_100               */
#define FD_ERR_CUDA     -3   /* a CUDA runtime / driver call failed                 */
#define FD_ERR_UNSUPP   -4   /* shape outside what the kernels were built for       */

/* element types */
#define FD_DTYPE_F32   0
#define FD_DTYPE_BF16  1

/* guide orders -- guidance.py:18-20 */
#define FD_GUIDE_ORDER_TEXT    0
#define FD_GUIDE_ORDER_ALIGN   1
#define FD_GUIDE_ORDER_DIRECT  2

/* per-blend status word written by fd_sim_blend */
#define FD_BLEND_OK            0
#define FD_BLEND_ZERO_DIVISION 1   /* two adjacent similarity peaks: the reference raises
                                      ZeroDivisionError in guidance.py:111-112 */
#define FD_BLEND_RANGE         2   /* a text embedding element with |x| >= 1023 (or non-finite): outside the
                                      range of the fp16 two-term split of the similarity GEMM; the row
                                      is not blended (CLIP hidden states stay below ~50)             */

/* ---- library ----------------------------------------------------------------------- */

/* ABI version of the loaded library (== FD_ABI_VERSION of the header it was built from). */
int fd_version(void);

/* Thread-local, human readable description of the last error on this thread. */
const char* fd_last_error_string(void);

/* 0 when `device` exists and is compute capability 10.x, FD_ERR_ARCH otherwise.
 * Never falls back to anything: the Python host raises on non-zero. */
int fd_arch_check(int device);

/* Number of SMs of the current device (grid sizing / bench reporting), <0 on error. */
int fd_sm_count(void);

/* ---- K4: classifier-free-guidance combine + scheduler update ------------------------ *
 * Replaces  pipeline/guide.py:61-63   eps = u + g * (c - u)
 *      and  the diffusers `scheduler.step(...)` called at pipeline/flex.py:280-285
 *      and  the LMS model-input pre-scale at pipeline/flex.py:270-274.
 * One vectorised elementwise pass:
 *      eps   = use_cfg ? u + g * (c - u) : c
 *      e'    = w[0]*eps + w[1]*h1 + w[2]*h2 + w[3]*h3        (PLMS / LMS multistep mix)
 *      x'    = a * x + b * e' + c_noise * noise             (DDIM / PLMS / LMS update)
 *      eps_out    = eps                 (optional, history for the multistep schedulers)
 *      scaled_out = x' * in_scale       (optional, next step's model input; bf16 or f32)
 * The scalar coefficients are computed on the host in fp64 by the scheduler state
 * machine (flexdiffuse_b200/schedulers.py) and rounded to fp32 here, as the reference's
 * Python-float * fp32-tensor products are.                                              */
typedef struct fd_sched_coeffs {
  float guidance;    /* g                                                    */
  int   use_cfg;     /* pipeline/guide.py:47  (guidance > 1.0)               */
  float w[4];        /* multistep weights; unused history => 0 and NULL ptr  */
  float a;           /* sample coefficient                                   */
  float b;           /* model-output coefficient                             */
  float c_noise;     /* DDIM eta>0 variance noise coefficient (0 = no noise) */
  float in_scale;    /* scaled_out = x' * in_scale                           */
} fd_sched_coeffs;

int fd_cfg_sched_step(const void*  eps_uncond_dev,  /* [n] eps_dtype, may be NULL when !use_cfg */
                      const void*  eps_cond_dev,    /* [n] eps_dtype                            */
                      int          eps_dtype,       /* FD_DTYPE_F32 | FD_DTYPE_BF16             */
                      const float* x_dev,           /* [n] current latents (or PLMS cur_sample) */
                      const float* h1_dev,          /* [n] history t-1 or NULL                  */
                      const float* h2_dev,          /* [n] history t-2 or NULL                  */
                      const float* h3_dev,          /* [n] history t-3 or NULL                  */
                      const float* noise_dev,       /* [n] or NULL                              */
                      const fd_sched_coeffs* coeffs,/* HOST pointer, copied at launch           */
                      int64_t      n_elem,          /* multiple of 4                            */
                      float*       x_out_dev,       /* [n] may alias x_dev                      */
                      float*       eps_out_dev,     /* [n] or NULL                              */
                      void*        scaled_out_dev,  /* [n] scaled_dtype or NULL                 */
                      int          scaled_dtype,
                      void*        stream);

/* ---- K9: regional noise composition --------------------------------------------------- *
 * Replaces composition/guide.py:66-87 (`CompositeGuide._guide_latents`): each entity's noise
 * prediction is lerped into the background prediction inside its rectangle (latent blocks),
 * entities in declaration order.  `eps` holds the UNet outputs for (uncond, background,
 * entity 0..E-1) on the SAME latents, NCHW.  Outputs the fp32 uncond copy and the composite
 * conditional prediction, ready for fd_cfg_sched_step.                                     */
typedef struct fd_entity_box {
  int   ox, oy;   /* EntityEmbeds.offset_blocks (x, y)   composition/embeds.py:13 */
  int   sx, sy;   /* EntityEmbeds.size_blocks   (w, h)   composition/embeds.py:14 */
  float blend;    /* EntityEmbeds.blend                   composition/embeds.py:15 */
} fd_entity_box;

int fd_composite_eps(const void* eps_dev,            /* [2 + n_entities, C, H, W] eps_dtype     */
                     int eps_dtype,
                     const fd_entity_box* boxes,     /* HOST [n_entities], <= 16                */
                     int n_entities, int C, int H, int W,
                     float* out_uncond_dev,          /* [C, H, W]                               */
                     float* out_cond_dev,            /* [C, H, W]                               */
                     void* stream);

/* ---- K1: text-token x guide-token similarity map + re-weighting + blend ------------- *
 * Replaces  guidance.py:23-85    _map_emb            (normalise, 100*cos, softmax over the
 *                                                     text tokens, header column dropped,
 *                                                     per-mode / reuse assignment)
 *           guidance.py:135-172  _clustered_guidance (+ _traverse_a_to_b :88-132)
 *           guidance.py:175-193  _blend_weights
 *           guidance.py:215-272  Tweener.tween       (avg similarity, threshold weights,
 *                                                     header cap, 3-way select / lerp)
 * for `n_text` prompts against one shared guide (guide_batch == 1) or one guide per
 * prompt (guide_batch == n_text), and `n_params` parameter sets per prompt.
 * The similarity GEMM runs on tcgen05 (kind::f16 on a two-term fp16 split of both operands,
 * three exact products => fp32-equivalent logits) with the accumulator in TMEM;
 * softmax is per guide token, the
 * column arg-max and the weight heuristics are warp-shuffle code in the same kernel.   */
typedef struct fd_tween_params {
  double threshold_floor;   /* Tweener.threshold_floor   guidance.py:205 */
  double threshold_mult;    /* Tweener.threshold_mult    guidance.py:206 */
  double clustered;         /* Tweener.clustered         guidance.py:209 */
  double max_guidance;      /* Tweener.max_guidance      guidance.py:210 */
  double header_max;        /* Tweener.header_max        guidance.py:211 */
  int    align_mode;        /* FD_GUIDE_ORDER_*          guidance.py:212 */
  int    mapping_reuse;     /* bool                      guidance.py:213 */
  int    blend_mode;        /* FD_BLEND_MODE_*: 0 = the reference's lerp (guidance.py:271);
                               1 = per-token slerp for the rows the reference lerps
                               (extension named by BASELINE.json north_star; the reference
                               has no slerp, so this mode's parity is pinned to
                               oracle/guidance_oracle.py:slerp_rows only)                 */
  int    reserved;          /* must be 0 */
} fd_tween_params;

#define FD_BLEND_MODE_LERP  0
#define FD_BLEND_MODE_SLERP 1
#define FD_SLERP_DOT_THRESHOLD 0.9995f  /* |cos| above this falls back to lerp */

int fd_sim_blend(const float* text_dev,      /* [n_text, T, D] fp32 base embeddings             */
                 const float* guide_dev,     /* [guide_batch, A, D] fp32 alt embeddings          */
                 int n_text, int guide_batch,
                 int T,                      /* text tokens  (77; 2 <= T <= 80)                  */
                 int A,                      /* guide tokens (257 image / 77 text; 1 <= A <= 384)*/
                 int D,                      /* embedding width (768; multiple of 64)            */
                 const fd_tween_params* params_dev, /* DEVICE [n_params] (8-byte aligned)           */
                 const float* linear_weights_dev,   /* [n_params, T] torch.linspace, gd.py:231   */
                 int n_params,
                 float*   out_dev,           /* [n_text, n_params, T, D] blended embeddings      */
                 float*   map_s_dev,         /* [n_text, n_params, T] similarity s  (or NULL)    */
                 int32_t* map_idx_dev,       /* [n_text, n_params, T] guide index   (or NULL)    */
                 float*   weights_dev,       /* [n_text, n_params, T] final alt_weights (or NULL)*/
                 int32_t* status_dev,        /* [n_text, n_params] FD_BLEND_*                    */
                 float*   sim_dev,           /* [n_text, A, T] softmax matrix P (debug, or NULL) */
                 void*    workspace_dev,     /* >= fd_sim_blend_workspace_bytes(...) scratch     */
                 int64_t  workspace_bytes,
                 const fd_tween_params* params_host, /* HOST copy of params_dev, or NULL.  With it the
                                                library may pick the batched kernel (many prompts, one
                                                guide, mappings with reuse / DIRECT); results are the
                                                same either way                                      */
                 void* stream);

/* Scratch needed by fd_sim_blend for the split planes and norms of the guide (an upper bound). */
int64_t fd_sim_blend_workspace_bytes(int guide_batch, int A, int D);

/* ---- K1P: CLIP visual_projection as the prologue of K1 --------------------------------- *
 * Replaces encode/clip.py:100  `clip.visual_projection(hidden_states)`  (Linear 1024 -> 768, no
 * bias, on all 257 post-layernorm vision tokens): out[M,N] = hidden[M,K] . weight[N,K]^T in fp32
 * accuracy (two-term fp16 split on tcgen05, like K1's similarity GEMM), so the guide embeddings K1
 * consumes never leave the hand-written path.  The weight planes live in `workspace_dev`; pass
 * weights_changed != 0 on the first call and whenever the weight tensor changes.             */
int64_t fd_visual_projection_workspace_bytes(int N, int K);
int fd_visual_projection(const float* hidden_dev,   /* [M, K] fp32, M = images * 257               */
                         const float* weight_dev,   /* [N, K] fp32 visual_projection.weight        */
                         float*       out_dev,      /* [M, N] fp32                                 */
                         int M, int N, int K,       /* K % 64 == 0, N % 4 == 0                     */
                         void* workspace_dev, int64_t workspace_bytes, int weights_changed,
                         void* stream);
/* 1 if an operand since the last call left the split's range (|activation| >= 1023, |weight| >= 63
 * or non-finite), clearing the flag; synchronises the device.                                 */
int fd_visual_projection_range_flag(void);

/* ---- K11: fp32-accurate Linear on tcgen05 for the CLIP towers --------------------------- *
 * Replaces the fp32 nn.Linear layers (q / k / v / out_proj, fc1, fc2) of the ViT-L/14 vision
 * tower and of the text tower that encode/clip.py:57-65 (`clip.text_model(ids)`) and :86-100
 * (`clip.vision_model...`) run through transformers' CLIPModel:
 *     out[M, N] = act( x[M, K] . W[N, K]^T + bias[N] )        fp32 in, fp32 out
 * Every operand row is scaled by a power of two chosen from its maximum and split into two fp16
 * terms; three exact fp16 products accumulate in fp32 in TMEM, so the result has fp32 accuracy
 * (no tf32 / bf16 rounding of the operands) at tensor-core speed.
 *   1. fd_linear_x3_split     fp32 [rows, K] -> operand buffer (weights once per version,
 *                             activations once per call; one buffer may feed several Linears)
 *   2. fd_linear_x3           the GEMM + bias + activation on two operand buffers             */
#define FD_LINEAR_ACT_NONE       0
#define FD_LINEAR_ACT_QUICK_GELU 1   /* x * sigmoid(1.702 x): CLIP's hidden_act               */
#define FD_LINEAR_ACT_GELU       2   /* erf GELU                                              */
int64_t fd_linear_x3_operand_bytes(int rows, int K);
int fd_linear_x3_split(const float* x_dev,      /* [rows, K] fp32 row-major, K % 64 == 0, <= 8192 */
                       int rows, int K,
                       void* operand_dev,       /* >= fd_linear_x3_operand_bytes, 256-B aligned  */
                       int64_t operand_bytes, void* stream);
/* LayerNorm fused with the split: operand of y = LayerNorm(x) * gamma + beta (y itself is never written). */
int fd_linear_x3_split_ln(const float* x_dev, int rows, int K,   /* K % 64 == 0, K <= 8192              */
                          const float* gamma_dev, const float* beta_dev, float eps,
                          void* operand_dev, int64_t operand_bytes, void* stream);
int fd_linear_x3(const void* act_operand_dev,   /* split [M, K]                                  */
                 int M,
                 const void* weight_operand_dev,/* split [N, K]                                  */
                 int N, int K,                  /* N % 4 == 0                                    */
                 const float* bias_dev,         /* [N] fp32 or NULL                              */
                 int act,                       /* FD_LINEAR_ACT_*                               */
                 const float* residual_dev,     /* [M, N] fp32 added after the activation, or NULL */
                 float* out_dev,                /* [M, N] fp32                                   */
                 int split_k,                   /* 1, 2, 4 or 8: K is cut into this many slices, one CTA of a thread-block
                                                   cluster each; the slices are added in ascending order over DSMEM
                                                   (bit-reproducible, no scratch in global memory)                  */
                 void* stream);
/* 1 if a non-finite operand was split since the last call, >= 16 if a pipeline wait gave up;
 * clears the flag; synchronises the device.                                                 */
int fd_linear_x3_flag(void);

/* ---- K12: exact-fp32 attention core for the CLIP towers' short sequences ---------------- *
 * Replaces the fp32 attention inside transformers' CLIPAttention (reached from encode/clip.py:57-65
 * and :86-100): out[b,t,h,:] = softmax_j(scale <q[b,t,h,:], k[b,j,h,:]> [+ causal mask]) . v[b,:,h,:].
 * q / k / v are [B * T] rows of `row_stride` floats with head h at column h * d (e.g. the three
 * column blocks of one fused q|k|v GEMM output); out is [B * T, H * d] contiguous.             */
int fd_attention_f32(const float* q_dev, const float* k_dev, const float* v_dev,
                     int64_t row_stride,            /* floats between consecutive tokens          */
                     float* out_dev,
                     int B, int T,                  /* T <= 288                                   */
                     int H, int d,                  /* d % 4 == 0, d <= 128                       */
                     float scale, int causal, void* stream);

/* ---- K13: feed-forward GEGLU projection with the activation in the GEMM epilogue -------- *
 * Replaces `GEGLU.proj` (nn.Linear(C, 8C)) + `hidden * F.gelu(gate)` of diffusers' FeedForward in every
 * BasicTransformerBlock of the UNet call at pipeline/guide.py:56-58:
 *     out[m, f] = (x[m,:] . w[f,:] + b[f]) * gelu(x[m,:] . w[F + f,:] + b[F + f])
 * tcgen05 cta_group::2 GEMM; the [M, 2F] projection never reaches memory.                       */
int fd_ff_geglu(const void* x_bf16_dev,     /* [M, K] bf16 row-major tokens                     */
                const void* w_bf16_dev,     /* [2F, K] bf16: value rows, then gate rows           */
                const void* bias_bf16_dev,  /* [2F] bf16                                          */
                void*       out_bf16_dev,   /* [M, F] bf16 (rows out_row_stride elements apart)    */
                int M, int F, int K,        /* F % 128 == 0, K % 64 == 0                          */
                int64_t out_row_stride,     /* 0 = F; % 8 == 0: the output may be a column block   *
                                             * of a wider matrix (see UNet "merged output GEMM")    */
                void* stream);

/* ---- K2: cross-attention K/V projection of the fixed context, hoisted out of the loop *
 * Replaces the 32 bias-free `to_k(context)` / `to_v(context)` Linears that diffusers'
 * CrossAttention.forward recomputes for each of the 16 attn2 layers at every step,
 * reached from pipeline/guide.py:56-58.   out[M,N] = ctx[M,K] . w[N,K]^T   (bf16 in,
 * fp32 accumulate in TMEM, bf16 out) -- TMA-fed tcgen05 GEMM.  `w` is the row-wise
 * concatenation of all to_k / to_v weights (N = 2 * sum(C_l) = 24960 for SD v1).        */
int fd_kv_project(const void* ctx_bf16_dev,  /* [M, K] row-major bf16, M = n_ctx * T_pad */
                  const void* w_bf16_dev,    /* [N, K] row-major bf16                     */
                  void*       out_bf16_dev,  /* [M, N] row-major bf16                     */
                  int M, int N, int K,       /* K % 64 == 0, N % 8 == 0                   */
                  void* stream);

/* ---- K3: cross-attention over the cached K/V ----------------------------------------- *
 * Replaces diffusers' CrossAttention._attention (softmax(Q K^T * scale) V over the 77
 * context tokens, no mask), reached from pipeline/guide.py:56-58 for each attn2 layer.
 * q / out are [n_samples, n_q, heads*d_head] bf16; K and V are column slices of the
 * K2 output: for sample s, rows ctx_index[s]*t_pad .. +t_pad-1, columns
 * k_col_off + h*d_head .. and v_col_off + h*d_head ..                                   */
int fd_cross_attn(const void* q_bf16_dev,
                  const void* kv_bf16_dev,      /* K2 output [n_ctx*t_pad, kv_row_stride]  */
                  int64_t kv_rows,              /* n_ctx * t_pad                            */
                  int64_t kv_row_stride,        /* elements per cache row (N of K2)         */
                  int k_col_off, int v_col_off, /* element offsets, multiples of 8          */
                  const int32_t* ctx_index_dev, /* [n_samples] context id of each sample    */
                  int n_samples, int n_q, int heads, int d_head, /* d_head in {40,80,160}   */
                  int t_valid,                  /* 77: keys >= t_valid are masked           */
                  int t_pad,                    /* 80                                       */
                  float scale,                  /* d_head ** -0.5                           */
                  void* out_bf16_dev,
                  void* stream);

/* ---- K3F: the whole attn2 layer in one launch ------------------------------------------ *
 * Replaces diffusers' CrossAttention.forward for the 16 attn2 layers reached from
 * pipeline/guide.py:56-58:   out = to_out( softmax( to_q(x) K^T * scale ) V ) + bias
 * with K / V from the K2 cache (fd_kv_project), i.e. the to_q GEMM, the attention and the
 * to_out GEMM (+ bias) that were three launches (cuBLAS, fd_cross_attn, cuBLAS) become one
 * TMA-fed tcgen05 kernel: Q lives only in TMEM, the attention output crosses L2 once
 * (`attn_bf16_dev`, also returned for inspection).  SURVEY 8d "fused variant": tensor-bound.
 * x / attn / out are [n_samples, n_q, C] bf16 with C = heads*d_head a multiple of 320;
 * wq / wo are the nn.Linear weights [C_out, C_in] row-major; bo the to_out bias [C].        */
int fd_cross_attn_fused(const void* x_bf16_dev,       /* [n_samples, n_q, C] hidden states    */
                        const void* wq_bf16_dev,      /* [C, C] to_q.weight                     */
                        const void* kv_bf16_dev,      /* K2 output [n_ctx*t_pad, kv_row_stride] */
                        int64_t kv_rows, int64_t kv_row_stride,
                        int k_col_off, int v_col_off, /* element offsets, multiples of 8        */
                        const int32_t* ctx_index_dev, /* [n_samples] context id of each sample  */
                        const void* wo_bf16_dev,      /* [C, C] to_out[0].weight                */
                        const void* bo_bf16_dev,      /* [C]    to_out[0].bias                  */
                        int n_samples, int n_q, int heads, int d_head, /* d_head in {40,80,160} */
                        int t_valid, int t_pad,       /* 77, 80                                 */
                        float scale,                  /* d_head ** -0.5                         */
                        void* attn_bf16_dev,          /* [n_samples, n_q, C] softmax(QK^T)V     */
                        void* out_bf16_dev,           /* [n_samples, n_q, C] layer output       */
                        void* stream);

/* ---- K5 / K6: normalisation + activation glue of the UNet forward --------------------- *
 * The UNet call at pipeline/guide.py:56-58 spends ~45 % of a B=1 denoising step in ATen's
 * GroupNorm path (layout copy, moments, params, apply, separate SiLU, separate time-embedding
 * add) and the GEGLU gelu/mul pair (profiles/r01/SUMMARY.md).  Fused, NHWC, bf16:
 *   y = act( GroupNorm(x + bias[n,c]) * gamma[c] + beta[c] ),  act = SiLU or identity
 * replaces diffusers ResnetBlock2D's  norm1 -> SiLU  and  (+ time_emb_proj) -> norm2 -> SiLU,
 * SpatialTransformer.norm and conv_norm_out -> SiLU.                                       */
int64_t fd_groupnorm_act_workspace_bytes(int N, int HW, int C, int G); /* size of `workspace_dev` */

int fd_groupnorm_act(const void* x_bf16_dev,     /* [N, HW, C] channels-last activations    */
                     const void* bias_bf16_dev,  /* [N, C] rows added before the norm, or NULL */
                     const void* gamma_bf16_dev, /* [C]                                     */
                     const void* beta_bf16_dev,  /* [C]                                     */
                     void*       workspace_dev,  /* fd_groupnorm_act_workspace_bytes(...),
                                                    zero-filled ONCE before its first use    */
                     void*       y_bf16_dev,     /* [N, HW, C]                              */
                     int N, int HW, int C, int G,/* G <= 32, C/G even, C % 64 == 0          */
                     float eps, int act_silu,
                     int64_t bias_row_stride,    /* elements between bias rows (>= C, even) */
                     void* stream);

/* K7 folded into the K5 that consumes its result: sum_out = x + h + rbias[c] (ResnetBlock2D's residual add, as
 * fd_add_bias_residual) and y = act(GroupNorm(sum_out) * gamma + beta) (the `norm` of the SpatialTransformer or the `norm1`
 * of the ResnetBlock2D that follows) in ONE launch where the cluster kernel applies; otherwise K7 + the streaming
 * GroupNorm.  Same arguments as fd_groupnorm_act without the per-(n, c) bias.                                              */
int fd_add_groupnorm_act(const void* x_bf16_dev, const void* h_bf16_dev,   /* [N, HW, C] each            */
                         const void* rbias_bf16_dev,                       /* [C]                        */
                         void* sum_out_bf16_dev,                           /* [N, HW, C]                 */
                         const void* gamma_bf16_dev, const void* beta_bf16_dev, void* workspace_dev,
                         void* y_bf16_dev, int N, int HW, int C, int G, float eps, int act_silu, void* stream);

/* y = x + h + bias[c]: ResnetBlock2D's residual add with conv2's bias folded in (NHWC bf16).
 * h == NULL: y = x + bias[c] (the bias of conv_in / Downsample2D / Upsample2D convolutions); y may alias x. */
int fd_add_bias_residual(const void* x_bf16_dev, const void* h_bf16_dev /* or NULL */,
                         const void* bias_bf16_dev,  /* [C]                                 */
                         void* y_bf16_dev, int64_t n_elem, int C, void* stream);

/* y[p, 0:Ca] = a[p, :], y[p, Ca:Ca+Cb] = b[p, :] on channels-last bf16 tensors: the skip-connection
 * `torch.cat([hidden, skip], dim=1)` of diffusers' up blocks inside the UNet call at pipeline/guide.py:56-58. */
int fd_concat_channels(const void* a_bf16_dev,   /* [pixels, Ca]                           */
                       const void* b_bf16_dev,   /* [pixels, Cb]                           */
                       void* y_bf16_dev,         /* [pixels, Ca + Cb]                      */
                       int64_t pixels, int Ca, int Cb, /* Ca % 8 == 0, Cb % 8 == 0         */
                       void* stream);

/* Nearest-neighbour 2x upsample, channels-last bf16 [N,H,W,C] -> [N,2H,2W,C]: `F.interpolate(x, scale_factor=2.0,
 * mode="nearest")` of diffusers' Upsample2D inside the UNet call at pipeline/guide.py:56-58. */
int fd_upsample_nearest2x(const void* x_bf16_dev, void* y_bf16_dev, int N, int H, int W,
                          int C /* % 8 == 0 */, void* stream);

/* s = x + y ; n = LayerNorm(s) * gamma + beta   (BasicTransformerBlock: the residual add of one
 * attention / feed-forward branch fused with the LayerNorm feeding the next one).  y == NULL:
 * plain LayerNorm of x.  Rows of C in {320, 640, 1280} bf16.                                 */
int fd_add_layernorm(const void* x_bf16_dev,       /* [M, C]                                  */
                     const void* y_bf16_dev,       /* [M, C] or NULL                          */
                     const void* gamma_bf16_dev, const void* beta_bf16_dev,   /* [C]          */
                     void* sum_out_bf16_dev,       /* [M, C] x + y, or NULL                   */
                     void* norm_out_bf16_dev,      /* [M, C]                                  */
                     int64_t M, int C, float eps,
                     int64_t sum_out_row_stride,   /* elements between rows of sum_out (0 = C):
                                                      the sum may land in a column block of a
                                                      wider matrix                             */
                     void* stream);

/* diffusers GEGLU: out[m, f] = in[m, f] * gelu(in[m, F + f]), exact (erf) GELU.           */
int fd_geglu(const void* in_bf16_dev,            /* [M, 2F]                                 */
             void*       out_bf16_dev,           /* [M, F]                                  */
             int64_t M, int F,                   /* F % 8 == 0                              */
             void* stream);

/* ---- K10: image post-processing tail of the VAE decode ---------------------------------- *
 * Replaces pipeline/flex.py:119-124 `(image / 2 + 0.5).clamp(0, 1)` -> `.cpu().permute(0,2,3,1)`
 * -> numpy_to_pil's `(images * 255).round().astype('uint8')`.  `x` is the decoder output in
 * channels-last memory order ([B, H, W, 3] flat); out[i] = round(clamp(x[i]/2 + 0.5, 0, 1) * 255)
 * with the reference's rounding (tensor-dtype ops, fp32 product, round-half-even).          */
int fd_image_tail_u8(const void* x_dev,      /* [n_elem] f32 or bf16, NHWC order               */
                     int dtype,              /* FD_DTYPE_F32 / FD_DTYPE_BF16                   */
                     int64_t n_elem,
                     void* out_u8_dev,       /* [n_elem] uint8                                 */
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLEXDIFFUSE_B200_H */
