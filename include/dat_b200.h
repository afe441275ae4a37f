/*
 * dat_b200 — C ABI of the B200-native deformable-attention block (DAT / DAT++).
 *
 * The reference (hehe717/DAT-Segmentation) is pure Python and has no FFI: the block
 * is `DAttentionBaseline.forward` (models/utils/dat_blocks.py:138-227), a straight
 * line of ~20 PyTorch operator calls.  This header is the boundary a maintainer binds
 * instead (ctypes stub in INTEGRATION.md): every entry point names the reference
 * lines it replaces.
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (PyTorch's allocator);
 *     nothing here allocates, frees or keeps a pointer after returning.
 *   - Activations are channel-last: x (B, H, W, C) == (B, HW, C), C = n_heads * 32.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it
 *     (CUDA-graph capturable; no host synchronisation inside).
 *   - Return value: DAT_OK (0) or a negative DAT_ERR_* code; never aborts.
 *     `dat_last_error()` returns a static, thread-local message for the last failure.
 *   - Re-entrant.  Global state: a launch counter (statistics) and, per device, ONE lazily
 *     created side stream with its events, used only inside dat_block_backward (weight /
 *     table gradients run next to the data-gradient chain) under a per-device mutex and
 *     joined back into `stream` before the call returns, on every exit path.
 *     DAT_B200_SERIAL_WGRAD=1 disables it.
 *   - Determinism: every reduction of the bf16 tensor-core path (the path bench.py times)
 *     has a fixed order, including d rpe_table (per-CTA partial tables + ordered sum): two
 *     runs give bit-identical outputs and gradients.  The fp32 / generic-shape CUDA-core
 *     attention backward and the in-kernel table scatter (DAT_B200_TABLE_SCATTER=1, maps
 *     the GEMM form does not cover) accumulate d rpe_table with fp32 atomics: those values
 *     can differ in the last bits from run to run.
 *   - dtype codes: DAT_F32 = 0, DAT_BF16 = 1.  Parameters and all gradients of
 *     parameters are fp32.  pos / lse / offsets are always fp32.
 */
#ifndef DAT_B200_H_
#define DAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAT_OK 0
#define DAT_ERR_ARG (-1)      /* bad shape / null pointer / unsupported configuration */
#define DAT_ERR_CUDA (-2)     /* a CUDA runtime call or kernel launch failed            */
#define DAT_ERR_UNSUPPORTED (-3)

#define DAT_F32 0
#define DAT_BF16 1

#define DAT_HEAD_DIM 32       /* n_head_channels; 32 in every DAT++ variant (dat.py:57) */

/* Position-encoding branch of the block (dat_blocks.py:84-104 ctor, :183-214,221-222 forward), in the
 * reference's precedence order: `use_pe and not no_off`, then dwc_pe / fixed_pe / log_cpb / rpe_table. */
#define DAT_PE_RPE 0          /* bilinear sample of rpe_table (n_heads, 2q_h-1, 2q_w-1) at (q_grid - pos)/2 */
#define DAT_PE_NONE 1         /* use_pe = False (or no_off): no bias                                        */
#define DAT_PE_DWC 2          /* dwc_pe: depthwise 3x3 conv of q added to the attention output (:185,221)    */
#define DAT_PE_FIXED 3        /* fixed_pe: dense table (n_heads, q_h*q_w, kv_h*kv_w), bilinearly resized
                                 (align_corners) to (HW, Ns) (:187-191)                                      */
#define DAT_PE_LOGCPB 4       /* log_cpb: Linear(2,32)-ReLU-Linear(32,hg) on sign(d)log2(|d|+1)/3,
                                 d = 4 (q_grid - pos) (:192-197)                                             */

/* Shape / hyper-parameters of one block call.  Mirrors the constructor arguments of
 * DAttentionBaseline (dat_blocks.py:21-50) plus the runtime (B, H, W). */
typedef struct dat_block_desc {
  int32_t B, H, W;            /* runtime query map; C = n_heads * 32                    */
  int32_t n_heads;            /* dat_blocks.py:33                                        */
  int32_t n_groups;           /* :38; Cg = C / n_groups, hg = n_heads / n_groups         */
  int32_t stride, ksize;      /* :46-47; pad = ksize/2 if ksize != stride else 0 (:50)   */
  int32_t table_h, table_w;   /* rpe_table spatial size 2*q_h-1, 2*q_w-1 (:101-103)      */
  float offset_range_factor;  /* :45; < 0 selects the clamp(-1, 1) branch (:159-162)     */
  int32_t x_dtype;            /* dtype of x and of dx                                    */
  int32_t act_dtype;          /* dtype of q, xs, k, v, o, y, dy and of d{o,q,k,v,xs}     */
  int32_t pe_mode;            /* DAT_PE_*; for DAT_PE_FIXED table_h = q_h*q_w, table_w = kv_h*kv_w;
                                 table_h/table_w are ignored by NONE / DWC / LOGCPB        */
  int32_t no_off;             /* :57-59,156-157,164-167: keys/values from avg_pool2d(x, stride), Hk = H/stride,
                                 offset network unused (its gradients are not written), no bias */
} dat_block_desc;

/* The 14 reference parameters (state-dict names in comments), fp32. */
typedef struct dat_block_params {
  const float* off_dw_w;   /* conv_offset.0.weight      (Cg, 1, k, k) */
  const float* off_dw_b;   /* conv_offset.0.bias        (Cg)          */
  const float* off_ln_g;   /* conv_offset.1.norm.weight (Cg)          */
  const float* off_ln_b;   /* conv_offset.1.norm.bias   (Cg)          */
  const float* off_pw_w;   /* conv_offset.3.weight      (2, Cg, 1, 1) */
  const float* wq; const float* bq;   /* proj_q.weight (C, C, 1, 1), proj_q.bias (C) */
  const float* wk; const float* bk;   /* proj_k.*                                     */
  const float* wv; const float* bv;   /* proj_v.*                                     */
  const float* wo; const float* bo;   /* proj_out.*                                   */
  const float* rpe_table;  /* RPE / FIXED: rpe_table (n_heads, table_h, table_w); DWC: rpe_table.weight (C,1,3,3);
                              LOGCPB: rpe_table.0.weight (32, 2); NONE: ignored (may be NULL)               */
  const float* pe_b;       /* DWC: rpe_table.bias (C); LOGCPB: rpe_table.0.bias (32); else ignored          */
  const float* pe_w2;      /* LOGCPB: rpe_table.2.weight (hg, 32); else ignored                             */
  /* Optional (may be NULL; bf16 activations only): bf16 copies (C, C) of proj_q / proj_k / proj_v / proj_out
   * .weight made by the caller once per step (dat_cast_bf16_multi).  The forward reads them as K-major and the
   * backward as MN-major tensor-core operands; when NULL the library casts into its workspace per call. */
  const void* wq_bf16; const void* wk_bf16; const void* wv_bf16; const void* wo_bf16;
} dat_block_params;

/* Gradients of the same 14 tensors (fp32, OVERWRITTEN, not accumulated). */
typedef struct dat_block_grads {
  float* off_dw_w; float* off_dw_b; float* off_ln_g; float* off_ln_b; float* off_pw_w;
  float* wq; float* bq; float* wk; float* bk; float* wv; float* bv; float* wo; float* bo;
  float* rpe_table; float* pe_b; float* pe_w2;   /* as in dat_block_params; unused ones may be NULL */
} dat_block_grads;

/* Activations kept between forward and backward; caller allocates each buffer.
 * Ns = Hk * Wk, G = n_groups. */
typedef struct dat_block_saved {
  void* q;          /* (B, HW, C)      act_dtype                                   */
  float* t_dw;      /* (B, G, Ns, Cg)  fp32  depthwise-conv output, pre-LayerNorm  */
  float* off_raw;   /* (B, G, Ns, 2)   fp32  output of conv_offset.3 (dy, dx)      */
  float* pos;       /* (B, G, Ns, 2)   fp32  sampling positions (y, x) in [-1, 1]  */
  void* xs;         /* (B, Ns, C)      act_dtype  sampled features                 */
  void* k;          /* (B, Ns, C)      act_dtype                                   */
  void* v;          /* (B, Ns, C)      act_dtype                                   */
  void* o;          /* (B, HW, C)      act_dtype  attention output before proj_out */
  float* lse;       /* (B, n_heads, HW) fp32 log-sum-exp of the score rows         */
} dat_block_saved;

/* ---- queries ------------------------------------------------------------------- */

/* Sample-grid size of the offset network's strided depthwise conv (dat_blocks.py:52,146). */
int dat_sample_grid(const dat_block_desc* d, int32_t* Hk, int32_t* Wk);

/* Bytes of scratch `dat_block_forward` / `dat_block_backward` need (0 is possible). */
size_t dat_block_fwd_workspace_bytes(const dat_block_desc* d);
size_t dat_block_bwd_workspace_bytes(const dat_block_desc* d);

const char* dat_last_error(void);
/* Build identification: "dat_b200 <git-free version> sm_100a". */
const char* dat_version(void);
/* Number of kernels this library has launched in this process (statistics; bench.py). */
uint64_t dat_launch_count(void);

/* ---- whole block (replaces DAttentionBaseline.forward, dat_blocks.py:138-227) ---- */

int dat_block_forward(const dat_block_desc* d, const dat_block_params* p,
                      const void* x,            /* (B, HW, C) x_dtype   */
                      void* y,                  /* (B, HW, C) act_dtype */
                      const dat_block_saved* s, /* all buffers filled   */
                      void* workspace, size_t workspace_bytes, void* stream);

/* Autograd of the above (what torch.autograd derives for the reference). */
int dat_block_backward(const dat_block_desc* d, const dat_block_params* p,
                       const void* x, const void* dy, const dat_block_saved* s,
                       float* dx,                /* (B, HW, C) fp32, overwritten */
                       const dat_block_grads* g,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- individual stages (unit-parity entry points) ------------------------------- */

/* 1x1 conv + bias, channel-last: Y[M,N] = X[M,K] W[N,K]^T + b[N].
 * Replaces proj_q/proj_k/proj_v/proj_out (dat_blocks.py:143,177-178,225). */
int dat_pointwise_fwd(const void* X, int32_t x_dtype, const float* W, const float* b,
                      void* Y, int32_t y_dtype, int64_t M, int32_t N, int32_t K, void* stream);

/* Same contraction on the tcgen05 tensor cores (TMA-fed, accumulator in tensor memory).
 * x_dtype DAT_F32: tf32 MMA straight on fp32 X and fp32 W (no cast pass);
 * x_dtype DAT_BF16: bf16 MMA, W must then point to a bf16 copy of the weight
 * (dat_cast_bf16).  Returns DAT_ERR_UNSUPPORTED for shapes it cannot tile
 * (the block driver then uses the CUDA-core path). */
int dat_pointwise_fwd_tc(const void* X, int32_t x_dtype, const void* W, const float* b,
                         void* Y, int32_t y_dtype, int64_t M, int32_t N, int32_t K, void* stream);
/* The same product with the residual add that follows it in the MLP branch (`x = drop_path(mlp(x)) + x`, dat.py:151-156)
 * in the epilogue: Y (fp32) = resid + scale[row / rows_per_sample] * bf16(X W^T + b).  resid, Y: (M, N) fp32; scale: one
 * float per sample (mask / keep_prob).  N must tile in 64-column groups (DAT_ERR_UNSUPPORTED otherwise). */
int dat_pointwise_fwd_tc_residual(const void* X, int32_t x_dtype, const void* W, const float* b, const float* resid,
                                  const float* scale, int64_t rows_per_sample, float* Y, int64_t M, int32_t N, int32_t K,
                                  void* stream);
/* Debug aid: globaltimer (ns) phase stamps of CTA (0,0) of the last dat_pointwise_fwd_tc
 * launch: entry, setup done, first TMA stage landed, MMAs issued, accumulator ready,
 * epilogue done (6 of 8 slots used).  Synchronises the device. */
int dat_debug_gemm_timing(uint64_t* out8);
/* Debug aid: SM-cycle phase counters of the tensor-core attention backward since the last call
 * (summed over CTAs): [0] tile loop, [1] wait for S/dP, [2] score loop, [3] d pos column sums,
 * [4] dQ wait + store, [5] per-tile setup, [6] number of CTAs.  Synchronises the device. */
int dat_debug_attn_bwd_timing(uint64_t* out8);
/* Tensor-core gradients of the 1x1 convolution (bf16 operands, fp32 accumulation):
 *   data gradient  dX[M,K] = dY[M,N] W[N,K]  ==  dat_pointwise_fwd_tc(dY, W^T) with the (K,N) bf16
 *                  transposed weight from dat_cast_transpose_bf16;
 *   weight gradient dW[N,K] = dY^T X (both bf16, read MN-major; N, K multiples of 64; deterministic
 *                  split reduction); the bias gradient db[N] = column sums of dY comes out of the
 *                  same pass over dY when db != NULL (an extra N = 16 MMA against a tile of ones).
 *                  dat_bias_grad is the stand-alone column sum. */
int dat_cast_transpose_bf16(const float* w, void* out, int32_t N, int32_t K, void* stream);
size_t dat_pointwise_wgrad_tc_workspace_bytes(int64_t M, int32_t N, int32_t K);
int dat_pointwise_wgrad_tc(const void* dY, const void* X, float* dW, float* db, int64_t M, int32_t N,
                           int32_t K, void* workspace, size_t workspace_bytes, void* stream);
/* Data gradient with the weight read in place: dX[M,K] (dx_dtype) = dY[M,N] W[N,K], dY and W bf16.  W is the
 * SAME bf16 copy the forward used (dat_cast_bf16): the tensor core takes it as an MN-major operand, so no
 * transposed copy exists.  DAT_ERR_UNSUPPORTED when K has no tile width that is a multiple of 64. */
int dat_pointwise_dgrad_tc(const void* dY, const void* W, void* dX, int32_t dx_dtype, int64_t M, int32_t N,
                           int32_t K, void* stream);
/* workspace >= 64 * N * 4 bytes */
int dat_bias_grad(const void* dY, int32_t dy_dtype, float* db, int64_t M, int32_t N, void* workspace,
                  size_t workspace_bytes, void* stream);
/* fp32 -> bf16 copy of n elements (n % 4 == 0), used for the weight operands above. */
int dat_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
/* The same for a whole table of tensors in ONE launch (every 1x1-conv weight of a model, once per step).
 * `items` is a DEVICE array of n_items entries; every n is a multiple of 4, pointers 16-byte aligned. */
typedef struct dat_cast_item {
  const float* src;
  void* dst;        /* bf16 */
  int64_t n;
} dat_cast_item;
int dat_cast_bf16_multi(const dat_cast_item* items, int32_t n_items, void* stream);

/* Offset network + reference points + range/clamp → pos
 * (dat_blocks.py:144-162 and _get_ref_points :108-121).
 * q (B,HW,C) act_dtype → t_dw, off_raw, pos as in dat_block_saved. */
int dat_offset_pos_fwd(const dat_block_desc* d, const dat_block_params* p, const void* q,
                       float* t_dw, float* off_raw, float* pos, void* stream);

/* Reference points alone, (Hk) + (Wk) floats: ((i + .5) / (n - 1)) * 2 - 1, bit-exact
 * with dat_blocks.py:111-118. */
int dat_ref_points(int32_t Hk, int32_t Wk, float* ref_y, float* ref_x, void* stream);

/* Bilinear gather of x at pos (F.grid_sample bilinear/zeros/align_corners=True,
 * dat_blocks.py:169-172).  taps (optional, may be NULL): (B, G, Ns, 2) int32 = (y0, x0)
 * north-west integer tap of every sample, for index bit-exactness tests. */
int dat_sample_fwd(const dat_block_desc* d, const void* x, const float* pos, void* xs,
                   int32_t* taps, void* stream);

/* The gather fused with the k / v projections (dat_blocks.py:169-178): xs = grid_sample(x, pos) (bf16, bit-identical
 * to dat_sample_fwd), k = xs Wk^T + bk, v = xs Wv^T + bv (bf16) in ONE launch - the sampled tile is built in shared
 * memory as the A operand of the tcgen05 GEMMs.  act_dtype must be DAT_BF16, C in {64, 128, 256, 512}, C / n_groups a
 * multiple of 8; wk_bf16 / wv_bf16 are bf16 (C, C) copies of the weights; bk / bv fp32 (may be NULL).
 * dat_block_forward uses it where it is faster than the three launches (C <= 128; DAT_B200_GATHER_KV_FUSION=0/1). */
int dat_gather_kv_fwd(const dat_block_desc* d, const void* x, const float* pos, const void* wk_bf16,
                      const void* wv_bf16, const float* bk, const float* bv, void* xs, void* k, void* v, void* stream);

/* QK^T*scale + bilinear rpe bias + softmax + PV (dat_blocks.py:180-223).
 * act_dtype DAT_BF16 with Ns in {64,128,256} runs the tcgen05 kernel and needs
 * `dat_attention_fwd_workspace_bytes` of scratch (packed rpe table); any other case runs
 * the fp32 CUDA-core kernel (workspace may be NULL).  `impl`: 0 = automatic,
 * 1 = force the CUDA-core kernel. */
size_t dat_attention_fwd_workspace_bytes(const dat_block_desc* d);
int dat_attention_fwd(const dat_block_desc* d, const void* q, const void* k, const void* v,
                      const float* pos, const float* rpe_table, void* o, float* lse,
                      void* workspace, size_t workspace_bytes, int32_t impl, void* stream);

/* ---- "next" row: the LayerNorm that feeds the block ------------ */

/* Channel-last LayerNorm over C, eps inside the sqrt, affine (LayerNormProxy,
 * dat_blocks.py:229-240; used at dat.py:147,151).  x (rows, C) -> y (rows, C); mean / rstd
 * (rows) fp32 are saved for the backward.  C even, <= 1024. */
int dat_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta,
                      void* y, int32_t y_dtype, float* mean, float* rstd, int64_t rows, int32_t C,
                      float eps, void* stream);
size_t dat_layernorm_bwd_workspace_bytes(int64_t rows, int32_t C);
/* dx has x's dtype; dgamma / dbeta (C) fp32 are overwritten (deterministic reduction).
 * dres (may be NULL; x's dtype, (rows, C)): the gradient that reaches x through the residual path
 * around the norm (dat.py:147-156) - it is added into dx in the same pass. */
int dat_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype,
                      const float* gamma, const float* mean, const float* rstd, void* dx,
                      const void* dres, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The residual add that precedes a norm, fused into it (dat.py:147-151: `x = drop_path(attn) + x` followed by
 * `layer_norms[2d+1](x)`): xout = x + a * scale[row / rows_per_sample] (x's dtype: the new residual stream),
 * y = LayerNorm(xout).  a has y's dtype; scale is a device array of rows / rows_per_sample floats (mask / keep_prob).
 * One pass over the stream instead of dat_scale_residual + dat_layernorm_fwd. */
int dat_residual_layernorm_fwd(const void* a, const float* scale, int64_t rows_per_sample, const void* x,
                               int32_t x_dtype, const float* gamma, const float* beta, void* xout, void* y,
                               int32_t y_dtype, float* mean, float* rstd, int64_t rows, int32_t C, float eps,
                               void* stream);
/* Its backward: dx = LayerNorm-backward(dy) + dres (gradient w.r.t. xout, which is also the gradient w.r.t. x) and the
 * branch gradient da = dx * scale[row / rows_per_sample] (dy's dtype) in the same pass.  x is the saved xout. */
int dat_residual_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* gamma,
                               const float* mean, const float* rstd, void* dx, const void* dres, void* da,
                               const float* scale, int64_t rows_per_sample, float* dgamma, float* dbeta, int64_t rows,
                               int32_t C, void* workspace, size_t workspace_bytes, void* stream);

/* Residual add with stochastic depth (dat.py:147-156, `x = drop_path(branch) + x`):
 * y[b, :] = x[b, :] + a[b, :] * scale[b]; x may be NULL ('X' blocks, dat.py:140-144).  a, x, y are
 * dense tensors of one layout whose outermost dimension is the sample (per_sample elements each,
 * a multiple of 4); scale is a device array of B floats (mask / keep_prob).  The branch gradient is
 * the same call: da = dy * scale[b] (a = dy, x = NULL). */
int dat_scale_residual(const void* a, int32_t a_dtype, const void* x, int32_t x_dtype, const float* scale,
                       void* y, int32_t y_dtype, int64_t B, int64_t per_sample, void* stream);

/* ---- "next" rows: channel-last depthwise convolutions ---------- */

/* Depthwise k x k conv, stride 1, padding k/2, x (B,H,W,C) channel-last, w (C,1,k,k) fp32.
 * mode 0: y = conv(x)+b ('X' mixer, dat.py:118-121); mode 1: y = conv(x)+b+x (LPU, dat.py:135-138);
 * mode 2: z = conv(x)+b+x, y = gelu(z), z stored to z_out (MLP middle, dat_blocks.py:338-343).
 * flip = 1 applies the spatially flipped filter (data gradient: dx = dwconv(dz, flip=1, mode 0/1)).
 * workspace >= dat_dwconv_workspace_bytes. */
size_t dat_dwconv_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t C, int32_t k);
int dat_dwconv_fwd(const void* x, int32_t x_dtype, const float* w, const float* bias, void* y, void* z_out,
                   int32_t y_dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, int32_t mode,
                   int32_t flip, void* workspace, size_t workspace_bytes, void* stream);
/* dz = dy * gelu'(z), n elements (n % 4 == 0), all of dtype `dtype`. */
int dat_gelu_bwd(const void* dy, const void* z, void* dz, int32_t dtype, int64_t n, void* stream);
/* dw (C,1,k,k), db (C; may be NULL) fp32, overwritten; k in {3,5,7}; deterministic. */
int dat_dwconv_wgrad(const void* x, int32_t x_dtype, const void* dz, int32_t dz_dtype, float* dw,
                     float* db, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- "next" row: the strided 3 x 3 convolutions around the stages ------------------
 * Conv stem (dat.py:213-218) and down-projections (dat.py:264-274), kernel 3, stride 2, padding 1, run as GEMMs on
 * the tensor-core kernels above:  cols = im2col(x);  Y = dat_pointwise_fwd_tc(cols, W2, b);  dW2 =
 * dat_pointwise_wgrad_tc(dY, cols);  dcols = dat_pointwise_dgrad_tc(dY, W2);  dx = col2im(dcols).
 * cols / dcols: (B * Ho * Wo, Kp) bf16, column t * C + c = tap t = kh * 3 + kw of channel c, Kp =
 * dat_conv3x3s2_kp(C) = 9 C rounded up to a multiple of 64 (zero columns); W2: (Cout, Kp) bf16 in the same column
 * order (dat_conv_weight_pack from the (Cout, C, 3, 3) fp32 parameter; dat_conv_weight_unpack maps dW2 back).
 * x / dx are channel-last (B, H, W, C), C % 8 == 0; nchw_rgb = 1 reads the fp32 NCHW image (C = 3) of the stem. */
int32_t dat_conv3x3s2_kp(int32_t C);
int dat_im2col3x3s2(const void* x, int32_t x_dtype, int32_t nchw_rgb, void* cols, int32_t B, int32_t H, int32_t W,
                    int32_t C, void* stream);
int dat_col2im3x3s2(const void* dcols, void* dx, int32_t dx_dtype, int32_t B, int32_t H, int32_t W, int32_t C,
                    void* stream);
int dat_conv_weight_pack(const float* w, void* w2, int32_t Cout, int32_t C, void* stream);
int dat_conv_weight_unpack(const float* dw2, float* dw, int32_t Cout, int32_t C, void* stream);
/* y = gelu(x) (exact erf form, nn.GELU of the stem, dat.py:215), n % 4 == 0; backward: dat_gelu_bwd. */
int dat_gelu_fwd(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, void* stream);
/* dx = dy * gelu'(x): dy of dy_dtype, x and dx of x_dtype. */
int dat_gelu_bwd_mixed(const void* dy, int32_t dy_dtype, const void* x, void* dx, int32_t x_dtype, int64_t n, void* stream);
/* y[b][c][p] = x[b][p][c] (B, P, C) -> (B, C, P): channel-last -> NCHW-contiguous copy of the backbone outputs
 * (dat.py:308-309); the gradient is the same call with P and C swapped. */
int dat_transpose_pc(const void* x, void* y, int32_t dtype, int32_t B, int32_t P, int32_t C, void* stream);
/* CUDA-core weight / bias gradient of a 1x1 convolution for shapes the tensor-core kernel does not tile (the
 * 32-channel stem convolution): dW[N,K] = dY^T X, db[N] = column sums of dY (may be NULL); deterministic. */
size_t dat_pointwise_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K);
int dat_pointwise_wgrad(const void* dY, int32_t dy_dtype, const void* X, int32_t x_dtype, float* dW, float* db,
                        int64_t M, int32_t N, int32_t K, void* workspace, size_t workspace_bytes, void* stream);

/* Fused backward of the 3 x 3 case (k must be 3, C even): one pass over dy, z (mode 2 only), x forms
 * dz = dy * gelu'(z) on the fly and produces dx (dtype of x), dw (C,1,3,3), db (C; may be NULL);
 * all overwritten, deterministic.  dy / z have dtype d_dtype.  Autograd of dat.py:135-138 and
 * dat_blocks.py:338-343.  workspace >= dat_dwconv_workspace_bytes(B, H, W, C, 3). */
int dat_dwconv_bwd(const void* x, int32_t x_dtype, const void* dy, const void* z, int32_t d_dtype,
                   const float* w, void* dx, float* dw, float* db, int32_t B, int32_t H, int32_t W,
                   int32_t C, int32_t k, int32_t mode, void* workspace, size_t workspace_bytes,
                   void* stream);

/* The rpe bias alone, (B, n_heads, HW, Ns) fp32 (dat_blocks.py:198-212); test hook. */
int dat_rpe_bias(const dat_block_desc* d, const float* pos, const float* rpe_table,
                 float* bias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAT_B200_H_ */
