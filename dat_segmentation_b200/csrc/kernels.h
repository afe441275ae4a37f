// Internal (non-ABI) launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/dat_b200.h"

namespace dat {

struct Shape;

// gemm_simt.cu — fp32-accumulate CUDA-core GEMMs (exact path)
int pointwise_fwd_simt(const void* X, int x_dt, const float* W, const float* b, void* Y, int y_dt,
                       long long M, int N, int K, cudaStream_t st);
int pointwise_dgrad_simt(const void* dY, int dy_dt, const float* W, void* dX, int dx_dt,
                         long long M, int N, int K, int accumulate, cudaStream_t st);
int wgrad_splits(long long M, int N, int K);
int colsum_splits(long long M);
size_t pointwise_wgrad_workspace(long long M, int N, int K);
int pointwise_wgrad_simt(const void* dY, int dy_dt, const void* X, int x_dt, float* dW, float* db,
                         long long M, int N, int K, void* ws, size_t ws_bytes, cudaStream_t st);
// gather fused with the k / v projections (gather_kv_tc.cu): xs, k, v (B, Ns, C) bf16 from x, pos and the bf16 weights
bool gather_kv_tc_supported(const Shape& s);
int gather_kv_tc(const Shape& s, const void* x, const float* pos, const void* wk, const void* wv, const float* bk,
                 const float* bv, void* xs, void* k, void* v, cudaStream_t st);
int reduce_partials_pair(const float* partA, const float* partB, int nsplit, long long count, void* outA, void* outB,
                         int out_dt, cudaStream_t st);
int reduce_partials(const float* part, int nsplit, long long count, void* out, int out_dt,
                    cudaStream_t st);

// offset_net.cu
int offset_pos_fwd(const Shape& s, const dat_block_params* p, const void* q, float* t_dw,
                   float* off_raw, float* pos, cudaStream_t st);
int ref_points(int Hk, int Wk, float* ry, float* rx, cudaStream_t st);
size_t offset_bwd_workspace(const Shape& s);
int offset_bwd(const Shape& s, const dat_block_params* p, const void* q, const float* t_dw,
               const float* off_raw, const float* dpos, void* dq, const dat_block_grads* g,
               void* ws, size_t ws_bytes, cudaStream_t st, cudaStream_t pst, cudaEvent_t fork_ev);

// gather.cu
int sample_fwd(const Shape& s, const void* x, const float* pos, void* xs, int32_t* taps,
               cudaStream_t st);
int sample_bwd_dpos(const Shape& s, const void* x, const float* pos, const void* dxs,
                    const float* dpos_bias_part, int qsplit, float* dpos, cudaStream_t st);
int sample_bwd_dx(const Shape& s, const float* pos, const void* dxs, float* dx, cudaStream_t st);

// attention_simt.cu
int attention_fwd_simt(const Shape& s, const void* q, const void* k, const void* v,
                       const float* pos, const float* table, void* o, float* lse,
                       cudaStream_t st, const float* bias = nullptr, long long bias_bstride = 0);
int rpe_bias(const Shape& s, const float* pos, const float* table, float* bias, cudaStream_t st);
int attention_bwd_qsplit(const Shape& s);
size_t attention_bwd_workspace(const Shape& s);
int attention_bwd_simt(const Shape& s, const void* q, const void* k, const void* v, const void* o,
                       const void* d_o, const float* lse, const float* pos, const float* table,
                       void* dq, void* dk, void* dv, float* d_table, float* dpos_part, void* ws,
                       size_t ws_bytes, cudaStream_t st, const float* bias = nullptr, long long bias_bstride = 0,
                       float* dbias = nullptr);

}  // namespace dat

namespace dat {
// gemm_tc.cu — tcgen05 / TMA GEMMs (bf16 hot path)
bool pointwise_fwd_tc_supported(int x_dt, long long M, int N, int K);
int cast_weights_bf16(const float* a, const float* b, const float* c, void* out, long long n,
                      cudaStream_t st);
int pointwise_fwd_tc(const void* X, int x_dt, const void* W, const float* b, void* Y, int y_dt,
                     long long M, int N, int K, cudaStream_t st);
}  // namespace dat

namespace dat {
// attention_tc.cu — tcgen05 fused attention forward (bf16, Ns in {64,128,256})
bool attention_fwd_tc_supported(const Shape& s);
size_t attention_fwd_tc_workspace(const Shape& s);
int attention_fwd_tc(const Shape& s, const void* q, const void* k, const void* v, const float* pos,
                     const float* table, void* o, float* lse, void* ws, size_t ws_bytes,
                     cudaStream_t st);
}  // namespace dat

namespace dat {
int debug_gemm_timing(unsigned long long* out8);
int debug_attn_bwd_timing(unsigned long long* out8);
}

namespace dat {
// offset_net_vec.cu — vectorised offset-network forward (Cg in {32,64,128,256})
bool offset_pos_fwd_vec_supported(const Shape& s);
int offset_pos_fwd_vec(const Shape& s, const dat_block_params* p, const void* q, float* t_dw,
                       float* off_raw, float* pos, cudaStream_t st);
}  // namespace dat

namespace dat {
int cast_transpose_weights_bf16(const float* w0, const float* w1, const float* w2, const float* w3,
                                void* out, int C, cudaStream_t st);
int pointwise_fwd_tc_dual(const void* X, const void* W, const void* X2, const void* W2, int x_dt,
                          const float* b, void* Y, int y_dt, long long M, int N, int K,
                          cudaStream_t st, bool w_mn = false, const float* b2 = nullptr, void* Y2 = nullptr,
                          const float* resid = nullptr, const float* rscale = nullptr, long long rows_per_sample = 0);
bool pointwise_fwd_tc_two_outputs_supported(int N);   // Y2 != NULL: tile width a multiple of 64
// dX[M, K] = dY[M, N] W[N, K] (+ dY2 W2): the bf16 weight is read in place as an MN-major operand (no transpose)
bool pointwise_dgrad_tc_supported(long long M, int N, int K);
int pointwise_dgrad_tc(const void* dY, const void* W, const void* dY2, const void* W2, void* dX, int dx_dt,
                       long long M, int N, int K, cudaStream_t st);
int cast_bf16_multi(const dat_cast_item* items_dev, int n_items, cudaStream_t st);
}  // namespace dat

namespace dat {
// attention_bwd_tc.cu — tcgen05 attention backward (bf16, Ns in {128, 256}, table fits smem)
int attention_delta(const Shape& s, const void* d_o, const void* o, float* delta, cudaStream_t st);
int attention_pack_table(const Shape& s, const float* table, void* out, cudaStream_t st);
bool attention_bwd_tc_supported(const Shape& s);
int attention_bwd_tc_chunks(const Shape& s);
int attention_bwd_tc(const Shape& s, const void* q, const void* k, const void* v, const void* d_o,
                     const float* lse, const float* delta, const float* pos, const void* tab_packed,
                     void* dq, float* dk_part, float* dv_part, float* d_table, float* dpos_part,
                     cudaStream_t st, void* ds_out = nullptr, void* dq_scratch = nullptr);
int attention_bwd_tc_sample_chunks(const Shape& s);   // CTAs along the samples (1 up to 256 samples)
size_t attention_bwd_tc_dq_scratch(const Shape& s);   // bytes of dQ slabs attention_bwd_tc needs as dq_scratch (0: none)
}  // namespace dat

namespace dat {
bool attention_bwd_tc_compact_table(const Shape& s, bool tbl = true);
int attention_pack_table_compact(const Shape& s, const float* table, void* out, cudaStream_t st);
int attention_pack_table2(const Shape& s, const float* table, void* out, cudaStream_t st);
int attention_bwd_pack_table(const Shape& s, const float* table, void* out, bool tbl, cudaStream_t st);
}  // namespace dat

namespace dat {
// gemm_tc_wgrad.cu / gemm_simt.cu
int bias_grad(const void* dY, int dy_dt, float* db, long long M, int N, void* ws, size_t ws_bytes,
              cudaStream_t st);
bool pointwise_wgrad_tc_supported(long long M, int N, int K);
size_t pointwise_wgrad_tc_workspace(long long M, int N, int K);
int pointwise_wgrad_tc(const void* dY, const void* X, float* dW, float* db, long long M, int N, int K,
                       void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace dat

namespace dat {
// layernorm.cu
size_t layernorm_bwd_workspace(long long rows, int C);
int residual_layernorm_fwd(const void* a, const float* scale, long long rows_per_sample, const void* x, int x_dt,
                           const float* gamma, const float* beta, void* xout, void* y, int y_dt, float* mean,
                           float* rstd, long long rows, int C, float eps, cudaStream_t st);
int residual_layernorm_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const float* gamma,
                           const float* mean, const float* rstd, void* dx, const void* dres, void* da,
                           const float* scale, long long rows_per_sample, float* dgamma, float* dbeta,
                           long long rows, int C, void* ws, size_t ws_bytes, cudaStream_t st);
int layernorm_fwd(const void* x, int x_dt, const float* gamma, const float* beta, void* y, int y_dt,
                  float* mean, float* rstd, long long rows, int C, float eps, cudaStream_t st);
int layernorm_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const float* gamma,
                  const float* mean, const float* rstd, void* dx, const void* dres, float* dgamma,
                  float* dbeta, long long rows, int C, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace dat

namespace dat {
// dwconv.cu
size_t dwconv_workspace(int B, int H, int W, int C, int k);
int dwconv_fwd(const void* x, int x_dt, const float* w, const float* bias, void* y, void* z_out, int y_dt,
               int B, int H, int W, int C, int k, int mode, int flip, void* ws, size_t ws_bytes,
               cudaStream_t st);
int gelu_bwd(const void* dy, const void* z, void* dz, int dt, long long n, cudaStream_t st);
int dwconv_wgrad(const void* x, int x_dt, const void* dz, int dz_dt, float* dw, float* db, int B, int H,
                 int W, int C, int k, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace dat

namespace dat {
int cast_transpose_bf16(const float* w, void* out, int N, int K, cudaStream_t st);
}

namespace dat {
// dwconv3.cu - register-window depthwise 3x3 (forward + fused backward)
bool dwconv3_supported(int C, int k);
size_t dwconv3_partial_bytes(int B, int H, int W, int C);
int dwconv3_fwd(const void* x, int x_dt, const float* w, const float* bias, void* y, void* z_out, int y_dt,
                int B, int H, int W, int C, int mode, int flip, cudaStream_t st);
int dwconv3_bwd(const void* x, int x_dt, const void* dy, const void* z, int d_dt, const float* w, void* dx,
                float* dw, float* db, int B, int H, int W, int C, int mode, void* ws, size_t ws_bytes,
                cudaStream_t st);
int dwconv_wgrad_reduce(const float* partial, int nsplit, int kk, int C, float* dw, float* db, cudaStream_t st);
}  // namespace dat

namespace dat {
// residual.cu - y = x + a * s[b] (residual + stochastic depth), also the branch gradient
int scale_residual(const void* a, int a_dt, const void* x, int x_dt, const float* s, void* y, int y_dt,
                   long long B, long long per_sample, cudaStream_t st);
}  // namespace dat

namespace dat {
// dwconv7.cu - input-row-stationary depthwise 7x7 (forward / data gradient, weight gradient)
bool dwconv7_supported(int C, int k);
size_t dwconv7_partial_bytes(int B, int H, int W, int C);
int dwconv7_fwd(const void* x, int x_dt, const float* w, const float* bias, void* y, int y_dt, int B, int H,
                int W, int C, int flip, cudaStream_t st);
int dwconv7_wgrad(const void* x, int x_dt, const void* dz, int dz_dt, float* dw, float* db, int B, int H,
                  int W, int C, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace dat

namespace dat {
// variants.cu - the block's variant branches (no_off, dwc_pe, fixed_pe, log_cpb)
int avgpool_fwd(const Shape& s, const void* x, void* xs, cudaStream_t st);
int avgpool_bwd(const Shape& s, const void* dxs, float* dx, cudaStream_t st);
int add2(const void* a, const void* b, void* y, int dt, long long n, cudaStream_t st);
int fixed_bias_fwd(const Shape& s, const float* table, float* bias, cudaStream_t st);
int fixed_bias_bwd(const Shape& s, const float* dbias, float* dtable, cudaStream_t st);
bool logcpb_supported(const Shape& s);
int logcpb_bias_fwd(const Shape& s, const float* pos, const float* w1, const float* b1, const float* w2, float* bias,
                    cudaStream_t st);
int logcpb_bias_bwd(const Shape& s, const float* dbias, const float* pos, const float* w1, const float* b1,
                    const float* w2, float* dw1, float* db1, float* dw2, float* dpos, cudaStream_t st);
}  // namespace dat

namespace dat {
// rpe_table_grad.cu - d rpe_table from the streamed dS as per-sample tensor-core GEMMs (mma.sync)
bool rpe_table_grad_mma_supported(const Shape& s);
size_t rpe_table_grad_mma_workspace(const Shape& s);
int rpe_table_grad_mma(const Shape& s, const void* ds, const float* pos, float* d_table, void* ws, size_t ws_bytes,
                       cudaStream_t st);
}  // namespace dat

namespace dat {
// conv_im2col.cu - data movement of the strided 3 x 3 convolutions (conv stem, down-projections) run as GEMMs
int conv3x3s2_kp(int C);
int im2col3x3s2(const void* x, int x_dt, int nchw_rgb, void* cols, int B, int H, int W, int C, cudaStream_t st);
int col2im3x3s2(const void* dcols, void* dx, int dx_dt, int B, int H, int W, int C, cudaStream_t st);
int conv_weight_pack(const float* w, void* w2, int Cout, int C, cudaStream_t st);
int conv_weight_unpack(const float* dw2, float* dw, int Cout, int C, cudaStream_t st);
int gelu_fwd(const void* x, int x_dt, void* y, int y_dt, long long n, cudaStream_t st);
int transpose_pc(const void* x, void* y, int dt, int B, int P, int C, cudaStream_t st);
int gelu_bwd_mixed(const void* dy, int dy_dt, const void* x, void* dx, int x_dt, long long n, cudaStream_t st);
}  // namespace dat
