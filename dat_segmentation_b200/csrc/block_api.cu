// C ABI (include/dat_b200.h): argument validation, workspace planning and the host-side
// sequencing of the kernels of one deformable-attention block.  No allocation, no host
// synchronisation: everything is enqueued on the caller's stream.
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace dat {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // statistics only (relaxed atomic increments)

bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("DAT_B200_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_shape(const dat_block_desc* d, Shape* s) {
  DAT_REQUIRE(d != nullptr, "descriptor is NULL");
  DAT_REQUIRE(d->B > 0 && d->H > 1 && d->W > 1, "bad B/H/W = %d/%d/%d (H, W must be > 1)", d->B, d->H, d->W);
  DAT_REQUIRE(d->n_heads > 0 && d->n_groups > 0 && d->n_heads % d->n_groups == 0,
              "n_heads=%d must be a positive multiple of n_groups=%d", d->n_heads, d->n_groups);
  DAT_REQUIRE(d->stride > 0 && d->ksize > 0, "bad stride/ksize");
  DAT_REQUIRE(d->pe_mode >= DAT_PE_RPE && d->pe_mode <= DAT_PE_LOGCPB, "bad pe_mode %d", d->pe_mode);
  s->no_off = d->no_off != 0;
  s->pe_mode = s->no_off ? DAT_PE_NONE : d->pe_mode;
  const bool has_table = s->pe_mode == DAT_PE_RPE || s->pe_mode == DAT_PE_FIXED;
  DAT_REQUIRE(!has_table || (d->table_h > 0 && d->table_w > 0), "bad rpe table size");
  DAT_REQUIRE((d->x_dtype == DAT_F32 || d->x_dtype == DAT_BF16) &&
                  (d->act_dtype == DAT_F32 || d->act_dtype == DAT_BF16), "bad dtype code");
  s->B = d->B; s->H = d->H; s->W = d->W; s->HW = d->H * d->W;
  s->heads = d->n_heads; s->G = d->n_groups;
  s->C = d->n_heads * DAT_HEAD_DIM;
  s->Cg = s->C / s->G; s->hg = s->heads / s->G;
  s->stride = d->stride; s->ksize = d->ksize;
  s->pad = d->ksize != d->stride ? d->ksize / 2 : 0;
  DAT_REQUIRE(d->H + 2 * s->pad >= d->ksize && d->W + 2 * s->pad >= d->ksize, "map smaller than the offset kernel");
  s->Hk = (d->H + 2 * s->pad - d->ksize) / d->stride + 1;
  s->Wk = (d->W + 2 * s->pad - d->ksize) / d->stride + 1;
  DAT_REQUIRE(s->Hk > 1 && s->Wk > 1, "sample grid %dx%d: the reference divides by (Hk-1), (Wk-1)", s->Hk, s->Wk);
  if (s->no_off) {   // F.avg_pool2d(x, stride, stride): floor division, no padding (dat_blocks.py:165-167)
    s->Hk = d->H / d->stride;
    s->Wk = d->W / d->stride;
    DAT_REQUIRE(s->Hk > 0 && s->Wk > 0, "no_off: map smaller than the pooling window");
  }
  s->Ns = s->Hk * s->Wk;
  s->Th = has_table ? d->table_h : 1; s->Tw = has_table ? d->table_w : 1;
  s->orf = d->offset_range_factor;
  s->x_dtype = d->x_dtype; s->act_dtype = d->act_dtype;
  DAT_REQUIRE((long long)s->B * s->HW * s->C < (1ll << 40), "tensor too large");
  return DAT_OK;
}

namespace {

// DAT_B200_DISABLE_TC=1 forces the CUDA-core kernels everywhere (A/B measurements, debugging).
bool tc_enabled() {
  static const int off = [] { const char* e = getenv("DAT_B200_DISABLE_TC"); return e && e[0] == '1' ? 1 : 0; }();
  return off == 0;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base((char*)p) {}
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  }
};

struct BwdPlan {
  void *d_o, *dq, *dk, *dv, *dxs;
  float *dpos_part, *dpos;
  void* wT;         // bf16 transposed copies of (wo, wk, wv, wq) for the tensor-core data gradients
  void* x_bf;       // bf16 copy of an fp32 x for the tensor-core weight gradient of proj_q
  void* sub;        // shared scratch of the individual stages (used one at a time)
  void* wg;         // scratch of the tensor-core weight gradients (they run on a side stream)
  // variant branches: dense bias + its gradient (fixed_pe, log_cpb); LePE, o + LePE, dq of the LePE conv and
  // the depthwise kernels' scratch (dwc_pe)
  float *bias, *dbias;
  void* ds_tab;     // bf16 dS (B * heads, HW, Ns) streamed out of the attention backward for the table gradient
  void* tg_part;    // per-CTA partial tables of the table-gradient GEMMs (fixed-order reduction)
  size_t tg_bytes;
  void *lepe, *o2, *dq_lepe, *dw_ws;
  size_t dw_ws_bytes;
  size_t sub_bytes, wg_bytes, total;
};

// scratch of the variant branches in the forward pass, placed after the bf16 weights / packed table
struct FwdVar {
  float* bias;
  void *lepe, *o2, *dw_ws;
  size_t dw_ws_bytes, total;
};
size_t dense_bias_bytes(const Shape& s) {
  if (s.pe_mode == DAT_PE_FIXED) return (size_t)s.heads * s.HW * s.Ns * 4;
  if (s.pe_mode == DAT_PE_LOGCPB) return (size_t)s.B * s.heads * s.HW * s.Ns * 4;
  return 0;
}
FwdVar plan_fwd_var(const Shape& s, void* base) {
  FwdVar f;
  Carver c(base);
  const size_t act = (size_t)s.B * s.HW * s.C * dtype_size(s.act_dtype);
  const bool dwc = s.pe_mode == DAT_PE_DWC;
  f.bias = (float*)c.take(dense_bias_bytes(s));
  f.lepe = c.take(dwc ? act : 0);
  f.o2 = c.take(dwc ? act : 0);
  f.dw_ws_bytes = dwc ? dwconv_workspace(s.B, s.H, s.W, s.C, 3) : 0;
  f.dw_ws = c.take(f.dw_ws_bytes);
  f.total = c.off;
  return f;
}
size_t fwd_fixed_bytes(const Shape& s) {   // bf16 weight copies + packed table of the tensor-core attention
  if (s.act_dtype != DAT_BF16) return 0;
  return align_up((size_t)4 * s.C * s.C * 2, 256) +
         align_up(attention_fwd_tc_supported(s) ? attention_fwd_tc_workspace(s) : 0, 256);
}

// The weight / bias gradients of the four projections are off the critical path (nothing later in
// the backward reads them).  They are enqueued on a library-owned side stream, forked from the
// caller's stream when their inputs are ready and joined before dat_block_backward returns, so they
// run next to the attention backward (which leaves SMs idle) and the data-gradient chain.  Under
// CUDA-graph capture the fork / join become graph edges.  DAT_B200_SERIAL_WGRAD=1 disables this.
struct SideStreams {
  cudaStream_t s;
  cudaEvent_t ev[5];
  pthread_mutex_t mu;   // one backward at a time per device enqueues on the stream / records the events
};
SideStreams* side_streams() {
  static SideStreams ctx[64];
  static bool made[64];
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  if (getenv("DAT_B200_SERIAL_WGRAD") != nullptr) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  pthread_mutex_lock(&mu);
  if (!made[dev]) {
    bool ok = cudaStreamCreateWithFlags(&ctx[dev].s, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 5 && ok; ++i)
      ok = cudaEventCreateWithFlags(&ctx[dev].ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      pthread_mutex_unlock(&mu);
      return nullptr;
    }
    pthread_mutex_init(&ctx[dev].mu, nullptr);
    made[dev] = true;
  }
  pthread_mutex_unlock(&mu);
  return &ctx[dev];
}
// Scope of one backward's use of the side stream: holds the per-device mutex (two host threads running backward
// on one device must not interleave record / wait on the shared events) and joins the side stream into the
// caller's stream on EVERY exit path, so an early error return never leaves the side stream reading the caller's
// buffers (and never leaves an unjoined fork inside a CUDA-graph capture).
struct SideScope {
  SideStreams* ss;
  cudaStream_t st;
  bool forked = false;
  SideScope(SideStreams* s_, cudaStream_t st_) : ss(s_), st(st_) {
    if (ss != nullptr) pthread_mutex_lock(&ss->mu);
  }
  int fork(int i) {
    if (ss != nullptr) {
      DAT_CUDA_OK(cudaEventRecord(ss->ev[i], st));
      DAT_CUDA_OK(cudaStreamWaitEvent(ss->s, ss->ev[i], 0));
      forked = true;
    }
    return DAT_OK;
  }
  ~SideScope() {
    if (ss == nullptr) return;
    if (forked) {
      if (cudaEventRecord(ss->ev[4], ss->s) == cudaSuccess) cudaStreamWaitEvent(st, ss->ev[4], 0);
    }
    pthread_mutex_unlock(&ss->mu);
  }
};

bool use_tc_attn_bwd(const Shape& s) { return tc_enabled() && attention_bwd_tc_supported(s); }
// d rpe_table from the streamed dS as tensor-core GEMMs (rpe_table_grad.cu) instead of the in-kernel scatter;
// DAT_B200_TABLE_SCATTER=1 keeps the scatter (A/B measurements)
bool use_mma_table_grad(const Shape& s) {
  static const int off = [] { const char* e = getenv("DAT_B200_TABLE_SCATTER"); return e && e[0] == '1' ? 1 : 0; }();
  return off == 0 && use_tc_attn_bwd(s) && rpe_table_grad_mma_supported(s);
}
// number of query splits whose partial dK / dV / dpos the backward produces
int bwd_qsplit(const Shape& s) {
  return use_tc_attn_bwd(s) ? attention_bwd_tc_chunks(s) : attention_bwd_qsplit(s);
}
// scratch of the tensor-core attention backward: delta | dk_part | dv_part | packed table | dQ slabs (> 256 samples)
size_t tc_attn_bwd_scratch(const Shape& s) {
  const size_t part = align_up((size_t)attention_bwd_tc_chunks(s) * s.B * s.Ns * s.C * 4, 256);
  return align_up((size_t)s.B * s.heads * s.HW * 4, 256) + 2 * part + align_up(attention_fwd_tc_workspace(s), 256) +
         attention_bwd_tc_dq_scratch(s);
}

BwdPlan plan_bwd(const Shape& s, void* ws) {
  BwdPlan p;
  Carver c(ws);
  const size_t e = dtype_size(s.act_dtype);
  p.d_o = c.take((size_t)s.B * s.HW * s.C * e);
  p.dq = c.take((size_t)s.B * s.HW * s.C * e);
  p.dk = c.take((size_t)s.B * s.Ns * s.C * e);
  p.dv = c.take((size_t)s.B * s.Ns * s.C * e);
  p.dxs = c.take((size_t)s.B * s.Ns * s.C * e);
  p.dpos_part = (float*)c.take((size_t)s.B * s.heads * bwd_qsplit(s) * s.Ns * 2 * 4);
  p.dpos = (float*)c.take((size_t)s.B * s.G * s.Ns * 2 * 4);
  p.wT = c.take((size_t)4 * s.C * s.C * 2);
  p.x_bf = c.take(s.act_dtype == DAT_BF16 && s.x_dtype == DAT_F32 ? (size_t)s.B * s.HW * s.C * 2 : 0);
  size_t sub = attention_bwd_workspace(s);
  if (use_tc_attn_bwd(s) && tc_attn_bwd_scratch(s) > sub) sub = tc_attn_bwd_scratch(s);
  size_t w1 = pointwise_wgrad_workspace((long long)s.B * s.HW, s.C, s.C);
  size_t w2 = pointwise_wgrad_workspace((long long)s.B * s.Ns, s.C, s.C);
  size_t w3 = offset_bwd_workspace(s);
  if (w1 > sub) sub = w1;
  if (w2 > sub) sub = w2;
  if (s.act_dtype == DAT_BF16 && pointwise_wgrad_tc_supported((long long)s.B * s.HW, s.C, s.C)) {
    size_t w4 = pointwise_wgrad_tc_workspace((long long)s.B * s.HW, s.C, s.C) + align_up((size_t)64 * s.C * 4, 256);
    if (w4 > sub) sub = w4;
  }
  if (w3 > sub) sub = w3;
  p.sub_bytes = sub;
  p.sub = c.take(sub);
  p.wg_bytes = 0;
  if (s.act_dtype == DAT_BF16 && pointwise_wgrad_tc_supported((long long)s.B * s.HW, s.C, s.C) &&
      pointwise_wgrad_tc_supported((long long)s.B * s.Ns, s.C, s.C)) {
    const size_t a = pointwise_wgrad_tc_workspace((long long)s.B * s.HW, s.C, s.C);
    const size_t b = pointwise_wgrad_tc_workspace((long long)s.B * s.Ns, s.C, s.C);
    p.wg_bytes = a > b ? a : b;
  }
  p.wg = c.take(p.wg_bytes);
  const bool dwc = s.pe_mode == DAT_PE_DWC;
  const size_t act = (size_t)s.B * s.HW * s.C * e;
  p.bias = (float*)c.take(dense_bias_bytes(s));
  p.dbias = (float*)c.take(dense_bias_bytes(s) ? (size_t)s.B * s.heads * s.HW * s.Ns * 4 : 0);
  p.ds_tab = c.take(use_mma_table_grad(s) ? (size_t)s.B * s.heads * s.HW * s.Ns * 2 : 0);
  p.tg_bytes = use_mma_table_grad(s) ? rpe_table_grad_mma_workspace(s) : 0;
  p.tg_part = c.take(p.tg_bytes);
  p.lepe = c.take(dwc ? act : 0);
  p.o2 = c.take(dwc ? act : 0);
  p.dq_lepe = c.take(dwc ? act : 0);
  p.dw_ws_bytes = dwc ? dwconv_workspace(s.B, s.H, s.W, s.C, 3) : 0;
  p.dw_ws = c.take(p.dw_ws_bytes);
  p.total = c.off;
  return p;
}

}  // namespace
}  // namespace dat

using namespace dat;

extern "C" {

const char* dat_last_error(void) { return g_err; }
const char* dat_version(void) { return "dat_b200 0.1 sm_100a"; }
uint64_t dat_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int dat_sample_grid(const dat_block_desc* d, int32_t* Hk, int32_t* Wk) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  if (Hk) *Hk = s.Hk;
  if (Wk) *Wk = s.Wk;
  return DAT_OK;
}

size_t dat_block_fwd_workspace_bytes(const dat_block_desc* d) {
  Shape s;
  if (make_shape(d, &s) != DAT_OK) return 0;
  // bf16 copies of (wk, wv, wo) and of wq for the tensor-core projections, the packed rpe table of the
  // tensor-core attention kernel, then the scratch of the variant branches
  return fwd_fixed_bytes(s) + plan_fwd_var(s, nullptr).total;
}

int dat_pointwise_fwd_tc_residual(const void* X, int32_t x_dtype, const void* W, const float* b, const float* resid,
                                  const float* scale, int64_t rows_per_sample, float* Y, int64_t M, int32_t N, int32_t K,
                                  void* stream) {
  DAT_REQUIRE(X && W && resid && scale && Y && rows_per_sample > 0, "pointwise_fwd_tc_residual: NULL pointer / bad size");
  if (!pointwise_fwd_tc_supported(x_dtype, M, N, K) || !pointwise_fwd_tc_two_outputs_supported(N)) {
    set_error("pointwise_fwd_tc_residual: shape M=%lld N=%d K=%d not tileable in 64-column groups", (long long)M, N, K);
    return DAT_ERR_UNSUPPORTED;
  }
  return pointwise_fwd_tc_dual(X, W, nullptr, nullptr, x_dtype, b, Y, DAT_F32, M, N, K, (cudaStream_t)stream, false, nullptr,
                               nullptr, resid, scale, rows_per_sample);
}

int dat_pointwise_fwd_tc(const void* X, int32_t x_dtype, const void* W, const float* b, void* Y,
                         int32_t y_dtype, int64_t M, int32_t N, int32_t K, void* stream) {
  DAT_REQUIRE(X && W && Y, "pointwise_fwd_tc: NULL pointer");
  if (!pointwise_fwd_tc_supported(x_dtype, M, N, K)) {
    set_error("pointwise_fwd_tc: shape M=%lld N=%d K=%d not tileable", (long long)M, N, K);
    return DAT_ERR_UNSUPPORTED;
  }
  return pointwise_fwd_tc(X, x_dtype, W, b, Y, y_dtype, M, N, K, (cudaStream_t)stream);
}

int dat_debug_attn_bwd_timing(uint64_t* out8) {
  DAT_REQUIRE(out8 != nullptr, "debug_attn_bwd_timing: NULL pointer");
  return debug_attn_bwd_timing((unsigned long long*)out8);
}

int dat_debug_gemm_timing(uint64_t* out8) {
  DAT_REQUIRE(out8 != nullptr, "debug_gemm_timing: NULL pointer");
  DAT_CUDA_OK(cudaDeviceSynchronize());
  return debug_gemm_timing((unsigned long long*)out8);
}

int dat_cast_transpose_bf16(const float* w, void* out, int32_t N, int32_t K, void* stream) {
  DAT_REQUIRE(w && out && N > 0 && K > 0, "cast_transpose_bf16: bad arguments");
  return cast_transpose_bf16(w, out, N, K, (cudaStream_t)stream);
}

size_t dat_pointwise_wgrad_tc_workspace_bytes(int64_t M, int32_t N, int32_t K) {
  return pointwise_wgrad_tc_supported(M, N, K) ? pointwise_wgrad_tc_workspace(M, N, K) : 0;
}

int dat_pointwise_wgrad_tc(const void* dY, const void* X, float* dW, float* db, int64_t M, int32_t N, int32_t K,
                           void* workspace, size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(dY && X && dW, "pointwise_wgrad_tc: NULL pointer");
  if (!pointwise_wgrad_tc_supported(M, N, K)) {
    set_error("pointwise_wgrad_tc: shape M=%lld N=%d K=%d not tileable", (long long)M, N, K);
    return DAT_ERR_UNSUPPORTED;
  }
  return pointwise_wgrad_tc(dY, X, dW, db, M, N, K, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dat_bias_grad(const void* dY, int32_t dy_dtype, float* db, int64_t M, int32_t N, void* workspace,
                  size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(dY && db && workspace, "bias_grad: NULL pointer");
  return bias_grad(dY, dy_dtype, db, M, N, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dat_cast_bf16_multi(const dat_cast_item* items, int32_t n_items, void* stream) {
  DAT_REQUIRE(items != nullptr && n_items >= 0, "cast_bf16_multi: bad arguments");
  return cast_bf16_multi(items, n_items, (cudaStream_t)stream);
}

int dat_pointwise_dgrad_tc(const void* dY, const void* W, void* dX, int32_t dx_dtype, int64_t M, int32_t N,
                           int32_t K, void* stream) {
  DAT_REQUIRE(dY && W && dX, "pointwise_dgrad_tc: NULL pointer");
  if (!pointwise_dgrad_tc_supported(M, N, K)) {
    set_error("pointwise_dgrad_tc: shape M=%lld N=%d K=%d not tileable with an MN-major weight", (long long)M, N, K);
    return DAT_ERR_UNSUPPORTED;
  }
  return pointwise_dgrad_tc(dY, W, nullptr, nullptr, dX, dx_dtype, M, N, K, (cudaStream_t)stream);
}

int32_t dat_conv3x3s2_kp(int32_t C) { return conv3x3s2_kp(C); }

int dat_im2col3x3s2(const void* x, int32_t x_dtype, int32_t nchw_rgb, void* cols, int32_t B, int32_t H, int32_t W,
                    int32_t C, void* stream) {
  DAT_REQUIRE(x && cols && B > 0 && H > 0 && W > 0 && C > 0, "im2col3x3s2: bad arguments");
  return im2col3x3s2(x, x_dtype, nchw_rgb, cols, B, H, W, C, (cudaStream_t)stream);
}

int dat_col2im3x3s2(const void* dcols, void* dx, int32_t dx_dtype, int32_t B, int32_t H, int32_t W, int32_t C,
                    void* stream) {
  DAT_REQUIRE(dcols && dx && B > 0 && H > 0 && W > 0 && C > 0, "col2im3x3s2: bad arguments");
  return col2im3x3s2(dcols, dx, dx_dtype, B, H, W, C, (cudaStream_t)stream);
}

int dat_conv_weight_pack(const float* w, void* w2, int32_t Cout, int32_t C, void* stream) {
  DAT_REQUIRE(w && w2 && Cout > 0 && C > 0, "conv_weight_pack: bad arguments");
  return conv_weight_pack(w, w2, Cout, C, (cudaStream_t)stream);
}

int dat_conv_weight_unpack(const float* dw2, float* dw, int32_t Cout, int32_t C, void* stream) {
  DAT_REQUIRE(dw2 && dw && Cout > 0 && C > 0, "conv_weight_unpack: bad arguments");
  return conv_weight_unpack(dw2, dw, Cout, C, (cudaStream_t)stream);
}

int dat_gelu_fwd(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, void* stream) {
  DAT_REQUIRE(x && y && n > 0, "gelu_fwd: bad arguments");
  return gelu_fwd(x, x_dtype, y, y_dtype, n, (cudaStream_t)stream);
}

int dat_gelu_bwd_mixed(const void* dy, int32_t dy_dtype, const void* x, void* dx, int32_t x_dtype, int64_t n, void* stream) {
  DAT_REQUIRE(dy && x && dx && n > 0, "gelu_bwd_mixed: bad arguments");
  return gelu_bwd_mixed(dy, dy_dtype, x, dx, x_dtype, n, (cudaStream_t)stream);
}

int dat_transpose_pc(const void* x, void* y, int32_t dtype, int32_t B, int32_t P, int32_t C, void* stream) {
  DAT_REQUIRE(x && y && B > 0 && P > 0 && C > 0, "transpose_pc: bad arguments");
  return transpose_pc(x, y, dtype, B, P, C, (cudaStream_t)stream);
}

size_t dat_pointwise_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K) { return pointwise_wgrad_workspace(M, N, K); }

int dat_pointwise_wgrad(const void* dY, int32_t dy_dtype, const void* X, int32_t x_dtype, float* dW, float* db,
                        int64_t M, int32_t N, int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(dY && X && dW && workspace, "pointwise_wgrad: NULL pointer");
  DAT_REQUIRE(workspace_bytes >= pointwise_wgrad_workspace(M, N, K), "pointwise_wgrad: workspace too small");
  return pointwise_wgrad_simt(dY, dy_dtype, X, x_dtype, dW, db, M, N, K, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dat_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  DAT_REQUIRE(src && dst && n > 0, "cast_bf16: bad arguments");
  return cast_weights_bf16(src, nullptr, nullptr, dst, n, (cudaStream_t)stream);
}

size_t dat_block_bwd_workspace_bytes(const dat_block_desc* d) {
  Shape s;
  if (make_shape(d, &s) != DAT_OK) return 0;
  return plan_bwd(s, nullptr).total;
}

int dat_pointwise_fwd(const void* X, int32_t x_dtype, const float* W, const float* b, void* Y,
                      int32_t y_dtype, int64_t M, int32_t N, int32_t K, void* stream) {
  DAT_REQUIRE(X && W && Y, "pointwise_fwd: NULL pointer");
  return pointwise_fwd_simt(X, x_dtype, W, b, Y, y_dtype, M, N, K, (cudaStream_t)stream);
}

int dat_offset_pos_fwd(const dat_block_desc* d, const dat_block_params* p, const void* q,
                       float* t_dw, float* off_raw, float* pos, void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(p && q && t_dw && off_raw && pos, "offset_pos_fwd: NULL pointer");
  return offset_pos_fwd(s, p, q, t_dw, off_raw, pos, (cudaStream_t)stream);
}

int dat_ref_points(int32_t Hk, int32_t Wk, float* ref_y, float* ref_x, void* stream) {
  DAT_REQUIRE(ref_y && ref_x, "ref_points: NULL pointer");
  return ref_points(Hk, Wk, ref_y, ref_x, (cudaStream_t)stream);
}

int dat_sample_fwd(const dat_block_desc* d, const void* x, const float* pos, void* xs,
                   int32_t* taps, void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(x && pos && xs, "sample_fwd: NULL pointer");
  return sample_fwd(s, x, pos, xs, taps, (cudaStream_t)stream);
}

int dat_gather_kv_fwd(const dat_block_desc* d, const void* x, const float* pos, const void* wk_bf16,
                      const void* wv_bf16, const float* bk, const float* bv, void* xs, void* k, void* v, void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(x && pos && wk_bf16 && wv_bf16 && xs && k && v, "gather_kv_fwd: NULL pointer");
  return gather_kv_tc(s, x, pos, wk_bf16, wv_bf16, bk, bv, xs, k, v, (cudaStream_t)stream);
}

size_t dat_attention_fwd_workspace_bytes(const dat_block_desc* d) {
  Shape s;
  if (make_shape(d, &s) != DAT_OK) return 0;
  return attention_fwd_tc_supported(s) ? attention_fwd_tc_workspace(s) : 0;
}

int dat_attention_fwd(const dat_block_desc* d, const void* q, const void* k, const void* v,
                      const float* pos, const float* rpe_table, void* o, float* lse,
                      void* workspace, size_t workspace_bytes, int32_t impl, void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(q && k && v && pos && rpe_table && o && lse, "attention_fwd: NULL pointer");
  if (impl == 0 && tc_enabled() && attention_fwd_tc_supported(s))
    return attention_fwd_tc(s, q, k, v, pos, rpe_table, o, lse, workspace, workspace_bytes,
                            (cudaStream_t)stream);
  return attention_fwd_simt(s, q, k, v, pos, rpe_table, o, lse, (cudaStream_t)stream);
}

int dat_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, void* y,
                      int32_t y_dtype, float* mean, float* rstd, int64_t rows, int32_t C, float eps,
                      void* stream) {
  DAT_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_fwd: NULL pointer");
  return layernorm_fwd(x, x_dtype, gamma, beta, y, y_dtype, mean, rstd, rows, C, eps, (cudaStream_t)stream);
}

size_t dat_layernorm_bwd_workspace_bytes(int64_t rows, int32_t C) { return layernorm_bwd_workspace(rows, C); }

int dat_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype,
                      const float* gamma, const float* mean, const float* rstd, void* dx,
                      const void* dres, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                      void* workspace, size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "layernorm_bwd: NULL pointer");
  return layernorm_bwd(dy, dy_dtype, x, x_dtype, gamma, mean, rstd, dx, dres, dgamma, dbeta, rows, C, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

int dat_residual_layernorm_fwd(const void* a, const float* scale, int64_t rows_per_sample, const void* x,
                               int32_t x_dtype, const float* gamma, const float* beta, void* xout, void* y,
                               int32_t y_dtype, float* mean, float* rstd, int64_t rows, int32_t C, float eps,
                               void* stream) {
  DAT_REQUIRE(a && scale && x && gamma && beta && xout && y && mean && rstd && rows_per_sample > 0,
              "residual_layernorm_fwd: NULL pointer / bad size");
  return residual_layernorm_fwd(a, scale, rows_per_sample, x, x_dtype, gamma, beta, xout, y, y_dtype, mean, rstd, rows, C,
                                eps, (cudaStream_t)stream);
}

int dat_residual_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* gamma,
                               const float* mean, const float* rstd, void* dx, const void* dres, void* da,
                               const float* scale, int64_t rows_per_sample, float* dgamma, float* dbeta, int64_t rows,
                               int32_t C, void* workspace, size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(dy && x && gamma && mean && rstd && dx && da && scale && dgamma && dbeta && workspace && rows_per_sample > 0,
              "residual_layernorm_bwd: NULL pointer / bad size");
  return residual_layernorm_bwd(dy, dy_dtype, x, x_dtype, gamma, mean, rstd, dx, dres, da, scale, rows_per_sample, dgamma,
                                dbeta, rows, C, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dat_scale_residual(const void* a, int32_t a_dtype, const void* x, int32_t x_dtype, const float* scale,
                       void* y, int32_t y_dtype, int64_t B, int64_t per_sample, void* stream) {
  DAT_REQUIRE(a && scale && y && B >= 0 && per_sample >= 0, "scale_residual: NULL pointer / bad size");
  return scale_residual(a, a_dtype, x, x_dtype, scale, y, y_dtype, B, per_sample, (cudaStream_t)stream);
}

size_t dat_dwconv_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t C, int32_t k) {
  return dwconv_workspace(B, H, W, C, k);
}

int dat_dwconv_fwd(const void* x, int32_t x_dtype, const float* w, const float* bias, void* y, void* z_out,
                   int32_t y_dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, int32_t mode,
                   int32_t flip, void* workspace, size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(x && w && y && workspace, "dwconv_fwd: NULL pointer");
  return dwconv_fwd(x, x_dtype, w, bias, y, z_out, y_dtype, B, H, W, C, k, mode, flip, workspace,
                    workspace_bytes, (cudaStream_t)stream);
}

int dat_gelu_bwd(const void* dy, const void* z, void* dz, int32_t dtype, int64_t n, void* stream) {
  DAT_REQUIRE(dy && z && dz, "gelu_bwd: NULL pointer");
  return gelu_bwd(dy, z, dz, dtype, n, (cudaStream_t)stream);
}

int dat_dwconv_wgrad(const void* x, int32_t x_dtype, const void* dz, int32_t dz_dtype, float* dw,
                     float* db, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k, void* workspace,
                     size_t workspace_bytes, void* stream) {
  DAT_REQUIRE(x && dz && dw && workspace, "dwconv_wgrad: NULL pointer");
  return dwconv_wgrad(x, x_dtype, dz, dz_dtype, dw, db, B, H, W, C, k, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

int dat_dwconv_bwd(const void* x, int32_t x_dtype, const void* dy, const void* z, int32_t d_dtype,
                   const float* w, void* dx, float* dw, float* db, int32_t B, int32_t H, int32_t W,
                   int32_t C, int32_t k, int32_t mode, void* workspace, size_t workspace_bytes,
                   void* stream) {
  DAT_REQUIRE(x && dy && w && dx && dw && workspace, "dwconv_bwd: NULL pointer");
  DAT_REQUIRE(dwconv3_supported(C, k), "dwconv_bwd: the fused backward needs k == 3 and an even C");
  return dwconv3_bwd(x, x_dtype, dy, z, d_dtype, w, dx, dw, db, B, H, W, C, mode, workspace,
                     workspace_bytes, (cudaStream_t)stream);
}

int dat_rpe_bias(const dat_block_desc* d, const float* pos, const float* rpe_table, float* bias,
                 void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(pos && rpe_table && bias, "rpe_bias: NULL pointer");
  return rpe_bias(s, pos, rpe_table, bias, (cudaStream_t)stream);
}

int dat_block_forward(const dat_block_desc* d, const dat_block_params* p, const void* x, void* y,
                      const dat_block_saved* sv, void* workspace, size_t workspace_bytes,
                      void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(p && x && y && sv, "block_forward: NULL pointer");
  DAT_REQUIRE(sv->q && sv->t_dw && sv->off_raw && sv->pos && sv->xs && sv->k && sv->v && sv->o && sv->lse,
              "block_forward: every dat_block_saved buffer must be provided");
  cudaStream_t st = (cudaStream_t)stream;
  const long long M = (long long)s.B * s.HW, Mk = (long long)s.B * s.Ns;
  const int C = s.C, adt = s.act_dtype;
  // bf16 mode: projections on the tensor cores when the shape tiles (else CUDA-core GEMM)
  const bool tc = adt == DAT_BF16 && tc_enabled() && workspace != nullptr &&
                  workspace_bytes >= (size_t)4 * C * C * 2 &&
                  pointwise_fwd_tc_supported(DAT_BF16, Mk, C, C) &&
                  pointwise_fwd_tc_supported(s.x_dtype, M, C, C);
  const size_t wsz = (size_t)C * C;
  bf16* wbf = (bf16*)workspace;   // [wk | wv | wo | wq]
  // bf16 operand copies of the weights: the caller's once-per-step copies when given, else cast here
  const bool pre = p->wk_bf16 && p->wv_bf16 && p->wo_bf16 && (s.x_dtype != DAT_BF16 || p->wq_bf16);
  const bf16* wk_b = pre ? (const bf16*)p->wk_bf16 : wbf;
  const bf16* wv_b = pre ? (const bf16*)p->wv_bf16 : wbf + wsz;
  const bf16* wo_b = pre ? (const bf16*)p->wo_bf16 : wbf + 2 * wsz;
  if (tc) {
    if (!pre) {
      DAT_FWD(cast_weights_bf16(p->wk, p->wv, p->wo, wbf, (long long)wsz, st));
      if (s.x_dtype == DAT_BF16) DAT_FWD(cast_weights_bf16(p->wq, nullptr, nullptr, wbf + 3 * wsz, (long long)wsz, st));
    }
    const void* wq = s.x_dtype == DAT_BF16 ? (pre ? p->wq_bf16 : (const void*)(wbf + 3 * wsz)) : (const void*)p->wq;
    DAT_FWD(pointwise_fwd_tc(x, s.x_dtype, wq, p->bq, sv->q, adt, M, C, C, st));
  } else {
    DAT_FWD(pointwise_fwd_simt(x, s.x_dtype, p->wq, p->bq, sv->q, adt, M, C, C, st));
  }
  bool fused_kv = false;
  if (s.no_off) {   // dat_blocks.py:156-157,164-167: offsets zeroed, keys / values from the average-pooled map
    DAT_CUDA_OK(cudaMemsetAsync(sv->pos, 0, (size_t)s.B * s.G * s.Ns * 2 * 4, st));
    DAT_FWD(avgpool_fwd(s, x, sv->xs, st));
  } else {
    DAT_FWD(offset_pos_fwd(s, p, sv->q, sv->t_dw, sv->off_raw, sv->pos, st));
    // the gather is the A-operand producer of the k / v projections when the shape allows (one launch, no round trip
    // of the sampled features between them); xs is still written: the backward reads it
    fused_kv = tc && gather_kv_tc_supported(s);
    if (fused_kv) DAT_FWD(gather_kv_tc(s, x, sv->pos, wk_b, wv_b, p->bk, p->bv, sv->xs, sv->k, sv->v, st));
    else DAT_FWD(sample_fwd(s, x, sv->pos, sv->xs, nullptr, st));
  }
  if (fused_kv) {
  } else if (tc && pointwise_fwd_tc_two_outputs_supported(C)) {
    // k and v = two products of the same sampled features: one launch of the persistent GEMM
    DAT_FWD(pointwise_fwd_tc_dual(sv->xs, wk_b, nullptr, wv_b, adt, p->bk, sv->k, adt, Mk, C, C, st, false, p->bv, sv->v));
  } else if (tc) {
    DAT_FWD(pointwise_fwd_tc(sv->xs, adt, wk_b, p->bk, sv->k, adt, Mk, C, C, st));
    DAT_FWD(pointwise_fwd_tc(sv->xs, adt, wv_b, p->bv, sv->v, adt, Mk, C, C, st));
  } else {
    DAT_FWD(pointwise_fwd_simt(sv->xs, adt, p->wk, p->bk, sv->k, adt, Mk, C, C, st));
    DAT_FWD(pointwise_fwd_simt(sv->xs, adt, p->wv, p->bv, sv->v, adt, Mk, C, C, st));
  }
  const size_t w_bytes = align_up((size_t)4 * C * C * 2, 256);
  // variant branches: scratch behind the fixed part of the workspace
  FwdVar fv = plan_fwd_var(s, nullptr);
  if (fv.total > 0) {
    DAT_REQUIRE(workspace != nullptr && workspace_bytes >= fwd_fixed_bytes(s) + fv.total,
                "block_forward: workspace %zu < %zu bytes", workspace_bytes, fwd_fixed_bytes(s) + fv.total);
    fv = plan_fwd_var(s, (char*)workspace + fwd_fixed_bytes(s));
  }
  long long bias_bstride = 0;
  if (s.pe_mode == DAT_PE_FIXED) {
    DAT_REQUIRE(p->rpe_table, "block_forward: fixed_pe needs rpe_table");
    DAT_FWD(fixed_bias_fwd(s, p->rpe_table, fv.bias, st));
  } else if (s.pe_mode == DAT_PE_LOGCPB) {
    DAT_REQUIRE(p->rpe_table && p->pe_b && p->pe_w2, "block_forward: log_cpb needs rpe_table.0.weight / .0.bias / .2.weight");
    DAT_FWD(logcpb_bias_fwd(s, sv->pos, p->rpe_table, p->pe_b, p->pe_w2, fv.bias, st));
    bias_bstride = (long long)s.heads * s.HW * s.Ns;
  }
  if (adt == DAT_BF16 && tc_enabled() && attention_fwd_tc_supported(s) && workspace != nullptr &&
      workspace_bytes >= w_bytes + attention_fwd_tc_workspace(s)) {
    DAT_REQUIRE(p->rpe_table, "block_forward: rpe_table is NULL");
    DAT_FWD(attention_fwd_tc(s, sv->q, sv->k, sv->v, sv->pos, p->rpe_table, sv->o, sv->lse,
                             (char*)workspace + w_bytes, attention_fwd_tc_workspace(s), st));
  } else {
    DAT_REQUIRE(s.pe_mode != DAT_PE_RPE || p->rpe_table, "block_forward: rpe_table is NULL");
    DAT_FWD(attention_fwd_simt(s, sv->q, sv->k, sv->v, sv->pos, p->rpe_table, sv->o, sv->lse, st, fv.bias, bias_bstride));
  }
  const void* o_in = sv->o;
  if (s.pe_mode == DAT_PE_DWC) {   // out = out + rpe_table(q), a depthwise 3x3 conv (dat_blocks.py:185-186,221-222)
    DAT_REQUIRE(p->rpe_table && p->pe_b, "block_forward: dwc_pe needs rpe_table.weight / rpe_table.bias");
    DAT_FWD(dwconv_fwd(sv->q, adt, p->rpe_table, p->pe_b, fv.lepe, nullptr, adt, s.B, s.H, s.W, C, 3, 0, 0, fv.dw_ws,
                       fv.dw_ws_bytes, st));
    DAT_FWD(add2(sv->o, fv.lepe, fv.o2, adt, M * C, st));
    o_in = fv.o2;
  }
  if (tc) DAT_FWD(pointwise_fwd_tc(o_in, adt, wo_b, p->bo, y, adt, M, C, C, st));
  else DAT_FWD(pointwise_fwd_simt(o_in, adt, p->wo, p->bo, y, adt, M, C, C, st));
  return DAT_OK;
}

int dat_block_backward(const dat_block_desc* d, const dat_block_params* p, const void* x,
                       const void* dy, const dat_block_saved* sv, float* dx,
                       const dat_block_grads* g, void* workspace, size_t workspace_bytes,
                       void* stream) {
  Shape s;
  DAT_FWD(make_shape(d, &s));
  DAT_REQUIRE(p && x && dy && sv && dx && g && workspace, "block_backward: NULL pointer");
  BwdPlan w = plan_bwd(s, workspace);
  DAT_REQUIRE(workspace_bytes >= w.total, "block_backward: workspace %zu < %zu bytes", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  const long long M = (long long)s.B * s.HW, Mk = (long long)s.B * s.Ns;
  const int C = s.C, adt = s.act_dtype;
  // bf16 mode: data gradients dX = dY W run on the tensor cores as K-major GEMMs against
  // transposed bf16 copies of the weights (one tiny transpose launch per backward)
  const bool tc = adt == DAT_BF16 && tc_enabled() && pointwise_fwd_tc_supported(DAT_BF16, M, C, C) &&
                  pointwise_fwd_tc_supported(DAT_BF16, Mk, C, C);
  bf16* wT = (bf16*)w.wT;          // [wo | wk | wv | wq]: transposed copies, or plain copies when read MN-major
  const size_t wsz = (size_t)C * C;
  // data gradients dX = dY W: the bf16 weight copy of the forward is read in place as an MN-major operand when C
  // tiles by 64; other widths (C = 96 ...) use K-major products against transposed copies
  const bool mn = tc && pointwise_dgrad_tc_supported(M, C, C) && pointwise_dgrad_tc_supported(Mk, C, C) &&
                  getenv("DAT_B200_DGRAD_TRANSPOSE") == nullptr;
  const bool pre = mn && p->wq_bf16 && p->wk_bf16 && p->wv_bf16 && p->wo_bf16;
  const bf16* wo_b = pre ? (const bf16*)p->wo_bf16 : wT;
  const bf16* wk_b = pre ? (const bf16*)p->wk_bf16 : wT + wsz;
  const bf16* wv_b = pre ? (const bf16*)p->wv_bf16 : wT + 2 * wsz;
  const bf16* wq_b = pre ? (const bf16*)p->wq_bf16 : wT + 3 * wsz;
  if (tc && mn && !pre) {
    DAT_FWD(cast_weights_bf16(p->wo, p->wk, p->wv, wT, (long long)wsz, st));
    DAT_FWD(cast_weights_bf16(p->wq, nullptr, nullptr, wT + 3 * wsz, (long long)wsz, st));
  } else if (tc && !mn) {
    DAT_FWD(cast_transpose_weights_bf16(p->wo, p->wk, p->wv, p->wq, wT, C, st));
  }
  auto dgrad = [&](const void* dY, const bf16* W, const void* dY2, const bf16* W2, void* dX, int dx_dt, long long rows) -> int {
    if (mn) return pointwise_dgrad_tc(dY, W, dY2, W2, dX, dx_dt, rows, C, C, st);
    return pointwise_fwd_tc_dual(dY, W, dY2, W2, DAT_BF16, nullptr, dX, dx_dt, rows, C, C, st);
  };
  // weight gradients dW = dY^T X on the tensor cores (both operands read MN-major), bias
  // gradients as column sums
  const bool tcw = tc && pointwise_wgrad_tc_supported(M, C, C) && pointwise_wgrad_tc_supported(Mk, C, C);
  SideStreams* ss = tcw && w.wg_bytes > 0 ? side_streams() : nullptr;
  cudaStream_t wst = ss != nullptr ? ss->s : st;       // stream of the tensor-core weight gradients
  SideScope scope(ss, st);
  auto fork = [&](int i) -> int { return scope.fork(i); };   // side stream waits for the work enqueued so far
  auto wgrad = [&](const void* dY, const void* X, int x_dt, float* dW, float* db, long long rows) -> int {
    if (!tcw) return pointwise_wgrad_simt(dY, adt, X, x_dt, dW, db, rows, C, C, w.sub, w.sub_bytes, st);
    return pointwise_wgrad_tc(dY, X, dW, db, rows, C, C, w.wg, w.wg_bytes, wst);
  };
  // dwc_pe: the input of proj_out was o + LePE(q); recompute it (two launches) instead of saving it
  const void* o_in = sv->o;
  if (s.pe_mode == DAT_PE_DWC) {
    DAT_REQUIRE(p->rpe_table && p->pe_b && g->rpe_table && g->pe_b, "block_backward: dwc_pe needs rpe_table.weight / .bias");
    DAT_FWD(dwconv_fwd(sv->q, adt, p->rpe_table, p->pe_b, w.lepe, nullptr, adt, s.B, s.H, s.W, C, 3, 0, 0, w.dw_ws,
                       w.dw_ws_bytes, st));
    DAT_FWD(add2(sv->o, w.lepe, w.o2, adt, M * C, st));
    o_in = w.o2;
  }
  // proj_out
  DAT_FWD(fork(0));
  DAT_FWD(wgrad(dy, o_in, adt, g->wo, g->bo, M));
  if (tc) DAT_FWD(dgrad(dy, wo_b, nullptr, nullptr, w.d_o, adt, M));
  else DAT_FWD(pointwise_dgrad_simt(dy, adt, p->wo, w.d_o, adt, M, C, C, 0, st));
  // attention core
  if (use_tc_attn_bwd(s)) {
    const int chunks = attention_bwd_tc_chunks(s);
    const size_t part = align_up((size_t)chunks * s.B * s.Ns * C * 4, 256);
    float* delta = (float*)w.sub;
    float* dk_part = (float*)((char*)w.sub + align_up((size_t)s.B * s.heads * s.HW * 4, 256));
    float* dv_part = (float*)((char*)dk_part + part);
    void* tabp = (char*)dv_part + part;
    DAT_FWD(attention_delta(s, w.d_o, sv->o, delta, st));
    const bool mma_table = use_mma_table_grad(s);
    if (!mma_table) DAT_CUDA_OK(cudaMemsetAsync(g->rpe_table, 0, (size_t)s.heads * s.Th * s.Tw * 4, st));
    DAT_FWD(attention_bwd_pack_table(s, p->rpe_table, tabp, !mma_table, st));
    void* dq_slabs = attention_bwd_tc_dq_scratch(s) ? (char*)tabp + align_up(attention_fwd_tc_workspace(s), 256) : nullptr;
    DAT_FWD(attention_bwd_tc(s, sv->q, sv->k, sv->v, w.d_o, sv->lse, delta, sv->pos, tabp, w.dq, dk_part,
                             dv_part, g->rpe_table, w.dpos_part, st, mma_table ? w.ds_tab : nullptr, dq_slabs));
    DAT_FWD(reduce_partials_pair(dk_part, dv_part, chunks, (long long)s.B * s.Ns * C, w.dk, w.dv, adt, st));
  } else {
    long long bias_bstride = 0;
    if (s.pe_mode == DAT_PE_FIXED) {
      DAT_REQUIRE(p->rpe_table && g->rpe_table, "block_backward: fixed_pe needs rpe_table");
      DAT_FWD(fixed_bias_fwd(s, p->rpe_table, w.bias, st));
    } else if (s.pe_mode == DAT_PE_LOGCPB) {
      DAT_REQUIRE(p->rpe_table && p->pe_b && p->pe_w2 && g->rpe_table && g->pe_b && g->pe_w2,
                  "block_backward: log_cpb needs rpe_table.0.weight / .0.bias / .2.weight and their gradients");
      DAT_FWD(logcpb_bias_fwd(s, sv->pos, p->rpe_table, p->pe_b, p->pe_w2, w.bias, st));
      bias_bstride = (long long)s.heads * s.HW * s.Ns;
    }
    if (s.pe_mode == DAT_PE_DWC)   // gradient of the LePE conv: dq_lepe, d weight, d bias from dO (= d(o + LePE))
      DAT_FWD(dwconv3_bwd(sv->q, adt, w.d_o, nullptr, adt, p->rpe_table, w.dq_lepe, g->rpe_table, g->pe_b, s.B, s.H,
                          s.W, C, 0, w.dw_ws, w.dw_ws_bytes, st));
    DAT_REQUIRE(s.pe_mode != DAT_PE_RPE || (p->rpe_table && g->rpe_table), "block_backward: rpe_table is NULL");
    DAT_FWD(attention_bwd_simt(s, sv->q, sv->k, sv->v, sv->o, w.d_o, sv->lse, sv->pos, p->rpe_table,
                               w.dq, w.dk, w.dv, g->rpe_table, w.dpos_part, w.sub, w.sub_bytes, st, w.bias,
                               bias_bstride, w.dbias));
    if (s.pe_mode == DAT_PE_FIXED) DAT_FWD(fixed_bias_bwd(s, w.dbias, g->rpe_table, st));
    if (s.pe_mode == DAT_PE_DWC) DAT_FWD(add2(w.dq, w.dq_lepe, w.dq, adt, M * C, st));
  }
  // proj_k / proj_v
  DAT_FWD(fork(1));
  if (use_mma_table_grad(s))   // d rpe_table: nothing later reads it - on the side stream when there is one
    DAT_FWD(rpe_table_grad_mma(s, w.ds_tab, sv->pos, g->rpe_table, w.tg_part, w.tg_bytes, wst));
  DAT_FWD(wgrad(w.dk, sv->xs, adt, g->wk, g->bk, Mk));
  DAT_FWD(wgrad(w.dv, sv->xs, adt, g->wv, g->bv, Mk));
  if (tc) {
    DAT_FWD(dgrad(w.dk, wk_b, w.dv, wv_b, w.dxs, adt, Mk));
  } else {
    DAT_FWD(pointwise_dgrad_simt(w.dk, adt, p->wk, w.dxs, adt, Mk, C, C, 0, st));
    DAT_FWD(pointwise_dgrad_simt(w.dv, adt, p->wv, w.dxs, adt, Mk, C, C, 1, st));
  }
  // sampling -> d pos; offset network -> dq  (no_off: the offset network is unused, its gradients are not written)
  if (!s.no_off) {
    DAT_FWD(sample_bwd_dpos(s, x, sv->pos, w.dxs, w.dpos_part, bwd_qsplit(s), w.dpos, st));
    if (s.pe_mode == DAT_PE_LOGCPB)   // the bias MLP's parameter gradients and its part of d pos
      DAT_FWD(logcpb_bias_bwd(s, w.dbias, sv->pos, p->rpe_table, p->pe_b, p->pe_w2, g->rpe_table, g->pe_b, g->pe_w2,
                              w.dpos, st));
    DAT_FWD(offset_bwd(s, p, sv->q, sv->t_dw, sv->off_raw, w.dpos, w.dq, g, w.sub, w.sub_bytes, st,
                       ss != nullptr ? ss->s : st, ss != nullptr ? ss->ev[3] : nullptr));
  }
  // proj_q, then the sampling scatter on top of its data gradient
  DAT_FWD(fork(2));
  if (tcw && s.x_dtype == DAT_F32) {
    DAT_FWD(cast_weights_bf16((const float*)x, nullptr, nullptr, w.x_bf, M * C, wst));
    DAT_FWD(wgrad(w.dq, w.x_bf, DAT_BF16, g->wq, g->bq, M));
  } else {
    DAT_FWD(wgrad(w.dq, x, s.x_dtype, g->wq, g->bq, M));
  }
  if (tc) DAT_FWD(dgrad(w.dq, wq_b, nullptr, nullptr, dx, DAT_F32, M));
  else DAT_FWD(pointwise_dgrad_simt(w.dq, adt, p->wq, dx, DAT_F32, M, C, C, 0, st));
  if (s.no_off) DAT_FWD(avgpool_bwd(s, w.dxs, dx, st));
  else DAT_FWD(sample_bwd_dx(s, sv->pos, w.dxs, dx, st));
  return DAT_OK;   // ~SideScope joins: the caller's stream owns every result again
}

}  // extern "C"
