// Variant branches of the block that no shipped DAT++ config enables:
//   no_off   (dat_blocks.py:57-59,156-157,164-167)  keys / values from avg_pool2d(x, stride)
//   dwc_pe   (:85-87,185-186,221-222)               out += depthwise3x3(q)   [dwconv3.cu kernels + add2 here]
//   fixed_pe (:88-92,187-191)                       dense table, bilinear align_corners resize to (HW, Ns)
//   log_cpb  (:93-99,192-197)                       Linear(2,32)-ReLU-Linear(32,hg) on the log-scaled displacement
// fixed_pe / log_cpb materialise the (.., heads, HW, Ns) fp32 bias like the reference does and feed the
// dense-bias mode of the CUDA-core attention kernels (attention_simt.cu); they are correctness paths,
// not tuned ones.
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int LC_HID = 32;       // hidden width of the log-CPB MLP (dat_blocks.py:95) == warp size
constexpr int LC_MAX_HG = 16;    // heads per group the log-CPB kernels keep in registers

// ---- no_off: average pooling, window = stride = s, floor grid, no padding --------------------------
template <typename TX, typename T>
__global__ void avgpool_fwd_kernel(const TX* __restrict__ x, T* __restrict__ xs, int H, int W, int C, int Hk, int Wk,
                                   int s, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, n, c4)
  if (idx >= total) return;
  const int c4 = (int)(idx % (C / 4));
  const long long bn = idx / (C / 4);
  const int n = (int)(bn % (Hk * Wk));
  const long long b = bn / (Hk * Wk);
  const int i = n / Wk, j = n % Wk;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int u = 0; u < s; ++u)
    for (int v = 0; v < s; ++v) {
      const float4 t = load4(x + ((b * H + i * s + u) * W + j * s + v) * C + 4 * c4);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
  const float inv = 1.0f / (float)(s * s);
  store4(xs + bn * C + 4 * c4, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
}

// dx (B, H, W, C) fp32 += dxs[b, (r / s, c / s)] / s^2 for the pixels inside the pooled region
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dxs, float* __restrict__ dx, int H, int W, int C, int Hk,
                                   int Wk, int s, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, pixel, c4)
  if (idx >= total) return;
  const int c4 = (int)(idx % (C / 4));
  const long long bp = idx / (C / 4);
  const int pix = (int)(bp % (H * W));
  const long long b = bp / (H * W);
  const int i = (pix / W) / s, j = (pix % W) / s;
  if (i >= Hk || j >= Wk) return;
  const float inv = 1.0f / (float)(s * s);
  const float4 g = load4(dxs + ((b * Hk + i) * Wk + j) * C + 4 * c4);
  float* o = dx + bp * C + 4 * c4;
  float4 cur = load4(o);
  store4(o, make_float4(cur.x + g.x * inv, cur.y + g.y * inv, cur.z + g.z * inv, cur.w + g.w * inv));
}

// y = a + b, elementwise (n % 4 == 0); y may alias a
template <typename T>
__global__ void add2_kernel(const T* a, const T* __restrict__ b, T* y, long long n4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 u = load4(a + 4 * i), v = load4(b + 4 * i);
  store4(y + 4 * i, make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w));
}

// ---- fixed_pe: F.interpolate(table[None], (HW, Ns), bilinear, align_corners=True) --------------------
struct Lin {
  int i0, i1;
  float l0, l1;
};
// ATen upsample index / weight rule for align_corners=True: src = dst * (in - 1) / (out - 1)
__device__ __forceinline__ Lin lin_index(int dst, int in, int out) {
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const float r = scale * (float)dst;
  Lin t;
  t.i0 = min((int)floorf(r), in - 1);
  t.l1 = fminf(fmaxf(r - (float)t.i0, 0.f), 1.f);
  t.l0 = 1.0f - t.l1;
  t.i1 = min(t.i0 + 1, in - 1);
  return t;
}

__global__ void fixed_bias_fwd_kernel(const float* __restrict__ table, float* __restrict__ bias, int heads, int Tq,
                                      int Tk, int HW, int Ns, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (eta, m, n)
  if (idx >= total) return;
  const int n = (int)(idx % Ns);
  const int m = (int)((idx / Ns) % HW);
  const int eta = (int)(idx / ((long long)Ns * HW));
  const Lin a = lin_index(m, Tq, HW), b = lin_index(n, Tk, Ns);
  const float* t = table + (long long)eta * Tq * Tk;
  const float top = b.l0 * t[(long long)a.i0 * Tk + b.i0] + b.l1 * t[(long long)a.i0 * Tk + b.i1];
  const float bot = b.l0 * t[(long long)a.i1 * Tk + b.i0] + b.l1 * t[(long long)a.i1 * Tk + b.i1];
  bias[idx] = a.l0 * top + a.l1 * bot;
}

// d table (zeroed by the caller) += resize^T (sum over the batch of dS)
__global__ void fixed_bias_bwd_kernel(const float* __restrict__ dbias, float* __restrict__ dtable, int B, int heads,
                                      int Tq, int Tk, int HW, int Ns, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (eta, m, n)
  if (idx >= total) return;
  const int n = (int)(idx % Ns);
  const int m = (int)((idx / Ns) % HW);
  const int eta = (int)(idx / ((long long)Ns * HW));
  float g = 0.f;
  for (int b = 0; b < B; ++b) g += dbias[(long long)b * total + idx];
  if (g == 0.f) return;
  const Lin a = lin_index(m, Tq, HW), bb = lin_index(n, Tk, Ns);
  float* t = dtable + (long long)eta * Tq * Tk;
  atomicAdd(t + (long long)a.i0 * Tk + bb.i0, g * a.l0 * bb.l0);
  atomicAdd(t + (long long)a.i0 * Tk + bb.i1, g * a.l0 * bb.l1);
  atomicAdd(t + (long long)a.i1 * Tk + bb.i0, g * a.l1 * bb.l0);
  atomicAdd(t + (long long)a.i1 * Tk + bb.i1, g * a.l1 * bb.l1);
}

// ---- log_cpb -------------------------------------------------------------------------------------
// t = sign(d) * log2(|d| + 1) / log2(8), d = (q_grid - pos) * 4   (dat_blocks.py:194-195)
__device__ __forceinline__ float logcpb_coord(float grid, float pos, float* dt_dpos) {
  const float d = __fmul_rn(__fsub_rn(grid, pos), 4.0f);
  const float ad = fabsf(d);
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  if (dt_dpos != nullptr)   // d t / d pos = -4 * sign^2 / ((|d| + 1) ln 2 * 3)
    *dt_dpos = d != 0.f ? -4.0f / ((ad + 1.0f) * 0.6931471805599453f * 3.0f) : 0.f;
  return sgn * log2f(ad + 1.0f) / 3.0f;
}

// bias (B, heads, HW, Ns): one thread per (b, g, m, n), all hg heads of the group
__global__ void logcpb_bias_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ w1,
                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                       float* __restrict__ bias, int H, int W, int G, int hg, int Ns, long long total) {
  pdl_enter();
  __shared__ float s_w1[LC_HID * 2], s_b1[LC_HID], s_w2[LC_MAX_HG * LC_HID];
  for (int i = threadIdx.x; i < LC_HID * 2; i += blockDim.x) s_w1[i] = w1[i];
  for (int i = threadIdx.x; i < LC_HID; i += blockDim.x) s_b1[i] = b1[i];
  for (int i = threadIdx.x; i < hg * LC_HID; i += blockDim.x) s_w2[i] = w2[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, g, m, n)
  if (idx >= total) return;
  const int HW = H * W;
  const int n = (int)(idx % Ns);
  const int m = (int)((idx / Ns) % HW);
  const long long bg = idx / ((long long)Ns * HW);
  const float* pp = pos + (bg * Ns + n) * 2;
  const float ty = logcpb_coord(query_point(m / W, H), pp[0], nullptr);
  const float tx = logcpb_coord(query_point(m % W, W), pp[1], nullptr);
  float out[LC_MAX_HG];
#pragma unroll
  for (int j = 0; j < LC_MAX_HG; ++j) out[j] = 0.f;
  for (int i = 0; i < LC_HID; ++i) {
    const float hid = fmaxf(fmaf(s_w1[2 * i + 1], tx, fmaf(s_w1[2 * i], ty, s_b1[i])), 0.f);
#pragma unroll
    for (int j = 0; j < LC_MAX_HG; ++j)
      if (j < hg) out[j] = fmaf(s_w2[j * LC_HID + i], hid, out[j]);
  }
  const long long b = bg / G;
  const int g = (int)(bg % G);
#pragma unroll
  for (int j = 0; j < LC_MAX_HG; ++j)
    if (j < hg) bias[((b * (G * hg) + g * hg + j) * HW + m) * (long long)Ns + n] = out[j];
}

// Backward of the MLP bias: one warp per (b, g, n), lane = hidden unit, loop over the queries.
// dW1 (32,2), db1 (32), dW2 (hg,32) accumulate in registers, are combined per CTA in shared memory and
// added to the (zeroed) global gradients; dpos[b, g, n] += the bias path's position gradient.
__global__ void __launch_bounds__(256)
logcpb_bias_bwd_kernel(const float* __restrict__ dbias, const float* __restrict__ pos, const float* __restrict__ w1,
                       const float* __restrict__ b1, const float* __restrict__ w2, float* __restrict__ dw1,
                       float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ dpos, int H, int W, int G,
                       int hg, int Ns, long long n_warps) {
  pdl_enter();
  __shared__ float acc[LC_HID * (3 + LC_MAX_HG)];
  for (int i = threadIdx.x; i < LC_HID * (3 + LC_MAX_HG); i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);   // (b, g, n)
  if (wid < n_warps) {
    const int HW = H * W;
    const int n = (int)(wid % Ns);
    const long long bg = wid / Ns;
    const long long b = bg / G;
    const int g = (int)(bg % G);
    const float py = pos[wid * 2], px = pos[wid * 2 + 1];
    const float w1y = w1[2 * lane], w1x = w1[2 * lane + 1], bb = b1[lane];
    float w2r[LC_MAX_HG], gw2[LC_MAX_HG];
#pragma unroll
    for (int j = 0; j < LC_MAX_HG; ++j) {
      w2r[j] = j < hg ? w2[j * LC_HID + lane] : 0.f;
      gw2[j] = 0.f;
    }
    float gw1y = 0.f, gw1x = 0.f, gb1 = 0.f, dpy = 0.f, dpx = 0.f;
    const float* ds_base = dbias + ((b * (G * hg) + g * hg) * HW) * (long long)Ns + n;
    for (int m = 0; m < HW; ++m) {
      float sy, sx;
      const float ty = logcpb_coord(query_point(m / W, H), py, &sy);
      const float tx = logcpb_coord(query_point(m % W, W), px, &sx);
      const float pre = fmaf(w1x, tx, fmaf(w1y, ty, bb));
      const float hid = fmaxf(pre, 0.f);
      float dh = 0.f;
#pragma unroll
      for (int j = 0; j < LC_MAX_HG; ++j)
        if (j < hg) {
          const float ds = ds_base[((long long)j * HW + m) * Ns];
          gw2[j] = fmaf(ds, hid, gw2[j]);
          dh = fmaf(ds, w2r[j], dh);
        }
      dh = pre > 0.f ? dh : 0.f;
      gb1 += dh;
      gw1y = fmaf(dh, ty, gw1y);
      gw1x = fmaf(dh, tx, gw1x);
      dpy = fmaf(warp_sum(dh * w1y), sy, dpy);
      dpx = fmaf(warp_sum(dh * w1x), sx, dpx);
    }
    if (lane == 0) {
      dpos[wid * 2] += dpy;
      dpos[wid * 2 + 1] += dpx;
    }
    atomicAdd(&acc[2 * lane], gw1y);
    atomicAdd(&acc[2 * lane + 1], gw1x);
    atomicAdd(&acc[2 * LC_HID + lane], gb1);
#pragma unroll
    for (int j = 0; j < LC_MAX_HG; ++j)
      if (j < hg) atomicAdd(&acc[3 * LC_HID + j * LC_HID + lane], gw2[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < LC_HID * (3 + hg); i += blockDim.x) {
    const float v = acc[i];
    if (v == 0.f) continue;
    if (i < 2 * LC_HID) atomicAdd(dw1 + i, v);
    else if (i < 3 * LC_HID) atomicAdd(db1 + (i - 2 * LC_HID), v);
    else atomicAdd(dw2 + (i - 3 * LC_HID), v);
  }
}

}  // namespace

int avgpool_fwd(const Shape& s, const void* x, void* xs, cudaStream_t st) {
  DAT_REQUIRE(s.C % 4 == 0, "avgpool: C must be a multiple of 4");
  const long long total = (long long)s.B * s.Ns * (s.C / 4);
  const int grid = ceil_div(total, 256);
  const bool xf = s.x_dtype == DAT_F32, af = s.act_dtype == DAT_F32;
  if (xf && af) launch_k(avgpool_fwd_kernel<float, float>, grid, 256, 0, st, (const float*)x, (float*)xs, s.H, s.W, s.C, s.Hk, s.Wk, s.stride, total);
  else if (xf) launch_k(avgpool_fwd_kernel<float, bf16>, grid, 256, 0, st, (const float*)x, (bf16*)xs, s.H, s.W, s.C, s.Hk, s.Wk, s.stride, total);
  else if (af) launch_k(avgpool_fwd_kernel<bf16, float>, grid, 256, 0, st, (const bf16*)x, (float*)xs, s.H, s.W, s.C, s.Hk, s.Wk, s.stride, total);
  else launch_k(avgpool_fwd_kernel<bf16, bf16>, grid, 256, 0, st, (const bf16*)x, (bf16*)xs, s.H, s.W, s.C, s.Hk, s.Wk, s.stride, total);
  DAT_LAUNCH_OK("avgpool_fwd_kernel");
  return DAT_OK;
}

int avgpool_bwd(const Shape& s, const void* dxs, float* dx, cudaStream_t st) {
  const long long total = (long long)s.B * s.HW * (s.C / 4);
  const int grid = ceil_div(total, 256);
  if (s.act_dtype == DAT_F32) launch_k(avgpool_bwd_kernel<float>, grid, 256, 0, st, (const float*)dxs, dx, s.H, s.W, s.C, s.Hk, s.Wk, s.stride, total);
  else launch_k(avgpool_bwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)dxs, dx, s.H, s.W, s.C, s.Hk, s.Wk, s.stride, total);
  DAT_LAUNCH_OK("avgpool_bwd_kernel");
  return DAT_OK;
}

int add2(const void* a, const void* b, void* y, int dt, long long n, cudaStream_t st) {
  DAT_REQUIRE(n % 4 == 0, "add2: n must be a multiple of 4");
  const long long n4 = n / 4;
  if (dt == DAT_F32) launch_k(add2_kernel<float>, ceil_div(n4, 256), 256, 0, st, (const float*)a, (const float*)b, (float*)y, n4);
  else launch_k(add2_kernel<bf16>, ceil_div(n4, 256), 256, 0, st, (const bf16*)a, (const bf16*)b, (bf16*)y, n4);
  DAT_LAUNCH_OK("add2_kernel");
  return DAT_OK;
}

int fixed_bias_fwd(const Shape& s, const float* table, float* bias, cudaStream_t st) {
  const long long total = (long long)s.heads * s.HW * s.Ns;
  launch_k(fixed_bias_fwd_kernel, ceil_div(total, 256), 256, 0, st, table, bias, s.heads, s.Th, s.Tw, s.HW, s.Ns, total);
  DAT_LAUNCH_OK("fixed_bias_fwd_kernel");
  return DAT_OK;
}

int fixed_bias_bwd(const Shape& s, const float* dbias, float* dtable, cudaStream_t st) {
  const long long total = (long long)s.heads * s.HW * s.Ns;
  DAT_CUDA_OK(cudaMemsetAsync(dtable, 0, (size_t)s.heads * s.Th * s.Tw * 4, st));
  launch_k(fixed_bias_bwd_kernel, ceil_div(total, 256), 256, 0, st, dbias, dtable, s.B, s.heads, s.Th, s.Tw, s.HW, s.Ns, total);
  DAT_LAUNCH_OK("fixed_bias_bwd_kernel");
  return DAT_OK;
}

bool logcpb_supported(const Shape& s) { return s.hg <= LC_MAX_HG; }

int logcpb_bias_fwd(const Shape& s, const float* pos, const float* w1, const float* b1, const float* w2, float* bias,
                    cudaStream_t st) {
  DAT_REQUIRE(logcpb_supported(s), "log_cpb: at most %d heads per group", LC_MAX_HG);
  const long long total = (long long)s.B * s.G * s.HW * s.Ns;
  launch_k(logcpb_bias_fwd_kernel, ceil_div(total, 256), 256, 0, st, pos, w1, b1, w2, bias, s.H, s.W, s.G, s.hg, s.Ns, total);
  DAT_LAUNCH_OK("logcpb_bias_fwd_kernel");
  return DAT_OK;
}

int logcpb_bias_bwd(const Shape& s, const float* dbias, const float* pos, const float* w1, const float* b1,
                    const float* w2, float* dw1, float* db1, float* dw2, float* dpos, cudaStream_t st) {
  DAT_REQUIRE(logcpb_supported(s), "log_cpb: at most %d heads per group", LC_MAX_HG);
  DAT_CUDA_OK(cudaMemsetAsync(dw1, 0, LC_HID * 2 * 4, st));
  DAT_CUDA_OK(cudaMemsetAsync(db1, 0, LC_HID * 4, st));
  DAT_CUDA_OK(cudaMemsetAsync(dw2, 0, (size_t)s.hg * LC_HID * 4, st));
  const long long n_warps = (long long)s.B * s.G * s.Ns;
  launch_k(logcpb_bias_bwd_kernel, ceil_div(n_warps, 8), 256, 0, st, dbias, pos, w1, b1, w2, dw1, db1, dw2, dpos, s.H, s.W, s.G,
                                                                s.hg, s.Ns, n_warps);
  DAT_LAUNCH_OK("logcpb_bias_bwd_kernel");
  return DAT_OK;
}

}  // namespace dat
