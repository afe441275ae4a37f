// The strided 3 x 3 convolutions of the DAT backbone around the block: the conv stem
// (dat.py:213-218: Conv 3->C/2 k3 s2 p1, LayerNorm, GELU, Conv C/2->C k3 s2 p1, LayerNorm) and the three
// down-projections (dat.py:264-274: Conv C_i -> C_{i+1} k3 s2 p1, no bias) as GEMMs on the tcgen05 kernels of
// gemm_tc.cu / gemm_tc_wgrad.cu:
//     cols[m, t * C + c] = x[b, 2 ho - 1 + kh, 2 wo - 1 + kw, c]   (m = (b, ho, wo), t = kh * 3 + kw, zero outside)
//     Y = cols W2^T + b,   dW2 = dY^T cols,   dcols = dY W2,   dx = col2im(dcols)
// with W2[co, t * C + c] = w[co, c, kh, kw].  This file holds the data-movement kernels around those GEMMs; all are
// channel-last, 16 bytes per thread access, HBM-bound.  K = 9 C is padded with zero columns to a multiple of 64 (the
// weight-gradient kernel's tile), which only matters for C = 3 (27 -> 64) and C = 32 (288 -> 320).
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 load8_bf16(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ uint4 load8_bf16(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  return make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
}

// thread = (output pixel m, column group of 8): cols[m][8 g .. 8 g + 8)
template <typename TI>
__global__ void im2col3x3s2_kernel(const TI* __restrict__ x, bf16* __restrict__ cols, int B, int H, int W, int C,
                                   int Ho, int Wo, int Kp, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int groups = Kp >> 3;
  const long long m = idx / groups;
  const int col = (int)(idx - m * groups) << 3;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (col < 9 * C) {
    const int t = col / C, c = col - t * C;                 // C % 8 == 0: a group never straddles two taps
    const int wo = (int)(m % Wo);
    const long long r = m / Wo;
    const int ho = (int)(r % Ho), b = (int)(r / Ho);
    const int hi = 2 * ho - 1 + t / 3, wi = 2 * wo - 1 + t % 3;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = load8_bf16(x + (((long long)b * H + hi) * W + wi) * C + c);
  }
  *reinterpret_cast<uint4*>(cols + m * Kp + col) = v;
}

// the RGB stem: x is the NCHW fp32 image (B, 3, H, W); thread = output pixel, writes its 27 taps (t * 3 + c) and the
// zero padding up to Kp = 64 as eight 16-byte pieces
__global__ void im2col3x3s2_rgb_kernel(const float* __restrict__ x, bf16* __restrict__ cols, int B, int H, int W,
                                       int Ho, int Wo, int Kp, long long M) {
  pdl_enter();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int wo = (int)(m % Wo);
  const long long r = m / Wo;
  const int ho = (int)(r % Ho), b = (int)(r / Ho);
  float v[28];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int hi = 2 * ho - 1 + t / 3, wi = 2 * wo - 1 + t % 3;
    const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[t * 3 + c] = ok ? x[(((long long)b * 3 + c) * H + hi) * W + wi] : 0.f;
  }
  v[27] = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(cols + m * Kp);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* p = v + 8 * g;
    dst[g] = g < 3 ? make_uint4(pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]), pack2(p[6], p[7]))
                   : make_uint4(pack2(p[0], p[1]), pack2(p[2], p[3]), 0u, 0u);
  }
  for (int g = 4; g < (Kp >> 3); ++g) dst[g] = make_uint4(0u, 0u, 0u, 0u);
}

// dx[b, h, w, c] = sum over the <= 4 (kh, kw) with (h + 1 - kh, w + 1 - kw) even and in range of
// dcols[(b, (h + 1 - kh) / 2, (w + 1 - kw) / 2), (kh * 3 + kw) * C + c]: gather form, no atomics.
// thread = (input pixel, 8 channels)
template <typename TO>
__global__ void col2im3x3s2_kernel(const bf16* __restrict__ dcols, TO* __restrict__ dx, int B, int H, int W, int C,
                                   int Ho, int Wo, int Kp, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int groups = C >> 3;
  const long long pix = idx / groups;
  const int c = (int)(idx - pix * groups) << 3;
  const int w = (int)(pix % W);
  const long long r = pix / W;
  const int h = (int)(r % H), b = (int)(r / H);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hh = h + 1 - kh;
    if (hh < 0 || (hh & 1) || (hh >> 1) >= Ho) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ww = w + 1 - kw;
      if (ww < 0 || (ww & 1) || (ww >> 1) >= Wo) continue;
      const long long m = ((long long)b * Ho + (hh >> 1)) * Wo + (ww >> 1);
      const uint4 raw = *reinterpret_cast<const uint4*>(dcols + m * Kp + (kh * 3 + kw) * C + c);
      const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(rw[i] << 16);
        acc[2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
      }
    }
  }
  TO* dst = dx + pix * C + c;
  store4(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
  store4(dst + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
}

// w (Cout, C, 3, 3) fp32 -> W2 (Cout, Kp) bf16, W2[co][t * C + c] = w[co][c][t], zero padding columns
__global__ void conv_weight_pack_kernel(const float* __restrict__ w, bf16* __restrict__ w2, int Cout, int C, int Kp) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Kp) return;
  const int co = idx / Kp, col = idx - co * Kp;
  float v = 0.f;
  if (col < 9 * C) {
    const int t = col / C, c = col - t * C;
    v = w[((long long)co * C + c) * 9 + t];
  }
  w2[idx] = __float2bfloat16_rn(v);
}
// dW2 (Cout, Kp) fp32 -> dw (Cout, C, 3, 3) fp32
__global__ void conv_weight_unpack_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int Cout, int C, int Kp) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * C * 9) return;
  const int t = idx % 9, c = (idx / 9) % C, co = idx / (9 * C);
  dw[idx] = dw2[(long long)co * Kp + t * C + c];
}

__device__ __forceinline__ float gelu_exact(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f)); }
template <typename TI, typename TO>
__global__ void gelu_fwd_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = load4(x + 4 * i);
  store4(y + 4 * i, make_float4(gelu_exact(v.x), gelu_exact(v.y), gelu_exact(v.z), gelu_exact(v.w)));
}

__device__ __forceinline__ float gelu_dexact(float z) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
  return cdf + z * expf(-0.5f * z * z) * 0.39894228040143267794f;
}
template <typename TD, typename TX, typename TO>
__global__ void gelu_bwd_mixed_kernel(const TD* __restrict__ dy, const TX* __restrict__ x, TO* __restrict__ dx, long long n4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 d = load4(dy + 4 * i), z = load4(x + 4 * i);
  store4(dx + 4 * i, make_float4(d.x * gelu_dexact(z.x), d.y * gelu_dexact(z.y), d.z * gelu_dexact(z.z), d.w * gelu_dexact(z.w)));
}

// y[b][c][p] = x[b][p][c]: 32 x 32 tiles through shared memory, both sides coalesced.  The four backbone outputs are
// returned NCHW-contiguous like the reference's (dat.py:308-309 `.contiguous()`); the same kernel with (P, C) swapped
// is the gradient.
template <typename T>
__global__ void transpose_pc_kernel(const T* __restrict__ x, T* __restrict__ y, int P, int C) {
  pdl_enter();
  __shared__ T tile[32][33];
  const long long base = (long long)blockIdx.z * P * C;
  const int p0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < P && c < C) tile[r][threadIdx.x] = x[base + (long long)p * C + c];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < P) y[base + (long long)c * P + p] = tile[threadIdx.x][r];
  }
}

}  // namespace

int transpose_pc(const void* x, void* y, int dt, int B, int P, int C, cudaStream_t st) {
  dim3 grid(ceil_div(C, 32), ceil_div(P, 32), B), block(32, 8);
  DAT_REQUIRE(B <= 65535 && grid.y <= 65535, "transpose: tensor too large");
  if (dt == DAT_F32) launch_k(transpose_pc_kernel<float>, grid, block, 0, st, (const float*)x, (float*)y, P, C);
  else launch_k(transpose_pc_kernel<bf16>, grid, block, 0, st, (const bf16*)x, (bf16*)y, P, C);
  DAT_LAUNCH_OK("transpose_pc_kernel");
  return DAT_OK;
}

// dx = dy * gelu'(x); dy bf16 or fp32, x and dx of one dtype (the stem: dy bf16, x / dx fp32)
int gelu_bwd_mixed(const void* dy, int dy_dt, const void* x, void* dx, int x_dt, long long n, cudaStream_t st) {
  DAT_REQUIRE(n % 4 == 0, "gelu_bwd: n must be a multiple of 4");
  const long long n4 = n / 4;
  const int grid = ceil_div(n4, 256);
  if (dy_dt == DAT_BF16 && x_dt == DAT_F32) launch_k(gelu_bwd_mixed_kernel<bf16, float, float>, grid, 256, 0, st, (const bf16*)dy, (const float*)x, (float*)dx, n4);
  else if (dy_dt == DAT_BF16) launch_k(gelu_bwd_mixed_kernel<bf16, bf16, bf16>, grid, 256, 0, st, (const bf16*)dy, (const bf16*)x, (bf16*)dx, n4);
  else if (x_dt == DAT_F32) launch_k(gelu_bwd_mixed_kernel<float, float, float>, grid, 256, 0, st, (const float*)dy, (const float*)x, (float*)dx, n4);
  else launch_k(gelu_bwd_mixed_kernel<float, bf16, bf16>, grid, 256, 0, st, (const float*)dy, (const bf16*)x, (bf16*)dx, n4);
  DAT_LAUNCH_OK("gelu_bwd_mixed_kernel");
  return DAT_OK;
}

int conv3x3s2_kp(int C) { return (9 * C + 63) / 64 * 64; }

int im2col3x3s2(const void* x, int x_dt, int nchw_rgb, void* cols, int B, int H, int W, int C, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, Kp = conv3x3s2_kp(C);
  const long long M = (long long)B * Ho * Wo;
  if (nchw_rgb) {
    DAT_REQUIRE(C == 3 && x_dt == DAT_F32, "im2col: the NCHW path is the fp32 RGB stem only");
    launch_k(im2col3x3s2_rgb_kernel, ceil_div(M, 256), 256, 0, st, (const float*)x, (bf16*)cols, B, H, W, Ho, Wo, Kp, M);
  } else {
    DAT_REQUIRE(C % 8 == 0, "im2col: C must be a multiple of 8");
    const long long total = M * (Kp >> 3);
    if (x_dt == DAT_F32)
      launch_k(im2col3x3s2_kernel<float>, ceil_div(total, 256), 256, 0, st, (const float*)x, (bf16*)cols, B, H, W, C, Ho, Wo, Kp, total);
    else
      launch_k(im2col3x3s2_kernel<bf16>, ceil_div(total, 256), 256, 0, st, (const bf16*)x, (bf16*)cols, B, H, W, C, Ho, Wo, Kp, total);
  }
  DAT_LAUNCH_OK("im2col3x3s2_kernel");
  return DAT_OK;
}

int col2im3x3s2(const void* dcols, void* dx, int dx_dt, int B, int H, int W, int C, cudaStream_t st) {
  DAT_REQUIRE(C % 8 == 0, "col2im: C must be a multiple of 8");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, Kp = conv3x3s2_kp(C);
  const long long total = (long long)B * H * W * (C >> 3);
  if (dx_dt == DAT_F32)
    launch_k(col2im3x3s2_kernel<float>, ceil_div(total, 256), 256, 0, st, (const bf16*)dcols, (float*)dx, B, H, W, C, Ho, Wo, Kp, total);
  else
    launch_k(col2im3x3s2_kernel<bf16>, ceil_div(total, 256), 256, 0, st, (const bf16*)dcols, (bf16*)dx, B, H, W, C, Ho, Wo, Kp, total);
  DAT_LAUNCH_OK("col2im3x3s2_kernel");
  return DAT_OK;
}

int conv_weight_pack(const float* w, void* w2, int Cout, int C, cudaStream_t st) {
  const int Kp = conv3x3s2_kp(C);
  launch_k(conv_weight_pack_kernel, ceil_div((long long)Cout * Kp, 256), 256, 0, st, w, (bf16*)w2, Cout, C, Kp);
  DAT_LAUNCH_OK("conv_weight_pack_kernel");
  return DAT_OK;
}
int conv_weight_unpack(const float* dw2, float* dw, int Cout, int C, cudaStream_t st) {
  const int Kp = conv3x3s2_kp(C);
  launch_k(conv_weight_unpack_kernel, ceil_div((long long)Cout * C * 9, 256), 256, 0, st, dw2, dw, Cout, C, Kp);
  DAT_LAUNCH_OK("conv_weight_unpack_kernel");
  return DAT_OK;
}

int gelu_fwd(const void* x, int x_dt, void* y, int y_dt, long long n, cudaStream_t st) {
  DAT_REQUIRE(n % 4 == 0, "gelu_fwd: n must be a multiple of 4");
  const long long n4 = n / 4;
  const int grid = ceil_div(n4, 256);
  if (x_dt == DAT_F32 && y_dt == DAT_F32) launch_k(gelu_fwd_kernel<float, float>, grid, 256, 0, st, (const float*)x, (float*)y, n4);
  else if (x_dt == DAT_F32) launch_k(gelu_fwd_kernel<float, bf16>, grid, 256, 0, st, (const float*)x, (bf16*)y, n4);
  else if (y_dt == DAT_F32) launch_k(gelu_fwd_kernel<bf16, float>, grid, 256, 0, st, (const bf16*)x, (float*)y, n4);
  else launch_k(gelu_fwd_kernel<bf16, bf16>, grid, 256, 0, st, (const bf16*)x, (bf16*)y, n4);
  DAT_LAUNCH_OK("gelu_fwd_kernel");
  return DAT_OK;
}

}  // namespace dat
