// Channel-last depthwise k x k convolution (stride 1, "same" padding) with the fusions the DAT
// backbone needs around the deformable-attention block:
//   mode 0:  y = dwconv(x) + b                 'X' mixer, dat.py:118-121 (k = 7)
//   mode 1:  y = dwconv(x) + b + x             local perception unit, dat.py:135-138 (k = 3)
//   mode 2:  z = dwconv(x) + b + x; y = gelu(z) MLP middle, dat_blocks.py:338-343 (k = 3); z is saved
// plus the gradients.  The data gradient is the same kernel run on dz with the flipped filter
// (no bias); `gelu_bwd` forms dz = dy * gelu'(z) for mode 2.
//
// All kernels are HBM-bound streaming kernels: a thread owns 4 consecutive channels of one pixel,
// every tap is a 16-byte (fp32) / 8-byte (bf16) vector read of a channel-last neighbour row, the
// (k*k, C)-transposed filter is read through the read-only cache.  Algorithmic bytes:
// forward R*C*(e_in + e_out) [+ R*C*e_z], data gradient R*C*(e_dz + e_dx), weight gradient
// R*C*(e_dz + e_x).  The k*k-fold spatial reuse is served by L1/L2.
// Weight gradient: per-CTA partial sums over a slice of the pixels, reduced in a fixed order
// (deterministic, no atomics).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

__device__ __forceinline__ float gelu_f(float z) {
  return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_df(float z) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
  const float pdf = expf(-0.5f * z * z) * 0.39894228040143267794f;
  return cdf + z * pdf;
}

// (C, k*k) -> (k*k, C), optionally spatially flipped (for the data gradient)
__global__ void dw_transpose_kernel(const float* __restrict__ w, float* __restrict__ wT, int C, int kk,
                                    int flip) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * kk) return;
  const int c = idx % C, uv = idx / C;
  wT[idx] = w[c * kk + (flip ? kk - 1 - uv : uv)];
}

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<bf16> { static constexpr int N = 8; };

template <int N> __device__ __forceinline__ void ldv(const float* p, float* o) {
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + i);
    o[i] = v.x; o[i + 1] = v.y; o[i + 2] = v.z; o[i + 3] = v.w;
  }
}
template <int N> __device__ __forceinline__ void ldv(const bf16* p, float* o) {
  static_assert(N == 8, "bf16 vectors are 8 wide");
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
template <int N> __device__ __forceinline__ void stv(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
template <int N> __device__ __forceinline__ void stv(bf16* p, const float* v) {
#pragma unroll
  for (int i = 0; i < N; i += 8) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
    *reinterpret_cast<uint4*>(p + i) = raw;
  }
}
// 4-wide bf16 store (8 bytes) for the fp32-in / bf16-out combination
__device__ __forceinline__ void stv4(bf16* p, const float* v) { store4(p, make_float4(v[0], v[1], v[2], v[3])); }
__device__ __forceinline__ void stv4(float* p, const float* v) { store4(p, make_float4(v[0], v[1], v[2], v[3])); }

// CTA = 8 x 8 pixel tile x 4 channel vectors (16 bytes of input each): every tap after the first
// touch of the 10 x 10 (k = 3) halo is an L1 hit, so L2/HBM see x about once.
// grid = (tiles_x * tiles_y * B, C / (4 * VEC)); K = 0 means a runtime k.
template <typename TI, typename TO, int MODE, int K>
__global__ void __launch_bounds__(256)
dwconv_cl_kernel(const TI* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ bias,
                 TO* __restrict__ y, TO* __restrict__ z_out, int B, int H, int W, int C, int k_rt,
                 int tiles_x, int tiles_y) {
  pdl_enter();
  constexpr int VEC = VecOf<TI>::N;
  const int k = K > 0 ? K : k_rt;
  const int p = k >> 1;
  const int tile = blockIdx.x % (tiles_x * tiles_y), b = blockIdx.x / (tiles_x * tiles_y);
  const int vec = threadIdx.x & 3, pix_l = threadIdx.x >> 2;
  const int yy = (tile / tiles_x) * 8 + (pix_l >> 3), xx = (tile % tiles_x) * 8 + (pix_l & 7);
  const int c = (blockIdx.y * 4 + vec) * VEC;
  if (yy >= H || xx >= W || c >= C) return;
  float acc[VEC];
  if (bias != nullptr) ldv<VEC>(bias + c, acc);
  else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  }
  const long long img = (long long)b * H * W;
#pragma unroll
  for (int u = 0; u < (K > 0 ? K : 15); ++u) {
    if (u >= k) break;
    const int y2 = yy + u - p;
    if (y2 < 0 || y2 >= H) continue;
#pragma unroll
    for (int v = 0; v < (K > 0 ? K : 15); ++v) {
      if (v >= k) break;
      const int x2 = xx + v - p;
      if (x2 < 0 || x2 >= W) continue;
      float xv[VEC], wv[VEC];
      ldv<VEC>(x + (img + (long long)y2 * W + x2) * C + c, xv);
      ldv<VEC>(wT + (u * k + v) * C + c, wv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(xv[i], wv[i], acc[i]);
    }
  }
  const long long off = (img + (long long)yy * W + xx) * C + c;
  if (MODE >= 1) {
    float xc[VEC];
    ldv<VEC>(x + off, xc);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += xc[i];
  }
  if (MODE == 2) {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) stv4(z_out + off + i, acc + i);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = gelu_f(acc[i]);
  }
#pragma unroll
  for (int i = 0; i < VEC; i += 4) stv4(y + off + i, acc + i);
}

template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, T* __restrict__ dz,
                                long long n4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 d = load4(dy + 4 * i), zz = load4(z + 4 * i);
  store4(dz + 4 * i, make_float4(d.x * gelu_df(zz.x), d.y * gelu_df(zz.y), d.z * gelu_df(zz.z), d.w * gelu_df(zz.w)));
}

// dw[c][uv] = sum_pix dz[pix][c] * x[pix + off(uv)][c];  db[c] = sum_pix dz[pix][c]
// grid = (ceil(C / (128 * CV)), nsplit); thread = CV consecutive channels (vector loads);
// partial[split][kk + 1][C]
template <int CV> __device__ __forceinline__ void ldc(const float* p, float* o) {
  if (CV == 4) { const float4 v = *reinterpret_cast<const float4*>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  else if (CV == 2) { const float2 v = *reinterpret_cast<const float2*>(p); o[0] = v.x; o[1] = v.y; }
  else o[0] = *p;
}
template <int CV> __device__ __forceinline__ void ldc(const bf16* p, float* o) {
  if (CV == 4) { const float4 v = load4(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  else if (CV == 2) { const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); o[0] = v.x; o[1] = v.y; }
  else o[0] = __bfloat162float(*p);
}

// block = 32 channel vectors x 4 pixel lanes: the 4 lanes stride the CTA's pixel slice (4x shorter
// serial chains than one thread per channel vector), then are summed in shared memory in a
// fixed order.  32-bit index arithmetic, interior pixels skip the per-tap bounds checks.
template <typename TX, typename TD, int KK, int CV>
__global__ void __launch_bounds__(128)
dwconv_wgrad_kernel(const TX* __restrict__ x, const TD* __restrict__ dz, float* __restrict__ partial,
                    int B, int H, int W, int C, int k, long long pix_per_split) {
  pdl_enter();
  constexpr int KS = KK == 9 ? 3 : (KK == 25 ? 5 : 7);   // compile-time filter size: constant tap offsets
  (void)k;
  constexpr int p = KS >> 1;
  __shared__ float red[3][32][(KK + 1) * CV + 1];
  const int cvec = threadIdx.x & 31, plane = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cvec) * CV;
  const bool active = c < C;
  const int npix = B * H * W;
  const int p0 = (int)(blockIdx.y * pix_per_split);
  const int p1 = min(npix, p0 + (int)pix_per_split);
  float acc[KK + 1][CV];
#pragma unroll
  for (int i = 0; i <= KK; ++i)
#pragma unroll
    for (int j = 0; j < CV; ++j) acc[i][j] = 0.f;
  if (active) {
    for (int pix = p0 + plane; pix < p1; pix += 4) {
      const int rem = pix % (H * W);
      const int yy = rem / W, xx = rem - yy * W;
      float d[CV];
      ldc<CV>(dz + (long long)pix * C + c, d);
#pragma unroll
      for (int j = 0; j < CV; ++j) acc[KK][j] += d[j];
      const TX* base = x + ((long long)pix - p * W - p) * C + c;
      const bool interior = yy >= p && yy < H - p && xx >= p && xx < W - p;
      if (interior) {
#pragma unroll
        for (int uv = 0; uv < KK; ++uv) {
          const int u = uv / KS, v = uv - u * KS;
          float xv[CV];
          ldc<CV>(base + ((long long)u * W + v) * C, xv);
#pragma unroll
          for (int j = 0; j < CV; ++j) acc[uv][j] = fmaf(d[j], xv[j], acc[uv][j]);
        }
      } else {
#pragma unroll
        for (int uv = 0; uv < KK; ++uv) {
          const int u = uv / KS, v = uv - u * KS;
          const int y2 = yy + u - p, x2 = xx + v - p;
          if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            float xv[CV];
            ldc<CV>(base + ((long long)u * W + v) * C, xv);
#pragma unroll
            for (int j = 0; j < CV; ++j) acc[uv][j] = fmaf(d[j], xv[j], acc[uv][j]);
          }
        }
      }
    }
  }
  if (plane > 0) {
#pragma unroll
    for (int i = 0; i <= KK; ++i)
#pragma unroll
      for (int j = 0; j < CV; ++j) red[plane - 1][cvec][i * CV + j] = acc[i][j];
  }
  __syncthreads();
  if (plane == 0 && active) {
    float* out = partial + (size_t)blockIdx.y * (KK + 1) * C + c;
#pragma unroll
    for (int i = 0; i <= KK; ++i)
#pragma unroll
      for (int j = 0; j < CV; ++j)
        out[(size_t)i * C + j] = ((acc[i][j] + red[0][cvec][i * CV + j]) + red[1][cvec][i * CV + j]) + red[2][cvec][i * CV + j];
  }
}

// block (32, 32): fixed-order reduction over the splits; writes dw (C, kk) and db (C)
__global__ void dwconv_wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int kk, int C,
                                           float* __restrict__ dw, float* __restrict__ db) {
  pdl_enter();
  __shared__ float red[32][33];
  const int idx = blockIdx.x * 32 + threadIdx.x;    // row * C + c, row in [0, kk]
  const int n_out = (kk + 1) * C;
  float s = 0.f;
  if (idx < n_out)
    for (int zz = threadIdx.y; zz < nsplit; zz += 32) s += partial[(size_t)zz * n_out + idx];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && idx < n_out) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) t += red[l][threadIdx.x];
    const int row = idx / C, c = idx % C;
    if (row < kk) dw[c * kk + row] = t;
    else if (db != nullptr) db[c] = t;
  }
}

int wgrad_splits(long long npix) {
  long long s = (npix + 127) / 128;
  return (int)(s < 1 ? 1 : (s > 1024 ? 1024 : s));
}

}  // namespace

size_t dwconv_workspace(int B, int H, int W, int C, int k) {
  const size_t wt = align_up((size_t)k * k * C * 4, 256);
  const size_t part = align_up((size_t)wgrad_splits((long long)B * H * W) * (k * k + 1) * C * 4, 256);
  size_t fused = dwconv3_supported(C, k) ? dwconv3_partial_bytes(B, H, W, C) : 0;
  if (dwconv7_supported(C, k)) fused = dwconv7_partial_bytes(B, H, W, C);
  return wt + part > fused ? wt + part : fused;
}

int dwconv_wgrad_reduce(const float* partial, int nsplit, int kk, int C, float* dw, float* db, cudaStream_t st) {
  launch_k(dwconv_wgrad_reduce_kernel, ceil_div((kk + 1) * C, 32), dim3(32, 32), 0, st, partial, nsplit, kk, C, dw, db);
  DAT_LAUNCH_OK("dwconv_wgrad_reduce_kernel");
  return DAT_OK;
}

// mode 0/1/2 as above; z_out only for mode 2.  ws >= k*k*C*4 bytes (transposed filter).
int dwconv_fwd(const void* x, int x_dt, const float* w, const float* bias, void* y, void* z_out, int y_dt,
               int B, int H, int W, int C, int k, int mode, int flip, void* ws, size_t ws_bytes,
               cudaStream_t st) {
  if (dwconv3_supported(C, k) && std::getenv("DAT_B200_DWCONV_GENERIC") == nullptr)
    return dwconv3_fwd(x, x_dt, w, bias, y, z_out, y_dt, B, H, W, C, mode, flip, st);
  if (dwconv7_supported(C, k) && mode == 0 && std::getenv("DAT_B200_DWCONV_GENERIC") == nullptr)
    return dwconv7_fwd(x, x_dt, w, bias, y, y_dt, B, H, W, C, flip, st);
  DAT_REQUIRE(C % 4 == 0 && (k & 1) == 1 && k >= 1 && k <= 15, "dwconv: C %% 4 == 0 and odd k <= 15 required");
  DAT_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "dwconv: 16-byte aligned tensors required");
  DAT_REQUIRE(ws_bytes >= (size_t)k * k * C * 4, "dwconv: workspace too small");
  DAT_REQUIRE(mode >= 0 && mode <= 2 && (mode != 2 || z_out != nullptr), "dwconv: bad mode");
  float* wT = (float*)ws;
  launch_k(dw_transpose_kernel, ceil_div(C * k * k, 256), 256, 0, st, w, wT, C, k * k, flip);
  DAT_LAUNCH_OK("dw_transpose_kernel");
  const int tiles_x = ceil_div(W, 8), tiles_y = ceil_div(H, 8);
  const int vec = x_dt == DAT_F32 ? 4 : 8;
  DAT_REQUIRE(C % vec == 0, "dwconv: C must be a multiple of %d for this dtype", vec);
  dim3 grid(tiles_x * tiles_y * B, ceil_div(C, 4 * vec));
#define LAUNCH(TI, TO, MD, KV)                                                                  \
  launch_k(dwconv_cl_kernel<TI, TO, MD, KV>, grid, 256, 0, st, (const TI*)x, wT, bias, (TO*)y, (TO*)z_out, B, \
                                                         H, W, C, k, tiles_x, tiles_y)
#define LAUNCH_K(TI, TO, MD)                                                  \
  do {                                                                        \
    if (k == 3) LAUNCH(TI, TO, MD, 3); else if (k == 7) LAUNCH(TI, TO, MD, 7); \
    else LAUNCH(TI, TO, MD, 0);                                               \
  } while (0)
#define LAUNCH_M(TI, TO)                                                   \
  do {                                                                     \
    if (mode == 0) LAUNCH_K(TI, TO, 0); else if (mode == 1) LAUNCH_K(TI, TO, 1); \
    else LAUNCH_K(TI, TO, 2);                                              \
  } while (0)
  if (x_dt == DAT_F32 && y_dt == DAT_F32) LAUNCH_M(float, float);
  else if (x_dt == DAT_F32) LAUNCH_M(float, bf16);
  else if (y_dt == DAT_F32) LAUNCH_M(bf16, float);
  else LAUNCH_M(bf16, bf16);
#undef LAUNCH_M
#undef LAUNCH_K
#undef LAUNCH
  DAT_LAUNCH_OK("dwconv_cl_kernel");
  return DAT_OK;
}

int gelu_bwd(const void* dy, const void* z, void* dz, int dt, long long n, cudaStream_t st) {
  DAT_REQUIRE(n % 4 == 0, "gelu_bwd: n %% 4 != 0");
  const long long n4 = n / 4;
  if (dt == DAT_F32) launch_k(gelu_bwd_kernel<float>, ceil_div(n4, 256), 256, 0, st, (const float*)dy, (const float*)z, (float*)dz, n4);
  else launch_k(gelu_bwd_kernel<bf16>, ceil_div(n4, 256), 256, 0, st, (const bf16*)dy, (const bf16*)z, (bf16*)dz, n4);
  DAT_LAUNCH_OK("gelu_bwd_kernel");
  return DAT_OK;
}

// dw (C, 1, k, k) and db (C) (db may be NULL), fp32, overwritten.
int dwconv_wgrad(const void* x, int x_dt, const void* dz, int dz_dt, float* dw, float* db, int B, int H,
                 int W, int C, int k, void* ws, size_t ws_bytes, cudaStream_t st) {
  DAT_REQUIRE(k == 3 || k == 5 || k == 7, "dwconv_wgrad: k must be 3, 5 or 7");
  DAT_REQUIRE(ws_bytes >= dwconv_workspace(B, H, W, C, k), "dwconv_wgrad: workspace too small");
  if (dwconv7_supported(C, k) && std::getenv("DAT_B200_DWCONV_GENERIC") == nullptr)
    return dwconv7_wgrad(x, x_dt, dz, dz_dt, dw, db, B, H, W, C, ws, ws_bytes, st);
  const long long npix = (long long)B * H * W;
  const int nsplit = wgrad_splits(npix);
  const long long pps = (npix + nsplit - 1) / nsplit;
  float* part = (float*)((char*)ws + align_up((size_t)k * k * C * 4, 256));
  const int cv = k == 3 ? 4 : (k == 5 ? 2 : 2);      // channels per thread (register budget (k*k+1)*cv)
  DAT_REQUIRE(C % cv == 0, "dwconv_wgrad: C must be a multiple of %d", cv);
  dim3 grid(ceil_div(C, 32 * cv), nsplit);
#define LAUNCH(TX, TD, KKV, CVV)                                                                    \
  launch_k(dwconv_wgrad_kernel<TX, TD, KKV, CVV>, grid, 128, 0, st, (const TX*)x, (const TD*)dz, part, B, H, W, C, k, pps)
#define LAUNCH_K(TX, TD)                                  \
  do {                                                    \
    if (k == 3) LAUNCH(TX, TD, 9, 4); else if (k == 5) LAUNCH(TX, TD, 25, 2); else LAUNCH(TX, TD, 49, 2); \
  } while (0)
  if (x_dt == DAT_F32 && dz_dt == DAT_F32) LAUNCH_K(float, float);
  else if (x_dt == DAT_F32) LAUNCH_K(float, bf16);
  else if (dz_dt == DAT_F32) LAUNCH_K(bf16, float);
  else LAUNCH_K(bf16, bf16);
#undef LAUNCH_K
#undef LAUNCH
  DAT_LAUNCH_OK("dwconv_wgrad_kernel");
  launch_k(dwconv_wgrad_reduce_kernel, ceil_div((k * k + 1) * C, 32), dim3(32, 32), 0, st, part, nsplit, k * k, C, dw, db);
  DAT_LAUNCH_OK("dwconv_wgrad_reduce_kernel");
  return DAT_OK;
}

}  // namespace dat
