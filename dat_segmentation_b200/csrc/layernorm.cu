// Channel-last LayerNorm over C (the reference's LayerNormProxy, dat_blocks.py:229-240:
// NCHW -> NHWC view -> nn.LayerNorm(C, eps=1e-5) -> back), forward and backward.
// the LayerNorm that feeds every block (dat.py:147-151).
//
// HBM-bound streaming kernels: one warp per pixel row, lanes stride the channels with
// 8-byte (fp32) / 4-byte (bf16) vector accesses, row statistics by warp shuffles, fp32 math.
// Algorithmic bytes: forward R*C*(e_in + e_out) + 8R; backward R*C*(e_dy + e_x + e_dx) + 8R.
// Backward: gamma / beta gradients are accumulated per lane over a grid-stride loop, reduced
// per CTA in shared memory and across CTAs by a second tiny kernel in a fixed order
// (deterministic, no atomics).
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXV = 16;      // C <= 32 * 2 * 16 = 1024

__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
// the value as the stream holds it after the store (fp32: unchanged, bf16: rounded)
__device__ __forceinline__ float2 ld2_round(const float*, float2 v) { return v; }
__device__ __forceinline__ float2 ld2_round(const bf16*, float2 v) {
  return __bfloat1622float2(__floats2bfloat162_rn(v.x, v.y));
}
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st2(bf16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}

// NV = number of 64-channel groups (C = 64 * NV exactly, or C == 32 handled with NV = 1 and half
// the lanes idle).  ROWS = rows a warp handles at once: all their loads are issued before the
// first reduction, so narrow rows (C = 64: 8 bytes per lane) still keep enough bytes in flight.
template <int NV> struct RowsOf { static constexpr int R = NV == 1 ? 4 : (NV == 2 ? 2 : 1); };

template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, TO* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, long long rows, int C, float eps,
                     const TO* __restrict__ a, const float* __restrict__ scale, long long rows_per_sample,
                     TI* __restrict__ xout) {
  // a != NULL: the residual add with stochastic depth that precedes the norm (dat.py:147-151) rides along -
  // xout = x + a * scale[sample] is formed in registers, stored as the new residual stream and normalised, so the
  // stream is not written and re-read by a separate kernel
  pdl_enter();
  constexpr int R = RowsOf<NV>::R;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  float2 v[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const TI* xr = x + (row0 + r) * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 2;
      v[r][i] = (c < C && row0 + r < rows) ? ld2(xr + c) : make_float2(0.f, 0.f);
    }
    if (a != nullptr && row0 + r < rows) {
      const float sc = scale[(row0 + r) / rows_per_sample];
      const TO* ar = a + (row0 + r) * C;
      TI* xo = xout + (row0 + r) * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 2;
        if (c < C) {
          const float2 av = ld2(ar + c);
          v[r][i] = make_float2(fmaf(av.x, sc, v[r][i].x), fmaf(av.y, sc, v[r][i].y));
          st2(xo + c, v[r][i]);
          v[r][i] = ld2_round(xo + c, v[r][i]);
        }
      }
    }
  }
  float2 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 2;
    g[i] = c < C ? ld2(gamma + c) : make_float2(0.f, 0.f);
    b[i] = c < C ? ld2(beta + c) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[r][i].x + v[r][i].y;
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 2;
      if (c < C) {
        const float a0 = v[r][i].x - mean, a1 = v[r][i].y - mean;
        q += a0 * a0 + a1 * a1;
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
    if (row < rows) {
      TO* yr = y + row * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 2;
        if (c < C)
          st2(yr + c, make_float2((v[r][i].x - mean) * rstd * g[i].x + b[i].x,
                                  (v[r][i].y - mean) * rstd * g[i].y + b[i].y));
      }
      if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
      }
    }
  }
}

template <typename TI, typename TDY, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const TDY* __restrict__ dy, const TI* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, TI* __restrict__ dx,
                     const TI* __restrict__ dres, float* __restrict__ partial, long long rows, int C,
                     TDY* __restrict__ da, const float* __restrict__ scale, long long rows_per_sample) {
  // da != NULL: the branch gradient of the fused residual add, da = dx * scale[sample], leaves in the same pass
  pdl_enter();
  constexpr int R = NV == 1 ? 2 : 1;   // 4 rows / iteration measured slower: registers -> occupancy
  extern __shared__ float red[];     // [LN_WARPS][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 gg[NV], gb[NV], gam[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 2;
    gg[i] = gb[i] = make_float2(0.f, 0.f);
    gam[i] = c < C ? ld2(gamma + c) : make_float2(0.f, 0.f);
  }
  const float inv_c = 1.0f / (float)C;
  for (long long row0 = ((long long)blockIdx.x * LN_WARPS + warp) * R; row0 < rows;
       row0 += (long long)gridDim.x * LN_WARPS * R) {
    float2 xv[R][NV], dv[R][NV], rv[R][NV];
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {          // every load of the R rows is in flight before the first use
      const long long row = row0 + r;
      const bool ok = row < rows;
      mean[r] = ok ? mean_in[row] : 0.f;
      rstd[r] = ok ? rstd_in[row] : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 2;
        const bool in = ok && c < C;
        xv[r][i] = in ? ld2(x + row * C + c) : make_float2(0.f, 0.f);
        dv[r][i] = in ? ld2(dy + row * C + c) : make_float2(0.f, 0.f);
        rv[r][i] = (in && dres != nullptr) ? ld2(dres + row * C + c) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
      float2 xh[NV], dg[NV];
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 2;
        xh[i] = dg[i] = make_float2(0.f, 0.f);
        if (c < C && row < rows) {
          const float2 d = dv[r][i];
          xh[i] = make_float2((xv[r][i].x - mean[r]) * rstd[r], (xv[r][i].y - mean[r]) * rstd[r]);
          gg[i].x = fmaf(d.x, xh[i].x, gg[i].x);
          gg[i].y = fmaf(d.y, xh[i].y, gg[i].y);
          gb[i].x += d.x;
          gb[i].y += d.y;
          dg[i] = make_float2(d.x * gam[i].x, d.y * gam[i].y);
          m1 += dg[i].x + dg[i].y;
          m2 += dg[i].x * xh[i].x + dg[i].y * xh[i].y;
        }
      }
      m1 = warp_sum(m1) * inv_c;
      m2 = warp_sum(m2) * inv_c;
      if (row < rows) {
        TI* dxr = dx + row * C;
        const float sc = da != nullptr ? scale[row / rows_per_sample] : 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 2;
          if (c < C) {
            // + gradient of the residual path through x (dat.py:147-156)
            const float2 g2 = make_float2(rstd[r] * (dg[i].x - m1 - xh[i].x * m2) + rv[r][i].x,
                                          rstd[r] * (dg[i].y - m1 - xh[i].y * m2) + rv[r][i].y);
            st2(dxr + c, g2);
            if (da != nullptr) {
              const float2 gr = ld2_round(dxr + c, g2);      // what a separate kernel would read back from dx
              st2(da + row * C + c, make_float2(gr.x * sc, gr.y * sc));
            }
          }
        }
      }
    }
  }
  float* mine = red + (size_t)warp * 2 * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 2;
    if (c < C) {
      mine[c] = gg[i].x; mine[c + 1] = gg[i].y;
      mine[C + c] = gb[i].x; mine[C + c + 1] = gb[i].y;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += red[(size_t)w * 2 * C + idx];
    partial[(size_t)blockIdx.x * 2 * C + idx] = s;
  }
}

// block (32, 32): threadIdx.x = output within a 32-wide slice, threadIdx.y = partial lane; fixed
// summation order (lane-strided partial sums, then lanes 0..31): deterministic.
__global__ void layernorm_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int C,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_enter();
  __shared__ float red[32][33];
  const int idx = blockIdx.x * 32 + threadIdx.x, zl = threadIdx.y;
  float s = 0.f;
  if (idx < 2 * C)
    for (int z = zl; z < nblocks; z += 32) s += partial[(size_t)z * 2 * C + idx];
  red[zl][threadIdx.x] = s;
  __syncthreads();
  if (zl == 0 && idx < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) t += red[l][threadIdx.x];
    if (idx < C) dgamma[idx] = t; else dbeta[idx - C] = t;
  }
}

int nv_of(int C) { return C <= 64 ? 1 : (C + 63) / 64; }

int ln_bwd_blocks(long long rows) {
  long long want = (rows + LN_WARPS - 1) / LN_WARPS;
  return (int)(want < 592 ? want : 592);    // 4 CTAs per SM
}

}  // namespace

size_t layernorm_bwd_workspace(long long rows, int C) {
  return align_up((size_t)ln_bwd_blocks(rows) * 2 * C * sizeof(float), 256);
}

int layernorm_fwd(const void* x, int x_dt, const float* gamma, const float* beta, void* y, int y_dt,
                  float* mean, float* rstd, long long rows, int C, float eps, cudaStream_t st) {
  return residual_layernorm_fwd(nullptr, nullptr, 1, x, x_dt, gamma, beta, nullptr, y, y_dt, mean, rstd, rows, C, eps, st);
}

// a (may be NULL; y's dtype): xout = x + a * scale[row / rows_per_sample] (x's dtype), y = LayerNorm(xout)
int residual_layernorm_fwd(const void* a, const float* scale, long long rows_per_sample, const void* x, int x_dt,
                           const float* gamma, const float* beta, void* xout, void* y, int y_dt, float* mean,
                           float* rstd, long long rows, int C, float eps, cudaStream_t st) {
  DAT_REQUIRE(rows > 0 && C >= 2 && C % 2 == 0 && C <= 64 * LN_MAXV, "layernorm: unsupported C=%d", C);
  DAT_REQUIRE(a == nullptr || (scale != nullptr && xout != nullptr && rows_per_sample > 0),
              "residual_layernorm_fwd: scale / xout missing");
  const int nv = nv_of(C);
  const int rows_per_warp = nv == 1 ? 4 : (nv == 2 ? 2 : 1);      // RowsOf<NV>::R
  const int grid = (int)ceil_div(rows, (long long)LN_WARPS * rows_per_warp);
#define LAUNCH(TI, TO, NVV)                                                                    \
  launch_k(layernorm_fwd_kernel<TI, TO, NVV>, grid, LN_WARPS * 32, 0, st, (const TI*)x, gamma, beta, \
           (TO*)y, mean, rstd, rows, C, eps, (const TO*)a, scale, rows_per_sample, (TI*)xout)
#define LAUNCH_NV(TI, TO)                                                      \
  do {                                                                         \
    if (nv == 1) LAUNCH(TI, TO, 1); else if (nv == 2) LAUNCH(TI, TO, 2);       \
    else if (nv <= 4) LAUNCH(TI, TO, 4); else if (nv <= 8) LAUNCH(TI, TO, 8);  \
    else LAUNCH(TI, TO, 16);                                                   \
  } while (0)
  if (x_dt == DAT_F32 && y_dt == DAT_F32) LAUNCH_NV(float, float);
  else if (x_dt == DAT_F32) LAUNCH_NV(float, bf16);
  else if (y_dt == DAT_F32) LAUNCH_NV(bf16, float);
  else LAUNCH_NV(bf16, bf16);
#undef LAUNCH_NV
#undef LAUNCH
  DAT_LAUNCH_OK("layernorm_fwd_kernel");
  return DAT_OK;
}

int layernorm_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const float* gamma,
                  const float* mean, const float* rstd, void* dx, const void* dres, float* dgamma,
                  float* dbeta, long long rows, int C, void* ws, size_t ws_bytes, cudaStream_t st) {
  return residual_layernorm_bwd(dy, dy_dt, x, x_dt, gamma, mean, rstd, dx, dres, nullptr, nullptr, 1, dgamma, dbeta,
                                rows, C, ws, ws_bytes, st);
}

// da (may be NULL; dy's dtype): da = dx * scale[row / rows_per_sample], the gradient of the branch of the fused add
int residual_layernorm_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const float* gamma,
                           const float* mean, const float* rstd, void* dx, const void* dres, void* da,
                           const float* scale, long long rows_per_sample, float* dgamma, float* dbeta,
                           long long rows, int C, void* ws, size_t ws_bytes, cudaStream_t st) {
  DAT_REQUIRE(rows > 0 && C >= 2 && C % 2 == 0 && C <= 64 * LN_MAXV, "layernorm: unsupported C=%d", C);
  DAT_REQUIRE(da == nullptr || (scale != nullptr && rows_per_sample > 0), "residual_layernorm_bwd: scale missing");
  DAT_REQUIRE(ws_bytes >= layernorm_bwd_workspace(rows, C), "layernorm_bwd: workspace too small");
  int nblk = ln_bwd_blocks(rows);
  const int nv = nv_of(C);
  const size_t smem = (size_t)LN_WARPS * 2 * C * sizeof(float);
  float* part = (float*)ws;
#define LAUNCH(TI, TD, NVV)                                                                       \
  do {                                                                                            \
    auto kern = layernorm_bwd_kernel<TI, TD, NVV>;                                                \
    if (smem > 48 * 1024)                                                                         \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    /* persistent grid = exactly the CTAs that are resident at once (one wave) */                 \
    int occ = 1;                                                                                  \
    DAT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LN_WARPS * 32, smem));  \
    if (occ < 1) occ = 1;                                                                         \
    if (nblk > 148 * occ) nblk = 148 * occ;                                                       \
    launch_k(kern, nblk, LN_WARPS * 32, smem, st, (const TD*)dy, (const TI*)x, gamma, mean, rstd, (TI*)dx, \
             (const TI*)dres, part, rows, C, (TD*)da, scale, rows_per_sample);                          \
  } while (0)
#define LAUNCH_NV(TI, TD)                                                      \
  do {                                                                         \
    if (nv == 1) LAUNCH(TI, TD, 1); else if (nv == 2) LAUNCH(TI, TD, 2);       \
    else if (nv <= 4) LAUNCH(TI, TD, 4); else if (nv <= 8) LAUNCH(TI, TD, 8);  \
    else LAUNCH(TI, TD, 16);                                                   \
  } while (0)
  if (x_dt == DAT_F32 && dy_dt == DAT_F32) LAUNCH_NV(float, float);
  else if (x_dt == DAT_F32) LAUNCH_NV(float, bf16);
  else if (dy_dt == DAT_F32) LAUNCH_NV(bf16, float);
  else LAUNCH_NV(bf16, bf16);
#undef LAUNCH_NV
#undef LAUNCH
  DAT_LAUNCH_OK("layernorm_bwd_kernel");
  launch_k(layernorm_bwd_reduce_kernel, ceil_div(2 * C, 32), dim3(32, 32), 0, st, part, nblk, C, dgamma, dbeta);
  DAT_LAUNCH_OK("layernorm_bwd_reduce_kernel");
  return DAT_OK;
}

}  // namespace dat
