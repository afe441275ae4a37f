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
#include <cstdlib>

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

// ---- rows of 64 / 128 / 256 channels: LPR = C / 8 lanes per row, 8 contiguous channels per lane ------------------------
// The warp-per-row kernels above give a lane 2 channels of a 64-channel row (8-byte accesses) and spend a full 5-level
// butterfly per statistic and row: ~100 warp instructions per 64 elements, 40 % (forward) / 61 % (backward) of the HBM
// peak at stage 0.  Here a warp works on 32 / LPR rows at once, a lane moves 32 (fp32) / 16 (bf16) bytes per access
// and a row's statistics take log2(LPR) shuffle levels.
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void round8(const float*, float (&)[8]) {}
__device__ __forceinline__ void round8(const bf16*, float (&v)[8]) {     // the values as a bf16 stream holds them
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
}
template <int LPR> __device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TI, typename TO, int LPR>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_sub_kernel(const TI* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         TO* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows,
                         float eps, const TO* __restrict__ a, const float* __restrict__ scale,
                         long long rows_per_sample, TI* __restrict__ xout) {
  pdl_enter();
  constexpr int C = LPR * 8, RPW = 32 / LPR, UNR = 2;
  const int lane = threadIdx.x & 31, sub = lane / LPR, cl = (lane % LPR) * 8;
  float g[8], b[8];
  ld8(gamma + cl, g);
  ld8(beta + cl, b);
  const float inv_c = 1.0f / (float)C;
  const long long row0 = ((long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5)) * (RPW * UNR) + sub;
  float v[UNR][8];
  bool ok[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const long long row = row0 + u * RPW;
    ok[u] = row < rows;
    if (ok[u]) {
      ld8(x + row * C + cl, v[u]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
    }
  }
  if (a != nullptr) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long row = row0 + u * RPW;
      if (ok[u]) {
        float av[8];
        ld8(a + row * C + cl, av);
        const float sc = scale[row / rows_per_sample];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[u][e] = fmaf(av[e], sc, v[u][e]);
        st8(xout + row * C + cl, v[u]);
        round8(xout, v[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const long long row = row0 + u * RPW;
    float s1 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s1 += v[u][e];
    const float mean = row_sum<LPR>(s1) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = v[u][e] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(row_sum<LPR>(q) * inv_c + eps);
    if (ok[u]) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[u][e] - mean) * rstd * g[e] + b[e];
      st8(y + row * C + cl, o);
      if (cl == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
      }
    }
  }
}

template <typename TI, typename TDY, int LPR>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_sub_kernel(const TDY* __restrict__ dy, const TI* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in, TI* __restrict__ dx,
                         const TI* __restrict__ dres, float* __restrict__ partial, long long rows,
                         TDY* __restrict__ da, const float* __restrict__ scale, long long rows_per_sample) {
  pdl_enter();
  constexpr int C = LPR * 8, RPW = 32 / LPR;
  extern __shared__ float red[];     // [LN_WARPS][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, cl = (lane % LPR) * 8;
  float gam[8], gg[8], gb[8];
  ld8(gamma + cl, gam);
#pragma unroll
  for (int e = 0; e < 8; ++e) gg[e] = gb[e] = 0.f;
  const float inv_c = 1.0f / (float)C;
  for (long long row = ((long long)blockIdx.x * LN_WARPS + warp) * RPW + sub; row - sub < rows;
       row += (long long)gridDim.x * LN_WARPS * RPW) {
    const bool ok = row < rows;
    float xv[8], dv[8], rv[8];
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) xv[e] = dv[e] = rv[e] = 0.f;
    if (ok) {
      mean = mean_in[row];
      rstd = rstd_in[row];
      ld8(x + row * C + cl, xv);
      ld8(dy + row * C + cl, dv);
      if (dres != nullptr) ld8(dres + row * C + cl, rv);
    }
    float xh[8], dg[8];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      xh[e] = ok ? (xv[e] - mean) * rstd : 0.f;
      gg[e] = fmaf(dv[e], xh[e], gg[e]);
      gb[e] += dv[e];
      dg[e] = dv[e] * gam[e];
      m1 += dg[e];
      m2 = fmaf(dg[e], xh[e], m2);
    }
    m1 = row_sum<LPR>(m1) * inv_c;
    m2 = row_sum<LPR>(m2) * inv_c;
    if (ok) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rstd * (dg[e] - m1 - xh[e] * m2) + rv[e];   // + the residual-path gradient
      st8(dx + row * C + cl, o);
      if (da != nullptr) {
        round8(dx, o);
        const float sc = scale[row / rows_per_sample];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] *= sc;
        st8(da + row * C + cl, o);
      }
    }
  }
  // the RPW row groups of the warp hold the same channels: add them (fixed order), then the warps of the CTA
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gg[e] += __shfl_xor_sync(0xffffffffu, gg[e], o);
      gb[e] += __shfl_xor_sync(0xffffffffu, gb[e], o);
    }
  }
  float* mine = red + (size_t)warp * 2 * C;
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mine[cl + e] = gg[e];
      mine[C + cl + e] = gb[e];
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
  static const int ln_v1 = [] { const char* e = std::getenv("DAT_B200_LN_V1"); return e && e[0] == '1' ? 1 : 0; }();
  if ((C == 64 || C == 128 || C == 256) && !ln_v1) {      // narrow rows: several rows per warp, 8 channels per lane
    const int rpw = 32 / (C / 8) * 2;         // rows per warp and pass x 2 passes in flight
    const int grid_s = (int)ceil_div(rows, (long long)LN_WARPS * rpw);
#define LAUNCH_S(TI, TO, L)                                                                                      \
  launch_k(layernorm_fwd_sub_kernel<TI, TO, L>, grid_s, LN_WARPS * 32, 0, st, (const TI*)x, gamma, beta, (TO*)y, \
           mean, rstd, rows, eps, (const TO*)a, scale, rows_per_sample, (TI*)xout)
#define LAUNCH_SL(TI, TO)                                                   \
  do {                                                                      \
    if (C == 64) LAUNCH_S(TI, TO, 8); else if (C == 128) LAUNCH_S(TI, TO, 16); else LAUNCH_S(TI, TO, 32); \
  } while (0)
    if (x_dt == DAT_F32 && y_dt == DAT_F32) LAUNCH_SL(float, float);
    else if (x_dt == DAT_F32) LAUNCH_SL(float, bf16);
    else if (y_dt == DAT_F32) LAUNCH_SL(bf16, float);
    else LAUNCH_SL(bf16, bf16);
#undef LAUNCH_SL
#undef LAUNCH_S
    DAT_LAUNCH_OK("layernorm_fwd_sub_kernel");
    return DAT_OK;
  }
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
  static const int ln_v1 = [] { const char* e = std::getenv("DAT_B200_LN_V1"); return e && e[0] == '1' ? 1 : 0; }();
  if ((C == 64 || C == 128 || C == 256) && !ln_v1) {
#define LAUNCH_S(TI, TD, L)                                                                                          \
  do {                                                                                                               \
    auto kern = layernorm_bwd_sub_kernel<TI, TD, L>;                                                                 \
    int occ = 1;                                                                                                     \
    DAT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LN_WARPS * 32, smem));                     \
    if (occ < 1) occ = 1;                                                                                            \
    if (nblk > 148 * occ) nblk = 148 * occ;                                                                          \
    launch_k(kern, nblk, LN_WARPS * 32, smem, st, (const TD*)dy, (const TI*)x, gamma, mean, rstd, (TI*)dx,           \
             (const TI*)dres, part, rows, (TD*)da, scale, rows_per_sample);                                          \
  } while (0)
#define LAUNCH_SL(TI, TD)                                                   \
  do {                                                                      \
    if (C == 64) LAUNCH_S(TI, TD, 8); else if (C == 128) LAUNCH_S(TI, TD, 16); else LAUNCH_S(TI, TD, 32); \
  } while (0)
    if (x_dt == DAT_F32 && dy_dt == DAT_F32) LAUNCH_SL(float, float);
    else if (x_dt == DAT_F32) LAUNCH_SL(float, bf16);
    else if (dy_dt == DAT_F32) LAUNCH_SL(bf16, float);
    else LAUNCH_SL(bf16, bf16);
#undef LAUNCH_SL
#undef LAUNCH_S
    DAT_LAUNCH_OK("layernorm_bwd_sub_kernel");
    launch_k(layernorm_bwd_reduce_kernel, ceil_div(2 * C, 32), dim3(32, 32), 0, st, part, nblk, C, dgamma, dbeta);
    DAT_LAUNCH_OK("layernorm_bwd_reduce_kernel");
    return DAT_OK;
  }
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
