// Offset network of the deformable-attention block, fused into one kernel:
//   depthwise k x k strided conv (+bias) -> LayerNorm over Cg -> exact GELU ->
//   1x1 conv Cg->2 (no bias) -> tanh*range*orf | clamp -> + reference point -> pos
// (dat_blocks.py:51-56, :144-162, _get_ref_points :108-121).
//
// Mapping: one warp per sample point (b, g, i, j); lanes stride the Cg channels, so every
// tap of the k x k window is one coalesced channel-last row read of q.  The transposed
// depthwise weights live in shared memory.  LayerNorm statistics and the 2-vector dot are
// warp-shuffle reductions; nothing but pos (and the small saved tensors) is written.
// HBM-bound: algorithmic bytes per launch = B*HW*C*e (q) + B*G*Ns*(Cg*4 + 16) (t, off, pos).
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int OFF_WARPS = 8;

struct OffsetArgs {
  int B, H, W, C, G, Cg, stride, ksize, pad, Hk, Wk, Ns;
  float orf, range_y, range_x;
  long long n_points;  // B * G * Ns
};

__device__ __forceinline__ float gelu_exact(float z) {
  return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad(float z) {
  float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
  float pdf = expf(-0.5f * z * z) * 0.39894228040143267794f;
  return cdf + z * pdf;
}

// dynamic smem: transposed depthwise weights [k*k][Cg]
template <typename TQ, int MAXCPL>
__global__ void __launch_bounds__(OFF_WARPS * 32)
offset_pos_fwd_kernel(const TQ* __restrict__ q, const float* __restrict__ w_dw,
                      const float* __restrict__ b_dw, const float* __restrict__ ln_g,
                      const float* __restrict__ ln_b, const float* __restrict__ w_pw,
                      float* __restrict__ t_dw, float* __restrict__ off_raw,
                      float* __restrict__ pos, OffsetArgs a) {
  pdl_enter();
  extern __shared__ float wsm[];
  const int kk = a.ksize * a.ksize;
  for (int idx = threadIdx.x; idx < kk * a.Cg; idx += blockDim.x) {
    int c = idx % a.Cg, uv = idx / a.Cg;
    wsm[idx] = w_dw[c * kk + uv];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long sp = (long long)blockIdx.x * OFF_WARPS + warp;
  if (sp >= a.n_points) return;
  const int n = (int)(sp % a.Ns);
  const int g = (int)((sp / a.Ns) % a.G);
  const int b = (int)(sp / ((long long)a.Ns * a.G));
  const int i = n / a.Wk, j = n % a.Wk;

  float acc[MAXCPL];
#pragma unroll
  for (int cc = 0; cc < MAXCPL; ++cc) {
    int c = lane + 32 * cc;
    acc[cc] = c < a.Cg ? b_dw[c] : 0.f;
  }
  for (int u = 0; u < a.ksize; ++u) {
    int y = i * a.stride - a.pad + u;
    if (y < 0 || y >= a.H) continue;
    for (int v = 0; v < a.ksize; ++v) {
      int x = j * a.stride - a.pad + v;
      if (x < 0 || x >= a.W) continue;
      const TQ* row = q + (((long long)b * a.H + y) * a.W + x) * a.C + g * a.Cg;
      const float* wrow = wsm + (u * a.ksize + v) * a.Cg;
#pragma unroll
      for (int cc = 0; cc < MAXCPL; ++cc) {
        int c = lane + 32 * cc;
        if (c < a.Cg) acc[cc] = fmaf(wrow[c], to_f32(row[c]), acc[cc]);
      }
    }
  }
  float s1 = 0.f;
#pragma unroll
  for (int cc = 0; cc < MAXCPL; ++cc) {
    int c = lane + 32 * cc;
    if (c < a.Cg) {
      t_dw[sp * a.Cg + c] = acc[cc];
      s1 += acc[cc];
    }
  }
  const float inv_n = 1.0f / (float)a.Cg;
  const float mean = warp_sum(s1) * inv_n;
  float s2 = 0.f;
#pragma unroll
  for (int cc = 0; cc < MAXCPL; ++cc) {
    int c = lane + 32 * cc;
    if (c < a.Cg) {
      float dlt = acc[cc] - mean;
      s2 = fmaf(dlt, dlt, s2);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(s2) * inv_n + 1e-5f);
  float oy = 0.f, ox = 0.f;
#pragma unroll
  for (int cc = 0; cc < MAXCPL; ++cc) {
    int c = lane + 32 * cc;
    if (c < a.Cg) {
      float z = (acc[cc] - mean) * rstd * ln_g[c] + ln_b[c];
      float act = gelu_exact(z);
      oy = fmaf(act, w_pw[c], oy);
      ox = fmaf(act, w_pw[a.Cg + c], ox);
    }
  }
  oy = warp_sum(oy);
  ox = warp_sum(ox);
  if (lane == 0) {
    off_raw[sp * 2 + 0] = oy;
    off_raw[sp * 2 + 1] = ox;
    float ry = ref_point(i, a.Hk), rx = ref_point(j, a.Wk);
    float py, px;
    if (a.orf >= 0.f) {
      py = __fadd_rn(__fmul_rn(__fmul_rn(tanhf(oy), a.range_y), a.orf), ry);
      px = __fadd_rn(__fmul_rn(__fmul_rn(tanhf(ox), a.range_x), a.orf), rx);
    } else {
      py = fminf(fmaxf(__fadd_rn(oy, ry), -1.0f), 1.0f);
      px = fminf(fmaxf(__fadd_rn(ox, rx), -1.0f), 1.0f);
    }
    pos[sp * 2 + 0] = py;
    pos[sp * 2 + 1] = px;
  }
}

__global__ void ref_points_kernel(int Hk, int Wk, float* ry, float* rx) {
  pdl_enter();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < Hk) ry[t] = ref_point(t, Hk);
  if (t < Wk) rx[t] = ref_point(t, Wk);
}

// ---- backward ---------------------------------------------------------------------
// Per sample point: d pos -> d raw offset -> 1x1^T -> GELU' -> LayerNorm backward -> dt
// (stored for the two depthwise-conv gradient kernels) plus per-channel parameter sums.
// Grid-stride over points; per-CTA partial sums [5][Cg] (pw_y, pw_x, ln_g, ln_b, dw_b),
// reduced in a fixed order by offset_bwd_reduce_kernel: deterministic, no atomics.
template <int MAXCPL>
__global__ void __launch_bounds__(OFF_WARPS * 32)
offset_bwd_point_kernel(const float* __restrict__ dpos, const float* __restrict__ off_raw,
                        const float* __restrict__ t_dw, const float* __restrict__ ln_g,
                        const float* __restrict__ ln_b, const float* __restrict__ w_pw,
                        float* __restrict__ dt, float* __restrict__ partial, OffsetArgs a) {
  pdl_enter();
  extern __shared__ float red[];  // [OFF_WARPS][5][Cg]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g_py[MAXCPL], g_px[MAXCPL], g_g[MAXCPL], g_b[MAXCPL], g_b0[MAXCPL];
#pragma unroll
  for (int cc = 0; cc < MAXCPL; ++cc) g_py[cc] = g_px[cc] = g_g[cc] = g_b[cc] = g_b0[cc] = 0.f;
  const float inv_n = 1.0f / (float)a.Cg;

  for (long long sp = (long long)blockIdx.x * OFF_WARPS + warp; sp < a.n_points;
       sp += (long long)gridDim.x * OFF_WARPS) {
    const int n = (int)(sp % a.Ns);
    const int i = n / a.Wk, j = n % a.Wk;
    float oy = off_raw[sp * 2], ox = off_raw[sp * 2 + 1];
    float dy = dpos[sp * 2], dx = dpos[sp * 2 + 1];
    if (a.orf >= 0.f) {
      float ty = tanhf(oy), tx = tanhf(ox);
      dy = dy * a.range_y * a.orf * (1.0f - ty * ty);
      dx = dx * a.range_x * a.orf * (1.0f - tx * tx);
    } else {
      float py = __fadd_rn(oy, ref_point(i, a.Hk)), px = __fadd_rn(ox, ref_point(j, a.Wk));
      if (!(py >= -1.0f && py <= 1.0f)) dy = 0.f;
      if (!(px >= -1.0f && px <= 1.0f)) dx = 0.f;
    }
    float tv[MAXCPL];
    float s1 = 0.f;
#pragma unroll
    for (int cc = 0; cc < MAXCPL; ++cc) {
      int c = lane + 32 * cc;
      tv[cc] = c < a.Cg ? t_dw[sp * a.Cg + c] : 0.f;
      s1 += tv[cc];
    }
    const float mean = warp_sum(s1) * inv_n;
    float s2 = 0.f;
#pragma unroll
    for (int cc = 0; cc < MAXCPL; ++cc) {
      int c = lane + 32 * cc;
      if (c < a.Cg) {
        float dlt = tv[cc] - mean;
        s2 = fmaf(dlt, dlt, s2);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(s2) * inv_n + 1e-5f);
    float dth[MAXCPL];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int cc = 0; cc < MAXCPL; ++cc) {
      int c = lane + 32 * cc;
      dth[cc] = 0.f;
      if (c < a.Cg) {
        float that = (tv[cc] - mean) * rstd;
        float gam = ln_g[c];
        float z = that * gam + ln_b[c];
        float act = gelu_exact(z);
        float dact = dy * w_pw[c] + dx * w_pw[a.Cg + c];
        float dln = dact * gelu_grad(z);
        g_py[cc] = fmaf(dy, act, g_py[cc]);
        g_px[cc] = fmaf(dx, act, g_px[cc]);
        g_g[cc] = fmaf(dln, that, g_g[cc]);
        g_b[cc] += dln;
        dth[cc] = dln * gam;
        m1 += dth[cc];
        m2 = fmaf(dth[cc], that, m2);
        tv[cc] = that;
      }
    }
    m1 = warp_sum(m1) * inv_n;
    m2 = warp_sum(m2) * inv_n;
#pragma unroll
    for (int cc = 0; cc < MAXCPL; ++cc) {
      int c = lane + 32 * cc;
      if (c < a.Cg) {
        float d = rstd * (dth[cc] - m1 - tv[cc] * m2);
        dt[sp * a.Cg + c] = d;
        g_b0[cc] += d;
      }
    }
  }
  // CTA reduction in a fixed warp order
#pragma unroll
  for (int cc = 0; cc < MAXCPL; ++cc) {
    int c = lane + 32 * cc;
    if (c < a.Cg) {
      float* r = red + (size_t)warp * 5 * a.Cg;
      r[0 * a.Cg + c] = g_py[cc];
      r[1 * a.Cg + c] = g_px[cc];
      r[2 * a.Cg + c] = g_g[cc];
      r[3 * a.Cg + c] = g_b[cc];
      r[4 * a.Cg + c] = g_b0[cc];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 5 * a.Cg; idx += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < OFF_WARPS; ++w) s += red[(size_t)w * 5 * a.Cg + idx];
    partial[(size_t)blockIdx.x * 5 * a.Cg + idx] = s;
  }
}

// block (32, 32): threadIdx.x = output within a 32-wide slice, threadIdx.y = partial lane; fixed summation order
// (lane-strided partial sums with 4 independent accumulators, then lanes 0..31): deterministic.  (The first version
// walked all per-CTA partials - 2048 at stage 2 - with one dependent load after the other: 15 us per launch.)
__global__ void offset_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int Cg,
                                         float* __restrict__ g_pw, float* __restrict__ g_ln_g,
                                         float* __restrict__ g_ln_b, float* __restrict__ g_dw_b) {
  pdl_enter();
  __shared__ float red[32][33];
  const int idx = blockIdx.x * 32 + threadIdx.x, zl = threadIdx.y;
  const size_t pitch = (size_t)5 * Cg;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (idx < 5 * Cg) {
    int z = zl;
    for (; z + 96 < nblocks; z += 128) {
      s0 += partial[(size_t)z * pitch + idx];
      s1 += partial[(size_t)(z + 32) * pitch + idx];
      s2 += partial[(size_t)(z + 64) * pitch + idx];
      s3 += partial[(size_t)(z + 96) * pitch + idx];
    }
    for (; z < nblocks; z += 32) s0 += partial[(size_t)z * pitch + idx];
  }
  red[zl][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (zl == 0 && idx < 5 * Cg) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) s += red[l][threadIdx.x];
    const int row = idx / Cg, c = idx % Cg;
    if (row < 2) g_pw[row * Cg + c] = s;
    else if (row == 2) g_ln_g[c] = s;
    else if (row == 3) g_ln_b[c] = s;
    else g_dw_b[c] = s;
  }
}

// Depthwise weight gradient: dw[c,u,v] = sum_points dt[point, c] * q[window(point,u,v), c].
// One CTA per chunk of sample points; thread = (channel, tap group) with up to WG_ACC taps
// accumulated in registers, so every q / dt access is a coalesced channel row and the taps
// give independent loads in flight.  partial[chunk][uv][c], reduced in a fixed order.
constexpr int WG_ACC = 24;
template <typename TQ>
__global__ void __launch_bounds__(256)
offset_bwd_wgrad_kernel(const TQ* __restrict__ q, const float* __restrict__ dt,
                        float* __restrict__ partial, long long pts_per_split, OffsetArgs a) {
  pdl_enter();
  const int kk = a.ksize * a.ksize;
  const int cthreads = a.Cg < (int)blockDim.x ? a.Cg : (int)blockDim.x;   // channels per pass
  const int uvg_n = blockDim.x / cthreads;                                // tap groups
  const int cl = threadIdx.x % cthreads, uvg = threadIdx.x / cthreads;
  const long long p0 = (long long)blockIdx.x * pts_per_split;
  const long long p1 = min(a.n_points, p0 + pts_per_split);
  if (uvg >= uvg_n) return;
  for (int c = cl; c < a.Cg; c += cthreads) {
    for (int uv0 = 0; uv0 < kk; uv0 += uvg_n * WG_ACC) {
      float acc[WG_ACC];
#pragma unroll
      for (int t = 0; t < WG_ACC; ++t) acc[t] = 0.f;
#pragma unroll 2
      for (long long sp = p0; sp < p1; ++sp) {
        const int n = (int)(sp % a.Ns);
        const int g = (int)((sp / a.Ns) % a.G);
        const int b = (int)(sp / ((long long)a.Ns * a.G));
        const int yb = (n / a.Wk) * a.stride - a.pad, xb = (n % a.Wk) * a.stride - a.pad;
        const float dtv = dt[sp * a.Cg + c];
        const TQ* qb = q + (long long)b * a.H * a.W * a.C + g * a.Cg + c;
#pragma unroll
        for (int t = 0; t < WG_ACC; ++t) {
          const int uv = uv0 + uvg + t * uvg_n;
          if (uv < kk) {
            const int y = yb + uv / a.ksize, x = xb + uv % a.ksize;
            if (y >= 0 && y < a.H && x >= 0 && x < a.W)
              acc[t] = fmaf(dtv, to_f32(qb[((long long)y * a.W + x) * a.C]), acc[t]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < WG_ACC; ++t) {
        const int uv = uv0 + uvg + t * uvg_n;
        if (uv < kk) partial[((size_t)blockIdx.x * kk + uv) * a.Cg + c] = acc[t];
      }
    }
  }
}

// block (32, 32): threadIdx.x = output within a 32-wide slice, threadIdx.y = partial lane.
// Fixed summation order (lane-strided partial sums, then lanes 0..31): deterministic.
__global__ void offset_bwd_wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit,
                                               int kk, int Cg, float* __restrict__ g_dw_w) {
  pdl_enter();
  __shared__ float red[32][33];
  const int idx = blockIdx.x * 32 + threadIdx.x;  // uv * Cg + c
  const int n_out = kk * Cg;
  float s = 0.f;
  if (idx < n_out)
    for (int z = threadIdx.y; z < nsplit; z += 32) s += partial[(size_t)z * n_out + idx];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && idx < n_out) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) t += red[l][threadIdx.x];
    const int uv = idx / Cg, c = idx % Cg;
    g_dw_w[c * kk + uv] = t;
  }
}

// Depthwise data gradient in gather form (no atomics): every q element sums the <=
// ceil(k/s)^2 sample points whose window covers it, and adds that to dq in place.
template <typename TQ>
__global__ void offset_bwd_dgrad_kernel(const float* __restrict__ dt,
                                        const float* __restrict__ w_dw, TQ* __restrict__ dq,
                                        long long total, OffsetArgs a) {
  pdl_enter();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cf = (int)(idx % a.C);
  const long long pix = idx / a.C;
  const int x = (int)(pix % a.W);
  const int y = (int)((pix / a.W) % a.H);
  const int b = (int)(pix / ((long long)a.W * a.H));
  const int g = cf / a.Cg, c = cf % a.Cg;
  const int kk = a.ksize * a.ksize;
  // i*s - p + u = y, 0 <= u < k  ->  (y + p - k + 1)/s <= i <= (y + p)/s
  int i_lo = y + a.pad - a.ksize + 1;
  i_lo = i_lo <= 0 ? 0 : (i_lo + a.stride - 1) / a.stride;
  int i_hi = min(a.Hk - 1, (y + a.pad) / a.stride);
  int j_lo = x + a.pad - a.ksize + 1;
  j_lo = j_lo <= 0 ? 0 : (j_lo + a.stride - 1) / a.stride;
  int j_hi = min(a.Wk - 1, (x + a.pad) / a.stride);
  float s = 0.f;
  for (int i = i_lo; i <= i_hi; ++i) {
    int u = y + a.pad - i * a.stride;
    for (int j = j_lo; j <= j_hi; ++j) {
      int v = x + a.pad - j * a.stride;
      long long sp = ((long long)b * a.G + g) * a.Ns + i * a.Wk + j;
      s = fmaf(w_dw[c * kk + u * a.ksize + v], dt[sp * a.Cg + c], s);
    }
  }
  dq[idx] = from_f32<TQ>(to_f32(dq[idx]) + s);
}

// ---- compile-time-k versions for Cg % 64 == 0 (every DAT / DAT++ configuration: Cg = 64) ------
// A lane owns 2 consecutive channels of a 64-channel chunk, so each q / dt / dq access of a warp
// is one contiguous 128-256 byte run, and all the window index arithmetic is warp-uniform with
// the divisions by k folded at compile time.

__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld_pair(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st_pair(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st_pair(bf16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}

constexpr int WGR_PG = 2;     // point groups per CTA (their sums are merged in shared memory)

// Weight gradient: warp (u, pg) owns tap row u for the points p0 + pg, p0 + pg + PG, ...:
// per point one dt load and K contiguous-pixel q loads, 2K FMAs into registers.
// partial[split][u * K + v][Cg]
template <typename TQ, int K>
__global__ void __launch_bounds__(K * WGR_PG * 32)
offset_bwd_wgrad_rows_kernel(const TQ* __restrict__ q, const float* __restrict__ dt,
                             float* __restrict__ partial, int pts_per_split, OffsetArgs a) {
  pdl_enter();
  __shared__ float2 red[(WGR_PG - 1) * K * K][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int u = warp % K, pg = warp / K;
  const int npts = (int)a.n_points;
  const int p0 = blockIdx.x * pts_per_split;
  const int p1 = min(npts, p0 + pts_per_split);
  for (int cc = 0; cc < a.Cg; cc += 64) {
    const int c = cc + lane * 2;
    float2 acc[K];
#pragma unroll
    for (int v = 0; v < K; ++v) acc[v] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int sp = p0 + pg; sp < p1; sp += WGR_PG) {
      const int n = sp % a.Ns, bg = sp / a.Ns;
      const int g = bg % a.G, b = bg / a.G;
      const int y = (n / a.Wk) * a.stride - a.pad + u, xb = (n % a.Wk) * a.stride - a.pad;
      if (y < 0 || y >= a.H) continue;
      const float2 d = ld_pair(dt + (long long)sp * a.Cg + c);
      const TQ* qrow = q + ((long long)(b * a.H + y) * a.W + xb) * a.C + g * a.Cg + c;
#pragma unroll
      for (int v = 0; v < K; ++v) {
        const int x = xb + v;
        if (x >= 0 && x < a.W) {
          const float2 qv = ld_pair(qrow + (long long)v * a.C);
          acc[v].x = fmaf(d.x, qv.x, acc[v].x);
          acc[v].y = fmaf(d.y, qv.y, acc[v].y);
        }
      }
    }
    if (pg > 0) {
#pragma unroll
      for (int v = 0; v < K; ++v) red[((pg - 1) * K + u) * K + v][lane] = acc[v];
    }
    __syncthreads();
    if (pg == 0) {
      float* out = partial + ((size_t)blockIdx.x * K * K + u * K) * a.Cg + c;
#pragma unroll
      for (int v = 0; v < K; ++v) {
        float2 t = acc[v];
#pragma unroll
        for (int m = 1; m < WGR_PG; ++m) { const float2 o = red[((m - 1) * K + u) * K + v][lane]; t.x += o.x; t.y += o.y; }
        *reinterpret_cast<float2*>(out + (size_t)v * a.Cg) = t;
      }
    }
    __syncthreads();
  }
}

// Data gradient, gather form: thread = one pixel x 8 channels (16-byte accesses to dq, 32-byte to dt and the
// transposed filter in shared memory); the <= ceil(k/s)^2 sample points whose window covers the pixel are
// enumerated per thread, all of them (<= 3 x 3) loaded before the first FMA.  A CTA walks image rows (the row's
// window range i_lo..i_hi is computed once per row).  The index arithmetic is paid once per 8 channels: the
// kernel was bound by it, not by memory (2 channels per lane: 118 us at stage 0 for 67 MB of traffic).
template <typename TQ, int K, int MW>   // MW: windows per dimension of the unrolled path, >= ceil(k / stride) for it to be taken
__global__ void __launch_bounds__(256)
offset_bwd_dgrad_warp_kernel(const float* __restrict__ dt, const float* __restrict__ w_dw,
                             TQ* __restrict__ dq, int n_rows, OffsetArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float wT[];          // [K * K][Cg]
  for (int i = threadIdx.x; i < K * K * a.Cg; i += blockDim.x) {
    const int uv = i / a.Cg, c = i - uv * a.Cg;
    wT[i] = w_dw[c * K * K + uv];
  }
  __syncthreads();
  const int c8s = a.C >> 3;              // 8-channel units per pixel
  const int units = a.W * c8s;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int b = row / a.H, y = row - b * a.H;
    // i*s - p + u = y, 0 <= u < k  ->  (y + p - k + 1)/s <= i <= (y + p)/s
    int i_lo = y + a.pad - K + 1;
    i_lo = i_lo <= 0 ? 0 : (i_lo + a.stride - 1) / a.stride;
    const int i_hi = min(a.Hk - 1, (y + a.pad) / a.stride);
    const int ni = i_hi - i_lo + 1;
    for (int unit = threadIdx.x; unit < units; unit += blockDim.x) {
      const int x = unit / c8s, cf = (unit - x * c8s) * 8;
      const int g = cf / a.Cg, c = cf - g * a.Cg;
      int j_lo = x + a.pad - K + 1;
      j_lo = j_lo <= 0 ? 0 : (j_lo + a.stride - 1) / a.stride;
      const int j_hi = min(a.Wk - 1, (x + a.pad) / a.stride);
      const int nj = j_hi - j_lo + 1;
      const float* dtb = dt + ((long long)(b * a.G + g) * a.Ns) * a.Cg + c;
      TQ* dst = dq + ((long long)row * a.W + x) * a.C + cf;
      const float4 o0 = load4(dst), o1 = load4(dst + 4);
      float acc[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      if (ni <= MW && nj <= MW) {
        // global loads of every covering window first (one round trip), the filter taps come from shared
        // memory inside the FMA loop
        float4 d[MW][MW][2];
#pragma unroll
        for (int ia = 0; ia < MW; ++ia)
#pragma unroll
          for (int jb = 0; jb < MW; ++jb) {
            const bool ok = ia < ni && jb < nj;
            const float* dp = dtb + (long long)((i_lo + ia) * a.Wk + j_lo + jb) * a.Cg;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            d[ia][jb][0] = ok ? *reinterpret_cast<const float4*>(dp) : z;
            d[ia][jb][1] = ok ? *reinterpret_cast<const float4*>(dp + 4) : z;
          }
#pragma unroll
        for (int ia = 0; ia < MW; ++ia)
#pragma unroll
          for (int jb = 0; jb < MW; ++jb) {
            if (ia < ni && jb < nj) {
              const int u = y + a.pad - (i_lo + ia) * a.stride, v = x + a.pad - (j_lo + jb) * a.stride;
              const float* wp = wT + (u * K + v) * a.Cg + c;
              const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
              acc[0] = fmaf(w0.x, d[ia][jb][0].x, acc[0]); acc[1] = fmaf(w0.y, d[ia][jb][0].y, acc[1]);
              acc[2] = fmaf(w0.z, d[ia][jb][0].z, acc[2]); acc[3] = fmaf(w0.w, d[ia][jb][0].w, acc[3]);
              acc[4] = fmaf(w1.x, d[ia][jb][1].x, acc[4]); acc[5] = fmaf(w1.y, d[ia][jb][1].y, acc[5]);
              acc[6] = fmaf(w1.z, d[ia][jb][1].z, acc[6]); acc[7] = fmaf(w1.w, d[ia][jb][1].w, acc[7]);
            }
          }
      } else {
        for (int i = i_lo; i <= i_hi; ++i) {
          const int u = y + a.pad - i * a.stride;
          for (int j = j_lo; j <= j_hi; ++j) {
            const int v = x + a.pad - j * a.stride;
            const float* dp = dtb + (long long)(i * a.Wk + j) * a.Cg;
            const float* wp = wT + (u * K + v) * a.Cg + c;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(wp[e], dp[e], acc[e]);
          }
        }
      }
      store4(dst, make_float4(acc[0], acc[1], acc[2], acc[3]));
      store4(dst + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
  }
}

bool offset_bwd_fast_supported(const Shape& s) {
  return s.Cg % 64 == 0 && s.Cg <= 128 && (s.ksize == 3 || s.ksize == 5 || s.ksize == 7 || s.ksize == 9) &&
         (long long)s.B * s.G * s.Ns < (1ll << 30) && (long long)s.B * s.HW * (s.C / 64) < (1ll << 31);
}

OffsetArgs make_args(const Shape& s) {
  OffsetArgs a;
  a.B = s.B; a.H = s.H; a.W = s.W; a.C = s.C; a.G = s.G; a.Cg = s.Cg;
  a.stride = s.stride; a.ksize = s.ksize; a.pad = s.pad; a.Hk = s.Hk; a.Wk = s.Wk; a.Ns = s.Ns;
  a.orf = s.orf;
  // the reference builds the range as a Python double rounded to fp32 (dat_blocks.py:150)
  a.range_y = (float)(1.0 / ((double)s.Hk - 1.0));
  a.range_x = (float)(1.0 / ((double)s.Wk - 1.0));
  a.n_points = (long long)s.B * s.G * s.Ns;
  return a;
}

int pick_cpl(int Cg) {
  int cpl = (Cg + 31) / 32;
  if (cpl <= 2) return 2;
  if (cpl <= 4) return 4;
  if (cpl <= 8) return 8;
  return 16;
}

}  // namespace

int offset_pos_fwd(const Shape& s, const dat_block_params* p, const void* q, float* t_dw,
                   float* off_raw, float* pos, cudaStream_t st) {
  if (offset_pos_fwd_vec_supported(s)) return offset_pos_fwd_vec(s, p, q, t_dw, off_raw, pos, st);
  DAT_REQUIRE(s.Cg <= 512, "offset net: Cg=%d > 512 unsupported", s.Cg);
  OffsetArgs a = make_args(s);
  size_t smem = (size_t)s.ksize * s.ksize * s.Cg * sizeof(float);
  DAT_REQUIRE(smem <= 200 * 1024, "offset net: k*k*Cg too large for shared memory");
  int grid = ceil_div(a.n_points, OFF_WARPS);
  int cpl = pick_cpl(s.Cg);
#define LAUNCH(TQ, CPL)                                                                        \
  do {                                                                                         \
    auto kern = offset_pos_fwd_kernel<TQ, CPL>;                                                \
    if (smem > 48 * 1024)                                                                      \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                       (int)smem));                                            \
    launch_k(kern, grid, OFF_WARPS * 32, smem, st, (const TQ*)q, p->off_dw_w, p->off_dw_b,           \
                                             p->off_ln_g, p->off_ln_b, p->off_pw_w, t_dw,      \
                                             off_raw, pos, a);                                 \
  } while (0)
#define LAUNCH_T(TQ)                                    \
  do {                                                  \
    if (cpl == 2) LAUNCH(TQ, 2);                        \
    else if (cpl == 4) LAUNCH(TQ, 4);                   \
    else if (cpl == 8) LAUNCH(TQ, 8);                   \
    else LAUNCH(TQ, 16);                                \
  } while (0)
  if (s.act_dtype == DAT_F32) LAUNCH_T(float); else LAUNCH_T(bf16);
#undef LAUNCH_T
#undef LAUNCH
  DAT_LAUNCH_OK("offset_pos_fwd_kernel");
  return DAT_OK;
}

int ref_points(int Hk, int Wk, float* ry, float* rx, cudaStream_t st) {
  DAT_REQUIRE(Hk > 1 && Wk > 1, "ref_points: Hk, Wk must be > 1");
  int n = Hk > Wk ? Hk : Wk;
  launch_k(ref_points_kernel, ceil_div(n, 128), 128, 0, st, Hk, Wk, ry, rx);
  DAT_LAUNCH_OK("ref_points_kernel");
  return DAT_OK;
}

static int offset_bwd_blocks(const Shape& s) {
  long long pts = (long long)s.B * s.G * s.Ns;
  long long want = (pts + OFF_WARPS - 1) / OFF_WARPS;
  return (int)(want < 296 ? want : 296);  // 2 CTAs per SM
}
static int offset_wgrad_splits(const Shape& s) {
  long long pts = (long long)s.B * s.G * s.Ns;
  if (offset_bwd_fast_supported(s)) {   // ~2.5 CTAs per SM; at least 8 points each
    long long per = (pts + 383) / 384;
    if (per < 8) per = 8;
    return (int)((pts + per - 1) / per);
  }
  long long sp = (pts + 7) / 8;      // ~8 points per CTA: latency-bound loop, so go wide
  return (int)(sp < 1 ? 1 : (sp > 4096 ? 4096 : sp));
}

size_t offset_bwd_workspace(const Shape& s) {
  size_t dt = align_up((size_t)s.B * s.G * s.Ns * s.Cg * sizeof(float), 256);
  size_t p1 = align_up((size_t)offset_bwd_blocks(s) * 5 * s.Cg * sizeof(float), 256);
  size_t p2 = align_up((size_t)offset_wgrad_splits(s) * s.ksize * s.ksize * s.Cg * sizeof(float), 256);
  return dt + p1 + p2;
}

// dpos (B,G,Ns,2) -> parameter grads of the offset net, and dq += depthwise data gradient.
int offset_bwd(const Shape& s, const dat_block_params* p, const void* q, const float* t_dw,
               const float* off_raw, const float* dpos, void* dq, const dat_block_grads* g,
               void* ws, size_t ws_bytes, cudaStream_t st, cudaStream_t pst, cudaEvent_t fork_ev) {
  DAT_REQUIRE(s.Cg <= 512, "offset net: Cg=%d > 512 unsupported", s.Cg);
  DAT_REQUIRE(ws_bytes >= offset_bwd_workspace(s), "offset_bwd: workspace too small");
  OffsetArgs a = make_args(s);
  const int kk = s.ksize * s.ksize;
  float* dt = (float*)ws;
  float* part1 = (float*)((char*)ws + align_up((size_t)a.n_points * s.Cg * sizeof(float), 256));
  int nblk = offset_bwd_blocks(s);
  float* part2 = (float*)((char*)part1 + align_up((size_t)nblk * 5 * s.Cg * sizeof(float), 256));
  size_t smem = (size_t)OFF_WARPS * 5 * s.Cg * sizeof(float);
  int cpl = pick_cpl(s.Cg);
#define LAUNCH(CPL)                                                                          \
  do {                                                                                       \
    auto kern = offset_bwd_point_kernel<CPL>;                                                \
    if (smem > 48 * 1024)                                                                    \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                       (int)smem));                                          \
    launch_k(kern, nblk, OFF_WARPS * 32, smem, st, dpos, off_raw, t_dw, p->off_ln_g, p->off_ln_b,  \
                                             p->off_pw_w, dt, part1, a);                     \
  } while (0)
  if (cpl == 2) LAUNCH(2); else if (cpl == 4) LAUNCH(4); else if (cpl == 8) LAUNCH(8); else LAUNCH(16);
#undef LAUNCH
  DAT_LAUNCH_OK("offset_bwd_point_kernel");
  // the parameter gradients (reductions, depthwise weight gradient) are off the critical path: they
  // go to the side stream `pst` (== st: no fork); only the data gradient below stays on `st`
  if (pst != st) {
    DAT_CUDA_OK(cudaEventRecord(fork_ev, st));
    DAT_CUDA_OK(cudaStreamWaitEvent(pst, fork_ev, 0));
  }
  launch_k(offset_bwd_reduce_kernel, ceil_div(5 * s.Cg, 32), dim3(32, 32), 0, pst,
           part1, nblk, s.Cg, g->off_pw_w, g->off_ln_g, g->off_ln_b, g->off_dw_b);
  DAT_LAUNCH_OK("offset_bwd_reduce_kernel");

  int nsplit = offset_wgrad_splits(s);
  long long pps = (a.n_points + nsplit - 1) / nsplit;
  const bool fast = offset_bwd_fast_supported(s);
  if (fast) {
#define LAUNCH_WG(TQ, KV) \
    launch_k(offset_bwd_wgrad_rows_kernel<TQ, KV>, nsplit, KV * WGR_PG * 32, 0, pst, (const TQ*)q, dt, part2, (int)pps, a)
#define LAUNCH_WG_K(TQ)                                                             \
    do {                                                                            \
      if (s.ksize == 3) LAUNCH_WG(TQ, 3); else if (s.ksize == 5) LAUNCH_WG(TQ, 5);  \
      else if (s.ksize == 7) LAUNCH_WG(TQ, 7); else LAUNCH_WG(TQ, 9);               \
    } while (0)
    if (s.act_dtype == DAT_F32) LAUNCH_WG_K(float); else LAUNCH_WG_K(bf16);
#undef LAUNCH_WG_K
#undef LAUNCH_WG
  } else if (s.act_dtype == DAT_F32)
    launch_k(offset_bwd_wgrad_kernel<float>, nsplit, 256, 0, pst, (const float*)q, dt, part2, pps, a);
  else
    launch_k(offset_bwd_wgrad_kernel<bf16>, nsplit, 256, 0, pst, (const bf16*)q, dt, part2, pps, a);
  DAT_LAUNCH_OK("offset_bwd_wgrad_kernel");
  launch_k(offset_bwd_wgrad_reduce_kernel, ceil_div(kk * s.Cg, 32), dim3(32, 32), 0, pst, part2, nsplit, kk, s.Cg,
                                                                                    g->off_dw_w);
  DAT_LAUNCH_OK("offset_bwd_wgrad_reduce_kernel");

  long long total = (long long)s.B * s.HW * s.C;
  if (fast) {
    const int n_items = s.B * s.H;            // image rows; a CTA's 8 warps share the (pixel, chunk) items of a row
    const size_t wsm = (size_t)kk * s.Cg * sizeof(float);
    const int dg_grid = n_items < 148 * 8 ? n_items : 148 * 8;
    const bool two = (s.ksize + s.stride - 1) / s.stride <= 2;   // covering windows per dimension
#define LAUNCH_DG(TQ, KV)                                                                             \
    do {                                                                                              \
      if (two) launch_k(offset_bwd_dgrad_warp_kernel<TQ, KV, 2>, dg_grid, 256, wsm, st, dt, p->off_dw_w, (TQ*)dq, n_items, a); \
      else launch_k(offset_bwd_dgrad_warp_kernel<TQ, KV, 3>, dg_grid, 256, wsm, st, dt, p->off_dw_w, (TQ*)dq, n_items, a);     \
    } while (0)
#define LAUNCH_DG_K(TQ)                                                             \
    do {                                                                            \
      if (s.ksize == 3) LAUNCH_DG(TQ, 3); else if (s.ksize == 5) LAUNCH_DG(TQ, 5);  \
      else if (s.ksize == 7) LAUNCH_DG(TQ, 7); else LAUNCH_DG(TQ, 9);               \
    } while (0)
    if (s.act_dtype == DAT_F32) LAUNCH_DG_K(float); else LAUNCH_DG_K(bf16);
#undef LAUNCH_DG_K
#undef LAUNCH_DG
    DAT_LAUNCH_OK("offset_bwd_dgrad_warp_kernel");
    return DAT_OK;
  }
  if (s.act_dtype == DAT_F32)
    launch_k(offset_bwd_dgrad_kernel<float>, ceil_div(total, 256), 256, 0, st, dt, p->off_dw_w, (float*)dq, total, a);
  else
    launch_k(offset_bwd_dgrad_kernel<bf16>, ceil_div(total, 256), 256, 0, st, dt, p->off_dw_w, (bf16*)dq, total, a);
  DAT_LAUNCH_OK("offset_bwd_dgrad_kernel");
  return DAT_OK;
}

}  // namespace dat
