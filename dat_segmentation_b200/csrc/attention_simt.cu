// Fused attention core of the block on CUDA cores, fp32 math throughout:
//   S = (Q K^T) * hc^-0.5 + bilinear(rpe_table, (q_grid - pos) / 2);  P = softmax(S);  O = P V
// (dat_blocks.py:180-223; bias branch :198-214, q grid :123-136), plus its backward.
//
// This is the exact-precision path (fp32 parity to 1e-5 needs fp32 products) and the
// shape-generic fallback; the bf16 hot path is the tcgen05 kernel in attention_tc.cu.
// The score matrix, displacement tensor and bias tensor of the reference (3 x 8.4 MB per
// image at stage 2) are never materialised: each thread owns one query, streams the
// sampled keys/values of its head through shared memory and evaluates the 4-tap bias
// on the fly from a shared-memory copy of the head's table (online softmax).
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int HC = DAT_HEAD_DIM;  // 32
constexpr int ATT_THREADS = 128;
constexpr int KV_CHUNK = 64;
constexpr int Q_CHUNK = 32;

struct AttnArgs {
  int B, H, W, HW, C, heads, G, hg, Ns, Th, Tw;
  float scale;
  int table_in_smem;
  // bias modes other than the rpe table (BM_DENSE): bias (Bb, heads, HW, Ns) fp32 added to the scaled
  // scores, batch stride 0 when shared by the batch; dbias (B, heads, HW, Ns) receives dS.
  const float* bias;
  long long bias_bstride;
  float* dbias;
};

// bias source of the score: BM_RPE bilinear rpe table (shipped configs), BM_NONE (use_pe False, dwc_pe,
// no_off), BM_DENSE a materialised bias tensor (fixed_pe, log_cpb — the reference materialises it too)
constexpr int BM_RPE = 0, BM_NONE = 1, BM_DENSE = 2;

struct BiasEval {
  float val, d_dix, d_diy;  // value, d/d(ix), d/d(iy) of the bilinear interpolant
  int i00;                  // y0 * Tw + x0
  float w00, w01, w10, w11; // tap weights with validity folded in (0 when out of range)
  float wx0, wx1, wy0, wy1; // separable factors (w00 = wx0 * wy0, ...)
  int vmask;                // validity bits: 1 = (y0,x0), 2 = (y0,x0+1), 4 = (y0+1,x0), 8 = (y0+1,x0+1)
};

// Bilinear sample of one head's table at displacement ((gy - py)/2, (gx - px)/2);
// same op order as the reference: sub, mul 0.5, then grid_sample's unnormalise.
template <bool WITH_GRAD>
__device__ __forceinline__ BiasEval rpe_bias_eval(const float* __restrict__ tab, int Th, int Tw,
                                                  float gy, float gx, float py, float px) {
  const float dy = __fmul_rn(__fsub_rn(gy, py), 0.5f);
  const float dx = __fmul_rn(__fsub_rn(gx, px), 0.5f);
  const Taps t = make_taps(dx, dy, Tw, Th);
  BiasEval r;
  r.i00 = t.y0 * Tw + t.x0;
  const bool v00 = t.vx0 && t.vy0, v01 = t.vx1 && t.vy0, v10 = t.vx0 && t.vy1, v11 = t.vx1 && t.vy1;
  const float t00 = v00 ? tab[r.i00] : 0.f;
  const float t01 = v01 ? tab[r.i00 + 1] : 0.f;
  const float t10 = v10 ? tab[r.i00 + Tw] : 0.f;
  const float t11 = v11 ? tab[r.i00 + Tw + 1] : 0.f;
  r.w00 = v00 ? __fmul_rn(t.wx0, t.wy0) : 0.f;
  r.w01 = v01 ? __fmul_rn(t.wx1, t.wy0) : 0.f;
  r.w10 = v10 ? __fmul_rn(t.wx0, t.wy1) : 0.f;
  r.w11 = v11 ? __fmul_rn(t.wx1, t.wy1) : 0.f;
  r.val = t00 * r.w00 + t01 * r.w01 + t10 * r.w10 + t11 * r.w11;
  r.wx0 = t.wx0; r.wx1 = t.wx1; r.wy0 = t.wy0; r.wy1 = t.wy1;
  r.vmask = (v00 ? 1 : 0) | (v01 ? 2 : 0) | (v10 ? 4 : 0) | (v11 ? 8 : 0);
  if (WITH_GRAD) {
    r.d_dix = (t01 - t00) * t.wy0 + (t11 - t10) * t.wy1;
    r.d_diy = (t10 - t00) * t.wx0 + (t11 - t01) * t.wx1;
  } else {
    r.d_dix = r.d_diy = 0.f;
  }
  return r;
}

// d table += ds * (4 tap weights), warp-aggregated.  Lanes are consecutive queries of one
// image row, so neighbouring lanes hit the same table cell (the table step per query is
// < 1 cell): runs of lanes with the same (row, north-west cell) are summed with a
// segmented shuffle reduction (windows of 8 lanes) and only the window leaders issue the
// shared-memory atomics - ~3x fewer, conflict-free within the warp.  Must be called by all
// 32 lanes.
__device__ __forceinline__ void table_grad_scatter(float* dtab, int Tw, const BiasEval& be,
                                                   float ds, int row) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int pk = __shfl_up_sync(FULL, be.i00, 1);
  const int pr = __shfl_up_sync(FULL, row, 1);
  const bool head = lane == 0 || pk != be.i00 || pr != row;
  const unsigned heads = __ballot_sync(FULL, head);
  const int start = 31 - __clz((int)(heads & (FULL >> (31 - lane))));
  float P = ds * be.wx0, Q = ds * be.wx1;
#pragma unroll
  for (int d = 1; d <= 4; d <<= 1) {
    const float oP = __shfl_down_sync(FULL, P, d), oQ = __shfl_down_sync(FULL, Q, d);
    const int os = __shfl_down_sync(FULL, start, d);
    if (lane + d < 32 && os == start) { P += oP; Q += oQ; }
  }
  if (((lane - start) & 7) == 0 && (P != 0.f || Q != 0.f)) {
    float* cell = dtab + be.i00;
    if (be.vmask & 1) atomicAdd(cell, P * be.wy0);
    if (be.vmask & 2) atomicAdd(cell + 1, Q * be.wy0);
    if (be.vmask & 4) atomicAdd(cell + Tw, P * be.wy1);
    if (be.vmask & 8) atomicAdd(cell + Tw + 1, Q * be.wy1);
  }
}

template <typename T>
__device__ __forceinline__ void load_row32(const T* __restrict__ p, float* dst) {
#pragma unroll
  for (int i = 0; i < HC / 4; ++i) {
    float4 v = load4(p + 4 * i);
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}
template <typename T>
__device__ __forceinline__ void store_row32(T* __restrict__ p, const float* src) {
#pragma unroll
  for (int i = 0; i < HC / 4; ++i)
    store4(p + 4 * i, make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]));
}

// cooperative stage of `rows` rows x 32 channels of one head into smem [rows][32]
template <typename T>
__device__ __forceinline__ void stage_rows(const T* __restrict__ src, long long row_stride,
                                           int rows_valid, int rows, float* dst) {
  for (int idx = threadIdx.x; idx < rows * (HC / 4); idx += blockDim.x) {
    int r = idx / (HC / 4), c4 = idx % (HC / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) v = load4(src + (long long)r * row_stride + 4 * c4);
    *reinterpret_cast<float4*>(dst + r * HC + 4 * c4) = v;
  }
}

// smem layout: [table Th*Tw | ks KV_CHUNK*32 | vs KV_CHUNK*32 | ps KV_CHUNK*2]
template <typename T, int BM>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                     const float* __restrict__ pos, const float* __restrict__ table,
                     T* __restrict__ o, float* __restrict__ lse, AttnArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  const int tsz = a.table_in_smem ? ((a.Th * a.Tw + 3) & ~3) : 0;
  float* tab_s = smem;
  float* ks = smem + tsz;
  float* vs = ks + KV_CHUNK * HC;
  float* ps = vs + KV_CHUNK * HC;
  const int bh = blockIdx.y, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  const float* tab_g = BM == BM_RPE ? table + (long long)eta * a.Th * a.Tw : nullptr;
  if (BM == BM_RPE && a.table_in_smem)
    for (int i = threadIdx.x; i < a.Th * a.Tw; i += blockDim.x) tab_s[i] = tab_g[i];
  const float* tab = a.table_in_smem ? tab_s : tab_g;

  const int m_raw = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool valid = m_raw < a.HW;
  const int m = valid ? m_raw : a.HW - 1;
  const float gy = query_point(m / a.W, a.H), gx = query_point(m % a.W, a.W);
  const float* brow = BM == BM_DENSE ? a.bias + (long long)b * a.bias_bstride + ((long long)eta * a.HW + m) * a.Ns : nullptr;
  float qr[HC], acc[HC];
  load_row32(q + ((long long)b * a.HW + m) * a.C + eta * HC, qr);
#pragma unroll
  for (int c = 0; c < HC; ++c) acc[c] = 0.f;
  float mx = -INFINITY, l = 0.f;

  const T* kbase = k + (long long)b * a.Ns * a.C + eta * HC;
  const T* vbase = v + (long long)b * a.Ns * a.C + eta * HC;
  const float* pbase = pos + ((long long)b * a.G + g) * a.Ns * 2;
  for (int n0 = 0; n0 < a.Ns; n0 += KV_CHUNK) {
    const int nv = min(KV_CHUNK, a.Ns - n0);
    __syncthreads();
    stage_rows(kbase + (long long)n0 * a.C, a.C, nv, KV_CHUNK, ks);
    stage_rows(vbase + (long long)n0 * a.C, a.C, nv, KV_CHUNK, vs);
    for (int i = threadIdx.x; i < 2 * nv; i += blockDim.x) ps[i] = pbase[2 * n0 + i];
    __syncthreads();
    for (int n = 0; n < nv; ++n) {
      const float* kr = ks + n * HC;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < HC; ++c) s = fmaf(qr[c], kr[c], s);
      s = s * a.scale;
      if (BM == BM_RPE) s += rpe_bias_eval<false>(tab, a.Th, a.Tw, gy, gx, ps[2 * n], ps[2 * n + 1]).val;
      if (BM == BM_DENSE) s += brow[n0 + n];
      if (s > mx) {
        const float corr = expf(mx - s);  // exp(-inf) = 0 on the first key
        l *= corr;
#pragma unroll
        for (int c = 0; c < HC; ++c) acc[c] *= corr;
        mx = s;
      }
      const float p = expf(s - mx);
      l += p;
      const float* vr = vs + n * HC;
#pragma unroll
      for (int c = 0; c < HC; ++c) acc[c] = fmaf(p, vr[c], acc[c]);
    }
  }
  if (valid) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int c = 0; c < HC; ++c) acc[c] *= inv;
    store_row32(o + ((long long)b * a.HW + m) * a.C + eta * HC, acc);
    lse[(long long)bh * a.HW + m] = mx + logf(l);
  }
}

// bias (B, heads, HW, Ns) alone — test hook for the rpe path
__global__ void rpe_bias_kernel(const float* __restrict__ pos, const float* __restrict__ table,
                                float* __restrict__ bias, AttnArgs a, long long total) {
  pdl_enter();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % a.Ns);
  const int m = (int)((idx / a.Ns) % a.HW);
  const int bh = (int)(idx / ((long long)a.Ns * a.HW));
  const int b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  const float* pp = pos + (((long long)b * a.G + g) * a.Ns + n) * 2;
  bias[idx] = rpe_bias_eval<false>(table + (long long)eta * a.Th * a.Tw, a.Th, a.Tw,
                                   query_point(m / a.W, a.H), query_point(m % a.W, a.W), pp[0],
                                   pp[1]).val;
}

// delta[b, eta, m] = sum_c dO * O over the head's 32 channels
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ d_o, const T* __restrict__ o,
                                  float* __restrict__ delta, int HW, int C, int heads,
                                  long long total) {
  pdl_enter();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, m, eta)
  if (idx >= total) return;
  const int eta = (int)(idx % heads);
  const long long bm = idx / heads;
  const int m = (int)(bm % HW);
  const int b = (int)(bm / HW);
  const T* pa = d_o + bm * C + eta * HC;
  const T* pb = o + bm * C + eta * HC;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HC / 4; ++i) {
    float4 x = load4(pa + 4 * i), y = load4(pb + 4 * i);
    s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  delta[((long long)b * heads + eta) * HW + m] = s;
}

// Query-parallel backward: dQ and the rpe-table gradient.
// The table gradient is accumulated in a per-CTA shared-memory copy (shared atomics) and
// flushed with one global atomicAdd per touched cell; d_table must be zeroed beforehand.
template <typename T, int BM>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                   const T* __restrict__ d_o, const float* __restrict__ lse,
                   const float* __restrict__ delta, const float* __restrict__ pos,
                   const float* __restrict__ table, T* __restrict__ dq,
                   float* __restrict__ d_table, AttnArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  const int tsz = (a.Th * a.Tw + 3) & ~3;
  float* tab_s = smem;
  float* dtab_s = smem + tsz;
  float* ks = dtab_s + tsz;
  float* vs = ks + KV_CHUNK * HC;
  float* ps = vs + KV_CHUNK * HC;
  const int bh = blockIdx.y, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  if (BM == BM_RPE) {
    const float* tab_g = table + (long long)eta * a.Th * a.Tw;
    for (int i = threadIdx.x; i < a.Th * a.Tw; i += blockDim.x) {
      tab_s[i] = tab_g[i];
      dtab_s[i] = 0.f;
    }
  }
  const int m_raw = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool valid = m_raw < a.HW;
  const int m = valid ? m_raw : a.HW - 1;
  const float gy = query_point(m / a.W, a.H), gx = query_point(m % a.W, a.W);
  const long long brow_off = ((long long)eta * a.HW + m) * a.Ns;
  const float* brow = BM == BM_DENSE ? a.bias + (long long)b * a.bias_bstride + brow_off : nullptr;
  float* dbrow = BM == BM_DENSE ? a.dbias + (long long)b * a.heads * a.HW * a.Ns + brow_off : nullptr;
  float qr[HC], dor[HC], acc[HC];
  load_row32(q + ((long long)b * a.HW + m) * a.C + eta * HC, qr);
  load_row32(d_o + ((long long)b * a.HW + m) * a.C + eta * HC, dor);
#pragma unroll
  for (int c = 0; c < HC; ++c) acc[c] = 0.f;
  const float lse_m = lse[(long long)bh * a.HW + m];
  const float dl_m = delta[(long long)bh * a.HW + m];

  const T* kbase = k + (long long)b * a.Ns * a.C + eta * HC;
  const T* vbase = v + (long long)b * a.Ns * a.C + eta * HC;
  const float* pbase = pos + ((long long)b * a.G + g) * a.Ns * 2;
  for (int n0 = 0; n0 < a.Ns; n0 += KV_CHUNK) {
    const int nv = min(KV_CHUNK, a.Ns - n0);
    __syncthreads();
    stage_rows(kbase + (long long)n0 * a.C, a.C, nv, KV_CHUNK, ks);
    stage_rows(vbase + (long long)n0 * a.C, a.C, nv, KV_CHUNK, vs);
    for (int i = threadIdx.x; i < 2 * nv; i += blockDim.x) ps[i] = pbase[2 * n0 + i];
    __syncthreads();
    for (int n = 0; n < nv; ++n) {
      const float* kr = ks + n * HC;
      const float* vr = vs + n * HC;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        s = fmaf(qr[c], kr[c], s);
        dp = fmaf(dor[c], vr[c], dp);
      }
      BiasEval be;
      float bval = 0.f;
      if (BM == BM_RPE) {
        be = rpe_bias_eval<false>(tab_s, a.Th, a.Tw, gy, gx, ps[2 * n], ps[2 * n + 1]);
        bval = be.val;
      }
      if (BM == BM_DENSE) bval = brow[n0 + n];
      const float p = expf(s * a.scale + bval - lse_m);
      const float ds = valid ? p * (dp - dl_m) : 0.f;
#pragma unroll
      for (int c = 0; c < HC; ++c) acc[c] = fmaf(ds, kr[c], acc[c]);
      if (BM == BM_RPE) table_grad_scatter(dtab_s, a.Tw, be, ds, m / a.W);
      if (BM == BM_DENSE && valid) dbrow[n0 + n] = ds;
    }
  }
  if (valid) {
#pragma unroll
    for (int c = 0; c < HC; ++c) acc[c] *= a.scale;
    store_row32(dq + ((long long)b * a.HW + m) * a.C + eta * HC, acc);
  }
  if (BM == BM_RPE) {
    __syncthreads();
    float* dt_g = d_table + (long long)eta * a.Th * a.Tw;
    for (int i = threadIdx.x; i < a.Th * a.Tw; i += blockDim.x) {
      float vv = dtab_s[i];
      if (vv != 0.f) atomicAdd(dt_g + i, vv);
    }
  }
}

// Key-parallel backward: dK, dV and the rpe-bias part of d pos.  Each thread owns one
// sampled key; queries of the CTA's q-split stream through shared memory.  Outputs are
// per-split partials (fixed-order reduction afterwards): no atomics.
//   dk_part/dv_part: (qsplit, B, Ns, C) fp32;  dpos_part: (B, heads, qsplit, Ns, 2) fp32
template <typename T, int BM>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const T* __restrict__ d_o, const float* __restrict__ lse,
                    const float* __restrict__ delta, const float* __restrict__ pos,
                    const float* __restrict__ table, float* __restrict__ dk_part,
                    float* __restrict__ dv_part, float* __restrict__ dpos_part, int q_per_split,
                    AttnArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float smem[];
  const int tsz = a.table_in_smem ? ((a.Th * a.Tw + 3) & ~3) : 0;
  float* tab_s = smem;
  float* qs = smem + tsz;
  float* dos = qs + Q_CHUNK * HC;
  float* ls = dos + Q_CHUNK * HC;     // lse
  float* dls = ls + Q_CHUNK;          // delta
  const int bh = blockIdx.y, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  const int z = blockIdx.z, qsplit = gridDim.z;
  const float* tab_g = BM == BM_RPE ? table + (long long)eta * a.Th * a.Tw : nullptr;
  if (BM == BM_RPE && a.table_in_smem)
    for (int i = threadIdx.x; i < a.Th * a.Tw; i += blockDim.x) tab_s[i] = tab_g[i];
  const float* tab = a.table_in_smem ? tab_s : tab_g;
  const float* bcol = BM == BM_DENSE ? a.bias + (long long)b * a.bias_bstride + (long long)eta * a.HW * a.Ns : nullptr;

  const int n_raw = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool valid = n_raw < a.Ns;
  const int n = valid ? n_raw : a.Ns - 1;
  float kr[HC], vr[HC], dk[HC], dv[HC];
  load_row32(k + ((long long)b * a.Ns + n) * a.C + eta * HC, kr);
  load_row32(v + ((long long)b * a.Ns + n) * a.C + eta * HC, vr);
#pragma unroll
  for (int c = 0; c < HC; ++c) dk[c] = dv[c] = 0.f;
  const float* pp = pos + (((long long)b * a.G + g) * a.Ns + n) * 2;
  const float py = pp[0], px = pp[1];
  float dpy = 0.f, dpx = 0.f;

  const int m_begin = z * q_per_split, m_end = min(a.HW, m_begin + q_per_split);
  const T* qbase = q + (long long)b * a.HW * a.C + eta * HC;
  const T* dobase = d_o + (long long)b * a.HW * a.C + eta * HC;
  for (int m0 = m_begin; m0 < m_end; m0 += Q_CHUNK) {
    const int mv = min(Q_CHUNK, m_end - m0);
    __syncthreads();
    stage_rows(qbase + (long long)m0 * a.C, a.C, mv, Q_CHUNK, qs);
    stage_rows(dobase + (long long)m0 * a.C, a.C, mv, Q_CHUNK, dos);
    for (int i = threadIdx.x; i < mv; i += blockDim.x) {
      ls[i] = lse[(long long)bh * a.HW + m0 + i];
      dls[i] = delta[(long long)bh * a.HW + m0 + i];
    }
    __syncthreads();
    for (int mi = 0; mi < mv; ++mi) {
      const int m = m0 + mi;
      const float* qr = qs + mi * HC;
      const float* dor = dos + mi * HC;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        s = fmaf(qr[c], kr[c], s);
        dp = fmaf(dor[c], vr[c], dp);
      }
      float bval = 0.f, d_dix = 0.f, d_diy = 0.f;
      if (BM == BM_RPE) {
        const BiasEval be = rpe_bias_eval<true>(tab, a.Th, a.Tw, query_point(m / a.W, a.H),
                                                query_point(m % a.W, a.W), py, px);
        bval = be.val; d_dix = be.d_dix; d_diy = be.d_diy;
      }
      if (BM == BM_DENSE) bval = bcol[(long long)m * a.Ns + n];
      const float p = expf(s * a.scale + bval - ls[mi]);
      const float ds = p * (dp - dls[mi]);
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        dv[c] = fmaf(p, dor[c], dv[c]);
        dk[c] = fmaf(ds, qr[c], dk[c]);
      }
      dpx = fmaf(ds, d_dix, dpx);
      dpy = fmaf(ds, d_diy, dpy);
    }
  }
  if (valid) {
#pragma unroll
    for (int c = 0; c < HC; ++c) dk[c] *= a.scale;
    const long long row = ((long long)z * a.B + b) * a.Ns + n;
    store_row32(dk_part + row * a.C + eta * HC, dk);
    store_row32(dv_part + row * a.C + eta * HC, dv);
    float* dpo = dpos_part + ((((long long)b * a.heads + eta) * qsplit + z) * a.Ns + n) * 2;
    // ix = ((d + 1)/2)(Tw - 1), d = (grid - pos)/2  ->  d ix / d pos = -(Tw - 1)/4
    dpo[0] = dpy * (-0.25f * (float)(a.Th - 1));
    dpo[1] = dpx * (-0.25f * (float)(a.Tw - 1));
  }
}

AttnArgs make_args(const Shape& s) {
  AttnArgs a;
  a.B = s.B; a.H = s.H; a.W = s.W; a.HW = s.HW; a.C = s.C; a.heads = s.heads; a.G = s.G;
  a.hg = s.hg; a.Ns = s.Ns; a.Th = s.Th; a.Tw = s.Tw;
  a.scale = 1.0f / sqrtf((float)HC);
  a.table_in_smem = ((size_t)s.Th * s.Tw * sizeof(float) <= 96 * 1024) ? 1 : 0;
  a.bias = nullptr; a.bias_bstride = 0; a.dbias = nullptr;
  if (s.pe_mode != DAT_PE_RPE) { a.Th = a.Tw = 0; a.table_in_smem = 1; }   // no table: no shared-memory copy
  return a;
}

int bias_mode(const Shape& s) {
  return s.pe_mode == DAT_PE_RPE ? BM_RPE : (s.pe_mode == DAT_PE_FIXED || s.pe_mode == DAT_PE_LOGCPB) ? BM_DENSE : BM_NONE;
}

template <typename K>
int set_smem(K kern, size_t smem) {
  if (smem > 48 * 1024)
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return DAT_OK;
}

// dtype x bias-mode dispatch: F is a generic lambda called with (T value, integral_constant<int, BM>)
template <typename F>
int dispatch(int act_dtype, int bm, F&& f) {
#define DAT_BM_CASE(T)                                                         \
  switch (bm) {                                                                \
    case BM_RPE: return f(T(), std::integral_constant<int, BM_RPE>());        \
    case BM_NONE: return f(T(), std::integral_constant<int, BM_NONE>());      \
    default: return f(T(), std::integral_constant<int, BM_DENSE>());          \
  }
  if (act_dtype == DAT_F32) { DAT_BM_CASE(float) }
  DAT_BM_CASE(bf16)
#undef DAT_BM_CASE
}

}  // namespace

int attention_fwd_simt(const Shape& s, const void* q, const void* k, const void* v,
                       const float* pos, const float* table, void* o, float* lse,
                       cudaStream_t st, const float* bias, long long bias_bstride) {
  AttnArgs a = make_args(s);
  a.bias = bias; a.bias_bstride = bias_bstride;
  const int bm = bias_mode(s);
  DAT_REQUIRE(bm != BM_DENSE || bias != nullptr, "attention_fwd: dense bias mode without a bias tensor");
  const size_t tsz = a.table_in_smem ? (size_t)((a.Th * a.Tw + 3) & ~3) : 0;
  size_t smem = (tsz + 2 * KV_CHUNK * HC + 2 * KV_CHUNK) * sizeof(float);
  dim3 grid(ceil_div(s.HW, ATT_THREADS), s.B * s.heads);
  DAT_FWD(dispatch(s.act_dtype, bm, [&](auto tv, auto bmc) -> int {
    using T = decltype(tv);
    constexpr int BM = decltype(bmc)::value;
    DAT_FWD(set_smem(attn_fwd_simt_kernel<T, BM>, smem));
    launch_k(attn_fwd_simt_kernel<T, BM>, grid, ATT_THREADS, smem, st, (const T*)q, (const T*)k, (const T*)v, pos, table,
                                                                   (T*)o, lse, a);
    return DAT_OK;
  }));
  DAT_LAUNCH_OK("attn_fwd_simt_kernel");
  return DAT_OK;
}

int rpe_bias(const Shape& s, const float* pos, const float* table, float* bias, cudaStream_t st) {
  AttnArgs a = make_args(s);
  long long total = (long long)s.B * s.heads * s.HW * s.Ns;
  launch_k(rpe_bias_kernel, ceil_div(total, 256), 256, 0, st, pos, table, bias, a, total);
  DAT_LAUNCH_OK("rpe_bias_kernel");
  return DAT_OK;
}

int attention_bwd_qsplit(const Shape& s) {
  long long base = (long long)s.B * s.heads * ceil_div(s.Ns, ATT_THREADS);
  long long want = (2 * 148 + base - 1) / base;
  long long cap = (s.HW + 255) / 256;
  long long r = want < cap ? want : cap;
  return (int)(r < 1 ? 1 : r);
}

// delta[b, eta, m] = sum_c dO * O (shared by the CUDA-core and tensor-core backward paths)
int attention_delta(const Shape& s, const void* d_o, const void* o, float* delta, cudaStream_t st) {
  long long tot = (long long)s.B * s.HW * s.heads;
  if (s.act_dtype == DAT_F32)
    launch_k(attn_delta_kernel<float>, ceil_div(tot, 256), 256, 0, st, (const float*)d_o, (const float*)o, delta, s.HW, s.C, s.heads, tot);
  else
    launch_k(attn_delta_kernel<bf16>, ceil_div(tot, 256), 256, 0, st, (const bf16*)d_o, (const bf16*)o, delta, s.HW, s.C, s.heads, tot);
  DAT_LAUNCH_OK("attn_delta_kernel");
  return DAT_OK;
}

size_t attention_bwd_workspace(const Shape& s) {
  int qs = attention_bwd_qsplit(s);
  size_t delta = align_up((size_t)s.B * s.heads * s.HW * 4, 256);
  size_t part = align_up((size_t)qs * s.B * s.Ns * s.C * 4, 256);
  return delta + 2 * part;
}

// dq, dk, dv (act dtype), d_table (fp32, overwritten; rpe mode only), dpos_part (B,heads,qsplit,Ns,2; zeros
// unless rpe mode).  Dense-bias mode: `bias` as in the forward, `dbias` (B, heads, HW, Ns) receives dS.
int attention_bwd_simt(const Shape& s, const void* q, const void* k, const void* v, const void* o,
                       const void* d_o, const float* lse, const float* pos, const float* table,
                       void* dq, void* dk, void* dv, float* d_table, float* dpos_part,
                       void* ws, size_t ws_bytes, cudaStream_t st, const float* bias, long long bias_bstride,
                       float* dbias) {
  DAT_REQUIRE(ws_bytes >= attention_bwd_workspace(s), "attention_bwd: workspace too small");
  AttnArgs a = make_args(s);
  a.bias = bias; a.bias_bstride = bias_bstride; a.dbias = dbias;
  const int bm = bias_mode(s);
  DAT_REQUIRE(bm != BM_DENSE || (bias != nullptr && dbias != nullptr), "attention_bwd: dense bias mode without bias / dbias");
  DAT_REQUIRE((size_t)a.Th * a.Tw * 4 <= 48 * 1024 + 32 * 1024,
              "attention_bwd: rpe table %dx%d too large for the shared-memory gradient copy", s.Th, s.Tw);
  const int qsplit = attention_bwd_qsplit(s);
  float* delta = (float*)ws;
  float* dk_part = (float*)((char*)ws + align_up((size_t)s.B * s.heads * s.HW * 4, 256));
  float* dv_part = (float*)((char*)dk_part + align_up((size_t)qsplit * s.B * s.Ns * s.C * 4, 256));

  DAT_FWD(attention_delta(s, d_o, o, delta, st));
  if (bm == BM_RPE) DAT_CUDA_OK(cudaMemsetAsync(d_table, 0, (size_t)s.heads * s.Th * s.Tw * 4, st));
  {
    const size_t tsz = (size_t)((a.Th * a.Tw + 3) & ~3);
    size_t smem = (2 * tsz + 2 * KV_CHUNK * HC + 2 * KV_CHUNK) * sizeof(float);
    dim3 grid(ceil_div(s.HW, ATT_THREADS), s.B * s.heads);
    DAT_FWD(dispatch(s.act_dtype, bm, [&](auto tv, auto bmc) -> int {
      using T = decltype(tv);
      constexpr int BM = decltype(bmc)::value;
      DAT_FWD(set_smem(attn_bwd_dq_kernel<T, BM>, smem));
      launch_k(attn_bwd_dq_kernel<T, BM>, grid, ATT_THREADS, smem, st, (const T*)q, (const T*)k, (const T*)v, (const T*)d_o, lse,
                                                                   delta, pos, table, (T*)dq, d_table, a);
      return DAT_OK;
    }));
    DAT_LAUNCH_OK("attn_bwd_dq_kernel");
  }
  {
    const size_t tsz = a.table_in_smem ? (size_t)((a.Th * a.Tw + 3) & ~3) : 0;
    size_t smem = (tsz + 2 * Q_CHUNK * HC + 2 * Q_CHUNK) * sizeof(float);
    dim3 grid(ceil_div(s.Ns, ATT_THREADS), s.B * s.heads, qsplit);
    int qps = ceil_div(ceil_div(s.HW, qsplit), Q_CHUNK) * Q_CHUNK;
    DAT_FWD(dispatch(s.act_dtype, bm, [&](auto tv, auto bmc) -> int {
      using T = decltype(tv);
      constexpr int BM = decltype(bmc)::value;
      DAT_FWD(set_smem(attn_bwd_dkv_kernel<T, BM>, smem));
      launch_k(attn_bwd_dkv_kernel<T, BM>, grid, ATT_THREADS, smem, st, (const T*)q, (const T*)k, (const T*)v, (const T*)d_o, lse,
                                                                    delta, pos, table, dk_part, dv_part, dpos_part, qps, a);
      return DAT_OK;
    }));
    DAT_LAUNCH_OK("attn_bwd_dkv_kernel");
  }
  long long cnt = (long long)s.B * s.Ns * s.C;
  DAT_FWD(reduce_partials(dk_part, qsplit, cnt, dk, s.act_dtype, st));
  DAT_FWD(reduce_partials(dv_part, qsplit, cnt, dv, s.act_dtype, st));
  return DAT_OK;
}

}  // namespace dat
