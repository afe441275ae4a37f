// Vectorised forward of the offset network (same math as offset_net.cu, which remains the
// generic-shape fallback): depthwise k x k strided conv -> LayerNorm -> GELU -> 1x1 (Cg->2)
// -> tanh*range | clamp -> + reference point (dat_blocks.py:51-56, :144-162, :108-121).
//
// One warp per sample point; each lane owns VEC = Cg/32 adjacent channels, so a tap is ONE
// vector load per lane (a whole channel-last row per warp, 128-512 B coalesced) and one
// vector LDS of the transposed weights.  (k/s)^2-fold window overlap is served by L2, HBM sees q once:
// algorithmic bytes = B*HW*C*e + B*G*Ns*(4*Cg + 16).
//
// offset_pos_fwd_unrolled_kernel<TQ, VEC, K> (k = 3, 5, 7, 9: every DAT / DAT++ stage): the first version walked the
// clipped window with a run-time loop, at most 3 loads in flight per warp - 27 dependent round trips to L2 for the
// 9 x 9 window, and ~25 us per launch at EVERY stage (7 - 22 % of the HBM roofline: latency-, not bandwidth-bound).
// Here a lane owns 8 channels (one 16-byte load per tap), so a warp works on 256 / Cg points at once (4 at Cg = 64);
// the taps of RB window rows are issued back to back as predicated loads (zero outside the image, which is what the
// zero padding contributes) before the first FMA, CTAs are persistent (a warp walks over point groups, the
// transposed filter is staged once per CTA with coalesced reads) and t_dw leaves as vector stores.
// Per channel the fp32 summation order (u, then v) is unchanged; LayerNorm / 1x1 sums are re-associated (8 channels
// per lane, then a butterfly over the point's lanes).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int WARPS = 8;

struct VArgs {
  int H, W, C, G, Cg, stride, ksize, pad, Hk, Wk, Ns;
  float orf, range_y, range_x;
  long long n_points;
};

template <int VEC> struct VecLoad;
template <> struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float* o) { o[0] = *p; }
  static __device__ __forceinline__ void ld(const bf16* p, float* o) { o[0] = __bfloat162float(*p); }
};
template <> struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float2 v = *reinterpret_cast<const float2*>(p); o[0] = v.x; o[1] = v.y;
  }
  static __device__ __forceinline__ void ld(const bf16* p, float* o) {
    float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); o[0] = v.x; o[1] = v.y;
  }
};
template <> struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void ld(const bf16* p, float* o) {
    float4 v = load4(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct VecLoad<8> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    VecLoad<4>::ld(p, o); VecLoad<4>::ld(p + 4, o + 4);
  }
  static __device__ __forceinline__ void ld(const bf16* p, float* o) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};

__device__ __forceinline__ float gelu_exact(float z) {
  return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
}

template <typename TQ, int VEC>
__global__ void __launch_bounds__(WARPS * 32)
offset_pos_fwd_vec_kernel(const TQ* __restrict__ q, const float* __restrict__ w_dw,
                          const float* __restrict__ b_dw, const float* __restrict__ ln_g,
                          const float* __restrict__ ln_b, const float* __restrict__ w_pw,
                          float* __restrict__ t_dw, float* __restrict__ off_raw,
                          float* __restrict__ pos, VArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) float wsm[];   // [k*k][Cg]
  const int kk = a.ksize * a.ksize;
  for (int idx = threadIdx.x; idx < kk * a.Cg; idx += blockDim.x) {
    int c = idx % a.Cg, uv = idx / a.Cg;
    wsm[idx] = w_dw[c * kk + uv];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long sp = (long long)blockIdx.x * WARPS + warp;
  if (sp >= a.n_points) return;
  const int n = (int)(sp % a.Ns);
  const int g = (int)((sp / a.Ns) % a.G);
  const int b = (int)(sp / ((long long)a.Ns * a.G));
  const int i = n / a.Wk, j = n - i * a.Wk;
  const int c0 = lane * VEC;

  float acc[VEC];
  VecLoad<VEC>::ld(b_dw + c0, acc);
  const int yb = i * a.stride - a.pad, xb = j * a.stride - a.pad;
  const int u_lo = max(0, -yb), u_hi = min(a.ksize, a.H - yb);
  const int v_lo = max(0, -xb), v_hi = min(a.ksize, a.W - xb);
  const TQ* prow = q + (((long long)b * a.H + yb + u_lo) * a.W + (xb + v_lo)) * a.C + g * a.Cg + c0;
  const long long row_step = (long long)a.W * a.C;
  for (int u = u_lo; u < u_hi; ++u) {
    const TQ* p = prow;
    const float* w = wsm + (u * a.ksize + v_lo) * a.Cg + c0;
#pragma unroll 3
    for (int v = v_lo; v < v_hi; ++v) {
      float qv[VEC], wv[VEC];
      VecLoad<VEC>::ld(p, qv);
      VecLoad<VEC>::ld(w, wv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] = fmaf(wv[e], qv[e], acc[e]);
      p += a.C;
      w += a.Cg;
    }
    prow += row_step;
  }
  float s1 = 0.f;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    t_dw[sp * a.Cg + c0 + e] = acc[e];
    s1 += acc[e];
  }
  const float inv_n = 1.0f / (float)a.Cg;
  const float mean = warp_sum(s1) * inv_n;
  float s2 = 0.f;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    float d = acc[e] - mean;
    s2 = fmaf(d, d, s2);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(s2) * inv_n + 1e-5f);
  float gam[VEC], bet[VEC], wy[VEC], wx[VEC];
  VecLoad<VEC>::ld(ln_g + c0, gam);
  VecLoad<VEC>::ld(ln_b + c0, bet);
  VecLoad<VEC>::ld(w_pw + c0, wy);
  VecLoad<VEC>::ld(w_pw + a.Cg + c0, wx);
  float oy = 0.f, ox = 0.f;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    float act = gelu_exact((acc[e] - mean) * rstd * gam[e] + bet[e]);
    oy = fmaf(act, wy[e], oy);
    ox = fmaf(act, wx[e], ox);
  }
  oy = warp_sum(oy);
  ox = warp_sum(ox);
  if (lane == 0) {
    off_raw[sp * 2 + 0] = oy;
    off_raw[sp * 2 + 1] = ox;
    const float ry = ref_point(i, a.Hk), rx = ref_point(j, a.Wk);
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

template <typename T, int VEC> struct RawVec;
template <int VEC> struct RawVec<float, VEC> {
  float v[VEC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = 0.f;
  }
  __device__ __forceinline__ void load(const float* p) { VecLoad<VEC>::ld(p, v); }
  __device__ __forceinline__ void get(float* o) const {
#pragma unroll
    for (int e = 0; e < VEC; ++e) o[e] = v[e];
  }
};
template <> struct RawVec<bf16, 1> {
  unsigned short r;
  __device__ __forceinline__ void zero() { r = 0; }
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const unsigned short*>(p); }
  __device__ __forceinline__ void get(float* o) const { o[0] = __uint_as_float((uint32_t)r << 16); }
};
template <> struct RawVec<bf16, 2> {
  uint32_t r;
  __device__ __forceinline__ void zero() { r = 0u; }
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint32_t*>(p); }
  __device__ __forceinline__ void get(float* o) const {
    o[0] = __uint_as_float(r << 16);
    o[1] = __uint_as_float(r & 0xffff0000u);
  }
};
template <> struct RawVec<bf16, 4> {
  uint2 r;
  __device__ __forceinline__ void zero() { r = make_uint2(0u, 0u); }
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void get(float* o) const {
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct RawVec<bf16, 8> {
  uint4 r;
  __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float* o) const {
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
    o[4] = __uint_as_float(r.z << 16); o[5] = __uint_as_float(r.z & 0xffff0000u);
    o[6] = __uint_as_float(r.w << 16); o[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};

template <int VEC> __device__ __forceinline__ void store_vec(float* p, const float* v) {
  if (VEC == 1) p[0] = v[0];
  else if (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  else {
#pragma unroll
    for (int e = 0; e < VEC; e += 4) *reinterpret_cast<float4*>(p + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
  }
}

// LPP lanes per point, each owning 8 adjacent channels (Cg = 8 * LPP): a warp works on 32 / LPP points at once
template <typename TQ, int LPP, int K>
__global__ void __launch_bounds__(WARPS * 32, sizeof(TQ) == 2 ? 3 : 2)
offset_pos_fwd_unrolled_kernel(const TQ* __restrict__ q, const float* __restrict__ w_dw,
                               const float* __restrict__ b_dw, const float* __restrict__ ln_g,
                               const float* __restrict__ ln_b, const float* __restrict__ w_pw,
                               float* __restrict__ t_dw, float* __restrict__ off_raw,
                               float* __restrict__ pos, VArgs a) {
  pdl_enter();
  constexpr int VEC = 8, PPW = 32 / LPP, KK = K * K;
  // raw taps in flight: rows per batch so that a batch holds at most 40 (bf16: 3 CTAs per SM) / 72 (fp32: 2) registers
  constexpr int REGS_PER_TAP = (int)(sizeof(TQ) * VEC) / 4;
  constexpr int RB_A = (sizeof(TQ) == 2 ? 40 : 72) / (K * REGS_PER_TAP), RB = RB_A < 1 ? 1 : (RB_A > K ? K : RB_A);
  extern __shared__ __align__(16) float wsm[];   // [K*K][Cg + 4]: transposed filter, pitch keeps 16-byte alignment
  const int pitch = a.Cg + 4;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < KK * a.Cg; idx += WARPS * 32) {   // coalesced read of (Cg, K*K)
    const int c = idx / KK, uv = idx - c * KK;
    wsm[uv * pitch + c] = w_dw[idx];
  }
  // per-channel parameters [bias | ln gamma | ln beta | 1x1 row y | 1x1 row x] behind the filter: fetched together with it
  // (one memory latency for everything), read back from shared memory where they are used (40 registers otherwise)
  float* spar = wsm + KK * pitch;
  for (int idx = threadIdx.x; idx < 5 * a.Cg; idx += WARPS * 32) {
    const int w = idx / a.Cg, c = idx - w * a.Cg;
    spar[idx] = w == 0 ? b_dw[c] : w == 1 ? ln_g[c] : w == 2 ? ln_b[c] : w_pw[(w - 3) * a.Cg + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPP, c0 = (lane % LPP) * VEC;
  const float inv_n = 1.0f / (float)a.Cg;
  const long long row_step = (long long)a.W * a.C;
  const int n_points = (int)a.n_points;
  const int n_items = (n_points + PPW - 1) / PPW;
  for (int item = blockIdx.x * WARPS + warp; item < n_items; item += gridDim.x * WARPS) {
    const int sp_raw = item * PPW + sub;
    const bool valid = sp_raw < n_points;
    const int sp = valid ? sp_raw : n_points - 1;       // idle sub-groups shadow the last point (shuffles stay uniform)
    const int n = sp % a.Ns;
    const int bg = sp / a.Ns, g = bg % a.G, b = bg / a.G;
    const int i = n / a.Wk, j = n - i * a.Wk;
    const int yb = i * a.stride - a.pad, xb = j * a.stride - a.pad;
    // the window's top-left tap may lie outside the image: the pointer is only dereferenced under the predicates
    const TQ* p0 = q + (((long long)b * a.H + yb) * a.W + xb) * a.C + g * a.Cg + c0;
    bool xok[K];
#pragma unroll
    for (int v = 0; v < K; ++v) xok[v] = (unsigned)(xb + v) < (unsigned)a.W;
    float acc[VEC];
    VecLoad<VEC>::ld(spar + c0, acc);
#pragma unroll 1
    for (int u0 = 0; u0 < K; u0 += RB) {     // not unrolled: the compiler would hoist all K * K loads (register spills)
      RawVec<TQ, VEC> raw[RB][K];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const bool yok = u0 + r < K && (unsigned)(yb + u0 + r) < (unsigned)a.H;
        const TQ* prow = p0 + (long long)(u0 + r) * row_step;
#pragma unroll
        for (int v = 0; v < K; ++v) {
          if (yok && xok[v]) raw[r][v].load(prow + v * a.C);
          else raw[r][v].zero();
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (u0 + r < K) {
#pragma unroll
          for (int v = 0; v < K; ++v) {
            float qv[VEC], wv[VEC];
            raw[r][v].get(qv);
            VecLoad<VEC>::ld(wsm + ((u0 + r) * K + v) * pitch + c0, wv);
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = fmaf(wv[e], qv[e], acc[e]);
          }
        }
      }
    }
    if (valid) store_vec<VEC>(t_dw + (long long)sp * a.Cg + c0, acc);
    float s1 = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) s1 += acc[e];
#pragma unroll
    for (int o = LPP / 2; o >= 1; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    const float mean = s1 * inv_n;
    float s2 = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float d = acc[e] - mean;
      s2 = fmaf(d, d, s2);
    }
#pragma unroll
    for (int o = LPP / 2; o >= 1; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float rstd = 1.0f / sqrtf(s2 * inv_n + 1e-5f);
    float oy = 0.f, ox = 0.f;
    float gam[VEC], bet[VEC], wy[VEC], wx[VEC];
    VecLoad<VEC>::ld(spar + a.Cg + c0, gam);
    VecLoad<VEC>::ld(spar + 2 * a.Cg + c0, bet);
    VecLoad<VEC>::ld(spar + 3 * a.Cg + c0, wy);
    VecLoad<VEC>::ld(spar + 4 * a.Cg + c0, wx);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float act = gelu_exact((acc[e] - mean) * rstd * gam[e] + bet[e]);
      oy = fmaf(act, wy[e], oy);
      ox = fmaf(act, wx[e], ox);
    }
#pragma unroll
    for (int o = LPP / 2; o >= 1; o >>= 1) {
      oy += __shfl_xor_sync(0xffffffffu, oy, o);
      ox += __shfl_xor_sync(0xffffffffu, ox, o);
    }
    if (valid && (lane % LPP) == 0) {
      const float ry = ref_point(i, a.Hk), rx = ref_point(j, a.Wk);
      float py, px;
      if (a.orf >= 0.f) {
        py = __fadd_rn(__fmul_rn(__fmul_rn(tanhf(oy), a.range_y), a.orf), ry);
        px = __fadd_rn(__fmul_rn(__fmul_rn(tanhf(ox), a.range_x), a.orf), rx);
      } else {
        py = fminf(fmaxf(__fadd_rn(oy, ry), -1.0f), 1.0f);
        px = fminf(fmaxf(__fadd_rn(ox, rx), -1.0f), 1.0f);
      }
      *reinterpret_cast<float2*>(off_raw + (long long)sp * 2) = make_float2(oy, ox);
      *reinterpret_cast<float2*>(pos + (long long)sp * 2) = make_float2(py, px);
    }
  }
}

}  // namespace

bool offset_pos_fwd_vec_supported(const Shape& s) {
  return s.Cg == 32 || s.Cg == 64 || s.Cg == 128 || s.Cg == 256;
}

int offset_pos_fwd_vec(const Shape& s, const dat_block_params* p, const void* q, float* t_dw,
                       float* off_raw, float* pos, cudaStream_t st) {
  VArgs a;
  a.H = s.H; a.W = s.W; a.C = s.C; a.G = s.G; a.Cg = s.Cg; a.stride = s.stride; a.ksize = s.ksize;
  a.pad = s.pad; a.Hk = s.Hk; a.Wk = s.Wk; a.Ns = s.Ns; a.orf = s.orf;
  a.range_y = (float)(1.0 / ((double)s.Hk - 1.0));   // Python double rounded to fp32, dat_blocks.py:150
  a.range_x = (float)(1.0 / ((double)s.Wk - 1.0));
  a.n_points = (long long)s.B * s.G * s.Ns;
  const bool unrolled = (s.ksize == 3 || s.ksize == 5 || s.ksize == 7 || s.ksize == 9) &&
                        std::getenv("DAT_B200_OFFSET_FWD_V1") == nullptr;
  const size_t smem = ((size_t)s.ksize * s.ksize * (s.Cg + (unrolled ? 4 : 0)) + (unrolled ? 5 * s.Cg : 0)) * sizeof(float);
  DAT_REQUIRE(smem <= 200 * 1024, "offset net: k*k*Cg too large for shared memory");
  const int lpp = s.Cg / 8;                                 // unrolled kernel: 8 channels per lane, 32 / lpp points per warp
  const int grid_u = ceil_div(ceil_div(a.n_points, 32 / (lpp > 0 ? lpp : 1)), WARPS);
  const int grid = unrolled ? (grid_u < 148 * 3 ? grid_u : 148 * 3) : ceil_div(a.n_points, WARPS);   // persistent: 3 CTAs / SM
  DAT_REQUIRE(!unrolled || a.n_points < (1ll << 30), "offset net: too many sample points");
#define LAUNCH_K(TQ, L, KV)                                                                    \
  do {                                                                                         \
    auto kern = offset_pos_fwd_unrolled_kernel<TQ, L, KV>;                                     \
    if (smem > 48 * 1024)                                                                      \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(kern, grid, WARPS * 32, smem, st, (const TQ*)q, p->off_dw_w, p->off_dw_b, p->off_ln_g,  \
                                         p->off_ln_b, p->off_pw_w, t_dw, off_raw, pos, a);     \
  } while (0)
#define LAUNCH_KK(TQ, L)                                                                       \
  do {                                                                                         \
    if (s.ksize == 3) LAUNCH_K(TQ, L, 3); else if (s.ksize == 5) LAUNCH_K(TQ, L, 5);           \
    else if (s.ksize == 7) LAUNCH_K(TQ, L, 7); else LAUNCH_K(TQ, L, 9);                        \
  } while (0)
#define LAUNCH(TQ, V)                                                                          \
  do {                                                                                         \
    if (unrolled) {                                                                            \
      if (lpp == 4) LAUNCH_KK(TQ, 4); else if (lpp == 8) LAUNCH_KK(TQ, 8);                     \
      else if (lpp == 16) LAUNCH_KK(TQ, 16); else LAUNCH_KK(TQ, 32);                           \
      break;                                                                                   \
    }                                                                                          \
    auto kern = offset_pos_fwd_vec_kernel<TQ, V>;                                              \
    if (smem > 48 * 1024)                                                                      \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(kern, grid, WARPS * 32, smem, st, (const TQ*)q, p->off_dw_w, p->off_dw_b, p->off_ln_g,  \
                                         p->off_ln_b, p->off_pw_w, t_dw, off_raw, pos, a);     \
  } while (0)
#define LAUNCH_T(TQ)                          \
  do {                                        \
    if (s.Cg == 32) LAUNCH(TQ, 1);            \
    else if (s.Cg == 64) LAUNCH(TQ, 2);       \
    else if (s.Cg == 128) LAUNCH(TQ, 4);      \
    else LAUNCH(TQ, 8);                       \
  } while (0)
  if (s.act_dtype == DAT_F32) LAUNCH_T(float); else LAUNCH_T(bf16);
#undef LAUNCH_T
#undef LAUNCH
#undef LAUNCH_K
#undef LAUNCH_KK
  DAT_LAUNCH_OK("offset_pos_fwd_vec_kernel");
  return DAT_OK;
}

}  // namespace dat
