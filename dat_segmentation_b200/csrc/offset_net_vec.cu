// Vectorised forward of the offset network (same math as offset_net.cu, which remains the
// generic-shape fallback): depthwise k x k strided conv -> LayerNorm -> GELU -> 1x1 (Cg->2)
// -> tanh*range | clamp -> + reference point (dat_blocks.py:51-56, :144-162, :108-121).
//
// One warp per sample point; each lane owns VEC = Cg/32 adjacent channels, so a tap is ONE
// vector load per lane (a whole channel-last row per warp, 128-512 B coalesced) and one
// vector LDS of the transposed weights.  The valid tap window is clipped once per point
// (no per-tap bounds checks) and pointers advance by constant strides, which cuts the
// instruction count ~6x against the scalar kernel; the kernel is then bound by how fast
// L2 can stream q ((k/s)^2-fold window overlap is served by L2, HBM sees q once:
// algorithmic bytes = B*HW*C*e + B*G*Ns*(4*Cg + 16)).
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
  const size_t smem = (size_t)s.ksize * s.ksize * s.Cg * sizeof(float);
  DAT_REQUIRE(smem <= 200 * 1024, "offset net: k*k*Cg too large for shared memory");
  const int grid = ceil_div(a.n_points, WARPS);
#define LAUNCH(TQ, V)                                                                          \
  do {                                                                                         \
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
  DAT_LAUNCH_OK("offset_pos_fwd_vec_kernel");
  return DAT_OK;
}

}  // namespace dat
