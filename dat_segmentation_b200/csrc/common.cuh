// Shared device/host helpers for the dat_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "../../include/dat_b200.h"

namespace dat {

// ---- error plumbing --------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();  // host-side counter behind dat_launch_count()

#define DAT_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::dat::set_error(__VA_ARGS__);           \
      return DAT_ERR_ARG;                      \
    }                                          \
  } while (0)

#define DAT_CUDA_OK(expr)                                                             \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::dat::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                       __FILE__, __LINE__);                                           \
      return DAT_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define DAT_LAUNCH_OK(name)                                                           \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::dat::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));     \
      return DAT_ERR_CUDA;                                                            \
    }                                                                                 \
    ::dat::count_launch();                                                            \
  } while (0)

#define DAT_FWD(expr)             \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != DAT_OK) return rc__; \
  } while (0)

// ---- derived shape ---------------------------------------------------------------
struct Shape {
  int B, H, W, HW, C, heads, G, Cg, hg, stride, ksize, pad, Hk, Wk, Ns, Th, Tw;
  float orf;
  int x_dtype, act_dtype;
  int pe_mode, no_off;       // DAT_PE_*; no_off forces DAT_PE_NONE and the avg-pool sample grid
};

// Validates a descriptor and fills the derived sizes; DAT_OK or DAT_ERR_ARG.
int make_shape(const dat_block_desc* d, Shape* s);

static inline size_t dtype_size(int dt) { return dt == DAT_BF16 ? 2 : 4; }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device helpers --------------------------------------------------------------
#ifdef __CUDACC__

// ---- programmatic dependent launch ---------------------------------------------------------
// Every kernel of the library starts with pdl_enter() and is launched through launch_k(): the launch carries the
// programmatic-stream-serialization attribute, so the grid may be scheduled while the previous kernel of the stream is
// still draining (its CTAs have all started and signalled `launch_dependents`); `griddepcontrol.wait` then blocks
// until that kernel has completed and its writes are visible.  Nothing is read or written before the wait, so the
// semantics are those of a plain in-order stream - only launch latency, CTA scheduling and the flush at the kernel
// boundary overlap the previous kernel's tail.  Captured into CUDA graphs as programmatic edges.
// No early `griddepcontrol.launch_dependents`: with it the next grid becomes resident (and holds registers / shared
// memory / tensor memory) while the current one still runs; measured on the training step that costs 9 % (816 vs
// 894 images/s, also with the side-branch streams excluded, 772 with a single stream; triggering AFTER the wait, which
// bounds the look-ahead to one grid, costs the same: 815 vs 877) where the implicit trigger at CTA exit gains 1.1 % over
// plain launches.
// DAT_B200_PDL=0 launches without the attribute (the wait is then a no-op).
__device__ __forceinline__ void pdl_enter() {
#ifdef DAT_PDL_EARLY_TRIGGER     // measured on B200: 816 vs 894 images/s - see the note above
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef DAT_PDL_TRIGGER_AFTER_WAIT
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);   // errors surface in DAT_LAUNCH_OK
}
// the same with a run-time thread-block cluster shape
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster.x;
  at[0].val.clusterDim.y = cluster.y;
  at[0].val.clusterDim.z = cluster.z;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements ↔ float4 (16-byte access for fp32, 8-byte for bf16).
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Grid coordinate → source index, align_corners=True:  ((g + 1) / 2) * (size - 1).
// Explicit round-to-nearest intrinsics: never contracted into an FMA, so the integer
// taps are bit-identical to ATen's (torch/include/ATen/native/GridSampler.h:27-36).
__device__ __forceinline__ float unnormalize(float g, int size) {
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

// Reference sample centre ((i + 0.5) / (n - 1)) * 2 - 1, every op rounded separately
// (dat_blocks.py:111-118).  True IEEE division.
__device__ __forceinline__ float ref_point(int i, int n) {
  return __fsub_rn(__fmul_rn(__fdiv_rn((float)i + 0.5f, (float)n - 1.0f), 2.0f), 1.0f);
}
// Query grid (r / (n - 1)) * 2 - 1 (dat_blocks.py:126-133).
__device__ __forceinline__ float query_point(int r, int n) {
  return __fsub_rn(__fmul_rn(__fdiv_rn((float)r, (float)n - 1.0f), 2.0f), 1.0f);
}

// One bilinear footprint: north-west tap, the four weights and validity bits
// (zeros padding).  Weight formulas follow ATen's CUDA kernel:
// nw = (x1 - ix)(y1 - iy), ne = (ix - x0)(y1 - iy), sw = (x1 - ix)(iy - y0), se = ...
struct Taps {
  int x0, y0;
  float wx0, wx1, wy0, wy1;  // wx0 = x1 - ix, wx1 = ix - x0
  bool vx0, vx1, vy0, vy1;
};
__device__ __forceinline__ Taps make_taps(float gx, float gy, int W, int H) {
  Taps t;
  float ix = unnormalize(gx, W), iy = unnormalize(gy, H);
  float fx = floorf(ix), fy = floorf(iy);
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  t.wx1 = __fsub_rn(ix, fx);
  t.wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);
  t.wy1 = __fsub_rn(iy, fy);
  t.wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
  t.vx0 = t.x0 >= 0 && t.x0 < W;
  t.vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.vy0 = t.y0 >= 0 && t.y0 < H;
  t.vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  return t;
}

#endif  // __CUDACC__

}  // namespace dat
