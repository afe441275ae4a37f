// Residual add with stochastic depth, one pass:  y[b, ...] = x[b, ...] + a[b, ...] * s[b]
// (TransformerStage._inner_forward, dat.py:147-156: `x = drop_path(branch) + x`; timm drop_path:
// per-sample mask / keep_prob).  x may be absent ('X' blocks have no residual around the MLP);
// the same kernel forms the branch gradient da = dy * s[b].  a, x, y share one dense layout whose
// outermost dimension is the sample, so the sample index is flat_index / per_sample.
// HBM-bound streaming kernel, 4 elements per thread; algorithmic bytes n * (e_a + e_x + e_y).
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

template <typename TA, typename TX, typename TY, bool HAS_X>
__global__ void __launch_bounds__(256)
scale_residual_kernel(const TA* __restrict__ a, const TX* __restrict__ x, const float* __restrict__ s,
                      TY* __restrict__ y, long long n4, long long per_sample4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float sc = s[i / per_sample4];
  const float4 av = load4(a + 4 * i);
  float4 o = make_float4(av.x * sc, av.y * sc, av.z * sc, av.w * sc);
  if (HAS_X) {
    const float4 xv = load4(x + 4 * i);
    o.x += xv.x; o.y += xv.y; o.z += xv.z; o.w += xv.w;
  }
  store4(y + 4 * i, o);
}

}  // namespace

int scale_residual(const void* a, int a_dt, const void* x, int x_dt, const float* s, void* y, int y_dt,
                   long long B, long long per_sample, cudaStream_t st) {
  DAT_REQUIRE(per_sample % 4 == 0, "scale_residual: elements per sample must be a multiple of 4");
  const long long n4 = B * per_sample / 4, ps4 = per_sample / 4;
  if (n4 == 0) return DAT_OK;
  const unsigned grid = (unsigned)ceil_div(n4, 256ll);
#define LAUNCH(TA, TX, TY, HX) \
  launch_k(scale_residual_kernel<TA, TX, TY, HX>, grid, 256, 0, st, (const TA*)a, (const TX*)x, s, (TY*)y, n4, ps4)
#define LAUNCH_Y(TA, TX, HX)                                                       \
  do {                                                                             \
    if (y_dt == DAT_F32) LAUNCH(TA, TX, float, HX); else LAUNCH(TA, TX, bf16, HX); \
  } while (0)
#define LAUNCH_X(TA)                                            \
  do {                                                          \
    if (x == nullptr) LAUNCH_Y(TA, float, false);               \
    else if (x_dt == DAT_F32) LAUNCH_Y(TA, float, true);        \
    else LAUNCH_Y(TA, bf16, true);                              \
  } while (0)
  if (a_dt == DAT_F32) LAUNCH_X(float); else LAUNCH_X(bf16);
#undef LAUNCH_X
#undef LAUNCH_Y
#undef LAUNCH
  DAT_LAUNCH_OK("scale_residual_kernel");
  return DAT_OK;
}

}  // namespace dat
