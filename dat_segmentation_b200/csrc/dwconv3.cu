// Channel-last depthwise 3 x 3 convolution (stride 1, padding 1) as register sliding windows:
// the local perception unit (dat.py:135-138) and the MLP middle (dat_blocks.py:338-343), forward
// and a single fused backward.
//
// A thread owns 2 consecutive channels of a strip of TW output columns and walks down `th` rows
// keeping rows y-1, y, y+1 of the (TW + 2)-column window in registers (statically rotated), so
// every input element is loaded about once per strip instead of once per tap; the 9 x 2 filter
// taps live in registers.  Consecutive lanes hold consecutive channel pairs: a warp's load of
// one pixel is one contiguous 128-byte (bf16) / 256-byte (fp32) segment of the channel-last row.
// Loads for row y + 2 are issued before row y is computed (software prefetch in registers).
//
//   forward   mode 0: y = conv(x) + b    mode 1: y = conv(x) + b + x
//             mode 2: z = conv(x) + b + x (stored), y = gelu(z)
//   backward  one kernel: dz = dy * gelu'(z) (mode 2) is formed on the fly while the window is
//             filled, then for every pixel q and tap (u, v), t = dz[q - (u-1, v-1)]:
//                 dx[q] += w[u][v] * t  (+ dz[q] for the residual modes)
//                 dw[u][v] += x[q] * t,  db += dz[q]
//             so dy, z, x are read once and dx written once (4 tensor passes instead of the 8 of
//             gelu_bwd + data gradient + weight gradient).  Per-strip partial dw/db are reduced in
//             a fixed order by dwconv_wgrad_reduce_kernel (deterministic, no atomics).
// GELU and its derivative use erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), which shares
// exp(-z^2/2) between the cdf and the pdf (13-14 instructions per element).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int TW = 4;          // output columns per strip
constexpr int D3_THREADS = 128;

template <typename T> struct Raw2;
template <> struct Raw2<float> {
  using type = float2;
  static __device__ __forceinline__ type zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ type load(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ float2 cvt(type r) { return r; }
  static __device__ __forceinline__ void store(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <> struct Raw2<bf16> {
  using type = uint32_t;
  static __device__ __forceinline__ type zero() { return 0u; }
  static __device__ __forceinline__ type load(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }
  static __device__ __forceinline__ float2 cvt(type r) {
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
  static __device__ __forceinline__ void store(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

// GELU pieces from erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7):
//   r(z) = 1/2 - Phi(-|z|) = 1/2 - (1/2) poly(t) exp(-z^2/2),  t = 1 / (1 + p |z| / sqrt 2)
//   Phi(z) = 1/2 + sgn(z) r,   gelu(z) = z / 2 + |z| r,   gelu'(z) = Phi(z) + z exp(-z^2/2) / sqrt(2 pi)
// MUFU.RCP / MUFU.EX2 in their flush-to-zero forms (no denormal fix-up code around them).
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_r(float z, float& e) {
  const float t = rcp_ftz(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(z), 1.0f));
  e = ex2_ftz(z * z * -0.72134752044448170368f);          // exp(-z^2 / 2)
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  return fmaf(-(p * t), e, 0.5f);
}
__device__ __forceinline__ float gelu_fast(float z) {
  float e;
  const float r = gelu_r(z, e);
  return fmaf(fabsf(z), r, 0.5f * z);
}
__device__ __forceinline__ float gelu_grad_fast(float z) {
  float e;
  const float r = gelu_r(z, e);
  return fmaf(z * 0.39894228040143267794f, e, copysignf(r, z)) + 0.5f;
}

struct Strip {
  int b, x0, y0, c;
  bool ok;
};
__device__ __forceinline__ Strip strip_of(int B, int C, int th, int strips_x, int strips_y) {
  const int nvec = C >> 1;
  const long long t = (long long)blockIdx.x * D3_THREADS + threadIdx.x;
  const long long strip = t / nvec;
  Strip s;
  s.c = (int)(t - strip * nvec) * 2;
  s.ok = strip < (long long)B * strips_x * strips_y;
  const int sx = (int)(strip % strips_x);
  const long long r = strip / strips_x;
  s.x0 = sx * TW;
  s.y0 = (int)(r % strips_y) * th;
  s.b = (int)(r / strips_y);
  return s;
}

// raw loads of one window row (columns x0-1 .. x0+TW); rp points at column x0 of that row, zeros
// outside the image.  ALIGNED (W % TW == 0): only the two halo columns need a column predicate,
// and those predicates (lval / rval) are per-thread constants.
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_row(const T* __restrict__ rp, bool row_ok, bool lval, bool rval, int C,
                                         int x0, int W, typename Raw2<T>::type (&raw)[TW + 2]) {
  if (ALIGNED) {
    raw[0] = (row_ok && lval) ? Raw2<T>::load(rp - C) : Raw2<T>::zero();
#pragma unroll
    for (int j = 1; j <= TW; ++j) raw[j] = row_ok ? Raw2<T>::load(rp + (j - 1) * C) : Raw2<T>::zero();
    raw[TW + 1] = (row_ok && rval) ? Raw2<T>::load(rp + TW * C) : Raw2<T>::zero();
  } else {
#pragma unroll
    for (int j = 0; j < TW + 2; ++j) {
      const int xx = x0 - 1 + j;
      raw[j] = (row_ok && xx >= 0 && xx < W) ? Raw2<T>::load(rp + (j - 1) * C) : Raw2<T>::zero();
    }
  }
}

template <typename TI, typename TO, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(D3_THREADS, 5)
dwconv3_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   TO* __restrict__ y, TO* __restrict__ z_out, int B, int H, int W, int C, int th,
                   int strips_x, int strips_y, int flip) {
  pdl_enter();
  using RI = Raw2<TI>;
  const Strip s = strip_of(B, C, th, strips_x, strips_y);
  if (!s.ok) return;
  float wr[9][2];
#pragma unroll
  for (int uv = 0; uv < 9; ++uv) {
    wr[uv][0] = w[s.c * 9 + (flip ? 8 - uv : uv)];
    wr[uv][1] = w[(s.c + 1) * 9 + (flip ? 8 - uv : uv)];
  }
  const float b0 = bias != nullptr ? bias[s.c] : 0.f, b1 = bias != nullptr ? bias[s.c + 1] : 0.f;
  const bool lval = s.x0 > 0, rval = s.x0 + TW < W;
  const int rstride = W * C;                               // < 2^31 elements per image
  const long long pix0 = (long long)s.b * H * rstride + (long long)s.y0 * rstride + s.x0 * C + s.c;
  const TI* rp = x + pix0 - rstride;                       // row being loaded: starts at y0 - 1
  TO* yp = y + pix0;
  TO* zp = MODE == 2 ? z_out + pix0 : nullptr;
  float win[3][TW + 2][2];
  typename RI::type raw[TW + 2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {            // rows y0-1, y0 -> slots 0, 1
    load_row<TI, ALIGNED>(rp, s.y0 - 1 + r >= 0, lval, rval, C, s.x0, W, raw);
    rp += rstride;
#pragma unroll
    for (int j = 0; j < TW + 2; ++j) { const float2 f = RI::cvt(raw[j]); win[r][j][0] = f.x; win[r][j][1] = f.y; }
  }
  load_row<TI, ALIGNED>(rp, s.y0 + 1 < H, lval, rval, C, s.x0, W, raw);      // row y0+1, converted in the loop
  rp += rstride;
  const int y_end = min(s.y0 + th, H);
  for (int yb = s.y0; yb < y_end; yb += 3) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int yy = yb + k;
      if (yy < y_end) {
        float (&prev)[TW + 2][2] = win[k];
        float (&cur)[TW + 2][2] = win[(k + 1) % 3];
        float (&next)[TW + 2][2] = win[(k + 2) % 3];
#pragma unroll
        for (int j = 0; j < TW + 2; ++j) { const float2 f = RI::cvt(raw[j]); next[j][0] = f.x; next[j][1] = f.y; }
        load_row<TI, ALIGNED>(rp, yy + 2 < H, lval, rval, C, s.x0, W, raw);   // prefetch for the next iteration
        rp += rstride;
#pragma unroll
        for (int i = 0; i < TW; ++i) {
          if (ALIGNED || s.x0 + i < W) {
            float a0 = b0, a1 = b1;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
              a0 = fmaf(wr[v][0], prev[i + v][0], a0);      a1 = fmaf(wr[v][1], prev[i + v][1], a1);
              a0 = fmaf(wr[3 + v][0], cur[i + v][0], a0);   a1 = fmaf(wr[3 + v][1], cur[i + v][1], a1);
              a0 = fmaf(wr[6 + v][0], next[i + v][0], a0);  a1 = fmaf(wr[6 + v][1], next[i + v][1], a1);
            }
            if (MODE >= 1) { a0 += cur[i + 1][0]; a1 += cur[i + 1][1]; }
            if (MODE == 2) {
              Raw2<TO>::store(zp + i * C, a0, a1);
              a0 = gelu_fast(a0);
              a1 = gelu_fast(a1);
            }
            Raw2<TO>::store(yp + i * C, a0, a1);
          }
        }
        yp += rstride;
        if (MODE == 2) zp += rstride;
      }
    }
  }
}

// fused backward; partial[strip or CTA][10][C] (rows 0-8 = dw taps, row 9 = db)
template <typename TX, typename TD, int MODE, bool ALIGNED>
__global__ void __launch_bounds__(D3_THREADS, 4)
dwconv3_bwd_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, const TD* __restrict__ z,
                   const float* __restrict__ w, TX* __restrict__ dx, float* __restrict__ partial, int B,
                   int H, int W, int C, int th, int strips_x, int strips_y, int spc) {
  pdl_enter();
  using RD = Raw2<TD>;
  using RX = Raw2<TX>;
  const Strip s = strip_of(B, C, th, strips_x, strips_y);
  float wr[9][2], dwa[10][2];
#pragma unroll
  for (int uv = 0; uv < 9; ++uv) {
    wr[uv][0] = w[s.c * 9 + uv];
    wr[uv][1] = w[(s.c + 1) * 9 + uv];
  }
#pragma unroll
  for (int uv = 0; uv < 10; ++uv) dwa[uv][0] = dwa[uv][1] = 0.f;
  if (!s.ok) H = 0;                       // every load of an idle thread is masked off
  const bool lval = s.x0 > 0, rval = s.x0 + TW < W;
  const int rstride = W * C;
  const long long pix0 = s.ok ? (long long)s.b * H * rstride + (long long)s.y0 * rstride + s.x0 * C + s.c : 0;
  const TD* dp = dy + pix0 - rstride;     // window row being loaded: starts at y0 - 1
  const TD* zp = MODE == 2 ? z + pix0 - rstride : nullptr;
  const TX* xp = x + pix0;                // centre row being loaded: starts at y0
  TX* dxp = dx + pix0;
  float win[3][TW + 2][2];
  typename RD::type rawd[TW + 2], rawz[TW + 2];
  typename RX::type rawx[TW];
  auto fill = [&](float (&row)[TW + 2][2]) {
#pragma unroll
    for (int j = 0; j < TW + 2; ++j) {
      float2 d = RD::cvt(rawd[j]);
      if (MODE == 2) {
        const float2 zz = RD::cvt(rawz[j]);
        d.x *= gelu_grad_fast(zz.x);
        d.y *= gelu_grad_fast(zz.y);
      }
      row[j][0] = d.x;
      row[j][1] = d.y;
    }
  };
  auto load_win = [&](int yy) {
    const bool ok = yy >= 0 && yy < H;
    load_row<TD, ALIGNED>(dp, ok, lval, rval, C, s.x0, W, rawd);
    dp += rstride;
    if (MODE == 2) {
      load_row<TD, ALIGNED>(zp, ok, lval, rval, C, s.x0, W, rawz);
      zp += rstride;
    }
  };
  auto load_x = [&](int yy) {
#pragma unroll
    for (int i = 0; i < TW; ++i)
      rawx[i] = (yy < H && (ALIGNED || s.x0 + i < W)) ? RX::load(xp + i * C) : RX::zero();
    xp += rstride;
  };
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    load_win(s.y0 - 1 + r);
    fill(win[r]);
  }
  load_win(s.y0 + 1);
  load_x(s.y0);
  const int y_end = s.ok ? min(s.y0 + th, H) : s.y0;     // threads past the last strip do no rows
  for (int yb = s.y0; yb < y_end; yb += 3) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int yy = yb + k;
      if (yy < y_end) {
        float (&prev)[TW + 2][2] = win[k];
        float (&cur)[TW + 2][2] = win[(k + 1) % 3];
        float (&next)[TW + 2][2] = win[(k + 2) % 3];
        fill(next);
        float xq[TW][2];
#pragma unroll
        for (int i = 0; i < TW; ++i) { const float2 f = RX::cvt(rawx[i]); xq[i][0] = f.x; xq[i][1] = f.y; }
        load_win(yy + 2);                                  // prefetch for the next iteration
        load_x(yy + 1);
#pragma unroll
        for (int i = 0; i < TW; ++i) {
          if (ALIGNED || s.x0 + i < W) {
            float a0 = MODE >= 1 ? cur[i + 1][0] : 0.f, a1 = MODE >= 1 ? cur[i + 1][1] : 0.f;
            dwa[9][0] += cur[i + 1][0];
            dwa[9][1] += cur[i + 1][1];
            // tap (u, v) pairs pixel q with dz[q - (u-1, v-1)]: u = 0 -> row y+1, u = 2 -> row y-1
#pragma unroll
            for (int v = 0; v < 3; ++v) {
              const int col = i + 2 - v;
              a0 = fmaf(wr[v][0], next[col][0], a0);        a1 = fmaf(wr[v][1], next[col][1], a1);
              a0 = fmaf(wr[3 + v][0], cur[col][0], a0);     a1 = fmaf(wr[3 + v][1], cur[col][1], a1);
              a0 = fmaf(wr[6 + v][0], prev[col][0], a0);    a1 = fmaf(wr[6 + v][1], prev[col][1], a1);
              dwa[v][0] = fmaf(xq[i][0], next[col][0], dwa[v][0]);         dwa[v][1] = fmaf(xq[i][1], next[col][1], dwa[v][1]);
              dwa[3 + v][0] = fmaf(xq[i][0], cur[col][0], dwa[3 + v][0]);  dwa[3 + v][1] = fmaf(xq[i][1], cur[col][1], dwa[3 + v][1]);
              dwa[6 + v][0] = fmaf(xq[i][0], prev[col][0], dwa[6 + v][0]); dwa[6 + v][1] = fmaf(xq[i][1], prev[col][1], dwa[6 + v][1]);
            }
            RX::store(dxp + i * C, a0, a1);
          }
        }
        dxp += rstride;
      }
    }
  }
  const int nvec = C >> 1;
  if (spc > 1) {
    // narrow tensors: the CTA's `spc` strips share their channels; sum them in a fixed order and
    // write one partial per CTA
    __shared__ float2 red[D3_THREADS][10];
#pragma unroll
    for (int uv = 0; uv < 10; ++uv) red[threadIdx.x][uv] = make_float2(dwa[uv][0], dwa[uv][1]);
    __syncthreads();
    if ((int)threadIdx.x < nvec) {
      float* out = partial + (long long)blockIdx.x * 10 * C + s.c;
#pragma unroll
      for (int uv = 0; uv < 10; ++uv) {
        float2 t = red[threadIdx.x][uv];
        for (int m = 1; m < spc; ++m) { const float2 o = red[threadIdx.x + m * nvec][uv]; t.x += o.x; t.y += o.y; }
        *reinterpret_cast<float2*>(out + (long long)uv * C) = t;
      }
    }
  } else if (s.ok) {
    const long long strip = ((long long)blockIdx.x * D3_THREADS + threadIdx.x) / nvec;
    float* out = partial + strip * 10 * C + s.c;
#pragma unroll
    for (int uv = 0; uv < 10; ++uv) *reinterpret_cast<float2*>(out + (long long)uv * C) = make_float2(dwa[uv][0], dwa[uv][1]);
  }
}

}  // namespace

// rows per strip: whole columns when that already fills the GPU, else halved until ~1.5 waves
// of threads exist (measured optimum; never below 8 rows: the 2-row halo is re-read and, in the
// backward, its gelu' recomputed per strip)
int dwconv3_rows_per_strip(int B, int H, int W, int C) {
  const long long per_row_block = (long long)B * ceil_div(W, TW) * (C / 2);
  if (const char* e = std::getenv("DAT_B200_DW3_TH")) { const int v = std::atoi(e); if (v > 0) return v < H ? v : H; }
  static const long long per_sm = [] { const char* e = std::getenv("DAT_B200_DW3_THREADS_PER_SM"); return e ? std::atoll(e) : 700ll; }();
  int th = H;
  while (th > 8 && per_row_block * ceil_div(H, th) < 148ll * per_sm) th = (th + 1) / 2;
  return th;
}

// strips per CTA whose partial sums are merged in shared memory (1 = one partial per strip)
static int strips_per_cta(int C) {
  const int nvec = C / 2;
  return (nvec < D3_THREADS && D3_THREADS % nvec == 0) ? D3_THREADS / nvec : 1;
}
static long long n_partials(int B, int H, int W, int C) {
  const int th = dwconv3_rows_per_strip(B, H, W, C);
  const long long nstrips = (long long)B * ceil_div(W, TW) * ceil_div(H, th);
  return strips_per_cta(C) > 1 ? ceil_div(nstrips * (C / 2), (long long)D3_THREADS) : nstrips;
}

size_t dwconv3_partial_bytes(int B, int H, int W, int C) {
  return align_up((size_t)n_partials(B, H, W, C) * 10 * C * 4, 256);
}

bool dwconv3_supported(int C, int k) { return k == 3 && C % 2 == 0; }

int dwconv3_fwd(const void* x, int x_dt, const float* w, const float* bias, void* y, void* z_out, int y_dt,
                int B, int H, int W, int C, int mode, int flip, cudaStream_t st) {
  DAT_REQUIRE(C % 2 == 0, "dwconv3: C must be even");
  DAT_REQUIRE(mode >= 0 && mode <= 2 && (mode != 2 || z_out != nullptr), "dwconv3: bad mode");
  int th = dwconv3_rows_per_strip(B, H, W, C);
  // the MLP middle (mode 2: z stored, GELU) runs best with half the strip height of the fused backward, whose optimum
  // the common rule above tracks: 115 -> 101 us (stage 0), 60 -> 53 (stage 1), 32.1 -> 29.7 (stage 2), DAT_B200_DW3_TH sweep
  if (mode == 2 && th >= 16 && std::getenv("DAT_B200_DW3_TH") == nullptr) th = th / 2;
  const int sx = ceil_div(W, TW), sy = ceil_div(H, th);
  const long long threads = (long long)B * sx * sy * (C / 2);
  const unsigned grid = (unsigned)ceil_div(threads, (long long)D3_THREADS);
#define LAUNCH_A(TI, TO, MD, AL)                                                                        \
  launch_k(dwconv3_fwd_kernel<TI, TO, MD, AL>, grid, D3_THREADS, 0, st, (const TI*)x, w, bias, (TO*)y, (TO*)z_out, B, \
                                                                   H, W, C, th, sx, sy, flip)
#define LAUNCH(TI, TO, MD)                                                    \
  do {                                                                        \
    if (W % TW == 0) LAUNCH_A(TI, TO, MD, true); else LAUNCH_A(TI, TO, MD, false); \
  } while (0)
#define LAUNCH_M(TI, TO)                                                                  \
  do {                                                                                    \
    if (mode == 0) LAUNCH(TI, TO, 0); else if (mode == 1) LAUNCH(TI, TO, 1); else LAUNCH(TI, TO, 2); \
  } while (0)
  if (x_dt == DAT_F32 && y_dt == DAT_F32) LAUNCH_M(float, float);
  else if (x_dt == DAT_F32) LAUNCH_M(float, bf16);
  else if (y_dt == DAT_F32) LAUNCH_M(bf16, float);
  else LAUNCH_M(bf16, bf16);
#undef LAUNCH_M
#undef LAUNCH
#undef LAUNCH_A
  DAT_LAUNCH_OK("dwconv3_fwd_kernel");
  return DAT_OK;
}

// dx (dtype of x), dw (C, 1, 3, 3), db (C, may be NULL): all overwritten.  dy / z have dtype d_dt;
// z only for mode 2.  ws >= dwconv3_partial_bytes.
int dwconv3_bwd(const void* x, int x_dt, const void* dy, const void* z, int d_dt, const float* w, void* dx,
                float* dw, float* db, int B, int H, int W, int C, int mode, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  DAT_REQUIRE(C % 2 == 0, "dwconv3: C must be even");
  DAT_REQUIRE(mode >= 0 && mode <= 2 && (mode != 2 || z != nullptr), "dwconv3_bwd: bad mode");
  DAT_REQUIRE(ws_bytes >= dwconv3_partial_bytes(B, H, W, C), "dwconv3_bwd: workspace too small");
  const int th = dwconv3_rows_per_strip(B, H, W, C);
  const int sx = ceil_div(W, TW), sy = ceil_div(H, th);
  const long long nstrips = (long long)B * sx * sy;
  const unsigned grid = (unsigned)ceil_div(nstrips * (C / 2), (long long)D3_THREADS);
  float* part = (float*)ws;
  const int spc = strips_per_cta(C);
#define LAUNCH_A(TX, TD, MD, AL)                                                                              \
  launch_k(dwconv3_bwd_kernel<TX, TD, MD, AL>, grid, D3_THREADS, 0, st, (const TX*)x, (const TD*)dy, (const TD*)z, w, \
                                                                   (TX*)dx, part, B, H, W, C, th, sx, sy, spc)
#define LAUNCH(TX, TD, MD)                                                    \
  do {                                                                        \
    if (W % TW == 0) LAUNCH_A(TX, TD, MD, true); else LAUNCH_A(TX, TD, MD, false); \
  } while (0)
#define LAUNCH_M(TX, TD)                                                                  \
  do {                                                                                    \
    if (mode == 0) LAUNCH(TX, TD, 0); else if (mode == 1) LAUNCH(TX, TD, 1); else LAUNCH(TX, TD, 2); \
  } while (0)
  if (x_dt == DAT_F32 && d_dt == DAT_F32) LAUNCH_M(float, float);
  else if (x_dt == DAT_F32) LAUNCH_M(float, bf16);
  else if (d_dt == DAT_F32) LAUNCH_M(bf16, float);
  else LAUNCH_M(bf16, bf16);
#undef LAUNCH_M
#undef LAUNCH
#undef LAUNCH_A
  DAT_LAUNCH_OK("dwconv3_bwd_kernel");
  return dwconv_wgrad_reduce(part, (int)n_partials(B, H, W, C), 9, C, dw, db, st);
}

}  // namespace dat
