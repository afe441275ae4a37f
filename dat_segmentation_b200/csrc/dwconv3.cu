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
// exp(-z^2/2) between the cdf and the pdf: the kernels stay HBM-bound.
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

// erf(|z| / sqrt 2) pieces (Abramowitz-Stegun 7.1.26): returns q = poly(t) * exp(-z^2 / 2), so that
// Phi(z) = z >= 0 ? 1 - q / 2 : q / 2, and e = exp(-z^2 / 2).
__device__ __forceinline__ float as_q(float z, float& e) {
  const float a = fabsf(z) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));
  e = __expf(-a * a);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return p * t * e;
}
__device__ __forceinline__ float gelu_fast(float z) {
  float e;
  const float q = 0.5f * as_q(z, e);
  return z * (z >= 0.f ? 1.0f - q : q);
}
__device__ __forceinline__ float gelu_grad_fast(float z) {
  float e;
  const float q = 0.5f * as_q(z, e);
  return (z >= 0.f ? 1.0f - q : q) + z * e * 0.39894228040143267794f;
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

// raw loads of window row yy (columns x0-1 .. x0+TW), zeros outside the image
template <typename T>
__device__ __forceinline__ void load_row(const T* __restrict__ img, int yy, int x0, int H, int W, int C,
                                         typename Raw2<T>::type (&raw)[TW + 2]) {
  const bool row_ok = yy >= 0 && yy < H;
  const T* rowp = img + ((long long)yy * W + (x0 - 1)) * C;
#pragma unroll
  for (int j = 0; j < TW + 2; ++j) {
    const int xx = x0 - 1 + j;
    raw[j] = (row_ok && xx >= 0 && xx < W) ? Raw2<T>::load(rowp + (long long)j * C) : Raw2<T>::zero();
  }
}

template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(D3_THREADS, 5)
dwconv3_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   TO* __restrict__ y, TO* __restrict__ z_out, int B, int H, int W, int C, int th,
                   int strips_x, int strips_y, int flip) {
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
  const long long img_off = (long long)s.b * H * W * C + s.c;
  const TI* img = x + img_off;
  float win[3][TW + 2][2];
  typename RI::type raw[TW + 2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {            // rows y0-1, y0 -> slots 0, 1
    load_row<TI>(img, s.y0 - 1 + r, s.x0, H, W, C, raw);
#pragma unroll
    for (int j = 0; j < TW + 2; ++j) { const float2 f = RI::cvt(raw[j]); win[r][j][0] = f.x; win[r][j][1] = f.y; }
  }
  load_row<TI>(img, s.y0 + 1, s.x0, H, W, C, raw);      // row y0+1, converted inside the loop
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
        load_row<TI>(img, yy + 2, s.x0, H, W, C, raw);   // prefetch: consumed by the next iteration
        TO* yrow = y + img_off + ((long long)yy * W + s.x0) * C;
        TO* zrow = MODE == 2 ? z_out + img_off + ((long long)yy * W + s.x0) * C : nullptr;
#pragma unroll
        for (int i = 0; i < TW; ++i) {
          if (s.x0 + i < W) {
            float a0 = b0, a1 = b1;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
              a0 = fmaf(wr[v][0], prev[i + v][0], a0);      a1 = fmaf(wr[v][1], prev[i + v][1], a1);
              a0 = fmaf(wr[3 + v][0], cur[i + v][0], a0);   a1 = fmaf(wr[3 + v][1], cur[i + v][1], a1);
              a0 = fmaf(wr[6 + v][0], next[i + v][0], a0);  a1 = fmaf(wr[6 + v][1], next[i + v][1], a1);
            }
            if (MODE >= 1) { a0 += cur[i + 1][0]; a1 += cur[i + 1][1]; }
            if (MODE == 2) {
              Raw2<TO>::store(zrow + (long long)i * C, a0, a1);
              a0 = gelu_fast(a0);
              a1 = gelu_fast(a1);
            }
            Raw2<TO>::store(yrow + (long long)i * C, a0, a1);
          }
        }
      }
    }
  }
}

// fused backward; partial[strip][10][C] (rows 0-8 = dw taps, row 9 = db)
template <typename TX, typename TD, int MODE>
__global__ void __launch_bounds__(D3_THREADS, 4)
dwconv3_bwd_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, const TD* __restrict__ z,
                   const float* __restrict__ w, TX* __restrict__ dx, float* __restrict__ partial, int B,
                   int H, int W, int C, int th, int strips_x, int strips_y, int spc) {
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
  const long long img_off = (long long)s.b * H * W * C + s.c;
  const TD* dimg = dy + img_off;
  const TD* zimg = MODE == 2 ? z + img_off : nullptr;
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
  if (!s.ok) H = 0;                       // every load of an idle thread is masked off
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    load_row<TD>(dimg, s.y0 - 1 + r, s.x0, H, W, C, rawd);
    if (MODE == 2) load_row<TD>(zimg, s.y0 - 1 + r, s.x0, H, W, C, rawz);
    fill(win[r]);
  }
  load_row<TD>(dimg, s.y0 + 1, s.x0, H, W, C, rawd);
  if (MODE == 2) load_row<TD>(zimg, s.y0 + 1, s.x0, H, W, C, rawz);
  const int y_end = s.ok ? min(s.y0 + th, H) : s.y0;     // threads past the last strip do no rows
  auto load_x = [&](int yy) {
    const TX* xrow = x + img_off + ((long long)yy * W + s.x0) * C;
#pragma unroll
    for (int i = 0; i < TW; ++i)
      rawx[i] = (yy < H && s.x0 + i < W) ? RX::load(xrow + (long long)i * C) : RX::zero();
  };
  load_x(s.y0);
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
        load_row<TD>(dimg, yy + 2, s.x0, H, W, C, rawd);   // prefetch for the next iteration
        if (MODE == 2) load_row<TD>(zimg, yy + 2, s.x0, H, W, C, rawz);
        load_x(yy + 1);
        TX* dxrow = dx + img_off + ((long long)yy * W + s.x0) * C;
#pragma unroll
        for (int i = 0; i < TW; ++i) {
          if (s.x0 + i < W) {
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
            RX::store(dxrow + (long long)i * C, a0, a1);
          }
        }
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

// rows per strip: whole columns when that already fills the GPU, else halved until ~1300
// threads per SM are in flight (never below 8 rows: the 2-row halo is re-read per strip)
int dwconv3_rows_per_strip(int B, int H, int W, int C) {
  const long long per_row_block = (long long)B * ceil_div(W, TW) * (C / 2);
  int th = H;
  while (th > 8 && per_row_block * ceil_div(H, th) < 148ll * 1300) th = (th + 1) / 2;
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
  const int th = dwconv3_rows_per_strip(B, H, W, C);
  const int sx = ceil_div(W, TW), sy = ceil_div(H, th);
  const long long threads = (long long)B * sx * sy * (C / 2);
  const unsigned grid = (unsigned)ceil_div(threads, (long long)D3_THREADS);
#define LAUNCH(TI, TO, MD)                                                                              \
  dwconv3_fwd_kernel<TI, TO, MD><<<grid, D3_THREADS, 0, st>>>((const TI*)x, w, bias, (TO*)y, (TO*)z_out, B, H, \
                                                               W, C, th, sx, sy, flip)
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
#define LAUNCH(TX, TD, MD)                                                                                \
  dwconv3_bwd_kernel<TX, TD, MD><<<grid, D3_THREADS, 0, st>>>((const TX*)x, (const TD*)dy, (const TD*)z, w, \
                                                               (TX*)dx, part, B, H, W, C, th, sx, sy, spc)
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
  DAT_LAUNCH_OK("dwconv3_bwd_kernel");
  return dwconv_wgrad_reduce(part, (int)n_partials(B, H, W, C), 9, C, dw, db, st);
}

}  // namespace dat
