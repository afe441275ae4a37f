// Channel-last depthwise 7 x 7 convolution (stride 1, padding 3): the 'X' mixer of the DAT++ stages
// (dat.py:118-121, 140-141), forward / data gradient and weight gradient.
//
// Same thread mapping as dwconv3.cu (2 consecutive channels x a strip of TW columns, walking down
// the rows, consecutive lanes = consecutive channel pairs), but 7 window rows do not fit in
// registers, so the kernels are organised around the INPUT row instead:
//   forward / dgrad  each loaded input row (TW + 6 columns) is scattered into 7 rolling output-row
//                    accumulators (7 x TW x 2 registers, statically rotated); an output row is
//                    stored when its last input row has passed.  The 49 x 2 filter taps of the
//                    thread sit in shared memory, one column per thread (conflict-free LDS.64).
//   wgrad            dw[u][v] = sum_p dz[p] x[p + (u-3, v-3)]: each loaded x row meets the dz rows
//                    r - u + 3 kept in a register ring.  The 49 taps are split over two launches
//                    (u = 0..3 and u = 4..6) so that accumulators + ring fit in registers.
// 49 FMAs per element and pass: these kernels sit between the HBM and the FP32 roofline.
// Per-strip partial sums are reduced in a fixed order (deterministic, no atomics).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int TW = 4;
constexpr int D7_THREADS = 128;
constexpr int RAD = 3;
constexpr int WIN = TW + 2 * RAD;   // 10 columns per input row

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

struct Strip {
  int b, x0, y0, c;
  bool ok;
};
__device__ __forceinline__ Strip strip_of(int B, int C, int th, int strips_x, int strips_y) {
  const int nvec = C >> 1;
  const long long t = (long long)blockIdx.x * D7_THREADS + threadIdx.x;
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

// one input row, columns x0-3 .. x0+TW+2; rp points at column x0.  ALIGNED (W % TW == 0): the left /
// right halo triples are valid or invalid as a whole (lval / rval, per-thread constants).
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_row(const T* __restrict__ rp, bool row_ok, bool lval, bool rval, int C,
                                         int x0, int W, typename Raw2<T>::type (&raw)[WIN]) {
#pragma unroll
  for (int j = 0; j < WIN; ++j) {
    bool ok;
    if (ALIGNED) ok = row_ok && (j < RAD ? lval : (j >= RAD + TW ? rval : true));
    else { const int xx = x0 - RAD + j; ok = row_ok && xx >= 0 && xx < W; }
    raw[j] = ok ? Raw2<T>::load(rp + (j - RAD) * C) : Raw2<T>::zero();
  }
}

// y = conv7(x) + b (flip = 0) or the data gradient dx = conv7(dy, flipped w) (flip = 1, bias NULL)
template <typename TI, typename TO, bool ALIGNED>
__global__ void __launch_bounds__(D7_THREADS, 4)
dwconv7_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   TO* __restrict__ y, int B, int H, int W, int C, int th, int strips_x, int strips_y,
                   int flip) {
  pdl_enter();
  using RI = Raw2<TI>;
  // The CTA's threads cover min(C, 256) consecutive channels (modulo C): their filters are staged cooperatively
  // with coalesced loads into w_s[49][256] (tap-major), each thread then reads its channel pair as one LDS.64.
  // (Per-thread staging - 98 scattered loads + 49 stores per thread - was ~40 % of the stall samples at stage 2.)
  // WPACK (bf16 activations, i.e. autocast: the library convolution runs on bf16-rounded weights too): a channel
  // pair's two taps are one bf16x2 word, [49][D7_THREADS] words.  The weight fetches were as expensive as the math: 49
  // LDS.64 per 196 FFMA2 and warp = 200 KB of shared-memory reads per round of the SM's 16 warps = 1568 cycles at
  // 128 B/clk, exactly the FMA pipe's 4 x 392 cycles per scheduler; halving the bytes leaves the FMA pipe as the bound.
  constexpr bool WPACK = sizeof(TI) == 2 || sizeof(TO) == 2;
  extern __shared__ __align__(16) float w_s[];             // [49][2 * D7_THREADS] floats or [49][D7_THREADS] words
  const int nst = C < 2 * D7_THREADS ? C : 2 * D7_THREADS;
  {
    const int c_first = (int)(((long long)blockIdx.x * D7_THREADS) % (C >> 1)) * 2;
    if (WPACK) {
      uint32_t* w_p = reinterpret_cast<uint32_t*>(w_s);
      for (int idx = threadIdx.x; idx < (nst >> 1) * 49; idx += D7_THREADS) {
        const int cp = idx / 49, uv = idx - cp * 49;
        int cg = c_first + 2 * cp;
        if (cg >= C) cg -= C;
        const int t = flip ? 48 - uv : uv;
        const __nv_bfloat162 pk = __floats2bfloat162_rn(w[cg * 49 + t], w[(cg + 1) * 49 + t]);
        w_p[uv * D7_THREADS + cp] = *reinterpret_cast<const uint32_t*>(&pk);
      }
    } else {
      for (int idx = threadIdx.x; idx < nst * 49; idx += D7_THREADS) {
        const int cc = idx / 49, uv = idx - cc * 49;
        int cg = c_first + cc;
        if (cg >= C) cg -= C;
        w_s[uv * (2 * D7_THREADS) + cc] = w[cg * 49 + (flip ? 48 - uv : uv)];   // row pitch fixed: immediate offsets below
      }
    }
  }
  __syncthreads();
  const Strip s = strip_of(B, C, th, strips_x, strips_y);
  if (!s.ok) return;
  const float* w_mine = w_s + (C < 2 * D7_THREADS ? (2 * (int)threadIdx.x) % C : 2 * (int)threadIdx.x);
  const uint32_t* w_mine_p = reinterpret_cast<const uint32_t*>(w_s) +
                             (C < 2 * D7_THREADS ? (int)threadIdx.x % (C >> 1) : (int)threadIdx.x);
  const float b0 = bias != nullptr ? bias[s.c] : 0.f, b1 = bias != nullptr ? bias[s.c + 1] : 0.f;
  const bool lval = s.x0 > 0, rval = s.x0 + TW < W;
  const int rstride = W * C;
  const long long pix0 = (long long)s.b * H * rstride + (long long)s.y0 * rstride + s.x0 * C + s.c;
  const TI* rp = x + pix0 - (long long)RAD * rstride;      // input row being loaded: starts at y0 - 3
  TO* yp = y + pix0;
  const int th_eff = min(th, H - s.y0);
  // channel pairs are kept as float2 and updated with the packed fp32 FMA (FFMA2): one issue slot per
  // two multiply-adds in this issue-bound kernel
  float2 acc[7][TW];
#pragma unroll
  for (int q = 0; q < 7; ++q)
#pragma unroll
    for (int i = 0; i < TW; ++i) acc[q][i] = make_float2(0.f, 0.f);
  typename RI::type raw[WIN];
  load_row<TI, ALIGNED>(rp, s.y0 - RAD >= 0, lval, rval, C, s.x0, W, raw);
  rp += rstride;
  const int steps = th_eff + 2 * RAD;                       // input rows y0-3 .. y0+th_eff+2
  for (int nb = 0; nb < steps; nb += 7) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int n = nb + k;
      if (n < steps) {
        const int r = s.y0 - RAD + n;
        float2 in[WIN];
#pragma unroll
        for (int j = 0; j < WIN; ++j) in[j] = RI::cvt(raw[j]);
        load_row<TI, ALIGNED>(rp, r + 1 >= 0 && r + 1 < H && n + 1 < steps, lval, rval, C, s.x0, W, raw);   // prefetch
        rp += rstride;
        if (r >= 0 && r < H) {
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            const int m = n - u;                             // output row (relative) fed through tap row u
            if (m >= 0 && m < th_eff) {
              float2 (&a)[TW] = acc[(k - u + 7) % 7];
#pragma unroll
              for (int v = 0; v < 7; ++v) {
                float2 wv;
                if (WPACK) {
                  const uint32_t pk = w_mine_p[(u * 7 + v) * D7_THREADS];
                  wv = make_float2(__uint_as_float(pk << 16), __uint_as_float(pk & 0xffff0000u));
                } else {
                  wv = *reinterpret_cast<const float2*>(w_mine + (u * 7 + v) * (2 * D7_THREADS));
                }
#pragma unroll
                for (int i = 0; i < TW; ++i) a[i] = __ffma2_rn(wv, in[i + v], a[i]);
              }
            }
          }
        }
        if (n >= 2 * RAD) {                                  // output row n - 6 has seen its last input row
          float2 (&a)[TW] = acc[(k + 1) % 7];
#pragma unroll
          for (int i = 0; i < TW; ++i) {
            if (ALIGNED || s.x0 + i < W) Raw2<TO>::store(yp + i * C, a[i].x + b0, a[i].y + b1);
            a[i] = make_float2(0.f, 0.f);
          }
          yp += rstride;
        }
      }
    }
  }
}

// taps u in [U0, U0 + NU): partial[strip or CTA][50][C] rows (U0 * 7 .. (U0 + NU) * 7) and, for U0 == 0,
// row 49 (bias gradient)
template <typename TX, typename TD, int U0, int NU, bool ALIGNED>
__global__ void __launch_bounds__(D7_THREADS, 3)
dwconv7_wgrad_kernel(const TX* __restrict__ x, const TD* __restrict__ dz, float* __restrict__ partial, int B,
                     int H, int W, int C, int th, int strips_x, int strips_y, int spc) {
  pdl_enter();
  using RX = Raw2<TX>;
  using RD = Raw2<TD>;
  const Strip s = strip_of(B, C, th, strips_x, strips_y);
  const int th_eff = s.ok ? min(th, H - s.y0) : 0;
  if (!s.ok) H = 0;
  const bool lval = s.x0 > 0, rval = s.x0 + TW < W;
  const int rstride = W * C;
  const long long pix0 = s.ok ? (long long)s.b * H * rstride + (long long)s.y0 * rstride + s.x0 * C + s.c : 0;
  const TX* xp = x + pix0 + (long long)(U0 - RAD) * rstride;    // x row of step j: y0 - 3 + U0 + j
  const TD* dp = dz + pix0;                                      // dz row j enters the ring at step j
  float acc[NU * 7][2], dbs[2] = {0.f, 0.f};
#pragma unroll
  for (int q = 0; q < NU * 7; ++q) acc[q][0] = acc[q][1] = 0.f;
  float ring[NU][TW][2];
  typename RX::type rawx[WIN];
  typename RD::type rawd[TW];
  auto load_dz = [&](int m) {
#pragma unroll
    for (int i = 0; i < TW; ++i)
      rawd[i] = (m < th_eff && (ALIGNED || s.x0 + i < W)) ? RD::load(dp + i * C) : RD::zero();
    dp += rstride;
  };
  const int steps = th_eff > 0 ? th_eff + NU - 1 : 0;
  {
    const int r0 = s.y0 - RAD + U0;
    load_row<TX, ALIGNED>(xp, r0 >= 0 && r0 < H, lval, rval, C, s.x0, W, rawx);
    xp += rstride;
    load_dz(0);
  }
  for (int jb = 0; jb < steps; jb += NU) {
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int j = jb + k;
      if (j < steps) {
        const int r = s.y0 - RAD + U0 + j;
        float in[WIN][2];
#pragma unroll
        for (int q = 0; q < WIN; ++q) { const float2 f = RX::cvt(rawx[q]); in[q][0] = f.x; in[q][1] = f.y; }
        {                                                    // dz row j -> ring slot j % NU
          float (&slot)[TW][2] = ring[k];
#pragma unroll
          for (int i = 0; i < TW; ++i) {
            const float2 f = RD::cvt(rawd[i]);
            slot[i][0] = f.x; slot[i][1] = f.y;
            if (U0 == 0) { dbs[0] += f.x; dbs[1] += f.y; }
          }
        }
        load_row<TX, ALIGNED>(xp, r + 1 >= 0 && r + 1 < H && j + 1 < steps, lval, rval, C, s.x0, W, rawx);   // prefetch
        xp += rstride;
        load_dz(j + 1);
        if (r >= 0 && r < H) {
#pragma unroll
          for (int ul = 0; ul < NU; ++ul) {
            const int m = j - ul;                            // dz row paired through tap row U0 + ul
            if (m >= 0 && m < th_eff) {
              float (&d)[TW][2] = ring[(k - ul + NU) % NU];
#pragma unroll
              for (int v = 0; v < 7; ++v) {
                float a0 = acc[ul * 7 + v][0], a1 = acc[ul * 7 + v][1];
#pragma unroll
                for (int i = 0; i < TW; ++i) {
                  a0 = fmaf(d[i][0], in[i + v][0], a0);
                  a1 = fmaf(d[i][1], in[i + v][1], a1);
                }
                acc[ul * 7 + v][0] = a0;
                acc[ul * 7 + v][1] = a1;
              }
            }
          }
        }
      }
    }
  }
  constexpr int NOUT = NU * 7 + (U0 == 0 ? 1 : 0);
  const int nvec = C >> 1;
  auto out_row = [&](int q) { return q < NU * 7 ? U0 * 7 + q : 49; };
  if (spc > 1) {
    __shared__ float2 red[D7_THREADS][NOUT];
#pragma unroll
    for (int q = 0; q < NOUT; ++q)
      red[threadIdx.x][q] = q < NU * 7 ? make_float2(acc[q < NU * 7 ? q : 0][0], acc[q < NU * 7 ? q : 0][1]) : make_float2(dbs[0], dbs[1]);
    __syncthreads();
    if ((int)threadIdx.x < nvec) {
      float* out = partial + (long long)blockIdx.x * 50 * C + s.c;
#pragma unroll
      for (int q = 0; q < NOUT; ++q) {
        float2 t = red[threadIdx.x][q];
        for (int mm = 1; mm < spc; ++mm) { const float2 o = red[threadIdx.x + mm * nvec][q]; t.x += o.x; t.y += o.y; }
        *reinterpret_cast<float2*>(out + (long long)out_row(q) * C) = t;
      }
    }
  } else if (s.ok) {
    const long long strip = ((long long)blockIdx.x * D7_THREADS + threadIdx.x) / nvec;
    float* out = partial + strip * 50 * C + s.c;
#pragma unroll
    for (int q = 0; q < NOUT; ++q)
      *reinterpret_cast<float2*>(out + (long long)out_row(q) * C) =
          q < NU * 7 ? make_float2(acc[q < NU * 7 ? q : 0][0], acc[q < NU * 7 ? q : 0][1]) : make_float2(dbs[0], dbs[1]);
  }
}

int strips_per_cta(int C) {
  const int nvec = C / 2;
  return (nvec < D7_THREADS && D7_THREADS % nvec == 0) ? D7_THREADS / nvec : 1;
}

// rows per strip: the split of the H rows into `sy` blocks that minimises waves x (rows + halo)
int rows_per_strip(int B, int H, int W, int C, int ctas_per_sm, int halo) {
  if (const char* e = std::getenv("DAT_B200_DW7_TH")) { const int v = std::atoi(e); if (v > 0) return v < H ? v : H; }
  const long long per_row_block = (long long)B * ceil_div(W, TW) * (C / 2);
  const long long resident = 148ll * ctas_per_sm * D7_THREADS;
  int best_th = H;
  long long best = -1;
  for (int sy = 1; sy <= H; ++sy) {
    const int th = ceil_div(H, sy);
    if (th < 4 && sy > 1) break;
    const long long waves = ceil_div(per_row_block * ceil_div(H, th), resident);
    const long long cost = waves * (th + halo);
    if (best < 0 || cost < best) { best = cost; best_th = th; }
  }
  return best_th;
}

long long n_partials(int B, int H, int W, int C, int th) {
  const long long nstrips = (long long)B * ceil_div(W, TW) * ceil_div(H, th);
  return strips_per_cta(C) > 1 ? ceil_div(nstrips * (C / 2), (long long)D7_THREADS) : nstrips;
}

}  // namespace

bool dwconv7_supported(int C, int k) { return k == 7 && C % 2 == 0; }

size_t dwconv7_partial_bytes(int B, int H, int W, int C) {
  const int th = rows_per_strip(B, H, W, C, 3, 5);
  return align_up((size_t)n_partials(B, H, W, C, th) * 50 * C * 4, 256);
}

// Rows per strip of the forward / data-gradient kernel.  Every input row costs the same whatever the number of
// output rows it feeds, so taller strips do less work per output ((th + 6) / th input rows each); measured
// optimum at batch 16 (DAT_B200_DW7_TH sweep, profiles/r01_kernel_rooflines.md): 32 rows for 128-row maps, 16
// for 64 / 32, 8 for 16 - i.e. as tall as possible while ~32 k threads (7 warps per SM) remain.
int fwd_rows_per_strip(int B, int H, int W, int C) {
  if (const char* e = std::getenv("DAT_B200_DW7_TH")) { const int v = std::atoi(e); if (v > 0) return v < H ? v : H; }
  const long long per_row_block = (long long)B * ceil_div(W, TW) * (C / 2);
  int th = H >= 128 ? 32 : (H >= 32 ? 16 : 8);
  if (th > H) th = H;
  while (th > 4 && per_row_block * ceil_div(H, th) < 24 * 1024) th = (th + 1) / 2;
  return th;
}

int dwconv7_fwd(const void* x, int x_dt, const float* w, const float* bias, void* y, int y_dt, int B, int H,
                int W, int C, int flip, cudaStream_t st) {
  DAT_REQUIRE(C % 2 == 0, "dwconv7: C must be even");
  const int th = fwd_rows_per_strip(B, H, W, C);
  const int sx = ceil_div(W, TW), sy = ceil_div(H, th);
  const unsigned grid = (unsigned)ceil_div((long long)B * sx * sy * (C / 2), (long long)D7_THREADS);
  constexpr int smem = 49 * D7_THREADS * (int)sizeof(float2);
#define LAUNCH_A(TI, TO, AL)                                                                                  \
  do {                                                                                                        \
    auto kern = dwconv7_fwd_kernel<TI, TO, AL>;                                                               \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));               \
    launch_k(kern, grid, D7_THREADS, smem, st, (const TI*)x, w, bias, (TO*)y, B, H, W, C, th, sx, sy, flip);        \
  } while (0)
#define LAUNCH(TI, TO)                                                \
  do {                                                                \
    if (W % TW == 0) LAUNCH_A(TI, TO, true); else LAUNCH_A(TI, TO, false); \
  } while (0)
  if (x_dt == DAT_F32 && y_dt == DAT_F32) LAUNCH(float, float);
  else if (x_dt == DAT_F32) LAUNCH(float, bf16);
  else if (y_dt == DAT_F32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
#undef LAUNCH_A
  DAT_LAUNCH_OK("dwconv7_fwd_kernel");
  return DAT_OK;
}

// dw (C, 1, 7, 7), db (C, may be NULL): overwritten.  ws >= dwconv7_partial_bytes.
int dwconv7_wgrad(const void* x, int x_dt, const void* dz, int dz_dt, float* dw, float* db, int B, int H,
                  int W, int C, void* ws, size_t ws_bytes, cudaStream_t st) {
  DAT_REQUIRE(C % 2 == 0, "dwconv7: C must be even");
  DAT_REQUIRE(ws_bytes >= dwconv7_partial_bytes(B, H, W, C), "dwconv7_wgrad: workspace too small");
  const int th = rows_per_strip(B, H, W, C, 3, 5);
  const int sx = ceil_div(W, TW), sy = ceil_div(H, th);
  const unsigned grid = (unsigned)ceil_div((long long)B * sx * sy * (C / 2), (long long)D7_THREADS);
  const int spc = strips_per_cta(C);
  float* part = (float*)ws;
#define LAUNCH_A(TX, TD, AL)                                                                                   \
  do {                                                                                                         \
    launch_k(dwconv7_wgrad_kernel<TX, TD, 0, 4, AL>, grid, D7_THREADS, 0, st, (const TX*)x, (const TD*)dz, part, B, H, W, \
                                                                        C, th, sx, sy, spc);                   \
    launch_k(dwconv7_wgrad_kernel<TX, TD, 4, 3, AL>, grid, D7_THREADS, 0, st, (const TX*)x, (const TD*)dz, part, B, H, W, \
                                                                        C, th, sx, sy, spc);                   \
  } while (0)
#define LAUNCH(TX, TD)                                                \
  do {                                                                \
    if (W % TW == 0) LAUNCH_A(TX, TD, true); else LAUNCH_A(TX, TD, false); \
  } while (0)
  if (x_dt == DAT_F32 && dz_dt == DAT_F32) LAUNCH(float, float);
  else if (x_dt == DAT_F32) LAUNCH(float, bf16);
  else if (dz_dt == DAT_F32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
#undef LAUNCH_A
  DAT_LAUNCH_OK("dwconv7_wgrad_kernel");
  return dwconv_wgrad_reduce(part, (int)n_partials(B, H, W, C, th), 49, C, dw, db, st);
}

}  // namespace dat
