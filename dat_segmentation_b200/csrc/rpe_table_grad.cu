// d rpe_table of the tensor-core attention backward, as tensor-core GEMMs instead of a per-score scatter.
//
// The bias of score (m, n) is the bilinear sample of the head's table at (iy, ix) with
//   ix = c * ax + (1 - pos_x[n]) * kx,   iy = r * ay + (1 - pos_y[n]) * ky,   m = (r, c)
// (dat_blocks.py:198-212; ax = (Tw-1) / (2 (W-1)), kx = (Tw-1) / 4), i.e. separable hat weights
//   w(m, n; y, x) = hat(iy - y) hat(ix - x),  hat(d) = max(0, 1 - |d|)
// (taps outside the table are simply never formed: zero padding).  For one sample n the table gradient is
//   dT_n = A_n^T dS_n B_n,   dS_n (H x W) = dS[(r, c), n],  A_n[r][y] = hat(iy - y),  B_n[c][x] = hat(ix - x)
// two small GEMMs whose hat operands are generated in registers.  A warp owns a 32 x 32 tile of the
// (transposed) table gradient in mma.sync accumulators and walks over samples:
//   E^T  (x, r) = B_n^T (x, c) . dS_n^T (c, r)      m16n8k16, B fragments = 32-bit loads of the staged dS
//   dT^T (x, y) += E^T (x, r) . A_n (r, y)          E^T accumulators re-used as the A fragments (bf16)
// Only the (c, r) ranges whose footprint meets the tile are visited.  dS arrives as bf16 in groups of 8
// samples, [b*heads][Ns/8][m][8], from attn_bwd_tc_kernel<.., TBL = false> (both sides move whole 16-byte
// pieces that are contiguous over the queries); a CTA stages [8 ng samples][rows][W] of it (transposed,
// padded rows: conflict-free fragment loads) per step.  ~5x fewer instructions than the scatter it replaces and
// off the critical path (d rpe_table is read by nothing later in the backward: side stream).
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int TG_THREADS = 256;   // 8 warps
constexpr int TG_NB = 8;          // samples staged per step (one 16-byte load per query)
constexpr int TG_TILE = 32;       // table tile edge per warp

struct TgArgs {
  int H, W, HW, heads, G, hg, Ns, Th, Tw;
  int rows_chunk, n_per_cta, tiles_x, ntiles, pitch, ng;   // ng: 8-sample groups staged per step
  float ax, ay, kx, ky;
};

__device__ __forceinline__ float hat(float d) { return __saturatef(1.0f - fabsf(d)); }   // one FADD.SAT
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// TPW: table tiles per warp (1: up to 8 tiles per table, warps share tiles and split the samples; 2: 9-16 tiles)
template <int TPW>
__global__ void __launch_bounds__(TG_THREADS)
rpe_table_grad_kernel(const bf16* __restrict__ ds, const float* __restrict__ pos, float* __restrict__ d_table,
                      TgArgs a) {
  extern __shared__ __align__(16) uint8_t tg_smem[];
  bf16* tile = reinterpret_cast<bf16*>(tg_smem);                 // [8 ng][rows_chunk][pitch]
  float* sred = reinterpret_cast<float*>(tg_smem);               // [Th * Tw], after the sample loop
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int bh = blockIdx.z, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  const int r_base = blockIdx.y * a.rows_chunk;
  const int rows = min(a.rows_chunk, a.H - r_base);              // multiple of 16
  const int n_begin = blockIdx.x * a.n_per_cta, n_end = min(a.Ns, n_begin + a.n_per_cta);
  // tile / sample ownership
  int my_tile[TPW];
  int n_phase = 0, n_step = 1;
  if (TPW == 1) {
    const int share = a.ntiles >= 8 ? 1 : 8 / a.ntiles;          // warps per tile (ntiles in {1, 2, 4, 8})
    my_tile[0] = a.ntiles >= 8 ? warp : warp % a.ntiles;
    n_phase = a.ntiles >= 8 ? 0 : warp / a.ntiles;
    n_step = share;
  } else {
#pragma unroll
    for (int i = 0; i < TPW; ++i) my_tile[i] = warp + 8 * i;     // may be >= ntiles: idle slot
  }
  float acc[TPW][2][4][4];
#pragma unroll
  for (int i = 0; i < TPW; ++i)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int yt = 0; yt < 4; ++yt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][mt][yt][e] = 0.f;

  const float* pbase = pos + ((long long)b * a.G + g) * a.Ns * 2;
  const bf16* dsb = ds + (long long)bh * a.HW * a.Ns + (long long)r_base * a.W * TG_NB;   // + group * HW * 8
  const int n_rows_px = rows * a.W;
  const int nstage = a.ng * TG_NB;
  const float inv_ax = 1.0f / a.ax, inv_ay = 1.0f / a.ay;

  for (int n0 = n_begin; n0 < n_end; n0 += nstage) {
    __syncthreads();
    // stage dS[rows x W queries][n0 .. n0 + 8 ng) transposed into tile[nl][r][c]
    const int groups = min(a.ng, (n_end - n0) / TG_NB);
    for (int i = threadIdx.x; i < n_rows_px * groups; i += TG_THREADS) {
      const int gi = i / n_rows_px, m = i - gi * n_rows_px;
      const uint4 v = *reinterpret_cast<const uint4*>(dsb + ((long long)(n0 / TG_NB + gi) * a.HW + m) * TG_NB);
      const int r = m / a.W, c = m - r * a.W;
      bf16* dst = tile + ((long long)gi * TG_NB * a.rows_chunk + r) * a.pitch + c;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        reinterpret_cast<uint16_t*>(dst + (long long)(2 * i) * a.rows_chunk * a.pitch)[0] = (uint16_t)(w[i] & 0xffffu);
        reinterpret_cast<uint16_t*>(dst + (long long)(2 * i + 1) * a.rows_chunk * a.pitch)[0] = (uint16_t)(w[i] >> 16);
      }
    }
    __syncthreads();
#pragma unroll
    for (int ti = 0; ti < TPW; ++ti) {
      if (my_tile[ti] >= a.ntiles) continue;
      const int x0 = (my_tile[ti] % a.tiles_x) * TG_TILE, y0 = (my_tile[ti] / a.tiles_x) * TG_TILE;
      const float x_last = (float)min(x0 + TG_TILE - 1, a.Tw - 1), y_last = (float)min(y0 + TG_TILE - 1, a.Th - 1);
      for (int nl = n_phase; nl < nstage && n0 + nl < n_end; nl += n_step) {
        const int n = n0 + nl;
        const float bx = (1.0f - pbase[2 * n + 1]) * a.kx, by = (1.0f - pbase[2 * n]) * a.ky;
        // queries whose footprint can meet the tile: x0 - 1 < ix < x_last + 1 (one cell of slack for rounding)
        int c_lo = (int)floorf(((float)x0 - 1.0f - bx) * inv_ax) - 1, c_hi = (int)ceilf((x_last + 1.0f - bx) * inv_ax) + 1;
        int r_lo = (int)floorf(((float)y0 - 1.0f - by) * inv_ay) - 1, r_hi = (int)ceilf((y_last + 1.0f - by) * inv_ay) + 1;
        c_lo = max(c_lo, 0); c_hi = min(c_hi, a.W - 1);
        r_lo = max(r_lo, r_base); r_hi = min(r_hi, r_base + rows - 1);
        if (c_lo > c_hi || r_lo > r_hi) continue;
        const int c16_lo = c_lo & ~15, c16_hi = c_hi | 15;
        const int rl_lo = (r_lo - r_base) & ~15, rl_hi = (r_hi - r_base) | 15;
        const bf16* tl = tile + (long long)nl * a.rows_chunk * a.pitch;
        // rows_chunk <= 32: at most two 16-row chunks per CTA.  The hat operands of the first GEMM depend on
        // (x, c) only, so they are generated once per 16-query chunk and used for both row chunks.
        const bool use_r[2] = {rl_lo == 0, rl_hi >= 16};
        float e[2][2][2][4];
#pragma unroll
        for (int rc = 0; rc < 2; ++rc)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int q = 0; q < 4; ++q) e[rc][mt][nt][q] = 0.f;
        for (int c0 = c16_lo; c0 < c16_hi; c0 += 16) {
          // A fragments: hat(ix(c) - x), x = x0 + mt * 16 + {gq, gq + 8}, c = c0 + {2t, 2t+1, 2t+8, 2t+9}
          const float i0 = fmaf((float)(c0 + 2 * t), a.ax, bx), i1 = i0 + a.ax;
          const float i2 = fmaf((float)(c0 + 2 * t + 8), a.ax, bx), i3 = i2 + a.ax;
          uint32_t af[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float xa = (float)(x0 + mt * 16 + gq), xb = xa + 8.0f;
            af[mt][0] = pack2(hat(i0 - xa), hat(i1 - xa));
            af[mt][1] = pack2(hat(i0 - xb), hat(i1 - xb));
            af[mt][2] = pack2(hat(i2 - xa), hat(i3 - xa));
            af[mt][3] = pack2(hat(i2 - xb), hat(i3 - xb));
          }
#pragma unroll
          for (int rc = 0; rc < 2; ++rc) {
            if (!use_r[rc]) continue;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const bf16* rowp = tl + (long long)(rc * 16 + nt * 8 + gq) * a.pitch + c0 + 2 * t;
              const uint32_t b0 = *reinterpret_cast<const uint32_t*>(rowp);
              const uint32_t b1 = *reinterpret_cast<const uint32_t*>(rowp + 8);
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) mma16816(e[rc][mt][nt], af[mt], b0, b1);
            }
          }
        }
#pragma unroll
        for (int rc = 0; rc < 2; ++rc) {
          if (!use_r[rc]) continue;
          // E^T accumulators -> A fragments of the second GEMM (contraction over the 16 rows r)
          uint32_t ef[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            ef[mt][0] = pack2(e[rc][mt][0][0], e[rc][mt][0][1]);
            ef[mt][1] = pack2(e[rc][mt][0][2], e[rc][mt][0][3]);
            ef[mt][2] = pack2(e[rc][mt][1][0], e[rc][mt][1][1]);
            ef[mt][3] = pack2(e[rc][mt][1][2], e[rc][mt][1][3]);
          }
          const float j0 = fmaf((float)(r_base + rc * 16 + 2 * t), a.ay, by), j1 = j0 + a.ay;
          const float j2 = fmaf((float)(r_base + rc * 16 + 2 * t + 8), a.ay, by), j3 = j2 + a.ay;
#pragma unroll
          for (int yt = 0; yt < 4; ++yt) {
            const float y = (float)(y0 + yt * 8 + gq);
            const uint32_t b0 = pack2(hat(j0 - y), hat(j1 - y)), b1 = pack2(hat(j2 - y), hat(j3 - y));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma16816(acc[ti][mt][yt], ef[mt], b0, b1);
          }
        }
      }
    }
  }
  // CTA reduction of the tile accumulators in shared memory, then one global atomic per touched cell
  __syncthreads();
  for (int i = threadIdx.x; i < a.Th * a.Tw; i += TG_THREADS) sred[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int ti = 0; ti < TPW; ++ti) {
    if (my_tile[ti] >= a.ntiles) continue;
    const int x0 = (my_tile[ti] % a.tiles_x) * TG_TILE, y0 = (my_tile[ti] / a.tiles_x) * TG_TILE;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int yt = 0; yt < 4; ++yt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int x = x0 + mt * 16 + gq + (q >> 1) * 8, y = y0 + yt * 8 + 2 * t + (q & 1);
          const float v = acc[ti][mt][yt][q];
          if (x < a.Tw && y < a.Th && v != 0.f) atomicAdd(&sred[y * a.Tw + x], v);
        }
  }
  __syncthreads();
  float* dt_g = d_table + (long long)eta * a.Th * a.Tw;
  for (int i = threadIdx.x; i < a.Th * a.Tw; i += TG_THREADS) {
    const float v = sred[i];
    if (v != 0.f) atomicAdd(dt_g + i, v);
  }
}

struct TgPlan {
  TgArgs a;
  dim3 grid;
  size_t smem;
  int tpw;
};
bool make_plan(const Shape& s, TgPlan* p) {
  if (s.W % 16 != 0 || s.H % 16 != 0 || s.Ns % TG_NB != 0 || s.W < 2 || s.H < 2) return false;
  TgArgs& a = p->a;
  a.H = s.H; a.W = s.W; a.HW = s.HW; a.heads = s.heads; a.G = s.G; a.hg = s.hg; a.Ns = s.Ns; a.Th = s.Th; a.Tw = s.Tw;
  a.tiles_x = (s.Tw + TG_TILE - 1) / TG_TILE;
  const int tiles_y = (s.Th + TG_TILE - 1) / TG_TILE;
  a.ntiles = a.tiles_x * tiles_y;
  if (a.ntiles > 16) return false;
  if (a.ntiles < 8 && (8 % a.ntiles) != 0) return false;     // warps share tiles evenly: 1, 2, 4 or >= 8 tiles
  p->tpw = a.ntiles > 8 ? 2 : 1;
  a.pitch = s.W + 8;
  // rows per CTA: 32 (two 16-row chunks, see the kernel) - the staged block [8][rows][W + 8] bf16 is <= 70 KB
  const int rows = s.H < 32 ? s.H : 32;
  if (s.W > 128) return false;
  a.rows_chunk = rows;
  const int r_chunks = (s.H + rows - 1) / rows;
  // 8-sample groups per staging step: as many as fit ~64 KB (small maps: fewer, longer steps)
  a.ng = (int)((64 * 1024) / ((size_t)TG_NB * rows * a.pitch * 2));
  if (a.ng < 1) a.ng = 1;
  if (a.ng > 8) a.ng = 8;
  // sample split: enough CTAs for ~3 per SM
  const long long base = (long long)s.B * s.heads * r_chunks;
  int n_split = (int)((3 * 148 + base - 1) / base);
  const int max_split = s.Ns / (TG_NB * a.ng) > 0 ? s.Ns / (TG_NB * a.ng) : 1;
  if (n_split > max_split) n_split = max_split;
  if (n_split < 1) n_split = 1;
  a.n_per_cta = ((s.Ns + n_split - 1) / n_split + TG_NB - 1) / TG_NB * TG_NB;
  n_split = (s.Ns + a.n_per_cta - 1) / a.n_per_cta;
  a.ax = 0.5f * (float)(s.Tw - 1) / (float)(s.W - 1);
  a.ay = 0.5f * (float)(s.Th - 1) / (float)(s.H - 1);
  a.kx = 0.25f * (float)(s.Tw - 1);
  a.ky = 0.25f * (float)(s.Th - 1);
  if (a.ax <= 0.f || a.ay <= 0.f) return false;
  p->grid = dim3(n_split, r_chunks, s.B * s.heads);
  size_t tile_b = (size_t)a.ng * TG_NB * rows * a.pitch * 2, red_b = (size_t)s.Th * s.Tw * 4;
  p->smem = tile_b > red_b ? tile_b : red_b;
  return p->smem <= 200 * 1024;
}

}  // namespace

bool rpe_table_grad_mma_supported(const Shape& s) {
  TgPlan p;
  return s.pe_mode == DAT_PE_RPE && make_plan(s, &p);
}

// d_table (heads, Th, Tw) fp32 is overwritten.  ds: bf16 [B * heads][Ns / 8][HW][8].
int rpe_table_grad_mma(const Shape& s, const void* ds, const float* pos, float* d_table, cudaStream_t st) {
  TgPlan p;
  DAT_REQUIRE(make_plan(s, &p), "rpe_table_grad_mma: unsupported shape");
  DAT_CUDA_OK(cudaMemsetAsync(d_table, 0, (size_t)s.heads * s.Th * s.Tw * 4, st));
  if (p.tpw == 1) {
    if (p.smem > 48 * 1024)
      DAT_CUDA_OK(cudaFuncSetAttribute(rpe_table_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    rpe_table_grad_kernel<1><<<p.grid, TG_THREADS, p.smem, st>>>((const bf16*)ds, pos, d_table, p.a);
  } else {
    if (p.smem > 48 * 1024)
      DAT_CUDA_OK(cudaFuncSetAttribute(rpe_table_grad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    rpe_table_grad_kernel<2><<<p.grid, TG_THREADS, p.smem, st>>>((const bf16*)ds, pos, d_table, p.a);
  }
  DAT_LAUNCH_OK("rpe_table_grad_kernel");
  return DAT_OK;
}

}  // namespace dat
