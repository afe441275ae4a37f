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

constexpr int TG_MAXN = 64;       // samples per staging step (8 groups of 8) at most
constexpr int TG_NB = 8;          // samples staged per step (one 16-byte load per query)
constexpr int TG_TILE = 32;       // table tile edge per warp

struct TgArgs {
  int H, W, HW, heads, G, hg, Ns, Th, Tw;
  int rows_chunk, n_per_cta, tiles_x, ntiles, pitch, ng;   // ng: 8-sample groups staged per step
  int log2w;                                               // W is a power of two (no divisions in the staging loop)
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

// NT threads = NW warps.  TPW: table tiles per warp (1: at most NW tiles, warps share tiles and split the
// samples; 2: up to 2 NW tiles)
template <int TPW, int NT>
__global__ void __launch_bounds__(NT)
rpe_table_grad_kernel(const bf16* __restrict__ ds, const float* __restrict__ pos, float* __restrict__ part,
                      TgArgs a) {
  pdl_enter();
  constexpr int NW = NT / 32;
  __shared__ float s_bx[TG_MAXN], s_by[TG_MAXN];                // (1 - pos) * k of the staged samples
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
    const int share = a.ntiles >= NW ? 1 : NW / a.ntiles;        // warps per tile (ntiles divides NW)
    my_tile[0] = a.ntiles >= NW ? warp : warp % a.ntiles;
    n_phase = a.ntiles >= NW ? 0 : warp / a.ntiles;
    n_step = share;
  } else {
#pragma unroll
    for (int i = 0; i < TPW; ++i) my_tile[i] = warp + NW * i;    // may be >= ntiles: idle slot
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
    if ((int)threadIdx.x < nstage && n0 + (int)threadIdx.x < n_end) {
      s_bx[threadIdx.x] = (1.0f - pbase[2 * (n0 + threadIdx.x) + 1]) * a.kx;
      s_by[threadIdx.x] = (1.0f - pbase[2 * (n0 + threadIdx.x)]) * a.ky;
    }
    // 4 independent 16-byte loads in flight per thread, then their 32 two-byte transposing stores
    // (n_rows_px is a multiple of 256, so a group never straddles a thread's batch unevenly)
    const int plane = a.rows_chunk * a.pitch;                      // elements per staged sample
    for (int gi = 0; gi < groups; ++gi) {
      const bf16* src = dsb + ((long long)(n0 / TG_NB + gi) * a.HW) * TG_NB;
      bf16* dstg = tile + gi * TG_NB * plane;
      for (int m0 = threadIdx.x; m0 < n_rows_px; m0 += 4 * NT) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = m0 + u * NT;
          v[u] = m < n_rows_px ? *reinterpret_cast<const uint4*>(src + m * TG_NB) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = m0 + u * NT;
          if (m < n_rows_px) {
            uint16_t* dst = reinterpret_cast<uint16_t*>(dstg + (m >> a.log2w) * a.pitch + (m & (a.W - 1)));
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              dst[(2 * i) * plane] = (uint16_t)(w[i] & 0xffffu);
              dst[(2 * i + 1) * plane] = (uint16_t)(w[i] >> 16);
            }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int ti = 0; ti < TPW; ++ti) {
      if (my_tile[ti] >= a.ntiles) continue;
      const int x0 = (my_tile[ti] % a.tiles_x) * TG_TILE, y0 = (my_tile[ti] / a.tiles_x) * TG_TILE;
      const float x_last = (float)min(x0 + TG_TILE - 1, a.Tw - 1), y_last = (float)min(y0 + TG_TILE - 1, a.Th - 1);
      // integer -> float conversions hoisted out of the sample loop
      const float xlo_f = (float)x0 - 1.0f, ylo_f = (float)y0 - 1.0f;
      const float xa_f[2] = {(float)(x0 + gq), (float)(x0 + 16 + gq)};
      float y_f[4];
#pragma unroll
      for (int yt = 0; yt < 4; ++yt) y_f[yt] = (float)(y0 + yt * 8 + gq);
      const float r_f = (float)(r_base + 2 * t);
      for (int nl = n_phase; nl < nstage && n0 + nl < n_end; nl += n_step) {
        const float bx = s_bx[nl], by = s_by[nl];
        // queries whose footprint can meet the tile: x0 - 1 < ix < x_last + 1 (one cell of slack for rounding)
        int c_lo = (int)floorf((xlo_f - bx) * inv_ax) - 1, c_hi = (int)ceilf((x_last + 1.0f - bx) * inv_ax) + 1;
        int r_lo = (int)floorf((ylo_f - by) * inv_ay) - 1, r_hi = (int)ceilf((y_last + 1.0f - by) * inv_ay) + 1;
        c_lo = max(c_lo, 0); c_hi = min(c_hi, a.W - 1);
        r_lo = max(r_lo, r_base); r_hi = min(r_hi, r_base + rows - 1);
        if (c_lo > c_hi || r_lo > r_hi) continue;
        const int c16_lo = c_lo & ~15, c16_hi = c_hi | 15;
        const int rl_lo = (r_lo - r_base) & ~15, rl_hi = (r_hi - r_base) | 15;
        const bf16* tl = tile + nl * (a.rows_chunk * a.pitch) + gq * a.pitch + 2 * t;   // + row * pitch + c0
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
        float c_f = (float)(c16_lo + 2 * t);
        const float ax8 = 8.0f * a.ax;
        for (int c0 = c16_lo; c0 < c16_hi; c0 += 16, c_f += 16.0f) {
          // A fragments: hat(ix(c) - x), x = x0 + mt * 16 + {gq, gq + 8}, c = c0 + {2t, 2t+1, 2t+8, 2t+9}
          const float i0 = fmaf(c_f, a.ax, bx), i1 = i0 + a.ax;
          const float i2 = i0 + ax8, i3 = i2 + a.ax;
          uint32_t af[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float xa = xa_f[mt], xb = xa + 8.0f;
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
              const bf16* rowp = tl + (rc * 16 + nt * 8) * a.pitch + c0;
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
          const float j0 = fmaf(r_f + (float)(rc * 16), a.ay, by), j1 = j0 + a.ay;
          const float j2 = fmaf(r_f + (float)(rc * 16 + 8), a.ay, by), j3 = j2 + a.ay;
#pragma unroll
          for (int yt = 0; yt < 4; ++yt) {
            const float y = y_f[yt];
            const uint32_t b0 = pack2(hat(j0 - y), hat(j1 - y)), b1 = pack2(hat(j2 - y), hat(j3 - y));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma16816(acc[ti][mt][yt], ef[mt], b0, b1);
          }
        }
      }
    }
  }
  // CTA reduction of the tile accumulators in shared memory in a fixed order (warps that share a tile add
  // their sample phases one after the other; within a phase every cell has exactly one owner), then the
  // CTA's partial table goes to its own slot of `part`; tg_reduce_kernel sums the slots in a fixed order.
  // No atomics anywhere: d rpe_table is bit-reproducible.
  __syncthreads();
  for (int i = threadIdx.x; i < a.Th * a.Tw; i += NT) sred[i] = 0.f;
  __syncthreads();
  for (int ph = 0; ph < n_step; ++ph) {
    if (ph == n_phase) {
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
              if (x < a.Tw && y < a.Th) sred[y * a.Tw + x] += acc[ti][mt][yt][q];
            }
      }
    }
    if (n_step > 1) __syncthreads();
  }
  __syncthreads();
  const long long slot = ((long long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  float* dst = part + (slot * a.heads + eta) * (long long)(a.Th * a.Tw);
  for (int i = threadIdx.x; i < a.Th * a.Tw; i += NT) dst[i] = sred[i];
}

// out[i] = sum_z part[z][i] in a fixed order: thread (e, zy) sums the slots z = zy, zy + 4, ... (4 loads in
// flight), the four partial sums are combined in shared memory in ascending zy.
__global__ void __launch_bounds__(256)
tg_reduce_kernel(const float* __restrict__ part, int nslots, int count, float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[4][64];
  const int e = blockIdx.x * 64 + (threadIdx.x & 63), zy = threadIdx.x >> 6;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (e < count) {
    int z = zy;
    for (; z + 12 < nslots; z += 16) {
      s0 += part[(long long)z * count + e];
      s1 += part[(long long)(z + 4) * count + e];
      s2 += part[(long long)(z + 8) * count + e];
      s3 += part[(long long)(z + 12) * count + e];
    }
    for (; z < nslots; z += 4) s0 += part[(long long)z * count + e];
  }
  red[zy][threadIdx.x & 63] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (zy == 0 && e < count) out[e] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

struct TgPlan {
  TgArgs a;
  dim3 grid;
  size_t smem;
  int tpw, threads;
};
bool make_plan(const Shape& s, TgPlan* p) {
  if (s.W % 16 != 0 || s.H % 16 != 0 || s.Ns % TG_NB != 0 || s.W < 2 || s.H < 2) return false;
  TgArgs& a = p->a;
  a.H = s.H; a.W = s.W; a.HW = s.HW; a.heads = s.heads; a.G = s.G; a.hg = s.hg; a.Ns = s.Ns; a.Th = s.Th; a.Tw = s.Tw;
  a.tiles_x = (s.Tw + TG_TILE - 1) / TG_TILE;
  const int tiles_y = (s.Th + TG_TILE - 1) / TG_TILE;
  a.ntiles = a.tiles_x * tiles_y;
  if (a.ntiles > 32) return false;
  // 8 warps per CTA, 16 for large tables (one 32 x 32 tile per warp up to 16 tiles, two beyond)
  p->threads = a.ntiles > 8 ? 512 : 256;
  const int nw = p->threads / 32;
  if (a.ntiles < nw && (nw % a.ntiles) != 0) return false;   // warps share tiles evenly: 1, 2, 4 or 8 tiles
  p->tpw = a.ntiles > nw ? 2 : 1;
  a.pitch = s.W + 8;
  // rows per CTA: 32 (two 16-row chunks, see the kernel) - the staged block [8][rows][W + 8] bf16 is <= 70 KB
  const int rows = s.H < 32 ? s.H : 32;
  if (s.W > 128 || (s.W & (s.W - 1)) != 0) return false;
  a.log2w = 0;
  while ((1 << a.log2w) < s.W) ++a.log2w;
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

size_t rpe_table_grad_mma_workspace(const Shape& s) {
  TgPlan p;
  if (s.pe_mode != DAT_PE_RPE || !make_plan(s, &p)) return 0;
  return align_up((size_t)p.grid.x * p.grid.y * s.B * s.heads * s.Th * s.Tw * 4, 256);
}

// d_table (heads, Th, Tw) fp32 is overwritten.  ds: bf16 [B * heads][Ns / 8][HW][8].  ws: one partial table per
// CTA (rpe_table_grad_mma_workspace bytes), summed in a fixed order by tg_reduce_kernel (deterministic).
int rpe_table_grad_mma(const Shape& s, const void* ds, const float* pos, float* d_table, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  TgPlan p;
  DAT_REQUIRE(make_plan(s, &p), "rpe_table_grad_mma: unsupported shape");
  DAT_REQUIRE(ws != nullptr && ws_bytes >= rpe_table_grad_mma_workspace(s), "rpe_table_grad_mma: workspace too small");
#define TG_LAUNCH(TPWV, NTV)                                                                                          \
  do {                                                                                                                \
    auto kern = rpe_table_grad_kernel<TPWV, NTV>;                                                                     \
    if (p.smem > 40 * 1024)   /* + 512 B static: stay clear of the 48 KB default limit */                           \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));              \
    launch_k(kern, p.grid, NTV, p.smem, st, (const bf16*)ds, pos, (float*)ws, p.a);                                         \
  } while (0)
  if (p.threads == 256) TG_LAUNCH(1, 256);
  else if (p.tpw == 1) TG_LAUNCH(1, 512);
  else TG_LAUNCH(2, 512);
#undef TG_LAUNCH
  DAT_LAUNCH_OK("rpe_table_grad_kernel");
  const int count = s.heads * s.Th * s.Tw;
  const int nslots = (int)(p.grid.x * p.grid.y) * s.B;
  launch_k(tg_reduce_kernel, ceil_div(count, 64), 256, 0, st, (const float*)ws, nslots, count, d_table);
  DAT_LAUNCH_OK("tg_reduce_kernel");
  return DAT_OK;
}

}  // namespace dat
