// Fused attention core of the block on the tcgen05 tensor cores (bf16 hot path):
//   S = Q K^T (tcgen05, fp32 in tensor memory)  ->  + bilinear rpe bias, softmax (CUDA cores)
//   ->  P (bf16, shared memory)  ->  O = P V (tcgen05)  ->  O / rowsum -> bf16, channel-last
// (dat_blocks.py:180-223; bias branch :198-214, q grid :123-136).
//
// One CTA owns one (batch, head) and walks a strided list of 128-query tiles.  K, V
// (Ns x 32 bf16 each) are TMA-loaded once per CTA; Q tiles are TMA double-buffered.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2-3 idle,
// 4-19 = softmax / epilogue (thread = one query row x a quarter of the Ns columns; the
// score row lives in registers between the two softmax passes; 16 warps = 4 per scheduler
// hide the LDS / MUFU latencies of the scalar work).
//
// The kernel is bound by the per-score scalar work, not by the tensor pipe (128 MMA FLOP
// vs ~25 CUDA-core instructions per score), so that work is cut to the bone:
//  * the displacement is separable, ix depends on (query column, n), iy on (query row, n):
//    the y part (row offset into the table, fraction) is tabulated once per tile per image
//    row in shared memory; the x part is one FADD from a per-n constant;
//  * floor() is the magic-number trick (no F2I / MUFU), the 4 taps are ONE 8-byte LDS from
//    a zero-padded table packed as bf16 {t00, t01-t00, t10, t11-t10} (pre-multiplied by
//    log2 e), so zero padding needs no bounds checks and the blend is 3 FMA + 1 FADD;
//  * scale * log2(e) is folded into the S -> exp2 FMA; normalisation is deferred to O.
// No score / bias / displacement tensor ever reaches HBM: algorithmic bytes per launch =
// B*HW*C*2 (q) + B*HW*C*2 (o) + 2*B*Ns*C*2 (k, v) + pos + table + lse.
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace dat {

namespace {

using namespace tc;

constexpr int TQ = 128;             // queries per tile (UMMA M)
constexpr int ATC_THREADS = 640;    // 4 control warps + 16 softmax warps
constexpr int SOFT_THREADS = 512;
constexpr int NPART = 4;            // column parts per query row (one per softmax warp group)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float MAGIC = 12582912.0f;           // 1.5 * 2^23: float -> integer rounding trick
constexpr int MAGIC_BITS = 0x4B400000;

struct AtcArgs {
  int B, H, W, HW, C, heads, G, hg, Th, Tw, Wp, Hp;
  int n_tiles, rows_max;
  float c1;        // hc^-0.5 * log2(e)
  float kx, ky;    // 0.25 * (Tw - 1), 0.25 * (Th - 1)
  float gsx, gsy;  // 2 / (W - 1), 2 / (H - 1): query grid step (no IEEE division in the kernel)
  // split-KV launches (Ns > 256, e.g. the 512 x 2048 evaluation shape: 16 x 64 = 1024 samples): CTA z handles
  // the samples [n_off0 + z * NS, + NS) of ns_total and writes its normalised partial output / log-sum-exp into
  // slot z0 + z (strides o_zstride / lse_zstride elements); attn_combine_kernel merges the slots.
  int ns_total, n_off0, z0;
  int force_xclamp;   // debug: -1 = per-CTA choice, 0 / 1 = force the unclamped / clamped pass (DAT_B200_ATTN_XCLAMP)
  long long o_zstride, lse_zstride;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void soft_bar_sync() {   // the 256 softmax threads only
  asm volatile("bar.sync 1, 512;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// rpe table -> zero-padded, tap-packed, log2(e)-scaled copy:
// entry (y, x), y in [-2, Th], x in [-2, Tw]  at  (y + 2) * Wp + (x + 2):
//   .x = bf16x2 {T[y][x],   T[y][x+1]   - T[y][x]}      .y = bf16x2 {T[y+1][x], T[y+1][x+1] - T[y+1][x]}
__global__ void pack_table_kernel(const float* __restrict__ table, uint2* __restrict__ out,
                                  int heads, int Th, int Tw) {
  pdl_enter();
  const int Wp = Tw + 3, Hp = Th + 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * Hp * Wp) return;
  const int eta = idx / (Hp * Wp), rem = idx % (Hp * Wp);
  const int y = rem / Wp - 2, x = rem % Wp - 2;
  const float* t = table + (long long)eta * Th * Tw;
  auto at = [&](int yy, int xx) {
    return (yy >= 0 && yy < Th && xx >= 0 && xx < Tw) ? t[yy * Tw + xx] * LOG2E : 0.f;
  };
  const float t00 = at(y, x), t01 = at(y, x + 1), t10 = at(y + 1, x), t11 = at(y + 1, x + 1);
  out[idx] = make_uint2(pack_bf16x2(t00, t01 - t00), pack_bf16x2(t10, t11 - t10));
}

struct SmemPlan {
  uint32_t q[2], k, v, p, tab, yt, xk, yk, red, bars, total;
};
__host__ __device__ inline SmemPlan plan_smem(int NS, int Hp, int Wp, int rows_max) {
  SmemPlan s;
  uint32_t off = 0;
  s.q[0] = off; off += TQ * 64;
  s.q[1] = off; off += TQ * 64;
  s.k = off; off += NS * 64;
  s.v = off; off += NS * 64;
  s.p = off; off += (NS / 64) * 16384;
  s.tab = off; off += ((uint32_t)(Hp * Wp) * 8 + 15) & ~15u;
  s.yt = off; off += (uint32_t)rows_max * NS * 8;
  s.xk = off; off += NS * 4;
  s.yk = off; off += NS * 4;
  s.red = off; off += 2 * NPART * TQ * 4;
  s.bars = off; off += 16 * 8;
  s.total = off + 1024;   // slack for the manual 1024-byte alignment
  return s;
}

template <int NS>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const float* __restrict__ pos,
                   const uint2* __restrict__ tab_packed, bf16* __restrict__ o,
                   float* __restrict__ lse, AtcArgs a) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  const SmemPlan sp = plan_smem(NS, a.Hp, a.Wp, a.rows_max);
  uint8_t* sQ0 = smem + sp.q[0];
  uint8_t* sK = smem + sp.k;
  uint8_t* sV = smem + sp.v;
  uint8_t* sP = smem + sp.p;
  uint2* sTab = reinterpret_cast<uint2*>(smem + sp.tab);
  int2* sYt = reinterpret_cast<int2*>(smem + sp.yt);
  float* sXk = reinterpret_cast<float*>(smem + sp.xk);
  float* sYk = reinterpret_cast<float*>(smem + sp.yk);
  float* sMax = reinterpret_cast<float*>(smem + sp.red);      // [NPART][128]
  float* sSum = sMax + NPART * TQ;                            // [NPART][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // [2]
  uint64_t* q_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* s_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  constexpr uint32_t TMEM_COLS = NS < 32 ? 32 : NS;   // NS is a power of two here
  const int n_off = a.n_off0 + (int)blockIdx.z * NS;   // first sample of this CTA's KV chunk
  o += (long long)(a.z0 + (int)blockIdx.z) * a.o_zstride;
  lse += (long long)(a.z0 + (int)blockIdx.z) * a.lse_zstride;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(&q_empty[0], 1);
    mbar_init(&q_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, SOFT_THREADS);
    mbar_init(o_full, 1);
    mbar_init(s_free, SOFT_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  {
    const uint2* src = tab_packed + (long long)eta * a.Hp * a.Wp;
    for (int i = threadIdx.x; i < a.Hp * a.Wp; i += ATC_THREADS) sTab[i] = src[i];
    const float* pp = pos + (((long long)b * a.G + g) * a.ns_total + n_off) * 2;
    for (int n = threadIdx.x; n < NS; n += ATC_THREADS) {
      sYk[n] = pp[2 * n] * a.ky;
      sXk[n] = pp[2 * n + 1] * a.kx;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      // ---- TMA producer ------------------------------------------------------------
      mbar_arrive_expect_tx(kv_full, 2u * NS * 64u);
      tma_load_2d(sK, &tmK, kv_full, eta * 32, b * a.ns_total + n_off);
      tma_load_2d(sV, &tmV, kv_full, eta * 32, b * a.ns_total + n_off);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int slot = it & 1;
        mbar_wait(&q_empty[slot], (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&q_full[slot], TQ * 64u);
        tma_load_2d(sQ0 + slot * (TQ * 64), &tmQ, &q_full[slot], eta * 32, b * a.HW + tile * TQ);
      }
    } else if (warp == 1 && lane == 0) {
      // ---- MMA issuer ---------------------------------------------------------------
      const uint32_t idesc_s = make_instr_desc(FMT_BF16, TQ, NS);
      const uint32_t idesc_o = make_instr_desc(FMT_BF16, TQ, 32, 0, 1);   // B (= V) is MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      mbar_wait(kv_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int slot = it & 1;
        mbar_wait(&q_full[slot], ((uint32_t)it >> 1) & 1u);
        if (it > 0) mbar_wait(s_free, (uint32_t)(it - 1) & 1u);
        tc_fence_after_sync();
        const uint32_t q_addr = smem_u32(sQ0 + slot * (TQ * 64));
#pragma unroll
        for (int k = 0; k < 2; ++k) {     // head dim 32 = 2 x K16
          const uint64_t ad = make_smem_desc(q_addr + k * 32, 16, 512, LAYOUT_SW64);
          const uint64_t bd = make_smem_desc(k_addr + k * 32, 16, 512, LAYOUT_SW64);
          mma_bf16_ss(tmem_base, ad, bd, idesc_s, (uint32_t)k);
        }
        tc_commit(&q_empty[slot]);
        tc_commit(s_full);
        mbar_wait(p_ready, (uint32_t)it & 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int kk = 0; kk < NS / 16; ++kk) {
          const uint64_t ad = make_smem_desc(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, LAYOUT_SW128);
          const uint64_t bd = make_smem_desc(v_addr + kk * 1024, 512, 512, LAYOUT_SW64);
          mma_bf16_ss(tmem_base, ad, bd, idesc_o, (uint32_t)(kk > 0));
        }
        tc_commit(o_full);
      }
    }
  } else {
    // ---- softmax + epilogue -------------------------------------------------------------
    const int quad = warp & 3, half = (warp - 4) >> 2;   // `half` = column part 0..NPART-1
    const int row = quad * 32 + lane;
    const int stid = threadIdx.x - 128;
    constexpr int NH = NS / NPART;                   // columns per thread
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int m = tile * TQ + row;
      const bool valid = m < a.HW;
      const int mm = valid ? m : a.HW - 1;
      const int r = mm / a.W, c = mm - r * a.W;
      const int r0 = (tile * TQ) / a.W;
      const int r_last = min(a.HW - 1, tile * TQ + TQ - 1) / a.W;
      // y part of the bias footprint, once per (image row of the tile, n)
      for (int e = stid; e < (r_last - r0 + 1) * NS; e += SOFT_THREADS) {
        const int rr = e / NS, n = e - rr * NS;
        const float gy = fmaf((float)(r0 + rr), a.gsy, -1.0f);
        const float ay = (gy * 0.25f + 0.5f) * (float)(a.Th - 1) - 0.5f;
        float u = ay - sYk[n];
        u = fminf(fmaxf(u, -1.5f), (float)a.Th - 0.5f);
        const float aa = u + MAGIC;
        const float fy = (u - (aa - MAGIC)) + 0.5f;
        const int y0 = __float_as_int(aa) - MAGIC_BITS;            // in [-2, Th]
        sYt[e] = make_int2((y0 + 2) * a.Wp + 2 - MAGIC_BITS, __float_as_int(fy));
      }
      soft_bar_sync();
      const float ax = (fmaf((float)c, a.gsx, -1.0f) * 0.25f + 0.5f) * (float)(a.Tw - 1) - 0.5f;
      const float xhi = (float)a.Tw - 0.5f;

      mbar_wait(s_full, (uint32_t)it & 1u);
      tc_fence_after_sync();
      uint32_t t[NH];
      if constexpr (NH >= 32) {
#pragma unroll
        for (int c4 = 0; c4 < NH / 32; ++c4)
          tmem_ld_32x32(t_lane + (uint32_t)(half * NH + c4 * 32),
                        *reinterpret_cast<uint32_t(*)[32]>(&t[c4 * 32]));
      } else {
        tmem_ld_32x16(t_lane + (uint32_t)(half * NH), *reinterpret_cast<uint32_t(*)[16]>(&t[0]));
      }
      tmem_wait_ld();

      // pass 1: t = s * scale*log2e + bias*log2e, running max
      const int2* yt = sYt + (r - r0) * NS + half * NH;
      const float* xk = sXk + half * NH;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        const int2 ye = yt[j];
        float u = ax - xk[j];
        u = fminf(fmaxf(u, -1.5f), xhi);
        const float aa = u + MAGIC;
        const float fx = (u - (aa - MAGIC)) + 0.5f;
        const uint2 e = sTab[ye.x + __float_as_int(aa)];
        const float t00 = __uint_as_float(e.x << 16), d0 = __uint_as_float(e.x & 0xffff0000u);
        const float t10 = __uint_as_float(e.y << 16), d1 = __uint_as_float(e.y & 0xffff0000u);
        const float top = fmaf(fx, d0, t00), bot = fmaf(fx, d1, t10);
        const float bias = fmaf(__int_as_float(ye.y), bot - top, top);
        const float val = fmaf(__uint_as_float(t[j]), a.c1, bias);
        t[j] = __float_as_uint(val);
        mx = fmaxf(mx, val);
      }
      sMax[half * TQ + row] = mx;
      soft_bar_sync();
#pragma unroll
      for (int pp = 0; pp < NPART; ++pp) mx = fmaxf(mx, sMax[pp * TQ + row]);

      // pass 2: p = 2^(t - max), row sum, bf16 -> shared memory (K-major, 128B swizzle)
      float l = 0.f;
      uint8_t* prow = sP + row * 128;
#pragma unroll
      for (int j = 0; j < NH; j += 8) {
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pv[i] = ex2(__uint_as_float(t[j + i]) - mx);
          l += pv[i];
        }
        const int n = half * NH + j;
        const int kb = n >> 6, ch = (n & 63) >> 3;
        uint4 w = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]),
                             pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
        *reinterpret_cast<uint4*>(prow + kb * 16384 + ((ch ^ (row & 7)) << 4)) = w;
      }
      sSum[half * TQ + row] = l;
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
      soft_bar_sync();
      l = 0.f;
#pragma unroll
      for (int pp = 0; pp < NPART; ++pp) l += sSum[pp * TQ + row];

      // epilogue: O (fp32, TMEM columns [0, 32) of the S buffer) / l -> bf16
      mbar_wait(o_full, (uint32_t)it & 1u);
      tc_fence_after_sync();
      uint32_t ov[16];
      if (half < 2) {                       // parts 0 and 1 write 16 output channels each
        tmem_ld_32x16(t_lane + (uint32_t)(half * 16), ov);
        tmem_wait_ld();
      }
      tc_fence_before_sync();
      mbar_arrive(s_free);
      if (valid && half < 2) {
        const float inv = __fdividef(1.0f, l);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pk[i] = pack_bf16x2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(o + ((long long)b * a.HW + m) * a.C + eta * 32 + half * 16);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (half == 0) lse[(long long)bh * a.HW + m] = (mx + __log2f(l)) * LN2;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// =====================================================================================================================
// Version 2 of the fused forward (default; DAT_B200_ATTN_V1=1 selects the kernel above).  Same mapping of the scores to
// threads, restructured so that neither the tensor pipe nor the softmax warps wait for each other and with roughly half
// the CUDA-core instructions per score:
//  * S is double-buffered in tensor memory (2 x 256 columns): QK^T of tile i + 1 is issued while the softmax warps work
//    on tile i; O = PV of tile i accumulates in columns [0, 32) of S buffer i after its scores have been read, and its
//    epilogue (normalise, store) runs in the middle of tile i + 1 - no thread ever waits for an MMA it has just caused.
//  * the row sums come out of the tensor core: one extra N = 16 MMA per K step multiplies P by a tile of ones into
//    columns [32, 48) (exactly the bf16 P that PV uses), so the softmax warps neither add nor exchange them.
//  * one bar.sync per tile (the row-max exchange); the per-(image row, sample) table of the y footprint is built one tile
//    ahead into a second buffer.
//  * per score: ONE 16-byte broadcast LDS brings (row address, y fraction as bf16x2, x constant), one LEA forms the tap
//    address from the magic-number float, ONE 8-byte LDS fetches the four taps stored as bf16 {mid, dif} of row y and
//    their differences to row y + 1, ONE HFMA2.BF16 blends in y, two shifts unpack, two FMAs blend in x and add the
//    scaled score.  Samples inside [-1, 1] (always the case when offset_range_factor < 0, dat_blocks.py:159-162) cannot
//    leave the padded table: a CTA whose samples all are runs the pass without per-score range clamps (run-time choice).
//  13 (15 with clamps) + 2.6 instructions per score instead of 22 + 3.6.
// =====================================================================================================================

constexpr int A2_THREADS = 576;     // warp 0: TMA producer, warp 1: TMEM allocator + MMA issuer, warps 2-17: softmax
constexpr int A2_SBUF = 256;        // tensor-memory columns per S buffer
constexpr int A2_ONES = 2048;       // bf16 ones read by the row-sum MMA

__device__ __forceinline__ void soft2_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ uint32_t hfma2_bf16(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

// rpe table -> zero-padded copy in (mid, dif) form, log2(e)-scaled, entry (y, x) at (y + 2) * Wp + (x + 2):
//   .x = bf16x2 {m0 = (T[y][x] + T[y][x+1]) / 2, d0 = T[y][x+1] - T[y][x]}      .y = bf16x2 {m1 - m0, d1 - d0} (row y + 1)
// bilinear value at (y + fy, x + 1/2 + fx'):  (m0 + fy (m1 - m0)) + fx' (d0 + fy (d1 - d0))
__global__ void pack_table2_kernel(const float* __restrict__ table, uint2* __restrict__ out, int heads, int Th, int Tw) {
  pdl_enter();
  const int Wp = Tw + 3, Hp = Th + 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * Hp * Wp) return;
  const int eta = idx / (Hp * Wp), rem = idx % (Hp * Wp);
  const int y = rem / Wp - 2, x = rem % Wp - 2;
  const float* t = table + (long long)eta * Th * Tw;
  auto at = [&](int yy, int xx) {
    return (yy >= 0 && yy < Th && xx >= 0 && xx < Tw) ? t[yy * Tw + xx] * LOG2E : 0.f;
  };
  const float t00 = at(y, x), t01 = at(y, x + 1), t10 = at(y + 1, x), t11 = at(y + 1, x + 1);
  const float m0 = 0.5f * (t00 + t01), d0 = t01 - t00, m1 = 0.5f * (t10 + t11), d1 = t11 - t10;
  out[idx] = make_uint2(pack_bf16x2(m0, d0), pack_bf16x2(m1 - m0, d1 - d0));
}

struct Smem2Plan {
  uint32_t q, k, v, p, ones, tab, yt[2], xk, yk, red, bars, total;
};
__host__ __device__ inline Smem2Plan plan_smem2(int NS, int Hp, int Wp, int rows_max) {
  Smem2Plan s;
  uint32_t off = 0;
  s.q = off; off += TQ * 64;
  s.k = off; off += NS * 64;
  s.v = off; off += NS * 64;
  s.p = off; off += (NS / 64) * 16384;
  s.ones = off; off += A2_ONES;
  s.tab = off; off += ((uint32_t)(Hp * Wp) * 8 + 15) & ~15u;
  s.yt[0] = off; off += (uint32_t)rows_max * NS * 16;
  s.yt[1] = off; off += (uint32_t)rows_max * NS * 16 + 64;   // + the 4-entry read-ahead of the last thread
  s.xk = off; off += NS * 4;
  s.yk = off; off += NS * 4;
  s.red = off; off += 2 * NPART * TQ * 4;   // row maxima, double-buffered by tile parity
  s.bars = off; off += 16 * 8;
  s.total = off + 1024;   // slack for the manual 1024-byte alignment
  return s;
}

// pass 1 over one thread's NH columns of the S buffer: val = s * scale*log2e + bias*log2e written back in place
// (tcgen05.st), returns the running maximum.  XCLAMP: per-score range clamps of the x footprint (needed only when some
// sample of the CTA lies outside [-1, 1]; decided per CTA at run time, so the result never depends on it).
template <int NH, bool XCLAMP>
__device__ __forceinline__ float attn2_pass1(uint32_t s_cols, const uint4* __restrict__ yt, float ax, float xhi, float c1) {
  // Software pipeline over groups of 4 scores: the per-sample parameters of group g + 1 are loaded while the four tap
  // loads of group g are in flight, and the blends of group g run after them - four independent LDS -> FADD -> LDS ->
  // HFMA2 -> FMA chains per thread instead of one or two (the profile of the plain loop was latency-bound:
  // short_scoreboard 2.7 + wait 1.9 warps per issue at 48 % issue utilisation).  The volatile tap loads keep their
  // program order, so the grouping survives the compiler's scheduling.
  constexpr int G = 4;
  float mx = -INFINITY;
  uint4 ye[G];
#pragma unroll
  for (int i = 0; i < G; ++i) ye[i] = yt[i];
#pragma unroll 1
  for (int ch = 0; ch < NH / 16; ++ch) {
    uint32_t t[16];
    tmem_ld_32x16(s_cols + (uint32_t)(ch * 16), t);
    tmem_wait_ld();
#pragma unroll
    for (int g = 0; g < 16 / G; ++g) {
      uint2 e[G];
      float fxp[G];
      uint32_t fy2[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        float u = ax - __uint_as_float(ye[i].z);
        if (XCLAMP) u = fminf(fmaxf(u, -1.5f), xhi);
        const float aa = u + MAGIC;
        e[i] = lds64(ye[i].x + (__float_as_uint(aa) << 3));
        fxp[i] = u - (aa - MAGIC);                                // x fraction - 1/2
        fy2[i] = ye[i].y;
      }
      // parameters of the next group (the last group of the last chunk reads 4 entries past the thread's range:
      // still inside the table buffer, never used)
#pragma unroll
      for (int i = 0; i < G; ++i) ye[i] = yt[ch * 16 + (g + 1) * G + i];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const uint32_t md = hfma2_bf16(fy2[i], e[i].y, e[i].x);   // blend in y: {mid, dif}
        const float bias = fmaf(fxp[i], __uint_as_float(md & 0xffff0000u), __uint_as_float(md << 16));
        const float val = fmaf(__uint_as_float(t[g * G + i]), c1, bias);
        t[g * G + i] = __float_as_uint(val);
        mx = fmaxf(mx, val);
      }
    }
    tmem_st_32x16(s_cols + (uint32_t)(ch * 16), t);
  }
  return mx;
}

template <int NS>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const float* __restrict__ pos,
                    const uint2* __restrict__ tab_packed, bf16* __restrict__ o, float* __restrict__ lse, AtcArgs a) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  const Smem2Plan sp = plan_smem2(NS, a.Hp, a.Wp, a.rows_max);
  uint8_t* sQ = smem + sp.q;
  uint8_t* sK = smem + sp.k;
  uint8_t* sV = smem + sp.v;
  uint8_t* sP = smem + sp.p;
  uint2* sTab = reinterpret_cast<uint2*>(smem + sp.tab);
  float* sXk = reinterpret_cast<float*>(smem + sp.xk);
  float* sYk = reinterpret_cast<float*>(smem + sp.yk);
  float* sMaxBase = reinterpret_cast<float*>(smem + sp.red);  // [2][NPART][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* s_full = bars + 3;    // [2]
  uint64_t* s_free = bars + 5;    // [2]
  uint64_t* p_ready = bars + 7;
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  uint32_t* oob_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  const int n_off = a.n_off0 + (int)blockIdx.z * NS;   // first sample of this CTA's KV chunk
  o += (long long)(a.z0 + (int)blockIdx.z) * a.o_zstride;
  lse += (long long)(a.z0 + (int)blockIdx.z) * a.lse_zstride;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA (>= 1)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(&s_free[0], SOFT_THREADS);
    mbar_init(&s_free[1], SOFT_THREADS);
    mbar_init(p_ready, SOFT_THREADS);
    mbar_init(o_full, 1);
    fence_barrier_init();
    *oob_flag = 0u;
  }
  __syncthreads();                    // the flag is zero before any thread may raise it
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  int oob = 0;
  {
    const uint2* src = tab_packed + (long long)eta * a.Hp * a.Wp;
    for (int i = threadIdx.x; i < a.Hp * a.Wp; i += A2_THREADS) sTab[i] = src[i];
    const float* pp = pos + (((long long)b * a.G + g) * a.ns_total + n_off) * 2;
    for (int n = threadIdx.x; n < NS; n += A2_THREADS) {
      const float py = pp[2 * n], px = pp[2 * n + 1];
      // positions inside [-1, 1] (always the case when offset_range_factor < 0, dat_blocks.py:159-162) keep every
      // tap index inside the padded table: the CTA then runs the pass without per-score range clamps
      if (!(fabsf(px) <= 1.0f)) oob = 1;
      sYk[n] = py * a.ky;
      sXk[n] = px * a.kx;
    }
    for (int i = threadIdx.x; i < A2_ONES / 4; i += A2_THREADS) reinterpret_cast<uint32_t*>(smem + sp.ones)[i] = 0x3f803f80u;
    fence_proxy_async_smem();          // the ones tile is read by the tensor core (async proxy)
  }
  tc_fence_before_sync();
  if (oob) *oob_flag = 1u;            // benign race: every writer stores the same value
  __syncthreads();
  const bool xclamp = a.force_xclamp >= 0 ? a.force_xclamp != 0 : *oob_flag != 0u;
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: K, V once; one Q slot (S is double-buffered in tensor memory, so the slot is free again
      //      long before the next tile's QK^T is due) ------------------------------------------------------------
      mbar_arrive_expect_tx(kv_full, 2u * NS * 64u);
      tma_load_2d(sK, &tmK, kv_full, eta * 32, b * a.ns_total + n_off);
      tma_load_2d(sV, &tmV, kv_full, eta * 32, b * a.ns_total + n_off);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(q_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full, TQ * 64u);
        tma_load_2d(sQ, &tmQ, q_full, eta * 32, b * a.HW + tile * TQ);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer -----------------------------------------------------------------------------------------
      const uint32_t idesc_s = make_instr_desc(FMT_BF16, TQ, NS);
      const uint32_t idesc_o = make_instr_desc(FMT_BF16, TQ, 32, 0, 1);   // B (= V) is MN-major
      const uint32_t idesc_1 = make_instr_desc(FMT_BF16, TQ, 16, 0, 1);   // B = ones
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
      const uint64_t onesd = make_smem_desc(smem_u32(smem + sp.ones), 8192, 1024, LAYOUT_SW128);
      mbar_wait(kv_full, 0);
      auto issue_s = [&](int j) {
        const int sb = j & 1;
        mbar_wait(q_full, (uint32_t)j & 1u);
        if (j >= 2) mbar_wait(&s_free[sb], ((uint32_t)(j - 2) >> 1) & 1u);   // epilogue of tile j - 2 has drained the buffer
        tc_fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)(sb * A2_SBUF);
#pragma unroll
        for (int k = 0; k < 2; ++k) {     // head dim 32 = 2 x K16
          const uint64_t ad = make_smem_desc(q_addr + k * 32, 16, 512, LAYOUT_SW64);
          const uint64_t bd = make_smem_desc(k_addr + k * 32, 16, 512, LAYOUT_SW64);
          mma_bf16_ss(d, ad, bd, idesc_s, (uint32_t)k);
        }
        tc_commit(q_empty);
        tc_commit(&s_full[sb]);
      };
      issue_s(0);
      for (int it = 0; it < n_my; ++it) {
        if (it + 1 < n_my) issue_s(it + 1);
        mbar_wait(p_ready, (uint32_t)it & 1u);
        tc_fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)((it & 1) * A2_SBUF);
#pragma unroll
        for (int kk = 0; kk < NS / 16; ++kk) {
          const uint64_t ad = make_smem_desc(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, LAYOUT_SW128);
          const uint64_t bd = make_smem_desc(v_addr + kk * 1024, 512, 512, LAYOUT_SW64);
          mma_bf16_ss(d, ad, bd, idesc_o, (uint32_t)(kk > 0));
          mma_bf16_ss(d + 32u, ad, onesd, idesc_1, (uint32_t)(kk > 0));   // row sums of P
        }
        tc_commit(o_full);
      }
    }
  } else {
    // ---- softmax + epilogue -------------------------------------------------------------------------------------
    const int quad = warp & 3, part = (warp - 2) >> 2;   // TMEM lane quadrant; column part 0..NPART-1
    const int row = quad * 32 + lane;
    const int stid = threadIdx.x - 64;
    constexpr int NH = NS / NPART;                   // columns per thread
    constexpr int LOG_NS = NS == 256 ? 8 : (NS == 128 ? 7 : 6);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tab_addr = smem_u32(sTab);
    const float xhi = (float)a.Tw - 0.5f, yhi = (float)a.Th - 0.5f;

    auto build_yt = [&](int tile, int buf) {
      uint4* dst = reinterpret_cast<uint4*>(smem + sp.yt[buf]);
      const int r0 = (tile * TQ) / a.W;
      const int r_last = min(a.HW - 1, tile * TQ + TQ - 1) / a.W;
      for (int e = stid; e < ((r_last - r0 + 1) << LOG_NS); e += SOFT_THREADS) {
        const int rr = e >> LOG_NS, n = e & (NS - 1);
        const float gy = fmaf((float)(r0 + rr), a.gsy, -1.0f);
        const float ay = (gy * 0.25f + 0.5f) * (float)(a.Th - 1) - 0.5f;
        float u = ay - sYk[n];
        u = fminf(fmaxf(u, -1.5f), yhi);
        const float aa = u + MAGIC;
        const float fy = (u - (aa - MAGIC)) + 0.5f;
        const int y0 = __float_as_int(aa) - MAGIC_BITS;            // in [-2, Th]
        const uint32_t ro8 = tab_addr + ((uint32_t)((y0 + 2) * a.Wp + 2) << 3) - ((uint32_t)MAGIC_BITS << 3);
        dst[e] = make_uint4(ro8, pack_bf16x2(fy, fy), __float_as_uint(sXk[n]), 0u);
      }
    };
    // epilogue of a finished tile: O (fp32, columns [0, 32) of its S buffer) / row sum (column 32) -> bf16
    auto epilogue = [&](int it_e, int tile_e, float mx_e) {
      mbar_wait(o_full, (uint32_t)it_e & 1u);
      tc_fence_after_sync();
      const uint32_t tb = t_lane + (uint32_t)((it_e & 1) * A2_SBUF);
      const int m = tile_e * TQ + row;
      uint32_t ov[16];
      uint32_t lv = 0u;
      if (part < 2) {                       // parts 0 and 1 write 16 output channels each
        tmem_ld_32x16(tb + (uint32_t)(part * 16), ov);
        lv = tmem_ld_32x1(tb + 32u);
        tmem_wait_ld();
      }
      tc_fence_before_sync();
      mbar_arrive(&s_free[it_e & 1]);
      if (part < 2 && m < a.HW) {
        const float l = __uint_as_float(lv);
        const float inv = __fdividef(1.0f, l);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pk[i] = pack_bf16x2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
        uint4* dst = reinterpret_cast<uint4*>(o + ((long long)b * a.HW + m) * a.C + eta * 32 + part * 16);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (part == 0) lse[(long long)bh * a.HW + m] = (mx_e + __log2f(l)) * LN2;
      }
    };

    build_yt(blockIdx.x, 0);
    soft2_bar_sync();
    float mx_prev = 0.f;
    int tile_prev = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int m = tile * TQ + row;
      const int mm = m < a.HW ? m : a.HW - 1;
      const int r = mm / a.W, c = mm - r * a.W;
      const int r0 = (tile * TQ) / a.W;
      const float ax = (fmaf((float)c, a.gsx, -1.0f) * 0.25f + 0.5f) * (float)(a.Tw - 1) - 0.5f;

      mbar_wait(&s_full[it & 1], ((uint32_t)it >> 1) & 1u);
      tc_fence_after_sync();
      // pass 1: val = s * scale*log2e + bias*log2e, running max.  The scores are processed 16 columns at a time and
      // the values go straight back into the S buffer (tcgen05.st): only 16 of them are ever live in registers, which
      // leaves the compiler room to interleave the dependent LDS -> FADD -> LDS -> HFMA2 -> FMA chains of neighbouring
      // scores (with all 64 values resident the kernel ran at the 96-register cap and issued them one after the other)
      const uint32_t s_cols = t_lane + (uint32_t)((it & 1) * A2_SBUF + part * NH);
      const uint4* yt = reinterpret_cast<const uint4*>(smem + sp.yt[it & 1]) + ((r - r0) << LOG_NS) + part * NH;
      float mx = xclamp ? attn2_pass1<NH, true>(s_cols, yt, ax, xhi, a.c1) : attn2_pass1<NH, false>(s_cols, yt, ax, xhi, a.c1);
      // y table of the next tile into the other buffer (its readers are two barriers away)
      if (tile + (int)gridDim.x < a.n_tiles) build_yt(tile + gridDim.x, (it + 1) & 1);
      // row maxima of the four column parts: the slots alternate with the tile parity, so a fast thread's write for
      // tile it + 1 cannot overtake a slow thread's read for tile it (one bar.sync per tile separates same-parity uses)
      float* sMax = sMaxBase + (it & 1) * (NPART * TQ);
      sMax[part * TQ + row] = mx;
      // the previous tile's PV is long done: its epilogue also frees its S buffer for QK^T of tile it + 1 and
      // guarantees that the tensor core has finished reading P before pass 2 overwrites it
      if (it > 0) epilogue(it - 1, tile_prev, mx_prev);
      tmem_wait_st();                 // this thread's values are back in tensor memory before pass 2 re-reads them
      soft2_bar_sync();
#pragma unroll
      for (int pp = 0; pp < NPART; ++pp) mx = fmaxf(mx, sMax[pp * TQ + row]);

      // pass 2: p = 2^(val - max), bf16 -> shared memory (K-major, 128B swizzle)
      uint8_t* prow = sP + row * 128;
#pragma unroll 1
      for (int c2 = 0; c2 < NH / 16; ++c2) {
        uint32_t t[16];
        tmem_ld_32x16(s_cols + (uint32_t)(c2 * 16), t);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pv[i] = ex2(__uint_as_float(t[j + i]) - mx);
          const int n = part * NH + c2 * 16 + j;
          const int kb = n >> 6, ch = (n & 63) >> 3;
          uint4 w = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]),
                               pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
          *reinterpret_cast<uint4*>(prow + kb * 16384 + ((ch ^ (row & 7)) << 4)) = w;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
      mx_prev = mx;
      tile_prev = tile;
    }
    epilogue(it - 1, tile_prev, mx_prev);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int rows_spanned_max(int HW, int W) {
  int n_tiles = (HW + TQ - 1) / TQ, best = 1;
  for (int t = 0; t < n_tiles; ++t) {
    int first = (t * TQ) / W;
    int last_m = t * TQ + TQ - 1;
    if (last_m > HW - 1) last_m = HW - 1;
    int rows = last_m / W - first + 1;
    if (rows > best) best = rows;
  }
  return best;
}

// number of CTAs per (batch, head): minimise waves * (tiles per CTA + prologue)
int pick_chunks(int pairs, int n_tiles) {
  double best_cost = 1e30;
  int best = 1;
  for (int ch = 1; ch <= n_tiles; ++ch) {
    long long ctas = (long long)pairs * ch;
    long long waves = (ctas + 147) / 148;
    int per = (n_tiles + ch - 1) / ch;
    double cost = (double)waves * (per + 0.6);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ch; }
  }
  return best;
}

AtcArgs make_args(const Shape& s) {
  AtcArgs a;
  a.B = s.B; a.H = s.H; a.W = s.W; a.HW = s.HW; a.C = s.C; a.heads = s.heads; a.G = s.G; a.hg = s.hg;
  a.Th = s.Th; a.Tw = s.Tw; a.Wp = s.Tw + 3; a.Hp = s.Th + 3;
  a.n_tiles = (s.HW + TQ - 1) / TQ;
  a.rows_max = rows_spanned_max(s.HW, s.W);
  a.c1 = (1.0f / sqrtf((float)DAT_HEAD_DIM)) * LOG2E;
  a.kx = 0.25f * (float)(s.Tw - 1);
  a.ky = 0.25f * (float)(s.Th - 1);
  a.gsx = 2.0f / (float)(s.W - 1);
  a.gsy = 2.0f / (float)(s.H - 1);
  a.ns_total = s.Ns; a.n_off0 = 0; a.z0 = 0; a.o_zstride = 0; a.lse_zstride = 0;
  const char* fx = getenv("DAT_B200_ATTN_XCLAMP");
  a.force_xclamp = fx != nullptr ? atoi(fx) : -1;
  return a;
}

// Merge of the split-KV partials: slot c holds softmax(S_c) V_c and lse_c of its sample chunk;
// O = sum_c w_c O_c / sum_c w_c, w_c = exp(lse_c - max lse), lse = max + log sum_c w_c.
__global__ void attn_combine_kernel(const bf16* __restrict__ o_part, const float* __restrict__ lse_part,
                                    bf16* __restrict__ o, float* __restrict__ lse, int nch, int HW, int C, int heads,
                                    long long o_zstride, long long lse_zstride, long long total) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, m, eta, 8-channel group)
  if (idx >= total) return;
  const int c8 = (int)(idx & 3);
  const int eta = (int)((idx >> 2) % heads);
  const long long bm = (idx >> 2) / heads;
  const int m = (int)(bm % HW);
  const long long b = bm / HW;
  const long long li = (b * heads + eta) * HW + m;
  float mx = -INFINITY;
  for (int c = 0; c < nch; ++c) mx = fmaxf(mx, lse_part[c * lse_zstride + li]);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float den = 0.f;
  const long long oi = bm * C + eta * 32 + c8 * 8;
  for (int c = 0; c < nch; ++c) {
    const float w = __expf(lse_part[c * lse_zstride + li] - mx);
    den += w;
    const uint4 raw = *reinterpret_cast<const uint4*>(o_part + c * o_zstride + oi);
    const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fmaf(w, __uint_as_float(rw[i] << 16), acc[2 * i]);
      acc[2 * i + 1] = fmaf(w, __uint_as_float(rw[i] & 0xffff0000u), acc[2 * i + 1]);
    }
  }
  const float inv = 1.0f / den;
  *reinterpret_cast<uint4*>(o + oi) = make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                                                  pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
  if (c8 == 0) lse[li] = mx + __logf(den);
}

constexpr int KV_MAX = 256;            // samples per CTA (one TMEM accumulator of 128 x 256 fp32)
constexpr int KV_CHUNKS_MAX = 32;      // Ns <= 8192
int kv_chunks(int Ns) { return Ns <= KV_MAX ? 1 : Ns / KV_MAX + ((Ns % KV_MAX) & 128 ? 1 : 0) + ((Ns % KV_MAX) & 64 ? 1 : 0); }
size_t table_bytes(const Shape& s) { return align_up((size_t)s.heads * (s.Th + 3) * (s.Tw + 3) * 8, 256); }

}  // namespace

// version 2 (double-buffered S, tensor-core row sums) whenever its shared-memory plan fits; DAT_B200_ATTN_V1=1 keeps
// the first kernel (A/B measurements)
bool attention_fwd_tc_v2(const Shape& s) {
  static const int v1 = [] { const char* e = getenv("DAT_B200_ATTN_V1"); return e && e[0] == '1' ? 1 : 0; }();
  if (v1) return false;
  const Smem2Plan sp = plan_smem2(s.Ns < KV_MAX ? s.Ns : KV_MAX, s.Th + 3, s.Tw + 3, rows_spanned_max(s.HW, s.W));
  return sp.total <= 227 * 1024;
}

size_t attention_fwd_tc_workspace(const Shape& s) {
  // packed table, then (Ns > 256 only) the split-KV partial outputs (bf16) and log-sum-exps
  size_t n = table_bytes(s);
  if (s.Ns > KV_MAX) {
    const size_t nch = (size_t)kv_chunks(s.Ns);
    n += align_up(nch * s.B * s.HW * s.C * 2, 256) + align_up(nch * s.B * s.heads * s.HW * 4, 256);
  }
  return n;
}

// packed table (see pack_table_kernel) into `out`, attention_fwd_tc_workspace(s) bytes
int attention_pack_table(const Shape& s, const float* table, void* out, cudaStream_t st) {
  const int ntab = s.heads * (s.Th + 3) * (s.Tw + 3);
  launch_k(pack_table_kernel, ceil_div(ntab, 256), 256, 0, st, table, (uint2*)out, s.heads, s.Th, s.Tw);
  DAT_LAUNCH_OK("pack_table_kernel");
  return DAT_OK;
}

// (mid, dif) packed table of the forward v2 / backward FAST kernels into `out`, attention_fwd_tc_workspace(s) bytes
int attention_pack_table2(const Shape& s, const float* table, void* out, cudaStream_t st) {
  const int ntab = s.heads * (s.Th + 3) * (s.Tw + 3);
  launch_k(pack_table2_kernel, ceil_div(ntab, 256), 256, 0, st, table, (uint2*)out, s.heads, s.Th, s.Tw);
  DAT_LAUNCH_OK("pack_table2_kernel");
  return DAT_OK;
}

bool attention_fwd_tc_supported(const Shape& s) {
  if (s.act_dtype != DAT_BF16 || s.pe_mode != DAT_PE_RPE) return false;
  // one CTA holds up to 256 samples; more are split into chunks of 256 (+ a 128 and / or a 64 remainder)
  if (s.Ns % 64 != 0 || kv_chunks(s.Ns) > KV_CHUNKS_MAX) return false;
  if (s.C % 8 != 0) return false;
  SmemPlan sp = plan_smem(s.Ns < KV_MAX ? s.Ns : KV_MAX, s.Th + 3, s.Tw + 3, rows_spanned_max(s.HW, s.W));
  return sp.total <= 227 * 1024;
}

int attention_fwd_tc(const Shape& s, const void* q, const void* k, const void* v, const float* pos,
                     const float* table, void* o, float* lse, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  DAT_REQUIRE(attention_fwd_tc_supported(s), "attention_fwd_tc: unsupported shape");
  DAT_REQUIRE(ws != nullptr && ws_bytes >= attention_fwd_tc_workspace(s), "attention_fwd_tc: workspace too small");
  AtcArgs a = make_args(s);
  const int ntab = s.heads * a.Hp * a.Wp;
  const bool v2 = attention_fwd_tc_v2(s);
  if (v2) launch_k(pack_table2_kernel, ceil_div(ntab, 256), 256, 0, st, table, (uint2*)ws, s.heads, s.Th, s.Tw);
  else launch_k(pack_table_kernel, ceil_div(ntab, 256), 256, 0, st, table, (uint2*)ws, s.heads, s.Th, s.Tw);
  DAT_LAUNCH_OK("pack_table_kernel");
  const int nch = kv_chunks(s.Ns);
  bf16* o_dst = (bf16*)o;
  float* lse_dst = lse;
  if (nch > 1) {   // partial slots behind the packed table
    o_dst = (bf16*)((char*)ws + table_bytes(s));
    lse_dst = (float*)((char*)o_dst + align_up((size_t)nch * s.B * s.HW * s.C * 2, 256));
    a.o_zstride = (long long)s.B * s.HW * s.C;
    a.lse_zstride = (long long)s.B * s.heads * s.HW;
  }
  CUtensorMap tmQ;
  DAT_FWD(tc::make_tmap_2d(&tmQ, q, 2, false, (uint64_t)s.B * s.HW, (uint64_t)s.C, (uint64_t)s.C * 2, TQ, 32, 64));
  const int gx = pick_chunks(s.B * s.heads * nch, a.n_tiles);
  // one launch per chunk size: all 256-sample chunks along grid.z, then the 128 / 64 remainders
  int n_off = 0, z = 0;
  for (int size : {256, 128, 64}) {
    int count = s.Ns <= KV_MAX ? (s.Ns == size ? 1 : 0) : (size == 256 ? s.Ns / 256 : ((s.Ns % 256) & size ? 1 : 0));
    if (count == 0) continue;
    CUtensorMap tmK, tmV;
    DAT_FWD(tc::make_tmap_2d(&tmK, k, 2, false, (uint64_t)s.B * s.Ns, (uint64_t)s.C, (uint64_t)s.C * 2, size, 32, 64));
    DAT_FWD(tc::make_tmap_2d(&tmV, v, 2, false, (uint64_t)s.B * s.Ns, (uint64_t)s.C, (uint64_t)s.C * 2, size, 32, 64));
    SmemPlan sp = plan_smem(size, a.Hp, a.Wp, a.rows_max);
    dim3 grid(gx, s.B * s.heads, count);
    a.n_off0 = n_off;
    a.z0 = z;
    if (v2) {
      const Smem2Plan sp2 = plan_smem2(size, a.Hp, a.Wp, a.rows_max);
#define LAUNCH2(NSV)                                                                              \
  do {                                                                                            \
    auto kern = attn_fwd_tc2_kernel<NSV>;                                                         \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp2.total)); \
    launch_k(kern, grid, A2_THREADS, sp2.total, st, tmQ, tmK, tmV, pos, (const uint2*)ws, o_dst, lse_dst, a); \
  } while (0)
      if (size == 256) LAUNCH2(256);
      else if (size == 128) LAUNCH2(128);
      else LAUNCH2(64);
#undef LAUNCH2
      DAT_LAUNCH_OK("attn_fwd_tc2_kernel");
      n_off += count * size;
      z += count;
      continue;
    }
#define LAUNCH(NSV)                                                                              \
  do {                                                                                           \
    auto kern = attn_fwd_tc_kernel<NSV>;                                                         \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.total)); \
    launch_k(kern, grid, ATC_THREADS, sp.total, st, tmQ, tmK, tmV, pos, (const uint2*)ws, o_dst, lse_dst, a); \
  } while (0)
    if (size == 256) LAUNCH(256);
    else if (size == 128) LAUNCH(128);
    else LAUNCH(64);
#undef LAUNCH
    DAT_LAUNCH_OK("attn_fwd_tc_kernel");
    n_off += count * size;
    z += count;
  }
  if (nch > 1) {
    const long long total = (long long)s.B * s.HW * s.heads * 4;
    launch_k(attn_combine_kernel, ceil_div(total, 256), 256, 0, st, o_dst, lse_dst, (bf16*)o, lse, nch, s.HW, s.C, s.heads,
                                                               a.o_zstride, a.lse_zstride, total);
    DAT_LAUNCH_OK("attn_combine_kernel");
  }
  return DAT_OK;
}

}  // namespace dat
