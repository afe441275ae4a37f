// Backward of the fused attention core on the tcgen05 tensor cores (bf16 hot path):
// autograd of  S = QK^T*scale + bias(pos, rpe_table);  P = softmax(S);  O = PV
// (dat_blocks.py:180-223).  Given dO, lse and delta = rowsum(dO*O):
//   dP = dO V^T;  dS = P*(dP - delta);  dQ = scale*dS K;  dK = scale*dS^T Q;  dV = P^T dO
//   d rpe_table += dS * (4 bilinear tap weights);   d pos = -k * sum_q dS * d bias/d(ix, iy)
//
// One CTA owns one (batch, head) and walks 128-query tiles; the Ns sampled keys are processed
// in halves of 128 columns.  Five GEMMs per half run on the tensor cores, all operands in
// shared memory, all accumulators in tensor memory:
//   S_h = Q K_h^T, dP_h = dO V_h^T          (K-major A and B, 64B swizzle)      -> TMEM [0,256)
//   dQ += dS_h K_h                           (A = dS_h K-major, B = K_h MN-major)
//   dK_h += dS_h^T Q,  dV_h += P_h^T dO      (A = the same dS_h / P_h tiles read MN-major,
//                                             B = Q / dO tiles read MN-major)
// dK/dV accumulate over all tiles of the CTA in TMEM and are written once at the end.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2-3 idle, 4-11 = compute
// (thread = query row x 64 columns of the half, in two 32-column steps): recompute the bias
// (same separable / packed-table scheme as the forward), P and dS -> bf16 tiles in shared
// memory (128B swizzle), and the two bias gradients:
//   * d pos: per-column sums over the 32 rows of a warp by a register transpose-reduction
//     (31 shuffles per 32 columns), accumulated in registers over all tiles - no atomics;
//   * d rpe_table: lanes are consecutive queries of one image row and hit the same table cell
//     in runs (table step per query < 1), so runs are pre-summed with a segmented shuffle
//     reduction and only run leaders issue shared-memory atomics into a per-CTA padded fp32
//     copy of the table, flushed with one global atomic per cell at the end.
#include <cstdlib>
#include <type_traits>

#include "kernels.h"
#include "tc_common.cuh"

namespace dat {

// Phase cycle counters of compute warp 4, summed over the CTAs of every launch (debug aid, read by
// dat_debug_attn_bwd_timing): [0] tile loop total, [1] wait for S / dP, [2] score loop (bias, P, dS,
// table gradient), [3] d pos column sums, [4] dQ wait + store, [5] per-tile setup, [6] CTAs.
__device__ unsigned long long g_attn_bwd_prof[8];

namespace {

using namespace tc;

constexpr int TQ = 128;
constexpr int NHC = 128;             // key columns per half
constexpr int BTC_THREADS = 384;     // 4 control + 8 compute warps
constexpr int COMP_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MAGIC = 12582912.0f;
constexpr int MAGIC_BITS = 0x4B400000;
constexpr unsigned FULL = 0xffffffffu;

// TMEM column map (512 columns allocated)
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_DK = 288, TM_DV = 352;

// Phase counters of the compute warps (tools/attn_bwd_phases.py): compiled in only with -DDAT_ATTN_BWD_PROFILE
// (DAT_B200_BUILD_DEFS, tagged build) - six 64-bit counters and a clock read per 32 scores cost the production kernel
// registers it does not have (168 per thread, spills).
#ifdef DAT_ATTN_BWD_PROFILE
#define PF(...) __VA_ARGS__
#else
#define PF(...)
#endif

struct BtcArgs {
  int B, H, W, HW, C, heads, G, hg, Th, Tw, Wp, Hp;
  int n_tiles, rows_max, chunks, nslots, light_table;
  int ns_total;            // samples of the block; a CTA works on the NS of them that start at blockIdx.z * NS
  long long dq_slab;       // elements between the dq outputs of consecutive sample chunks (B * HW * C)
  float c1, scale, kx, ky, gsx, gsy;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void comp_bar_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
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

// sum over the 32 lanes of v[lane-th column]: lane L returns sum_rows v_row[L]  (31 shuffles)
__device__ __forceinline__ float column_sums32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = upper ? v[i] : v[i + w];
      const float keep = upper ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, w);
    }
  }
  return v[0];
}

// the same for bf16x2-packed pairs: lane L returns {sum_rows lo[L], sum_rows hi[L]} (31 shuffles, 31 packed adds)
__device__ __forceinline__ uint32_t hadd2_bf16(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t column_sums32_bf16x2(uint32_t (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const uint32_t send = upper ? v[i] : v[i + w];
      const uint32_t keep = upper ? v[i + w] : v[i];
      v[i] = hadd2_bf16(keep, __shfl_xor_sync(FULL, send, w));
    }
  }
  return v[0];
}
__device__ __forceinline__ uint32_t hfma2_bf16_b(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint2 lds64_b(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// compact packed table: entry (y, x) at (y + 2) * Wp + (x + 2) = bf16x2 {T[y][x], T[y][x+1]} * log2(e),
// zero outside the table
__global__ void pack_table_compact_kernel(const float* __restrict__ table, uint32_t* __restrict__ out,
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
  out[idx] = pack_bf16x2(at(y, x), at(y, x + 1));
}

struct SmemPlanB {
  uint32_t q[2], d_o[2], k, v, p, ds, tab, dtab, yt, xk, yk, dpos, bars, total;
};
// yt_bytes: 8 = {row offset, y fraction} (scatter / compact variants), 16 = {row address, y fraction bf16x2, x constant, -}
// (FAST variant: one broadcast LDS.128 per score, as in the forward v2)
__host__ __device__ inline SmemPlanB plan_smem_b(int NS, int Hp, int Wp, int rows_max, bool compact,
                                                 int ndt, int nslots, int yt_bytes = 8) {
  SmemPlanB s;
  uint32_t off = 0;
  s.q[0] = off; off += TQ * 64;          // Q slots are contiguous: slot i at q[0] + i * 8 KB
  s.q[1] = off; off += (nslots > 1 ? TQ * 64 : 0);
  s.d_o[0] = off; off += TQ * 64;
  s.d_o[1] = off; off += (nslots > 1 ? TQ * 64 : 0);
  s.k = off; off += NS * 64;
  s.v = off; off += NS * 64;
  s.p = off; off += 2 * 16384;      // [128 x 128] bf16, two 64-column K-blocks
  s.ds = off; off += 2 * 16384;
  s.tab = off; off += ((uint32_t)(Hp * Wp) * (compact ? 4 : 8) + 15) & ~15u;
  s.dtab = off; off += ((uint32_t)(ndt * Hp * Wp) * 4 + 15) & ~15u;
  s.yt = off; off += 2u * (uint32_t)rows_max * NS * yt_bytes;   // double-buffered: built one tile ahead
  s.xk = off; off += NS * 4;
  s.yk = off; off += NS * 4;
  s.dpos = s.p;                            // per-warp column sums [8][NS][2] alias P/dS at the end
  s.bars = off; off += 16 * 8;
  s.total = off + 1024;
  return s;
}

// COMPACT: table entries are 4 bytes {T[y][x], T[y][x+1]} (two LDS per score) instead of the
//          8-byte 4-tap entries: halves the table footprint for the 111 x 111 stage-0 table.
// PRIV:    every compute warp owns a private padded copy of the table gradient and updates it
//          with plain read-modify-writes (no atomics); needs W % 32 == 0 (a warp's 32 queries
//          lie in one image row, so its run leaders hit distinct cells).
// TBL:     true = the table gradient is accumulated inside this kernel (shared-memory scatter, above);
//          false = dS (bf16, [b*heads][NS/8][m][8]) is streamed to `ds_out` instead and the table gradient is
//          formed from it by rpe_table_grad_mma (rpe_table_grad.cu) as small tensor-core GEMMs - the
//          per-score scatter is 50-75 % of this kernel's time, the 16-byte stores are ~2 %.
template <int NS, bool COMPACT, bool PRIV, bool TBL>
__global__ void __launch_bounds__(BTC_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const float* __restrict__ pos, const void* __restrict__ tab_packed,
                   const float* __restrict__ lse, const float* __restrict__ delta,
                   bf16* __restrict__ dq, float* __restrict__ dk_part, float* __restrict__ dv_part,
                   float* __restrict__ d_table, float* __restrict__ dpos_part, bf16* __restrict__ ds_out,
                   BtcArgs a) {
  pdl_enter();
  constexpr int NHALF = NS / NHC;
  constexpr int NDT = !TBL ? 0 : (PRIV ? 8 : 1);          // copies of the table gradient
  // FAST (dS streamed out, 8-byte table entries - every DAT++ stage but the 111 x 111 table of stage 0): the bias and
  // its two derivatives use the forward-v2 scheme ((mid, dif) table, one LDS.128 of per-sample parameters, HFMA2 blend
  // in y, range clamps only when a sample of the CTA lies outside [-1, 1]) and the d pos column sums run on bf16x2 pairs
  constexpr bool FAST = !TBL && !COMPACT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  const SmemPlanB sp = plan_smem_b(NS, a.Hp, a.Wp, a.rows_max, COMPACT, NDT, a.nslots, FAST ? 16 : 8);
  uint8_t* sQ0 = smem + sp.q[0];
  uint8_t* sDO0 = smem + sp.d_o[0];
  uint8_t* sK = smem + sp.k;
  uint8_t* sV = smem + sp.v;
  uint8_t* sP = smem + sp.p;
  uint8_t* sDS = smem + sp.ds;
  uint2* sTab = reinterpret_cast<uint2*>(smem + sp.tab);
  uint32_t* sTabC = reinterpret_cast<uint32_t*>(smem + sp.tab);
  float* sDTab = reinterpret_cast<float*>(smem + sp.dtab);
  int2* sYt = reinterpret_cast<int2*>(smem + sp.yt);
  float* sXk = reinterpret_cast<float*>(smem + sp.xk);
  float* sYk = reinterpret_cast<float*>(smem + sp.yk);
  float* sDpos = reinterpret_cast<float*>(smem + sp.p);   // aliases the P / dS tiles (used after the last MMA)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_ready = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_free = bars + 8;
  uint64_t* pds_free = bars + 11;   // (bars + 9, + 10 hold the tensor-memory slot and the out-of-range flag)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  uint32_t* oob_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / a.heads, eta = bh % a.heads, g = eta / a.hg;
  // More than 256 samples (a 512 x 2048 crop has 1024): the saved log-sum-exp and delta make sample chunks independent
  // in the backward - P = exp(S - lse) needs no running maximum - so chunk blockIdx.z is simply another CTA: its dK / dV /
  // d pos / dS cover disjoint samples, its dQ goes to its own slab and the slabs are summed afterwards.
  const int n_off = (int)blockIdx.z * NS;

  if (threadIdx.x == 0) {
    *oob_flag = 0u;
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_ready, COMP_THREADS);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, COMP_THREADS);
    mbar_init(pds_free, 1);
    fence_barrier_init();
  }
  if (FAST) __syncthreads();            // the out-of-range flag is zero before any thread may raise it
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  {
    if (COMPACT) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(tab_packed) + (long long)eta * a.Hp * a.Wp;
      for (int i = threadIdx.x; i < a.Hp * a.Wp; i += BTC_THREADS) sTabC[i] = src[i];
    } else {
      const uint2* src = reinterpret_cast<const uint2*>(tab_packed) + (long long)eta * a.Hp * a.Wp;
      for (int i = threadIdx.x; i < a.Hp * a.Wp; i += BTC_THREADS) sTab[i] = src[i];
    }
    for (int i = threadIdx.x; i < NDT * a.Hp * a.Wp; i += BTC_THREADS) sDTab[i] = 0.f;
    const float* pp = pos + (((long long)b * a.G + g) * a.ns_total + n_off) * 2;
    for (int n = threadIdx.x; n < NS; n += BTC_THREADS) {
      const float px = pp[2 * n + 1];
      if (FAST && !(fabsf(px) <= 1.0f)) *oob_flag = 1u;     // benign race: every writer stores the same value
      sYk[n] = pp[2 * n] * a.ky;
      sXk[n] = px * a.kx;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const bool xclamp = FAST ? *oob_flag != 0u : true;      // samples inside [-1, 1] never leave the padded table

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ------------------------------------------------------------------
      mbar_arrive_expect_tx(kv_full, 2u * NS * 64u);
      tma_load_2d(sK, &tmK, kv_full, eta * 32, b * a.ns_total + n_off);
      tma_load_2d(sV, &tmV, kv_full, eta * 32, b * a.ns_total + n_off);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int slot = it % a.nslots;
        mbar_wait(&qdo_empty[slot], ((uint32_t)(it / a.nslots) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&qdo_full[slot], 2u * TQ * 64u);
        tma_load_2d(sQ0 + slot * (TQ * 64), &tmQ, &qdo_full[slot], eta * 32, b * a.HW + tile * TQ);
        tma_load_2d(sDO0 + slot * (TQ * 64), &tmDO, &qdo_full[slot], eta * 32, b * a.HW + tile * TQ);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ---------------------------------------------------------------------
      const uint32_t idesc_s = make_instr_desc(FMT_BF16, TQ, NHC);             // S, dP
      const uint32_t idesc_q = make_instr_desc(FMT_BF16, TQ, 32, 0, 1);        // dQ: B MN-major
      const uint32_t idesc_kv = make_instr_desc(FMT_BF16, NHC, 32, 1, 1);      // dK, dV: A and B MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      mbar_wait(kv_full, 0);
      const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int E = n_my * NHALF;                       // halves this CTA processes
      // S_h = Q K_h^T, dP_h = dO V_h^T into the (single) S / dP buffer
      auto issue_sdp = [&](int e) {
        const int it_e = e / NHALF, h = e % NHALF, slot = it_e % a.nslots;
        if (h == 0) mbar_wait(&qdo_full[slot], (uint32_t)(it_e / a.nslots) & 1u);
        tc_fence_after_sync();
        const uint32_t q_addr = smem_u32(sQ0 + slot * (TQ * 64));
        const uint32_t do_addr = smem_u32(sDO0 + slot * (TQ * 64));
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          mma_bf16_ss(tmem_base + TM_S, make_smem_desc(q_addr + k * 32, 16, 512, LAYOUT_SW64),
                      make_smem_desc(k_addr + h * NHC * 64 + k * 32, 16, 512, LAYOUT_SW64), idesc_s, (uint32_t)k);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          mma_bf16_ss(tmem_base + TM_DP, make_smem_desc(do_addr + k * 32, 16, 512, LAYOUT_SW64),
                      make_smem_desc(v_addr + h * NHC * 64 + k * 32, 16, 512, LAYOUT_SW64), idesc_s, (uint32_t)k);
        }
        tc_commit(sdp_full);
      };
      issue_sdp(0);
      for (int e = 0; e < E; ++e) {
        const int it = e / NHALF, h = e % NHALF, slot = it % a.nslots;
        const uint32_t q_addr = smem_u32(sQ0 + slot * (TQ * 64));
        const uint32_t do_addr = smem_u32(sDO0 + slot * (TQ * 64));
        mbar_wait(pds_ready, (uint32_t)e & 1u);        // P_h / dS_h are in shared memory, S / dP have been read
        // The next half's S / dP go FIRST: the compute warps get their next scores after 4 MMAs instead of after
        // the 24 gradient MMAs of this half (which they wait for only before overwriting the P / dS tiles: pds_free).
        // A single Q / dO slot cannot do that across a tile boundary (the next tile's Q arrives only after this
        // tile's last MMAs have released the slot).
        const bool early = e + 1 < E && !(a.nslots == 1 && h == NHALF - 1);
        if (early) issue_sdp(e + 1);
        if (h == 0 && it > 0) mbar_wait(dq_free, (uint32_t)(it - 1) & 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int j = 0; j < NHC / 16; ++j) {     // contraction over the 128 queries of the tile
          const uint64_t a_p = make_smem_desc(p_addr + j * 2048, 16384, 1024, LAYOUT_SW128);
          const uint64_t a_ds = make_smem_desc(ds_addr + j * 2048, 16384, 1024, LAYOUT_SW128);
          const uint64_t b_do = make_smem_desc(do_addr + j * 1024, 512, 512, LAYOUT_SW64);
          const uint64_t b_q = make_smem_desc(q_addr + j * 1024, 512, 512, LAYOUT_SW64);
          const uint32_t acc = (uint32_t)((it | j) != 0);
          mma_bf16_ss(tmem_base + TM_DV + h * 32, a_p, b_do, idesc_kv, acc);
          mma_bf16_ss(tmem_base + TM_DK + h * 32, a_ds, b_q, idesc_kv, acc);
        }
#pragma unroll
        for (int kk = 0; kk < NHC / 16; ++kk) {  // contraction over the 128 keys of the half
          const uint64_t a_ds = make_smem_desc(ds_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, LAYOUT_SW128);
          const uint64_t b_k = make_smem_desc(k_addr + (h * NHC + kk * 16) * 64, 512, 512, LAYOUT_SW64);
          mma_bf16_ss(tmem_base + TM_DQ, a_ds, b_k, idesc_q, (uint32_t)((h | kk) != 0));
        }
        tc_commit(pds_free);                     // the P / dS tiles of this half have been consumed
        if (h == NHALF - 1) {
          tc_commit(dq_full);
          tc_commit(&qdo_empty[slot]);
        }
        if (!early && e + 1 < E) issue_sdp(e + 1);
      }
    }
  } else if (warp >= 4) {
    // ---- compute warps ------------------------------------------------------------------------
    const int quad = warp & 3, chalf = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int ctid = threadIdx.x - 128;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    float dpx_acc[NHALF * 2], dpy_acc[NHALF * 2];
#pragma unroll
    for (int i = 0; i < NHALF * 2; ++i) dpx_acc[i] = dpy_acc[i] = 0.f;
    int it = 0;
    PF(long long pf_total = clock64(), pf_wait = 0, pf_score = 0, pf_col = 0, pf_dq = 0, pf_setup = 0;)
    // per-(image row of the tile, sample) y footprint, double-buffered: the table of tile i + 1 is built after the score
    // loops of tile i, so a tile starts with ONE barrier and no table build on its critical path
    const uint32_t yt_stride = (uint32_t)a.rows_max * NS * (FAST ? 16u : 8u);
    auto build_yt = [&](int tile_b, int buf) {
      uint8_t* dstb = reinterpret_cast<uint8_t*>(sYt) + (size_t)buf * yt_stride;
      const int r0b = (tile_b * TQ) / a.W;
      const int r_lastb = min(a.HW - 1, tile_b * TQ + TQ - 1) / a.W;
      for (int e = ctid; e < (r_lastb - r0b + 1) * NS; e += COMP_THREADS) {
        const int rr = e / NS, n = e - rr * NS;
        const float gy = fmaf((float)(r0b + rr), a.gsy, -1.0f);
        const float ay = (gy * 0.25f + 0.5f) * (float)(a.Th - 1) - 0.5f;
        float u = ay - sYk[n];
        u = fminf(fmaxf(u, -1.5f), (float)a.Th - 0.5f);
        const float aa = u + MAGIC;
        const float fy = (u - (aa - MAGIC)) + 0.5f;
        const int y0 = __float_as_int(aa) - MAGIC_BITS;
        if (FAST) {
          const uint32_t ro8 = smem_u32(sTab) + ((uint32_t)((y0 + 2) * a.Wp + 2) << 3) - ((uint32_t)MAGIC_BITS << 3);
          reinterpret_cast<uint4*>(dstb)[e] = make_uint4(ro8, pack_bf16x2(fy, fy), __float_as_uint(sXk[n]), 0u);
        } else {
          reinterpret_cast<int2*>(dstb)[e] = make_int2((y0 + 2) * a.Wp + 2 - MAGIC_BITS, __float_as_int(fy));
        }
      }
    };
    // dQ of a finished tile: TMEM -> * scale -> bf16 -> global (each compute warp writes 16 channels).  Deferred into
    // the first half of the NEXT tile: its MMAs are long done by then, so no thread waits for the tensor pipe
    auto dq_epilogue = [&](int it_e, int m_e, bool valid_e) {
      mbar_wait(dq_full, (uint32_t)it_e & 1u);
      tc_fence_after_sync();
      uint32_t qv[16];
      tmem_ld_32x16(t_lane + TM_DQ + (uint32_t)(chalf * 16), qv);
      tmem_wait_ld();
      tc_fence_before_sync();
      mbar_arrive(dq_free);
      if (valid_e) {
        uint32_t pk8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pk8[i] = pack_bf16x2(__uint_as_float(qv[2 * i]) * a.scale, __uint_as_float(qv[2 * i + 1]) * a.scale);
        uint4* dst = reinterpret_cast<uint4*>(dq + (long long)blockIdx.z * a.dq_slab + ((long long)b * a.HW + m_e) * a.C + eta * 32 + chalf * 16);
        dst[0] = make_uint4(pk8[0], pk8[1], pk8[2], pk8[3]);
        dst[1] = make_uint4(pk8[4], pk8[5], pk8[6], pk8[7]);
      }
    };
    build_yt(blockIdx.x, 0);
    int m_prev = 0;
    bool valid_prev = false;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      PF(long long pf_t = clock64();)
      const int m = tile * TQ + row;
      const bool valid = m < a.HW;
      const int mm = valid ? m : a.HW - 1;
      const int r = mm / a.W, c = mm - r * a.W;
      const int r0 = (tile * TQ) / a.W;
      comp_bar_sync();     // this tile's y table is complete; the other buffer is free for the next tile's
      const uint8_t* yt_base = reinterpret_cast<const uint8_t*>(sYt) + (size_t)(it & 1) * yt_stride;
      const float ax = (fmaf((float)c, a.gsx, -1.0f) * 0.25f + 0.5f) * (float)(a.Tw - 1) - 0.5f;
      const float xhi = (float)a.Tw - 0.5f;
      // FAST: rows beyond HW get lse = +inf, i.e. p = 2^(-inf) = 0, instead of a select per score
      const float lse2 = (FAST && !valid) ? INFINITY : lse[(long long)bh * a.HW + mm] * LOG2E;
      const float dl = delta[(long long)bh * a.HW + mm];
      const int r_up = __shfl_up_sync(FULL, r, 1);
      const bool row_head = lane == 0 || r_up != r;
      const int r_up2 = __shfl_up_sync(FULL, r, 2), r_dn = __shfl_down_sync(FULL, r, 1);
      const bool rs1 = lane >= 1 && r_up == r, rs2 = lane >= 2 && r_up2 == r;   // lane-1 / lane-2 in my image row
      const bool row_end = lane == 31 || r_dn != r;
      float* mytab = sDTab + (PRIV ? (warp - 4) * a.Hp * a.Wp : 0);
      uint8_t* prow_p = sP + chalf * 16384 + row * 128;
      uint8_t* prow_d = sDS + chalf * 16384 + row * 128;
      // dS for the table gradient: groups of 8 samples, [bh][NS / 8][m][8] - a warp's 32 rows write 512 contiguous bytes
      bf16* ds_row = TBL ? nullptr : ds_out + (long long)bh * a.HW * a.ns_total + (long long)mm * 8;

      PF(pf_setup += clock64() - pf_t;)
#pragma unroll 1
      for (int h = 0; h < NHALF; ++h) {
        const uint32_t e_idx = (uint32_t)(it * NHALF + h);
        PF(pf_t = clock64();)
        mbar_wait(sdp_full, e_idx & 1u);
        tc_fence_after_sync();
        PF(pf_wait += clock64() - pf_t;)
        if constexpr (FAST) {
          auto fast_body = [&](auto xc_tag) {
            constexpr bool XC = decltype(xc_tag)::value;
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
              const int col0 = chalf * 64 + sub * 32;          // column within the half
              uint32_t sv[32], dpv[32];
              PF(pf_t = clock64();)
              tmem_ld_32x32(t_lane + TM_S + (uint32_t)col0, sv);
              tmem_ld_32x32(t_lane + TM_DP + (uint32_t)col0, dpv);
              tmem_wait_ld();
              const int nbase = h * NHC + col0;
              const uint4* yt = reinterpret_cast<const uint4*>(yt_base) + (r - r0) * NS + nbase;
              uint32_t g2[32];
              uint32_t pp[4], dd[4];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint4 ye = yt[j];
                float u = ax - __uint_as_float(ye.z);
                if (XC) u = fminf(fmaxf(u, -1.5f), xhi);
                const float aa = u + MAGIC;
                const float fxp = u - (aa - MAGIC);                               // x fraction - 1/2
                const uint2 e = lds64_b(ye.x + (__float_as_uint(aa) << 3));
                const uint32_t md = hfma2_bf16_b(ye.y, e.y, e.x);                 // blend in y: {mid, dif}
                const float dxb = __uint_as_float(md & 0xffff0000u);              // d bias / d ix  (x log2e)
                const float bias = fmaf(fxp, dxb, __uint_as_float(md << 16));
                const float dyb = fmaf(fxp, __uint_as_float(e.y & 0xffff0000u), __uint_as_float(e.y << 16));   // d bias / d iy
                const float tv = fmaf(__uint_as_float(sv[j]), a.c1, bias);
                const float p = ex2(tv - lse2);
                const float ds = p * (__uint_as_float(dpv[j]) - dl);
                g2[j] = pack_bf16x2(ds * dxb, ds * dyb);
                if (j & 1) {
                  pp[(j >> 1) & 3] = pack_bf16x2(__uint_as_float(sv[j - 1]), p);
                  dd[(j >> 1) & 3] = pack_bf16x2(__uint_as_float(dpv[j - 1]), ds);
                } else {                                             // park p, ds until the pair is complete
                  sv[j] = __float_as_uint(p);
                  dpv[j] = __float_as_uint(ds);
                }
                if ((j & 7) == 7) {
                  // first store of this half: the tensor core must have consumed the previous half's tiles
                  if (j == 7 && sub == 0 && e_idx > 0) mbar_wait(pds_free, (e_idx - 1u) & 1u);
                  const int ch = sub * 4 + (j >> 3);
                  const uint32_t sw = (uint32_t)((ch ^ (row & 7)) << 4);
                  *reinterpret_cast<uint4*>(prow_p + sw) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
                  *reinterpret_cast<uint4*>(prow_d + sw) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
                  if (valid)      // the same 8 dS values, [m][n] layout, for the tensor-core table gradient
                    *reinterpret_cast<uint4*>(ds_row + (long long)((n_off + nbase + j - 7) >> 3) * a.HW * 8) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
                }
              }
              // d pos: column sums over this warp's 32 rows on bf16x2 pairs {x part, y part} (lane L ends up with
              // column L); the running sums over tiles stay fp32
              PF(const long long pf_m = clock64();)
              PF(pf_score += pf_m - pf_t;)
              const uint32_t cs = column_sums32_bf16x2(g2, lane);
              dpx_acc[h * 2 + sub] += __uint_as_float(cs << 16);
              dpy_acc[h * 2 + sub] += __uint_as_float(cs & 0xffff0000u);
              PF(pf_col += clock64() - pf_m;)
            }
          };
          if (xclamp) fast_body(std::true_type{});
          else fast_body(std::false_type{});
        } else {
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          const int col0 = chalf * 64 + sub * 32;          // column within the half
          uint32_t sv[32], dpv[32];
          PF(pf_t = clock64();)
          tmem_ld_32x32(t_lane + TM_S + (uint32_t)col0, sv);
          tmem_ld_32x32(t_lane + TM_DP + (uint32_t)col0, dpv);
          tmem_wait_ld();
          const int nbase = h * NHC + col0;
          const int2* yt = reinterpret_cast<const int2*>(yt_base) + (r - r0) * NS + nbase;
          const float* xk = sXk + nbase;
          float gxs[32], gys[32];
          uint32_t pp[16], dd[16];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int2 ye = yt[j];
            const float ur = ax - xk[j];
            const float u = fminf(fmaxf(ur, -1.5f), xhi);
            const float aa = u + MAGIC;
            const float fx = (u - (aa - MAGIC)) + 0.5f;
            const float fy = __int_as_float(ye.y);
            const int idx = ye.x + __float_as_int(aa);
            float t00, d0, t10, d1;
            if (COMPACT) {
              const uint32_t e0 = sTabC[idx], e1 = sTabC[idx + a.Wp];
              t00 = __uint_as_float(e0 << 16);
              d0 = __uint_as_float(e0 & 0xffff0000u) - t00;
              t10 = __uint_as_float(e1 << 16);
              d1 = __uint_as_float(e1 & 0xffff0000u) - t10;
            } else {
              const uint2 e = sTab[idx];
              t00 = __uint_as_float(e.x << 16); d0 = __uint_as_float(e.x & 0xffff0000u);
              t10 = __uint_as_float(e.y << 16); d1 = __uint_as_float(e.y & 0xffff0000u);
            }
            const float top = fmaf(fx, d0, t00), bot = fmaf(fx, d1, t10);
            const float dyb = bot - top;                         // d bias / d iy  (x log2e)
            const float bias = fmaf(fy, dyb, top);
            const float dxb = fmaf(fy, d1 - d0, d0);             // d bias / d ix  (x log2e)
            const float tv = fmaf(__uint_as_float(sv[j]), a.c1, bias);
            const float p = valid ? ex2(tv - lse2) : 0.f;
            const float ds = p * (__uint_as_float(dpv[j]) - dl);
            gxs[j] = ds * dxb;
            gys[j] = ds * dyb;
            if (j & 1) {
              pp[(j >> 1) & 3] = pack_bf16x2(__uint_as_float(sv[j - 1]), p);
              dd[(j >> 1) & 3] = pack_bf16x2(__uint_as_float(dpv[j - 1]), ds);
            } else {                                             // park p, ds until the pair is complete
              sv[j] = __float_as_uint(p);
              dpv[j] = __float_as_uint(ds);
            }
            if ((j & 7) == 7) {
              // P_h, dS_h -> shared memory (K-major, 128B swizzle), 16 bytes per 8 columns
              if (j == 7 && sub == 0 && e_idx > 0) mbar_wait(pds_free, (e_idx - 1u) & 1u);   // previous tiles consumed
              const int ch = sub * 4 + (j >> 3);
              const uint32_t sw = (uint32_t)((ch ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(prow_p + sw) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
              *reinterpret_cast<uint4*>(prow_d + sw) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
              if (!TBL && valid)      // the same 8 dS values, [m][n] layout, for the tensor-core table gradient
                *reinterpret_cast<uint4*>(ds_row + (long long)((n_off + nbase + j - 7) >> 3) * a.HW * 8) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
            }
            // ---- d rpe_table ----------------------------------------------------------------
            // Lanes = consecutive queries of an image row; the table step per query is < 1 cell,
            // so lanes form runs with the same north-west cell.  (1) segmented shuffle reduction
            // of P = ds(1-fx) -> cell x0 and Q = ds*fx -> cell x0+1 over each run (windows of 8);
            // (2) a run head also takes the Q of the previous run when that run is its left
            // neighbour (same image row, cell x0-1, single window), so a run usually issues 2
            // updates (rows y0, y0+1 of one cell) instead of 4.
            if (TBL) {
              float* cell = mytab + idx;                  // padded index: no bounds checks
              const float wy1 = fy, wy0 = 1.0f - fy;
              // Light path (0.29 <= table step per query <= 1, nothing clamped: every shipped config):
              // runs are at most 4 lanes and consecutive runs hit consecutive cells.  Two-level
              // segmented prefix scans leave the run totals in the run's LAST lane, which sits next
              // to the head of the following run: the fx part of run k moves to run k + 1 with one
              // shuffle, so a run issues 2 updates (8 shuffles, no ballots / bit scans).
              if (PRIV && a.light_table && !__any_sync(FULL, u != ur)) {   // (measured slower with atomics)
                const int i_u1 = __shfl_up_sync(FULL, idx, 1), i_u2 = __shfl_up_sync(FULL, idx, 2);
                const int i_d1 = __shfl_down_sync(FULL, idx, 1);
                const bool s1 = rs1 && i_u1 == idx, s2 = rs2 && i_u2 == idx;
                const bool tail = row_end || i_d1 != idx;
                float Qs = ds * fx, Ps = ds - Qs, t;
                t = __shfl_up_sync(FULL, Qs, 1); Qs += s1 ? t : 0.f;
                t = __shfl_up_sync(FULL, Qs, 2); Qs += s2 ? t : 0.f;
                const float qp = __shfl_up_sync(FULL, Qs, 1);          // previous run's total (from its last lane)
                Ps += (!s1 && rs1 && i_u1 + 1 == idx) ? qp : 0.f;
                t = __shfl_up_sync(FULL, Ps, 1); Ps += s1 ? t : 0.f;
                t = __shfl_up_sync(FULL, Ps, 2); Ps += s2 ? t : 0.f;
                const bool absorbed = !row_end && i_d1 == idx + 1;     // the next run's head took my fx part
                if (tail) { cell[0] += Ps * wy0; cell[a.Wp] += Ps * wy1; }
                __syncwarp();
                if (tail && !absorbed) { cell[1] += Qs * wy0; cell[a.Wp + 1] += Qs * wy1; }
                __syncwarp();
              } else {
                const int pk = __shfl_up_sync(FULL, idx, 1);
                const bool head = row_head || pk != idx;
                const unsigned heads = __ballot_sync(FULL, head);
                const int start = 31 - __clz((int)(heads & (FULL >> (31 - lane))));
                float Qw = ds * fx, Pw = ds - Qw;
#pragma unroll
                for (int dsh = 1; dsh <= 4; dsh <<= 1) {
                  const float oP = __shfl_down_sync(FULL, Pw, dsh), oQ = __shfl_down_sync(FULL, Qw, dsh);
                  const int os = __shfl_down_sync(FULL, start, dsh);
                  const bool same = (lane + dsh < 32) && (os == start);
                  Pw += same ? oP : 0.f;
                  Qw += same ? oQ : 0.f;
                }
                const unsigned above = (heads >> lane) >> 1;              // heads strictly above my lane
                const int nh = above ? lane + __ffs((int)above) : 32;    // next run head (32: none)
                const int src_n = nh & 31;
                const int idx_n = __shfl_sync(FULL, idx, src_n);
                const int r_n = __shfl_sync(FULL, r, src_n);
                const bool is_head = lane == start;
                // my Q is absorbed by the next run's head
                const bool absorbed = is_head && nh < 32 && (nh - lane) <= 8 && idx_n == idx + 1 && r_n == r;
                // I (a run head) absorb the previous run's Q: previous head lane
                const unsigned below = heads & ((1u << lane) - 1u);
                const int ph = below ? 31 - __clz((int)below) : 0;
                const float q_prev = __shfl_sync(FULL, Qw, ph);
                const bool take = is_head && below != 0u && (lane - ph) <= 8 && pk + 1 == idx && r_up == r;
                const float Uw = Pw + (take ? q_prev : 0.f);
                const bool leader = ((lane - start) & 7) == 0;
                if (PRIV) {
                  if (leader) { cell[0] += Uw * wy0; cell[a.Wp] += Uw * wy1; }
                  __syncwarp();
                  if (leader && !absorbed) { cell[1] += Qw * wy0; cell[a.Wp + 1] += Qw * wy1; }
                  __syncwarp();
                } else if (leader) {
                  atomicAdd(cell, Uw * wy0);
                  atomicAdd(cell + a.Wp, Uw * wy1);
                  if (!absorbed) {
                    atomicAdd(cell + 1, Qw * wy0);
                    atomicAdd(cell + a.Wp + 1, Qw * wy1);
                  }
                }
              }
            }
          }
          // d pos: column sums over this warp's 32 rows (lane L ends up with column L)
          PF(const long long pf_m = clock64();)
          PF(pf_score += pf_m - pf_t;)
          dpx_acc[h * 2 + sub] += column_sums32(gxs, lane);
          dpy_acc[h * 2 + sub] += column_sums32(gys, lane);
          PF(pf_col += clock64() - pf_m;)
        }
        }
        if (h == 0 && it > 0) {         // the previous tile's dQ (its MMAs finished long ago); frees the accumulator
          PF(pf_t = clock64();)  // before this half's dQ MMAs are issued (they follow pds_ready)
          dq_epilogue(it - 1, m_prev, valid_prev);
          PF(pf_dq += clock64() - pf_t;)
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(pds_ready);
      }
      PF(pf_t = clock64();)
      if (tile + (int)gridDim.x < a.n_tiles) build_yt(tile + gridDim.x, (it + 1) & 1);
      PF(pf_setup += clock64() - pf_t;)
      m_prev = m;
      valid_prev = valid;
    }
    {
      PF(const long long pf_t = clock64();)
      dq_epilogue(it - 1, m_prev, valid_prev);      // the last tile's dQ; also: every MMA of the CTA has completed
      PF(pf_dq += clock64() - pf_t;)
    }
#ifdef DAT_ATTN_BWD_PROFILE
    if (warp == 4 && lane == 0) {
      atomicAdd(&g_attn_bwd_prof[0], (unsigned long long)(clock64() - pf_total));
      atomicAdd(&g_attn_bwd_prof[1], (unsigned long long)pf_wait);
      atomicAdd(&g_attn_bwd_prof[2], (unsigned long long)pf_score);
      atomicAdd(&g_attn_bwd_prof[3], (unsigned long long)pf_col);
      atomicAdd(&g_attn_bwd_prof[4], (unsigned long long)pf_dq);
      atomicAdd(&g_attn_bwd_prof[5], (unsigned long long)pf_setup);
      atomicAdd(&g_attn_bwd_prof[6], 1ull);
    }
#endif

    // ---- end of CTA: dK / dV accumulators, d pos, d table --------------------------------------
    // every MMA of this CTA has completed: the last dq_full commit covers all earlier MMAs
    {
      const uint32_t acc_base = chalf == 0 ? TM_DK : TM_DV;
      float* outp = chalf == 0 ? dk_part : dv_part;
      const float mul = chalf == 0 ? a.scale : 1.0f;
#pragma unroll
      for (int h = 0; h < NHALF; ++h) {
        const int n = h * NHC + row;
        float* dst = outp + (((long long)blockIdx.x * a.B + b) * a.ns_total + n_off + n) * a.C + eta * 32;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t kv[16];
          tmem_ld_32x16(t_lane + acc_base + (uint32_t)(h * 32 + c2 * 16), kv);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(dst + c2 * 16 + i) =
                make_float4(__uint_as_float(kv[i]) * mul, __uint_as_float(kv[i + 1]) * mul,
                            __uint_as_float(kv[i + 2]) * mul, __uint_as_float(kv[i + 3]) * mul);
        }
      }
    }
    // per-warp column sums -> shared memory -> sum over the 4 row-quads
    {
      float* mine = sDpos + (size_t)(warp - 4) * NS * 2;
#pragma unroll
      for (int h = 0; h < NHALF; ++h)
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int n = h * NHC + chalf * 64 + sub * 32 + lane;
          mine[2 * n] = dpy_acc[h * 2 + sub];
          mine[2 * n + 1] = dpx_acc[h * 2 + sub];
        }
    }
    comp_bar_sync();
    for (int n = ctid; n < NS; n += COMP_THREADS) {
      const int ch = (n % NHC) / 64;                 // which column-half group of warps owns n
      float sy = 0.f, sx = 0.f;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const float* src = sDpos + (size_t)(ch * 4 + qd) * NS * 2;
        sy += src[2 * n];
        sx += src[2 * n + 1];
      }
      float* dpo = dpos_part + ((((long long)b * a.heads + eta) * a.chunks + blockIdx.x) * a.ns_total + n_off + n) * 2;
      dpo[0] = sy * (-a.ky / LOG2E);
      dpo[1] = sx * (-a.kx / LOG2E);
    }
    // padded shared-memory table gradient -> global (one atomic per touched in-range cell)
    float* dt_g = d_table + (long long)eta * a.Th * a.Tw;
    for (int i = ctid; TBL && i < a.Hp * a.Wp; i += COMP_THREADS) {
      const int y = i / a.Wp - 2, x = i - (i / a.Wp) * a.Wp - 2;
      float vv = 0.f;
#pragma unroll
      for (int cpy = 0; cpy < NDT; ++cpy) vv += sDTab[cpy * a.Hp * a.Wp + i];
      if (vv != 0.f && y >= 0 && y < a.Th && x >= 0 && x < a.Tw) atomicAdd(dt_g + y * a.Tw + x, vv);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int rows_spanned_max_b(int HW, int W) {
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

}  // namespace

// samples per CTA: the block's samples in chunks of 256 (128 when that is all that divides them)
static int sample_chunk(int Ns) {
  if (Ns <= 256) return Ns;
  static const int off = [] { const char* e = std::getenv("DAT_B200_ATTN_BWD_SIMT_LARGE_NS"); return e && e[0] == '1' ? 1 : 0; }();
  if (off) return 0;            // A/B: the CUDA-core backward for more than 256 samples (round 1)
  return Ns % 256 == 0 ? 256 : (Ns % 128 == 0 ? 128 : 0);
}
int attention_bwd_tc_sample_chunks(const Shape& s) { return sample_chunk(s.Ns) > 0 ? s.Ns / sample_chunk(s.Ns) : 1; }

// sum of the per-sample-chunk dQ slabs (bf16) in fp32, 8 elements per thread
__global__ void __launch_bounds__(256) sum_slabs_bf16_kernel(const bf16* __restrict__ slabs, int nslab, long long count,
                                                             bf16* __restrict__ out) {
  pdl_enter();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= count) return;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int z = 0; z < nslab; ++z) {
    const uint4 raw = *reinterpret_cast<const uint4*>(slabs + (long long)z * count + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      acc[2 * e] += f.x;
      acc[2 * e + 1] += f.y;
    }
  }
  uint4 o;
  __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) ho[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
  *reinterpret_cast<uint4*>(out + i) = o;
}

// scratch for the dQ slabs of a multi-chunk backward (0 when the samples fit one chunk)
size_t attention_bwd_tc_dq_scratch(const Shape& s) {
  const int nsc = attention_bwd_tc_sample_chunks(s);
  return nsc > 1 ? align_up((size_t)nsc * s.B * s.HW * s.C * 2, 256) : 0;
}

int attention_bwd_tc_chunks(const Shape& s) {
  const int pairs = s.B * s.heads * attention_bwd_tc_sample_chunks(s), n_tiles = (s.HW + TQ - 1) / TQ;
  double best_cost = 1e30;
  int best = 1;
  for (int ch = 1; ch <= n_tiles && ch <= 16; ++ch) {
    long long ctas = (long long)pairs * ch;
    long long waves = (ctas + 147) / 148;
    int per = (n_tiles + ch - 1) / ch;
    double cost = (double)waves * (per + 1.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ch; }
  }
  return best;
}

// kernel variant for a shape: private table-gradient copies and 8-byte table entries when they
// fit, else shared copy (atomics) / compact table / single Q,dO slot.  variant < 0: unsupported.
struct BtcVariant { int ok, compact, priv, nslots; uint32_t smem; };
BtcVariant pick_variant(const Shape& s, bool tbl = true) {
  BtcVariant v = {0, 0, 0, 2, 0};
  const int NSc = sample_chunk(s.Ns);
  if (s.act_dtype != DAT_BF16 || !(NSc == 128 || NSc == 256) || s.C % 8 != 0) return v;
  const int rows = rows_spanned_max_b(s.HW, s.W);
  const uint32_t lim = 227 * 1024;
  // private copies: a warp's queries lie in one image row, and runs of equal cells stay within
  // one 8-lane reduction window (table step per query >= 0.2 cells)
  const bool priv_ok = (s.W % 32) == 0 && (s.Tw - 1) * 5 >= (s.W - 1);
  const int tries[4][3] = {{0, 1, 2}, {0, 0, 2}, {1, 0, 2}, {1, 0, 1}};   // {compact, priv, nslots}
  for (int t = 0; t < 4; ++t) {
    if (tries[t][1] && (!priv_ok || !tbl)) continue;
    SmemPlanB sp = plan_smem_b(NSc, s.Th + 3, s.Tw + 3, rows, tries[t][0] != 0, !tbl ? 0 : (tries[t][1] ? 8 : 1), tries[t][2],
                               (!tbl && tries[t][0] == 0) ? 16 : 8);
    if (sp.total <= lim) {
      v.ok = 1; v.compact = tries[t][0]; v.priv = tries[t][1]; v.nslots = tries[t][2]; v.smem = sp.total;
      return v;
    }
  }
  return v;
}

int debug_attn_bwd_timing(unsigned long long* out8) {
  DAT_CUDA_OK(cudaMemcpyFromSymbol(out8, g_attn_bwd_prof, sizeof(unsigned long long) * 8));
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  DAT_CUDA_OK(cudaMemcpyToSymbol(g_attn_bwd_prof, zero, sizeof(zero)));
  return DAT_OK;
}

bool attention_bwd_tc_supported(const Shape& s) { return s.pe_mode == DAT_PE_RPE && pick_variant(s).ok != 0; }
bool attention_bwd_tc_compact_table(const Shape& s, bool tbl) { return pick_variant(s, tbl).compact != 0; }

// packed table in the format the kernel variant of this shape reads: compact {T, T_x+1} entries, the forward-v2
// (mid, dif) entries (dS streamed out, FAST score loop) or the 4-tap entries of the scatter variants
int attention_bwd_pack_table(const Shape& s, const float* table, void* out, bool tbl, cudaStream_t st) {
  const BtcVariant var = pick_variant(s, tbl);
  if (var.compact) return attention_pack_table_compact(s, table, out, st);
  if (!tbl) return attention_pack_table2(s, table, out, st);
  return attention_pack_table(s, table, out, st);
}

int attention_pack_table_compact(const Shape& s, const float* table, void* out, cudaStream_t st) {
  const int ntab = s.heads * (s.Th + 3) * (s.Tw + 3);
  launch_k(pack_table_compact_kernel, ceil_div(ntab, 256), 256, 0, st, table, (uint32_t*)out, s.heads, s.Th, s.Tw);
  DAT_LAUNCH_OK("pack_table_compact_kernel");
  return DAT_OK;
}

// dq (bf16); dk_part / dv_part: (chunks, B, Ns, C) fp32; dpos_part: (B, heads, chunks, Ns, 2);
// d_table must be zeroed by the caller; tab_packed from attention_pack_table().
int attention_bwd_tc(const Shape& s, const void* q, const void* k, const void* v, const void* d_o,
                     const float* lse, const float* delta, const float* pos, const void* tab_packed,
                     void* dq, float* dk_part, float* dv_part, float* d_table, float* dpos_part,
                     cudaStream_t st, void* ds_out, void* dq_scratch) {
  const bool tbl = ds_out == nullptr;      // dS streamed out: the table gradient is formed by rpe_table_grad_mma
  const BtcVariant var = pick_variant(s, tbl);
  DAT_REQUIRE(var.ok, "attention_bwd_tc: unsupported shape");
  const int NSc = sample_chunk(s.Ns), nsc = s.Ns / NSc;
  DAT_REQUIRE(nsc == 1 || dq_scratch != nullptr, "attention_bwd_tc: more than one sample chunk needs the dQ scratch");
  BtcArgs a;
  a.ns_total = s.Ns;
  a.dq_slab = (long long)s.B * s.HW * s.C;
  a.nslots = var.nslots;
  {   // light table-gradient path: 0.29 <= table cells per query step <= 1 (runs of <= 3.5 lanes)
    const float step = 0.5f * (float)(s.Tw - 1) / (float)(s.W - 1);
    a.light_table = step >= 0.2857f && step <= 1.0f && std::getenv("DAT_B200_ATTN_BWD_GENERIC_TABLE") == nullptr;
  }
  a.B = s.B; a.H = s.H; a.W = s.W; a.HW = s.HW; a.C = s.C; a.heads = s.heads; a.G = s.G; a.hg = s.hg;
  a.Th = s.Th; a.Tw = s.Tw; a.Wp = s.Tw + 3; a.Hp = s.Th + 3;
  a.n_tiles = (s.HW + TQ - 1) / TQ;
  a.rows_max = rows_spanned_max_b(s.HW, s.W);
  a.chunks = attention_bwd_tc_chunks(s);
  a.scale = 1.0f / sqrtf((float)DAT_HEAD_DIM);
  a.c1 = a.scale * LOG2E;
  a.kx = 0.25f * (float)(s.Tw - 1);
  a.ky = 0.25f * (float)(s.Th - 1);
  a.gsx = 2.0f / (float)(s.W - 1);
  a.gsy = 2.0f / (float)(s.H - 1);
  CUtensorMap tmQ, tmDO, tmK, tmV;
  const uint64_t pitch = (uint64_t)s.C * 2;
  DAT_FWD(tc::make_tmap_2d(&tmQ, q, 2, false, (uint64_t)s.B * s.HW, (uint64_t)s.C, pitch, TQ, 32, 64));
  DAT_FWD(tc::make_tmap_2d(&tmDO, d_o, 2, false, (uint64_t)s.B * s.HW, (uint64_t)s.C, pitch, TQ, 32, 64));
  DAT_FWD(tc::make_tmap_2d(&tmK, k, 2, false, (uint64_t)s.B * s.Ns, (uint64_t)s.C, pitch, NSc, 32, 64));
  DAT_FWD(tc::make_tmap_2d(&tmV, v, 2, false, (uint64_t)s.B * s.Ns, (uint64_t)s.C, pitch, NSc, 32, 64));
  dim3 grid(a.chunks, s.B * s.heads, nsc);
  void* dq_dst = nsc > 1 ? dq_scratch : dq;
#define LAUNCH(NSV, CP, PV, TB)                                                                  \
  do {                                                                                           \
    auto kern = attn_bwd_tc_kernel<NSV, CP, PV, TB>;                                             \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var.smem)); \
    launch_k(kern, grid, BTC_THREADS, var.smem, st, tmQ, tmDO, tmK, tmV, pos, tab_packed, lse, delta,  \
                                              (bf16*)dq_dst, dk_part, dv_part, d_table, dpos_part, (bf16*)ds_out, a); \
  } while (0)
#define LAUNCH_NS(NSV)                                   \
  do {                                                   \
    if (!tbl && var.compact) LAUNCH(NSV, true, false, false);  \
    else if (!tbl) LAUNCH(NSV, false, false, false);     \
    else if (var.compact) LAUNCH(NSV, true, false, true); \
    else if (var.priv) LAUNCH(NSV, false, true, true);   \
    else LAUNCH(NSV, false, false, true);                \
  } while (0)
  if (NSc == 256) LAUNCH_NS(256); else LAUNCH_NS(128);
#undef LAUNCH_NS
#undef LAUNCH
  DAT_LAUNCH_OK("attn_bwd_tc_kernel");
  if (nsc > 1) {
    const long long count = a.dq_slab;
    DAT_REQUIRE(count % 8 == 0, "attention_bwd_tc: B * HW * C must be a multiple of 8");
    launch_k(sum_slabs_bf16_kernel, ceil_div(count / 8, 256), 256, 0, st, (const bf16*)dq_scratch, nsc, count, (bf16*)dq);
    DAT_LAUNCH_OK("sum_slabs_bf16_kernel");
  }
  return DAT_OK;
}

}  // namespace dat
