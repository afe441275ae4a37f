// Bilinear gather of the deformed key / value features FUSED with the k / v projections (dat_blocks.py:169-178:
// `x_sampled = F.grid_sample(x, pos)`, `k = proj_k(x_sampled)`, `v = proj_v(x_sampled)`): one launch instead of three.
//
// A CTA owns 128 consecutive samples (rows of the (B * Ns, C) sampled matrix).  Its eight gather / epilogue warps
// build the A operand directly in shared memory - every (sample, 8-channel chunk) is the 4-tap blend of channel-last
// rows of x (exact ATen index arithmetic, the same code as sample_fwd_kernel: x_sampled stays bit-identical), rounded
// to bf16 and stored as one 16-byte piece at its place in the canonical K-major 128B-swizzled layout tcgen05.mma
// reads (C / 64 blocks of [128 rows x 128 bytes]); the same 16 bytes go to `xs` in global memory, which the backward
// needs (weight gradients of proj_k / proj_v, d pos).  The sampled tile never makes the HBM round trip between the
// gather and the projections.  Then both projections run against the resident tile: the weight panels of W_k and W_v
// stream through a TMA ring, accumulators are double-buffered in tensor memory, and k / v leave through the TMA-store
// epilogue of gemm_tc.cu (bias, bf16, 128B-swizzled staging, one cp.async.bulk.tensor store per warp and 64 columns).
// Warp roles: 0 = TMA producer (weights), 1 = TMEM allocation + MMA issuer, 2-9 = gather, then epilogue.
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace dat {

namespace {

using namespace tc;

constexpr int GK_THREADS = 320;
constexpr int GK_GWARPS = 8;
constexpr int GK_BM = 128;

struct GkArgs {
  int B, H, W, C, G, Cg, Ns;
  long long M;          // B * Ns sampled rows
  int KC;               // C / 64: K chunks (128-byte rows) of the resident A tile
  int BN, NT;           // output-column tile and tiles per projection (C / BN)
  int stages, tmem_cols, bias_bytes;
};

__device__ __forceinline__ uint32_t gk_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void gk_load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void gk_load8(const bf16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename TX>
__global__ void __launch_bounds__(GK_THREADS, 1)
gather_kv_tc_kernel(const __grid_constant__ CUtensorMap tmWk, const __grid_constant__ CUtensorMap tmWv,
                    const __grid_constant__ CUtensorMap tmKo, const __grid_constant__ CUtensorMap tmVo,
                    const TX* __restrict__ x, const float* __restrict__ pos, const float* __restrict__ bk,
                    const float* __restrict__ bv, bf16* __restrict__ xs, GkArgs a) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int b_stage_bytes = a.BN * 128;
  // [barriers 1 KB][bias][A: KC x 16 KB][B ring][epilogue staging 8 x 4 KB (first: the tile's sampling positions)]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + a.stages;
  uint64_t* acc_full = empty + a.stages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* a_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);
  float* sBias = reinterpret_cast<float*>(smem + 1024);
  uint8_t* sA = smem + 1024 + a.bias_bytes;
  uint8_t* sB = sA + a.KC * (GK_BM * 128);
  uint8_t* sStage = sB + a.stages * b_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GK_BM;
  // blockIdx.y = 0: the k projection (this CTA also writes xs), 1: the v projection.  Both CTAs gather the tile (the
  // taps come from L2 the second time): with one CTA per 128 samples for both projections only B * Ns / 128 = 32 SMs
  // worked at the DAT-T++ shapes and the fused launch was slower than the three it replaces.
  const bool is_k = blockIdx.y == 0;
  const int n_tiles = a.NT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWk);
    tma_prefetch_desc(&tmWv);
    tma_prefetch_desc(&tmKo);
    tma_prefetch_desc(&tmVo);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], GK_GWARPS);
    }
    mbar_init(a_full, GK_GWARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const CUtensorMap* tm = is_k ? &tmWk : &tmWv;
        const int n0 = t * a.BN;
        for (int kc = 0; kc < a.KC; ++kc, ++it) {
          const int s = it % a.stages;
          mbar_wait(&empty[s], ((uint32_t)(it / a.stages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full[s], (uint32_t)b_stage_bytes);
          tma_load_2d(sB + s * b_stage_bytes, tm, &full[s], kc * 64, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(FMT_BF16, GK_BM, (uint32_t)a.BN);
      mbar_wait(a_full, 0);                          // the gathered tile is complete (and visible to the async proxy)
      tc_fence_after_sync();
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&acc_empty[buf], (uint32_t)((t >> 1) & 1) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * a.BN);
        for (int kc = 0; kc < a.KC; ++kc, ++it) {
          const int s = it % a.stages;
          mbar_wait(&full[s], (uint32_t)(it / a.stages) & 1u);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + kc * (GK_BM * 128));
          const uint32_t b_addr = smem_u32(sB + s * b_stage_bytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t ad = make_smem_desc(a_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
            const uint64_t bd = make_smem_desc(b_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
            mma_bf16_ss(d_tmem, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[buf]);
      }
    }
  } else {
    // ---- gather: (row, 8-channel chunk) units, four per thread and pass (16-32 tap loads in flight) --------------------
    const int gt = threadIdx.x - 64;                 // 0 .. 255
    {
      const float* bsrc = is_k ? bk : bv;
      for (int i = gt; i < a.C; i += 32 * GK_GWARPS) sBias[i] = bsrc != nullptr ? bsrc[i] : 0.f;
    }
    // sampling positions of the tile's rows, all groups: one round trip to memory for the whole tile instead of one per
    // unit in front of its tap loads.  sPos[row][g] lives in the epilogue staging area (free until the first store).
    float2* sPos = reinterpret_cast<float2*>(sStage);
    for (int i = gt; i < GK_BM * a.G; i += 32 * GK_GWARPS) {
      const int row = i / a.G, g = i - row * a.G;
      const long long m = (long long)m0 + row;
      float2 pp = make_float2(0.f, 0.f);
      if (m < a.M) {
        const int b = (int)(m / a.Ns), n = (int)(m - (long long)b * a.Ns);
        pp = *reinterpret_cast<const float2*>(pos + (((long long)b * a.G + g) * a.Ns + n) * 2);
      }
      sPos[i] = pp;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // positions and bias staged
    const int lpr = a.C >> 3;                        // chunks per row
    const int units = GK_BM * lpr;
    // Per pass a thread owns UQ units.  All their tap loads are issued before the first blend: the addresses of
    // out-of-range taps are clamped into the image and the loaded VALUE is replaced by zero (fmaf(0, w, acc) == acc,
    // so x_sampled stays bit-identical to sample_fwd_kernel, which skips such taps).  With a branch per tap the 16 tap
    // loads of a pass formed one dependent chain (~10 us per pass: the first version of this kernel took 50 / 88 us
    // at stages 2 / 3 where the three separate launches take 22 / 29).
    constexpr int UQ = 2;
    for (int u0 = gt; u0 < units; u0 += UQ * 32 * GK_GWARPS) {
      float acc[UQ][8], tv[UQ][4][8], tw[UQ][4];
      bool tok[UQ][4];
      int row_[UQ], ch_[UQ];
      bool live[UQ];
#pragma unroll
      for (int q = 0; q < UQ; ++q) {
        const int u = u0 + q * 32 * GK_GWARPS;
        const int row = u / lpr, ch = u - row * lpr;
        row_[q] = row; ch_[q] = ch;
        const long long m = (long long)m0 + row;
        live[q] = u < units && m < a.M;
        const int b = live[q] ? (int)(m / a.Ns) : 0;
        const int c = ch * 8, g = live[q] ? c / a.Cg : 0;
        const float2 pp = live[q] ? sPos[row * a.G + g] : make_float2(0.f, 0.f);
        const Taps t = make_taps(pp.y, pp.x, a.W, a.H);
        const TX* xb = x + (long long)b * a.H * a.W * a.C + (live[q] ? c : 0);
        const int yc0 = min(max(t.y0, 0), a.H - 1), yc1 = min(max(t.y0 + 1, 0), a.H - 1);
        const int xc0 = min(max(t.x0, 0), a.W - 1), xc1 = min(max(t.x0 + 1, 0), a.W - 1);
        tok[q][0] = live[q] && t.vx0 && t.vy0; tw[q][0] = __fmul_rn(t.wx0, t.wy0);
        tok[q][1] = live[q] && t.vx1 && t.vy0; tw[q][1] = __fmul_rn(t.wx1, t.wy0);
        tok[q][2] = live[q] && t.vx0 && t.vy1; tw[q][2] = __fmul_rn(t.wx0, t.wy1);
        tok[q][3] = live[q] && t.vx1 && t.vy1; tw[q][3] = __fmul_rn(t.wx1, t.wy1);
        gk_load8(xb + ((long long)yc0 * a.W + xc0) * a.C, tv[q][0]);
        gk_load8(xb + ((long long)yc0 * a.W + xc1) * a.C, tv[q][1]);
        gk_load8(xb + ((long long)yc1 * a.W + xc0) * a.C, tv[q][2]);
        gk_load8(xb + ((long long)yc1 * a.W + xc1) * a.C, tv[q][3]);
      }
#pragma unroll
      for (int q = 0; q < UQ; ++q) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[q][e] = 0.f;
#pragma unroll
        for (int tp = 0; tp < 4; ++tp)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[q][e] = fmaf(tok[q][tp] ? tv[q][tp][e] : 0.f, tw[q][tp], acc[q][e]);
      }
#pragma unroll
      for (int q = 0; q < UQ; ++q) {
        if (u0 + q * 32 * GK_GWARPS >= units) continue;
        const uint4 pk = make_uint4(gk_pack(acc[q][0], acc[q][1]), gk_pack(acc[q][2], acc[q][3]),
                                    gk_pack(acc[q][4], acc[q][5]), gk_pack(acc[q][6], acc[q][7]));
        const int row = row_[q], ch = ch_[q];
        // K chunk ch / 8, 16-byte piece ch % 8 of the row, 128B swizzle: piece ^ (row % 8)
        uint8_t* dst = sA + (ch >> 3) * (GK_BM * 128) + row * 128 + (((ch & 7) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = pk;         // rows beyond M are zero: their outputs are clipped by the store
        if (live[q] && is_k) *reinterpret_cast<uint4*>(xs + ((long long)m0 + row) * a.C + ch * 8) = pk;
      }
    }
    fence_proxy_async_smem();                        // generic-proxy writes of the tile -> visible to tcgen05.mma
    __syncwarp();
    if (lane == 0) mbar_arrive(a_full);
    asm volatile("bar.sync 1, 256;" ::: "memory");   // every warp has read its positions: the staging area is free

    // ---- epilogue (as gemm_tc_persistent_kernel: TMEM -> + bias -> bf16 -> swizzled staging -> TMA store) -------------
    const int quad = warp & 3, chalf = (warp - 2) >> 2;
    uint8_t* tstage = sStage + (warp - 2) * 4096;
    const uint32_t srow = smem_u32(tstage) + (uint32_t)lane * 128u;
    const uint32_t sxor = (uint32_t)(lane & 7);
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int n0 = t * a.BN;
      mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * a.BN);
      const int cg_first = chalf * 64;
      if (cg_first >= a.BN) {              // a single 64-column group: this warp only hands the buffer back
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      for (int cg = cg_first; cg < a.BN; cg += 128) {
        uint32_t r[2][32];
        tmem_ld_32x32(t_addr + (uint32_t)cg, r[0]);
        tmem_ld_32x32(t_addr + (uint32_t)(cg + 32), r[1]);
        tmem_wait_ld();
        if (cg + 128 >= a.BN) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
        const float* bp = sBias + n0 + cg;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(bp + c * 32 + j);
            const float4 b1 = *reinterpret_cast<const float4*>(bp + c * 32 + j + 4);
            const uint32_t chunk = (uint32_t)(c * 4 + (j >> 3));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((chunk ^ sxor) << 4)),
                         "r"(gk_pack(__uint_as_float(r[c][j]) + b0.x, __uint_as_float(r[c][j + 1]) + b0.y)),
                         "r"(gk_pack(__uint_as_float(r[c][j + 2]) + b0.z, __uint_as_float(r[c][j + 3]) + b0.w)),
                         "r"(gk_pack(__uint_as_float(r[c][j + 4]) + b1.x, __uint_as_float(r[c][j + 5]) + b1.y)),
                         "r"(gk_pack(__uint_as_float(r[c][j + 6]) + b1.z, __uint_as_float(r[c][j + 7]) + b1.w))
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && (long long)m0 + quad * 32 < a.M) {
          tma_store_2d(is_k ? &tmKo : &tmVo, tstage, n0 + cg, m0 + quad * 32);
          bulk_commit_group();
        }
      }
    }
    // the staging tile must have been read before the CTA's shared memory goes away; the global writes themselves
      // complete asynchronously (they are ordered before the end of the grid, like any other store)
      if (lane == 0) bulk_wait_group_read<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

bool gk_plan(const Shape& s, GkArgs* a, size_t* smem) {
  if (s.act_dtype != DAT_BF16 || s.no_off) return false;
  if (!(s.C == 64 || s.C == 128 || s.C == 256 || s.C == 512) || s.Cg % 8 != 0) return false;
  a->B = s.B; a->H = s.H; a->W = s.W; a->C = s.C; a->G = s.G; a->Cg = s.Cg; a->Ns = s.Ns;
  a->M = (long long)s.B * s.Ns;
  if (a->M >= (1ll << 31) - GK_BM) return false;
  a->KC = s.C / 64;
  a->BN = s.C <= 256 ? s.C : 128;        // C = 512: the resident tile takes 128 KB, narrower weight tiles keep a 3-deep ring
  a->NT = s.C / a->BN;
  a->bias_bytes = (int)align_up((size_t)s.C * 4, 1024);
  const size_t fixed = 1024 + 1024 + (size_t)a->bias_bytes + (size_t)a->KC * GK_BM * 128 + (size_t)GK_GWARPS * 4096;
  const size_t bstage = (size_t)a->BN * 128;
  if (fixed + 2 * bstage > 226 * 1024) return false;
  int st = (int)((226 * 1024 - fixed) / bstage);
  if (st > 4) st = 4;
  a->stages = st;
  a->tmem_cols = 32;
  while (a->tmem_cols < 2 * a->BN) a->tmem_cols <<= 1;
  *smem = fixed + (size_t)st * bstage;
  return true;
}

}  // namespace

// Measured on B200 (ncu, cold L2, batch 16, DAT-T++ stages 0-3: C = 64 / 128 / 256 / 512): this kernel 11.9 / 16.5 /
// 27.0 / 53.5 us against 16.0 / 19.3 / 22.4 / 28.7 us for sample_fwd + the two GEMM launches.  A 128-sample tile pulls
// C x 2 KB of fp32 taps into ONE SM (1 MB at C = 512) at the 20-30 GB/s a single SM gets out of a latency-bound
// gather, where the stand-alone gather spreads the same bytes over all 148 SMs - and B * Ns / 128 x 2 = 64 CTAs is
// all the parallelism the 4096 sampled rows of a batch offer.  So the fused path is the default where it wins
// (C <= 128); DAT_B200_GATHER_KV_FUSION=1 forces it for every supported width, =0 switches it off.
bool gather_kv_tc_supported(const Shape& s) {
  static const int mode = [] { const char* e = std::getenv("DAT_B200_GATHER_KV_FUSION"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
  GkArgs a;
  size_t smem;
  if (mode == 0 || !gk_plan(s, &a, &smem)) return false;
  return mode == 1 || s.C <= 128;
}

// xs (B, Ns, C) bf16 = bilinear samples of x at pos; k = xs Wk^T + bk, v = xs Wv^T + bv (bf16).  wk / wv: bf16 (C, C).
int gather_kv_tc(const Shape& s, const void* x, const float* pos, const void* wk, const void* wv, const float* bk,
                 const float* bv, void* xs, void* k, void* v, cudaStream_t st) {
  GkArgs a;
  size_t smem;
  DAT_REQUIRE(gk_plan(s, &a, &smem), "gather_kv_tc: unsupported shape");
  CUtensorMap tmWk, tmWv, tmKo, tmVo;
  const uint64_t C = (uint64_t)s.C;
  DAT_FWD(tc::make_tmap_2d(&tmWk, wk, 2, false, C, C, C * 2, (uint32_t)a.BN, 64, 128));
  DAT_FWD(tc::make_tmap_2d(&tmWv, wv, 2, false, C, C, C * 2, (uint32_t)a.BN, 64, 128));
  DAT_FWD(tc::make_tmap_2d(&tmKo, k, 2, false, (uint64_t)a.M, C, C * 2, 32, 64, 128));
  DAT_FWD(tc::make_tmap_2d(&tmVo, v, 2, false, (uint64_t)a.M, C, C * 2, 32, 64, 128));
  const dim3 grid((unsigned)((a.M + GK_BM - 1) / GK_BM), 2);
#define LAUNCH(TX)                                                                                              \
  do {                                                                                                          \
    auto kern = gather_kv_tc_kernel<TX>;                                                                        \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    launch_k(kern, grid, GK_THREADS, smem, st, tmWk, tmWv, tmKo, tmVo, (const TX*)x, pos, bk, bv, (bf16*)xs, a); \
  } while (0)
  if (s.x_dtype == DAT_F32) LAUNCH(float); else LAUNCH(bf16);
#undef LAUNCH
  DAT_LAUNCH_OK("gather_kv_tc_kernel");
  return DAT_OK;
}

}  // namespace dat
