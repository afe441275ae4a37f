// 1x1-convolution GEMMs of the block on the 5th-generation tensor cores (tcgen05):
//     Y[M, N] = X[M, K] * W[N, K]^T + b[N]        (proj_q / proj_k / proj_v / proj_out,
//                                                  dat_blocks.py:143,177-178,225)
// Both operands are K-major and are staged by TMA (128-byte swizzle) into a 4-stage
// shared-memory ring; one elected thread issues tcgen05.mma (M = 128, N = tile width,
// K step = 32 bytes) with the fp32 accumulator in tensor memory; four epilogue warps read
// it back with tcgen05.ld, add the bias, convert and store channel-last rows.
//
// Two operand kinds share the code:
//   bf16 : X and W are bf16 (W pre-cast once per call into the workspace).
//   tf32 : X is the fp32 block input straight from the preceding LayerNorm and W the fp32
//          parameter — no cast pass at all; x is read exactly once at 4 B/element, which
//          is what bounds proj_q (HBM-bound: 4 B in + 2 B out per element, ~64 FLOP/B).
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace dat {

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;   // benign race: every thread resolves the same pointer
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, bool is_float32,
                 uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows,
                 uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn enc = get_encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return DAT_ERR_CUDA;
  }
  DAT_REQUIRE(((uintptr_t)base & 15) == 0 && (pitch_bytes & 15) == 0, "TMA: base / pitch must be 16-byte aligned");
  DAT_REQUIRE(box_rows <= 256 && box_cols <= 256, "TMA: box too large");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dt = is_float32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return DAT_ERR_CUDA;
  }
  return DAT_OK;
}

}  // namespace tc

// Phase timestamps (globaltimer ns) of CTA (0,0) of the last tensor-core GEMM launch:
// [0] entry, [1] setup done, [2] first stage landed, [3] MMAs issued, [4] accumulator
// ready, [5] epilogue done.  Debug aid read back by dat_debug_gemm_timing().
__device__ unsigned long long g_gemm_timing[8];

namespace {

using namespace tc;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TSTAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_gemm_timing[i] = gtime(); } while (0)

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 192;
constexpr int TCP_THREADS = 320;          // persistent kernel: producer, MMA issuer, 8 epilogue warps
constexpr int TCP_EPI_WARPS = 8;
constexpr int CHUNK_BYTES = 128;          // K bytes per pipeline stage row (one swizzle atom)
constexpr int A_STAGE_BYTES = TC_BM * CHUNK_BYTES;
constexpr int EPI_COLS = 128;            // epilogue column group (staging area per warp: 32 x 128 outputs)
constexpr int SMEM_BUDGET = 108 * 1024;   // per CTA: two CTAs share an SM

template <bool TF32, typename TOut>
__global__ void __launch_bounds__(TC_THREADS, 4)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const float* __restrict__ bias, TOut* __restrict__ Y, int M, int N, int k_chunks,
               int k_chunks1, int BN, int stages, int tmem_cols) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int b_stage_bytes = BN * CHUNK_BYTES;
  // [barriers: 1 KB][A stages][B stages]; the stage buffers double as the epilogue's
  // output staging area once the accumulator is complete
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sBias = reinterpret_cast<float*>(smem + 1024);   // BN floats (<= 1 KB)
  uint8_t* sA = smem + 2048;
  uint8_t* sB = sA + stages * A_STAGE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  constexpr int CHUNK_ELEMS = TF32 ? 32 : 64;
  if (threadIdx.x == 0) TSTAMP(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (k_chunks1 < k_chunks) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  // bias -> shared memory once: a global load inside the epilogue loop cannot be hoisted
  // above the tcgen05.ld waits and cost one L2 round trip per 32-column chunk (5.5 us/tile)
  for (int i = threadIdx.x; i < BN; i += TC_THREADS) sBias[i] = bias != nullptr ? bias[n0 + i] : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TSTAMP(1);

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < k_chunks; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (uint32_t)(kc / stages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], (uint32_t)(A_STAGE_BYTES + b_stage_bytes));
        // optional second operand pair: Y = X1 W1^T + X2 W2^T as one K-concatenated GEMM
        const bool second = kc >= k_chunks1;
        const int kcol = (second ? kc - k_chunks1 : kc) * CHUNK_ELEMS;
        tma_load_2d(sA + s * A_STAGE_BYTES, second ? &tmA2 : &tmA, &full[s], kcol, m0);
        tma_load_2d(sB + s * b_stage_bytes, second ? &tmB2 : &tmB, &full[s], kcol, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(TF32 ? FMT_TF32 : FMT_BF16, TC_BM, (uint32_t)BN);
      for (int kc = 0; kc < k_chunks; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (uint32_t)(kc / stages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        if (kc == 0) TSTAMP(2);
        const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(sB + s * b_stage_bytes);
#pragma unroll
        for (int k4 = 0; k4 < CHUNK_BYTES / 32; ++k4) {
          const uint64_t ad = make_smem_desc(a_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
          const uint64_t bd = make_smem_desc(b_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
          if (TF32) mma_tf32_ss(tmem_base, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
          else mma_bf16_ss(tmem_base, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
        }
        tc_commit(&empty[s]);        // frees the smem slot when these MMAs have read it
      }
      tc_commit(tmem_full);          // accumulator complete
      TSTAMP(3);
    }
  } else {
    const int quad = warp & 3;       // TMEM lane quadrant this warp may access
    const int row = m0 + quad * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after_sync();
    if (warp == 2 && lane == 0) TSTAMP(4);
    // TMEM -> registers (+bias, convert) -> this warp's 32-row staging tile in shared memory (row
    // pitch padded by 16 B: conflict-free 16-byte accesses) -> global memory one row segment per
    // store instruction (coalesced 256-512 B runs).  Column groups of EPI_COLS keep the staging
    // area small enough for two CTAs per SM, whose load / MMA / store phases then overlap.
    const int gcols = BN < EPI_COLS ? BN : EPI_COLS;
    const int seg_bytes = gcols * (int)sizeof(TOut);
    const int pitch = seg_bytes + 16;
    uint8_t* stage = sA + (warp - 2) * 32 * pitch;
    (void)row;
    const int rows_here = min(32, M - (m0 + quad * 32));
    for (int cg = 0; cg < BN; cg += gcols) {
      for (int c = 0; c < gcols / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cg + c * 32), r);
        tmem_wait_ld();
        const float* bp = sBias + cg + c * 32;
        TOut* dst = reinterpret_cast<TOut*>(stage + lane * pitch) + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                 __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          const float4 bb = *reinterpret_cast<const float4*>(bp + j);
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          store4(dst + j, v);
        }
      }
      __syncwarp();
      for (int rr = 0; rr < rows_here; ++rr) {
        uint8_t* grow = reinterpret_cast<uint8_t*>(Y + (long long)(m0 + quad * 32 + rr) * N + n0 + cg);
        const uint8_t* srow = stage + rr * pitch;
        for (int off = lane * 16; off < seg_bytes; off += 512)
          *reinterpret_cast<uint4*>(grow + off) = *reinterpret_cast<const uint4*>(srow + off);
      }
      __syncwarp();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) TSTAMP(5);
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store8(float* p, float4 a, float4 b) {
  *reinterpret_cast<float4*>(p) = a;
  *reinterpret_cast<float4*>(p + 4) = b;
}
__device__ __forceinline__ void store8(bf16* p, float4 a, float4 b) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(a.x, a.y); h[1] = __floats2bfloat162_rn(a.z, a.w);
  h[2] = __floats2bfloat162_rn(b.x, b.y); h[3] = __floats2bfloat162_rn(b.z, b.w);
  *reinterpret_cast<uint4*>(p) = raw;
}

// Persistent version: one CTA per SM walks the output tiles (m fastest, so neighbouring CTAs share
// the weight tile in L2).  The TMA producer never drains its ring between tiles, the fp32
// accumulator is double-buffered in TMEM (2 x BN columns), and the eight epilogue warps store tile
// i while the MMA warp accumulates tile i + 1: loads, MMAs and stores of a CTA overlap instead of
// alternating (the one-tile kernel idles its loads ~60 % of a CTA's lifetime, profiles/r01_ncu_kernels.md).
//   barriers: full/empty per ring stage, acc_full/acc_empty per TMEM buffer.
// column group of the persistent kernel's epilogue: 64 columns when the tile is a multiple of 64 wide (two or four
// groups shared by a quadrant's two warps), else the whole tile (32 / 96 columns, one warp)
__host__ __device__ inline int persistent_gcols(int BN) { return BN % 64 == 0 ? 64 : BN; }

// Epilogue: 8 warps, two per TMEM lane quadrant, taking alternate column groups of a tile.  Tiles that are a multiple
// of 64 columns wide leave through TMA stores (TMEM -> + bias [-> fused residual] -> bf16 / fp32 -> 128B-swizzled
// staging -> one cp.async.bulk.tensor store per warp and group); the round-1 epilogue (padded staging, ld.shared +
// st.global write-out) remains for 32 / 96-column tiles.  The write-out bounded every wide-N / small-K shape: with the
// epilogue switched off the stage-2 fc1 GEMM takes 12.7 us, with the old one 21.0, with TMA stores 15.2
// (profiles/r02_gemm_decompose.md).  One launch can also form two products of the same A (N1 < N: the k and v
// projections).
// BMN: the B operand is MN-major - a (K, N) row-major matrix, i.e. the weight W (N_fwd, K_fwd) itself when the
// product is the data gradient dX = dY W: 64 x 64 TMA boxes (128B swizzle) are the canonical MN-major core-matrix
// layout (8-row groups 1 KB apart, 64-column blocks 8 KB apart, gemm_tc_wgrad.cu), so no transposed copy of the
// weight is ever made.  bf16 only, BN a multiple of 64.
template <bool TF32, typename TOut, bool BMN = false>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                          const __grid_constant__ CUtensorMap tmY, const float* __restrict__ bias,
                          TOut* __restrict__ Y, int M, int N, int k_chunks, int k_chunks1, int BN, int stages,
                          int tmem_cols, int m_tiles, int total_tiles, int stage_pitch, int bias_bytes, int dbg,
                          const __grid_constant__ CUtensorMap tmY2, const float* __restrict__ bias2, int N1,
                          const float* __restrict__ resid, const float* __restrict__ rscale, int rows_per_sample) {
  // resid != NULL (fp32 output): the residual add with stochastic depth that follows the product rides along in the
  // epilogue - Y = resid + rscale[row / rows_per_sample] * bf16(X W^T + bias), i.e. `x = drop_path(mlp(x)) + x`
  // (dat.py:151-156) with the branch rounded to bf16 as the autocast convolution returns it.
  // N1 < N: TWO products of the same A in one launch (the k and v projections of the sampled features): output
  // columns [0, N1) are X W^T + bias into tmY, columns [N1, N) are X W2^T + bias2 into tmY2 (W2 behind tmB2, which is
  // otherwise the second K-concatenated source; the two uses exclude each other).  N1 is a multiple of BN.
  pdl_enter();
  // dbg (DAT_B200_GEMM_DBG, timing decomposition only - results are wrong): 1 = no global stores, 2 = B loaded for the
  // CTA's first tile only, 4 = A loaded for the first tile only, 8 = epilogue only hands the accumulator back
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int b_stage_bytes = BN * CHUNK_BYTES;
  // [barriers 1 KB][bias: all N floats, staged once per CTA][ring: A stages | B stages (1 KB aligned)][epilogue staging]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sBias = reinterpret_cast<float*>(smem + 1024);
  uint8_t* sA = smem + 1024 + bias_bytes;
  uint8_t* sB = sA + stages * A_STAGE_BYTES;
  uint8_t* sStage = sB + stages * b_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CHUNK_ELEMS = TF32 ? 32 : 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (k_chunks1 < k_chunks) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    if (BN % 64 == 0) tma_prefetch_desc(&tmY);
    if (N1 < N) tma_prefetch_desc(&tmY2);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], TCP_EPI_WARPS);       // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t % m_tiles) * TC_BM, n0 = (t / m_tiles) * BN;
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const bool first_tile = t == (int)blockIdx.x;
          const bool ld_a = first_tile || !(dbg & 4), ld_b = first_tile || !(dbg & 2);
          if (!ld_a && !ld_b) { mbar_arrive(&full[s]); continue; }
          mbar_arrive_expect_tx(&full[s], (uint32_t)((ld_a ? A_STAGE_BYTES : 0) + (ld_b ? b_stage_bytes : 0)));
          const bool second = kc >= k_chunks1;
          const bool out2 = n0 >= N1;                       // second product: its weights sit behind tmB2
          const int n0w = out2 ? n0 - N1 : n0;
          const int kcol = (second ? kc - k_chunks1 : kc) * CHUNK_ELEMS;
          if (ld_a) tma_load_2d(sA + s * A_STAGE_BYTES, second ? &tmA2 : &tmA, &full[s], kcol, m0);
          if (!ld_b) continue;
          if (BMN) {
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sB + s * b_stage_bytes + i * 8192, (second || out2) ? &tmB2 : &tmB, &full[s], n0w + 64 * i, kcol);
          } else {
            tma_load_2d(sB + s * b_stage_bytes, (second || out2) ? &tmB2 : &tmB, &full[s], kcol, n0w);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(TF32 ? FMT_TF32 : FMT_BF16, TC_BM, (uint32_t)BN, 0, BMN ? 1u : 0u);
      int it = 0, li = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
        const int buf = li & 1;
        mbar_wait(&acc_empty[buf], (uint32_t)((li >> 1) & 1) ^ 1u);   // epilogue has drained this buffer
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * b_stage_bytes);
#pragma unroll
          for (int k4 = 0; k4 < CHUNK_BYTES / 32; ++k4) {
            const uint64_t ad = make_smem_desc(a_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
            const uint64_t bd = BMN ? make_smem_desc(b_addr + k4 * 2048, 8192, 1024, LAYOUT_SW128)
                                    : make_smem_desc(b_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
            if (TF32) mma_tf32_ss(d_tmem, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
            else mma_bf16_ss(d_tmem, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[buf]);
      }
    }
  } else {
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;   // which of the quadrant's two warps: takes column groups chalf, chalf + 2, ...
    const int gcols = persistent_gcols(BN);
    const int seg_bytes = gcols * (int)sizeof(TOut);
    uint8_t* stage = sStage + (warp - 2) * 32 * stage_pitch;
    int li = 0;
    // the whole bias vector is staged once per CTA by the epilogue warps alone: the producer and the MMA issuer start
    // their first tile without waiting for this global load
    for (int i = threadIdx.x - 64; i < N; i += 32 * TCP_EPI_WARPS) {
      const float* bsrc = i < N1 ? bias : bias2;
      sBias[i] = bsrc != nullptr ? bsrc[i < N1 ? i : i - N1] : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");     // the eight epilogue warps
    if (BN % 64 == 0) {
      // TMA-store epilogue: a warp's 32 rows x 64 columns go TMEM -> registers -> (+ bias, convert) -> a 128B-swizzled
      // shared-memory tile (conflict-free 16-byte stores) -> ONE cp.async.bulk.tensor store issued by lane 0.  The
      // store is asynchronous: the warp only waits (wait_group.read) until the tile has been read out of shared memory
      // before it stages the next group, and rows beyond M are clipped by the tensor map.
      constexpr int BOX_COLS = 128 / (int)sizeof(TOut);               // 64 bf16 / 32 fp32 columns = one 128-byte row
      constexpr int NBOX = 64 / BOX_COLS;
      uint8_t* tstage = sStage + (warp - 2) * (NBOX * 4096);
      const uint32_t srow = smem_u32(tstage) + (uint32_t)lane * 128u;
      const uint32_t sxor = (uint32_t)(lane & 7);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
        const int buf = li & 1;
        const int m0 = (t % m_tiles) * TC_BM, n0 = (t / m_tiles) * BN;
        mbar_wait(&acc_full[buf], (uint32_t)((li >> 1) & 1));
        tc_fence_after_sync();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
        const int cg_first = chalf * 64;
        // fused residual: this lane's row of the residual stream (NULL beyond M) and its sample's scale
        const int my_row = m0 + quad * 32 + lane;
        const float* rrow = (resid != nullptr && my_row < M) ? resid + (long long)my_row * N + n0 : nullptr;
        const float rsc = (resid != nullptr && my_row < M) ? rscale[my_row / rows_per_sample] : 0.f;
        if (cg_first >= BN || (dbg & 8)) {   // a single column group: this warp only hands the buffer back
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          continue;
        }
        for (int cg = cg_first; cg < BN; cg += 128) {
          uint32_t r[2][32];
          tmem_ld_32x32(t_addr + (uint32_t)cg, r[0]);
          tmem_ld_32x32(t_addr + (uint32_t)(cg + 32), r[1]);
          tmem_wait_ld();
          if (cg + 128 >= BN) {           // this warp's share of the accumulator is read: hand the TMEM buffer back early
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
          if (lane == 0) bulk_wait_group_read<0>();      // the previous store has read the staging tile
          __syncwarp();
          const float* bp = sBias + n0 + cg;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 b0 = *reinterpret_cast<const float4*>(bp + c * 32 + j);
              const float4 b1 = *reinterpret_cast<const float4*>(bp + c * 32 + j + 4);
              float v0 = __uint_as_float(r[c][j]) + b0.x, v1 = __uint_as_float(r[c][j + 1]) + b0.y;
              float v2 = __uint_as_float(r[c][j + 2]) + b0.z, v3 = __uint_as_float(r[c][j + 3]) + b0.w;
              float v4 = __uint_as_float(r[c][j + 4]) + b1.x, v5 = __uint_as_float(r[c][j + 5]) + b1.y;
              float v6 = __uint_as_float(r[c][j + 6]) + b1.z, v7 = __uint_as_float(r[c][j + 7]) + b1.w;
              if (sizeof(TOut) == 4 && resid != nullptr) {
                const float4 x0 = rrow != nullptr ? *reinterpret_cast<const float4*>(rrow + cg + c * 32 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 x1 = rrow != nullptr ? *reinterpret_cast<const float4*>(rrow + cg + c * 32 + j + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                auto rb = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
                v0 = fmaf(rb(v0), rsc, x0.x); v1 = fmaf(rb(v1), rsc, x0.y); v2 = fmaf(rb(v2), rsc, x0.z); v3 = fmaf(rb(v3), rsc, x0.w);
                v4 = fmaf(rb(v4), rsc, x1.x); v5 = fmaf(rb(v5), rsc, x1.y); v6 = fmaf(rb(v6), rsc, x1.z); v7 = fmaf(rb(v7), rsc, x1.w);
              }
              if (sizeof(TOut) == 2) {     // 8 columns = one 16-byte chunk of the 64-column box
                const uint32_t chunk = (uint32_t)(c * 4 + (j >> 3));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((chunk ^ sxor) << 4)),
                             "r"(pack_bf16x2(v0, v1)), "r"(pack_bf16x2(v2, v3)), "r"(pack_bf16x2(v4, v5)),
                             "r"(pack_bf16x2(v6, v7))
                             : "memory");
              } else {                     // 8 columns = two 16-byte chunks of box c (32 fp32 columns)
                const uint32_t chunk = (uint32_t)(j >> 2);
                const uint32_t bbase = srow + (uint32_t)c * 4096u;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bbase + ((chunk ^ sxor) << 4)), "f"(v0),
                             "f"(v1), "f"(v2), "f"(v3)
                             : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bbase + (((chunk + 1) ^ sxor) << 4)),
                             "f"(v4), "f"(v5), "f"(v6), "f"(v7)
                             : "memory");
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(dbg & 1) && m0 + quad * 32 < M) {
#pragma unroll
            const bool out2 = n0 >= N1;
            const int ncol = (out2 ? n0 - N1 : n0) + cg;
            for (int bx = 0; bx < NBOX; ++bx)
              tma_store_2d(out2 ? &tmY2 : &tmY, tstage + bx * 4096, ncol + bx * BOX_COLS, m0 + quad * 32);
            bulk_commit_group();
          }
        }
      }
      // the staging tile must have been read before the CTA's shared memory goes away; the global writes themselves
      // complete asynchronously (they are ordered before the end of the grid, like any other store)
      if (lane == 0) bulk_wait_group_read<0>();
    } else
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
      const int buf = li & 1;
      const int m0 = (t % m_tiles) * TC_BM, n0 = (t / m_tiles) * BN;
      const float* bvec = sBias + n0;
      mbar_wait(&acc_full[buf], (uint32_t)((li >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      const int rows_here = min(32, M - (m0 + quad * 32));
      const int nch = gcols / 32;
      const int cg_first = chalf * gcols;
      if (cg_first >= BN || (dbg & 8)) {   // a single column group: this warp only hands the buffer back
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      for (int cg = cg_first; cg < BN; cg += 2 * gcols) {
        // the whole column group (<= 96 columns) is fetched with back-to-back tcgen05.ld and one
        // wait: one TMEM round trip per group instead of one per 32 columns
        uint32_t r[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < nch) tmem_ld_32x32(t_addr + (uint32_t)(cg + c * 32), r[c]);
        tmem_wait_ld();
        if (cg + 2 * gcols >= BN) {      // this warp's share of the accumulator is read: hand the TMEM buffer back early
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < nch) {
            const float* bp = bvec + cg + c * 32;
            TOut* dst = reinterpret_cast<TOut*>(stage + lane * stage_pitch) + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 b0 = *reinterpret_cast<const float4*>(bp + j);
              const float4 b1 = *reinterpret_cast<const float4*>(bp + j + 4);
              const float4 v0 = make_float4(__uint_as_float(r[c][j]) + b0.x, __uint_as_float(r[c][j + 1]) + b0.y,
                                            __uint_as_float(r[c][j + 2]) + b0.z, __uint_as_float(r[c][j + 3]) + b0.w);
              const float4 v1 = make_float4(__uint_as_float(r[c][j + 4]) + b1.x, __uint_as_float(r[c][j + 5]) + b1.y,
                                            __uint_as_float(r[c][j + 6]) + b1.z, __uint_as_float(r[c][j + 7]) + b1.w);
              store8(dst + j, v0, v1);    // 16-byte shared stores: conflict-free with the +16 B row pitch
            }
          }
        }
        __syncwarp();
        // 512 bytes per store instruction: 512 / seg_bytes rows at a time
        const int lanes_per_row = seg_bytes / 16;
        const int rows_per_it = 32 / lanes_per_row;
        const int rsub = lane / lanes_per_row, off = (lane % lanes_per_row) * 16;
        for (int rr = rsub; rr < ((dbg & 1) ? 0 : rows_here); rr += rows_per_it) {
          uint8_t* grow = reinterpret_cast<uint8_t*>(Y + (long long)(m0 + quad * 32 + rr) * N + n0 + cg);
          *reinterpret_cast<uint4*>(grow + off) = *reinterpret_cast<const uint4*>(stage + rr * stage_pitch + off);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// CTA-pair version (cta_group::2): two SMs of a TPC own one 256 x BN output tile.  Each CTA streams its own 128 rows of
// A and only HALF of the B panel (BN / 2 weight rows); one M = 256 tcgen05.mma issued by the even CTA reads both shared
// memories and accumulates each CTA's 128 rows in its own tensor memory.  The single-CTA kernel is bound by the
// L2 -> shared-memory fill (every 128-row tile re-reads the whole weight panel: 98 MB of fills for the 42 MB stage-2
// fc1 GEMM); pairing halves the weight-panel traffic per output element, which is what the kernel waits for.
//   barriers: full[s] lives in the even CTA (2 producer arrivals + the bytes of both CTAs' loads); empty[s],
//   acc_full[b] are multicast by tcgen05.commit to both CTAs; acc_empty[b] (even CTA) collects the epilogue warps of both.
template <bool TF32, typename TOut, bool BMN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TCP_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const float* __restrict__ bias, TOut* __restrict__ Y, int M, int N, int k_chunks,
                    int k_chunks1, int BN, int stages, int tmem_cols, int m_pairs, int total_tiles,
                    int stage_pitch) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int BH = BN / 2;                              // weight rows (output columns) staged by this CTA
  const int b_stage_bytes = BH * CHUNK_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sBias = reinterpret_cast<float*>(smem + 1024);
  uint8_t* sA = smem + 3072;
  uint8_t* sB = sA + stages * A_STAGE_BYTES;
  uint8_t* sStage = sB + stages * b_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  constexpr int CHUNK_ELEMS = TF32 ? 32 : 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (k_chunks1 < k_chunks) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 2);                          // the producers of both CTAs
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 2 * TCP_EPI_WARPS);     // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = cid; t < total_tiles; t += n_clusters) {
        const int m0 = (t % m_pairs) * (2 * TC_BM) + (int)rank * TC_BM, n0 = (t / m_pairs) * BN + (int)rank * BH;
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);               // the pair's MMAs have read this slot (commit multicast)
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)(A_STAGE_BYTES + b_stage_bytes));
          else mbar_arrive_leader(&full[s]);
          const bool second = kc >= k_chunks1;
          const int kcol = (second ? kc - k_chunks1 : kc) * CHUNK_ELEMS;
          tma_load_2d_pair(sA + s * A_STAGE_BYTES, second ? &tmA2 : &tmA, &full[s], kcol, m0);
          if (BMN) {
            for (int i = 0; i < BH / 64; ++i)
              tma_load_2d_pair(sB + s * b_stage_bytes + i * 8192, second ? &tmB2 : &tmB, &full[s], n0 + 64 * i, kcol);
          } else {
            tma_load_2d_pair(sB + s * b_stage_bytes, second ? &tmB2 : &tmB, &full[s], kcol, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_instr_desc(TF32 ? FMT_TF32 : FMT_BF16, 2 * TC_BM, (uint32_t)BN, 0, BMN ? 1u : 0u);
      int it = 0, li = 0;
      for (int t = cid; t < total_tiles; t += n_clusters, ++li) {
        const int buf = li & 1;
        mbar_wait(&acc_empty[buf], (uint32_t)((li >> 1) & 1) ^ 1u);   // both CTAs' epilogues have drained this buffer
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + s * b_stage_bytes);
#pragma unroll
          for (int k4 = 0; k4 < CHUNK_BYTES / 32; ++k4) {
            const uint64_t ad = make_smem_desc(a_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
            const uint64_t bd = BMN ? make_smem_desc(b_addr + k4 * 2048, 8192, 1024, LAYOUT_SW128)
                                    : make_smem_desc(b_addr + k4 * 32, 16, 1024, LAYOUT_SW128);
            if (TF32) mma_tf32_ss_pair(d_tmem, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
            else mma_bf16_ss_pair(d_tmem, ad, bd, idesc, (uint32_t)((kc | k4) != 0));
          }
          tc_commit_pair(&empty[s], 3);
        }
        tc_commit_pair(&acc_full[buf], 3);
      }
    }
  } else {
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;   // which of the quadrant's two warps: takes column groups chalf, chalf + 2, ...
    const int epi_tid = threadIdx.x - 64;
    const int gcols = persistent_gcols(BN);
    const int seg_bytes = gcols * (int)sizeof(TOut);
    uint8_t* stage = sStage + (warp - 2) * 32 * stage_pitch;
    int li = 0;
    for (int t = cid; t < total_tiles; t += n_clusters, ++li) {
      const int buf = li & 1;
      const int m0 = (t % m_pairs) * (2 * TC_BM) + (int)rank * TC_BM, n0 = (t / m_pairs) * BN;
      float* bvec = sBias + buf * 256;
      for (int i = epi_tid; i < BN; i += 32 * TCP_EPI_WARPS) bvec[i] = bias != nullptr ? bias[n0 + i] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");     // the eight epilogue warps
      mbar_wait(&acc_full[buf], (uint32_t)((li >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      const int rows_here = min(32, M - (m0 + quad * 32));
      const int nch = gcols / 32;
      const int cg_first = chalf * gcols;
      if (cg_first >= BN) {              // a single column group: this warp only hands the buffer back
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
        continue;
      }
      for (int cg = cg_first; cg < BN; cg += 2 * gcols) {
        uint32_t r[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < nch) tmem_ld_32x32(t_addr + (uint32_t)(cg + c * 32), r[c]);
        tmem_wait_ld();
        if (cg + 2 * gcols >= BN) {      // this warp's share of the accumulator is read: hand the TMEM buffer back early
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < nch) {
            const float* bp = bvec + cg + c * 32;
            TOut* dst = reinterpret_cast<TOut*>(stage + lane * stage_pitch) + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 b0 = *reinterpret_cast<const float4*>(bp + j);
              const float4 b1 = *reinterpret_cast<const float4*>(bp + j + 4);
              const float4 v0 = make_float4(__uint_as_float(r[c][j]) + b0.x, __uint_as_float(r[c][j + 1]) + b0.y,
                                            __uint_as_float(r[c][j + 2]) + b0.z, __uint_as_float(r[c][j + 3]) + b0.w);
              const float4 v1 = make_float4(__uint_as_float(r[c][j + 4]) + b1.x, __uint_as_float(r[c][j + 5]) + b1.y,
                                            __uint_as_float(r[c][j + 6]) + b1.z, __uint_as_float(r[c][j + 7]) + b1.w);
              store8(dst + j, v0, v1);
            }
          }
        }
        __syncwarp();
        const int lanes_per_row = seg_bytes / 16;
        const int rows_per_it = 32 / lanes_per_row;
        const int rsub = lane / lanes_per_row, off = (lane % lanes_per_row) * 16;
        for (int rr = rsub; rr < rows_here; rr += rows_per_it) {
          uint8_t* grow = reinterpret_cast<uint8_t*>(Y + (long long)(m0 + quad * 32 + rr) * N + n0 + cg);
          *reinterpret_cast<uint4*>(grow + off) = *reinterpret_cast<const uint4*>(stage + rr * stage_pitch + off);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();                    // neither CTA leaves (or frees tensor memory) while the pair's MMAs may touch it
  if (warp == 1) tmem_dealloc2(tmem_base, (uint32_t)tmem_cols);
}

// fp32 (C x C) -> transposed bf16 for up to 4 matrices in one launch: out[z][k][n] = w_z[n][k].
// The data-gradient GEMMs dX = dY W are then plain K-major products with "weight" W^T.
__global__ void cast_transpose_bf16_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                           const float* __restrict__ w2, const float* __restrict__ w3,
                                           bf16* __restrict__ out, int N, int K) {
  pdl_enter();
  __shared__ float tile[32][33];          // w_z is (N, K) row-major, out[z] is (K, N)
  const float* src = blockIdx.z == 0 ? w0 : (blockIdx.z == 1 ? w1 : (blockIdx.z == 2 ? w2 : w3));
  if (src == nullptr) return;
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int n = n0 + r, k = k0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && k < K) ? src[(long long)n * K + k] : 0.f;
  }
  __syncthreads();
  bf16* dst = out + (long long)blockIdx.z * N * K;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int k = k0 + r, n = n0 + threadIdx.x;
    if (k < K && n < N) dst[(long long)k * N + n] = __float2bfloat16_rn(tile[threadIdx.x][r]);
  }
}

// fp32 -> bf16 for up to 3 equally sized matrices in one launch (weights of a block)
__global__ void cast_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                 const float* __restrict__ c, bf16* __restrict__ out, long long n) {
  pdl_enter();
  const float* src = blockIdx.y == 0 ? a : (blockIdx.y == 1 ? b : c);
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (src == nullptr || i >= n) return;
  store4(out + (long long)blockIdx.y * n + i, load4(src + i));
}

// fp32 -> bf16 for a whole table of tensors in one launch (the bf16 operand copies of every 1x1-conv weight of a
// model, once per step): blockIdx.y = table entry, blockIdx.x strides over its elements, 4 per thread.
__global__ void cast_bf16_multi_kernel(const dat_cast_item* __restrict__ items) {
  pdl_enter();
  const dat_cast_item it = items[blockIdx.y];
  const float* __restrict__ src = it.src;
  bf16* __restrict__ dst = reinterpret_cast<bf16*>(it.dst);
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < it.n; i += (long long)gridDim.x * blockDim.x * 4)
    store4(dst + i, load4(src + i));
}

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e != nullptr && std::atoi(e) > 0 ? std::atoi(e) : dflt;
}

int pick_bn(int N) {
  const int cap = env_int("DAT_B200_GEMM_BN", 256);     // tuning knob: widest output tile
  if (cap < 256 && N % cap == 0) return cap;
  // widest tile that divides N; tiles wider than one epilogue column group (128) must be a multiple of it
  for (int bn : {256, 128, 96, 64, 32})
    if (N % bn == 0) return bn;
  return 0;
}

// MN-major weight operand: 64-column TMA boxes, so the tile width is a multiple of 64
int pick_bn_mn(int N) {
  for (int bn : {256, 128, 64})
    if (N % bn == 0) return bn;
  return 0;
}

}  // namespace

int debug_gemm_timing(unsigned long long* out8) {
  DAT_CUDA_OK(cudaMemcpyFromSymbol(out8, g_gemm_timing, sizeof(unsigned long long) * 8));
  return DAT_OK;
}

bool pointwise_fwd_tc_supported(int x_dt, long long M, int N, int K) {
  if (M <= 0 || pick_bn(N) == 0) return false;
  return x_dt == DAT_F32 ? (K % 32 == 0) : (K % 8 == 0);
}

int cast_weights_bf16(const float* a, const float* b, const float* c, void* out, long long n,
                      cudaStream_t st) {
  DAT_REQUIRE(n % 4 == 0, "cast_weights: n must be a multiple of 4");
  dim3 grid(ceil_div(n / 4, 256), 3);
  launch_k(cast_bf16_kernel, grid, 256, 0, st, a, b, c, (bf16*)out, n);
  DAT_LAUNCH_OK("cast_bf16_kernel");
  return DAT_OK;
}

int cast_transpose_weights_bf16(const float* w0, const float* w1, const float* w2, const float* w3,
                                void* out, int C, cudaStream_t st) {
  dim3 grid(ceil_div(C, 32), ceil_div(C, 32), 4), block(32, 8);
  launch_k(cast_transpose_bf16_kernel, grid, block, 0, st, w0, w1, w2, w3, (bf16*)out, C, C);
  DAT_LAUNCH_OK("cast_transpose_bf16_kernel");
  return DAT_OK;
}

// one (N, K) fp32 matrix -> (K, N) bf16
int cast_transpose_bf16(const float* w, void* out, int N, int K, cudaStream_t st) {
  dim3 grid(ceil_div(K, 32), ceil_div(N, 32), 1), block(32, 8);
  launch_k(cast_transpose_bf16_kernel, grid, block, 0, st, w, nullptr, nullptr, nullptr, (bf16*)out, N, K);
  DAT_LAUNCH_OK("cast_transpose_bf16_kernel");
  return DAT_OK;
}

bool pointwise_fwd_tc_two_outputs_supported(int N) {
  return pick_bn(N) != 0 && pick_bn(N) % 64 == 0 && std::getenv("DAT_B200_GEMM_LEGACY") == nullptr &&
         std::getenv("DAT_B200_KV_TWO_LAUNCHES") == nullptr;
}

int pointwise_fwd_tc(const void* X, int x_dt, const void* W, const float* b, void* Y, int y_dt,
                     long long M, int N, int K, cudaStream_t st) {
  return pointwise_fwd_tc_dual(X, W, nullptr, nullptr, x_dt, b, Y, y_dt, M, N, K, st);
}

int cast_bf16_multi(const dat_cast_item* items_dev, int n_items, cudaStream_t st) {
  if (n_items <= 0) return DAT_OK;
  dim3 grid(32, n_items);
  launch_k(cast_bf16_multi_kernel, grid, 256, 0, st, items_dev);
  DAT_LAUNCH_OK("cast_bf16_multi_kernel");
  return DAT_OK;
}

// data gradient dX[M, K] = dY[M, N] W[N, K] (+ dY2 W2) with the bf16 weight read in place as an MN-major B operand
bool pointwise_dgrad_tc_supported(long long M, int N, int K) {
  return M > 0 && pick_bn_mn(K) != 0 && N % 8 == 0;
}
int pointwise_dgrad_tc(const void* dY, const void* W, const void* dY2, const void* W2, void* dX, int dx_dt,
                       long long M, int N, int K, cudaStream_t st) {
  DAT_REQUIRE(pointwise_dgrad_tc_supported(M, N, K), "pointwise_dgrad_tc: unsupported shape M=%lld N=%d K=%d", M, N, K);
  return pointwise_fwd_tc_dual(dY, W, dY2, W2, DAT_BF16, nullptr, dX, dx_dt, M, K, N, st, true);
}

// Y = X W^T (+ X2 W2^T) + b.  W: fp32 when x_dt == DAT_F32 (tf32 MMA), bf16 when x_dt == DAT_BF16.
// w_mn: W (and W2) are (K, N) row-major bf16 matrices read as MN-major B operands (see the kernel's BMN flag).
// Y2 != NULL (two products of one X in one launch): Y = X W^T + b and Y2 = X W2^T + b2, both (M, N); needs a tile width
// that is a multiple of 64, no second K source and K-major weights.
int pointwise_fwd_tc_dual(const void* X, const void* W, const void* X2, const void* W2, int x_dt,
                          const float* b, void* Y, int y_dt, long long M, int N, int K,
                          cudaStream_t st, bool w_mn, const float* b2, void* Y2, const float* resid, const float* rscale,
                          long long rows_per_sample) {
  DAT_REQUIRE(pointwise_fwd_tc_supported(x_dt, M, N, K), "pointwise_fwd_tc: unsupported shape M=%lld N=%d K=%d", M, N, K);
  const bool tf32 = x_dt == DAT_F32;
  const bool two_out = Y2 != nullptr;
  DAT_REQUIRE(resid == nullptr || (y_dt == DAT_F32 && !two_out && rscale != nullptr && rows_per_sample > 0 &&
                                   pick_bn(N) % 64 == 0 && !w_mn && std::getenv("DAT_B200_GEMM_LEGACY") == nullptr),
              "pointwise_fwd_tc: the fused residual needs an fp32 output and a tile width that is a multiple of 64");
  DAT_REQUIRE(!two_out || (X2 == nullptr && W2 != nullptr && !w_mn && pick_bn(N) % 64 == 0 &&
                           std::getenv("DAT_B200_GEMM_LEGACY") == nullptr),
              "pointwise_fwd_tc: two outputs need K-major weights, one K source and a tile width that is a multiple of 64");
  DAT_REQUIRE(!w_mn || (!tf32 && pick_bn_mn(N) != 0 && std::getenv("DAT_B200_GEMM_LEGACY") == nullptr),
              "pointwise_fwd_tc: the MN-major weight operand needs bf16 and a tile width that is a multiple of 64");
  const int eb = tf32 ? 4 : 2;
  const int chunk_elems = CHUNK_BYTES / eb;
  const int BN = w_mn ? pick_bn_mn(N) : pick_bn(N);
  const int k_chunks1 = (K + chunk_elems - 1) / chunk_elems;
  const int k_chunks = X2 != nullptr ? 2 * k_chunks1 : k_chunks1;
  CUtensorMap tmA, tmB, tmA2, tmB2;
  DAT_FWD(tc::make_tmap_2d(&tmA, X, eb, tf32, (uint64_t)M, (uint64_t)K, (uint64_t)K * eb, TC_BM, chunk_elems, 128));
  auto map_w = [&](CUtensorMap* tm, const void* w) -> int {
    if (w_mn) return tc::make_tmap_2d(tm, w, 2, false, (uint64_t)K, (uint64_t)N, (uint64_t)N * 2, 64, 64, 128);
    return tc::make_tmap_2d(tm, w, eb, tf32, (uint64_t)N, (uint64_t)K, (uint64_t)K * eb, BN, chunk_elems, 128);
  };
  DAT_FWD(map_w(&tmB, W));
  tmA2 = tmA;
  tmB2 = tmB;
  if (X2 != nullptr) {
    DAT_FWD(tc::make_tmap_2d(&tmA2, X2, eb, tf32, (uint64_t)M, (uint64_t)K, (uint64_t)K * eb, TC_BM, chunk_elems, 128));
    DAT_FWD(map_w(&tmB2, W2));
  }
  if (two_out) DAT_FWD(map_w(&tmB2, W2));
  const int stage_bytes = A_STAGE_BYTES + BN * CHUNK_BYTES;
  // CTA pairs (cta_group::2, 256 x BN tiles): tiles of 128 or 256 columns, at least one full pair of row tiles.
  // Opt-in (DAT_B200_GEMM_PAIR=1): measured on B200 it halves the weight-panel fills but is no faster than the
  // single-CTA kernel at any DAT-T++ shape (20.8 vs 20.9 us stage-2 fc1, 16.7 vs 16.8 us fc2, 16.6 vs 14.8 us stage-3
  // fc1: profiles/r02_gemm_pair.md) - the kernel is bound by ring depth x fill latency and by its epilogue, not by
  // L2 -> shared-memory bandwidth.
  static const int pair_off = [] { const char* e = std::getenv("DAT_B200_GEMM_PAIR"); return e && e[0] == '1' ? 0 : 1; }();
  if (!pair_off && !two_out && resid == nullptr && std::getenv("DAT_B200_GEMM_LEGACY") == nullptr && (BN == 256 || BN == 128) && M >= 2 * TC_BM) {
    const int gcols = persistent_gcols(BN);
    const int stage_pitch = gcols * (int)dtype_size(y_dt) + 16;
    const int staging = TCP_EPI_WARPS * 32 * stage_pitch;
    const int pstage_bytes = A_STAGE_BYTES + (BN / 2) * CHUNK_BYTES;
    int stages = (224 * 1024 - 1024 - 3072 - staging) / pstage_bytes;
    if (stages > 8) stages = 8;
    DAT_REQUIRE(stages >= 2, "pointwise_fwd_tc: tile does not fit shared memory");
    const size_t smem = 1024 + 3072 + (size_t)stages * pstage_bytes + staging;
    int tmem_cols = 32;
    while (tmem_cols < 2 * BN) tmem_cols <<= 1;
    const int m_pairs = (int)ceil_div(M, (long long)(2 * TC_BM)), total = m_pairs * (N / BN);
    const int clusters = total < 74 ? total : 74;
    CUtensorMap tmBp = tmB, tmB2p = tmB2;
    if (!w_mn) {     // K-major weight: this CTA's half of the panel is one box of BN / 2 rows
      DAT_FWD(tc::make_tmap_2d(&tmBp, W, eb, tf32, (uint64_t)N, (uint64_t)K, (uint64_t)K * eb, BN / 2, chunk_elems, 128));
      tmB2p = tmBp;
      if (X2 != nullptr)
        DAT_FWD(tc::make_tmap_2d(&tmB2p, W2, eb, tf32, (uint64_t)N, (uint64_t)K, (uint64_t)K * eb, BN / 2, chunk_elems, 128));
    }
#define LAUNCH_PAIR(TF, TO, MN)                                                                   \
  do {                                                                                            \
    auto kern = gemm_tc_pair_kernel<TF, TO, MN>;                                                  \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(kern, 2 * clusters, TCP_THREADS, smem, st, tmA, tmBp, tmA2, tmB2p, b, (TO*)Y, (int)M, N, k_chunks, \
                                                  k_chunks1, BN, stages, tmem_cols, m_pairs, total, stage_pitch); \
  } while (0)
    if (w_mn && y_dt == DAT_F32) LAUNCH_PAIR(false, float, true);
    else if (w_mn) LAUNCH_PAIR(false, bf16, true);
    else if (tf32 && y_dt == DAT_F32) LAUNCH_PAIR(true, float, false);
    else if (tf32) LAUNCH_PAIR(true, bf16, false);
    else if (y_dt == DAT_F32) LAUNCH_PAIR(false, float, false);
    else LAUNCH_PAIR(false, bf16, false);
#undef LAUNCH_PAIR
    DAT_LAUNCH_OK("gemm_tc_pair_kernel");
    return DAT_OK;
  }
  if (std::getenv("DAT_B200_GEMM_LEGACY") == nullptr) {
    // persistent kernel: ring + separate epilogue staging in up to 224 KB, one CTA per SM
    const int gcols = persistent_gcols(BN);
    const int stage_pitch = gcols * (int)dtype_size(y_dt) + 16;
    // tiles that are a multiple of 64 columns wide leave through TMA stores: one 128B-swizzled [32 rows x 128 bytes]
    // box (bf16) or two (fp32) per epilogue warp; other widths use the padded, row-coalesced st.global staging
    const bool tma_out = BN % 64 == 0;
    const int staging = tma_out ? TCP_EPI_WARPS * 4096 * (int)(dtype_size(y_dt) / 2) : TCP_EPI_WARPS * 32 * stage_pitch;
    const int N1 = N;                                  // columns of the first output
    const int Ntot = two_out ? 2 * N : N;              // the kernel's N: both outputs side by side
    const int bias_bytes = (int)align_up((size_t)Ntot * 4, 1024);
    int stages = (226 * 1024 - 1024 - 1024 - bias_bytes - staging) / stage_bytes;
    if (stages > 8) stages = 8;
    DAT_REQUIRE(stages >= 2, "pointwise_fwd_tc: tile does not fit shared memory");
    const size_t smem = 1024 + 1024 + (size_t)bias_bytes + (size_t)stages * stage_bytes + staging;
    int tmem_cols = 32;
    while (tmem_cols < 2 * BN) tmem_cols <<= 1;
    const int m_tiles = (int)ceil_div(M, (long long)TC_BM), total = m_tiles * (Ntot / BN);
    const int grid = total < 148 ? total : 148;
    CUtensorMap tmY = tmA, tmY2 = tmA;
    if (tma_out) {
      const int eo = (int)dtype_size(y_dt);
      DAT_FWD(tc::make_tmap_2d(&tmY, Y, eo, y_dt == DAT_F32, (uint64_t)M, (uint64_t)N, (uint64_t)N * eo, 32, 128 / eo, 128));
      tmY2 = tmY;
      if (two_out)
        DAT_FWD(tc::make_tmap_2d(&tmY2, Y2, eo, y_dt == DAT_F32, (uint64_t)M, (uint64_t)N, (uint64_t)N * eo, 32, 128 / eo, 128));
    }
    static const int gemm_dbg = [] { const char* e = std::getenv("DAT_B200_GEMM_DBG"); return e ? std::atoi(e) : 0; }();
#define LAUNCH_P(TF, TO, MN)                                                                      \
  do {                                                                                            \
    auto kern = gemm_tc_persistent_kernel<TF, TO, MN>;                                            \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(kern, grid, TCP_THREADS, smem, st, tmA, tmB, tmA2, tmB2, tmY, b, (TO*)Y, (int)M, Ntot, k_chunks,  \
             k_chunks1, BN, stages, tmem_cols, m_tiles, total, stage_pitch, bias_bytes, gemm_dbg, tmY2, b2, N1,    \
             resid, rscale, (int)rows_per_sample);                                                               \
  } while (0)
    if (w_mn && y_dt == DAT_F32) LAUNCH_P(false, float, true);
    else if (w_mn) LAUNCH_P(false, bf16, true);
    else if (tf32 && y_dt == DAT_F32) LAUNCH_P(true, float, false);
    else if (tf32) LAUNCH_P(true, bf16, false);
    else if (y_dt == DAT_F32) LAUNCH_P(false, float, false);
    else LAUNCH_P(false, bf16, false);
#undef LAUNCH_P
    DAT_LAUNCH_OK("gemm_tc_persistent_kernel");
    return DAT_OK;
  }
  const int budget = env_int("DAT_B200_GEMM_SMEM_KB", SMEM_BUDGET / 1024) * 1024;   // tuning knob: CTAs per SM
  int stages = (budget - 3072) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages > k_chunks) stages = k_chunks;
  if (stages < 1) stages = 1;
  const size_t out_stage = (size_t)4 * 32 * ((BN < EPI_COLS ? BN : EPI_COLS) * dtype_size(y_dt) + 16);
  size_t buf = (size_t)stages * stage_bytes;
  if (out_stage > buf) buf = out_stage;
  size_t smem = 1024 /*alignment slack*/ + 2048 /*barriers + bias*/ + buf;
  int tmem_cols = 32;
  while (tmem_cols < BN) tmem_cols <<= 1;
  dim3 grid(ceil_div(M, TC_BM), N / BN);
#define LAUNCH(TF, TO)                                                                          \
  do {                                                                                          \
    auto kern = gemm_tc_kernel<TF, TO>;                                                         \
    DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_k(kern, grid, TC_THREADS, smem, st, tmA, tmB, tmA2, tmB2, b, (TO*)Y, (int)M, N, k_chunks,    \
                                         k_chunks1, BN, stages, tmem_cols);                       \
  } while (0)
  if (tf32 && y_dt == DAT_F32) LAUNCH(true, float);
  else if (tf32) LAUNCH(true, bf16);
  else if (y_dt == DAT_F32) LAUNCH(false, float);
  else LAUNCH(false, bf16);
#undef LAUNCH
  DAT_LAUNCH_OK("gemm_tc_kernel");
  return DAT_OK;
}

}  // namespace dat
