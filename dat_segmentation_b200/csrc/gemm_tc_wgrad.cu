// Weight gradients of the 1x1 convolutions on the tcgen05 tensor cores:
//     dW[N, K] = sum_m dY[m, n] * X[m, k]          (autograd of dat_blocks.py:143,177-178,225)
// The contraction runs over the pixels m, i.e. over the ROWS of both channel-last matrices,
// so both operands are MN-major: TMA boxes of 64 rows x 64 columns (128B swizzle) are used
// directly as the canonical MN-major core-matrix layout (8-row groups 1 KB apart, 64-column
// blocks 8 KB apart) - no transposed copy of either activation is ever made.
// One CTA = one 128 (out-channel) x BNK (in-channel, <= 256) tile of dW over one slice of the
// pixels; fp32 partial tiles per slice are reduced in a fixed order afterwards (deterministic).
// The bias gradient db[n] = sum_m dY[m, n] rides along: the CTAs of the first in-channel tile issue
// one extra N = 16 MMA per step against a shared-memory tile of ones (accumulated in 16 spare
// TMEM columns), so dY is not read a second time by a column-sum kernel.
// Warp roles as in gemm_tc.cu: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2-5 = epilogue.
#include "kernels.h"
#include "tc_common.cuh"

namespace dat {

namespace {

using namespace tc;

constexpr int WG_THREADS = 192;
constexpr int WG_BM = 128;           // out-channel rows of dW per CTA (UMMA M)
constexpr int WG_CHUNK = 64;         // pixels per pipeline stage
constexpr int BOX_BYTES = 64 * 128;  // one 64-row x 64-column bf16 box
constexpr int ONES_BYTES = 2048;     // bf16 ones read by the N = 16 bias-gradient MMA

__global__ void __launch_bounds__(WG_THREADS, 2)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     float* __restrict__ partial, float* __restrict__ db_partial, int N, int K, int BNK,
                     int rows_per_split, long long M, int stages, int tmem_cols) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int a_bytes = 2 * BOX_BYTES, b_bytes = (BNK / 64) * BOX_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  uint8_t* sOnes = smem + 1024;
  uint8_t* sA = sOnes + ONES_BYTES;
  uint8_t* sB = sA + stages * a_bytes;
  const bool with_db = db_partial != nullptr && blockIdx.y == 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * WG_BM, k0 = blockIdx.y * BNK, split = blockIdx.z;
  const long long m_begin = (long long)split * rows_per_split;
  long long m_end = m_begin + rows_per_split;
  if (m_end > M) m_end = M;
  const int chunks = m_end > m_begin ? (int)((m_end - m_begin + WG_CHUNK - 1) / WG_CHUNK) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (with_db) {
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += WG_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3f803f80u;
    fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor-core (async) proxy
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (uint32_t)(kc / stages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        // a 64-channel tail tile loads one dY box only: rows 64-127 of the accumulator then hold
        // products of stale shared memory, stay in their own rows and are never stored
        const int a_boxes = (N - n0 >= WG_BM) ? 2 : 1;
        mbar_arrive_expect_tx(&full[s], (uint32_t)(a_boxes * BOX_BYTES + b_bytes));
        // rows past m_end belong to the next slice: they are loaded but must not count, so the
        // last chunk of a slice is clipped by construction (rows_per_split is a multiple of 64)
        const int mrow = (int)(m_begin + (long long)kc * WG_CHUNK);
        for (int i = 0; i < a_boxes; ++i)
          tma_load_2d(sA + s * a_bytes + i * BOX_BYTES, &tmDY, &full[s], n0 + 64 * i, mrow);
        for (int i = 0; i < BNK / 64; ++i)
          tma_load_2d(sB + s * b_bytes + i * BOX_BYTES, &tmX, &full[s], k0 + 64 * i, mrow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(FMT_BF16, WG_BM, (uint32_t)BNK, 1, 1);   // A, B MN-major
      const uint32_t idesc1 = make_instr_desc(FMT_BF16, WG_BM, 16u, 1, 1);
      const uint64_t onesd = make_smem_desc(smem_u32(sOnes), BOX_BYTES, 1024, LAYOUT_SW128);
      for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (uint32_t)(kc / stages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(sA + s * a_bytes), b_addr = smem_u32(sB + s * b_bytes);
#pragma unroll
        for (int j = 0; j < WG_CHUNK / 16; ++j) {      // 16 pixels per MMA
          const uint64_t ad = make_smem_desc(a_addr + j * 2048, BOX_BYTES, 1024, LAYOUT_SW128);
          const uint64_t bd = make_smem_desc(b_addr + j * 2048, BOX_BYTES, 1024, LAYOUT_SW128);
          mma_bf16_ss(tmem_base, ad, bd, idesc, (uint32_t)((kc | j) != 0));
          if (with_db) mma_bf16_ss(tmem_base + (uint32_t)BNK, ad, onesd, idesc1, (uint32_t)((kc | j) != 0));
        }
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  } else {
    const int quad = warp & 3;
    float* out = partial + ((long long)split * N + n0 + quad * 32) * K + k0;
    // column groups of 128 keep the staging area at 66 KB: two CTAs per SM overlap their phases
    const int gcols = BNK < 128 ? BNK : 128;
    const int seg_bytes = gcols * 4, pitch = seg_bytes + 16;
    uint8_t* stage = sA + (warp - 2) * 32 * pitch;
    if (chunks > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after_sync();
    }
    const int rows_here = min(32, N - (n0 + quad * 32));
    if (with_db) {                       // column BNK of this lane's row = sum over the slice's pixels
      uint32_t r[32];
      if (chunks > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)BNK, r);
        tmem_wait_ld();
      } else {
        r[0] = 0u;
      }
      if (lane < rows_here) db_partial[(long long)split * N + n0 + quad * 32 + lane] = __uint_as_float(r[0]);
    }
    for (int cg = 0; cg < BNK; cg += gcols) {
      for (int c = 0; c < gcols / 32; ++c) {
        uint32_t r[32];
        if (chunks > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cg + c * 32), r);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        float* dst = reinterpret_cast<float*>(stage + lane * pitch) + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      __syncwarp();
      for (int rr = 0; rr < rows_here; ++rr) {
        uint8_t* grow = reinterpret_cast<uint8_t*>(out + (long long)rr * K + cg);
        const uint8_t* srow = stage + rr * pitch;
        for (int off = lane * 16; off < seg_bytes; off += 512)
          *reinterpret_cast<uint4*>(grow + off) = *reinterpret_cast<const uint4*>(srow + off);
      }
      __syncwarp();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// in-channel tile: <= 128 so that accumulator + 16 bias-gradient columns fit 256 TMEM columns
// (two CTAs per SM share the 512)
int wg_bnk(int C) { return C % 128 == 0 ? 128 : 64; }

}  // namespace

bool pointwise_wgrad_tc_supported(long long M, int N, int K) {
  // N = 64 (and N = 32: the RGB stem convolution) run as partly empty 128-row tiles (the out-of-range part of the
  // dY box is zero-filled by TMA, the rows beyond N are never stored)
  return M >= 64 && (N % 64 == 0 || N == 32) && K % 64 == 0 && (K % wg_bnk(K)) == 0 && M < (1ll << 31);
}

int pointwise_wgrad_tc_splits(long long M, int N, int K) {
  const int tiles = ((N + WG_BM - 1) / WG_BM) * (K / wg_bnk(K));
  long long want = (2 * 148 + tiles - 1) / tiles;   // two CTAs per SM
  long long cap = (M + 127) / 128;            // at least 2 chunks per slice
  long long s = want < cap ? want : cap;
  return (int)(s < 1 ? 1 : s);
}

size_t pointwise_wgrad_tc_workspace(long long M, int N, int K) {
  const size_t splits = (size_t)pointwise_wgrad_tc_splits(M, N, K);
  return align_up(splits * N * K * 4, 256) + align_up(splits * N * 4, 256);
}

namespace {
// fixed-order sum over the splits of dW and db partials in one launch
__global__ void wgrad_reduce_kernel(const float* __restrict__ wpart, const float* __restrict__ bpart,
                                    int splits, long long nk, int n, float* __restrict__ dW,
                                    float* __restrict__ db) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nk) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += wpart[(long long)z * nk + i];
    dW[i] = s;
  } else if (i < nk + n && db != nullptr) {
    const int c = (int)(i - nk);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += bpart[(long long)z * n + c];
    db[c] = s;
  }
}
}  // namespace

// dW[N,K] (fp32, overwritten) = dY[M,N]^T X[M,K] and, when db != NULL, db[N] = column sums of dY;
// dY, X bf16 channel-last.
int pointwise_wgrad_tc(const void* dY, const void* X, float* dW, float* db, long long M, int N, int K,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  DAT_REQUIRE(pointwise_wgrad_tc_supported(M, N, K), "pointwise_wgrad_tc: unsupported shape");
  DAT_REQUIRE(ws_bytes >= pointwise_wgrad_tc_workspace(M, N, K), "pointwise_wgrad_tc: workspace too small");
  const int BNK = wg_bnk(K);
  const int splits = pointwise_wgrad_tc_splits(M, N, K);
  long long rps = (M + splits - 1) / splits;
  rps = (rps + WG_CHUNK - 1) / WG_CHUNK * WG_CHUNK;
  CUtensorMap tmDY, tmX;
  DAT_FWD(tc::make_tmap_2d(&tmDY, dY, 2, false, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, 64, 64, 128));
  DAT_FWD(tc::make_tmap_2d(&tmX, X, 2, false, (uint64_t)M, (uint64_t)K, (uint64_t)K * 2, 64, 64, 128));
  const int stage_bytes = 2 * BOX_BYTES + (BNK / 64) * BOX_BYTES;
  int stages = 100 * 1024 / stage_bytes;      // two CTAs per SM
  if (stages > 6) stages = 6;
  size_t buf = (size_t)stages * stage_bytes;
  const size_t out_stage = (size_t)4 * 32 * ((BNK < 128 ? BNK : 128) * 4 + 16);
  if (out_stage > buf) buf = out_stage;
  const size_t smem = 1024 + 1024 + ONES_BYTES + buf;
  int tmem_cols = 32;
  while (tmem_cols < BNK + 32) tmem_cols <<= 1;      // + the bias-gradient columns (read 32 wide)
  float* part = splits > 1 ? (float*)ws : dW;
  float* bpart = nullptr;
  if (db != nullptr)
    bpart = splits > 1 ? (float*)((char*)ws + align_up((size_t)splits * N * K * 4, 256)) : db;
  dim3 grid((N + WG_BM - 1) / WG_BM, K / BNK, splits);
  DAT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_k(gemm_tc_wgrad_kernel, grid, WG_THREADS, smem, st, tmDY, tmX, part, bpart, N, K, BNK, (int)rps, M, stages,
                                                       tmem_cols);
  DAT_LAUNCH_OK("gemm_tc_wgrad_kernel");
  if (splits > 1) {
    const long long nk = (long long)N * K;
    launch_k(wgrad_reduce_kernel, (unsigned)ceil_div(nk + N, 256ll), 256, 0, st, part, bpart, splits, nk, N, dW, db);
    DAT_LAUNCH_OK("wgrad_reduce_kernel");
  }
  return DAT_OK;
}

}  // namespace dat
