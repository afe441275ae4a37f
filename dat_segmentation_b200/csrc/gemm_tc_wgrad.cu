// Weight gradients of the 1x1 convolutions on the tcgen05 tensor cores:
//     dW[N, K] = sum_m dY[m, n] * X[m, k]          (autograd of dat_blocks.py:143,177-178,225)
// The contraction runs over the pixels m, i.e. over the ROWS of both channel-last matrices,
// so both operands are MN-major: TMA boxes of 64 rows x 64 columns (128B swizzle) are used
// directly as the canonical MN-major core-matrix layout (8-row groups 1 KB apart, 64-column
// blocks 8 KB apart) - no transposed copy of either activation is ever made.
// One CTA = one 128 (out-channel) x BNK (in-channel, <= 128) tile of dW over one slice of the
// pixels.  The CTAs of up to 8 consecutive slices form a thread-block cluster: every CTA parks its fp32
// accumulator tile in its own shared memory and, after a cluster barrier, CTA r sums rows [r * 128 / CS, ...) of all
// CS tiles through distributed shared memory in rank order (deterministic) and writes ONE partial tile per cluster -
// 8 x fewer partial bytes to write and to reduce afterwards, and no reduction launch at all when one cluster covers
// the pixels.  Remaining per-cluster partials are summed in a fixed order by wgrad_reduce_kernel.
// The bias gradient db[n] = sum_m dY[m, n] rides along: the CTAs of the first in-channel tile issue
// one extra N = 16 MMA per step against a shared-memory tile of ones (accumulated in 16 spare
// TMEM columns), so dY is not read a second time by a column-sum kernel.
// Warp roles as in gemm_tc.cu: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2-5 = epilogue.
#include <cstdlib>

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
                     int rows_per_split, long long M, int stages, int tmem_cols, int CS) {
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
    // accumulator tile -> this CTA's shared memory (the ring is idle once tmem_full has fired), row pitch + 16 bytes:
    // conflict-free 16-byte stores with one row per lane
    const int quad = warp & 3;
    if (chunks > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after_sync();
    }
    const int pitch = BNK * 4 + 16;
    uint8_t* myrow = sA + (quad * 32 + lane) * pitch;
    for (int c = 0; c < BNK / 32; ++c) {
      uint32_t r[32];
      if (chunks > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(myrow + (c * 32 + j) * 4) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
    if (db_partial != nullptr) {           // column BNK of this lane's row = sum of dY over the slice's pixels
      uint32_t r[32];
      r[0] = 0u;
      if (chunks > 0 && with_db) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)BNK, r);
        tmem_wait_ld();
      }
      reinterpret_cast<float*>(sA + 128 * pitch)[quad * 32 + lane] = __uint_as_float(r[0]);
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();                      // every tile of the cluster is parked (release / acquire at cluster scope)
  {
    const int crank = (int)cluster_ctarank();
    const int pitch = BNK * 4 + 16;
    const int rows_per = WG_BM / CS, r0 = crank * rows_per;
    const int part_idx = split / CS;       // one partial tile per cluster
    const int n4 = BNK / 4;
    const uint32_t my = smem_u32(sA);
    for (int i = threadIdx.x; i < rows_per * n4; i += WG_THREADS) {
      const int row = r0 + i / n4, c4 = i - (i / n4) * n4;
      if (n0 + row >= N) continue;         // 64-row tail tile
      const uint32_t local = my + (uint32_t)(row * pitch + c4 * 16);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int z = 0; z < CS; ++z) {       // rank order: the sum does not depend on scheduling
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(z));
        float4 v;
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(remote)
                     : "memory");
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(partial + ((long long)part_idx * N + n0 + row) * K + k0 + c4 * 4) = acc;
    }
    if (with_db && (int)threadIdx.x < rows_per && n0 + r0 + (int)threadIdx.x < N) {
      const int row = r0 + threadIdx.x;
      const uint32_t local = my + (uint32_t)(128 * pitch + row * 4);
      float acc = 0.f;
      for (int z = 0; z < CS; ++z) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(z));
        float v;
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
        acc += v;
      }
      db_partial[(long long)part_idx * N + n0 + row] = acc;
    }
  }
  cluster_sync_all();                      // nobody leaves while a peer still reads its tile
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

// CTAs of clusters of size c that are resident at once with two CTAs per SM.  A cluster lives inside one GPC (16 - 20
// of the 148 SMs): 4 clusters of 8 per GPC but 8 - 10 clusters of 4.  Measured on B200: 32 clusters of 8 run as one wave
// (stage-2 fc2 weight gradient 24.8 -> 22.7 us), 36 or 38 take two (stage-1: 29.8 -> 35.8 us).
// (cudaOccupancyMaxActiveClusters reports one CTA per SM for this kernel - 15 / 33 / 74 clusters of 8 / 4 / 2 - and
// was not usable for this plan.)
static int wg_cluster_slots(int c) { return c >= 8 ? 256 : c == 4 ? 272 : c == 2 ? 288 : 296; }

// slices of the pixels and cluster size (slices per cluster): one wave of resident CTAs; the cluster size trades
// resident CTAs (fewer with larger clusters) against partial tiles to write and reduce (one per cluster)
static void wg_plan(long long M, int N, int K, int* splits, int* cs) {
  const int BNK = wg_bnk(K);
  const int tiles = ((N + WG_BM - 1) / WG_BM) * (K / BNK);
  const long long cap = (M + 127) / 128;            // at least 2 chunks per slice
  double best = 1e30;
  *splits = 1;
  *cs = 1;
  static const int no_cluster = [] { const char* e = std::getenv("DAT_B200_WG_CLUSTER"); return e && e[0] == '0' ? 1 : 0; }();
  for (int c = no_cluster ? 1 : 8; c >= 1; c >>= 1) {     // DAT_B200_WG_CLUSTER=0: one partial per slice (round 1), for A/B
    const int slots = wg_cluster_slots(c);
    long long s = slots / tiles / c * c;            // whole clusters, one wave
    if (s > cap) s = cap / c * c;
    if (s < c) {
      if (c > 1) continue;
      s = 1;
    }
    const double chunks = (double)((M + s - 1) / s + WG_CHUNK - 1) / WG_CHUNK;
    const double t_mma = chunks * (BNK == 128 ? 370.0 : 190.0);                       // cycles per 64-pixel chunk
    const double parts = (double)(s / c);
    const double t_red = parts > 1 ? parts * N * K * 8.0 / 2000.0 + 4000.0 : 0.0;     // partial write + read, + a launch
    if (t_mma + t_red < best) {
      best = t_mma + t_red;
      *splits = (int)s;
      *cs = c;
    }
  }
  if (std::getenv("DAT_B200_WG_DEBUG"))
    fprintf(stderr, "wgrad plan M=%lld N=%d K=%d: %d tiles x %d slices, clusters of %d\n", M, N, K, tiles, *splits, *cs);
}
int pointwise_wgrad_tc_splits(long long M, int N, int K) {
  int s, c;
  wg_plan(M, N, K, &s, &c);
  return s;
}

size_t pointwise_wgrad_tc_workspace(long long M, int N, int K) {
  int s, c;
  wg_plan(M, N, K, &s, &c);
  const size_t parts = (size_t)(s / c);             // one partial tile per cluster
  return align_up(parts * N * K * 4, 256) + align_up(parts * N * 4, 256);
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
  int splits, CS;
  wg_plan(M, N, K, &splits, &CS);
  const int parts = splits / CS;
  long long rps = (M + splits - 1) / splits;
  rps = (rps + WG_CHUNK - 1) / WG_CHUNK * WG_CHUNK;
  CUtensorMap tmDY, tmX;
  DAT_FWD(tc::make_tmap_2d(&tmDY, dY, 2, false, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, 64, 64, 128));
  DAT_FWD(tc::make_tmap_2d(&tmX, X, 2, false, (uint64_t)M, (uint64_t)K, (uint64_t)K * 2, 64, 64, 128));
  const int stage_bytes = 2 * BOX_BYTES + (BNK / 64) * BOX_BYTES;
  int stages = 100 * 1024 / stage_bytes;      // two CTAs per SM
  if (stages > 6) stages = 6;
  size_t buf = (size_t)stages * stage_bytes;
  const size_t out_stage = (size_t)WG_BM * (BNK * 4 + 16) + WG_BM * 4;   // parked accumulator tile + bias-gradient column
  if (out_stage > buf) buf = out_stage;
  const size_t smem = 1024 + 1024 + ONES_BYTES + buf;
  int tmem_cols = 32;
  while (tmem_cols < BNK + 32) tmem_cols <<= 1;      // + the bias-gradient columns (read 32 wide)
  float* part = parts > 1 ? (float*)ws : dW;
  float* bpart = nullptr;
  if (db != nullptr)
    bpart = parts > 1 ? (float*)((char*)ws + align_up((size_t)parts * N * K * 4, 256)) : db;
  dim3 grid((N + WG_BM - 1) / WG_BM, K / BNK, splits);
  DAT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_k_cluster(gemm_tc_wgrad_kernel, grid, WG_THREADS, smem, st, dim3(1, 1, CS), tmDY, tmX, part, bpart, N, K, BNK,
                   (int)rps, M, stages, tmem_cols, CS);
  DAT_LAUNCH_OK("gemm_tc_wgrad_kernel");
  if (parts > 1) {
    const long long nk = (long long)N * K;
    launch_k(wgrad_reduce_kernel, (unsigned)ceil_div(nk + N, 256ll), 256, 0, st, part, bpart, parts, nk, N, dW, db);
    DAT_LAUNCH_OK("wgrad_reduce_kernel");
  }
  return DAT_OK;
}

}  // namespace dat
