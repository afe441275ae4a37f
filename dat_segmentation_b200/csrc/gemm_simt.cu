// fp32-accumulate CUDA-core GEMMs for the four 1x1 convolutions of the block
// (proj_q / proj_k / proj_v / proj_out, dat_blocks.py:61-79) and their gradients.
//
// This is the exact-precision path (fp32 mode must match the reference to 1e-5
// relative, which rules out bf16/tf32 tensor-core math) and the generic fallback shape
// handler.  The bf16 hot path uses the tcgen05 kernels in gemm_tc.cu instead.
//
// One kernel covers the three contraction forms through strides:
//     C[i, j] = sum_l A(i, l) * B(l, j)
//   forward   (TN): i = pixel m, j = out-channel n, l = in-channel k
//   data grad (NN): i = pixel m, j = in-channel k,  l = out-channel n
//   weight grad(TT): i = out-channel n, j = in-channel k, l = pixel m (split over l)
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;
constexpr int GEMM_THREADS = 256;

// A_LC: A is contiguous along l (else along i).  B_LC: B contiguous along l (else j).
template <typename TA, typename TB, typename TC, bool A_LC, bool B_LC>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(const TA* __restrict__ A, long long sai, long long sal,
                 const TB* __restrict__ Bm, long long sbl, long long sbj,
                 TC* __restrict__ C, long long ldc, const float* __restrict__ bias,
                 int accumulate, int I, int J, int L, int l_chunk,
                 float* __restrict__ partial) {
  pdl_enter();
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int l_begin = blockIdx.z * l_chunk;
  const int l_end = min(L, l_begin + l_chunk);

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int lt = l_begin; lt < l_end; lt += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + GEMM_THREADS * e;
      int li, ii;
      if (A_LC) { li = idx & (BK - 1); ii = idx >> 4; } else { ii = idx & (BM - 1); li = idx >> 6; }
      int gi = i0 + ii, gl = lt + li;
      float v = 0.f;
      if (gi < I && gl < l_end) v = to_f32(A[(long long)gi * sai + (long long)gl * sal]);
      As[li][ii] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + GEMM_THREADS * e;
      int li, jj;
      if (B_LC) { li = idx & (BK - 1); jj = idx >> 4; } else { jj = idx & (BN - 1); li = idx >> 6; }
      int gj = j0 + jj, gl = lt + li;
      float v = 0.f;
      if (gj < J && gl < l_end) v = to_f32(Bm[(long long)gl * sbl + (long long)gj * sbj]);
      Bs[li][jj] = v;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < BK; ++l) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[l][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[l][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w};
      float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int gi = i0 + ty * 4 + a;
    if (gi >= I) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int gj = j0 + tx * 4 + b;
      if (gj >= J) continue;
      float v = acc[a][b];
      if (partial != nullptr) {
        partial[((long long)blockIdx.z * I + gi) * J + gj] = v;
      } else {
        if (bias != nullptr) v += bias[gj];
        TC* dst = C + (long long)gi * ldc + gj;
        if (accumulate) v += to_f32(*dst);
        *dst = from_f32<TC>(v);
      }
    }
  }
}

// out[i, j] = sum_z partial[z][i][j] (+ bias[j]); fixed summation order (deterministic).
template <typename TC>
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nsplit,
                                       long long count, int J, const float* __restrict__ bias,
                                       TC* __restrict__ out) {
  pdl_enter();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  float v = 0.f;
  for (int z = 0; z < nsplit; ++z) v += partial[(long long)z * count + idx];
  if (bias != nullptr) v += bias[idx % J];
  out[idx] = from_f32<TC>(v);
}

// Two reductions of the same shape in one launch (dK and dV partials of the attention backward), 4 elements per thread:
// out{A,B}[i] = sum_z part{A,B}[z][i]; blockIdx.y selects the pair member.  count % 4 == 0.
template <typename TC>
__global__ void reduce_partials_pair_kernel(const float* __restrict__ partA, const float* __restrict__ partB, int nsplit,
                                            long long count, TC* __restrict__ outA, TC* __restrict__ outB) {
  pdl_enter();
  const float* part = blockIdx.y == 0 ? partA : partB;
  TC* out = blockIdx.y == 0 ? outA : outB;
  const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (idx >= count) return;
  float4 v = *reinterpret_cast<const float4*>(part + idx);
  for (int z = 1; z < nsplit; ++z) {
    const float4 o = *reinterpret_cast<const float4*>(part + (long long)z * count + idx);
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  store4(out + idx, v);
}

// Column sums of a (M, N) row-major matrix over a row range: partial[y][n].
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ X, long long M, int N, long long rows_per_block,
                              float* __restrict__ partial) {
  pdl_enter();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + cx;
  long long m0 = (long long)blockIdx.y * rows_per_block;
  long long m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (long long m = m0 + ry; m < m1; m += 8) s += to_f32(X[m * N + n]);
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][cx];
    partial[(long long)blockIdx.y * N + n] = t;
  }
}

template <typename TA, typename TB, typename TC, bool A_LC, bool B_LC>
int launch_gemm(const void* A, long long sai, long long sal, const void* Bm, long long sbl,
                long long sbj, void* C, long long ldc, const float* bias, int accumulate, int I,
                int J, int L, int nsplit, float* partial, cudaStream_t st) {
  dim3 grid(ceil_div(J, BN), ceil_div(I, BM), nsplit);
  int l_chunk = ceil_div(ceil_div(L, nsplit), BK) * BK;
  launch_k(gemm_simt_kernel<TA, TB, TC, A_LC, B_LC>, grid, GEMM_THREADS, 0, st, 
      (const TA*)A, sai, sal, (const TB*)Bm, sbl, sbj, (TC*)C, ldc, bias, accumulate, I, J, L,
      l_chunk, partial);
  DAT_LAUNCH_OK("gemm_simt_kernel");
  return DAT_OK;
}

#define DISPATCH2(dtA, dtC, CALL)                                       \
  do {                                                                  \
    if (dtA == DAT_F32 && dtC == DAT_F32) { CALL(float, float); }       \
    else if (dtA == DAT_F32 && dtC == DAT_BF16) { CALL(float, bf16); }  \
    else if (dtA == DAT_BF16 && dtC == DAT_F32) { CALL(bf16, float); }  \
    else { CALL(bf16, bf16); }                                          \
  } while (0)

}  // namespace

int pointwise_fwd_simt(const void* X, int x_dt, const float* W, const float* b, void* Y, int y_dt,
                       long long M, int N, int K, cudaStream_t st) {
  DAT_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31), "pointwise_fwd: bad sizes M=%lld N=%d K=%d", M, N, K);
#define CALL(TA, TC)                                                                          \
  return launch_gemm<TA, float, TC, true, true>(X, K, 1, W, 1, K, Y, N, b, 0, (int)M, N, K, 1, \
                                                nullptr, st)
  DISPATCH2(x_dt, y_dt, CALL);
#undef CALL
}

// dX[M,K] (+)= dY[M,N] W[N,K]
int pointwise_dgrad_simt(const void* dY, int dy_dt, const float* W, void* dX, int dx_dt,
                         long long M, int N, int K, int accumulate, cudaStream_t st) {
  DAT_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31), "pointwise_dgrad: bad sizes");
#define CALL(TA, TC)                                                                       \
  return launch_gemm<TA, float, TC, true, false>(dY, N, 1, W, K, 1, dX, K, nullptr,        \
                                                 accumulate, (int)M, K, N, 1, nullptr, st)
  DISPATCH2(dy_dt, dx_dt, CALL);
#undef CALL
}

size_t pointwise_wgrad_workspace(long long M, int N, int K) {
  int nsplit = wgrad_splits(M, N, K);
  size_t a = (size_t)nsplit * N * K * sizeof(float);
  size_t b = (size_t)colsum_splits(M) * N * sizeof(float);
  return align_up(a, 256) + align_up(b, 256);
}

int wgrad_splits(long long M, int N, int K) {
  // enough CTAs for ~2 waves of 148 SMs, each split at least 256 rows deep
  long long tiles = (long long)ceil_div(N, BM) * ceil_div(K, BN);
  long long want = (2 * 148 + tiles - 1) / tiles;
  long long cap = (M + 255) / 256;
  long long s = want < cap ? want : cap;
  return (int)(s < 1 ? 1 : s);
}
int colsum_splits(long long M) {
  long long s = (M + 511) / 512;
  return (int)(s > 64 ? 64 : (s < 1 ? 1 : s));
}

// dW[N,K] = dY[M,N]^T X[M,K];  db[N] = column sums of dY (db may be NULL).
int pointwise_wgrad_simt(const void* dY, int dy_dt, const void* X, int x_dt, float* dW, float* db,
                         long long M, int N, int K, void* ws, size_t ws_bytes, cudaStream_t st) {
  DAT_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31), "pointwise_wgrad: bad sizes");
  DAT_REQUIRE(ws_bytes >= pointwise_wgrad_workspace(M, N, K), "pointwise_wgrad: workspace too small");
  int nsplit = wgrad_splits(M, N, K);
  float* part = (float*)ws;
  float* cpart = (float*)((char*)ws + align_up((size_t)nsplit * N * K * sizeof(float), 256));
  int rc;
#define CALL(TA, TB)                                                                          \
  rc = launch_gemm<TA, TB, float, false, false>(dY, 1, N, X, K, 1, dW, K, nullptr, 0, N, K,   \
                                                (int)M, nsplit, nsplit > 1 ? part : nullptr, st)
  DISPATCH2(dy_dt, x_dt, CALL);
#undef CALL
  DAT_FWD(rc);
  if (nsplit > 1) {
    long long count = (long long)N * K;
    launch_k(reduce_partials_kernel<float>, ceil_div(count, 256), 256, 0, st, part, nsplit, count, K,
                                                                        nullptr, dW);
    DAT_LAUNCH_OK("reduce_partials_kernel");
  }
  if (db != nullptr) {
    int cs = colsum_splits(M);
    long long rows = (M + cs - 1) / cs;
    dim3 grid(ceil_div(N, 32), cs);
    if (dy_dt == DAT_F32)
      launch_k(colsum_kernel<float>, grid, 256, 0, st, (const float*)dY, M, N, rows, cpart);
    else
      launch_k(colsum_kernel<bf16>, grid, 256, 0, st, (const bf16*)dY, M, N, rows, cpart);
    DAT_LAUNCH_OK("colsum_kernel");
    launch_k(reduce_partials_kernel<float>, ceil_div(N, 256), 256, 0, st, cpart, cs, N, N, nullptr, db);
    DAT_LAUNCH_OK("reduce_partials_kernel");
  }
  return DAT_OK;
}

// db[N] = column sums of dY[M,N] (two-stage, fixed order).  ws: colsum_splits(M) * N floats.
int bias_grad(const void* dY, int dy_dt, float* db, long long M, int N, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  const int cs = colsum_splits(M);
  DAT_REQUIRE(ws_bytes >= (size_t)cs * N * sizeof(float), "bias_grad: workspace too small");
  float* cpart = (float*)ws;
  const long long rows = (M + cs - 1) / cs;
  dim3 grid(ceil_div(N, 32), cs);
  if (dy_dt == DAT_F32) launch_k(colsum_kernel<float>, grid, 256, 0, st, (const float*)dY, M, N, rows, cpart);
  else launch_k(colsum_kernel<bf16>, grid, 256, 0, st, (const bf16*)dY, M, N, rows, cpart);
  DAT_LAUNCH_OK("colsum_kernel");
  launch_k(reduce_partials_kernel<float>, ceil_div(N, 256), 256, 0, st, cpart, cs, N, N, nullptr, db);
  DAT_LAUNCH_OK("reduce_partials_kernel");
  return DAT_OK;
}

int reduce_partials_pair(const float* partA, const float* partB, int nsplit, long long count, void* outA, void* outB,
                         int out_dt, cudaStream_t st) {
  DAT_REQUIRE(count % 4 == 0, "reduce_partials_pair: count must be a multiple of 4");
  dim3 grid(ceil_div(count / 4, 256), 2);
  if (out_dt == DAT_F32)
    launch_k(reduce_partials_pair_kernel<float>, grid, 256, 0, st, partA, partB, nsplit, count, (float*)outA, (float*)outB);
  else
    launch_k(reduce_partials_pair_kernel<bf16>, grid, 256, 0, st, partA, partB, nsplit, count, (bf16*)outA, (bf16*)outB);
  DAT_LAUNCH_OK("reduce_partials_pair_kernel");
  return DAT_OK;
}

// Generic deterministic reduction used by other stages: out[idx] = sum_z part[z][idx].
int reduce_partials(const float* part, int nsplit, long long count, void* out, int out_dt,
                    cudaStream_t st) {
  if (out_dt == DAT_F32)
    launch_k(reduce_partials_kernel<float>, ceil_div(count, 256), 256, 0, st, part, nsplit, count, 1,
                                                                        nullptr, (float*)out);
  else
    launch_k(reduce_partials_kernel<bf16>, ceil_div(count, 256), 256, 0, st, part, nsplit, count, 1,
                                                                       nullptr, (bf16*)out);
  DAT_LAUNCH_OK("reduce_partials_kernel");
  return DAT_OK;
}

}  // namespace dat
