// Bilinear sampling of the key/value feature map at the deformed points, and its
// gradient (F.grid_sample, bilinear / zeros / align_corners=True; dat_blocks.py:169-172).
//
// Forward: features are channel-last, so each of the 4 taps of a sample is one contiguous
// Cg-channel row; a thread owns 4 consecutive channels (16-byte loads for fp32, 8-byte for
// bf16) and consecutive threads own consecutive chunks, so every tap row is read with
// fully coalesced vector loads.  HBM/L2-bound: algorithmic bytes per launch =
// 4*B*Ns*C*e (tap rows) + 8*B*G*Ns (pos) + B*Ns*C*e (xs).
//
// Backward (d x): atomics-free and deterministic.  Per (batch, group) the <= 4*Ns
// (sample, tap) entries are sorted by destination pixel in shared memory (bitonic sort of
// packed 32-bit keys); each run of equal pixels is summed by one warp in ascending
// (sample, tap) order and added to that pixel's dx row exactly once.
#include "common.cuh"
#include "kernels.h"

namespace dat {

namespace {

template <typename TX, typename TO>
__global__ void sample_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ pos,
                                  TO* __restrict__ xs, int32_t* __restrict__ taps, int B, int H,
                                  int W, int C, int G, int Cg, int Ns, long long total) {
  pdl_enter();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4n = C >> 2;
  const int c = (int)(idx % c4n) * 4;
  const int n = (int)((idx / c4n) % Ns);
  const int b = (int)(idx / ((long long)c4n * Ns));
  const int g = c / Cg;
  const long long sp = ((long long)b * G + g) * Ns + n;
  const float py = pos[sp * 2], px = pos[sp * 2 + 1];
  const Taps t = make_taps(px, py, W, H);
  if (taps != nullptr && (c % Cg) == 0) {
    taps[sp * 2] = t.y0;
    taps[sp * 2 + 1] = t.x0;
  }
  const TX* base = x + (long long)b * H * W * C + c;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto tap = [&](bool ok, int yy, int xx, float w) {
    if (!ok) return;
    float4 v = load4(base + ((long long)yy * W + xx) * C);
    acc.x = fmaf(v.x, w, acc.x);
    acc.y = fmaf(v.y, w, acc.y);
    acc.z = fmaf(v.z, w, acc.z);
    acc.w = fmaf(v.w, w, acc.w);
  };
  tap(t.vx0 && t.vy0, t.y0, t.x0, __fmul_rn(t.wx0, t.wy0));
  tap(t.vx1 && t.vy0, t.y0, t.x0 + 1, __fmul_rn(t.wx1, t.wy0));
  tap(t.vx0 && t.vy1, t.y0 + 1, t.x0, __fmul_rn(t.wx0, t.wy1));
  tap(t.vx1 && t.vy1, t.y0 + 1, t.x0 + 1, __fmul_rn(t.wx1, t.wy1));
  store4(xs + ((long long)b * Ns + n) * C + c, acc);
}

// d pos of the feature sampling + reduction of the rpe-bias part.
// One warp per sample point (b, g, n); lanes stride 4-channel chunks of the group.
//   dpos_bias_part: (B, heads, qsplit, Ns, 2) partial sums from the attention backward.
template <typename TX, typename TD>
__global__ void sample_bwd_dpos_kernel(const TX* __restrict__ x, const float* __restrict__ pos,
                                       const TD* __restrict__ dxs,
                                       const float* __restrict__ dpos_bias_part, int qsplit,
                                       float* __restrict__ dpos, int B, int H, int W, int C, int G,
                                       int Cg, int hg, int Ns, long long n_points) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long sp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sp >= n_points) return;
  const int n = (int)(sp % Ns);
  const int g = (int)((sp / Ns) % G);
  const int b = (int)(sp / ((long long)Ns * G));
  const float py = pos[sp * 2], px = pos[sp * 2 + 1];
  const Taps t = make_taps(px, py, W, H);
  const TX* base = x + (long long)b * H * W * C + g * Cg;
  const TD* drow = dxs + ((long long)b * Ns + n) * C + g * Cg;
  float gix = 0.f, giy = 0.f;
  for (int ch = lane * 4; ch < Cg; ch += 128) {
    float4 d = load4(drow + ch);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nw = (t.vx0 && t.vy0) ? load4(base + ((long long)t.y0 * W + t.x0) * C + ch) : z;
    float4 ne = (t.vx1 && t.vy0) ? load4(base + ((long long)t.y0 * W + t.x0 + 1) * C + ch) : z;
    float4 sw = (t.vx0 && t.vy1) ? load4(base + ((long long)(t.y0 + 1) * W + t.x0) * C + ch) : z;
    float4 se = (t.vx1 && t.vy1) ? load4(base + ((long long)(t.y0 + 1) * W + t.x0 + 1) * C + ch) : z;
    float top = d.x * (ne.x - nw.x) + d.y * (ne.y - nw.y) + d.z * (ne.z - nw.z) + d.w * (ne.w - nw.w);
    float bot = d.x * (se.x - sw.x) + d.y * (se.y - sw.y) + d.z * (se.z - sw.z) + d.w * (se.w - sw.w);
    float lef = d.x * (sw.x - nw.x) + d.y * (sw.y - nw.y) + d.z * (sw.z - nw.z) + d.w * (sw.w - nw.w);
    float rig = d.x * (se.x - ne.x) + d.y * (se.y - ne.y) + d.z * (se.z - ne.z) + d.w * (se.w - ne.w);
    gix += top * t.wy0 + bot * t.wy1;
    giy += lef * t.wx0 + rig * t.wx1;
  }
  gix = warp_sum(gix);
  giy = warp_sum(giy);
  if (lane == 0) {
    float gy = giy * (0.5f * (float)(H - 1));
    float gx = gix * (0.5f * (float)(W - 1));
    if (dpos_bias_part != nullptr) {
      const int heads = G * hg;
      for (int hh = 0; hh < hg; ++hh)
        for (int z = 0; z < qsplit; ++z) {
          const float* pp = dpos_bias_part +
                            ((((long long)b * heads + g * hg + hh) * qsplit + z) * Ns + n) * 2;
          gy += pp[0];
          gx += pp[1];
        }
    }
    dpos[sp * 2] = gy;
    dpos[sp * 2 + 1] = gx;
  }
}

// d x scatter, sorted-run form.  grid = (B*G, Cg/32); block = min(P, 1024) threads: one (sample, tap) entry per thread,
// so the read-modify-write of every run head is ONE round trip to memory for the whole CTA (with 256 threads a thread
// walked 4 entries one after the other, each a dependent load -> FMA -> store chain through the same dx array).
// dynamic smem: keys[P] (uint32) + weight[4*Ns] (float), P = pow2 >= 4*Ns.
// key = pixel << 14 | entry  (entry = n*4 + tap < 2^14, pixel < 2^18); invalid = ~0u.
template <typename TD>
__global__ void __launch_bounds__(1024)
sample_bwd_dx_kernel(const float* __restrict__ pos, const TD* __restrict__ dxs,
                     float* __restrict__ dx, int H, int W, int C, int G, int Cg, int Ns, int P) {
  pdl_enter();
  extern __shared__ uint32_t smem_u[];
  uint32_t* keys = smem_u;
  float* wts = reinterpret_cast<float*>(smem_u + P);
  const int bg = blockIdx.x, b = bg / G, g = bg % G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int e = tid; e < P; e += blockDim.x) {
    uint32_t key = 0xffffffffu;
    if (e < 4 * Ns) {
      const int n = e >> 2, tp = e & 3;
      const long long sp = (long long)bg * Ns + n;
      const Taps t = make_taps(pos[sp * 2 + 1], pos[sp * 2], W, H);
      const int xx = t.x0 + (tp & 1), yy = t.y0 + (tp >> 1);
      const bool ok = ((tp & 1) ? t.vx1 : t.vx0) && ((tp >> 1) ? t.vy1 : t.vy0);
      const float w = __fmul_rn((tp & 1) ? t.wx1 : t.wx0, (tp >> 1) ? t.wy1 : t.wy0);
      wts[e] = w;
      if (ok) key = ((uint32_t)(yy * W + xx) << 14) | (uint32_t)e;
    }
    keys[e] = key;
  }
  __syncthreads();
  if (P == (int)blockDim.x) {
    // one key per thread: the compare-exchange stages with a partner inside the warp (j < 32: 40 of the 55 stages at
    // P = 1024) run on registers with one shuffle each; only the cross-warp stages go through shared memory and a
    // barrier (ncu: the all-shared-memory sort was ~35 % of this kernel's samples)
    uint32_t key = keys[tid];
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        uint32_t other;
        if (j >= 32) {
          keys[tid] = key;
          __syncthreads();
          other = keys[tid ^ j];
          __syncthreads();
        } else {
          other = __shfl_xor_sync(0xffffffffu, key, j);
        }
        const bool take_min = ((tid & k) == 0) == ((tid & j) == 0);
        key = take_min ? min(key, other) : max(key, other);
      }
    }
    keys[tid] = key;
    __syncthreads();
  } else {
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P; i += blockDim.x) {
          int ixj = i ^ j;
          if (ixj > i) {
            uint32_t a = keys[i], c2 = keys[ixj];
            bool up = (i & k) == 0;
            if ((a > c2) == up) { keys[i] = c2; keys[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
  }
  // One THREAD per run head, 32 channels (this CTA's slice) as 8 independent 4-wide
  // vector accumulators: every global access is a full 32-byte+ segment and the runs of
  // different threads overlap their memory latency (the previous warp-per-run loop
  // serialised ~1 us of latency per run).
  (void)lane; (void)warp; (void)nwarps;
  const int c = g * Cg + blockIdx.y * 32;
  const TD* dbase = dxs + (long long)b * Ns * C + c;
  float* xbase = dx + (long long)b * H * W * C + c;
  for (int p = tid; p < P; p += blockDim.x) {
    const uint32_t key = keys[p];
    if (key == 0xffffffffu) break;          // sorted: only invalid entries follow
    const uint32_t pix = key >> 14;
    if (p > 0 && (keys[p - 1] >> 14) == pix) continue;  // not a run head
    float4* xrow = reinterpret_cast<float4*>(xbase + (long long)pix * C);
    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = xrow[i];
    int r = p;
    uint32_t kr = key;
    do {
      const int e = (int)(kr & 0x3fffu);
      const float w = wts[e];
      const TD* drow = dbase + (long long)(e >> 2) * C;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 d = load4(drow + 4 * i);
        acc[i].x = fmaf(w, d.x, acc[i].x);
        acc[i].y = fmaf(w, d.y, acc[i].y);
        acc[i].z = fmaf(w, d.z, acc[i].z);
        acc[i].w = fmaf(w, d.w, acc[i].w);
      }
      ++r;
      kr = r < P ? keys[r] : 0xffffffffu;
    } while (kr != 0xffffffffu && (kr >> 14) == pix);
#pragma unroll
    for (int i = 0; i < 8; ++i) xrow[i] = acc[i];
  }
}

}  // namespace

int sample_fwd(const Shape& s, const void* x, const float* pos, void* xs, int32_t* taps,
               cudaStream_t st) {
  long long total = (long long)s.B * s.Ns * (s.C / 4);
  int grid = ceil_div(total, 256);
#define LAUNCH(TX, TO)                                                                        \
  launch_k(sample_fwd_kernel<TX, TO>, grid, 256, 0, st, (const TX*)x, pos, (TO*)xs, taps, s.B, s.H, \
                                                  s.W, s.C, s.G, s.Cg, s.Ns, total)
  if (s.x_dtype == DAT_F32 && s.act_dtype == DAT_F32) LAUNCH(float, float);
  else if (s.x_dtype == DAT_F32) LAUNCH(float, bf16);
  else if (s.act_dtype == DAT_F32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
  DAT_LAUNCH_OK("sample_fwd_kernel");
  return DAT_OK;
}

int sample_bwd_dpos(const Shape& s, const void* x, const float* pos, const void* dxs,
                    const float* dpos_bias_part, int qsplit, float* dpos, cudaStream_t st) {
  long long pts = (long long)s.B * s.G * s.Ns;
  int grid = ceil_div(pts, 8);
#define LAUNCH(TX, TD)                                                                         \
  launch_k(sample_bwd_dpos_kernel<TX, TD>, grid, 256, 0, st, (const TX*)x, pos, (const TD*)dxs,      \
                                                       dpos_bias_part, qsplit, dpos, s.B, s.H, \
                                                       s.W, s.C, s.G, s.Cg, s.hg, s.Ns, pts)
  if (s.x_dtype == DAT_F32 && s.act_dtype == DAT_F32) LAUNCH(float, float);
  else if (s.x_dtype == DAT_F32) LAUNCH(float, bf16);
  else if (s.act_dtype == DAT_F32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
  DAT_LAUNCH_OK("sample_bwd_dpos_kernel");
  return DAT_OK;
}

// dx (fp32, already holding the proj_q path gradient) += scatter of dxs through the taps.
int sample_bwd_dx(const Shape& s, const float* pos, const void* dxs, float* dx, cudaStream_t st) {
  DAT_REQUIRE(4 * s.Ns <= (1 << 14), "sample_bwd: Ns=%d > 4096 unsupported", s.Ns);
  DAT_REQUIRE(s.HW <= (1 << 18), "sample_bwd: H*W=%d > 262144 unsupported", s.HW);
  int P = 1;
  while (P < 4 * s.Ns) P <<= 1;
  size_t smem = (size_t)P * 4 + (size_t)4 * s.Ns * 4;
  dim3 grid(s.B * s.G, s.Cg / 32);
#define LAUNCH(TD)                                                                             \
  do {                                                                                         \
    auto kern = sample_bwd_dx_kernel<TD>;                                                      \
    if (smem > 48 * 1024)                                                                      \
      DAT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                       (int)smem));                                            \
    launch_k(kern, grid, P < 1024 ? P : 1024, smem, st, pos, (const TD*)dxs, dx, s.H, s.W, s.C, s.G, s.Cg, s.Ns, P); \
  } while (0)
  if (s.act_dtype == DAT_F32) LAUNCH(float); else LAUNCH(bf16);
#undef LAUNCH
  DAT_LAUNCH_OK("sample_bwd_dx_kernel");
  return DAT_OK;
}

}  // namespace dat
