// K2d - fused diagonal LNCDE (D-LNCDE): log-signature rows -> flows -> product scan in ONE pass.
//
// Replaces the diagonal branch of models/LinearNeuralCDEs.py:302-316 with its scanner (:355-386):
//   f_t[h] = 1 + sum_c logsig[t, 1 + c] * vf_A[c, h]   (level-1 columns only, quirk B7)
//   y_t    = f_t * y_{t-1},  t = 1 .. T-1               (flow 0 is never applied, quirk B4)
// The generic K2 path (linear_scan.cu) materialises the flows F[B,T,H] in HBM (flow-build GEMM), reads them
// back in the scan and, in the backward pass, overwrites them with dF for a second GEMM: 3 x 4 B per
// (sample, step, state) of avoidable traffic each way.  Here a thread owns one (sample, state) pair, keeps
// its C vector-field coefficients in registers and rebuilds f_t from the C level-1 coordinates of row t, which
// a CTA stages through shared memory in chunks (packed 8 floats per step: two broadcast 128-bit loads per
// step for the whole warp, next chunk prefetched into registers while the current one is consumed).
// Compulsory traffic per (sample, step): forward 4 H B written (ys) + the logsig row; backward 8 H B read
// (ys, g_ys) + the row - the kernel is HBM-bound once B x H threads fill the machine.
// The reference's associative-scan chunking (parallel_steps) only changes the fp32 association order.
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"

namespace lncde {

constexpr int kDThreads = 128;  // states per CTA
constexpr int kDChunk = 64;     // time steps staged per chunk
constexpr int kDPerThread = (kDChunk * 8 + kDThreads - 1) / kDThreads;  // packed floats each thread stages

// level-1 coordinates of rows [t0, t0 + nt) of one sample, packed [step][8] (columns >= C are zero)
template <int C>
__device__ __forceinline__ void fetch_rows(const float* __restrict__ ls, int D, int t0, int nt, float (&r)[kDPerThread]) {
#pragma unroll
  for (int k = 0; k < kDPerThread; ++k) {
    const int e = threadIdx.x + k * kDThreads, t = e >> 3, c = e & 7;
    r[k] = (t < nt && c < C) ? __ldg(ls + (size_t)(t0 + t) * D + 1 + c) : 0.f;
  }
}
__device__ __forceinline__ void put_rows(float* sm, const float (&r)[kDPerThread]) {
#pragma unroll
  for (int k = 0; k < kDPerThread; ++k) sm[threadIdx.x + k * kDThreads] = r[k];
}

template <int C>
__device__ __forceinline__ float flow_of(const float (&A)[C], const float* row8) {
  const float4 a = *reinterpret_cast<const float4*>(row8);
  const float4 b = *reinterpret_cast<const float4*>(row8 + 4);
  const float l[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float f = 1.f;
#pragma unroll
  for (int c = 0; c < C; ++c) f = fmaf(l[c], A[c], f);
  return f;
}

// MEAN: also y_mean[b][h] = mean_t ys[b][t][h] (the LNCDE classification read-out, LinearNeuralCDEs.py:389) - summed per
// chunk of 64 steps and then over the chunks, in time order, by the thread that owns (b, h): no second pass over ys.
template <int C, bool MEAN>
__global__ void __launch_bounds__(kDThreads) dlncde_fwd_kernel(const float* __restrict__ vf_At,
                                                               const float* __restrict__ logsig,
                                                               const float* __restrict__ y0, float* __restrict__ ys,
                                                               float* __restrict__ y_mean, int T, int D, int H) {
  __shared__ __align__(16) float rows[kDChunk * 8];
  const int b = blockIdx.y, h = blockIdx.x * kDThreads + threadIdx.x;
  const bool live = h < H;
  const float* ls = logsig + (size_t)b * T * D;
  float* yb = ys + (size_t)b * T * H + (live ? h : 0);
  float A[C];
#pragma unroll
  for (int c = 0; c < C; ++c) A[c] = live ? vf_At[(size_t)h * C + c] : 0.f;
  float y = live ? y0[(size_t)b * H + h] : 0.f;
  if (live) yb[0] = y;
  float total = y;
  float pre[kDPerThread];
  fetch_rows<C>(ls, D, 1, min(kDChunk, T - 1), pre);
  for (int t0 = 1; t0 < T; t0 += kDChunk) {
    const int nt = min(kDChunk, T - t0);
    __syncthreads();  // everybody is done with the previous chunk
    put_rows(rows, pre);
    __syncthreads();
    if (t0 + kDChunk < T) fetch_rows<C>(ls, D, t0 + kDChunk, min(kDChunk, T - t0 - kDChunk), pre);
    float csum = 0.f;
#pragma unroll 4
    for (int t = 0; t < nt; ++t) {
      y *= flow_of<C>(A, rows + t * 8);
      if (MEAN) csum += y;
      if (live) yb[(size_t)(t0 + t) * H] = y;
    }
    total += csum;
  }
  if (MEAN && live) y_mean[(size_t)b * H + h] = total / (float)T;
}

// lam_{T-1} = g_{T-1};  t = T-1 .. 1:  dF_t = lam_t * y_{t-1},  dA[c] += dF_t * logsig[t, 1 + c],
// lam_{t-1} = f_t * lam_t + g_{t-1};  g_y0 = lam_0.  partial[b][h][c] is summed over b afterwards.
// MEAN: g_ys[b][t][h] = g_mean[b][h] / T for every t (the cotangent of the mean over time): `g_ys` then points at the
// [B][H] array and nothing of size [B][T][H] is written or read for it.
// (Capping the kernel at 32 registers for 16 CTAs per SM - one wave at B = 2048 - was measured SLOWER, 1.30 vs 0.98 ms
// per step: the unrolled loop needs its registers to keep four steps' loads in flight.)
template <int C, bool MEAN>
__global__ void __launch_bounds__(kDThreads) dlncde_bwd_kernel(const float* __restrict__ vf_At,
                                                               const float* __restrict__ logsig,
                                                               const float* __restrict__ ys,
                                                               const float* __restrict__ g_ys, float* __restrict__ g_y0,
                                                               float* __restrict__ partial, int T, int D, int H) {
  __shared__ __align__(16) float rows[kDChunk * 8];
  const int b = blockIdx.y, h = blockIdx.x * kDThreads + threadIdx.x;
  const bool live = h < H;
  const float* ls = logsig + (size_t)b * T * D;
  const float* yb = ys + (size_t)b * T * H + (live ? h : 0);
  const float* gb = MEAN ? g_ys : g_ys + (size_t)b * T * H + (live ? h : 0);
  const float gm = (MEAN && live) ? g_ys[(size_t)b * H + h] / (float)T : 0.f;
  float A[C], acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    A[c] = live ? vf_At[(size_t)h * C + c] : 0.f;
    acc[c] = 0.f;
  }
  float lam = MEAN ? gm : (live ? gb[(size_t)(T - 1) * H] : 0.f);
  // chunks cover steps [t_lo, t_hi]; walked from the last chunk to the first, each from its last step down
  const int n_chunks = (T - 1 + kDChunk - 1) / kDChunk;
  float pre[kDPerThread];
  if (n_chunks > 0) {
    const int t_lo = 1 + (n_chunks - 1) * kDChunk;
    fetch_rows<C>(ls, D, t_lo, T - t_lo, pre);
  }
  for (int ch = n_chunks - 1; ch >= 0; --ch) {
    const int t_lo = 1 + ch * kDChunk, nt = min(kDChunk, T - t_lo);
    __syncthreads();
    put_rows(rows, pre);
    __syncthreads();
    if (ch > 0) fetch_rows<C>(ls, D, t_lo - kDChunk, kDChunk, pre);
#pragma unroll 4
    for (int t = nt - 1; t >= 0; --t) {
      const float* row8 = rows + t * 8;
      const float yp = live ? yb[(size_t)(t_lo + t - 1) * H] : 0.f;
      const float gp = MEAN ? gm : (live ? gb[(size_t)(t_lo + t - 1) * H] : 0.f);
      const float4 r0 = *reinterpret_cast<const float4*>(row8);
      const float4 r1 = *reinterpret_cast<const float4*>(row8 + 4);
      const float l[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      const float dF = lam * yp;
      float f = 1.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        f = fmaf(l[c], A[c], f);
        acc[c] = fmaf(dF, l[c], acc[c]);
      }
      lam = fmaf(f, lam, gp);
    }
  }
  if (live) {
    g_y0[(size_t)b * H + h] = lam;
    float* p = partial + ((size_t)b * H + h) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] = acc[c];
  }
}

// out[e] = sum_b partial[b][e] in two levels (B = 2048 would otherwise be a 2048-long chain on H C = 896 threads):
// chunks of kDRedChunk samples are summed IN PLACE into the chunk's first row, then the chunk sums are added in order.
constexpr int kDRedChunk = 32;
__global__ void dlncde_reduce_chunks_kernel(float* __restrict__ partial, int n, int B) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x, b0 = blockIdx.y * kDRedChunk;
  if (e >= n) return;
  float s = 0.f;
  for (int b = b0; b < min(B, b0 + kDRedChunk); ++b) s += partial[(size_t)b * n + e];
  partial[(size_t)b0 * n + e] = s;
}
__global__ void dlncde_reduce_small_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int B) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += partial[(size_t)b * n + e];
    out[e] = s;
  }
}
__global__ void dlncde_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int B) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; b += kDRedChunk) s += partial[(size_t)b * n + e];
    out[e] = s;
  }
}

}  // namespace lncde

template <bool MEAN>
static int dlncde_fwd_launch(cudaStream_t cs, const float* vf_At, const float* logsig, const float* y0, float* ys,
                             float* y_mean, int B, int T, int D, int H, int C) {
  using namespace lncde;
  const dim3 grid((H + kDThreads - 1) / kDThreads, B);
  switch (C) {
    case 1: dlncde_fwd_kernel<1, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    case 2: dlncde_fwd_kernel<2, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    case 3: dlncde_fwd_kernel<3, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    case 4: dlncde_fwd_kernel<4, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    case 5: dlncde_fwd_kernel<5, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    case 6: dlncde_fwd_kernel<6, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    case 7: dlncde_fwd_kernel<7, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
    default: dlncde_fwd_kernel<8, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, y0, ys, y_mean, T, D, H); break;
  }
  return (int)cudaGetLastError();
}

template <bool MEAN>
static int dlncde_bwd_launch(cudaStream_t cs, const float* vf_At, const float* logsig, const float* ys, const float* g,
                             float* g_vf_At, float* g_y0, int B, int T, int D, int H, int C, float* partial) {
  using namespace lncde;
  const dim3 grid((H + kDThreads - 1) / kDThreads, B);
  switch (C) {
    case 1: dlncde_bwd_kernel<1, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    case 2: dlncde_bwd_kernel<2, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    case 3: dlncde_bwd_kernel<3, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    case 4: dlncde_bwd_kernel<4, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    case 5: dlncde_bwd_kernel<5, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    case 6: dlncde_bwd_kernel<6, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    case 7: dlncde_bwd_kernel<7, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
    default: dlncde_bwd_kernel<8, MEAN><<<grid, kDThreads, 0, cs>>>(vf_At, logsig, ys, g, g_y0, partial, T, D, H); break;
  }
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
  const int n = H * C;
  if (B > kDRedChunk) {
    const dim3 rgrid((n + 255) / 256, (B + kDRedChunk - 1) / kDRedChunk);
    dlncde_reduce_chunks_kernel<<<rgrid, 256, 0, cs>>>(partial, n, B);
  }
  int blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  // B <= kDRedChunk: the first pass is skipped and the second adds the samples one by one (stride 1)
  if (B > kDRedChunk) dlncde_reduce_kernel<<<blocks, 256, 0, cs>>>(partial, g_vf_At, n, B);
  else dlncde_reduce_small_kernel<<<blocks, 256, 0, cs>>>(partial, g_vf_At, n, B);
  return (int)cudaGetLastError();
}

extern "C" {

size_t lncde_dlncde_bwd_workspace_bytes(int B, int H, int C) {
  if (B < 1 || H < 1 || C < 1) return 0;
  return (size_t)B * H * C * sizeof(float);
}

static int dlncde_check(int B, int T, int D, int H, int C) {
  if (B < 1 || T < 1 || H < 1 || C < 1 || D < 1 + C) return LNCDE_ERR_SHAPE;
  if (C > 8 || B > 65535) return LNCDE_ERR_UNSUPPORTED;
  return LNCDE_OK;
}

int lncde_dlncde_fwd(void* stream, const float* vf_At, const float* logsig, const float* y0, float* ys, int B, int T,
                     int D, int H, int C) {
  if (!vf_At || !logsig || !y0 || !ys) return LNCDE_ERR_NULL;
  int st = dlncde_check(B, T, D, H, C);
  if (st != LNCDE_OK) return st;
  return dlncde_fwd_launch<false>((cudaStream_t)stream, vf_At, logsig, y0, ys, nullptr, B, T, D, H, C);
}

int lncde_dlncde_fwd_mean(void* stream, const float* vf_At, const float* logsig, const float* y0, float* ys,
                          float* y_mean, int B, int T, int D, int H, int C) {
  if (!vf_At || !logsig || !y0 || !ys || !y_mean) return LNCDE_ERR_NULL;
  int st = dlncde_check(B, T, D, H, C);
  if (st != LNCDE_OK) return st;
  return dlncde_fwd_launch<true>((cudaStream_t)stream, vf_At, logsig, y0, ys, y_mean, B, T, D, H, C);
}

static int dlncde_bwd_args(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f,
                           const void* w, int B, int T, int D, int H, int C, size_t workspace_bytes) {
  if (!a || !b || !c || !d || !e || !f || !w) return LNCDE_ERR_NULL;
  int st = dlncde_check(B, T, D, H, C);
  if (st != LNCDE_OK) return st;
  if (workspace_bytes < lncde_dlncde_bwd_workspace_bytes(B, H, C)) return LNCDE_ERR_WORKSPACE;
  return LNCDE_OK;
}

int lncde_dlncde_bwd(void* stream, const float* vf_At, const float* logsig, const float* ys, const float* g_ys,
                     float* g_vf_At, float* g_y0, int B, int T, int D, int H, int C, void* workspace,
                     size_t workspace_bytes) {
  int st = dlncde_bwd_args(vf_At, logsig, ys, g_ys, g_vf_At, g_y0, workspace, B, T, D, H, C, workspace_bytes);
  if (st != LNCDE_OK) return st;
  return dlncde_bwd_launch<false>((cudaStream_t)stream, vf_At, logsig, ys, g_ys, g_vf_At, g_y0, B, T, D, H, C,
                                  (float*)workspace);
}

int lncde_dlncde_bwd_mean(void* stream, const float* vf_At, const float* logsig, const float* ys, const float* g_mean,
                          float* g_vf_At, float* g_y0, int B, int T, int D, int H, int C, void* workspace,
                          size_t workspace_bytes) {
  int st = dlncde_bwd_args(vf_At, logsig, ys, g_mean, g_vf_At, g_y0, workspace, B, T, D, H, C, workspace_bytes);
  if (st != LNCDE_OK) return st;
  return dlncde_bwd_launch<true>((cudaStream_t)stream, vf_At, logsig, ys, g_mean, g_vf_At, g_y0, B, T, D, H, C,
                                 (float*)workspace);
}

}  // extern "C"
