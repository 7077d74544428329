// K2 time-parallel scan for small blocks (D / BD-LNCDE, bs <= 8) - tuning "k2_pscan" (written blind, opt-in).
//
// The reference evaluates the product of the per-interval block matrices with jax.lax.associative_scan inside chunks of
// `parallel_steps` and a sequential lax.scan over the chunks (LogLinearCDE scanner, models/LinearNeuralCDEs.py:355-386).
// The default kernels of linear_scan.cu apply the flows one after the other (one thread per state element, 1 499
// dependent steps for EigenWorms) - at batch 32 that is 4 096 threads on a 148-SM machine.  Here the time axis is cut
// into chunks of kPscanChunk steps and the recurrence runs on two levels, with all chunks of all samples in flight:
//
//   forward   1. pscan_prod_kernel      P_c = F_{hi-1} ... F_{lo}            one thread per (sample, chunk, block, column)
//             2. scan_fwd_warp_kernel   Y_{c+1} = P_c Y_c over the chunks    (the default kernel, on the chunk products)
//             3. pscan_apply_kernel     y_t = F_t y_{t-1} inside each chunk, started from Y_c
//   reverse   1. pscan_prod_kernel (again: the call must not depend on what the forward left in the workspace)
//             2. pscan_bwd_local_kernel<false>   q_c = the chunk's own cotangent contributions (incoming lam = 0)
//             3. pscan_bwd_boundary_kernel       Lam_c = q_c + P_c^T Lam_{c+1} over the chunks
//             4. pscan_bwd_local_kernel<true>    lam_{t-1} = g_{t-1} + F_t^T lam_t, dF_t = lam_t (x) y_{t-1} in place
//
// The dependent chain drops from T - 1 steps to 2 kPscanChunk + (T - 1) / kPscanChunk (111 for EigenWorms); the price is
// bs + 2 times the multiply-adds and reading the flows twice - both far below what the machine has idle at this batch.
// The association differs from the sequential order, so results agree with the default kernels to rounding (not bit for
// bit); chunk boundaries are taken from the product recurrence and every state in between from its chunk's own sweep.
// Thread mapping inside a block is that of the default kernels (the bs lanes of a block are neighbours in a warp,
// exchange by shuffles); every loop has a warp-uniform trip count, lanes beyond their range only skip loads / stores.
#pragma once
#include <cuda_runtime.h>

namespace lncde {

constexpr int kPscanChunk = 32;   // steps per chunk
constexpr int kPscanMaxBlock = 8; // block sizes served (the chunk products hold a bs-vector per thread and read bs^2 per step)

// chunk c covers the flows t in [1 + c L, min(T, 1 + (c + 1) L)); flow 0 is never applied
__host__ __device__ inline int pscan_chunks(int T) { return (T - 1 + kPscanChunk - 1) / kPscanChunk; }
inline bool pscan_shape(int T, int bs) { return bs <= kPscanMaxBlock && (32 % bs) == 0 && T - 1 > 2 * kPscanChunk; }
// floats behind the flows / partials: P [B][nc + 1][H][bs] | Yb [B][nc + 1][H] | Q [B][nc][H] | Lam [B][nc + 1][H]
inline size_t pscan_floats(int B, int T, int H, int bs) {
  const size_t nc = (size_t)pscan_chunks(T);
  return (size_t)B * H * ((nc + 1) * bs + (nc + 1) + nc + (nc + 1));
}

// thread (b, c, h = blk * BS + j) owns COLUMN j of the chunk product.  P is stored as the "flow" of step c + 1 of a
// sequence of nc + 1 steps, so that the boundary recurrence is the default forward scan on it.
template <int BS>
__global__ void __launch_bounds__(256) pscan_prod_kernel(const float* __restrict__ F, float* __restrict__ P, int B, int T,
                                                          int H, int nc) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)B * nc * H) return;  // no collectives in this kernel
  const int h = (int)(gid % H);
  const long long bc = gid / H;
  const int c = (int)(bc % nc), b = (int)(bc / nc);
  const int j = h % BS, blk0 = h - j;
  const int lo = 1 + c * kPscanChunk, hi = min(T, lo + kPscanChunk);
  float p[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) p[i] = (i == j) ? 1.f : 0.f;
  const float* Fb = F + ((size_t)b * T * H + blk0) * BS;  // block of step 0: BS * BS contiguous floats
  for (int t = lo; t < hi; ++t) {
    const float* Ft = Fb + (size_t)t * H * BS;
    float q[BS];
    if constexpr (BS % 4 == 0) {  // the block is BS * BS contiguous floats at a 16-byte aligned offset: 128-bit loads
#pragma unroll
      for (int i = 0; i < BS; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < BS; k += 4) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(Ft + i * BS + k));
          acc = fmaf(f.x, p[k], acc);
          acc = fmaf(f.y, p[k + 1], acc);
          acc = fmaf(f.z, p[k + 2], acc);
          acc = fmaf(f.w, p[k + 3], acc);
        }
        q[i] = acc;
      }
    } else {
#pragma unroll
      for (int i = 0; i < BS; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < BS; ++k) acc = fmaf(__ldg(Ft + i * BS + k), p[k], acc);
        q[i] = acc;
      }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i) p[i] = q[i];
  }
  float* Pc = P + (((size_t)b * (nc + 1) + c + 1) * H + blk0) * BS;
#pragma unroll
  for (int i = 0; i < BS; ++i) Pc[i * BS + j] = p[i];
  if (c == 0) {  // step 0 of the product sequence is never applied; keep it defined
    float* P0 = P + ((size_t)b * (nc + 1) * H + blk0) * BS;
#pragma unroll
    for (int i = 0; i < BS; ++i) P0[i * BS + j] = (i == j) ? 1.f : 0.f;
  }
}

// thread (b, c, h) owns ROW h of its block: y_t[h] = sum_j F_t[h][j] y_{t-1}[blk, j]
template <int BS>
__global__ void __launch_bounds__(256) pscan_apply_kernel(const float* __restrict__ F, const float* __restrict__ Yb,
                                                           float* __restrict__ ys, int B, int T, int H, int nc) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < (long long)B * nc * H;
  const int h = live ? (int)(gid % H) : 0;
  const long long bc = live ? gid / H : 0;
  const int c = (int)(bc % nc), b = (int)(bc / nc);
  const int lane = threadIdx.x & 31, base = lane & ~(BS - 1);
  const int lo = 1 + c * kPscanChunk, hi = live ? min(T, lo + kPscanChunk) : 0;
  const float* Fb = F + ((size_t)b * T * H + h) * BS;
  float* yb = ys + (size_t)b * T * H + h;
  float y = live ? Yb[((size_t)b * (nc + 1) + c) * H + h] : 0.f;
  if (live && c == 0) yb[0] = y;
  const size_t step = (size_t)H * BS;
  for (int u = 0; u < kPscanChunk; ++u) {  // warp-uniform trip count
    const int t = lo + u;
    const bool on = t < hi;
    float row[BS];
#pragma unroll
    for (int j = 0; j < BS; ++j) row[j] = on ? __ldg(Fb + (size_t)t * step + j) : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < BS; ++j) acc = fmaf(row[j], __shfl_sync(0xffffffffu, y, base + j), acc);
    if (on) {
      y = acc;
      yb[(size_t)t * H] = y;
    }
  }
}

// thread (b, c, h = blk * BS + j) owns COLUMN j: lam_{t-1}[j] = g_{t-1}[j] + sum_i F_t[i][j] lam_t[i].
// WRITE = false: incoming lam = 0, result q_c (the chunk's own contributions);  WRITE = true: incoming Lam_{c+1}, and
// dF_t[i][j] = lam_t[i] y_{t-1}[j] overwrites the flows in place.
template <int BS, bool WRITE>
__global__ void __launch_bounds__(256) pscan_bwd_local_kernel(float* __restrict__ F, const float* __restrict__ ys,
                                                               const float* __restrict__ g_ys, const float* __restrict__ Lam,
                                                               float* __restrict__ Q, int B, int T, int H, int nc) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < (long long)B * nc * H;
  const int h = live ? (int)(gid % H) : 0;
  const long long bc = live ? gid / H : 0;
  const int c = (int)(bc % nc), b = (int)(bc / nc);
  const int lane = threadIdx.x & 31, base = lane & ~(BS - 1), j = lane & (BS - 1);
  const int lo = 1 + c * kPscanChunk, hi = live ? min(T, lo + kPscanChunk) : 0;
  const size_t step = (size_t)H * BS;
  float* Fb = F + ((size_t)b * T * H + (h - j)) * BS + j;  // F[blk][0][j] of step 0
  const float* yb = ys + (size_t)b * T * H + h;
  const float* gb = g_ys + (size_t)b * T * H + h;
  float lam = 0.f;
  if (WRITE) {
    lam = live ? Lam[((size_t)b * (nc + 1) + c + 1) * H + h] : 0.f;
    if (live && c == 0) {
#pragma unroll
      for (int i = 0; i < BS; ++i) Fb[i * BS] = 0.f;  // flow 0 is never applied
    }
  }
  for (int u = 0; u < kPscanChunk; ++u) {  // t = hi - 1 - u down to lo; warp-uniform trip count
    const int t = hi - 1 - u;
    const bool on = live && t >= lo;
    float col[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) col[i] = on ? Fb[(size_t)t * step + i * BS] : 0.f;
    const float yp = (WRITE && on) ? yb[(size_t)(t - 1) * H] : 0.f;
    float acc = on ? gb[(size_t)(t - 1) * H] : 0.f;
#pragma unroll
    for (int i = 0; i < BS; ++i) {
      const float li = __shfl_sync(0xffffffffu, lam, base + i);
      acc = fmaf(col[i], li, acc);
      if (WRITE && on) Fb[(size_t)t * step + i * BS] = li * yp;
    }
    if (on) lam = acc;
  }
  if (!WRITE && live) Q[((size_t)b * nc + c) * H + h] = lam;
}

// thread (b, h = blk * BS + j): Lam_nc = g_{T-1};  Lam_c[j] = q_c[j] + sum_i P_c[i][j] Lam_{c+1}[i];  g_y0 = Lam_0
template <int BS>
__global__ void __launch_bounds__(256) pscan_bwd_boundary_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                                  const float* __restrict__ g_ys, float* __restrict__ Lam,
                                                                  float* __restrict__ g_y0, int B, int T, int H, int nc) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < (long long)B * H;
  const int b = live ? (int)(gid / H) : 0, h = live ? (int)(gid - (long long)b * H) : 0;
  const int lane = threadIdx.x & 31, base = lane & ~(BS - 1), j = lane & (BS - 1);
  float lam = live ? g_ys[((size_t)b * T + (T - 1)) * H + h] : 0.f;
  if (live) Lam[((size_t)b * (nc + 1) + nc) * H + h] = lam;
  for (int c = nc - 1; c >= 0; --c) {
    const float* Pc = P + (((size_t)b * (nc + 1) + c + 1) * H + (h - j)) * BS + j;  // column j of P_c
    float acc = live ? Q[((size_t)b * nc + c) * H + h] : 0.f;
#pragma unroll
    for (int i = 0; i < BS; ++i) {
      const float li = __shfl_sync(0xffffffffu, lam, base + i);
      acc = fmaf(live ? Pc[i * BS] : 0.f, li, acc);
    }
    lam = acc;
    if (live) Lam[((size_t)b * (nc + 1) + c) * H + h] = lam;
  }
  if (live) g_y0[(size_t)b * H + h] = lam;
}

}  // namespace lncde
