// K2 – structured linear CDE (D / BD / DE-LNCDE): flows F_t = I + sum_l logsig[t,l] A_l and
// the recurrence h_t = F_t h_{t-1}, t = 1..T-1 (flow 0 is never applied, quirk B4).
//
// Replaces models/LinearNeuralCDEs.py:302-316 (diagonal), :318-353 (block) and the scanner
// :355-386.  The reference materialises nothing it does not need either way; here the work is
// split by what bounds it on B200:
//   1. flow_build_kernel   – [B*T, n_basis] x [n_basis, H*bs] FP32 GEMM (+I), all SMs busy,
//                            flows written once to HBM (D: 24 MB, BD: 98 MB, DE: 3.1 GB per
//                            32 x 1499 batch – small against 180 GB / 6.5 TB/s)
//   2. scan_fwd_kernel     – per (sample) CTA streams its flows once, block mat-vec per step
//   3. scan_bwd_kernel     – reverse sweep lam_{t-1} = F_t^T lam_t + g_{t-1}; overwrites F_t in
//                            place with dF_t = lam_t (x) h_{t-1} (block-restricted)
//   4. outer_gemm_kernel   – dLie = dF^T @ logsig, split-K over B*T, deterministic reduce
// The associative-scan chunking of the reference (parallel_steps) changes only the fp32
// association order; the sequential order used here is the parallel_steps = 1 path.
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"
#include "linear_cluster.cuh"
#include "linear_de.cuh"
#include "linear_pscan.cuh"
#include "outer_gemm.cuh"
#include "ptx.cuh"

namespace lncde {

// ---------------------------------------------------------------- 1. flow build
// F[m][n] = diag(n) + sum_l ls[m][1+l] * lie[n][l];  m = b*T + t, n = (blk,i,j) flattened.
template <int TM, int TN>
__global__ void __launch_bounds__(256) flow_build_kernel(const float* __restrict__ lie,
                                                         const float* __restrict__ logsig, float* __restrict__ F,
                                                         long long Mtot, int N, int D, int n_basis, int bs) {
  constexpr int KT = 32;
  __shared__ float As[KT][16 * TM + 1];  // logsig tile, k-major
  __shared__ float Bs[KT][16 * TN + 1];  // lie tile, k-major
  const long long m0 = (long long)blockIdx.y * 16 * TM;
  const int n0 = blockIdx.x * 16 * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < n_basis; k0 += KT) {
    for (int e = threadIdx.x; e < KT * 16 * TM; e += 256) {
      const int m = e / KT, kk = e - m * KT;  // consecutive threads walk k: contiguous in logsig rows
      As[kk][m] = (m0 + m < Mtot && k0 + kk < n_basis) ? logsig[(m0 + m) * D + 1 + k0 + kk] : 0.f;
    }
    for (int e = threadIdx.x; e < KT * 16 * TN; e += 256) {
      const int n = e / KT, kk = e - n * KT;
      Bs[kk][n] = (n0 + n < N && k0 + kk < n_basis) ? lie[(size_t)(n0 + n) * n_basis + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int bs2 = bs * bs;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty + 16 * i;
    if (m >= Mtot) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      const int r = n % bs2;
      const float eye = (r / bs == r % bs) ? 1.f : 0.f;
      F[m * N + n] = acc[i][j] + eye;
    }
  }
}

// ---------------------------------------------------------------- 2. forward scan
// One CTA per sample, H threads... generalised: thread groups of `lanes` threads per row.
// small blocks (bs <= 32): one thread per row; big blocks: one warp per row group, lanes over j.
__global__ void scan_fwd_small_kernel(const float* __restrict__ F, const float* __restrict__ y0,
                                      float* __restrict__ ys, int T, int H, int bs) {
  extern __shared__ float sh[];  // y double buffer [2][H]
  const int b = blockIdx.x;
  const int N = H * bs;
  const float* Fb = F + (size_t)b * T * N;
  float* yb = ys + (size_t)b * T * H;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float v = y0[(size_t)b * H + h];
    sh[h] = v;
    yb[h] = v;
  }
  __syncthreads();
  int cur = 0;
  for (int t = 1; t < T; ++t) {
    const float* Ft = Fb + (size_t)t * N;
    const float* yc = sh + cur * H;
    float* yn = sh + (cur ^ 1) * H;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
      const int blk = h / bs;
      const float* row = Ft + (size_t)h * bs;
      const float* yv = yc + blk * bs;
      float acc = 0.f;
      for (int j = 0; j < bs; ++j) acc = fmaf(__ldg(row + j), yv[j], acc);
      yn[h] = acc;
      yb[(size_t)t * H + h] = acc;
    }
    __syncthreads();
    cur ^= 1;
  }
}

// big blocks: warp per row, lanes across the columns of the block (coalesced row reads)
__global__ void scan_fwd_big_kernel(const float* __restrict__ F, const float* __restrict__ y0,
                                    float* __restrict__ ys, int T, int H, int bs) {
  extern __shared__ float sh[];
  const int b = blockIdx.x;
  const int N = H * bs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* Fb = F + (size_t)b * T * N;
  float* yb = ys + (size_t)b * T * H;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float v = y0[(size_t)b * H + h];
    sh[h] = v;
    yb[h] = v;
  }
  __syncthreads();
  int cur = 0;
  for (int t = 1; t < T; ++t) {
    const float* Ft = Fb + (size_t)t * N;
    const float* yc = sh + cur * H;
    float* yn = sh + (cur ^ 1) * H;
    for (int h = warp; h < H; h += nw) {
      const int blk = h / bs;
      const float* row = Ft + (size_t)h * bs;
      const float* yv = yc + blk * bs;
      float acc = 0.f;
      for (int j = lane; j < bs; j += 32) acc = fmaf(__ldg(row + j), yv[j], acc);
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
      if (lane == 0) {
        yn[h] = acc;
        yb[(size_t)t * H + h] = acc;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

// ---------------------------------------------------------------- 3. backward scan
// lam_{T-1} = g_{T-1};  for t = T-1 .. 1:  dF_t = lam_t (x) h_{t-1} (in place over F_t),
// lam_{t-1} = F_t^T lam_t + g_{t-1};  g_y0 = lam_0.  Thread per column j of the state.
__global__ void scan_bwd_kernel(float* __restrict__ F, const float* __restrict__ ys,
                                const float* __restrict__ g_ys, float* __restrict__ g_y0, int T, int H, int bs) {
  extern __shared__ float sh[];  // lam double buffer [2][H], yprev [H]
  const int b = blockIdx.x;
  const int N = H * bs;
  float* Fb = F + (size_t)b * T * N;
  const float* yb = ys + (size_t)b * T * H;
  const float* gb = g_ys + (size_t)b * T * H;
  float* yprev = sh + 2 * H;
  for (int h = threadIdx.x; h < H; h += blockDim.x) sh[h] = gb[(size_t)(T - 1) * H + h];
  // flow 0 is never applied: its gradient is zero
  for (int e = threadIdx.x; e < N; e += blockDim.x) Fb[e] = 0.f;
  __syncthreads();
  int cur = 0;
  for (int t = T - 1; t >= 1; --t) {
    float* Ft = Fb + (size_t)t * N;
    const float* lam = sh + cur * H;
    float* ln = sh + (cur ^ 1) * H;
    for (int h = threadIdx.x; h < H; h += blockDim.x) yprev[h] = yb[(size_t)(t - 1) * H + h];
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
      const int blk = h / bs, j = h - blk * bs;
      float* col = Ft + (size_t)blk * bs * bs + j;  // F[blk][i][j], i strided by bs
      const float* lv = lam + blk * bs;
      const float yp = yprev[h];
      float acc = gb[(size_t)(t - 1) * H + h];
      for (int i = 0; i < bs; ++i) {
        const float f = col[(size_t)i * bs];
        acc = fmaf(f, lv[i], acc);
        col[(size_t)i * bs] = lv[i] * yp;
      }
      ln[h] = acc;
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int h = threadIdx.x; h < H; h += blockDim.x) g_y0[(size_t)b * H + h] = sh[cur * H + h];
}


// ---------------------------------------------------------------- 2b/3b. warp-shuffle scans
// Block sizes that divide 32: the bs threads of one block sit in one warp, the state lives in
// registers and is exchanged with shuffles - no shared memory, no barriers; flows are
// prefetched PF steps ahead so the dependent chain per step is bs x (SHFL + FMA).
template <int BS, int PF>
__global__ void __launch_bounds__(256) scan_fwd_warp_kernel(const float* __restrict__ F, const float* __restrict__ y0,
                                                            float* __restrict__ ys, int B, int T, int H) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < (long long)B * H;
  const int b = live ? (int)(gid / H) : 0, h = live ? (int)(gid - (long long)b * H) : 0;
  const int lane = threadIdx.x & 31, base = lane & ~(BS - 1);
  const float* Fb = F + ((size_t)b * T) * H * BS + (size_t)h * BS;  // row h of step 0
  float* yb = ys + (size_t)b * T * H + h;
  float y = live ? y0[(size_t)b * H + h] : 0.f;
  if (live) yb[0] = y;
  float row[PF][BS];
  const size_t step = (size_t)H * BS;
#pragma unroll
  for (int u = 0; u < PF; ++u)
#pragma unroll
    for (int j = 0; j < BS; ++j) row[u][j] = (live && 1 + u < T) ? __ldg(Fb + (size_t)(1 + u) * step + j) : 0.f;
  for (int t0 = 1; t0 < T; t0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int t = t0 + u;
      if (t < T) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < BS; ++j) acc = fmaf(row[u][j], __shfl_sync(0xffffffffu, y, base + j), acc);
        y = acc;
        if (live) yb[(size_t)t * H] = y;
        // refill this slot with the flow PF steps ahead
        const int tn = t + PF;
#pragma unroll
        for (int j = 0; j < BS; ++j) row[u][j] = (live && tn < T) ? __ldg(Fb + (size_t)tn * step + j) : 0.f;
      }
    }
  }
}

// thread (b, h = blk*BS + j) owns COLUMN j of its block:
// lam_{t-1}[j] = g_{t-1}[j] + sum_i F_t[i][j] lam_t[i];  dF_t[i][j] = lam_t[i] * y_{t-1}[j]
template <int BS, int PF>
__global__ void __launch_bounds__(256) scan_bwd_warp_kernel(float* __restrict__ F, const float* __restrict__ ys,
                                                            const float* __restrict__ g_ys, float* __restrict__ g_y0,
                                                            int B, int T, int H) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < (long long)B * H;
  const int b = live ? (int)(gid / H) : 0, h = live ? (int)(gid - (long long)b * H) : 0;
  const int lane = threadIdx.x & 31, base = lane & ~(BS - 1), j = lane & (BS - 1);
  const size_t step = (size_t)H * BS;
  float* Fb = F + ((size_t)b * T) * H * BS + (size_t)(h - j) * BS + j;  // F[blk][0][j] of step 0
  const float* yb = ys + (size_t)b * T * H + h;
  const float* gb = g_ys + (size_t)b * T * H + h;
  float lam = live ? gb[(size_t)(T - 1) * H] : 0.f;
  if (live) {
#pragma unroll
    for (int i = 0; i < BS; ++i) Fb[i * BS] = 0.f;  // flow 0 is never applied
  }
  float col[PF][BS], yp[PF], gp[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) {
    const int t = T - 1 - u;
#pragma unroll
    for (int i = 0; i < BS; ++i) col[u][i] = (live && t >= 1) ? Fb[(size_t)t * step + i * BS] : 0.f;
    yp[u] = (live && t >= 1) ? yb[(size_t)(t - 1) * H] : 0.f;
    gp[u] = (live && t >= 1) ? gb[(size_t)(t - 1) * H] : 0.f;
  }
  for (int t0 = T - 1; t0 >= 1; t0 -= PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int t = t0 - u;
      if (t >= 1) {
        float acc = gp[u];
        float* Ft = Fb + (size_t)t * step;
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          const float li = __shfl_sync(0xffffffffu, lam, base + i);
          acc = fmaf(col[u][i], li, acc);
          if (live) Ft[i * BS] = li * yp[u];
        }
        lam = acc;
        const int tn = t - PF;
#pragma unroll
        for (int i = 0; i < BS; ++i) col[u][i] = (live && tn >= 1) ? Fb[(size_t)tn * step + i * BS] : 0.f;
        yp[u] = (live && tn >= 1) ? yb[(size_t)(tn - 1) * H] : 0.f;
        gp[u] = (live && tn >= 1) ? gb[(size_t)(tn - 1) * H] : 0.f;
      }
    }
  }
  if (live) g_y0[(size_t)b * H + h] = lam;
}

template <int BS, int PF>
static void launch_scan_warp(bool bwd, cudaStream_t cs, float* F, const float* y0, float* ys, const float* g_ys,
                             float* g_y0, int B, int T, int H) {
  const long long n = (long long)B * H;
  const int threads = 256, grid = (int)((n + threads - 1) / threads);
  if (!bwd)
    scan_fwd_warp_kernel<BS, PF><<<grid, threads, 0, cs>>>(F, y0, ys, B, T, H);
  else
    scan_bwd_warp_kernel<BS, PF><<<grid, threads, 0, cs>>>(F, ys, g_ys, g_y0, B, T, H);
}
static bool scan_warp_dispatch(bool bwd, cudaStream_t cs, float* F, const float* y0, float* ys, const float* g_ys,
                               float* g_y0, int B, int T, int H, int bs) {
  switch (bs) {
    case 1: launch_scan_warp<1, 8>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H); return true;
    case 2: launch_scan_warp<2, 8>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H); return true;
    case 4: launch_scan_warp<4, 6>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H); return true;
    case 8: launch_scan_warp<8, 4>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H); return true;
    case 16: launch_scan_warp<16, 2>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H); return true;
    case 32: launch_scan_warp<32, 2>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H); return true;
  }
  return false;
}

// ---------------------------------------------------------------- 2b'. time-parallel scans ("k2_pscan", linear_pscan.cuh)
template <int BS, int PF>
static void launch_pscan(bool bwd, cudaStream_t cs, float* F, const float* y0, float* ys, const float* g_ys, float* g_y0,
                         int B, int T, int H, float* scratch) {
  const int nc = pscan_chunks(T);
  float* P = scratch;                                  // [B][nc + 1][H][BS]
  float* Yb = P + (size_t)B * (nc + 1) * H * BS;       // [B][nc + 1][H]
  float* Q = Yb + (size_t)B * (nc + 1) * H;            // [B][nc][H]
  float* Lam = Q + (size_t)B * nc * H;                 // [B][nc + 1][H]
  const int threads = 256;
  const int grid_c = (int)(((long long)B * nc * H + threads - 1) / threads);
  const int grid_b = (int)(((long long)B * H + threads - 1) / threads);
  pscan_prod_kernel<BS><<<grid_c, threads, 0, cs>>>(F, P, B, T, H, nc);
  if (!bwd) {
    scan_fwd_warp_kernel<BS, PF><<<grid_b, threads, 0, cs>>>(P, y0, Yb, B, nc + 1, H);
    pscan_apply_kernel<BS><<<grid_c, threads, 0, cs>>>(F, Yb, ys, B, T, H, nc);
  } else {
    pscan_bwd_local_kernel<BS, false><<<grid_c, threads, 0, cs>>>(F, ys, g_ys, Lam, Q, B, T, H, nc);
    pscan_bwd_boundary_kernel<BS><<<grid_b, threads, 0, cs>>>(P, Q, g_ys, Lam, g_y0, B, T, H, nc);
    pscan_bwd_local_kernel<BS, true><<<grid_c, threads, 0, cs>>>(F, ys, g_ys, Lam, Q, B, T, H, nc);
  }
}
static bool pscan_dispatch(bool bwd, cudaStream_t cs, float* F, const float* y0, float* ys, const float* g_ys, float* g_y0,
                           int B, int T, int H, int bs, float* scratch, unsigned flags) {
  if ((flags & LNCDE_FLAG_GENERIC) || !pscan_shape(T, bs)) return false;
  if (((long long)B * pscan_chunks(T) * H + 255) / 256 > 0x7fffffffLL) return false;
  switch (bs) {
    case 1: launch_pscan<1, 8>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H, scratch); return true;
    case 2: launch_pscan<2, 8>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H, scratch); return true;
    case 4: launch_pscan<4, 6>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H, scratch); return true;
    case 8: launch_pscan<8, 4>(bwd, cs, F, y0, ys, g_ys, g_y0, B, T, H, scratch); return true;
  }
  return false;
}

// ---------------------------------------------------------------- 2c/3c. big-block scans (DE-LNCDE)
// One CTA per sample.  The flows do not depend on the state, so they are streamed through a ring
// of shared-memory stages by 1-D TMA bulk copies (cp.async.bulk + mbarrier, issued NS-1 steps
// ahead by one thread); only the H-vector recurrence is on the dependent chain.
constexpr int kBigThreads = 512;

template <int NS>
__global__ void __launch_bounds__(kBigThreads, 1) scan_fwd_tma_kernel(const float* __restrict__ F,
                                                                       const float* __restrict__ y0,
                                                                       float* __restrict__ ys, int T, int H, int bs) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int N = H * bs;
  const unsigned step_bytes = (unsigned)N * 4;
  float* tiles = reinterpret_cast<float*>(smraw);                 // [NS][N]
  float* ybuf = tiles + (size_t)NS * N;                           // [2][H]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ybuf + 2 * H);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kBigThreads / 32;
  const float* Fb = F + (size_t)b * T * N;
  float* yb = ys + (size_t)b * T * H;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
    fence_mbar_init();
  }
  for (int h = tid; h < H; h += kBigThreads) {
    const float v = y0[(size_t)b * H + h];
    ybuf[h] = v;
    yb[h] = v;
  }
  __syncthreads();
  if (tid == 0) {
    for (int u = 0; u < NS - 1 && 1 + u < T; ++u) {
      mbar_expect_tx(bars + (u % NS), step_bytes);
      tma_load_1d(tiles + (size_t)(u % NS) * N, Fb + (size_t)(1 + u) * N, step_bytes, bars + (u % NS));
    }
  }
  int cur = 0;
  for (int t = 1; t < T; ++t) {
    const int it = t - 1, s = it % NS;
    // keep NS-1 steps in flight: the stage freed at the end of the previous step is refilled now
    if (tid == 0) {
      const int tn = t + NS - 1;
      if (tn < T) {
        const int sn = (it + NS - 1) % NS;
        fence_async_proxy();
        mbar_expect_tx(bars + sn, step_bytes);
        tma_load_1d(tiles + (size_t)sn * N, Fb + (size_t)tn * N, step_bytes, bars + sn);
      }
    }
    mbar_wait(bars + s, (unsigned)((it / NS) & 1));
    const float* Ft = tiles + (size_t)s * N;
    const float* yc = ybuf + cur * H;
    float* yn = ybuf + (cur ^ 1) * H;
    // warp per group of 4 rows, lanes across the columns
    for (int h0 = warp * 4; h0 < H; h0 += nw * 4) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int blk = h0 / bs;
      const float* yv = yc + blk * bs;
      for (int j4 = lane; j4 < bs / 4; j4 += 32) {
        const float4 y4 = *reinterpret_cast<const float4*>(yv + j4 * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 f4 = *reinterpret_cast<const float4*>(Ft + (size_t)(h0 + r) * bs + j4 * 4);
          v[r] = fmaf(f4.x, y4.x, v[r]);
          v[r] = fmaf(f4.y, y4.y, v[r]);
          v[r] = fmaf(f4.z, y4.z, v[r]);
          v[r] = fmaf(f4.w, y4.w, v[r]);
        }
      }
      // transpose-reduce 4 rows over 32 lanes
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool up = lane & 16;
        const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      {
        const bool up = lane & 8;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      if ((lane & 7) == 0) {
        const int r = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
        yn[h0 + r] = v[0];
        yb[(size_t)t * H + h0 + r] = v[0];
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

template <int NS>
__global__ void __launch_bounds__(kBigThreads, 1) scan_bwd_tma_kernel(float* __restrict__ F,
                                                                       const float* __restrict__ ys,
                                                                       const float* __restrict__ g_ys,
                                                                       float* __restrict__ g_y0, int T, int H, int bs) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int N = H * bs;
  const unsigned step_bytes = (unsigned)N * 4;
  float* tiles = reinterpret_cast<float*>(smraw);      // [NS][N]
  float* lamb = tiles + (size_t)NS * N;                // [2][H]
  float* yprev = lamb + 2 * H;                         // [2][H]  (double buffered prefetch)
  float* red = yprev + 2 * H;                          // [RQ][H]
  const int RQ = kBigThreads / H;                      // row groups (H <= 512)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + (size_t)RQ * H);
  const int b = blockIdx.x, tid = threadIdx.x;
  float* Fb = F + (size_t)b * T * N;
  const float* yb = ys + (size_t)b * T * H;
  const float* gb = g_ys + (size_t)b * T * H;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
    fence_mbar_init();
  }
  for (int h = tid; h < H; h += kBigThreads) {
    lamb[h] = gb[(size_t)(T - 1) * H + h];
    if (T >= 2) yprev[h] = yb[(size_t)(T - 2) * H + h];
  }
  for (int e = tid; e < N; e += kBigThreads) Fb[e] = 0.f;  // flow 0 is never applied: zero gradient
  __syncthreads();
  if (tid == 0) {
    for (int u = 0; u < NS - 1 && T - 1 - u >= 1; ++u) {
      mbar_expect_tx(bars + (u % NS), step_bytes);
      tma_load_1d(tiles + (size_t)(u % NS) * N, Fb + (size_t)(T - 1 - u) * N, step_bytes, bars + (u % NS));
    }
  }
  const int h = tid % H, rq = tid / H;
  const bool act = rq < RQ;
  const int blk = h / bs, jj = h - blk * bs;
  int cur = 0;
  for (int t = T - 1; t >= 1; --t) {
    const int it = T - 1 - t, s = it % NS;
    if (tid == 0) {
      const int tn = t - (NS - 1);
      if (tn >= 1) {
        const int sn = (it + NS - 1) % NS;
        fence_async_proxy();
        mbar_expect_tx(bars + sn, step_bytes);
        tma_load_1d(tiles + (size_t)sn * N, Fb + (size_t)tn * N, step_bytes, bars + sn);
      }
    }
    // prefetch y_{t-2} and g for the next step into the other halves
    float y_next = 0.f;
    if (tid < H && t >= 2) y_next = yb[(size_t)(t - 2) * H + tid];
    const float g_prev = tid < H ? gb[(size_t)(t - 1) * H + tid] : 0.f;
    mbar_wait(bars + s, (unsigned)((it / NS) & 1));
    const float* Ft = tiles + (size_t)s * N;
    const float* lam = lamb + cur * H;
    const float* yp = yprev + cur * H;
    float* Fg = Fb + (size_t)t * N;
    float acc = 0.f;
    if (act) {
      const float ypj = yp[h];
      const float* col = Ft + (size_t)blk * bs * bs + jj;
      float* gcol = Fg + (size_t)blk * bs * bs + jj;
      const float* lv = lam + blk * bs;
      for (int i = rq; i < bs; i += RQ) {
        const float li = lv[i];
        acc = fmaf(col[(size_t)i * bs], li, acc);
        gcol[(size_t)i * bs] = li * ypj;  // dF_t[i][j] = lam_t[i] * y_{t-1}[j], in place
      }
      red[rq * H + h] = acc;
    }
    __syncthreads();
    if (tid < H) {
      float sum = g_prev;
      for (int q = 0; q < RQ; ++q) sum += red[q * H + tid];
      lamb[(cur ^ 1) * H + tid] = sum;
      yprev[(cur ^ 1) * H + tid] = y_next;
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int hh = tid; hh < H; hh += kBigThreads) g_y0[(size_t)b * H + hh] = lamb[cur * H + hh];
}

static bool scan_tma_dispatch(bool bwd, cudaStream_t cs, float* F, const float* y0, float* ys, const float* g_ys,
                              float* g_y0, int B, int T, int H, int bs) {
  if (bs % 4 != 0 || H > kBigThreads || kBigThreads % H != 0 || H % 4 != 0) return false;
  const size_t step_bytes = (size_t)H * bs * 4;
  if (step_bytes % 16 != 0) return false;
  const size_t extra = (size_t)(4 * H + (kBigThreads / H) * H) * 4 + 64;
  int ns = 0;
  for (int cand = 4; cand >= 2; --cand)
    if (cand * step_bytes + extra <= 220 * 1024) { ns = cand; break; }
  if (ns == 0) return false;
  const size_t smem = ns * step_bytes + extra;
  auto launch = [&](auto kf, auto kb) {
    if (!bwd) {
      cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kf<<<B, kBigThreads, smem, cs>>>(F, y0, ys, T, H, bs);
    } else {
      cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kb<<<B, kBigThreads, smem, cs>>>(F, ys, g_ys, g_y0, T, H, bs);
    }
  };
  if (ns == 4) launch(scan_fwd_tma_kernel<4>, scan_bwd_tma_kernel<4>);
  else if (ns == 3) launch(scan_fwd_tma_kernel<3>, scan_bwd_tma_kernel<3>);
  else launch(scan_fwd_tma_kernel<2>, scan_bwd_tma_kernel<2>);
  return true;
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int ksplit) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < ksplit; ++k) s += part[(size_t)k * n + e];
    out[e] = s;
  }
}

template <int TM>
static void launch_build_tn(int tn, dim3 grid, cudaStream_t st, const float* lie, const float* ls, float* F,
                            long long Mtot, int N, int D, int nbasis, int bs) {
  switch (tn) {
    case 1: flow_build_kernel<TM, 1><<<grid, 256, 0, st>>>(lie, ls, F, Mtot, N, D, nbasis, bs); break;
    case 2: flow_build_kernel<TM, 2><<<grid, 256, 0, st>>>(lie, ls, F, Mtot, N, D, nbasis, bs); break;
    case 4: flow_build_kernel<TM, 4><<<grid, 256, 0, st>>>(lie, ls, F, Mtot, N, D, nbasis, bs); break;
    default: flow_build_kernel<TM, 8><<<grid, 256, 0, st>>>(lie, ls, F, Mtot, N, D, nbasis, bs); break;
  }
}

constexpr int kScanKsplit = 64;

static size_t flows_floats(int B, int T, int nb, int bs) { return (size_t)B * T * nb * bs * bs; }

// "k2_de_v2" applies to big blocks only (the warp-shuffle scans of bs | 32 keep their kernels)
bool dlie_tcgen05_shape(int bs, int n_basis);
// shapes for which SOME kernel level runs the reverse sweep that stores lam_t only (sizes the workspace: shape-only)
static bool de_v2_shape(int bs, int n_basis) {
  return bs >= 64 && bs % 4 == 0 && (((n_basis + 3) & ~3) <= 36 || dlie_tcgen05_shape(bs, n_basis));
}
// ... and whether this call's level has a bracket-gradient kernel for it: the mma.sync / FFMA triple products are
// compiled for n_basis <= 36, the tcgen05 kernel (level 3) takes any n_basis
static bool lam_path(int level, int bs, int n_basis) {
  if (level < 2 || bs < 64 || bs % 4 != 0) return false;
  if (((n_basis + 3) & ~3) <= 36) return true;
  return level >= 3 && dlie_tcgen05_shape(bs, n_basis);
}


// floats in front of the k2_pscan scratch: flows | split-K partials | lam (dense only) | packed operands (dense only)
static size_t pscan_scratch_offset(int B, int T, int nb, int bs, int n_basis);

}  // namespace lncde

namespace lncde {
// flow_tcgen05.cu: the dense flow build on tcgen05 (hand-written; 3xTF32 split, fp32 accuracy), "k2_de_v2" = 3
size_t flow_tcgen05_scratch_floats(long long Mtot, int N, int n_basis);
bool flow_tcgen05_shape(int N, int bs);
int dlie_tcgen05(cudaStream_t cs, const float* lam, const float* ys, const float* logsig, float* part, int B, int T,
                 int D, int H, int bs, int n_basis, int max_ksplit, int* err, int* ksplit_out);
int flow_build_tcgen05(cudaStream_t cs, const float* lie, const float* logsig, float* F, long long Mtot, int N, int D,
                       int n_basis, int bs, float* scratch, size_t scratch_floats);

static size_t pscan_scratch_offset(int B, int T, int nb, int bs, int n_basis) {
  const size_t flows = (flows_floats(B, T, nb, bs) + 3) & ~(size_t)3;
  const size_t part = ((size_t)nb * bs * bs * n_basis * kScanKsplit + 3) & ~(size_t)3;
  const size_t lam = de_v2_shape(bs, n_basis) ? (size_t)B * T * nb * bs : 0;
  const size_t tc5 = flow_tcgen05_shape(nb * bs * bs, bs)
                         ? ((flow_tcgen05_scratch_floats((long long)B * T, nb * bs * bs, n_basis) + 3) & ~(size_t)3)
                         : 0;
  return flows + part + ((lam + 3) & ~(size_t)3) + tc5;
}
// Kernel level of the dense model for this call: 0 = generic FFMA kernels (flows and dF materialised), 2 = warp-level
// tensor pipe (mma.sync, 3xTF32) flow build + lam-only reverse sweep + triple-product bracket gradient, 3 = the same
// with the flow build and the bracket gradient on tcgen05 (flow_tcgen05.cu, dlie_tcgen05.cu).  Measured on B200,
// EigenWorms DE step with the cluster sweeps: 11.0 / 5.3 / 2.7 ms.
static int dense_level(unsigned flags, int N, int bs) {
  if ((flags & LNCDE_FLAG_GENERIC) || bs < 64) return 0;
  if (!(flags & LNCDE_FLAG_K2_MMA_SYNC) && flow_tcgen05_shape(N, bs)) return 3;
  return 2;
}

// the sweeps of the dense model run on thread-block clusters (linear_cluster.cuh) unless the caller asks for one CTA per series
static bool cluster_sweeps(unsigned flags, int bs) {
  return !(flags & (LNCDE_FLAG_GENERIC | LNCDE_FLAG_K2_ONE_CTA)) && bs >= 64;
}

}  // namespace lncde

extern "C" {

size_t lncde_linear_scan_workspace_bytes(int B, int T, int nb, int bs, int n_basis) {
  using namespace lncde;
  if (B < 1 || T < 1 || nb < 1 || bs < 1 || n_basis < 1) return 0;
  // flows | split-K partials | lam_t of the "k2_de_v2" reverse sweep | packed operands of the tcgen05 flow build
  // ("k2_de_v2" = 3) | chunk products and boundary states of the time-parallel scans ("k2_pscan") - every term is a
  // function of the shape only, so that the size does not depend on the tuning state at the time of the query
  const size_t pscan = pscan_shape(T, bs) ? ((pscan_floats(B, T, nb * bs, bs) + 3) & ~(size_t)3) : 0;
  return (pscan_scratch_offset(B, T, nb, bs, n_basis) + pscan) * 4;
}

static int check_scan_args(int B, int T, int D, int nb, int bs, int n_basis) {
  if (B < 1 || T < 1 || D < 2 || nb < 1 || bs < 1 || n_basis < 1) return LNCDE_ERR_SHAPE;
  if (n_basis > D - 1) return LNCDE_ERR_SHAPE;
  if ((long long)nb * bs > 1024) return LNCDE_ERR_UNSUPPORTED;
  return LNCDE_OK;
}

int lncde_linear_scan_fwd(void* stream, const float* lie, const float* logsig, const float* y0, float* ys,
                          int B, int T, int D, int nb, int bs, int n_basis, void* workspace,
                          size_t workspace_bytes, unsigned flags) {
  using namespace lncde;
  if (!lie || !logsig || !y0 || !ys || !workspace) return LNCDE_ERR_NULL;
  int st = check_scan_args(B, T, D, nb, bs, n_basis);
  if (st != LNCDE_OK) return st;
  if (((uintptr_t)workspace & 15) != 0) return LNCDE_ERR_ALIGN;
  if (workspace_bytes < lncde_linear_scan_workspace_bytes(B, T, nb, bs, n_basis)) return LNCDE_ERR_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  float* F = (float*)workspace;
  const int H = nb * bs, N = H * bs;
  const long long Mtot = (long long)B * T;
  // 1. flows
  const int de_v2 = dense_level(flags, N, bs);
  if (de_v2 >= 3) {  // tcgen05 (3xTF32 split): packed operands behind flows | part | lam
    const size_t flows = (flows_floats(B, T, nb, bs) + 3) & ~(size_t)3;
    const size_t part_n = ((size_t)N * n_basis * kScanKsplit + 3) & ~(size_t)3;
    const size_t lam_n = de_v2_shape(bs, n_basis) ? (size_t)B * T * nb * bs : 0;
    float* scratch = F + flows + part_n + lam_n;
    st = flow_build_tcgen05(cs, lie, logsig, F, Mtot, N, D, n_basis, bs, scratch,
                         flow_tcgen05_scratch_floats(Mtot, N, n_basis));
    if (st != 0) return st;
  } else if (de_v2 >= 2) {  // warp-level tensor pipe (3xTF32), any n_basis
    dim3 grid((unsigned)((Mtot + 127) / 128), (N + 127) / 128);
    cudaFuncSetAttribute(flow_build_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    flow_build_tc_kernel<<<grid, 256, kTcSmemBytes, cs>>>(lie, logsig, F, Mtot, N, D, n_basis, bs);
  } else {
    const int tm = 4, tn = round_tile(N) > 4 ? 4 : round_tile(N);
    if ((Mtot + 16 * tm - 1) / (16 * tm) > 65535) return LNCDE_ERR_UNSUPPORTED;  // grid.y limit: B * T <= 4.19 M rows
    dim3 grid((N + 16 * tn - 1) / (16 * tn), (unsigned)((Mtot + 16 * tm - 1) / (16 * tm)));
    launch_build_tn<4>(tn, grid, cs, lie, logsig, F, Mtot, N, D, n_basis, bs);
  }
  st = (int)cudaGetLastError();
  if (st != 0) return st;
  // 2. scan
  float* pscan_scratch = F + pscan_scratch_offset(B, T, nb, bs, n_basis);
  if (pscan_dispatch(false, cs, F, y0, ys, nullptr, nullptr, B, T, H, bs, pscan_scratch, flags)) {
  } else if (scan_warp_dispatch(false, cs, F, y0, ys, nullptr, nullptr, B, T, H, bs)) {
  } else if (cluster_sweeps(flags, bs) && scan_cluster_dispatch(false, cs, F, y0, ys, nullptr, B, T, H, bs)) {
  } else if (scan_tma_dispatch(false, cs, F, y0, ys, nullptr, nullptr, B, T, H, bs)) {
  } else if (bs <= 16) {
    int threads = H < 32 ? 32 : (H > 1024 ? 1024 : ((H + 31) & ~31));
    scan_fwd_small_kernel<<<B, threads, 2 * H * sizeof(float), cs>>>(F, y0, ys, T, H, bs);
  } else {
    int threads = H >= 32 ? 1024 : 32 * H;
    if (threads > 32 * H) threads = 32 * H;
    scan_fwd_big_kernel<<<B, threads, 2 * H * sizeof(float), cs>>>(F, y0, ys, T, H, bs);
  }
  return (int)cudaGetLastError();
}

int lncde_linear_scan_bwd(void* stream, const float* lie, const float* logsig, const float* ys,
                          const float* g_ys, float* g_lie, float* g_y0, int B, int T, int D, int nb, int bs,
                          int n_basis, void* workspace, size_t workspace_bytes, unsigned flags) {
  using namespace lncde;
  if (!lie || !logsig || !ys || !g_ys || !g_lie || !g_y0 || !workspace) return LNCDE_ERR_NULL;
  int st = check_scan_args(B, T, D, nb, bs, n_basis);
  if (st != LNCDE_OK) return st;
  if (((uintptr_t)workspace & 15) != 0) return LNCDE_ERR_ALIGN;
  if (workspace_bytes < lncde_linear_scan_workspace_bytes(B, T, nb, bs, n_basis)) return LNCDE_ERR_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  float* F = (float*)workspace;  // still holds the forward flows of the SAME (lie, logsig)
  const int H = nb * bs, N = H * bs;
  const int de_v2 = dense_level(flags, N, bs);
  if (lam_path(de_v2, bs, n_basis) && T >= 2) {
    // reverse sweep stores lam_t only; dLie = sum lam_t (x) y_{t-1} (x) logsig_t from the small operands
    const size_t flows = (flows_floats(B, T, nb, bs) + 3) & ~(size_t)3;
    const size_t part_n = ((size_t)N * n_basis * kScanKsplit + 3) & ~(size_t)3;
    float* part = F + flows;
    float* lam = part + part_n;
    if ((cluster_sweeps(flags, bs) && scan_cluster_dispatch(true, cs, F, g_ys, lam, g_y0, B, T, H, bs)) ||
        scan_bwd_lam_dispatch(cs, F, g_ys, lam, g_y0, B, T, H, bs)) {
      st = (int)cudaGetLastError();
      if (st != 0) return st;
      int ksplit = kScanKsplit;
      if (de_v2 >= 3 && dlie_tcgen05_shape(bs, n_basis)) {
        // tcgen05 (3xTF32 split); its error word is the tail of the flow build's scratch (same shape rule)
        const size_t lam_n = (size_t)B * T * nb * bs;
        float* tc5_scratch = F + flows + part_n + lam_n;
        int* err = reinterpret_cast<int*>(tc5_scratch + flow_tcgen05_scratch_floats((long long)B * T, N, n_basis) - 4);
        st = dlie_tcgen05(cs, lam, ys, logsig, part, B, T, D, H, bs, n_basis, kScanKsplit, err, &ksplit);
        if (st != 0) return st;
      } else {
        if (!dlie_triple_dispatch(cs, lam, ys, logsig, part, B, T, D, H, bs, n_basis, kScanKsplit, de_v2 >= 2))
          return LNCDE_ERR_UNSUPPORTED;
        st = (int)cudaGetLastError();
        if (st != 0) return st;
      }
      reduce_partials_kernel<<<148 * 2, 256, 0, cs>>>(part, g_lie, N * n_basis, ksplit);
      return (int)cudaGetLastError();
    }
  }
  if (!pscan_dispatch(true, cs, F, nullptr, const_cast<float*>(ys), g_ys, g_y0, B, T, H, bs,
                      F + pscan_scratch_offset(B, T, nb, bs, n_basis), flags) &&
      !scan_warp_dispatch(true, cs, F, nullptr, const_cast<float*>(ys), g_ys, g_y0, B, T, H, bs) &&
      !scan_tma_dispatch(true, cs, F, nullptr, const_cast<float*>(ys), g_ys, g_y0, B, T, H, bs)) {
    int threads = H < 32 ? 32 : (H > 1024 ? 1024 : ((H + 31) & ~31));
    scan_bwd_kernel<<<B, threads, 3 * H * sizeof(float), cs>>>(F, ys, g_ys, g_y0, T, H, bs);
  }
  st = (int)cudaGetLastError();
  if (st != 0) return st;
  // dLie[n][l] = sum_m dF[m][n] * logsig[m][1+l]
  const size_t flows = (flows_floats(B, T, nb, bs) + 3) & ~(size_t)3;
  GemmParams g;
  g.ksplit = kScanKsplit;
  g.part = F + flows;
  g.n_tasks = 1;
  GemmTask& t = g.task[0];
  t.M = N; t.N = n_basis; t.out_off = 0; t.part_off = 0; t.ksplit = kScanKsplit;
  t.seg[0] = {F, logsig + 1, (long long)N, (long long)D, (long long)B * T};
  t.seg[1] = {nullptr, nullptr, 0, 0, 0};
  launch_gemm(g, 0, 1, cs);
  st = (int)cudaGetLastError();
  if (st != 0) return st;
  reduce_partials_kernel<<<148 * 2, 256, 0, cs>>>(g.part, g_lie, N * n_basis, kScanKsplit);
  return (int)cudaGetLastError();
}

}  // extern "C"
