// Split-K FP32 "sum of outer products" GEMM shared by the Log-NCDE weight gradient (K3b)
// and the LNCDE bracket-matrix gradient (K2 backward):
//
//   out[m][n] = sum over segments  sum_k  left[k*ldl + m] * right[k*ldr + n]
//
// K (the batch x time axis) is huge, M and N are small/medium: each CTA owns one
// (task, m-tile, n-tile, k-slice), accumulates a (16*TM) x (16*TN) tile in registers and
// writes a partial; a second pass sums the k-slices in a fixed order (deterministic).
#pragma once
#include <cuda_runtime.h>

namespace lncde {

struct GemmSeg {
  const float* left;
  const float* right;
  long long ldl, ldr, count;
};
struct GemmTask {
  GemmSeg seg[2];
  int M, N;
  int ksplit;          // k-slices of this task (all tasks of one launch share it)
  long long part_off;  // float offset of this task's partials: [ksplit][M*N]
  int out_off;         // offset in the output buffer
};
constexpr int kMaxTasks = 16;
struct GemmParams {
  GemmTask task[kMaxTasks];
  int n_tasks, ksplit;
  int m_tiles, n_tiles;  // tiles per task (same for every task of one launch)
  float* part;
};

template <int TM, int TN>
__global__ void __launch_bounds__(256) outer_gemm_kernel(const GemmParams p) {
  constexpr int KT = 16;
  __shared__ float Ls[KT][16 * TM];
  __shared__ float Rs[KT][16 * TN];
  int bid = blockIdx.x;
  const int ks = bid % p.ksplit; bid /= p.ksplit;
  const int nt = bid % p.n_tiles; bid /= p.n_tiles;
  const int mt = bid % p.m_tiles; bid /= p.m_tiles;
  const GemmTask& t = p.task[bid];
  const int m0 = mt * 16 * TM, n0 = nt * 16 * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int sgi = 0; sgi < 2; ++sgi) {
    const GemmSeg& sg = t.seg[sgi];
    if (sg.count == 0) continue;
    const long long per = (sg.count + p.ksplit - 1) / p.ksplit;
    const long long k_lo = per * ks, k_hi = min(sg.count, k_lo + per);
    for (long long k0 = k_lo; k0 < k_hi; k0 += KT) {
      const int kn = (int)min((long long)KT, k_hi - k0);
      for (int e = threadIdx.x; e < KT * 16 * TM; e += 256) {
        const int kk = e / (16 * TM), m = e - kk * 16 * TM;
        Ls[kk][m] = (kk < kn && m0 + m < t.M) ? sg.left[(k0 + kk) * sg.ldl + m0 + m] : 0.f;
      }
      for (int e = threadIdx.x; e < KT * 16 * TN; e += 256) {
        const int kk = e / (16 * TN), n = e - kk * 16 * TN;
        Rs[kk][n] = (kk < kn && n0 + n < t.N) ? sg.right[(k0 + kk) * sg.ldr + n0 + n] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        float a[TM], bb[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = Ls[kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bb[j] = Rs[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* out = p.part + t.part_off + (size_t)ks * t.M * t.N;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < t.M && n < t.N) out[(size_t)m * t.N + n] = acc[i][j];
    }
}

static inline int round_tile(int v) {  // ceil(v/16) rounded up to 1,2,4,8
  int t = (v + 15) / 16;
  if (t <= 1) return 1;
  if (t <= 2) return 2;
  if (t <= 4) return 4;
  return 8;
}

template <int TM>
static inline void launch_gemm_tn(int tn, const GemmParams& sub, int grid, cudaStream_t st) {
  switch (tn) {
    case 1: outer_gemm_kernel<TM, 1><<<grid, 256, 0, st>>>(sub); break;
    case 2: outer_gemm_kernel<TM, 2><<<grid, 256, 0, st>>>(sub); break;
    case 4: outer_gemm_kernel<TM, 4><<<grid, 256, 0, st>>>(sub); break;
    default: outer_gemm_kernel<TM, 8><<<grid, 256, 0, st>>>(sub); break;
  }
}
// Launch tasks [first, first+count) of g; all must share M, N.
static inline void launch_gemm(const GemmParams& g, int first, int count, cudaStream_t st) {
  GemmParams sub = g;
  sub.n_tasks = count;
  for (int i = 0; i < count; ++i) sub.task[i] = g.task[first + i];
  const int M = sub.task[0].M, N = sub.task[0].N;
  sub.ksplit = sub.task[0].ksplit;
  const int tm = round_tile(M), tn = round_tile(N);
  sub.m_tiles = (M + 16 * tm - 1) / (16 * tm);
  sub.n_tiles = (N + 16 * tn - 1) / (16 * tn);
  const int grid = count * sub.m_tiles * sub.n_tiles * sub.ksplit;
  switch (tm) {
    case 1: launch_gemm_tn<1>(tn, sub, grid, st); break;
    case 2: launch_gemm_tn<2>(tn, sub, grid, st); break;
    case 4: launch_gemm_tn<4>(tn, sub, grid, st); break;
    default: launch_gemm_tn<8>(tn, sub, grid, st); break;
  }
}

}  // namespace lncde
