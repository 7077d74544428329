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

#include <cstdint>

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

// index of the i-th row (column) owned by thread coordinate t: contiguous groups of 4 so that the
// operand reads from shared memory are conflict-free 128-bit (64-bit for T = 2) loads
template <int T>
__device__ __forceinline__ int own_index(int t, int i) {
  return T >= 4 ? (i >> 2) * 64 + t * 4 + (i & 3) : t * T + i;
}

template <int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256) outer_gemm_kernel(const GemmParams p) {
  constexpr int KT = 32, MT = 16 * TM, NT_ = 16 * TN;
  __shared__ __align__(16) float Ls[KT][MT];
  __shared__ __align__(16) float Rs[KT][NT_];
  int bid = blockIdx.x;
  const int ks = bid % p.ksplit; bid /= p.ksplit;
  const int nt = bid % p.n_tiles; bid /= p.n_tiles;
  const int mt = bid % p.m_tiles; bid /= p.m_tiles;
  const GemmTask& t = p.task[bid];
  const int m0 = mt * MT, n0 = nt * NT_;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  // global -> register staging: LV (RV) 128-bit loads per thread cover one K-tile
  constexpr int LV = (KT * MT / 4 + 255) / 256, RV = (KT * NT_ / 4 + 255) / 256;
  float4 lreg[LV], rreg[RV];
  for (int sgi = 0; sgi < 2; ++sgi) {
    const GemmSeg& sg = t.seg[sgi];
    if (sg.count == 0) continue;
    const long long per = (sg.count + p.ksplit - 1) / p.ksplit;
    const long long k_lo = per * ks, k_hi = min(sg.count, k_lo + per);
    auto fetch = [&](long long k0) {
#pragma unroll
      for (int u = 0; u < LV; ++u) {
        const int e = threadIdx.x + u * 256;
        const int kk = e / (MT / 4), m = (e - kk * (MT / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < KT * MT / 4 && k0 + kk < k_hi) {
          const float* src = sg.left + (k0 + kk) * sg.ldl + m0 + m;
          if (VEC) {
            if (m0 + m < t.M) v = *reinterpret_cast<const float4*>(src);
          } else {
            if (m0 + m + 0 < t.M) v.x = src[0];
            if (m0 + m + 1 < t.M) v.y = src[1];
            if (m0 + m + 2 < t.M) v.z = src[2];
            if (m0 + m + 3 < t.M) v.w = src[3];
          }
        }
        lreg[u] = v;
      }
#pragma unroll
      for (int u = 0; u < RV; ++u) {
        const int e = threadIdx.x + u * 256;
        const int kk = e / (NT_ / 4), n = (e - kk * (NT_ / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < KT * NT_ / 4 && k0 + kk < k_hi) {
          const float* src = sg.right + (k0 + kk) * sg.ldr + n0 + n;
          if (VEC) {
            if (n0 + n < t.N) v = *reinterpret_cast<const float4*>(src);
          } else {
            if (n0 + n + 0 < t.N) v.x = src[0];
            if (n0 + n + 1 < t.N) v.y = src[1];
            if (n0 + n + 2 < t.N) v.z = src[2];
            if (n0 + n + 3 < t.N) v.w = src[3];
          }
        }
        rreg[u] = v;
      }
    };
    if (k_lo < k_hi) fetch(k_lo);
    for (long long k0 = k_lo; k0 < k_hi; k0 += KT) {
#pragma unroll
      for (int u = 0; u < LV; ++u) {
        const int e = threadIdx.x + u * 256;
        if (e < KT * MT / 4) *reinterpret_cast<float4*>(&Ls[0][0] + e * 4) = lreg[u];
      }
#pragma unroll
      for (int u = 0; u < RV; ++u) {
        const int e = threadIdx.x + u * 256;
        if (e < KT * NT_ / 4) *reinterpret_cast<float4*>(&Rs[0][0] + e * 4) = rreg[u];
      }
      __syncthreads();
      if (k0 + KT < k_hi) fetch(k0 + KT);  // next tile in flight while this one is consumed
#pragma unroll 8
      for (int kk = 0; kk < KT; ++kk) {
        float a[TM], bb[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = Ls[kk][own_index<TM>(ty, i)];
#pragma unroll
        for (int j = 0; j < TN; ++j) bb[j] = Rs[kk][own_index<TN>(tx, j)];
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
      const int m = m0 + own_index<TM>(ty, i), n = n0 + own_index<TN>(tx, j);
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

template <int TM, bool VEC>
static inline void launch_gemm_tn(int tn, const GemmParams& sub, int grid, cudaStream_t st) {
  switch (tn) {
    case 1: outer_gemm_kernel<TM, 1, VEC><<<grid, 256, 0, st>>>(sub); break;
    case 2: outer_gemm_kernel<TM, 2, VEC><<<grid, 256, 0, st>>>(sub); break;
    case 4: outer_gemm_kernel<TM, 4, VEC><<<grid, 256, 0, st>>>(sub); break;
    default: outer_gemm_kernel<TM, 8, VEC><<<grid, 256, 0, st>>>(sub); break;
  }
}
template <bool VEC>
static inline void launch_gemm_tm(int tm, int tn, const GemmParams& sub, int grid, cudaStream_t st) {
  switch (tm) {
    case 1: launch_gemm_tn<1, VEC>(tn, sub, grid, st); break;
    case 2: launch_gemm_tn<2, VEC>(tn, sub, grid, st); break;
    case 4: launch_gemm_tn<4, VEC>(tn, sub, grid, st); break;
    default: launch_gemm_tn<8, VEC>(tn, sub, grid, st); break;
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
  // 128-bit global loads when every operand row is 16 B aligned and M, N are multiples of 4
  bool vec = (M % 4 == 0) && (N % 4 == 0);
  for (int i = 0; i < count && vec; ++i)
    for (int sg = 0; sg < 2; ++sg) {
      const GemmSeg& g2 = sub.task[i].seg[sg];
      if (g2.count == 0) continue;
      vec = vec && ((uintptr_t)g2.left % 16 == 0) && ((uintptr_t)g2.right % 16 == 0) && (g2.ldl % 4 == 0) &&
            (g2.ldr % 4 == 0);
    }
  if (vec) launch_gemm_tm<true>(tm, tn, sub, grid, st);
  else launch_gemm_tm<false>(tm, tn, sub, grid, st);
}

}  // namespace lncde
