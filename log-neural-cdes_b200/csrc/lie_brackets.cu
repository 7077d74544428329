// K2a: the bracket matrices of the structured linear CDE - LogLinearCDE.log_ode, models/LinearNeuralCDEs.py:192-232 -
// and their reverse-mode derivative, as kernels (round 1 left this to batched torch.matmul / cuBLAS on the step).
//
//   A_k = vf[k]                       for the letters k < C
//   A_[u,v] = A_v A_u - A_u A_v       for every Hall basis element [u, v] of degree >= 2 (:225-227), per diagonal block
//   lie[blk, i, j, k] = A_k[blk, i, j]   (:230-232: stacked on the last axis)
//
// Parameters only (no batch axis): 21 brackets x 2 products of 128^3 at depth 2 for the dense model (0.18 GFLOP), 133
// brackets at depth 3 (1.1 GFLOP) - small, but on every step, forward and backward.  The elements are computed level
// by level (a degree-d element only needs lower degrees), one launch per level; the reverse pass walks the levels
// downwards and GATHERS per parent (every (parent, block, tile) is owned by one CTA, which adds the contributions of its
// children in table order), so the gradient is deterministic without atomics:
//   G_v += G_k A_u^T - A_u^T G_k,   G_u += A_v^T G_k - G_k A_v^T      for k = [u, v].
// The Hall tables travel in the kernel parameter block (<= 160 basis elements: width 7 at depth 3 has 140).
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"

namespace lncde {

constexpr int kMaxBasis = 160, kMaxLevels = 6, kBrTile = 32;

struct BracketTables {
  int n_basis, C, nb, bs;
  short left[kMaxBasis], right[kMaxBasis];          // parents of element k (0-based indices into A), -1 for letters
  short level_lo[kMaxLevels + 2];                   // elements of degree d are level_lo[d] .. level_lo[d + 1] - 1 (d >= 1)
  // reverse pass: for degree d, the parents that have children of that degree, and per (d, parent) its children
  short par_lo[kMaxLevels + 2];                     // parents of degree-d children: par[par_lo[d] .. par_lo[d + 1])
  short par[2 * kMaxBasis];
  short child_lo[2 * kMaxBasis + 1];                // children of par[q]: child[child_lo[q] .. child_lo[q + 1])
  short child[2 * kMaxBasis];                       // element index k
  signed char role[2 * kMaxBasis];                  // 0: the parent is u (left) of k, 1: it is v (right)
};

// acc[a][b] += sign * sum_m op(P)[i0 + ty*2 + a][m] * op(Q)[m][j0 + tx*2 + b], op = transpose if t* is set; bs x bs matrices
__device__ __forceinline__ void tile_term(float (&acc)[2][2], const float* __restrict__ P, const float* __restrict__ Q, bool tp,
                                          bool tq, float sign, int bs, int i0, int j0, float (*Xs)[kBrTile + 1],
                                          float (*Ys)[kBrTile + 1]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int m0 = 0; m0 < bs; m0 += kBrTile) {
    for (int e = threadIdx.x; e < kBrTile * kBrTile; e += 256) {
      const int r = e / kBrTile, c = e % kBrTile;
      const int i = i0 + r, m = m0 + c;  // Xs[r][c] = op(P)[i][m]
      Xs[r][c] = (i < bs && m < bs) ? (tp ? P[(size_t)m * bs + i] : P[(size_t)i * bs + m]) : 0.f;
      const int mm = m0 + r, j = j0 + c;  // Ys[r][c] = op(Q)[mm][j]
      Ys[r][c] = (mm < bs && j < bs) ? (tq ? Q[(size_t)j * bs + mm] : Q[(size_t)mm * bs + j]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < kBrTile; ++m) {
      const float x0 = Xs[ty * 2][m], x1 = Xs[ty * 2 + 1][m];
      const float y0 = Ys[m][tx * 2], y1 = Ys[m][tx * 2 + 1];
      acc[0][0] = fmaf(sign * x0, y0, acc[0][0]);
      acc[0][1] = fmaf(sign * x0, y1, acc[0][1]);
      acc[1][0] = fmaf(sign * x1, y0, acc[1][0]);
      acc[1][1] = fmaf(sign * x1, y1, acc[1][1]);
    }
    __syncthreads();
  }
}

// A: [n_basis][nb][bs][bs].  grid (tiles, nb, elements of this level)
__global__ void __launch_bounds__(256) bracket_fwd_kernel(float* __restrict__ A, const BracketTables t, int degree) {
  __shared__ float Xs[kBrTile][kBrTile + 1], Ys[kBrTile][kBrTile + 1];
  const int k = t.level_lo[degree] + blockIdx.z, blk = blockIdx.y;
  const int tiles = (t.bs + kBrTile - 1) / kBrTile;
  const int i0 = (blockIdx.x / tiles) * kBrTile, j0 = (blockIdx.x % tiles) * kBrTile;
  const size_t mat = (size_t)t.bs * t.bs, stride = (size_t)t.nb * mat;
  const float* Au = A + t.left[k] * stride + blk * mat;
  const float* Av = A + t.right[k] * stride + blk * mat;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  tile_term(acc, Av, Au, false, false, 1.f, t.bs, i0, j0, Xs, Ys);   // A_v A_u
  tile_term(acc, Au, Av, false, false, -1.f, t.bs, i0, j0, Xs, Ys);  // - A_u A_v
  float* out = A + k * stride + blk * mat;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty * 2 + a, j = j0 + tx * 2 + b;
      if (i < t.bs && j < t.bs) out[(size_t)i * t.bs + j] = acc[a][b];
    }
}

// G: [n_basis][nb][bs][bs], G_parent += contributions of its degree-`degree` children.  grid (tiles, nb, parents)
__global__ void __launch_bounds__(256) bracket_bwd_kernel(const float* __restrict__ A, float* __restrict__ G,
                                                           const BracketTables t, int degree) {
  __shared__ float Xs[kBrTile][kBrTile + 1], Ys[kBrTile][kBrTile + 1];
  const int q = t.par_lo[degree] + blockIdx.z, p = t.par[q], blk = blockIdx.y;
  const int tiles = (t.bs + kBrTile - 1) / kBrTile;
  const int i0 = (blockIdx.x / tiles) * kBrTile, j0 = (blockIdx.x % tiles) * kBrTile;
  const size_t mat = (size_t)t.bs * t.bs, stride = (size_t)t.nb * mat;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int c = t.child_lo[q]; c < t.child_lo[q + 1]; ++c) {
    const int k = t.child[c];
    const float* Gk = G + k * stride + blk * mat;
    if (t.role[c] == 1) {  // p = v:  G_k A_u^T - A_u^T G_k
      const float* Au = A + t.left[k] * stride + blk * mat;
      tile_term(acc, Gk, Au, false, true, 1.f, t.bs, i0, j0, Xs, Ys);
      tile_term(acc, Au, Gk, true, false, -1.f, t.bs, i0, j0, Xs, Ys);
    } else {               // p = u:  A_v^T G_k - G_k A_v^T
      const float* Av = A + t.right[k] * stride + blk * mat;
      tile_term(acc, Av, Gk, true, false, 1.f, t.bs, i0, j0, Xs, Ys);
      tile_term(acc, Gk, Av, false, true, -1.f, t.bs, i0, j0, Xs, Ys);
    }
  }
  float* out = G + p * stride + blk * mat;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty * 2 + a, j = j0 + tx * 2 + b;
      if (i < t.bs && j < t.bs) out[(size_t)i * t.bs + j] += acc[a][b];
    }
}

// kmajor [n_basis][n] <-> kminor [n][n_basis] (n = nb * bs * bs): the scans read lie[n][l] rows
__global__ void __launch_bounds__(256) basis_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int n,
                                                               int n_basis, int to_kminor) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    // read with the source's contiguous index on tx
    if (to_kminor) {
      const int k = k0 + r, e = n0 + tx;
      tile[r][tx] = (k < n_basis && e < n) ? src[(size_t)k * n + e] : 0.f;
    } else {
      const int e = n0 + r, k = k0 + tx;
      tile[r][tx] = (k < n_basis && e < n) ? src[(size_t)e * n_basis + k] : 0.f;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    if (to_kminor) {
      const int e = n0 + r, k = k0 + tx;
      if (k < n_basis && e < n) dst[(size_t)e * n_basis + k] = tile[tx][r];
    } else {
      const int k = k0 + r, e = n0 + tx;
      if (k < n_basis && e < n) dst[(size_t)k * n + e] = tile[tx][r];
    }
  }
}

// pairs: HallSet.data[1:] as (left, right) element numbers, 1-based, left == 0 for the letters (hall_set.py:60-119)
static int make_tables(BracketTables* t, const int32_t* pairs, int C, int n_basis, int nb, int bs) {
  if (C < 1 || n_basis < C || nb < 1 || bs < 1) return LNCDE_ERR_SHAPE;
  if (n_basis > kMaxBasis) return LNCDE_ERR_UNSUPPORTED;
  t->n_basis = n_basis; t->C = C; t->nb = nb; t->bs = bs;
  int deg[kMaxBasis];
  int max_deg = 1;
  for (int k = 0; k < n_basis; ++k) {
    const int l = pairs[2 * k], r = pairs[2 * k + 1];
    if (l == 0) {
      if (k >= C || r != k + 1) return LNCDE_ERR_SHAPE;  // the letters come first, in order
      deg[k] = 1;
      t->left[k] = t->right[k] = -1;
    } else {
      if (l < 1 || r < 1 || l > k || r > k) return LNCDE_ERR_SHAPE;  // parents precede their children
      deg[k] = deg[l - 1] + deg[r - 1];
      t->left[k] = (short)(l - 1);
      t->right[k] = (short)(r - 1);
    }
    if (k > 0 && deg[k] < deg[k - 1]) return LNCDE_ERR_SHAPE;  // Hall order is by degree
    if (deg[k] > kMaxLevels) return LNCDE_ERR_UNSUPPORTED;
    if (deg[k] > max_deg) max_deg = deg[k];
  }
  for (int d = 0; d <= kMaxLevels + 1; ++d) t->level_lo[d] = (short)n_basis;
  for (int k = n_basis - 1; k >= 0; --k) t->level_lo[deg[k]] = (short)k;
  t->level_lo[0] = 0;
  for (int d = kMaxLevels; d >= 1; --d)
    if (t->level_lo[d] > t->level_lo[d + 1]) t->level_lo[d] = t->level_lo[d + 1];
  // gather tables of the reverse pass
  int np = 0, nc = 0;
  for (int d = 0; d <= kMaxLevels + 1; ++d) t->par_lo[d] = 0;
  for (int d = 2; d <= max_deg; ++d) {
    t->par_lo[d] = (short)np;
    for (int p = 0; p < n_basis; ++p) {
      const int c0 = nc;
      for (int k = t->level_lo[d]; k < t->level_lo[d + 1]; ++k) {
        if (t->left[k] == p) { t->child[nc] = (short)k; t->role[nc++] = 0; }
        if (t->right[k] == p) { t->child[nc] = (short)k; t->role[nc++] = 1; }
      }
      if (nc > c0) {
        t->par[np] = (short)p;
        t->child_lo[np++] = (short)c0;
      }
    }
  }
  for (int d = max_deg + 1; d <= kMaxLevels + 1; ++d) t->par_lo[d] = (short)np;
  t->child_lo[np] = (short)nc;
  return max_deg;
}

}  // namespace lncde

extern "C" {

size_t lncde_lie_brackets_workspace_bytes(int n_basis, int nb, int bs) {
  if (n_basis < 1 || nb < 1 || bs < 1) return 0;
  return (size_t)n_basis * nb * bs * bs * sizeof(float);
}

int lncde_lie_brackets_fwd(void* stream, const float* vf, const int32_t* pairs, float* lie, int C, int n_basis, int nb,
                           int bs, void* workspace, size_t workspace_bytes) {
  using namespace lncde;
  if (!vf || !pairs || !lie || !workspace) return LNCDE_ERR_NULL;
  BracketTables t;
  const int max_deg = make_tables(&t, pairs, C, n_basis, nb, bs);
  if (max_deg < 0) return max_deg;
  if (workspace_bytes < lncde_lie_brackets_workspace_bytes(n_basis, nb, bs)) return LNCDE_ERR_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  float* A = (float*)workspace;
  const size_t n = (size_t)nb * bs * bs;
  int st = (int)cudaMemcpyAsync(A, vf, (size_t)C * n * sizeof(float), cudaMemcpyDeviceToDevice, cs);
  if (st != 0) return st;
  const int tiles = (bs + kBrTile - 1) / kBrTile;
  for (int d = 2; d <= max_deg; ++d) {
    const int cnt = t.level_lo[d + 1] - t.level_lo[d];
    if (cnt <= 0) continue;
    bracket_fwd_kernel<<<dim3(tiles * tiles, nb, cnt), 256, 0, cs>>>(A, t, d);
  }
  basis_transpose_kernel<<<dim3((unsigned)((n + 31) / 32), (n_basis + 31) / 32), 256, 0, cs>>>(A, lie, (int)n, n_basis, 1);
  return (int)cudaGetLastError();
}

/* workspace: the A [n_basis][nb][bs][bs] the forward call left behind; workspace_g: as many bytes again (scratch) */
int lncde_lie_brackets_bwd(void* stream, const float* g_lie, const int32_t* pairs, float* g_vf, int C, int n_basis, int nb,
                           int bs, const void* workspace, void* workspace_g, size_t workspace_bytes) {
  using namespace lncde;
  if (!g_lie || !pairs || !g_vf || !workspace || !workspace_g) return LNCDE_ERR_NULL;
  BracketTables t;
  const int max_deg = make_tables(&t, pairs, C, n_basis, nb, bs);
  if (max_deg < 0) return max_deg;
  if (workspace_bytes < lncde_lie_brackets_workspace_bytes(n_basis, nb, bs)) return LNCDE_ERR_WORKSPACE;
  cudaStream_t cs = (cudaStream_t)stream;
  const float* A = (const float*)workspace;
  float* G = (float*)workspace_g;
  const size_t n = (size_t)nb * bs * bs;
  basis_transpose_kernel<<<dim3((unsigned)((n + 31) / 32), (n_basis + 31) / 32), 256, 0, cs>>>(g_lie, G, (int)n, n_basis, 0);
  const int tiles = (bs + kBrTile - 1) / kBrTile;
  for (int d = max_deg; d >= 2; --d) {
    const int cnt = t.par_lo[d + 1] - t.par_lo[d];
    if (cnt <= 0) continue;
    bracket_bwd_kernel<<<dim3(tiles * tiles, nb, cnt), 256, 0, cs>>>(A, G, t, d);
  }
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
  return (int)cudaMemcpyAsync(g_vf, G, (size_t)C * n * sizeof(float), cudaMemcpyDeviceToDevice, cs);
}

}  // extern "C"
