// Shared device code of the Log-NCDE solve kernels (K3 forward, K3b backward sweep).
//
// One CTA integrates one sample.  The Log-ODE field of models/LogNeuralCDEs.py:113-138
// is evaluated in the restructured but algebraically identical form
//
//   G(y) = sum_i l_i f_i(y) + sum_j [ Df(y) . w_j ]_{block j},   w_j = sum_i Lam_ij f_i(y)
//
// with f = reshape(MLP(y), (C, H)) and Lam the antisymmetric matrix of the level-2
// Hall coordinates (pairs (i,j), i<j, lexicographic = HallSet order).  Each of the C
// tangents therefore needs only block j of the last layer: 1 primal + C "thin" tangent
// passes instead of the reference's 1 + C full MLP passes.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lncde {

constexpr int kMaxLayers = 5;   // n_hidden <= 4
constexpr int kSolveThreads = 512;

struct MlpDesc {
  int H, C, width, n_hidden, n_layers, depth;
  int in_dim[kMaxLayers], rows[kMaxLayers];
  int w_off[kMaxLayers], b_off[kMaxLayers];  // offsets in the flat parameter buffer
  int stride[kMaxLayers];                     // row stride of the weight copy the kernel reads
  int sm_off[kMaxLayers];                     // offset of the layer in the smem weight region, -1: global
  int kq_log2[kMaxLayers];                    // lanes cooperating on one output row
  int gvec[kMaxLayers];                       // layer read from global memory with coalesced 128-bit rows
  int n_params;
  int weights_in_smem;
  int smem_weight_floats;
};

// per-stage intermediates in shared memory (floats)
struct StageLayout {
  int a, s, z;       // [n_hidden][width]
  int o, w, ul;      // [C*H]
  int u, p;          // [n_hidden][C][width]
  int yin;           // [H]
  int total;
};

__host__ __device__ inline StageLayout make_stage_layout(const MlpDesc& d) {
  StageLayout L;
  int off = 0;
  const int hw = d.n_hidden * d.width, ch = d.C * d.H, cw = d.n_hidden * d.C * d.width;
  L.a = off; off += hw;
  L.s = off; off += hw;
  L.z = off; off += hw;
  L.o = off; off += ch;
  L.w = off; off += ch;
  L.ul = off; off += ch;
  L.u = off; off += cw;
  L.p = off; off += cw;
  L.yin = off; off += d.H;
  L.total = (off + 3) & ~3;
  return L;
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

// out[row] for a batch of NB input vectors; KQ = 2^kq_log2 lanes share one row.
// emit(row, acc[NB]) is called by the lane with q == 0.
template <int NB, typename Emit>
__device__ __forceinline__ void matvec_rows(const float* __restrict__ W, int stride, int rows, int in_dim,
                                            int kq_log2, const float* __restrict__ in, int in_stride,
                                            Emit&& emit) {
  const int KQ = 1 << kq_log2;
  const int items = rows << kq_log2;
  for (int base = 0; base < items; base += kSolveThreads) {
    const int item = base + threadIdx.x;
    const int row = item >> kq_log2, q = item & (KQ - 1);
    const bool valid = row < rows;
    float acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = 0.f;
    if (valid) {
      const float* wr = W + (size_t)row * stride;
      for (int c = q; c < in_dim; c += KQ) {
        const float wv = wr[c];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = fmaf(wv, in[j * in_stride + c], acc[j]);
      }
    }
    for (int m = KQ >> 1; m >= 1; m >>= 1) {
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], m);
    }
    if (valid && q == 0) emit(row, acc);
  }
}

// last layer, tangent j uses only block j:  emit(j, h, dot(W[j*H+h,:], in[j]))
template <typename Emit>
__device__ __forceinline__ void matvec_rows_blocked(const float* __restrict__ W, int stride, int C, int H,
                                                    int in_dim, int kq_log2, const float* __restrict__ in,
                                                    int in_stride, Emit&& emit) {
  const int KQ = 1 << kq_log2;
  const int rows = C * H;
  const int items = rows << kq_log2;
  for (int base = 0; base < items; base += kSolveThreads) {
    const int item = base + threadIdx.x;
    const int row = item >> kq_log2, q = item & (KQ - 1);
    const bool valid = row < rows;
    float acc = 0.f;
    int j = 0;
    if (valid) {
      j = row / H;
      const float* wr = W + (size_t)row * stride;
      const float* v = in + j * in_stride;
      for (int c = q; c < in_dim; c += KQ) acc = fmaf(wr[c], v[c], acc);
    }
    for (int m = KQ >> 1; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (valid && q == 0) emit(j, row - j * H, acc);
  }
}


__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Global-memory weights (layer not resident in shared memory): one warp per output row, lanes
// across the columns with coalesced 128-bit loads, butterfly reduction.  Requires in_dim % 4 == 0
// and 16-byte aligned rows (checked by the host when it picks this path).
template <int NB, typename Emit>
__device__ __forceinline__ void matvec_rows_global(const float* __restrict__ W, int rows, int in_dim,
                                                   const float* __restrict__ in, int in_stride, Emit&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const int nq = in_dim >> 2;
  for (int row = warp; row < rows; row += nw) {
    float acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = 0.f;
    const float* wr = W + (size_t)row * in_dim;
    for (int q = lane; q < nq; q += 32) {
      const float4 w4 = ldg4(wr + 4 * q);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float4 x4 = *reinterpret_cast<const float4*>(in + j * in_stride + 4 * q);
        acc[j] = fmaf(w4.x, x4.x, acc[j]);
        acc[j] = fmaf(w4.y, x4.y, acc[j]);
        acc[j] = fmaf(w4.z, x4.z, acc[j]);
        acc[j] = fmaf(w4.w, x4.w, acc[j]);
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], m);
    }
    if (lane == 0) emit(row, acc);
  }
}

// 4 rows per warp at a time, each row against its own block's input vector
template <typename Emit>
__device__ __forceinline__ void matvec_rows_blocked_global(const float* __restrict__ W, int C, int H, int in_dim,
                                                           const float* __restrict__ in, int in_stride, Emit&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const int nq = in_dim >> 2, rows = C * H;
  for (int r0 = warp * 4; r0 < rows; r0 += nw * 4) {  // H % 4 == 0: the 4 rows share one block
    const int j = r0 / H;
    const float* v = in + j * in_stride;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = lane; q < nq; q += 32) {
      const float4 x4 = *reinterpret_cast<const float4*>(v + 4 * q);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 w4 = ldg4(W + (size_t)(r0 + r) * in_dim + 4 * q);
        acc[r] = fmaf(w4.x, x4.x, acc[r]);
        acc[r] = fmaf(w4.y, x4.y, acc[r]);
        acc[r] = fmaf(w4.z, x4.z, acc[r]);
        acc[r] = fmaf(w4.w, x4.w, acc[r]);
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], m);
    }
    if (lane < 4) emit(j, r0 + lane - j * H, lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]);
  }
}

// primal last layer from global memory: 4 rows per warp at a time, one shared input vector
template <typename Emit>
__device__ __forceinline__ void matvec_rows4_global(const float* __restrict__ W, int rows, int in_dim,
                                                    const float* __restrict__ in, Emit&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const int nq = in_dim >> 2;
  for (int r0 = warp * 4; r0 < rows; r0 += nw * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = lane; q < nq; q += 32) {
      const float4 x4 = *reinterpret_cast<const float4*>(in + 4 * q);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r0 + r < rows) {
          const float4 w4 = ldg4(W + (size_t)(r0 + r) * in_dim + 4 * q);
          acc[r] = fmaf(w4.x, x4.x, acc[r]);
          acc[r] = fmaf(w4.y, x4.y, acc[r]);
          acc[r] = fmaf(w4.z, x4.z, acc[r]);
          acc[r] = fmaf(w4.w, x4.w, acc[r]);
        }
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], m);
    }
    if (lane < 4 && r0 + lane < rows) {
      float one[1] = {lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]};
      emit(r0 + lane, one);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// "Streamed" forms of the global-weight mat-vecs (template flag PF of the generic kernels, tuning "k3_stream").
// The functions above wait for every group of weight rows before they ask for the next one: per warp one round
// trip to L2 per 4 rows (and with in_dim <= 64 half of the lanes issue nothing), so a 896 x 64 layer costs 14
// dependent L2 latencies.  Here a warp covers 32 / LPR groups of 4 rows per pass (LPR = lanes a row needs) and
// the loads of pass k + 1 are issued before pass k is reduced.  Same partial sums in the same order (idle lanes
// contributed exact zeros to the butterflies above): bit-identical results.
__device__ __forceinline__ int lanes_per_row(int nq) {
  int l = 1;
  while (l < nq && l < 32) l <<= 1;
  return l;
}

// rows of W [rows][in_dim] against one vector (BLOCKED: the 4-row group of block j = r0 / H against in + j*in_stride).
// emit(row, value) is called by one lane per row.
template <bool BLOCKED, typename Emit>
__device__ __forceinline__ void matvec_rows4_stream(const float* __restrict__ W, int rows, int H, int in_dim,
                                                    const float* __restrict__ in, int in_stride, Emit&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const int nq = in_dim >> 2, nch = (nq + 31) >> 5;
  const int LPR = lanes_per_row(nq), GPW = 32 / LPR;  // groups of 4 rows a warp covers per pass
  const int sub = lane / LPR, ql = lane - sub * LPR;
  const int ngroups = (rows + 3) >> 2, stride_g = nw * GPW;
  auto load = [&](int g, int ch, float4 (&w)[4]) {
    const int q = ql + 32 * ch, r0 = 4 * g;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      w[r] = (g < ngroups && q < nq && r0 + r < rows) ? ldg4(W + (size_t)(r0 + r) * in_dim + 4 * q)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 cur[4], nxt[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) nxt[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  int g = warp * GPW + sub, ch = 0;  // g0 = the warp's first group of the pass: loop control is warp-uniform
  load(g, 0, cur);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int g0 = warp * GPW; g0 < ngroups; ) {
    int g2 = g, ch2 = ch + 1, g02 = g0;
    if (ch2 == nch) { ch2 = 0; g2 = g + stride_g; g02 = g0 + stride_g; }
    if (g02 < ngroups) load(g2, ch2, nxt);
    const int q = ql + 32 * ch;
    if (g < ngroups && q < nq) {
      const float* v = BLOCKED ? in + ((4 * g) / H) * in_stride : in;
      const float4 x4 = *reinterpret_cast<const float4*>(v + 4 * q);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r] = fmaf(cur[r].x, x4.x, acc[r]);
        acc[r] = fmaf(cur[r].y, x4.y, acc[r]);
        acc[r] = fmaf(cur[r].z, x4.z, acc[r]);
        acc[r] = fmaf(cur[r].w, x4.w, acc[r]);
      }
    }
    if (ch == nch - 1) {
      for (int m = LPR >> 1; m >= 1; m >>= 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], m);
      }
      if (g < ngroups && ql < 4 && 4 * g + ql < rows)
        emit(4 * g + ql, ql == 0 ? acc[0] : ql == 1 ? acc[1] : ql == 2 ? acc[2] : acc[3]);
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = 0.f;
    }
    g = g2; ch = ch2; g0 = g02;
#pragma unroll
    for (int r = 0; r < 4; ++r) cur[r] = nxt[r];
  }
}

// one row per LPR lanes against NB vectors (tangent passes through a hidden layer held in global memory)
template <int NB, typename Emit>
__device__ __forceinline__ void matvec_rows_stream(const float* __restrict__ W, int rows, int in_dim,
                                                   const float* __restrict__ in, int in_stride, Emit&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const int nq = in_dim >> 2, nch = (nq + 31) >> 5;
  const int LPR = lanes_per_row(nq), RPW = 32 / LPR;
  const int sub = lane / LPR, ql = lane - sub * LPR;
  const int stride_r = nw * RPW;
  auto load = [&](int row, int ch) {
    const int q = ql + 32 * ch;
    return (row < rows && q < nq) ? ldg4(W + (size_t)row * in_dim + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  int row = warp * RPW + sub, ch = 0;
  float4 cur = load(row, 0), nxt = make_float4(0.f, 0.f, 0.f, 0.f);
  float acc[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j] = 0.f;
  for (int r0 = warp * RPW; r0 < rows; ) {
    int row2 = row, ch2 = ch + 1, r02 = r0;
    if (ch2 == nch) { ch2 = 0; row2 = row + stride_r; r02 = r0 + stride_r; }
    if (r02 < rows) nxt = load(row2, ch2);
    const int q = ql + 32 * ch;
    if (row < rows && q < nq) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float4 x4 = *reinterpret_cast<const float4*>(in + j * in_stride + 4 * q);
        acc[j] = fmaf(cur.x, x4.x, acc[j]);
        acc[j] = fmaf(cur.y, x4.y, acc[j]);
        acc[j] = fmaf(cur.z, x4.z, acc[j]);
        acc[j] = fmaf(cur.w, x4.w, acc[j]);
      }
    }
    if (ch == nch - 1) {
      for (int m = LPR >> 1; m >= 1; m >>= 1) {
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], m);
      }
      if (row < rows && ql == 0) emit(row, acc);
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[j] = 0.f;
    }
    row = row2; ch = ch2; r0 = r02;
    cur = nxt;
  }
}

// Transposed product with W in global memory (row-major [*][in_dim], in_dim % 4 == 0, 16-byte aligned rows):
// out[j*out_stride + c] = sum_r W[(row0_j + r)*in_dim + c] * v[j*v_stride + r], row0_j = j*nrows if BLOCKED else 0.
// Lanes own 4 columns (128-bit loads, a warp reads 32 / LPR whole rows per pass, several passes in flight); the rows
// of a warp are combined by shuffles, the warps through `scratch` ((kSolveThreads / 32) * NB * in_dim floats).
// Ends with the result visible to all threads.  Same mathematics as matvec_cols, different association order.
template <int NB, bool BLOCKED>
__device__ __forceinline__ void matvec_cols_stream(const float* __restrict__ W, int nrows, int in_dim,
                                                   const float* __restrict__ v, int v_stride, float* out,
                                                   int out_stride, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSolveThreads >> 5;
  const int nq = in_dim >> 2, nch = (nq + 31) >> 5;
  const int LPR = lanes_per_row(nq), RPW = 32 / LPR;
  const int sub = lane / LPR, ql = lane - sub * LPR;
  for (int ch = 0; ch < nch; ++ch) {
    const int q = ql + 32 * ch;
    float4 acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
      if (BLOCKED) {
#pragma unroll 2
        for (int r = warp * RPW + sub; r < nrows; r += nw * RPW) {
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const float4 w4 = ldg4(W + (size_t)(j * nrows + r) * in_dim + 4 * q);
            const float vv = v[j * v_stride + r];
            acc[j].x = fmaf(w4.x, vv, acc[j].x); acc[j].y = fmaf(w4.y, vv, acc[j].y);
            acc[j].z = fmaf(w4.z, vv, acc[j].z); acc[j].w = fmaf(w4.w, vv, acc[j].w);
          }
        }
      } else {
#pragma unroll 8
        for (int r = warp * RPW + sub; r < nrows; r += nw * RPW) {
          const float4 w4 = ldg4(W + (size_t)r * in_dim + 4 * q);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const float vv = v[j * v_stride + r];
            acc[j].x = fmaf(w4.x, vv, acc[j].x); acc[j].y = fmaf(w4.y, vv, acc[j].y);
            acc[j].z = fmaf(w4.z, vv, acc[j].z); acc[j].w = fmaf(w4.w, vv, acc[j].w);
          }
        }
      }
    }
    for (int m = LPR; m < 32; m <<= 1) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, m);
        acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, m);
        acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, m);
        acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, m);
      }
    }
    if (sub == 0 && q < nq) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float* dst = scratch + (size_t)(warp * NB + j) * in_dim + 4 * q;
        dst[0] = acc[j].x; dst[1] = acc[j].y; dst[2] = acc[j].z; dst[3] = acc[j].w;
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NB * in_dim; t += kSolveThreads) {
    const int j = t / in_dim, cc = t - j * in_dim;
    float sum = 0.f;
    for (int w = 0; w < nw; ++w) sum += scratch[(size_t)(w * NB + j) * in_dim + cc];
    out[j * out_stride + cc] = sum;
  }
  __syncthreads();
}

// Transposed product.  out[j*out_stride + c] = sum_r W[(row0_j + r)*stride + c] * v[j*v_stride + r]
// row0_j = j*nrows if BLOCKED else 0.  `scratch` needs kSolveThreads*NB floats.  Ends with the
// result visible to all threads (contains __syncthreads).
template <int NB, bool BLOCKED>
__device__ __forceinline__ void matvec_cols(const float* __restrict__ W, int stride, int nrows, int in_dim,
                                            const float* __restrict__ v, int v_stride, float* out,
                                            int out_stride, float* scratch) {
  // column-parallel: CP = in_dim rounded up to a multiple of 32 lanes, RQ row groups
  const int CP = (in_dim + 31) & ~31;
  const int RQ = max(1, kSolveThreads / CP);
  const int c = threadIdx.x % CP, rq = threadIdx.x / CP;
  float acc[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j] = 0.f;
  if (rq < RQ && c < in_dim) {
    for (int r = rq; r < nrows; r += RQ) {
      if (BLOCKED) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
          acc[j] = fmaf(W[(size_t)(j * nrows + r) * stride + c], v[j * v_stride + r], acc[j]);
      } else {
        const float wv = W[(size_t)r * stride + c];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = fmaf(wv, v[j * v_stride + r], acc[j]);
      }
    }
  }
  if (rq < RQ && c < in_dim) {
#pragma unroll
    for (int j = 0; j < NB; ++j) scratch[(rq * NB + j) * in_dim + c] = acc[j];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NB * in_dim; t += kSolveThreads) {
    const int j = t / in_dim, cc = t - j * in_dim;
    float sum = 0.f;
    for (int g = 0; g < RQ; ++g) sum += scratch[(g * NB + j) * in_dim + cc];
    out[j * out_stride + cc] = sum;
  }
  __syncthreads();
}

struct StageCtx {
  const MlpDesc* d;
  const float* params;     // flat parameter buffer in global memory (biases are read from here)
  const float* wbase[kMaxLayers];  // weight matrix the kernel reads (smem copy or global)
  float* G;                // [H] field value (without the stage scale)
  float* lam;              // [C*C] antisymmetric level-2 matrix, [C] level-1 at lam1
  float* lam1;
};

// Load level-1 / level-2 coordinates of a logsig row into shared memory.
template <int C>
__device__ __forceinline__ void load_logsig_row(const StageCtx& cx, const float* __restrict__ row) {
  const int depth = cx.d->depth;
  if (threadIdx.x < C) cx.lam1[threadIdx.x] = row[1 + threadIdx.x];
  if (depth == 2 && threadIdx.x < C * C) {
    const int i = threadIdx.x / C, j = threadIdx.x - i * C;
    float v = 0.f;
    if (i != j) {
      const int a = i < j ? i : j, b = i < j ? j : i;
      // index of pair (a,b), a<b, lexicographic
      const int p = a * C - (a * (a + 1)) / 2 + (b - a - 1);
      v = row[1 + C + p];
      if (i > j) v = -v;
    }
    cx.lam[threadIdx.x] = v;
  }
}

// Evaluate the field at buf.yin.  On exit G[h] holds the (unscaled) field and all
// intermediates needed by the reverse sweep are in `buf`.
template <int C, bool PF = false>
__device__ __noinline__ void stage_forward(const StageCtx& cx, float* buf, const StageLayout& L,
                                              const float* __restrict__ ls_row) {
  const MlpDesc& d = *cx.d;
  const int H = d.H, Wd = d.width, NL = d.n_hidden;
  load_logsig_row<C>(cx, ls_row);
  // ---- primal pass --------------------------------------------------------
  const float* in = buf + L.yin;
  for (int l = 0; l < NL; ++l) {
    const float* bias = cx.params + d.b_off[l];
    float* a = buf + L.a + l * Wd;
    float* s = buf + L.s + l * Wd;
    float* z = buf + L.z + l * Wd;
    auto fin = [&](int row, const float* acc) {
      const float zz = acc[0] + bias[row];
      const float sg = sigmoidf_(zz);
      z[row] = zz;
      a[row] = zz * sg;
      s[row] = sg * (1.f + zz * (1.f - sg));
    };
    if (d.gvec[l]) {
      if constexpr (PF)
        matvec_rows4_stream<false>(cx.wbase[l], d.rows[l], 1, d.in_dim[l], in, 0, [&](int row, float v) { fin(row, &v); });
      else
        matvec_rows4_global(cx.wbase[l], d.rows[l], d.in_dim[l], in, fin);
    } else matvec_rows<1>(cx.wbase[l], d.stride[l], d.rows[l], d.in_dim[l], d.kq_log2[l], in, 0, fin);
    __syncthreads();
    in = a;
  }
  {
    const float* bias = cx.params + d.b_off[NL];
    float* o = buf + L.o;
    auto fin = [&](int row, const float* acc) { o[row] = tanhf(acc[0] + bias[row]); };
    if (d.gvec[NL]) {
      if constexpr (PF)
        matvec_rows4_stream<false>(cx.wbase[NL], d.rows[NL], 1, d.in_dim[NL], in, 0, [&](int row, float v) { fin(row, &v); });
      else
        matvec_rows4_global(cx.wbase[NL], d.rows[NL], d.in_dim[NL], in, fin);
    } else matvec_rows<1>(cx.wbase[NL], d.stride[NL], d.rows[NL], d.in_dim[NL], d.kq_log2[NL], in, 0, fin);
  }
  __syncthreads();
  // ---- G = sum_i l_i o_i ; w_j = sum_i Lam_ij o_i ---------------------------
  {
    const float* o = buf + L.o;
    float* w = buf + L.w;
    for (int h = threadIdx.x; h < H; h += kSolveThreads) {
      float oi[C];
      float g = 0.f;
#pragma unroll
      for (int i = 0; i < C; ++i) {
        oi[i] = o[i * H + h];
        g = fmaf(cx.lam1[i], oi[i], g);
      }
      cx.G[h] = g;
      if (d.depth == 2) {
#pragma unroll
        for (int j = 0; j < C; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < C; ++i) acc = fmaf(cx.lam[i * C + j], oi[i], acc);
          w[j * H + h] = acc;
        }
      }
    }
  }
  __syncthreads();
  if (d.depth != 2) return;
  // ---- tangent pass: C tangents through the hidden layers -------------------
  const float* tin = buf + L.w;
  int tin_stride = H;
  for (int l = 0; l < NL; ++l) {
    const float* s = buf + L.s + l * Wd;
    float* u = buf + L.u + l * C * Wd;
    float* p = buf + L.p + l * C * Wd;
    auto fin = [&](int row, const float* acc) {
      const float sv = s[row];
#pragma unroll
      for (int j = 0; j < C; ++j) {
        u[j * Wd + row] = acc[j];
        p[j * Wd + row] = sv * acc[j];
      }
    };
    if (d.gvec[l]) {
      if constexpr (PF) matvec_rows_stream<C>(cx.wbase[l], d.rows[l], d.in_dim[l], tin, tin_stride, fin);
      else matvec_rows_global<C>(cx.wbase[l], d.rows[l], d.in_dim[l], tin, tin_stride, fin);
    } else matvec_rows<C>(cx.wbase[l], d.stride[l], d.rows[l], d.in_dim[l], d.kq_log2[l], tin, tin_stride, fin);
    __syncthreads();
    tin = p;
    tin_stride = Wd;
  }
  {
    float* ul = buf + L.ul;
    auto fin = [&](int j, int h, float acc) { ul[j * H + h] = acc; };
    if (d.gvec[NL]) {
      if constexpr (PF)
        matvec_rows4_stream<true>(cx.wbase[NL], C * H, H, d.in_dim[NL], tin, tin_stride, [&](int row, float v) {
          const int j = row / H;
          fin(j, row - j * H, v);
        });
      else
        matvec_rows_blocked_global(cx.wbase[NL], C, H, d.in_dim[NL], tin, tin_stride, fin);
    } else matvec_rows_blocked(cx.wbase[NL], d.stride[NL], C, H, d.in_dim[NL], d.kq_log2[NL], tin, tin_stride, fin);
  }
  __syncthreads();
  {
    const float* o = buf + L.o;
    const float* ul = buf + L.ul;
    for (int h = threadIdx.x; h < H; h += kSolveThreads) {
      float g = cx.G[h];
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float ov = o[j * H + h];
        g = fmaf(1.f - ov * ov, ul[j * H + h], g);
      }
      cx.G[h] = g;
    }
  }
  __syncthreads();
}

// Copy the MLP weights into shared memory with the padded strides of `d`.
__device__ __forceinline__ void stage_weights(const MlpDesc& d, const float* __restrict__ params,
                                              float* smem_w, const float** wbase) {
  for (int l = 0; l < d.n_layers; ++l) {
    if (d.sm_off[l] >= 0) {
      float* dst = smem_w + d.sm_off[l];
      const float* src = params + d.w_off[l];
      const int n = d.rows[l] * d.in_dim[l];
      for (int t = threadIdx.x; t < n; t += kSolveThreads) {
        const int r = t / d.in_dim[l], c = t - r * d.in_dim[l];
        dst[r * d.stride[l] + c] = src[t];
      }
      wbase[l] = dst;
    } else {
      wbase[l] = params + d.w_off[l];
    }
  }
}

// host helper
int make_mlp_desc(MlpDesc* d, int H, int C, int width, int n_hidden, int depth, size_t other_smem_bytes);

}  // namespace lncde
