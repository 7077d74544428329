// K1 – batched per-interval truncated log-signature in the Hall basis (sm_100a).
//
// Replaces data_dir/generate_paths.py:22-77 (calc_paths -> hall_basis_logsig ->
// signax signature/log/flatten -> t2l projection).  Two kernels:
//
//  * logsig_levy_kernel<C, DEPTH2>  (C <= 8, depth <= 2): HBM-bound streaming
//    kernel.  A CTA stages a contiguous tile of one sample's path
//    (WT windows x stepsize points x C channels, read once, 128-bit coalesced)
//    into shared memory with an odd window stride (conflict-free), one thread
//    then walks one window keeping the C running coordinates and the C(C-1)/2
//    Levy areas in registers, results are transposed through shared memory and
//    written back as one contiguous, coalesced block.  The Hall projection at
//    depth 2 is the identity on the upper triangle (t2l folds to "take L2[i,j],
//    i<j"), so no table is read.
//  * logsig_tensor_kernel (any C, depth <= 4): one warp per window; truncated
//    signature levels live in shared memory, Chen/Horner update per increment,
//    log and the sparse t2l map (CSR from K0) are folded into the epilogue.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "lncde_b200.h"
#include "ptx.cuh"

namespace lncde {

// ----------------------------------------------------------------------------
// geometry shared by both kernels.  After the reference's zero-prepend the
// padded path has n = L+1 points xt[0] = 0, xt[m] = x[m-1].  Full window k
// (0 <= k < q) holds xt[k*s .. k*s+s-1] with basepoint xt[k*s-1] (0 for k = 0):
// in ORIGINAL rows that is basepoint row k*s-2 and new rows k*s-1 .. k*s+s-2,
// rows < 0 reading as 0.  The remainder window (r = n % s != 0) holds original
// rows L-r-1 .. L-1 (its true predecessor first) and, in compat mode, one more
// leading point: the last row of the previous sample (generate_paths.py:59).
// ----------------------------------------------------------------------------

template <int C, bool DEPTH2>
struct LevyState {
  float base[C];
  float prev[C];   // previous point minus base
  float area[DEPTH2 ? (C * (C - 1)) / 2 + 1 : 1];

  __device__ __forceinline__ void init(const float* b) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      base[c] = b[c];
      prev[c] = 0.f;
    }
    if (DEPTH2) {
#pragma unroll
      for (int p = 0; p < (C * (C - 1)) / 2; ++p) area[p] = 0.f;
    }
  }
  // advance by one point x (absolute coordinates)
  __device__ __forceinline__ void step(const float* x) {
    float cur[C], inc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      cur[c] = x[c] - base[c];
      inc[c] = cur[c] - prev[c];
    }
    if (DEPTH2) {
      int p = 0;
#pragma unroll
      for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = i + 1; j < C; ++j) {
          // 2*dA_ij = a_i * dx_j - a_j * dx_i, a = position before the increment
          area[p] = fmaf(prev[i], inc[j], area[p]);
          area[p] = fmaf(-prev[j], inc[i], area[p]);
          ++p;
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) prev[c] = cur[c];
  }
  // write the D-float row [0 | level1 | level2]
  __device__ __forceinline__ void store(float* out) const {
    out[0] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) out[1 + c] = prev[c];
    if (DEPTH2) {
#pragma unroll
      for (int p = 0; p < (C * (C - 1)) / 2; ++p) out[1 + C + p] = 0.5f * area[p];
    }
  }
};

struct LevyParams {
  const float* data;
  float* out;
  int B, L, s, q, r, K, D, WT, tiles, compat, chunk;
  int stride;         // shared-memory floats per window (s*C, made odd)
  uint32_t sc_magic;  // ceil(2^32 / (s*C))
};

// Stage rows k0*s-2 .. k0*s-2 + nw*s of one sample (nw*s + 1 rows) into shared memory:
// tile_base[0,C) = basepoint row of window k0; window w's s rows at win + w*stride (odd stride).
__device__ __forceinline__ void stage_windows(const LevyParams& p, const int C, const float* __restrict__ xb,
                                              const int k0, const int nw, float* tile_base, float* win) {
  const int sC = p.s * C;
  // ---- stage rows k0*s-2 .. k0*s-2 + nw*s  (nw*s + 1 rows) ------------------
  // smem: [0,C) tile basepoint row; window w's s rows at C + w*stride.
  const long long row0 = (long long)k0 * p.s - 2;  // basepoint row of window k0
  if (threadIdx.x < C) {
    tile_base[threadIdx.x] = row0 >= 0 ? ld_stream(xb + row0 * C + threadIdx.x) : 0.f;
  }
  {
    const long long g_begin = (row0 + 1) * C;               // first float wanted (may be < 0)
    const int n_f = nw * sC;                                 // floats wanted
    const long long g_lo = g_begin < 0 ? 0 : g_begin;        // clamp: row -1 is the zero point
    // zero-fill the part that lies before the sample start (only tile 0: C floats)
    for (int t = threadIdx.x; t < (int)(g_lo - g_begin); t += blockDim.x) win[t] = 0.f;
    const float* src = xb + g_lo;
    const int t_off = (int)(g_lo - g_begin);
    const int n_ld = n_f - t_off;
    // 128-bit body on the 16 B aligned part, scalar head/tail
    const int head = min(n_ld, (int)(((16 - ((uintptr_t)src & 15)) & 15) >> 2));
    const int n_v4 = (n_ld - head) >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src + head);
    // window index once per 128-bit load; the +1 pad per window boundary is branch-free
    const int pad = p.stride - sC;  // 0 or 1
    auto put = [&](int t, float v) {
      const uint32_t w = __umulhi((uint32_t)t, p.sc_magic);
      win[t + (int)w * pad] = v;
    };
    auto put4 = [&](int t, const float4& v) {
      if (sC < 4) {  // a 128-bit load may span several windows: scalar placement
        put(t, v.x); put(t + 1, v.y); put(t + 2, v.z); put(t + 3, v.w);
        return;
      }
      const int w = (int)__umulhi((uint32_t)t, p.sc_magic);
      const int r = t - w * sC;          // position inside window w
      float* dst = win + t + w * pad;
      const int lim = sC - r;            // elements e >= lim belong to window w+1
      dst[0] = v.x;
      dst[1 + (1 >= lim ? pad : 0)] = v.y;
      dst[2 + (2 >= lim ? pad : 0)] = v.z;
      dst[3 + (3 >= lim ? pad : 0)] = v.w;
    };
    if (threadIdx.x < head) put(t_off + threadIdx.x, ld_stream(src + threadIdx.x));
    for (int i = threadIdx.x; i < n_v4; i += blockDim.x * 6) {
      float4 v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u)
        if (i + u * (int)blockDim.x < n_v4) v[u] = ld_stream4(src4 + i + u * blockDim.x);
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int ii = i + u * (int)blockDim.x;
        if (ii < n_v4) put4(t_off + head + 4 * ii, v[u]);
      }
    }
    const int tail0 = head + 4 * n_v4;
    if (threadIdx.x < n_ld - tail0) put(t_off + tail0 + threadIdx.x, ld_stream(src + tail0 + threadIdx.x));
  }
}

template <int C, bool DEPTH2>
__global__ void __launch_bounds__(128) logsig_levy_kernel(const LevyParams p) {
  extern __shared__ float smem[];
  constexpr int D = 1 + C + (DEPTH2 ? (C * (C - 1)) / 2 : 0);
  const int b = blockIdx.x / p.tiles;
  const int tile = blockIdx.x - b * p.tiles;
  const int k0 = tile * p.WT;
  const int nw = min(p.WT, p.q - k0);  // full windows in this tile (may be 0 for K == remainder only)
  const int sC = p.s * C;
  const float* xb = p.data + (size_t)b * p.L * C;

  float* tile_base = smem;
  float* win = smem + C;
  stage_windows(p, C, xb, k0, nw, tile_base, win);
  __syncthreads();

  // ---- one thread per window ------------------------------------------------
  LevyState<C, DEPTH2> st;
  const int w = threadIdx.x;
  if (w < nw) {
    const float* bp = (w == 0) ? tile_base : (win + (w - 1) * p.stride + (p.s - 1) * C);
    float tmp[C];
#pragma unroll
    for (int c = 0; c < C; ++c) tmp[c] = bp[c];
    st.init(tmp);
    const float* rows = win + w * p.stride;
#pragma unroll 4
    for (int l = 0; l < p.s; ++l) {
      float xs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) xs[c] = rows[l * C + c];
      st.step(xs);
    }
  }
  __syncthreads();
  // ---- transpose through shared memory, contiguous coalesced store ---------
  // staged at an offset that makes shared source and global destination 16 B aligned together
  float* dst = p.out + ((size_t)b * p.K + k0) * D;
  const int OUT_SHIFT = (int)(((uintptr_t)dst >> 2) & 3);
  if (w < nw) st.store(smem + OUT_SHIFT + w * D);
  __syncthreads();
  {
    const int n_o = nw * D;
    // 128-bit body: smem offset `oh` floats makes source and destination 16 B aligned together
    const int oh = min(n_o, (int)(((16 - ((uintptr_t)dst & 15)) & 15) >> 2));
    const int n4 = (n_o - oh) >> 2;
    if (threadIdx.x < oh) dst[threadIdx.x] = smem[OUT_SHIFT + threadIdx.x];
    const float4* s4 = reinterpret_cast<const float4*>(smem + OUT_SHIFT + oh);
    float4* d4 = reinterpret_cast<float4*>(dst + oh);
    for (int t = threadIdx.x; t < n4; t += blockDim.x) d4[t] = s4[t];
    const int ot = oh + 4 * n4;
    if (threadIdx.x < n_o - ot) dst[ot + threadIdx.x] = smem[OUT_SHIFT + ot + threadIdx.x];
  }

  // ---- remainder window: handled by the sample's last tile, one thread -----
  if (p.r != 0 && tile == p.tiles - 1 && threadIdx.x == 0) {
    float tmp[C];
    const bool extra = p.compat != 0;
    const float* first = xb + (size_t)(p.L - p.r - 1) * C;  // true predecessor row
    if (extra) {
      const bool has_prev = (b % p.chunk) != 0;
#pragma unroll
      for (int c = 0; c < C; ++c) tmp[c] = has_prev ? xb[-C + c] : 0.f;  // last row of sample b-1
      st.init(tmp);
#pragma unroll
      for (int c = 0; c < C; ++c) tmp[c] = first[c];
      st.step(tmp);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) tmp[c] = first[c];
      st.init(tmp);
    }
    for (int l = 1; l <= p.r; ++l) {
#pragma unroll
      for (int c = 0; c < C; ++c) tmp[c] = first[l * C + c];
      st.step(tmp);
    }
    float row[D];
    st.store(row);
    float* dst = p.out + ((size_t)b * p.K + p.q) * D;
#pragma unroll
    for (int c = 0; c < D; ++c) dst[c] = row[c];
  }
}


// ----------------------------------------------------------------------------
// logsig_levy_tma_kernel<C, DEPTH2>: same closed form and the same LevyState arithmetic as
// logsig_levy_kernel (bit-identical results), with the data movement handed to the TMA unit:
//   * the tile's input slab (32 windows of one sample, contiguous in global memory) arrives in
//     shared memory with ONE 1-D bulk copy (cp.async.bulk + mbarrier) issued by lane 0, instead of
//     ~85 scalar shared-memory stores per window plus their index arithmetic (40 % of the
//     instructions of the kernel above; ncu: issue slots were the co-limiter at 64 % of HBM);
//   * when s*C is a multiple of 4 every window of the tile has the same 16-byte phase SH, so a
//     thread walks its window with 128-bit shared loads at compile-time register offsets
//     (template on SH); window stride s*C/4 chunks: conflict-free when odd, at most 2-/4-way
//     otherwise, which still leaves the shared-memory pipe far below the HBM time of a tile;
//     other strides walk with scalar loads on the same unpadded tile;
//   * the 32 x D result rows leave through one bulk store of the 16-byte aligned body
//     (cp.async.bulk.global.shared::cta) plus <= 3 head / tail floats.
// Single-warp CTAs: the only synchronisation is __syncwarp and the mbarrier wait.
// ----------------------------------------------------------------------------
constexpr int kLevyPad = 16;  // floats in front of the bulk-copied slab: room for the two zero rows of tile 0

template <int C, bool DEPTH2, int SH>
__device__ __forceinline__ void levy_walk_vec(LevyState<C, DEPTH2>& st, const float* slot, int s) {
  // stream position i of this window (basepoint row, then the s rows) is element SH + i of the
  // 16-byte chunks starting at `slot`
  constexpr int NB = (SH + C + 3) / 4;
  float b[4 * NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const float4 v = reinterpret_cast<const float4*>(slot)[j];
    b[4 * j] = v.x; b[4 * j + 1] = v.y; b[4 * j + 2] = v.z; b[4 * j + 3] = v.w;
  }
  st.init(&b[SH]);
  constexpr int R = (C % 4 == 0) ? 1 : (C % 2 == 0) ? 2 : 4;  // rows per group: R*C floats = G whole chunks
  constexpr int G = R * C / 4;
  constexpr int O = (SH + C) % 4, Q0 = (SH + C) / 4;
  constexpr int NV = G + (O ? 1 : 0);
  const float4* q = reinterpret_cast<const float4*>(slot) + Q0;
  const int groups = s / R;  // s*C % 4 == 0 implies s % R == 0
  for (int g = 0; g < groups; ++g, q += G) {
    float v[4 * NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 t = q[j];
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) st.step(&v[O + r * C]);
  }
}

// Geometry of tile t = (sample b, windows k0 .. k0 + nw - 1): floats [start_f, end_f) of the sample - basepoint row
// k0*s - 2 .. last row (k0 + nw)*s - 2, rows < 0 read as 0 - and the 16-byte aligned range [a0, a1) inside it.
struct LevyTile {
  int b, tile, k0, nw, sh0, lead;
  unsigned bytes;
  const float *xb, *a0, *t_lo;
  int n_tail;
};
__device__ __forceinline__ LevyTile levy_tile(const LevyParams& p, int C, long long t) {
  LevyTile g;
  g.b = (int)(t / p.tiles);
  g.tile = (int)(t - (long long)g.b * p.tiles);
  g.k0 = g.tile * 32;
  g.nw = min(32, p.q - g.k0);
  g.xb = p.data + (size_t)g.b * p.L * C;
  const long long start_f = ((long long)g.k0 * p.s - 2) * C;
  const long long end_f = ((long long)(g.k0 + g.nw) * p.s - 1) * C;
  const long long lo_f = start_f < 0 ? 0 : start_f;
  const float* g_lo = g.xb + lo_f;
  const float* g_end = g.xb + end_f;
  g.a0 = reinterpret_cast<const float*>((uintptr_t)g_lo & ~(uintptr_t)15);
  const float* a1 = reinterpret_cast<const float*>((uintptr_t)g_end & ~(uintptr_t)15);
  g.sh0 = (int)(g_lo - g.a0);
  g.bytes = (g.nw > 0 && a1 > g.a0) ? (unsigned)((a1 - g.a0) * 4) : 0u;
  g.t_lo = a1 > g_lo ? a1 : g_lo;                       // <= 3 floats between the 16-byte floor of the end and the end
  g.n_tail = g.nw > 0 ? (int)(g_end - g.t_lo) : 0;
  g.lead = (int)(lo_f - start_f);                       // 2C zero floats in front of tile 0, else 0
  return g;
}
// start the load of a tile into `buf`: one bulk copy (lane 0) + the tail floats
__device__ __forceinline__ void levy_issue(const LevyTile& g, float* buf, unsigned long long* bar, int lane) {
  float* slab = buf + kLevyPad;  // smem image of a0[...]
  if (lane == 0 && g.bytes) {
    mbar_expect_tx(bar, g.bytes);
    tma_load_1d(slab, g.a0, g.bytes, bar);
  }
  if (lane < g.n_tail) slab[(g.t_lo - g.a0) + lane] = ld_stream(g.t_lo + lane);
}

// One single-warp CTA per tile.  (A persistent two-buffer form of this kernel - grid sized to the machine, the bulk
// copy of tile k + 1 in flight while tile k is walked - measured 7 % SLOWER on B200, 0.216 vs 0.201 ms at the roofline
// shape, and was removed in round 2: with ~20 CTAs resident per SM the hardware scheduler already overlaps the phases.)
template <int C, bool DEPTH2>
__global__ void __launch_bounds__(32) logsig_levy_tma_kernel(const LevyParams p, int buf_floats) {
  extern __shared__ __align__(128) float smem[];
  constexpr int D = 1 + C + (DEPTH2 ? (C * (C - 1)) / 2 : 0);
  const int lane = threadIdx.x;
  const int sC = p.s * C;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
  float* bufs = smem + 4;  // 16 bytes in; per buffer: [kLevyPad floats | bulk-copied slab | tail floats]
  const long long n_tiles = (long long)p.B * p.tiles;
  if (lane == 0) {
    mbar_init(bars, 1);
    fence_mbar_init();
  }
  __syncwarp();
  long long t = blockIdx.x;
  if (t >= n_tiles) return;
  levy_issue(levy_tile(p, C, t), bufs, bars, lane);
  {
    float* buf = bufs;
    const LevyTile g = levy_tile(p, C, t);
    LevyState<C, DEPTH2> st;
    if (g.nw > 0) {
      float* slab = buf + kLevyPad;
      __syncwarp();
      if (g.bytes) {
        mbar_wait(bars, 0u);
      }
      if (lane < g.lead) slab[g.sh0 - g.lead + lane] = 0.f;
      __syncwarp();
      if (lane < g.nw) {
        const int base_idx = kLevyPad + g.sh0 - g.lead + lane * sC;  // basepoint row of this lane's window
        if ((sC & 3) == 0) {
          const float* slot = buf + (base_idx & ~3);
          switch (base_idx & 3) {  // warp-uniform: kLevyPad, lane*sC are multiples of 4
            case 0: levy_walk_vec<C, DEPTH2, 0>(st, slot, p.s); break;
            case 1: levy_walk_vec<C, DEPTH2, 1>(st, slot, p.s); break;
            case 2: levy_walk_vec<C, DEPTH2, 2>(st, slot, p.s); break;
            default: levy_walk_vec<C, DEPTH2, 3>(st, slot, p.s); break;
          }
        } else {
          const float* rows = buf + base_idx;
          float tmp[C];
#pragma unroll
          for (int c = 0; c < C; ++c) tmp[c] = rows[c];
          st.init(tmp);
          rows += C;
#pragma unroll 2
          for (int l = 0; l < p.s; ++l) {
#pragma unroll
            for (int c = 0; c < C; ++c) tmp[c] = rows[l * C + c];
            st.step(tmp);
          }
        }
      }
      __syncwarp();  // every lane is done reading the slab: reuse it for the result rows
      float* dst = p.out + ((size_t)g.b * p.K + g.k0) * D;
      const int os = (int)(((uintptr_t)dst >> 2) & 3);  // smem phase = global phase: body 16 B aligned in both
      float* stage = buf + os;
      if (lane < g.nw) st.store(stage + lane * D);
      fence_async_proxy();  // generic-proxy writes above -> visible to the bulk store below
      __syncwarp();
      const int n_o = g.nw * D;
      const int head = min(n_o, (4 - os) & 3);
      const int body = (n_o - head) & ~3;
      if (lane == 0 && body > 0) {
        tma_store_1d(dst + head, stage + head, (unsigned)body * 4);
        tma_store_commit();
      }
      if (lane < head) dst[lane] = stage[lane];
      const int tail0 = head + body;
      if (lane < n_o - tail0) dst[tail0 + lane] = stage[tail0 + lane];
    }

    // ---- remainder window: handled with the sample's last tile, one thread -----
    if (p.r != 0 && g.tile == p.tiles - 1 && lane == 0) {
      float tmp[C];
      const float* xb = g.xb;
      const bool extra = p.compat != 0;
      const float* first = xb + (size_t)(p.L - p.r - 1) * C;  // true predecessor row
      if (extra) {
        const bool has_prev = (g.b % p.chunk) != 0;
#pragma unroll
        for (int c = 0; c < C; ++c) tmp[c] = has_prev ? xb[-C + c] : 0.f;  // last row of sample b-1
        st.init(tmp);
#pragma unroll
        for (int c = 0; c < C; ++c) tmp[c] = first[c];
        st.step(tmp);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) tmp[c] = first[c];
        st.init(tmp);
      }
      for (int l = 1; l <= p.r; ++l) {
#pragma unroll
        for (int c = 0; c < C; ++c) tmp[c] = first[l * C + c];
        st.step(tmp);
      }
      float row[D];
      st.store(row);
      float* dst = p.out + ((size_t)g.b * p.K + p.q) * D;
#pragma unroll
      for (int c = 0; c < D; ++c) dst[c] = row[c];
    }
  }
  if (lane == 0) tma_store_wait_read();  // shared memory must outlive the bulk read (no-op without a pending store)
}

// ----------------------------------------------------------------------------
// depth 3, C <= 8: thread (window, i) keeps "row i" of the truncated signature in registers
// (S1[i], S2[i][.], S3[i][.][.]): the Chen/Horner update of row i needs only the increment, so
// the sweep over the window is communication-free.  log() and the sparse Hall projection (CSR
// from K0) run once per window out of shared memory.
// ----------------------------------------------------------------------------
struct D3Params {
  LevyParams lv;
  const int32_t* rowptr;
  const int32_t* cols;
  const float* vals;
};

template <int C>
__global__ void __launch_bounds__(32 * C) logsig_d3_kernel(const D3Params pp) {
  extern __shared__ float smem[];
  const LevyParams& p = pp.lv;
  constexpr int WT = 32, C2 = C * C, C3 = C * C * C, E = C + C2 + C3;
  const int b = blockIdx.x / p.tiles, tile = blockIdx.x - b * p.tiles;
  const int k0 = tile * WT;
  const int nw = min(WT, p.q - k0);
  const float* xb = p.data + (size_t)b * p.L * C;
  float* ex = smem;                       // [WT][E] : S1 | S2 | L3 per window
  float* tile_base = smem + WT * E;
  float* win = tile_base + C;
  stage_windows(p, C, xb, k0, nw, tile_base, win);
  __syncthreads();
  const int w = threadIdx.x / C, i = threadIdx.x - w * C;
  float s1 = 0.f, S2r[C], S3r[C][C];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    S2r[j] = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) S3r[j][k] = 0.f;
  }
  if (w < nw) {
    const float* bp = (w == 0) ? tile_base : (win + (w - 1) * p.stride + (p.s - 1) * C);
    float base[C], prev[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      base[c] = bp[c];
      prev[c] = 0.f;
    }
    const float base_i = bp[i];
    float prev_i = 0.f;
    const float* rows = win + w * p.stride;
#pragma unroll 2
    for (int l = 0; l < p.s; ++l) {
      float z[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float cur = rows[l * C + c] - base[c];
        z[c] = cur - prev[c];
        prev[c] = cur;
      }
      const float cur_i = rows[l * C + i] - base_i;
      const float zi = cur_i - prev_i;
      prev_i = cur_i;
      const float a3 = fmaf(zi, 1.f / 3.f, s1);
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float bj = fmaf(a3 * 0.5f, z[j], S2r[j]);
#pragma unroll
        for (int k = 0; k < C; ++k) S3r[j][k] = fmaf(bj, z[k], S3r[j][k]);
      }
      const float a2 = fmaf(zi, 0.5f, s1);
#pragma unroll
      for (int j = 0; j < C; ++j) S2r[j] = fmaf(a2, z[j], S2r[j]);
      s1 += zi;
    }
    float* e = ex + w * E;
    e[i] = s1;
#pragma unroll
    for (int j = 0; j < C; ++j) e[C + i * C + j] = S2r[j];
  }
  __syncthreads();
  if (w < nw) {
    // log(1+S) level 3, row i:  S3 - (S1 (x) S2 + S2 (x) S1)/2 + S1^3/3
    float* e = ex + w * E;
    float S1v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) S1v[c] = e[c];
#pragma unroll
    for (int j = 0; j < C; ++j)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float s2jk = e[C + j * C + k];
        float v = S3r[j][k] - 0.5f * (s1 * s2jk + S2r[j] * S1v[k]);
        v = fmaf((1.f / 3.f) * s1 * S1v[j], S1v[k], v);
        e[C + C2 + (i * C + j) * C + k] = v;
      }
  }
  __syncthreads();
  if (w < nw) {  // level 2 of the log in place: S2 - S1 (x) S1 / 2 (all readers of S2 are done)
    float* e = ex + w * E;
    const float s1i = e[i];
#pragma unroll
    for (int j = 0; j < C; ++j) e[C + i * C + j] -= 0.5f * s1i * e[j];
  }
  // sparse t2l (CSR, a few KB) into shared memory: it is re-read for every window of the tile
  const int D = p.D;
  int* s_rowptr = reinterpret_cast<int*>(win + WT * p.stride);
  const int nnz = pp.rowptr[D];
  int* s_cols = s_rowptr + D + 1;
  float* s_vals = reinterpret_cast<float*>(s_cols + nnz);
  for (int t = threadIdx.x; t <= D; t += blockDim.x) s_rowptr[t] = pp.rowptr[t];
  for (int t = threadIdx.x; t < nnz; t += blockDim.x) {
    s_cols[t] = pp.cols[t] - 1;  // flat tensor index without the scalar slot
    s_vals[t] = pp.vals[t];
  }
  __syncthreads();
  // Hall projection: one thread per Hall row, looping over the windows of the tile; for a fixed
  // window consecutive threads write consecutive floats
  {
    float* dst = p.out + ((size_t)b * p.K + k0) * D;
    for (int h = threadIdx.x; h < D; h += blockDim.x) {
      const int lo = s_rowptr[h], hi = s_rowptr[h + 1];
      for (int w0 = 0; w0 < nw; w0 += 8) {  // 8 windows at a time: independent accumulators
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int t = lo; t < hi; ++t) {
          const float v = s_vals[t];
          const float* e = ex + w0 * E + s_cols[t];
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = fmaf(v, e[u * E], acc[u]);  // windows beyond nw hold stale but finite data
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (w0 + u < nw) dst[(size_t)(w0 + u) * D + h] = acc[u];
      }
    }
  }
}

// ----------------------------------------------------------------------------
// generic kernel: any width, depth <= 4, one warp per window.
// ----------------------------------------------------------------------------
struct TensorParams {
  const float* data;
  float* out;
  const int32_t* rowptr;
  const int32_t* cols;
  const float* vals;
  int B, L, C, s, q, r, K, D, depth, compat, chunk;
  int tdim;  // C + C^2 + ... + C^depth
  int k_begin;  // only windows k >= k_begin are processed (the depth-3 kernel leaves the remainder here)
};

// read point m of the (virtual) window `k` of sample b.  Point 0 is the basepoint.
__device__ __forceinline__ float window_point(const TensorParams& p, const float* xb, int b, int k,
                                              int m, int c) {
  if (k < p.q) {
    const long long row = (long long)k * p.s - 2 + m;
    return row >= 0 ? xb[row * p.C + c] : 0.f;
  }
  // remainder window
  if (p.compat) {
    if (m == 0) return (b % p.chunk) != 0 ? xb[-p.C + c] : 0.f;
    return xb[(size_t)(p.L - p.r - 2 + m) * p.C + c];
  }
  return xb[(size_t)(p.L - p.r - 1 + m) * p.C + c];
}

// entry (word of length n, flat index e within its level) of log(1+S), depth <= 4
__device__ float log_entry(const float* S, const int* off, int C, int n, int e) {
  // off[k] = offset of level k (1-based) inside S
  if (n == 1) return S[off[1] + e];
  if (n == 2) {
    const int i = e / C, j = e - i * C;
    return S[off[2] + e] - 0.5f * S[off[1] + i] * S[off[1] + j];
  }
  if (n == 3) {
    const int i = e / (C * C), jk = e - i * C * C, j = jk / C, k = jk - j * C;
    const float a = S[off[1] + i], b = S[off[1] + j], c = S[off[1] + k];
    return S[off[3] + e] - 0.5f * (a * S[off[2] + j * C + k] + S[off[2] + i * C + j] * c) +
           (1.f / 3.f) * a * b * c;
  }
  // n == 4
  const int C2 = C * C, C3 = C2 * C;
  const int i = e / C3, r1 = e - i * C3, j = r1 / C2, r2 = r1 - j * C2, k = r2 / C, l = r2 - k * C;
  const float s1i = S[off[1] + i], s1j = S[off[1] + j], s1k = S[off[1] + k], s1l = S[off[1] + l];
  const float s2ij = S[off[2] + i * C + j], s2jk = S[off[2] + j * C + k], s2kl = S[off[2] + k * C + l];
  const float s3ijk = S[off[3] + (i * C + j) * C + k], s3jkl = S[off[3] + (j * C + k) * C + l];
  return S[off[4] + e] - 0.5f * (s1i * s3jkl + s2ij * s2kl + s3ijk * s1l) +
         (1.f / 3.f) * (s1i * s1j * s2kl + s1i * s2jk * s1l + s2ij * s1k * s1l) -
         0.25f * s1i * s1j * s1k * s1l;
}

__global__ void __launch_bounds__(128) logsig_tensor_kernel(const TensorParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float* S = smem + (size_t)warp * (p.tdim + 2 * p.C);
  float* z = S + p.tdim;       // current increment
  float* prev = z + p.C;       // previous point
  int off[5];
  off[1] = 0;
  for (int k = 2, pw = p.C; k <= 4; ++k, pw *= p.C) off[k] = off[k - 1] + pw;
  const int C = p.C;

  const int nk = p.K - p.k_begin;
  const long long n_win = (long long)p.B * nk;
  for (long long wi = (long long)blockIdx.x * wpb + warp; wi < n_win; wi += (long long)gridDim.x * wpb) {
    const int b = (int)(wi / nk), k = p.k_begin + (int)(wi - (long long)b * nk);
    const long long wid = (long long)b * p.K + k;
    const float* xb = p.data + (size_t)b * p.L * C;
    const int n_inc = k < p.q ? p.s : (p.compat ? p.r + 1 : p.r);
    for (int e = lane; e < p.tdim; e += 32) S[e] = 0.f;
    for (int c = lane; c < C; c += 32) prev[c] = window_point(p, xb, b, k, 0, c);
    __syncwarp();
    for (int m = 1; m <= n_inc; ++m) {
      for (int c = lane; c < C; c += 32) {
        const float x = window_point(p, xb, b, k, m, c);
        z[c] = x - prev[c];
        prev[c] = x;
      }
      __syncwarp();
      // Chen/Horner, top level first so lower levels are still the old ones
      int lvl_size = 1;
      for (int d = 1; d < p.depth; ++d) lvl_size *= C;
      lvl_size *= C;
      for (int n = p.depth; n >= 1; --n) {
        for (int e = lane; e < lvl_size; e += 32) {
          // decode word digits, most significant first
          int idx[4];
          int rem = e;
          for (int t = n - 1; t >= 0; --t) {
            idx[t] = rem % C;
            rem /= C;
          }
          float cur = S[off[1] + idx[0]] + z[idx[0]] / (float)n;
          int pre = idx[0];
          for (int t = 1; t < n; ++t) {
            pre = pre * C + idx[t];
            cur = fmaf(cur, z[idx[t]] / (float)(n - t), S[off[t + 1] + pre]);
          }
          // note: for t = n-1 the fma added S_n[e] itself (pre == e)
          if (n == 1) cur = S[off[1] + e] + z[e];
          S[off[n] + e] = cur;
        }
        __syncwarp();
        lvl_size /= C;
      }
    }
    // log + Hall projection; row 0 is identically 0
    float* dst = p.out + (size_t)wid * p.D;
    for (int h = lane; h < p.D; h += 32) {
      float acc = 0.f;
      const int lo = p.rowptr[h], hi = p.rowptr[h + 1];
      for (int t = lo; t < hi; ++t) {
        int col = p.cols[t] - 1;  // drop the scalar slot
        int n = 1, pw = C;
        while (col >= pw) {
          col -= pw;
          pw *= C;
          ++n;
        }
        acc = fmaf(p.vals[t], log_entry(S, off, C, n, col), acc);
      }
      dst[h] = acc;
    }
    __syncwarp();
  }
}

template <int C>
static int launch_levy_tma(cudaStream_t st, LevyParams p, int depth) {
  const int D = 1 + C + (depth == 2 ? (C * (C - 1)) / 2 : 0);
  p.WT = 32;
  p.tiles = (p.q + 31) / 32;
  if (p.tiles < 1) p.tiles = 1;
  // buffer: [kLevyPad | 32 windows + basepoint row + <= 3 alignment floats | 4 floats of over-read slack], or the
  // 32 x D result rows staged in the same place
  const size_t in_f = kLevyPad + (size_t)32 * p.s * C + C + 3 + 4;
  const size_t out_f = 3 + (size_t)32 * D + 4;
  const size_t buf_f = ((in_f > out_f ? in_f : out_f) + 3) & ~(size_t)3;
  const size_t smem = (4 + buf_f) * 4;  // 16 bytes of mbarrier in front
  const long long grid = (long long)p.B * p.tiles;
  auto go = [&](auto k) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(int)grid, 32, smem, st>>>(p, (int)buf_f);
  };
  if (depth == 2) go(logsig_levy_tma_kernel<C, true>);
  else go(logsig_levy_tma_kernel<C, false>);
  return (int)cudaGetLastError();
}

template <int C>
static int launch_levy(cudaStream_t st, LevyParams p, int depth, size_t smem) {
  const int grid = p.B * p.tiles;
  if (depth == 2) {
    auto k = logsig_levy_kernel<C, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, p.WT, smem, st>>>(p);
  } else {
    auto k = logsig_levy_kernel<C, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, p.WT, smem, st>>>(p);
  }
  return (int)cudaGetLastError();
}

}  // namespace lncde

extern "C" {

int lncde_logsig_num_windows(int L, int stepsize) {
  if (L < 1 || stepsize < 1) return LNCDE_ERR_SHAPE;
  const int n = L + 1;
  const int s = stepsize > n ? n : stepsize;
  return n / s + (n % s != 0 ? 1 : 0);
}

int lncde_logsig_fwd(void* stream, const float* data, float* logsig, int B, int L, int C, int stepsize,
                     int depth, unsigned flags, int chunk, const int32_t* t2l_rowptr,
                     const int32_t* t2l_cols, const float* t2l_vals) {
  using namespace lncde;
  if (!data || !logsig) return LNCDE_ERR_NULL;
  if (B < 1 || L < 1 || C < 1 || stepsize < 1) return LNCDE_ERR_SHAPE;
  if (depth < 1 || depth > 4) return LNCDE_ERR_UNSUPPORTED;
  const int compat = (flags & LNCDE_FLAG_COMPAT) ? 1 : 0;
  if (compat && chunk < 1) return LNCDE_ERR_SHAPE;
  const int D = lncde_hall_size(C, depth);
  if (D < 0) return D;
  const int n = L + 1;
  const int s = stepsize > n ? n : stepsize;
  const int q = n / s, r = n % s;
  const int K = q + (r ? 1 : 0);
  cudaStream_t st = (cudaStream_t)stream;

  // closed-form streaming kernel: C <= 8, depth <= 2, and a tile of >= 32
  // windows must fit comfortably in shared memory
  const long long sC = (long long)s * C;
  if (C <= 8 && depth <= 2 && sC <= 640) {
    LevyParams p;
    p.data = data;
    p.out = logsig;
    p.B = B; p.L = L; p.s = s; p.q = q; p.r = r; p.K = K; p.D = D;
    p.compat = compat;
    p.chunk = compat ? chunk : 1;
    p.stride = (int)(sC | 1);
    p.sc_magic = (uint32_t)((0x100000000ull + sC - 1) / sC);
    // Bulk-copy (TMA) kernel or classic staging.  Measured on B200 (round 2, scripts/k1_sweep.py): TMA 6 903 vs 4 178 GB/s
    // for EigenWorms (1 499 windows per path), 6 756 vs 4 230 for PPG, but 2 814 vs 3 541 for SCP1 (56 windows per
    // path: two tiles, the second one three-quarters full) - so the choice is by windows per path.  The TMA kernel
    // needs a 16-byte aligned `data` (the 16-byte floor of a tile's first float must not precede the buffer).
    // LNCDE_FLAG_K1_TMA forces it, LNCDE_FLAG_GENERIC forces the classic kernel; both give bit-identical results.
    const bool want_tma = (flags & LNCDE_FLAG_GENERIC) ? false : ((flags & LNCDE_FLAG_K1_TMA) != 0 || q >= 128);
    if (want_tma && ((uintptr_t)data & 15) == 0) {
      switch (C) {
        case 1: return launch_levy_tma<1>(st, p, depth);
        case 2: return launch_levy_tma<2>(st, p, depth);
        case 3: return launch_levy_tma<3>(st, p, depth);
        case 4: return launch_levy_tma<4>(st, p, depth);
        case 5: return launch_levy_tma<5>(st, p, depth);
        case 6: return launch_levy_tma<6>(st, p, depth);
        case 7: return launch_levy_tma<7>(st, p, depth);
        case 8: return launch_levy_tma<8>(st, p, depth);
      }
    }
    // measured on B200 (scripts/k1_sweep.py): 32-window, single-warp CTAs give the best overlap of
    // the load / walk / store phases across the ~21 CTAs resident per SM (63.8 % vs 59.7 % at 128)
    int WT = 32;
    while (WT > 32 && (size_t)(C + (size_t)WT * p.stride) * 4 > 80 * 1024) WT >>= 1;
    if (WT > 128) WT = 128;
    p.WT = WT;
    p.tiles = (q + WT - 1) / WT;
    if (p.tiles < 1) p.tiles = 1;
    size_t smem = (size_t)(C + (size_t)WT * p.stride) * 4;
    const size_t smem_out = ((size_t)WT * D + 4) * 4;
    if (smem_out > smem) smem = smem_out;
    switch (C) {
      case 1: return launch_levy<1>(st, p, depth, smem);
      case 2: return launch_levy<2>(st, p, depth, smem);
      case 3: return launch_levy<3>(st, p, depth, smem);
      case 4: return launch_levy<4>(st, p, depth, smem);
      case 5: return launch_levy<5>(st, p, depth, smem);
      case 6: return launch_levy<6>(st, p, depth, smem);
      case 7: return launch_levy<7>(st, p, depth, smem);
      case 8: return launch_levy<8>(st, p, depth, smem);
    }
  }

  if (!t2l_rowptr || !t2l_cols || !t2l_vals) return LNCDE_ERR_NULL;
  bool d3_done = false;
  if (depth == 3 && C >= 2 && C <= 8 && sC <= 640) {
    D3Params dp;
    LevyParams& lp = dp.lv;
    lp.data = data; lp.out = logsig;
    lp.B = B; lp.L = L; lp.s = s; lp.q = q; lp.r = r; lp.K = K; lp.D = D;
    lp.compat = compat; lp.chunk = compat ? chunk : 1;
    lp.stride = (int)(sC | 1);
    lp.sc_magic = (uint32_t)((0x100000000ull + sC - 1) / sC);
    lp.WT = 32;
    lp.tiles = (q + 31) / 32;
    dp.rowptr = t2l_rowptr; dp.cols = t2l_cols; dp.vals = t2l_vals;
    const size_t E = (size_t)C + C * C + C * C * C;
    const int nnz_h = lncde_hall_t2l_nnz(C, 3);
    const size_t smem = (32 * E + C + (size_t)32 * lp.stride + (size_t)(D + 1) + 2 * (size_t)nnz_h + 4) * 4;
    const int grid = B * lp.tiles;
    auto go = [&](auto k) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<grid, 32 * C, smem, st>>>(dp);
    };
    switch (C) {
      case 2: go(logsig_d3_kernel<2>); break;
      case 3: go(logsig_d3_kernel<3>); break;
      case 4: go(logsig_d3_kernel<4>); break;
      case 5: go(logsig_d3_kernel<5>); break;
      case 6: go(logsig_d3_kernel<6>); break;
      case 7: go(logsig_d3_kernel<7>); break;
      case 8: go(logsig_d3_kernel<8>); break;
    }
    const int e = (int)cudaGetLastError();
    if (e != 0) return e;
    if (r == 0) return 0;
    d3_done = true;  // the remainder windows (one per sample) go through the generic kernel below
  }
  TensorParams p;
  p.data = data; p.out = logsig;
  p.rowptr = t2l_rowptr; p.cols = t2l_cols; p.vals = t2l_vals;
  p.B = B; p.L = L; p.C = C; p.s = s; p.q = q; p.r = r; p.K = K; p.D = D; p.depth = depth;
  p.k_begin = d3_done ? q : 0;
  p.compat = compat;
  p.chunk = compat ? chunk : 1;
  const long long tdim = lncde_tensor_dim(C, depth) - 1;
  p.tdim = (int)tdim;
  int wpb = 4;
  size_t per_warp = (size_t)(tdim + 2 * C) * 4;
  while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
  const size_t smem = per_warp * wpb;
  if (smem > 227 * 1024) return LNCDE_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(logsig_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long n_win = (long long)B * (K - p.k_begin);
  long long grid = (n_win + wpb - 1) / wpb;
  if (grid > 148 * 64) grid = 148 * 64;
  logsig_tensor_kernel<<<(int)grid, wpb * 32, smem, st>>>(p);
  return (int)cudaGetLastError();
}

}  // extern "C"
