// K2 dense (DE-LNCDE) bracket gradient on the 5th-generation tensor cores - hand-written tcgen05 / TMEM kernel.
//
//   dLie[blk, i, j, l] = sum_m lam[m, blk bs + i] * y_prev[m, blk bs + j] * logsig[m, 1 + l]      m = (sample, step >= 1)
//   (the adjoint of flows = I + einsum(lie_brackets, logsigs[:, 1:]) followed by y_t = F_t y_{t-1},
//    models/LinearNeuralCDEs.py:318-329; dF_t = lam_t y_{t-1}^T is never materialised)
//
// As a GEMM:  D[i, (j, l)] = sum_m A[i, m] * Bm[(j, l), m]   with A = lam^T and Bm[(j, l), m] = y_prev[m, j] * logsig[m, 1 + l]
// formed on the fly: M = 128 rows i (the TMEM lanes), N = 16 j x 32 l (tiles of 32 basis elements, the last one padded) =
// 512 fp32 accumulator columns = all of TMEM, K = the (sample, step) pairs, split over CTAs (split-K partials, reduced by reduce_partials_kernel).
// 44 GFLOP x 3 for EigenWorms against 0.1 GB of operands: tensor-pipe work.  3xTF32 as in flow_tcgen05.cu: both operands
// are split into hi = tf32(x), lo = tf32(x - hi) on their way into shared memory and
// a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi, accumulated in fp32 in TMEM.
//
// Kernel (one CTA per (block, 128-row i tile, 16-column j group, 32-element basis tile, K split), 9 warps):
//   warps 0-7  producers: per stage of 16 m's read lam / y_prev / logsig rows straight from global memory (L2 resident,
//              one stage ahead in registers), form the products, split, and write the K-major / no-swizzle operand
//              image tcgen05.mma reads (core matrix = 8 rows x 16 B; hi and lo planes): 16 KB of A + 64 KB of B per stage,
//              2 stages; fence.proxy.async + arrive on full[s].  Afterwards the same warps are the epilogue: tcgen05.ld
//              (32 lanes x 32 columns = one (i, j) pair's basis row per thread) -> part[ks][blk][i][j][0 .. n_basis).
//   warp 8     TMEM owner + MMA issuer (one thread): per stage 2 K-steps x 2 column halves x 3 tcgen05.mma
//              (kind::tf32, M = 128, N = 256, K = 8), tcgen05.commit -> empty[s]; after the last stage commit -> done.
// Every mbarrier wait is bounded (tcgen05.cuh): a protocol error ends the kernel with a code in *err.
//
// Under the host emulator (LNCDE_EMU) the tensor pipe does not exist: the emulated build runs the SAME operand-forming
// code (row bookkeeping across samples, zero fill, split, operand image) and evaluates the three partial products from
// the shared-memory image in plain fp32 - everything but the MMA unit and TMEM.
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"
#include "ptx.cuh"
#if !defined(LNCDE_EMU)
#include "tcgen05.cuh"
#endif

namespace lncde {
namespace tc5d {

constexpr int kM = 128;                  // rows i of a tile = TMEM lanes
constexpr int kJ = 16;                   // columns j of a tile
constexpr int kL = 32;                   // basis elements per (i, j), padded
constexpr int kN = kJ * kL;              // accumulator columns (all 512 TMEM columns)
constexpr int kStageK = 16;              // (sample, step) pairs per stage
constexpr int kChunks = kStageK / 4;     // 16-byte K chunks per stage
constexpr int kAPlane = kChunks * kM * 4;  // floats: 8 KB
constexpr int kBPlane = kChunks * kN * 4;  // floats: 32 KB
constexpr int kStageFloats = 2 * kAPlane + 2 * kBPlane;  // A hi | A lo | B hi | B lo: 80 KB
constexpr int kStages = 2;
constexpr int kProducers = 256;
constexpr int kThreads = kProducers + 32;
constexpr size_t kSmemBytes = (size_t)kStages * kStageFloats * 4 + 128;

struct DlieParams {
  const float* lam;  // [B][T][H]  lam_t of the reverse sweep (row t >= 1 used)
  const float* ys;   // [B][T][H]
  const float* ls;   // [B][T][D]
  float* part;       // [ksplit][nb][bs][bs][n_basis]
  int T, D, H, bs, n_basis, ksplit, i_tiles, j_tiles, l_tiles;
  unsigned Ktot, per;  // B (T - 1) pairs, pairs per K split (multiple of kStageK)
  size_t n_out;
  int* err;
};

struct Tile {
  int ks, blk, i0, j0, l0;  // l0: first of the (at most) 32 basis elements of this tile
  unsigned k_lo, k_hi;
};
__device__ __forceinline__ Tile tile_of(const DlieParams& p, int bid) {
  Tile t;
  t.ks = bid % p.ksplit; bid /= p.ksplit;
  t.l0 = (bid % p.l_tiles) * kL; bid /= p.l_tiles;
  t.j0 = (bid % p.j_tiles) * kJ; bid /= p.j_tiles;
  t.i0 = (bid % p.i_tiles) * kM; bid /= p.i_tiles;
  t.blk = bid;
  t.k_lo = min(p.Ktot, p.per * (unsigned)t.ks);
  t.k_hi = min(p.Ktot, t.k_lo + p.per);
  return t;
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(cvt_tf32(x));
  lo = __uint_as_float(cvt_tf32(x - hi));
}

// What one producer thread reads for one stage.  Thread (warp w, lane l) owns the operand rows (j0 + w, l) and
// (j0 + w + 8, l) of Bm and row i = tid % 128 of A for the K chunks 2 (tid / 128), 2 (tid / 128) + 1.
struct StageRegs {
  float ls[kStageK], ya[kStageK], yb[kStageK], lam[8];
};

// pair k = (sample b, step 1 + rem): lam / logsig row k + b + 1, y_prev row k + b.  (b, rem) of the stage's first pair
// are carried by the caller; pairs at or beyond k_hi read as zero.
__device__ __forceinline__ void load_stage(StageRegs& r, const DlieParams& p, const Tile& t, unsigned k0, unsigned b,
                                           unsigned rem, int tid) {
  const int lane = tid & 31, warp = tid >> 5, half = tid >> 7;
  const unsigned Tm1 = (unsigned)p.T - 1u;
  const float* lam = p.lam + (size_t)t.blk * p.bs + t.i0 + (tid & 127);
  const float* ya = p.ys + (size_t)t.blk * p.bs + t.j0 + warp;
  const float* ls = p.ls + 1 + t.l0 + lane;
  const bool l_ok = t.l0 + lane < p.n_basis;
#pragma unroll
  for (int kk = 0; kk < kStageK; ++kk) {
    const unsigned k = k0 + kk;
    const bool ok = k < t.k_hi;
    const size_t row = (size_t)k + b + 1;
    r.ls[kk] = (ok && l_ok) ? __ldg(ls + row * p.D) : 0.f;
    r.ya[kk] = ok ? __ldg(ya + (row - 1) * p.H) : 0.f;
    r.yb[kk] = ok ? __ldg(ya + (row - 1) * p.H + 8) : 0.f;
    if ((kk >> 3) == half) r.lam[kk & 7] = ok ? __ldg(lam + row * p.H) : 0.f;
    if (++rem == Tm1) {
      rem = 0;
      ++b;
    }
  }
}

// operand image of one stage: plane[(chunk * rows + row) * 4 + k % 4], rows = kM (A) / kN (Bm)
__device__ __forceinline__ void form_stage(const StageRegs& r, float* stage, int tid) {
  const int lane = tid & 31, warp = tid >> 5, half = tid >> 7, i = tid & 127;
  float* a_hi = stage;
  float* a_lo = a_hi + kAPlane;
  float* b_hi = a_lo + kAPlane;
  float* b_lo = b_hi + kBPlane;
#pragma unroll
  for (int c2 = 0; c2 < 2; ++c2) {
    float hi[4], lo[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) split_tf32(r.lam[c2 * 4 + x], hi[x], lo[x]);
    const int at = ((half * 2 + c2) * kM + i) * 4;
    *reinterpret_cast<float4*>(a_hi + at) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(a_lo + at) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      float hi[4], lo[4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
        split_tf32((jh ? r.yb[c * 4 + x] : r.ya[c * 4 + x]) * r.ls[c * 4 + x], hi[x], lo[x]);
      const int at = (c * kN + (warp + 8 * jh) * kL + lane) * 4;
      *reinterpret_cast<float4*>(b_hi + at) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(b_lo + at) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// (b, rem) of pair k
__device__ __forceinline__ void pair_of(unsigned k, unsigned Tm1, unsigned& b, unsigned& rem) {
  b = k / Tm1;
  rem = k - b * Tm1;
}
__device__ __forceinline__ void advance_pairs(unsigned n, unsigned Tm1, unsigned& b, unsigned& rem) {
  rem += n;
  while (rem >= Tm1) {
    rem -= Tm1;
    ++b;
  }
}

// part[ks][blk][i][j][l0 ..]: the tile's basis elements of pair (i, j); n_l of them exist
__device__ __forceinline__ float* out_row(const DlieParams& p, const Tile& t, int i, int jl) {
  return p.part + (size_t)t.ks * p.n_out + (((size_t)t.blk * p.bs + t.i0 + i) * p.bs + t.j0 + jl) * p.n_basis + t.l0;
}
__device__ __forceinline__ int tile_basis(const DlieParams& p, const Tile& t) { return min(kL, p.n_basis - t.l0); }

#if defined(LNCDE_EMU)
// emulator stand-in for the tensor pipe: the same stage images, the three partial products in plain fp32
__global__ void __launch_bounds__(kProducers) dlie_tc5_reference_kernel(const DlieParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x;
  const Tile t = tile_of(p, blockIdx.x);
  const int n_l = tile_basis(p, t);
  for (int e = tid; e < kM * kJ; e += kProducers)
    for (int l = 0; l < n_l; ++l) out_row(p, t, e / kJ, e % kJ)[l] = 0.f;
  unsigned b, rem;
  pair_of(t.k_lo, (unsigned)p.T - 1u, b, rem);
  for (unsigned k0 = t.k_lo; k0 < t.k_hi; k0 += kStageK) {
    StageRegs r;
    load_stage(r, p, t, k0, b, rem, tid);
    advance_pairs(kStageK, (unsigned)p.T - 1u, b, rem);
    __syncthreads();
    form_stage(r, stage, tid);
    __syncthreads();
    const float* a_hi = stage;
    const float* a_lo = a_hi + kAPlane;
    const float* b_hi = a_lo + kAPlane;
    const float* b_lo = b_hi + kBPlane;
    for (int e = tid; e < kM * kJ; e += kProducers) {
      const int i = e / kJ, jl = e % kJ;
      float* o = out_row(p, t, i, jl);
      for (int l = 0; l < n_l; ++l) {
        float acc = o[l];
        for (int k = 0; k < kStageK; ++k) {
          const int ia = ((k >> 2) * kM + i) * 4 + (k & 3), ib = ((k >> 2) * kN + jl * kL + l) * 4 + (k & 3);
          acc += a_lo[ia] * b_hi[ib];
          acc += a_hi[ia] * b_lo[ib];
          acc += a_hi[ia] * b_hi[ib];
        }
        o[l] = acc;
      }
    }
  }
}
#else

__global__ void __launch_bounds__(kThreads, 1) dlie_tcgen05_kernel(const DlieParams p) {
  using namespace tc5;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)kStages * kStageFloats * 4);
  unsigned long long* full = bars;       // [kStages]  producers -> MMA (one arrival per producer thread)
  unsigned long long* empty = bars + 2;  // [kStages]  MMA -> producers
  unsigned long long* done = bars + 4;   //            MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Tile t = tile_of(p, blockIdx.x);
  const unsigned n_stages = (t.k_hi - t.k_lo + kStageK - 1) / kStageK;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, kProducers);
      mbar_init(empty + s, 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, kN);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ producers
    const unsigned Tm1 = (unsigned)p.T - 1u;
    unsigned b, rem;
    pair_of(t.k_lo, Tm1, b, rem);
    StageRegs cur, nxt;
    if (n_stages > 0) load_stage(cur, p, t, t.k_lo, b, rem, tid);
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (unsigned st = 0; st < n_stages; ++st) {
      if (st + 1 < n_stages) {
        advance_pairs(kStageK, Tm1, b, rem);
        load_stage(nxt, p, t, t.k_lo + (st + 1) * kStageK, b, rem, tid);
      }
      if (!(ok = mbar_wait_bounded(empty + s, ph ^ 1u, p.err, 11))) break;
      form_stage(cur, stages + (size_t)s * kStageFloats, tid);
      fence_async_proxy();  // my generic-proxy writes -> visible to the tensor pipe's reads
      mbar_arrive(full + s);
      cur = nxt;
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    // ------------------------------------------------------------------ epilogue: thread = TMEM lane = row i
    if (ok) ok = mbar_wait_bounded(done, 0, p.err, 12);
    ok = __all_sync(0xffffffffu, ok);
    if (ok) {
      fence_after_thread_sync();
      const int quarter = warp & 3, i = quarter * 32 + lane, jbase = (warp >> 2) * (kJ / 2);
      const int n_l = tile_basis(p, t);
      for (int jj = 0; jj < kJ / 2; ++jj) {
        float v[32];
        if (n_stages > 0) {
          tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((jbase + jj) * kL), v);
        } else {
#pragma unroll
          for (int l = 0; l < 32; ++l) v[l] = 0.f;
        }
        float* o = out_row(p, t, i, jbase + jj);
        if ((p.n_basis & 3) == 0) {  // rows of the partial buffer are 16-byte aligned, n_l is a multiple of 4
#pragma unroll
          for (int l = 0; l < 32; l += 4)
            if (l < n_l) *reinterpret_cast<float4*>(o + l) = make_float4(v[l], v[l + 1], v[l + 2], v[l + 3]);
        } else {
#pragma unroll
          for (int l = 0; l < 32; ++l)
            if (l < n_l) o[l] = v[l];
        }
      }
    }
  } else if (lane == 0) {
    // -------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = instr_desc_tf32(kM, 256);
    constexpr uint32_t kLboA = kM * 16, kLboB = kN * 16, kSbo = 128;
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (unsigned st = 0; st < n_stages; ++st) {
      if (!(ok = mbar_wait_bounded(full + s, ph, p.err, 13))) break;
      fence_after_thread_sync();
      const uint32_t a_hi = smem_u32(stages + (size_t)s * kStageFloats);
      const uint32_t a_lo = a_hi + kAPlane * 4, b_hi = a_lo + kAPlane * 4, b_lo = b_hi + kBPlane * 4;
#pragma unroll
      for (int kk = 0; kk < kStageK / 8; ++kk) {
        const uint64_t dah = smem_desc_kmajor_noswizzle(a_hi + kk * 2 * kLboA, kLboA, kSbo);
        const uint64_t dal = smem_desc_kmajor_noswizzle(a_lo + kk * 2 * kLboA, kLboA, kSbo);
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
          const uint32_t off = kk * 2 * kLboB + nh * 256 * 16;
          const uint64_t dbh = smem_desc_kmajor_noswizzle(b_hi + off, kLboB, kSbo);
          const uint64_t dbl = smem_desc_kmajor_noswizzle(b_lo + off, kLboB, kSbo);
          const uint32_t d = tmem + (uint32_t)nh * 256;
          mma_tf32_ss(d, dal, dbh, idesc, (st | (unsigned)kk) != 0);  // small terms first
          mma_tf32_ss(d, dah, dbl, idesc, 1);
          mma_tf32_ss(d, dah, dbh, idesc, 1);
        }
      }
      mma_commit(empty + s);
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    if (ok) mma_commit(done);
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 8) {
    fence_after_thread_sync();
    tmem_dealloc(tmem, kN);
  }
}
#endif

}  // namespace tc5d

// shapes the kernel takes: whole 128-row i tiles and 16-column j groups inside a block; any number of basis elements
// (tiles of 32: depth 3 of 7 channels has 140)
bool dlie_tcgen05_shape(int bs, int n_basis) { return bs >= tc5d::kM && bs % tc5d::kM == 0 && n_basis >= 1; }

// K splits used for this shape (<= max_ksplit, the size of the caller's partial buffer)
int dlie_tcgen05_ksplit(int B, int T, int H, int bs, int n_basis, int max_ksplit) {
  const long long tiles = (long long)(H / bs) * (bs / tc5d::kM) * (bs / tc5d::kJ) * ((n_basis + tc5d::kL - 1) / tc5d::kL);
  const long long stages = ((long long)B * (T - 1) + tc5d::kStageK - 1) / tc5d::kStageK;
  long long ks = 148 / (tiles < 1 ? 1 : tiles);
  if (ks > stages) ks = stages;
  if (ks > max_ksplit) ks = max_ksplit;
  return (int)(ks < 1 ? 1 : ks);
}

// part[ksplit][nb bs bs n_basis] (every element written) from lam / ys [B][T][H] and logsig [B][T][D]; err: one int of
// device scratch.  Returns the number of K splits written through *ksplit_out.
int dlie_tcgen05(cudaStream_t cs, const float* lam, const float* ys, const float* logsig, float* part, int B, int T,
                 int D, int H, int bs, int n_basis, int max_ksplit, int* err, int* ksplit_out) {
  using namespace tc5d;
  if (!dlie_tcgen05_shape(bs, n_basis) || T < 2 || H % bs != 0) return LNCDE_ERR_UNSUPPORTED;
  if ((long long)B * (T - 1) + B + 1 >= (1ll << 31)) return LNCDE_ERR_UNSUPPORTED;
  DlieParams p;
  p.lam = lam; p.ys = ys; p.ls = logsig; p.part = part;
  p.T = T; p.D = D; p.H = H; p.bs = bs; p.n_basis = n_basis;
  p.ksplit = dlie_tcgen05_ksplit(B, T, H, bs, n_basis, max_ksplit);
  p.i_tiles = bs / kM; p.j_tiles = bs / kJ; p.l_tiles = (n_basis + kL - 1) / kL;
  p.Ktot = (unsigned)B * (unsigned)(T - 1);
  const unsigned per = (p.Ktot + p.ksplit - 1) / p.ksplit;
  p.per = (per + kStageK - 1) / kStageK * kStageK;
  p.n_out = (size_t)H * bs * n_basis;
  p.err = err;
  *ksplit_out = p.ksplit;
  const long long grid_ll = (long long)(H / bs) * p.i_tiles * p.j_tiles * p.l_tiles * p.ksplit;
  if (grid_ll > 0x7fffffffLL) return LNCDE_ERR_UNSUPPORTED;
  const int grid = (int)grid_ll;
#if defined(LNCDE_EMU)
  const size_t smem = (size_t)kStageFloats * 4;
  cudaFuncSetAttribute(dlie_tc5_reference_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dlie_tc5_reference_kernel<<<grid, kProducers, smem, cs>>>(p);
#else
  int st = (int)cudaMemsetAsync(err, 0, sizeof(int), cs);
  if (st != 0) return st;
  cudaFuncSetAttribute(dlie_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  dlie_tcgen05_kernel<<<grid, kThreads, kSmemBytes, cs>>>(p);
#endif
  return (int)cudaGetLastError();
}

}  // namespace lncde
