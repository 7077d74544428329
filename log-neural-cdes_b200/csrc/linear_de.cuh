// K2 dense (DE-LNCDE, block size >= 64): the kernels that dominate the step besides the scans.
//
// ncu launch list of round 1 (EigenWorms DE, 32 x 1499 x 128 x 128, n_basis 28): flow build 3.8 ms for 44 GFLOP
// and 3.1 GB of output (16 % of the FP32 peak - its 4 x 4 register tile issues one shared-memory load per two
// FMAs), bracket-gradient GEMM 1.7 ms reading 3.1 GB of dF that the reverse scan had to write first.
//
//  * (an FFMA flow build with 8 x 8 register tiles was measured in round 2 - 7.5 ms per EigenWorms DE step against
//    7.3 ms with the mma.sync build below and 5.7 ms with the tcgen05 build of flow_tcgen05.cu - and removed.)
//  * dlie_triple_kernel: dLie[(i,j)][l] = sum_{b,t>=1} lam_t[i] * y_{t-1}[j] * ls[t][1+l] straight from the small
//    operands (lam, y: 24 MB each, ls: 5 MB) - the reverse scan then only has to store lam_t (512 B per step)
//    instead of the outer product dF_t (64 KB per step), and nothing reads 3.1 GB back.  Same FLOPs as the GEMM it
//    replaces (the rank-1 factor costs 2 extra multiplies per 56 FMAs), split-K with a deterministic reduce.
//  * flow_build_tc_kernel: the flow build on the warp-level tensor pipe (mma.sync, 3xTF32 split precision, see below) -
//    used for dense block sizes the tcgen05 kernel does not take (not a multiple of 128) or with LNCDE_FLAG_K2_MMA_SYNC.
//  * scan_bwd_tma_lam_kernel: the reverse sweep of scan_bwd_tma_kernel (linear_scan.cu) with the in-place dF_t store
//    replaced by a 512-byte store of lam_t - same ring of TMA bulk copies, same order of the sums, so lam, g_y0 are
//    bit-identical to the default kernel; the flows in the workspace stay intact.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx.cuh"

namespace lncde {

// ---------------------------------------------------------------- flow build on the tensor pipe (3xTF32)
// Same contraction on the warp-level TF32 MMA (mma.sync m16n8k8) with the split-precision scheme
//   a b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi,   x_hi = tf32(x), x_lo = tf32(x - x_hi)
// (the dropped a_lo b_lo term and the rounding of x_lo are ~2^-22 relative: fp32-grade flows, inside the 1e-5
// parity bar; plain TF32 would not be).  CTA tile 128 x 128, 8 warps as 2 (M) x 4 (N), warp tile 64 x 32 =
// 4 x 4 MMA tiles with fp32 accumulators in registers; the operands are split once while they are staged into
// shared memory (row stride 36 words: the fragment reads of a warp hit 32 different banks).  K = n_basis is
// walked in tiles of 32 (zero padded): depth 2 needs one tile, depth 3 (n_basis up to 140) five.  With K this
// small the kernel is bound by writing the flows (64 KB per CTA), which is the point: the FFMA kernels above
// are issue-bound at 2-3x the time the bytes need.
constexpr int kTcTile = 128, kTcK = 32, kTcStride = kTcK + 4;
constexpr size_t kTcSmemBytes = (size_t)4 * kTcTile * kTcStride * sizeof(uint32_t);

__global__ void __launch_bounds__(256, 2) flow_build_tc_kernel(const float* __restrict__ lie,
                                                               const float* __restrict__ logsig, float* __restrict__ F,
                                                               long long Mtot, int N, int D, int n_basis, int bs) {
  extern __shared__ __align__(16) unsigned char tc_raw[];
  uint32_t* Ah = reinterpret_cast<uint32_t*>(tc_raw);  // [128][36] logsig tile, hi parts
  uint32_t* Al = Ah + kTcTile * kTcStride;             // lo parts
  uint32_t* Bh = Al + kTcTile * kTcStride;             // [128][36] lie tile (row n, column k)
  uint32_t* Bl = Bh + kTcTile * kTcStride;
  const long long m0 = (long long)blockIdx.x * kTcTile;  // row tiles on grid.x (no 65 535 limit)
  const int n0 = blockIdx.y * kTcTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
  for (int k0 = 0; k0 < n_basis; k0 += kTcK) {
    // consecutive threads read consecutive k of one row: contiguous in global memory, conflict-free in shared
    for (int e = threadIdx.x; e < kTcTile * kTcK; e += 256) {
      const int r = e >> 5, kk = e & 31;
      const bool kin = k0 + kk < n_basis;
      const float a = (kin && m0 + r < Mtot) ? __ldg(logsig + (m0 + r) * D + 1 + k0 + kk) : 0.f;
      const float b = (kin && n0 + r < N) ? __ldg(lie + (size_t)(n0 + r) * n_basis + k0 + kk) : 0.f;
      const uint32_t ah = cvt_tf32(a), bh = cvt_tf32(b);
      Ah[r * kTcStride + kk] = ah;
      Al[r * kTcStride + kk] = cvt_tf32(a - __uint_as_float(ah));
      Bh[r * kTcStride + kk] = bh;
      Bl[r * kTcStride + kk] = cvt_tf32(b - __uint_as_float(bh));
    }
    __syncthreads();
    const int ksteps = (min(kTcK, n_basis - k0) + 7) >> 3;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int kc = ks * 8 + t;
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int o = (wn + nt * 8 + g) * kTcStride + kc;
        bh[nt][0] = Bh[o]; bh[nt][1] = Bh[o + 4];
        bl[nt][0] = Bl[o]; bl[nt][1] = Bl[o + 4];
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int o = (wm + mt * 16 + g) * kTcStride + kc;
        const uint32_t ah[4] = {Ah[o], Ah[o + 8 * kTcStride], Ah[o + 4], Ah[o + 8 * kTcStride + 4]};
        const uint32_t al[4] = {Al[o], Al[o + 8 * kTcStride], Al[o + 4], Al[o + 8 * kTcStride + 4]};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_m16n8k8_tf32(acc[mt][nt], al, bh[nt]);  // small terms first
          mma_m16n8k8_tf32(acc[mt][nt], ah, bl[nt]);
          mma_m16n8k8_tf32(acc[mt][nt], ah, bh[nt]);
        }
      }
    }
    __syncthreads();
  }
  const int bs2 = bs * bs;
  const bool vec2 = (N & 1) == 0;  // rows of F 8-byte aligned (the workspace is 16-byte aligned)
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const long long m = m0 + wm + mt * 16 + g + 8 * half;
      if (m >= Mtot) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = n0 + wn + nt * 8 + 2 * t;
        float v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = (n + j) % bs2;
          v[j] = acc[mt][nt][2 * half + j] + ((r / bs == r % bs) ? 1.f : 0.f);
        }
        float* dst = F + m * N + n;
        if (vec2 && n + 1 < N) {
          *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
        } else {
          if (n < N) dst[0] = v[0];
          if (n + 1 < N) dst[1] = v[1];
        }
      }
    }
}

// ---------------------------------------------------------------- bracket gradient without dF
// out[ks][((blk*bs + i)*bs + j)*n_basis + l] = sum over the k-slice of lam[b,t,blk*bs+i] * y[b,t-1,blk*bs+j] * ls[b,t,1+l]
// with k = b*(T-1) + (t-1), t = 1..T-1.  CTA tile: 16 (i) x 32 (j) x all l; thread (ty = i, tx) owns j = tx and tx + 16.
template <int NBP>  // n_basis rounded up to a multiple of 4 (<= 36)
__global__ void __launch_bounds__(256) dlie_triple_kernel(const float* __restrict__ lam, const float* __restrict__ ys,
                                                          const float* __restrict__ logsig, float* __restrict__ part,
                                                          int B, int T, int D, int H, int bs, int n_basis, int ksplit,
                                                          int i_tiles, int j_tiles) {
  constexpr int KC = 32, TI = 16, TJ = 32;
  __shared__ __align__(16) float lam_s[KC][TI];
  __shared__ __align__(16) float y_s[KC][TJ];
  __shared__ __align__(16) float ls_s[KC][NBP];
  int bid = blockIdx.x;
  const int ks = bid % ksplit; bid /= ksplit;
  const int jt = bid % j_tiles; bid /= j_tiles;
  const int it = bid % i_tiles; bid /= i_tiles;
  const int blk = bid;
  const int i0 = it * TI, j0 = jt * TJ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const unsigned Tm1 = (unsigned)(T - 1);
  const unsigned Ktot = (unsigned)B * Tm1;  // < 2^31 (checked by the dispatcher): 32-bit index arithmetic
  const unsigned per = (Ktot + ksplit - 1) / ksplit;
  const unsigned k_lo = min(Ktot, per * ks), k_hi = min(Ktot, k_lo + per);
  float acc0[NBP], acc1[NBP];
#pragma unroll
  for (int l = 0; l < NBP; ++l) acc0[l] = acc1[l] = 0.f;
  for (unsigned kc = k_lo; kc < k_hi; kc += KC) {
    for (int e = threadIdx.x; e < KC * TI; e += 256) {
      const int kk = e / TI, i = e - kk * TI;
      const unsigned k = kc + kk;
      float v = 0.f;
      if (k < k_hi && i0 + i < bs) {
        const unsigned b = k / Tm1, t = 1 + (k - b * Tm1);
        v = __ldg(lam + ((size_t)b * T + t) * H + blk * bs + i0 + i);
      }
      lam_s[kk][i] = v;
    }
    for (int e = threadIdx.x; e < KC * TJ; e += 256) {
      const int kk = e / TJ, j = e - kk * TJ;
      const unsigned k = kc + kk;
      float v = 0.f;
      if (k < k_hi && j0 + j < bs) {
        const unsigned b = k / Tm1, t = 1 + (k - b * Tm1);
        v = __ldg(ys + ((size_t)b * T + t - 1) * H + blk * bs + j0 + j);
      }
      y_s[kk][j] = v;
    }
    for (int e = threadIdx.x; e < KC * NBP; e += 256) {
      const int kk = e / NBP, l = e - kk * NBP;
      const unsigned k = kc + kk;
      float v = 0.f;
      if (k < k_hi && l < n_basis) {
        const unsigned b = k / Tm1, t = 1 + (k - b * Tm1);
        v = __ldg(logsig + ((size_t)b * T + t) * D + 1 + l);
      }
      ls_s[kk][l] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int kk = 0; kk < KC; ++kk) {
      const float li = lam_s[kk][ty];
      const float w0 = li * y_s[kk][tx], w1 = li * y_s[kk][tx + 16];
#pragma unroll
      for (int q = 0; q < NBP / 4; ++q) {
        const float4 s4 = *reinterpret_cast<const float4*>(&ls_s[kk][4 * q]);
        acc0[4 * q + 0] = fmaf(w0, s4.x, acc0[4 * q + 0]);
        acc0[4 * q + 1] = fmaf(w0, s4.y, acc0[4 * q + 1]);
        acc0[4 * q + 2] = fmaf(w0, s4.z, acc0[4 * q + 2]);
        acc0[4 * q + 3] = fmaf(w0, s4.w, acc0[4 * q + 3]);
        acc1[4 * q + 0] = fmaf(w1, s4.x, acc1[4 * q + 0]);
        acc1[4 * q + 1] = fmaf(w1, s4.y, acc1[4 * q + 1]);
        acc1[4 * q + 2] = fmaf(w1, s4.z, acc1[4 * q + 2]);
        acc1[4 * q + 3] = fmaf(w1, s4.w, acc1[4 * q + 3]);
      }
    }
    __syncthreads();
  }
  const size_t n_out = (size_t)(H / bs) * bs * bs * n_basis;
  float* out = part + (size_t)ks * n_out;
  const int i = i0 + ty;
  if (i < bs) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = j0 + tx + 16 * half;
      if (j >= bs) continue;
      float* o = out + (((size_t)blk * bs + i) * bs + j) * n_basis;
#pragma unroll
      for (int l = 0; l < NBP; ++l)
        if (l < n_basis) o[l] = half ? acc1[l] : acc0[l];
    }
  }
}

// ---------------------------------------------------------------- bracket gradient on the tensor pipe (3xTF32)
// The same triple product as a GEMM whose left operand is formed on the fly: rows (i, j) of A[(i,j)][k] =
// lam[k][i] * y[k][j] are built in registers (one multiply, then the hi / lo split), B[k][l] = logsig row k is split
// once per staged chunk.  CTA tile as above (16 i x 32 j x all l <= 32), 8 warps x (2 i x 32 j) = 4 MMA row tiles of
// 16 (i fixed, 16 consecutive j) x 4 column tiles of 8 basis elements; per k-step of 8 and warp 48 HMMA against
// 16 multiplies + 32 conversions - the FFMA kernel above spends 56 FMAs per (i, j, k) instead of 4 ALU operations.
constexpr int kTcYs = 40, kTcLs = 40;  // row strides (words) of the staged y / logsig chunks: = 8 (mod 32), conflict-free fragments

__global__ void __launch_bounds__(256, 2) dlie_triple_tc_kernel(const float* __restrict__ lam, const float* __restrict__ ys,
                                                                const float* __restrict__ logsig, float* __restrict__ part,
                                                                int B, int T, int D, int H, int bs, int n_basis, int ksplit,
                                                                int i_tiles, int j_tiles) {
  constexpr int KC = 32, TI = 16, TJ = 32;
  __shared__ __align__(16) float lam_s[KC][TI];
  __shared__ __align__(16) float y_s[KC][kTcYs];
  __shared__ __align__(16) uint32_t lsh[KC][kTcLs];
  __shared__ __align__(16) uint32_t lsl[KC][kTcLs];
  int bid = blockIdx.x;
  const int ks = bid % ksplit; bid /= ksplit;
  const int jt = bid % j_tiles; bid /= j_tiles;
  const int it = bid % i_tiles; bid /= i_tiles;
  const int blk = bid;
  const int i0 = it * TI, j0 = jt * TJ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const unsigned Tm1 = (unsigned)(T - 1);
  const unsigned Ktot = (unsigned)B * Tm1;
  const unsigned per = (Ktot + ksplit - 1) / ksplit;
  const unsigned k_lo = min(Ktot, per * ks), k_hi = min(Ktot, k_lo + per);
  float acc[4][4][4];  // [m-tile = 2 * ii + jh][n-tile][fragment]
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[a][b][r] = 0.f;
  for (unsigned kc = k_lo; kc < k_hi; kc += KC) {
    for (int e = threadIdx.x; e < KC * TI; e += 256) {
      const int kk = e / TI, i = e - kk * TI;
      const unsigned k = kc + kk;
      float v = 0.f;
      if (k < k_hi && i0 + i < bs) {
        const unsigned b = k / Tm1, tt = 1 + (k - b * Tm1);
        v = __ldg(lam + ((size_t)b * T + tt) * H + blk * bs + i0 + i);
      }
      lam_s[kk][i] = v;
    }
    for (int e = threadIdx.x; e < KC * TJ; e += 256) {
      const int kk = e / TJ, j = e - kk * TJ;
      const unsigned k = kc + kk;
      float v = 0.f;
      if (k < k_hi && j0 + j < bs) {
        const unsigned b = k / Tm1, tt = 1 + (k - b * Tm1);
        v = __ldg(ys + ((size_t)b * T + tt - 1) * H + blk * bs + j0 + j);
      }
      y_s[kk][j] = v;
    }
    for (int e = threadIdx.x; e < KC * 32; e += 256) {
      const int kk = e >> 5, l = e & 31;
      const unsigned k = kc + kk;
      float v = 0.f;
      if (k < k_hi && l < n_basis) {
        const unsigned b = k / Tm1, tt = 1 + (k - b * Tm1);
        v = __ldg(logsig + ((size_t)b * T + tt) * D + 1 + l);
      }
      const uint32_t hi = cvt_tf32(v);
      lsh[kk][l] = hi;
      lsl[kk][l] = cvt_tf32(v - __uint_as_float(hi));
    }
    __syncthreads();
#pragma unroll 1
    for (int kst = 0; kst < KC / 8; ++kst) {
      const int ka = kst * 8 + t, kb = ka + 4;
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        bh[nt][0] = lsh[ka][nt * 8 + g]; bh[nt][1] = lsh[kb][nt * 8 + g];
        bl[nt][0] = lsl[ka][nt * 8 + g]; bl[nt][1] = lsl[kb][nt * 8 + g];
      }
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const float y00 = y_s[ka][16 * jh + g], y01 = y_s[ka][16 * jh + g + 8];
        const float y10 = y_s[kb][16 * jh + g], y11 = y_s[kb][16 * jh + g + 8];
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
          const float la = lam_s[ka][2 * warp + ii], lb = lam_s[kb][2 * warp + ii];
          const float a[4] = {la * y00, la * y01, lb * y10, lb * y11};
          uint32_t ah[4], al[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            ah[r] = cvt_tf32(a[r]);
            al[r] = cvt_tf32(a[r] - __uint_as_float(ah[r]));
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            mma_m16n8k8_tf32(acc[2 * ii + jh][nt], al, bh[nt]);
            mma_m16n8k8_tf32(acc[2 * ii + jh][nt], ah, bl[nt]);
            mma_m16n8k8_tf32(acc[2 * ii + jh][nt], ah, bh[nt]);
          }
        }
      }
    }
    __syncthreads();
  }
  const size_t n_out = (size_t)(H / bs) * bs * bs * n_basis;
  float* out = part + (size_t)ks * n_out;
#pragma unroll
  for (int ii = 0; ii < 2; ++ii) {
    const int i = i0 + 2 * warp + ii;
    if (i >= bs) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = j0 + 16 * jh + g + 8 * half;
        if (j >= bs) continue;
        float* o = out + (((size_t)blk * bs + i) * bs + j) * n_basis;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int l = nt * 8 + 2 * t + c;
            if (l < n_basis) o[l] = acc[2 * ii + jh][nt][2 * half + c];
          }
      }
  }
}

template <int NBP>
static inline void launch_dlie_triple(cudaStream_t st, int grid, const float* lam, const float* ys, const float* ls,
                                      float* part, int B, int T, int D, int H, int bs, int n_basis, int ksplit, int it,
                                      int jt) {
  dlie_triple_kernel<NBP><<<grid, 256, 0, st>>>(lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt);
}

// returns false if n_basis is outside the compiled range
static inline bool dlie_triple_dispatch(cudaStream_t st, const float* lam, const float* ys, const float* ls, float* part,
                                        int B, int T, int D, int H, int bs, int n_basis, int ksplit, bool tensor) {
  if (T < 2 || (long long)B * (T - 1) >= (1ll << 31) - 64) return false;
  const int nb = H / bs, it = (bs + 15) / 16, jt = (bs + 31) / 32;
  const int grid = nb * it * jt * ksplit;
  if (tensor && n_basis <= 32) {
    dlie_triple_tc_kernel<<<grid, 256, 0, st>>>(lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt);
    return true;
  }
  const int nbp = (n_basis + 3) & ~3;
  switch (nbp) {
    case 4: launch_dlie_triple<4>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 8: launch_dlie_triple<8>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 12: launch_dlie_triple<12>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 16: launch_dlie_triple<16>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 20: launch_dlie_triple<20>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 24: launch_dlie_triple<24>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 28: launch_dlie_triple<28>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 32: launch_dlie_triple<32>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
    case 36: launch_dlie_triple<36>(st, grid, lam, ys, ls, part, B, T, D, H, bs, n_basis, ksplit, it, jt); return true;
  }
  return false;
}

// ---------------------------------------------------------------- reverse scan that stores lam_t only
// lam_{T-1} = g_{T-1};  for t = T-1 .. 1:  lam_out[b,t,:] = lam_t;  lam_{t-1} = F_t^T lam_t + g_{t-1};  g_y0 = lam_0.
constexpr int kLamThreads = 512;

template <int NS>
__global__ void __launch_bounds__(kLamThreads, 1) scan_bwd_tma_lam_kernel(const float* __restrict__ F,
                                                                           const float* __restrict__ g_ys,
                                                                           float* __restrict__ lam_out,
                                                                           float* __restrict__ g_y0, int T, int H, int bs) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int N = H * bs;
  const unsigned step_bytes = (unsigned)N * 4;
  float* tiles = reinterpret_cast<float*>(smraw);  // [NS][N]
  float* lamb = tiles + (size_t)NS * N;            // [2][H]
  float* red = lamb + 2 * H;                       // [RQ][H]
  const int RQ = kLamThreads / H;                  // row groups (H <= 512, H | 512)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + (size_t)RQ * H);
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* Fb = F + (size_t)b * T * N;
  const float* gb = g_ys + (size_t)b * T * H;
  float* lo = lam_out + (size_t)b * T * H;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
    fence_mbar_init();
  }
  for (int h = tid; h < H; h += kLamThreads) lamb[h] = gb[(size_t)(T - 1) * H + h];
  __syncthreads();
  if (tid == 0) {
    for (int u = 0; u < NS - 1 && T - 1 - u >= 1; ++u) {
      mbar_expect_tx(bars + (u % NS), step_bytes);
      tma_load_1d(tiles + (size_t)(u % NS) * N, Fb + (size_t)(T - 1 - u) * N, step_bytes, bars + (u % NS));
    }
  }
  const int h = tid % H, rq = tid / H;
  const int blk = h / bs, jj = h - blk * bs;
  int cur = 0;
  for (int t = T - 1; t >= 1; --t) {
    const int it = T - 1 - t, s = it % NS;
    if (tid == 0) {
      const int tn = t - (NS - 1);
      if (tn >= 1) {
        // the stage being refilled was last read (generic proxy) before the barrier that closed the previous step
        const int sn = (it + NS - 1) % NS;
        fence_async_proxy();
        mbar_expect_tx(bars + sn, step_bytes);
        tma_load_1d(tiles + (size_t)sn * N, Fb + (size_t)tn * N, step_bytes, bars + sn);
      }
    }
    const float g_prev = tid < H ? gb[(size_t)(t - 1) * H + tid] : 0.f;
    const float* lam = lamb + cur * H;
    if (tid < H) lo[(size_t)t * H + tid] = lam[tid];
    mbar_wait(bars + s, (unsigned)((it / NS) & 1));
    const float* col = tiles + (size_t)s * N + (size_t)blk * bs * bs + jj;
    const float* lv = lam + blk * bs;
    float acc = 0.f;
    for (int i = rq; i < bs; i += RQ) acc = fmaf(col[(size_t)i * bs], lv[i], acc);
    red[rq * H + h] = acc;
    __syncthreads();
    if (tid < H) {
      float sum = g_prev;
      for (int q = 0; q < RQ; ++q) sum += red[q * H + tid];
      lamb[(cur ^ 1) * H + tid] = sum;
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int hh = tid; hh < H; hh += kLamThreads) g_y0[(size_t)b * H + hh] = lamb[cur * H + hh];
}

// false if the shape is outside what the kernel handles (same conditions as scan_tma_dispatch in linear_scan.cu)
static inline bool scan_bwd_lam_dispatch(cudaStream_t cs, const float* F, const float* g_ys, float* lam_out, float* g_y0,
                                         int B, int T, int H, int bs) {
  if (bs % 4 != 0 || H > kLamThreads || kLamThreads % H != 0 || H % 4 != 0) return false;
  const size_t step_bytes = (size_t)H * bs * 4;
  if (step_bytes % 16 != 0) return false;
  const size_t extra = (size_t)(2 * H + (kLamThreads / H) * H) * 4 + 64;
  int ns = 0;
  for (int cand = 4; cand >= 2; --cand)
    if (cand * step_bytes + extra <= 220 * 1024) { ns = cand; break; }
  if (ns == 0) return false;
  const size_t smem = ns * step_bytes + extra;
  if (ns == 4) {
    cudaFuncSetAttribute(scan_bwd_tma_lam_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_bwd_tma_lam_kernel<4><<<B, kLamThreads, smem, cs>>>(F, g_ys, lam_out, g_y0, T, H, bs);
  } else if (ns == 3) {
    cudaFuncSetAttribute(scan_bwd_tma_lam_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_bwd_tma_lam_kernel<3><<<B, kLamThreads, smem, cs>>>(F, g_ys, lam_out, g_y0, T, H, bs);
  } else {
    cudaFuncSetAttribute(scan_bwd_tma_lam_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_bwd_tma_lam_kernel<2><<<B, kLamThreads, smem, cs>>>(F, g_ys, lam_out, g_y0, T, H, bs);
  }
  return true;
}

}  // namespace lncde
