// K2 dense (DE-LNCDE) flow build on the 5th-generation tensor cores - hand-written tcgen05 / TMEM / TMA kernel
// (tuning "k2_de_v2" = 3).  No library code: the PTX wrappers are in tcgen05.cuh.
//
//   F[m, n] = I[n] + sum_l logsig[m, 1 + l] * lie[n, l]       m = (sample, step), n = (block, i, j)
//   (LogLinearCDE.__call__ block branch, models/LinearNeuralCDEs.py:318-329: flows = I + einsum(lie_brackets, logsigs[:, 1:]))
//
// a [B T, n_basis] x [n_basis, nb bs bs] GEMM: 44 GFLOP at depth 2 / 220 GFLOP at depth 3 for the EigenWorms shape
// against 3.1 GB of flows written (14 / 70 flop per byte): FP32-issue bound as an FFMA kernel (3.8 ms measured at
// depth 2), bound by WRITING the flows on the tensor pipe.  The parity bar is 1e-5 in fp32 - a plain TF32 product
// is 5e-4 - so the contraction runs as 3xTF32: every operand x is split once, on the way into the packed operand
// buffers, into hi = tf32(x), lo = tf32(x - hi) (both exactly representable), and a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi
// with fp32 accumulation in TMEM (error ~2^-21 |a b|, the level of the fp32 rounding of the FFMA kernels).
//
// Data movement.  de_pack_operand_kernel writes both operands in the shared-memory image tcgen05.mma reads - K-major,
// no swizzle: per (128-row tile, 32-wide K stage, hi | lo plane) 8 K-chunks x 128 rows x 16 bytes = 16 KB, hi and lo
// planes adjacent - so one stage of one operand tile is ONE 32 KB 1-D bulk copy (cp.async.bulk, SASS UBLKCP; no tensor
// map needed).  The packed operands are 12 MB (logsig) + 4 MB (lie) at depth 2 and stay in L2.
//
// Kernel (persistent, one CTA per SM, 6 warps):
//   warp 0   producer: per stage wait empty[s], expect_tx 64 KB on full[s], two bulk copies (A tile, B tile); 2 stages
//   warp 1   TMEM owner + MMA issuer: per stage 4 K-steps x 3 tcgen05.mma (kind::tf32, M = N = 128, K = 8),
//            tcgen05.commit -> empty[s]; after the tile's last stage tcgen05.commit -> tmem_full[a]; 2 accumulators
//            (2 x 128 TMEM columns), so the MMAs of tile i + 1 overlap the epilogue of tile i
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> + identity -> padded row in shared memory
//            (pitch 528 B: conflict-free 128-bit stores) -> arrive on tmem_empty[a] -> every thread stores ITS OWN
//            output row segment (512 B, contiguous in F) with one bulk store (cp.async.bulk.global.shared::cta): full
//            128-byte lines leave the SM, no partially written sectors.
// Every mbarrier wait is bounded (tcgen05.cuh: mbar_wait_bounded): a protocol error ends the kernel with a code in
// *err instead of hanging the device; lncde_linear_scan_fwd's caller can read it back from the scratch tail.
//
// Under the host emulator (LNCDE_EMU) the tensor pipe does not exist: the emulated build runs the packing kernel
// and a plain fp32 evaluation of the same three partial products from the PACKED operands (which checks the operand
// image, the K padding, the identity and the workspace layout - everything but the MMA unit itself).
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"
#include "ptx.cuh"
#if !defined(LNCDE_EMU)
#include "tcgen05.cuh"
#endif

namespace lncde {
namespace tc5 {

constexpr int kTile = 128;                       // rows of an operand tile = M and N of one MMA
constexpr int kStageK = 32;                      // K floats per pipeline stage (4 MMAs of K = 8)
constexpr int kChunks = kStageK / 4;             // 16-byte K chunks per stage
constexpr int kPlaneFloats = kChunks * kTile * 4;  // one plane (hi or lo) of one stage of one tile: 16 KB
constexpr int kStageFloats = 2 * kPlaneFloats;     // hi | lo: 32 KB
constexpr int kStages = 2;
constexpr int kRowPitch = kTile + 4;             // floats per staged output row (528 B: conflict-free v4 stores)
constexpr int kThreads = 192;
constexpr size_t kSmemOperands = (size_t)kStages * 2 * kStageFloats * 4;  // 128 KB
constexpr size_t kSmemStaging = (size_t)kTile * kRowPitch * 4;             // 66 KB
constexpr size_t kSmemBytes = kSmemOperands + kSmemStaging + 128;          // + barriers, TMEM pointer

__host__ __device__ inline int k_stages_of(int n_basis) { return (n_basis + kStageK - 1) / kStageK; }
__host__ __device__ inline long long tiles_of(long long rows) { return (rows + kTile - 1) / kTile; }

// packed operand: [tile][k stage][plane][chunk][row][4]
__host__ __device__ inline size_t packed_floats(long long rows, int n_basis) {
  return (size_t)tiles_of(rows) * k_stages_of(n_basis) * kStageFloats;
}

// round-to-nearest TF32 split of one fp32 value: x = hi + lo + O(2^-22 |x|), hi and lo exactly representable in TF32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(cvt_tf32(x));
  lo = __uint_as_float(cvt_tf32(x - hi));
}

// one thread per (tile, stage, chunk, row): reads 4 consecutive K values of a source row, writes the hi and lo chunks
__global__ void __launch_bounds__(256) de_pack_operand_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                               long long rows, int src_pitch, int src_off, int n_basis) {
  const int nks = k_stages_of(n_basis);
  const long long total = tiles_of(rows) * nks * kChunks * kTile;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e % kTile);
    const int chunk = (int)((e / kTile) % kChunks);
    const long long ts = e / (kTile * kChunks);  // tile * nks + stage
    const int ks = (int)(ts % nks);
    const long long r = (ts / nks) * kTile + row;
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = ks * kStageK + chunk * 4 + j;
      const float x = (r < rows && k < n_basis) ? __ldg(src + r * src_pitch + src_off + k) : 0.f;
      split_tf32(x, hi[j], lo[j]);
    }
    float* p = dst + (size_t)ts * kStageFloats + ((size_t)chunk * kTile + row) * 4;
    *reinterpret_cast<float4*>(p) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(p + kPlaneFloats) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// 1 on the diagonal of the (bs x bs) block that flat column n belongs to
__device__ __forceinline__ float identity_at(long long n, int bs) {
  const int q = (int)(n % ((long long)bs * bs));
  return (q / bs == q % bs) ? 1.f : 0.f;
}

#if defined(LNCDE_EMU)
// emulator stand-in for the tensor pipe: the same three partial products, read from the PACKED operand images
__global__ void __launch_bounds__(256) flow_tc5_reference_kernel(const float* __restrict__ Ap, const float* __restrict__ Bp,
                                                                  float* __restrict__ F, long long Mtot, int N, int n_basis,
                                                                  int bs) {
  const int nks = k_stages_of(n_basis);
  const long long total = Mtot * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long m = e / N;
    const int n = (int)(e - m * N);
    const float* a = Ap + (size_t)(m / kTile) * nks * kStageFloats;
    const float* b = Bp + (size_t)(n / kTile) * nks * kStageFloats;
    const int ra = (int)(m % kTile), rb = n % kTile;
    float acc = 0.f;
    for (int ks = 0; ks < nks; ++ks)
      for (int c = 0; c < kChunks; ++c)
        for (int j = 0; j < 4; ++j) {
          const size_t ia = (size_t)ks * kStageFloats + ((size_t)c * kTile + ra) * 4 + j;
          const size_t ib = (size_t)ks * kStageFloats + ((size_t)c * kTile + rb) * 4 + j;
          acc += a[ia + kPlaneFloats] * b[ib];
          acc += a[ia] * b[ib + kPlaneFloats];
          acc += a[ia] * b[ib];
        }
    F[e] = acc + identity_at(n, bs);
  }
}
#else

struct FlowParams {
  const float* Ap;  // packed logsig rows
  const float* Bp;  // packed bracket matrices
  float* F;
  long long Mtot;
  int N, nks, bs;
  int m_tiles, n_tiles;
  int* err;
};

__global__ void __launch_bounds__(kThreads, 1) flow_build_tcgen05_kernel(const FlowParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* opnd = reinterpret_cast<float*>(smem_raw);                                   // [stage][A | B][hi | lo]
  float* stg = reinterpret_cast<float*>(smem_raw + kSmemOperands);                    // [128][kRowPitch]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kSmemOperands + kSmemStaging);
  unsigned long long* full = bars;        // [kStages]  producer -> MMA
  unsigned long long* empty = bars + 2;   // [kStages]  MMA -> producer
  unsigned long long* tfull = bars + 4;   // [2]        MMA -> epilogue
  unsigned long long* tempty = bars + 6;  // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull + a, 1);
      mbar_init(tempty + a, 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const long long n_tiles_total = (long long)p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- producer
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (long long t = blockIdx.x; ok && t < n_tiles_total; t += gridDim.x) {
        const long long mt = t / p.n_tiles, nt = t % p.n_tiles;
        for (int ks = 0; ks < p.nks; ++ks) {
          if (!(ok = mbar_wait_bounded(empty + s, ph ^ 1u, p.err, 1))) break;
          float* dstA = opnd + (size_t)s * 2 * kStageFloats;
          mbar_expect_tx(full + s, 2u * kStageFloats * 4u);
          tma_load_1d(dstA, p.Ap + ((size_t)mt * p.nks + ks) * kStageFloats, kStageFloats * 4u, full + s);
          tma_load_1d(dstA + kStageFloats, p.Bp + ((size_t)nt * p.nks + ks) * kStageFloats, kStageFloats * 4u, full + s);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = instr_desc_tf32(kTile, kTile);
      constexpr uint32_t kLbo = kTile * 16, kSbo = 128;  // next K chunk: 128 rows x 16 B; next 8 rows: 128 B
      int s = 0;
      uint32_t ph = 0;
      long long it = 0;
      bool ok = true;
      for (long long t = blockIdx.x; ok && t < n_tiles_total; t += gridDim.x, ++it) {
        const int a = (int)(it & 1);
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        if (!(ok = mbar_wait_bounded(tempty + a, aph ^ 1u, p.err, 2))) break;
        fence_after_thread_sync();
        const uint32_t d = tmem + (uint32_t)a * kTile;
        for (int ks = 0; ks < p.nks; ++ks) {
          if (!(ok = mbar_wait_bounded(full + s, ph, p.err, 3))) break;
          fence_after_thread_sync();
          const uint32_t a_hi = smem_u32(opnd + (size_t)s * 2 * kStageFloats);
          const uint32_t a_lo = a_hi + kPlaneFloats * 4, b_hi = a_hi + kStageFloats * 4, b_lo = b_hi + kPlaneFloats * 4;
#pragma unroll
          for (int kk = 0; kk < kStageK / 8; ++kk) {
            const uint32_t off = (uint32_t)kk * 2u * kLbo;  // two K chunks per MMA
            const uint64_t dah = smem_desc_kmajor_noswizzle(a_hi + off, kLbo, kSbo);
            const uint64_t dal = smem_desc_kmajor_noswizzle(a_lo + off, kLbo, kSbo);
            const uint64_t dbh = smem_desc_kmajor_noswizzle(b_hi + off, kLbo, kSbo);
            const uint64_t dbl = smem_desc_kmajor_noswizzle(b_lo + off, kLbo, kSbo);
            mma_tf32_ss(d, dal, dbh, idesc, (ks | kk) != 0);  // small terms first
            mma_tf32_ss(d, dah, dbl, idesc, 1);
            mma_tf32_ss(d, dah, dbh, idesc, 1);
          }
          mma_commit(empty + s);  // the stage may be refilled once these MMAs have read it
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        if (ok) mma_commit(tfull + a);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;            // the TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;     // row of the tile = TMEM lane
    float* my = stg + (size_t)row * kRowPitch;
    long long it = 0;
    bool ok = true;
    for (long long t = blockIdx.x; ok && t < n_tiles_total; t += gridDim.x, ++it) {
      const long long mt = t / p.n_tiles, nt = t % p.n_tiles;
      const int a = (int)(it & 1);
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      ok = mbar_wait_bounded(tfull + a, aph, p.err, 4);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      fence_after_thread_sync();
      tma_store_wait_read();  // my bulk store of the previous tile has finished reading my staging row
      const long long n0 = nt * kTile;
      // column of this tile that carries the identity for my row's block (bs % 128 == 0: at most one per row)
      int id_col = -1;
      {
        const int q = (int)(n0 % ((long long)p.bs * p.bs));
        const int i = q / p.bs, j0 = q % p.bs;
        if (i >= j0 && i < j0 + kTile) id_col = i - j0;
      }
#pragma unroll
      for (int c = 0; c < kTile / 32; ++c) {
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * kTile + c * 32), v);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] += (c * 32 + k == id_col) ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(my + c * 32 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
      fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + a);  // the accumulator may be overwritten
      fence_async_proxy();                     // my generic-proxy writes -> visible to the bulk-copy unit
      const long long m = mt * kTile + row;
      if (m < p.Mtot) tma_store_1d(p.F + m * p.N + n0, my, kTile * 4u);
      tma_store_commit();
    }
    tma_store_wait_all();
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_thread_sync();
    tmem_dealloc(tmem, 256);
  }
}
#endif

}  // namespace tc5

// floats of scratch behind the flows: packed logsig rows | packed bracket matrices | error word (padded to 4 floats)
size_t flow_tcgen05_scratch_floats(long long Mtot, int N, int n_basis) {
  return tc5::packed_floats(Mtot, n_basis) + tc5::packed_floats(N, n_basis) + 4;
}

// which shapes the kernel takes: whole 128-column tiles that sit inside one row of a (bs x bs) block
bool flow_tcgen05_shape(int N, int bs) { return bs >= tc5::kTile && bs % tc5::kTile == 0 && N % tc5::kTile == 0; }

// F[Mtot][N] (fully overwritten) from lie [N][n_basis] and logsig [Mtot][D]; scratch as sized above, 16-byte aligned.
int flow_build_tcgen05(cudaStream_t cs, const float* lie, const float* logsig, float* F, long long Mtot, int N, int D,
                       int n_basis, int bs, float* scratch, size_t scratch_floats) {
  using namespace tc5;
  if (Mtot < 1 || N < 1 || n_basis < 1 || bs < 1 || Mtot > 0x7fffffffLL) return LNCDE_ERR_SHAPE;
  if (!flow_tcgen05_shape(N, bs)) return LNCDE_ERR_UNSUPPORTED;
  if (((uintptr_t)scratch & 15) != 0 || ((uintptr_t)F & 15) != 0) return LNCDE_ERR_ALIGN;
  if (scratch_floats < flow_tcgen05_scratch_floats(Mtot, N, n_basis)) return LNCDE_ERR_WORKSPACE;
  float* Ap = scratch;
  float* Bp = Ap + packed_floats(Mtot, n_basis);
  int* err = reinterpret_cast<int*>(Bp + packed_floats(N, n_basis));
  const int grid = 148 * 8;
  de_pack_operand_kernel<<<grid, 256, 0, cs>>>(logsig, Ap, Mtot, D, 1, n_basis);
  de_pack_operand_kernel<<<grid, 256, 0, cs>>>(lie, Bp, (long long)N, n_basis, 0, n_basis);
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
#if defined(LNCDE_EMU)
  (void)err;
  flow_tc5_reference_kernel<<<grid, 256, 0, cs>>>(Ap, Bp, F, Mtot, N, n_basis, bs);
  return (int)cudaGetLastError();
#else
  st = (int)cudaMemsetAsync(err, 0, sizeof(int), cs);
  if (st != 0) return st;
  FlowParams p;
  p.Ap = Ap; p.Bp = Bp; p.F = F; p.Mtot = Mtot; p.N = N; p.nks = k_stages_of(n_basis); p.bs = bs;
  p.m_tiles = (int)tiles_of(Mtot); p.n_tiles = N / kTile; p.err = err;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (long long)p.m_tiles * p.n_tiles;
  const int ctas = (int)(tiles < sms ? tiles : sms);
  cudaFuncSetAttribute(flow_build_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  flow_build_tcgen05_kernel<<<ctas, kThreads, kSmemBytes, cs>>>(p);
  return (int)cudaGetLastError();
#endif
}

}  // namespace lncde
