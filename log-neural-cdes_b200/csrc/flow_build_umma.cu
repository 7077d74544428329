// K2 dense (DE-LNCDE) flow build on the 5th-generation tensor cores ("k2_de_v2" = 3; written blind, opt-in).
//
//   F[m, n] = I[n] + sum_l logsig[m, 1 + l] * lie[n, l]       m = (sample, step), n = (block, i, j)
//   (LogLinearCDE.__call__ block branch, models/LinearNeuralCDEs.py:318-329: flows = I + einsum(lie_brackets, logsigs[:, 1:]))
//
// is a [B T, n_basis] x [n_basis, H bs] GEMM: 44 GFLOP at depth 2 and 220 GFLOP at depth 3 for the EigenWorms shape
// against 3.1 GB of flows written, i.e. 14 / 70 flop per byte - compute-bound on the FP32 pipe (3.8 ms measured for the
// FFMA kernel at depth 2), output-bound on the tensor pipe.  fp32 accuracy is required (1e-5 parity; plain TF32 / BF16
// products are 5e-4 / 4e-3), so the contraction runs as the 9xBF16 split scheme of CUTLASS' sm_100 "FastFP32"
// collective: each fp32 operand is split into three BF16 terms in shared memory by transform warps, the nine partial
// products are issued as tcgen05.mma (SASS UTCHMMA) with fp32 accumulators in TMEM, operands arrive by TMA
// (UTMALDG), and the accumulators are read back with tcgen05.ld for the epilogue.  The collective and kernel layer
// are CUTLASS templates (header tree vendored with the image); the operand packing below is ours:
//
//   A'[m, :] = (logsig[m, 1 .. n_basis], 1, 0 ...)     K' = n_basis + 1 rounded up to 16 (the K tile)
//   B'[n, :] = (lie[n, 0 .. n_basis - 1], I[n], 0 ...)  -> the identity rides along as one more K column, and both
//   operands get the 16-byte row pitch TMA needs (logsig rows are D = 29 floats: 116 bytes).
//
// The bracket gradient of the default reverse sweep, dLie[n, l] = sum_m dF[m, n] logsig[m, 1 + l] (220 GFLOP at depth 3,
// where n_basis = 140 is beyond the triple-product kernels), runs through the same collective with both operands
// MN-major: A = dF viewed as [N x B T] (n contiguous), B = the packed logsig rows [B T x K'] (l contiguous), fp32
// accumulation over all B T rows in TMEM, output [N x K'] into the (then free) packed-lie scratch, copied to g_lie.
//
// Under the host emulator (LNCDE_EMU) and where the CUTLASS headers are missing (LNCDE_NO_CUTLASS) the GEMM itself is
// not available: the emulator runs the packing kernels and a plain host-order fp32 contraction of the PACKED
// operands (which checks the packing, the identity column and the workspace layout); without CUTLASS the entry point
// returns LNCDE_ERR_UNSUPPORTED and the caller keeps the FFMA / mma.sync kernels.
#include <cuda_runtime.h>

#include <cstdint>

#include "lncde_b200.h"

#if !defined(LNCDE_EMU) && !defined(LNCDE_NO_CUTLASS)
#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#endif

namespace lncde {

constexpr int kUmmaKTile = 16;                       // K tile of the collective (floats)
constexpr size_t kUmmaGemmWorkspaceBytes = 1 << 20;  // reserved for the CUTLASS kernel's own workspace

__host__ __device__ inline int umma_kpad(int n_basis) { return (n_basis + 1 + kUmmaKTile - 1) / kUmmaKTile * kUmmaKTile; }

// floats of scratch behind the flows: A' [Mtot][Kp] | B' [N][Kp] | CUTLASS workspace
size_t flow_build_umma_scratch_floats(long long Mtot, int N, int n_basis) {
  const size_t kp = (size_t)umma_kpad(n_basis);
  return (size_t)Mtot * kp + (size_t)N * kp + kUmmaGemmWorkspaceBytes / 4;
}

// one thread per float4 of the packed row: consecutive threads write consecutive 16-byte chunks
__global__ void __launch_bounds__(256) umma_pack_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              long long rows, int src_pitch, int src_off, int n_basis,
                                                              int kp, int bs) {
  const int chunks = kp >> 2;
  const long long total = rows * chunks;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / chunks;
    const int k0 = (int)(e - r * chunks) * 4;
    // the extra K column: 1 for the logsig rows (bs == 0), the flattened identity of the block for the lie rows
    float one = 1.f;
    if (bs > 0) {
      const int q = (int)(r % ((long long)bs * bs));
      one = (q / bs == q % bs) ? 1.f : 0.f;
    }
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      v[j] = k < n_basis ? __ldg(src + r * src_pitch + src_off + k) : (k == n_basis ? one : 0.f);
    }
    *reinterpret_cast<float4*>(dst + r * kp + k0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// g_lie[n][l] = Dp[n][l] for l < n_basis (the padded columns of the GEMM output are dropped)
__global__ void __launch_bounds__(256) umma_unpad_kernel(const float* __restrict__ Dp, float* __restrict__ g_lie, int N,
                                                          int n_basis, int kp) {
  const long long total = (long long)N * n_basis;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / n_basis;
    const int l = (int)(e - n * n_basis);
    g_lie[e] = Dp[n * kp + l];
  }
}

#if defined(LNCDE_EMU)
// emulator stand-in for the bracket-gradient GEMM: Dp[n][l] = sum_m dF[m][n] A'[m][l], m ascending
__global__ void __launch_bounds__(256) umma_reference_dlie_kernel(const float* __restrict__ dF, const float* __restrict__ A,
                                                                   float* __restrict__ Dp, long long Mtot, int N, int kp) {
  const long long total = (long long)N * kp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / kp;
    const int l = (int)(e - n * kp);
    float acc = 0.f;
    for (long long m = 0; m < Mtot; ++m) acc = fmaf(dF[m * N + n], A[m * kp + l], acc);
    Dp[e] = acc;
  }
}

// emulator stand-in for the tensor-pipe GEMM: F = A' B'^T in fp32, k ascending (see the header comment)
__global__ void __launch_bounds__(256) umma_reference_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                   float* __restrict__ F, long long Mtot, int N, int kp) {
  const long long total = Mtot * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long m = e / N;
    const int n = (int)(e - m * N);
    float acc = 0.f;
    for (int k = 0; k < kp; ++k) acc = fmaf(A[m * kp + k], Bm[(size_t)n * kp + k], acc);
    F[e] = acc;
  }
}
#elif !defined(LNCDE_NO_CUTLASS)
namespace umma {
using namespace cute;
using ElementA = float;
using ElementB = float;
using ElementC = float;
using Accumulator = float;
using LayoutA = cutlass::layout::RowMajor;     // A'[m][k], k contiguous
using LayoutB = cutlass::layout::ColumnMajor;  // B'[n][k], k contiguous
using LayoutC = cutlass::layout::RowMajor;     // F[m][n]
constexpr int kAlign = 4;                      // 16 bytes
using TileShape = Shape<_128, _128, Int<kUmmaKTile>>;
using ClusterShape = Shape<_1, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, TileShape, ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, Accumulator, Accumulator, ElementC, LayoutC, kAlign, ElementC,
    LayoutC, kAlign, cutlass::epilogue::FastF32NoSmemWarpSpecialized1Sm>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlign, ElementB, LayoutB, kAlign,
    Accumulator, TileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

// bracket gradient: both operands MN-major (dF: n contiguous; packed logsig rows: l contiguous), reduction over B T
using LayoutAg = cutlass::layout::ColumnMajor;
using LayoutBg = cutlass::layout::RowMajor;
using MainloopG = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutAg, kAlign, ElementB, LayoutBg, kAlign,
    Accumulator, TileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>::CollectiveOp;
using GemmKernelG = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, MainloopG, CollectiveEpilogue, void>;
using GemmG = cutlass::gemm::device::GemmUniversalAdapter<GemmKernelG>;
}  // namespace umma
#endif

// F[Mtot][N] (fully overwritten) from lie [N][n_basis] and logsig [Mtot][D]; scratch as sized above, 16-byte aligned.
int flow_build_umma(cudaStream_t cs, const float* lie, const float* logsig, float* F, long long Mtot, int N, int D,
                    int n_basis, int bs, float* scratch, size_t scratch_floats) {
#if !defined(LNCDE_EMU) && defined(LNCDE_NO_CUTLASS)
  (void)cs; (void)lie; (void)logsig; (void)F; (void)Mtot; (void)N; (void)D; (void)n_basis; (void)bs; (void)scratch;
  (void)scratch_floats;
  return LNCDE_ERR_UNSUPPORTED;
#else
  if (Mtot < 1 || N < 1 || n_basis < 1 || bs < 1 || Mtot > 0x7fffffffLL) return LNCDE_ERR_SHAPE;
  if ((N & 3) != 0) return LNCDE_ERR_UNSUPPORTED;  // rows of F must keep the 16-byte alignment of the epilogue stores
  if (((uintptr_t)scratch & 15) != 0 || ((uintptr_t)F & 15) != 0) return LNCDE_ERR_ALIGN;
  if (scratch_floats < flow_build_umma_scratch_floats(Mtot, N, n_basis)) return LNCDE_ERR_WORKSPACE;
  const int kp = umma_kpad(n_basis);
  float* Ap = scratch;
  float* Bp = Ap + (size_t)Mtot * kp;
  void* gemm_ws = Bp + (size_t)N * kp;
  const int grid = 148 * 8;
  umma_pack_rows_kernel<<<grid, 256, 0, cs>>>(logsig, Ap, Mtot, D, 1, n_basis, kp, 0);
  umma_pack_rows_kernel<<<grid, 256, 0, cs>>>(lie, Bp, (long long)N, n_basis, 0, n_basis, kp, bs);
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
#if defined(LNCDE_EMU)
  (void)gemm_ws;
  umma_reference_gemm_kernel<<<grid, 256, 0, cs>>>(Ap, Bp, F, Mtot, N, kp);
  return (int)cudaGetLastError();
#else
  using namespace umma;
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  const int M = (int)Mtot;
  StrideA sA = cute::make_stride((int64_t)kp, cute::Int<1>{}, (int64_t)0);
  StrideB sB = cute::make_stride((int64_t)kp, cute::Int<1>{}, (int64_t)0);
  StrideC sC = cute::make_stride((int64_t)N, cute::Int<1>{}, (int64_t)0);
  StrideD sD = cute::make_stride((int64_t)N, cute::Int<1>{}, (int64_t)0);
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, kp, 1}, {Ap, sA, Bp, sB},
                                {{1.f, 0.f}, F, sC, F, sD}};
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return LNCDE_ERR_UNSUPPORTED;
  if (Gemm::get_workspace_size(args) > kUmmaGemmWorkspaceBytes) return LNCDE_ERR_WORKSPACE;
  if (gemm.initialize(args, gemm_ws, cs) != cutlass::Status::kSuccess) return LNCDE_ERR_UNSUPPORTED;
  if (gemm.run(cs) != cutlass::Status::kSuccess) {
    st = (int)cudaGetLastError();
    return st != 0 ? st : LNCDE_ERR_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
#endif
#endif
}

// g_lie[N][n_basis] (fully overwritten) = dF^T logsig[:, 1:] with dF [Mtot][N] (the in-place result of the default reverse
// sweep) and logsig [Mtot][D]; same scratch as flow_build_umma (the packed-lie region receives the padded output).
int dlie_umma(cudaStream_t cs, const float* dF, const float* logsig, float* g_lie, long long Mtot, int N, int D,
              int n_basis, float* scratch, size_t scratch_floats) {
#if !defined(LNCDE_EMU) && defined(LNCDE_NO_CUTLASS)
  (void)cs; (void)dF; (void)logsig; (void)g_lie; (void)Mtot; (void)N; (void)D; (void)n_basis; (void)scratch;
  (void)scratch_floats;
  return LNCDE_ERR_UNSUPPORTED;
#else
  if (Mtot < 1 || N < 1 || n_basis < 1 || Mtot > 0x7fffffffLL) return LNCDE_ERR_SHAPE;
  if ((N & 3) != 0) return LNCDE_ERR_UNSUPPORTED;
  if (((uintptr_t)scratch & 15) != 0 || ((uintptr_t)dF & 15) != 0) return LNCDE_ERR_ALIGN;
  if (scratch_floats < flow_build_umma_scratch_floats(Mtot, N, n_basis)) return LNCDE_ERR_WORKSPACE;
  const int kp = umma_kpad(n_basis);
  float* Ap = scratch;                 // packed logsig rows [Mtot][kp] (re-packed: the caller may have changed nothing, but
  float* Dp = Ap + (size_t)Mtot * kp;  // the call must not depend on what an earlier call left here); output [N][kp]
  void* gemm_ws = Dp + (size_t)N * kp;
  const int grid = 148 * 8;
  umma_pack_rows_kernel<<<grid, 256, 0, cs>>>(logsig, Ap, Mtot, D, 1, n_basis, kp, 0);
  int st = (int)cudaGetLastError();
  if (st != 0) return st;
#if defined(LNCDE_EMU)
  (void)gemm_ws;
  umma_reference_dlie_kernel<<<grid, 256, 0, cs>>>(dF, Ap, Dp, Mtot, N, kp);
#else
  using namespace umma;
  using StrideA = typename GemmG::GemmKernel::StrideA;
  using StrideB = typename GemmG::GemmKernel::StrideB;
  using StrideC = typename GemmG::GemmKernel::StrideC;
  using StrideD = typename GemmG::GemmKernel::StrideD;
  StrideA sA = cute::make_stride(cute::Int<1>{}, (int64_t)N, (int64_t)0);   // A(n, m) = dF[m * N + n]
  StrideB sB = cute::make_stride(cute::Int<1>{}, (int64_t)kp, (int64_t)0);  // B(l, m) = Ap[m * kp + l]
  StrideC sC = cute::make_stride((int64_t)kp, cute::Int<1>{}, (int64_t)0);
  StrideD sD = cute::make_stride((int64_t)kp, cute::Int<1>{}, (int64_t)0);
  typename GemmG::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {N, kp, (int)Mtot, 1}, {dF, sA, Ap, sB},
                                 {{1.f, 0.f}, Dp, sC, Dp, sD}};
  GemmG gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return LNCDE_ERR_UNSUPPORTED;
  if (GemmG::get_workspace_size(args) > kUmmaGemmWorkspaceBytes) return LNCDE_ERR_WORKSPACE;
  if (gemm.initialize(args, gemm_ws, cs) != cutlass::Status::kSuccess) return LNCDE_ERR_UNSUPPORTED;
  if (gemm.run(cs) != cutlass::Status::kSuccess) {
    st = (int)cudaGetLastError();
    return st != 0 ? st : LNCDE_ERR_UNSUPPORTED;
  }
#endif
  st = (int)cudaGetLastError();
  if (st != 0) return st;
  umma_unpad_kernel<<<grid, 256, 0, cs>>>(Dp, g_lie, N, n_basis, kp);
  return (int)cudaGetLastError();
#endif
}

}  // namespace lncde
