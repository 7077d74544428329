// Inline-PTX wrappers for the 5th-generation tensor cores (tcgen05.* / TMEM) of sm_100a, used by flow_tcgen05.cu.
// Not visible to the host emulator (tests/emu cannot model TMEM or the MMA unit): include only under !LNCDE_EMU.
//
// Operand descriptors (64-bit shared-memory matrix descriptor, PTX ISA "tcgen05 matrix descriptor"):
//   [ 0,14) start address >> 4      [16,30) leading-dimension byte offset >> 4      [32,46) stride-dimension byte offset >> 4
//   [46,48) version = 1 (Blackwell) [49,52) base offset = 0    [52] LBO mode = 0     [61,64) swizzle: 0 = none
// K-major operand without swizzle ("interleaved" canonical layout): a core matrix is 8 rows x 16 bytes stored
// contiguously (row pitch 16 B); LBO = distance between the two core matrices adjacent in K of one MMA (K = 8 tf32 =
// 2 x 16 B), SBO = distance between core matrices adjacent in M / N (next 8 rows).
// Instruction descriptor (32-bit, kind::tf32 / f16):
//   [4,6) D format: 1 = f32   [7,10) A format, [10,13) B format: 2 = tf32   [15] A major, [16] B major: 0 = K-major
//   [17,23) N >> 3            [24,29) M >> 4
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lncde {
namespace tc5 {

__device__ __forceinline__ uint64_t smem_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  return d;
}

__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (one full warp, .sync.aligned) ---------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- MMA: D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread ----------------------------
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on `bar` when they have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns ---------------------
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier pieces not in ptx.cuh ----------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A wait that cannot hang the device: gives up after `limit` clocks, records `code` in *err (global memory) and
// returns false - the role that called it leaves its loop and the kernel drains (flow_tcgen05.cu).
__device__ __forceinline__ bool mbar_wait_bounded(unsigned long long* bar, uint32_t parity, int* err, int code,
                                                  long long limit = 400000000LL) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > limit) {
      atomicExch(err, code);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tc5
}  // namespace lncde
