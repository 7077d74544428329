// Host-emulator counterpart of log-neural-cdes_b200/csrc/ptx.cuh (same names, same contracts).
// Asynchronous copies complete at issue; the mbarrier bookkeeping (expect_tx / complete_tx / phase
// parity) and the alignment + size rules of the bulk copies are modelled and checked.
#pragma once
#include <cstdint>
#include <cstring>

namespace lncde {

inline float ld_stream(const float* p) { return *p; }
inline float4 ld_stream4(const float4* p) {
  if ((uintptr_t)p & 15u) {
    std::fprintf(stderr, "cuda-emu: misaligned 128-bit global load\n");
    std::abort();
  }
  return *p;
}

inline void cp_async4(float* smem_dst, const float* gsrc) { *smem_dst = *gsrc; }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}

inline void mbar_init(unsigned long long* bar, unsigned count) { emu::mbar_init(bar, count); }
inline void fence_mbar_init() {}
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu::mbar_expect_tx(bar, bytes); }
inline void mbar_wait(unsigned long long* bar, unsigned parity) { emu::mbar_wait(bar, parity); }
inline void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  emu::check_bulk(smem_dst, gsrc, bytes, "cp.async.bulk global->shared");
  std::memcpy(smem_dst, gsrc, bytes);
  emu::mbar_complete_tx(bar, bytes);
}
inline void tma_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
  emu::check_bulk(smem_src, gdst, bytes, "cp.async.bulk shared->global");
  std::memcpy(gdst, smem_src, bytes);
}
inline void tma_store_commit() {}
inline void tma_store_wait_read() {}
inline void fence_async_proxy() {}

}  // namespace lncde
