// Host-emulator counterpart of log-neural-cdes_b200/csrc/ptx.cuh (same names, same contracts).
// Asynchronous copies complete at issue or, in the emulator's lazy mode, at the consumer's wait; the mbarrier bookkeeping (expect_tx / complete_tx / phase
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

inline void cp_async4(float* smem_dst, const float* gsrc) { emu::cp_async(smem_dst, gsrc, 4); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() { emu::cp_async_wait(); }

inline void mbar_init(unsigned long long* bar, unsigned count) { emu::mbar_init(bar, count); }
inline void fence_mbar_init() {}
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu::mbar_expect_tx(bar, bytes); }
inline void mbar_wait(unsigned long long* bar, unsigned parity) { emu::mbar_wait(bar, parity); }
inline void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  emu::check_bulk(smem_dst, gsrc, bytes, "cp.async.bulk global->shared");
  emu::async_load(smem_dst, gsrc, bytes, bar);
}
inline void tma_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
  emu::check_bulk(smem_src, gdst, bytes, "cp.async.bulk shared->global");
  emu::async_store(gdst, smem_src, bytes);
}
inline void tma_store_commit() {}
inline void tma_store_wait_read() { emu::async_store_wait(); }
inline void fence_async_proxy() {}

// clusters: the emulator runs a cluster of ONE CTA (csrc/linear_cluster.cuh launches CS = 1 under LNCDE_EMU), so the only
// peer is the CTA itself; the byte counting on the mbarrier is the real protocol
inline unsigned cluster_ctarank() { return 0; }
inline void cluster_sync_all() { emu::sync_threads(); }
inline void st_async_f32(float* dst, float v, unsigned long long* bar, unsigned rank) {
  if (rank != 0 || !emu::in_dyn_smem(dst, 4)) {
    std::fprintf(stderr, "cuda-emu: st.async to a peer that does not exist / outside shared memory\n");
    std::abort();
  }
  *dst = v;
  emu::count_peer_store();
  emu::mbar_complete_tx(bar, 4);
}

// TF32: keep 10 mantissa bits, round to nearest with ties away from zero (cvt.rna)
inline uint32_t cvt_tf32(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
  return u & 0xffffe000u;
}
// warp-collective model of mma.sync.m16n8k8 (tf32 x tf32 -> fp32) with the fragment layout documented in csrc/ptx.cuh:
// every lane deposits its fragments, then forms its four outputs from the gathered 16x8 / 8x8 operand tiles
inline void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  uint64_t fa[2][32], fb[32];
  const uint64_t* s = emu::warp_exchange(0xffffffffu, (uint64_t)a[0] | ((uint64_t)a[1] << 32));
  std::memcpy(fa[0], s, sizeof(fa[0]));
  s = emu::warp_exchange(0xffffffffu, (uint64_t)a[2] | ((uint64_t)a[3] << 32));
  std::memcpy(fa[1], s, sizeof(fa[1]));
  s = emu::warp_exchange(0xffffffffu, (uint64_t)b[0] | ((uint64_t)b[1] << 32));
  std::memcpy(fb, s, sizeof(fb));
  auto as_f = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
  auto A = [&](int m, int k) {  // owner lane (m % 8) * 4 + k % 4, register (m >= 8) + 2 * (k >= 4)
    const uint64_t w = fa[k >= 4][(m & 7) * 4 + (k & 3)];
    return as_f((uint32_t)(m >= 8 ? (w >> 32) : w));
  };
  auto Bm = [&](int k, int n) {  // owner lane n * 4 + k % 4, register (k >= 4)
    const uint64_t w = fb[n * 4 + (k & 3)];
    return as_f((uint32_t)(k >= 4 ? (w >> 32) : w));
  };
  const int lane = emu::lane_id(), g = lane >> 2, t = lane & 3;
  for (int i = 0; i < 4; ++i) {
    const int m = g + (i >= 2 ? 8 : 0), n = 2 * t + (i & 1);
    double acc = d[i];
    for (int k = 0; k < 8; ++k) acc += (double)A(m, k) * (double)Bm(k, n);
    d[i] = (float)acc;
  }
}

}  // namespace lncde
