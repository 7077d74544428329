// Host emulator of the CUDA execution model for the CPU test-suite (TEST INFRASTRUCTURE ONLY).
//
// tests/emu/build_emu.py compiles the UNMODIFIED kernel sources of log-neural-cdes_b200/csrc with g++
// against this header (force-included) into tests/emu/_build/liblncde_emu.so, so that
// `pytest -m "not gpu"` can execute the real kernel code - every thread of every CTA as a
// cooperative fiber, with __syncthreads / warp shuffles / mbarrier + bulk-copy semantics - and
// compare it with the oracle when no GPU is available.  The product package never loads this
// library (lncde/_lib.py only knows liblncde_b200.so and refuses CPU tensors).
//
// What it checks: the arithmetic, indexing, barrier structure (deadlocks and mismatched
// collectives abort), shared-memory bounds (canaries), 16-byte alignment and size rules of the
// TMA bulk copies, the 48 KB opt-in rule and the 227 KB limit of dynamic shared memory.
// What it cannot check: data races between barriers (fibers of a CTA run one after another),
// the accuracy of the fast-math units (__expf etc. map to libm), performance.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <type_traits>

// ---- function / variable space qualifiers ------------------------------------------------
#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __constant__
#undef __launch_bounds__
#undef __forceinline__
#undef __noinline__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __forceinline__ inline
#define __noinline__ /* empty: libstdc++ spells the attribute __attribute__((__noinline__)) */

namespace emu {

struct Idx {
  unsigned x, y, z;
};
extern Idx g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

void* dyn_smem();
void sync_threads();
// warp collective: every lane named in `mask` deposits `v`; returns the 32 deposited values
const uint64_t* warp_exchange(unsigned mask, uint64_t v);
void launch(const void* func, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

template <class F>
inline const void* fptr(F* f) {
  return reinterpret_cast<const void*>(f);
}
// kernel<<<grid, block, smem, stream>>>(args) after the textual rewrite of build_emu.py
inline void launch_k(const void* func, const std::function<void()>& body, dim3 grid, dim3 block, size_t smem = 0,
                     cudaStream_t = 0) {
  launch(func, grid, block, smem, body);
}

// mbarrier / bulk-copy model (used by tests/emu/ptx.cuh)
void mbar_init(void* bar, unsigned count);
void mbar_expect_tx(void* bar, unsigned bytes);
void mbar_complete_tx(void* bar, unsigned bytes);
void mbar_wait(void* bar, unsigned parity);
void check_bulk(const void* smem_ptr, const void* gptr, unsigned bytes, const char* what);
// asynchronous copies: delivered at issue (eager) or at the consumer's wait (lazy), see emu_set_async_mode
void async_load(void* smem_dst, const void* gsrc, unsigned bytes, void* bar);
void async_store(void* gdst, const void* smem_src, unsigned bytes);
void async_store_wait();
void cp_async(void* smem_dst, const void* gsrc, unsigned bytes);
void cp_async_wait();
bool in_dyn_smem(const void* p, size_t bytes);
void count_peer_store();

template <class T>
inline uint64_t pack(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t u = 0;
  std::memcpy(&u, &v, sizeof(T));
  return u;
}
template <class T>
inline T unpack(uint64_t u) {
  T v;
  std::memcpy(&v, &u, sizeof(T));
  return v;
}
inline int lane_id() { return (int)(g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z)) & 31; }

}  // namespace emu

// cuda_runtime.h only declares the typed overload under nvcc
template <class T>
inline cudaError_t cudaFuncSetAttribute(T* entry, cudaFuncAttribute attr, int value) {
  return ::cudaFuncSetAttribute(reinterpret_cast<const void*>(entry), attr, value);
}

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

// ---- synchronisation and warp intrinsics ---------------------------------------------------
inline void __syncthreads() { emu::sync_threads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_exchange(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* s = emu::warp_exchange(mask, emu::pack(v));
  const int base = lane & ~(width - 1);
  return emu::unpack<T>(s[base + (src & (width - 1))]);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* s = emu::warp_exchange(mask, emu::pack(v));
  const int tgt = lane ^ lane_mask;
  return emu::unpack<T>(s[(tgt & ~(width - 1)) == (lane & ~(width - 1)) ? tgt : lane]);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* s = emu::warp_exchange(mask, emu::pack(v));
  const int tgt = lane + (int)delta;
  return emu::unpack<T>(s[tgt < ((lane & ~(width - 1)) + width) ? tgt : lane]);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const uint64_t* s = emu::warp_exchange(mask, emu::pack(v));
  const int tgt = lane - (int)delta;
  return emu::unpack<T>(s[tgt >= (lane & ~(width - 1)) ? tgt : lane]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  const uint64_t* s = emu::warp_exchange(mask, pred ? 1 : 0);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (((mask >> l) & 1u) && s[l]) r |= 1u << l;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
inline unsigned __activemask() { return 0xffffffffu; }

// ---- device math / memory intrinsics --------------------------------------------------------
template <class T>
inline T __ldg(const T* p) {
  return *p;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline int __mulhi(int a, int b) { return (int)(((long long)a * b) >> 32); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return ::expf(x); }
inline float __logf(float x) { return ::logf(x); }
inline float __fmaf_rn(float a, float b, float c) { return ::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100): each half rounds like the scalar instruction
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(::fmaf(a.x, b.x, c.x), ::fmaf(a.y, b.y, c.y)); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float __int_as_float(int i) { return emu::unpack<float>(emu::pack(i)); }
inline int __float_as_int(float f) { return emu::unpack<int>(emu::pack(f)); }
inline float __uint_as_float(unsigned u) { return emu::unpack<float>(emu::pack(u)); }
inline unsigned __float_as_uint(float f) { return emu::unpack<unsigned>(emu::pack(f)); }
inline long long clock64() { return 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline float rsqrtf(float x) { return 1.0f / ::sqrtf(x); }

inline float atomicAdd(float* p, float v) {
  const float o = *p;
  *p = o + v;
  return o;
}
inline int atomicAdd(int* p, int v) {
  const int o = *p;
  *p = o + v;
  return o;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) {
  const unsigned o = *p;
  *p = o + v;
  return o;
}

// CUDA's global min/max overload set (mixed integer widths are common in kernel code)
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long min(int a, long long b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }
inline long long max(int a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
