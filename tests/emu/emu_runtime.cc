// Runtime of the host CUDA emulator (see cuda_emu.h): fibers, barriers, warp collectives,
// mbarrier / bulk-copy bookkeeping and the few cudart entry points the host wrappers call.
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <map>
#include <vector>

#include "cuda_emu.h"

// Minimal cooperative context switch (callee-saved registers + stack pointer).  glibc's
// swapcontext makes a signal-mask system call per switch, which dominated the run time.
#if !defined(__x86_64__)
#error "the CUDA host emulator's context switch is written for x86-64"
#endif
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

namespace emu {

Idx g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

namespace {

constexpr size_t kStack = 512 * 1024;
constexpr size_t kCanary = 256;  // bytes either side of dynamic shared memory
constexpr unsigned char kCanaryByte = 0xA5;

enum Wait { RUN = 0, BAR, WARP, MBAR, DONE };

struct PendingCopy {
  void* dst;
  const void* src;
  unsigned bytes;
  void* bar;  // bulk loads: the mbarrier that counts the bytes
};

struct Fiber {
  void* sp = nullptr;
  Wait wait = RUN;
  long gen = 0;        // generation being waited for (BAR / WARP)
  void* mbar = nullptr;
  unsigned parity = 0;
  std::vector<PendingCopy> stores;  // lazy mode: bulk stores not yet waited for (cp.async.bulk.wait_group.read)
  std::vector<PendingCopy> cps;     // lazy mode: cp.async copies not yet waited for (cp.async.wait_group)
};

struct Warp {
  unsigned arrived = 0, mask = 0;
  long gen = 0;
  uint64_t slot[32];
  uint64_t snap[2][32];
};

struct MBar {
  unsigned count = 0;   // arrivals per phase
  int pending = 0;      // arrivals still missing in the current phase
  long long tx = 0;     // outstanding transaction bytes
  unsigned phase = 0;
};

struct Block {
  int n = 0, live = 0, cur = 0;
  std::vector<Fiber> fib;
  std::vector<Warp> warps;
  void* sched_sp = nullptr;
  int bar_arrived = 0;
  long bar_gen = 0;
  std::map<void*, MBar> mbars;
  std::vector<PendingCopy> loads;   // lazy mode: bulk loads whose data has not been delivered yet
  unsigned char* smem_raw = nullptr;
  size_t smem_bytes = 0;
  const std::function<void()>* body = nullptr;
  dim3 bdim;
};

Block* g_blk = nullptr;
// Schedule of the fibers inside a sweep (test hook emu_set_schedule): 0 = ascending thread index, 1 = descending,
// 2 = a fresh pseudo-random permutation per sweep, 3 / 4 = warp-greedy (warps ascending / descending, each running
// until none of its lanes can continue).  A kernel that is free of shared-memory races between barriers
// gives the same result under every schedule; a missing barrier shows up as an order-dependent result.
// Asynchronous copies (test hook emu_set_async_mode).  Eager (default): the data moves when the copy is issued - the
// earliest the hardware may do it (exposes a destination or source that is still in use).  Lazy: the data moves when
// the consumer performs the wait the programming model requires (mbarrier wait for bulk loads, wait_group.read for
// bulk stores, cp.async.wait_group for cp.async) - the latest the hardware may do it (exposes a missing or misplaced
// wait: the reader finds the NaN fill of uninitialised shared memory).  A correct kernel is bit-identical under both.
bool g_lazy_async = false;
bool g_strict_masks = true;  // warp collectives must not name exited lanes
int g_sched_mode = 0;
unsigned g_sched_state = 1;
bool g_blocks_reversed = false;
std::vector<int> g_perm;
char* g_stacks = nullptr;
size_t g_stack_cap = 0;
std::map<const void*, int> g_max_dyn_smem;
int g_last_error = 0;

[[noreturn]] void die(const char* msg) {
  std::fprintf(stderr, "cuda-emu: %s (block %u,%u,%u thread %u)\n", msg, g_blockIdx.x, g_blockIdx.y, g_blockIdx.z,
               g_blk ? (unsigned)g_blk->cur : 0u);
  std::abort();
}

void set_thread(Block& b, int t) {
  b.cur = t;
  g_threadIdx.x = t % b.bdim.x;
  g_threadIdx.y = (t / b.bdim.x) % b.bdim.y;
  g_threadIdx.z = t / (b.bdim.x * b.bdim.y);
}

void yield() {
  Block& b = *g_blk;
  emu_switch(&b.fib[b.cur].sp, b.sched_sp);
}

void trampoline() {
  Block& b = *g_blk;
  (*b.body)();
  if (!b.fib[b.cur].stores.empty())
    die("thread exits with bulk stores it never waited for (cp.async.bulk.wait_group.read): shared memory may be "
        "released while the copy engine still reads it");
  if (!b.fib[b.cur].cps.empty()) die("thread exits with cp.async copies it never waited for");
  b.fib[b.cur].wait = DONE;
  emu_switch(&b.fib[b.cur].sp, b.sched_sp);
  std::abort();  // a finished fiber is never resumed
}

bool runnable(Block& b, int t) {
  Fiber& f = b.fib[t];
  switch (f.wait) {
    case RUN: return true;
    case BAR: return b.bar_gen > f.gen;
    case WARP: return b.warps[t >> 5].gen > f.gen;
    case MBAR: return b.mbars[f.mbar].phase != f.parity;
    case DONE: return false;
  }
  return false;
}

unsigned live_mask(Block& b, int warp) {
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) {
    const int t = warp * 32 + l;
    if (t < b.n && b.fib[t].wait != DONE) m |= 1u << l;
  }
  return m;
}

void complete_warp(Warp& w) {
  std::memcpy(w.snap[w.gen & 1], w.slot, sizeof(w.slot));
  w.arrived = 0;
  w.gen++;
}

void run_block(Block& b) {
  for (int t = 0; t < b.n; ++t) {
    Fiber& f = b.fib[t];
    f.wait = RUN;
    // initial frame: six callee-saved registers, then `ret` into trampoline with the stack
    // 8 mod 16 as after a call
    uintptr_t top = ((uintptr_t)(g_stacks + (size_t)(t + 1) * kStack)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of trampoline
    *--sp = (void*)&trampoline;
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    f.sp = sp;
  }
  b.live = b.n;
  if ((int)g_perm.size() != b.n) {
    g_perm.resize(b.n);
    for (int i = 0; i < b.n; ++i) g_perm[i] = i;
  }
  // one fiber segment: resume thread t until it blocks or exits
  auto run_one = [&](int t) {
    b.fib[t].wait = RUN;
    set_thread(b, t);
    emu_switch(&b.sched_sp, b.fib[t].sp);
    if (b.fib[t].wait == DONE) {
      b.live--;
      // exited threads no longer count towards barriers / warp collectives
      if (b.bar_arrived > 0 && b.bar_arrived == b.live) {
        b.bar_arrived = 0;
        b.bar_gen++;
      }
      Warp& w = b.warps[t >> 5];
      const unsigned need = w.mask & live_mask(b, t >> 5);
      if (w.arrived != 0 && (w.arrived & need) == need) complete_warp(w);
    }
  };
  while (b.live > 0) {
    bool progress = false;
    if (g_sched_mode >= 3) {
      // warp-greedy: each warp in turn runs for as long as any of its lanes can (maximal skew between warps:
      // a warp reaches the next block barrier before the others have left the previous phase)
      const int nw = (b.n + 31) / 32;
      for (int k = 0; k < nw; ++k) {
        const int w = g_sched_mode == 3 ? k : nw - 1 - k;
        bool again = true;
        while (again) {
          again = false;
          for (int l = 0; l < 32; ++l) {
            const int t = w * 32 + l;
            if (t >= b.n || !runnable(b, t)) continue;
            run_one(t);
            again = progress = true;
          }
        }
      }
      if (!progress) die("deadlock: no runnable thread (mismatched __syncthreads / collective / mbarrier wait)");
      continue;
    }
    if (g_sched_mode == 2) {  // Fisher-Yates with a small LCG: reproducible for a given seed
      for (int i = b.n - 1; i > 0; --i) {
        g_sched_state = g_sched_state * 1664525u + 1013904223u;
        std::swap(g_perm[i], g_perm[(g_sched_state >> 8) % (unsigned)(i + 1)]);
      }
    }
    for (int k = 0; k < b.n; ++k) {
      const int t = g_sched_mode == 0 ? k : g_sched_mode == 1 ? b.n - 1 - k : g_perm[k];
      if (!runnable(b, t)) continue;
      run_one(t);
      progress = true;
    }
    if (!progress) die("deadlock: no runnable thread (mismatched __syncthreads / collective / mbarrier wait)");
  }
}

}  // namespace

void* dyn_smem() { return g_blk->smem_raw + kCanary; }

bool in_dyn_smem(const void* p, size_t bytes) {
  const unsigned char* q = (const unsigned char*)p;
  const unsigned char* lo = g_blk->smem_raw + kCanary;
  return q >= lo && q + bytes <= lo + g_blk->smem_bytes;
}

void sync_threads() {
  Block& b = *g_blk;
  Fiber& f = b.fib[b.cur];
  const long my = b.bar_gen;
  if (++b.bar_arrived == b.live) {
    b.bar_arrived = 0;
    b.bar_gen++;
    return;
  }
  f.wait = BAR;
  f.gen = my;
  yield();
}

const uint64_t* warp_exchange(unsigned mask, uint64_t v) {
  Block& b = *g_blk;
  const int t = b.cur, wi = t >> 5, lane = t & 31;
  Warp& w = b.warps[wi];
  if (!((mask >> lane) & 1u)) die("warp collective: calling lane is not in its own mask");
  const unsigned lm = live_mask(b, wi);
  // a collective that names a lane which has already exited is undefined behaviour on the hardware (lanes beyond the
  // end of a partial last warp do not exist and are ignored, as before)
  if (g_strict_masks) {
    const int in_warp = b.n - wi * 32 >= 32 ? 32 : b.n - wi * 32;
    const unsigned exist = in_warp >= 32 ? 0xffffffffu : ((1u << in_warp) - 1u);
    if (mask & exist & ~lm) die("warp collective names a lane that has already exited (undefined behaviour on the GPU)");
  }
  if (w.arrived == 0) w.mask = mask;
  else if (w.mask != mask) die("warp collective: lanes of one warp passed different masks");
  if ((w.arrived >> lane) & 1u) die("warp collective: lane arrived twice");
  w.slot[lane] = v;
  w.arrived |= 1u << lane;
  const long my = w.gen;
  const unsigned need = mask & lm;
  if ((w.arrived & need) == need) {
    complete_warp(w);
  } else {
    Fiber& f = b.fib[t];
    f.wait = WARP;
    f.gen = my;
    yield();
  }
  return w.snap[my & 1];
}

void mbar_init(void* bar, unsigned count) {
  if (!in_dyn_smem(bar, 8) || ((uintptr_t)bar & 7)) die("mbarrier.init: not an 8-byte aligned shared-memory address");
  MBar& m = g_blk->mbars[bar];
  m.count = count;
  m.pending = (int)count;
  m.tx = 0;
  m.phase = 0;
}
static void mbar_maybe_complete(MBar& m) {
  if (m.pending == 0 && m.tx == 0) {
    m.phase ^= 1u;
    m.pending = (int)m.count;
  }
}
void mbar_expect_tx(void* bar, unsigned bytes) {
  auto it = g_blk->mbars.find(bar);
  if (it == g_blk->mbars.end()) die("mbarrier.arrive.expect_tx on an uninitialised barrier");
  MBar& m = it->second;
  if (m.pending <= 0) die("mbarrier: more arrivals than the initialised count");
  m.tx += bytes;
  m.pending--;
  mbar_maybe_complete(m);
}
void mbar_complete_tx(void* bar, unsigned bytes) {
  auto it = g_blk->mbars.find(bar);
  if (it == g_blk->mbars.end()) die("bulk copy signals an uninitialised mbarrier");
  MBar& m = it->second;
  m.tx -= bytes;
  mbar_maybe_complete(m);
}
void mbar_wait(void* bar, unsigned parity) {
  Block& b = *g_blk;
  auto it = b.mbars.find(bar);
  if (it == b.mbars.end()) die("mbarrier.try_wait on an uninitialised barrier");
  // lazy mode: the bytes counted on this barrier arrive now, at the wait
  for (size_t i = 0; i < b.loads.size();) {
    if (b.loads[i].bar == bar) {
      const PendingCopy c = b.loads[i];
      b.loads.erase(b.loads.begin() + i);
      std::memcpy(c.dst, c.src, c.bytes);
      mbar_complete_tx(bar, c.bytes);
    } else {
      ++i;
    }
  }
  if (it->second.phase != (parity & 1u)) return;
  Fiber& f = b.fib[b.cur];
  f.wait = MBAR;
  f.mbar = bar;
  f.parity = parity & 1u;
  yield();
}
void async_load(void* smem_dst, const void* gsrc, unsigned bytes, void* bar) {
  bool waited_for = false;  // a consumer already blocked on this barrier: "its wait" is now
  for (const Fiber& f : g_blk->fib) waited_for = waited_for || (f.wait == MBAR && f.mbar == bar);
  if (g_lazy_async && !waited_for) {
    g_blk->loads.push_back({smem_dst, gsrc, bytes, bar});
  } else {
    std::memcpy(smem_dst, gsrc, bytes);
    mbar_complete_tx(bar, bytes);
  }
}
void async_store(void* gdst, const void* smem_src, unsigned bytes) {
  if (g_lazy_async) g_blk->fib[g_blk->cur].stores.push_back({gdst, smem_src, bytes, nullptr});
  else std::memcpy(gdst, smem_src, bytes);
}
void async_store_wait() {
  Fiber& f = g_blk->fib[g_blk->cur];
  for (const PendingCopy& c : f.stores) std::memcpy(c.dst, c.src, c.bytes);
  f.stores.clear();
}
void cp_async(void* smem_dst, const void* gsrc, unsigned bytes) {
  if (g_lazy_async) g_blk->fib[g_blk->cur].cps.push_back({smem_dst, gsrc, bytes, nullptr});
  else std::memcpy(smem_dst, gsrc, bytes);
}
void cp_async_wait() {
  Fiber& f = g_blk->fib[g_blk->cur];
  for (const PendingCopy& c : f.cps) std::memcpy(c.dst, c.src, c.bytes);
  f.cps.clear();
}
long g_bulk_copies = 0;
long g_peer_stores = 0;
void count_peer_store() { ++g_peer_stores; }
void check_bulk(const void* smem_ptr, const void* gptr, unsigned bytes, const char* what) {
  ++g_bulk_copies;
  if (bytes == 0 || (bytes & 15u)) die((std::string(what) + ": size is not a positive multiple of 16 bytes").c_str());
  if ((uintptr_t)smem_ptr & 15u) die((std::string(what) + ": shared-memory address not 16-byte aligned").c_str());
  if ((uintptr_t)gptr & 15u) die((std::string(what) + ": global address not 16-byte aligned").c_str());
  if (!in_dyn_smem(smem_ptr, bytes)) die((std::string(what) + ": shared-memory range outside the CTA's allocation").c_str());
}

void launch(const void* func, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const size_t n = (size_t)block.x * block.y * block.z;
  if (n == 0 || n > 1024 || (size_t)grid.x * grid.y * grid.z == 0 || grid.y > 65535 || grid.z > 65535 || grid.x > 2147483647u ||
      block.z > 64) {
    g_last_error = (int)cudaErrorInvalidConfiguration;
    return;
  }
  int allowed = 48 * 1024;
  auto it = g_max_dyn_smem.find(func);
  if (it != g_max_dyn_smem.end()) allowed = it->second;
  if (smem > (size_t)allowed || smem > 227 * 1024) {
    std::fprintf(stderr, "cuda-emu: launch with %zu B dynamic shared memory (allowed %d)\n", smem, allowed);
    g_last_error = (int)cudaErrorInvalidValue;
    return;
  }
  if (g_stack_cap < n) {
    if (g_stacks) munmap(g_stacks, g_stack_cap * kStack);
    g_stacks = (char*)mmap(nullptr, n * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_stacks == MAP_FAILED) {
      std::perror("cuda-emu mmap");
      std::abort();
    }
    g_stack_cap = n;
  }
  Block b;
  b.n = (int)n;
  b.bdim = block;
  b.fib.resize(n);
  b.warps.resize((n + 31) / 32);
  b.body = &body;
  b.smem_bytes = smem;
  // Shared-memory window: its END sits against an inaccessible guard page (reads or writes past the
  // allocation fault, to within the 128-byte alignment of the start), a canary guards the front.
  const size_t page = 4096;
  const size_t span = ((kCanary + smem + 127) / 128 * 128 + page - 1) / page * page;
  unsigned char* region = (unsigned char*)mmap(nullptr, span + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (region == MAP_FAILED) {
    std::perror("cuda-emu mmap");
    std::abort();
  }
  mprotect(region + span, page, PROT_NONE);
  unsigned char* win = region + span - (smem + 127) / 128 * 128;  // 128-byte aligned start like the hardware window
  unsigned char* base = win - kCanary;
  b.smem_raw = base;
  g_blockDim = {block.x, block.y, block.z};
  g_gridDim = {grid.x, grid.y, grid.z};
  Block* saved = g_blk;
  g_blk = &b;
  for (unsigned zz = 0; zz < grid.z; ++zz)
    for (unsigned yy = 0; yy < grid.y; ++yy)
      for (unsigned xx = 0; xx < grid.x; ++xx) {
        const unsigned x = g_blocks_reversed ? grid.x - 1 - xx : xx, y = g_blocks_reversed ? grid.y - 1 - yy : yy,
                       z = g_blocks_reversed ? grid.z - 1 - zz : zz;
        g_blockIdx = {x, y, z};
        std::memset(base, kCanaryByte, kCanary);
        std::memset(win, 0xFF, region + span - win);  // uninitialised shared memory reads as NaN
        b.bar_arrived = 0;
        b.mbars.clear();
        b.loads.clear();
        for (auto& f : b.fib) {
          f.stores.clear();
          f.cps.clear();
        }
        for (auto& w : b.warps) w = Warp();
        run_block(b);
        for (size_t i = 0; i < kCanary; ++i)
          if (base[i] != kCanaryByte) die("dynamic shared memory written out of bounds (before the window)");
      }
  munmap(region, span + page);
  g_blk = saved;
}

}  // namespace emu

// ---- the cudart entry points used by the host wrappers ----------------------------------------
namespace {
void on_ill(int) {
  static const char msg[] = "cuda-emu: misaligned vector access (float2 / float4 / 64-bit dereference of an address that is "
                            "not a multiple of the type's alignment) - a fault on the GPU\n";
  ssize_t r = write(2, msg, sizeof(msg) - 1);
  (void)r;
  _exit(135);
}
struct InstallTrapHandler {
  InstallTrapHandler() { signal(SIGILL, on_ill); }
} g_install_trap_handler;
void on_segv(int) {
  static const char msg[] = "cuda-emu: memory access outside a guarded allocation (out-of-bounds global or shared access)\n";
  ssize_t r = write(2, msg, sizeof(msg) - 1);
  (void)r;
  _exit(134);
}
}  // namespace

extern "C" {
// Test hook: "device" allocation whose end (tight_end != 0) or start (tight_end == 0) touches an
// inaccessible guard page, so that out-of-bounds global reads and writes of the kernels fault.
// start_mod16: byte offset of the returned pointer modulo 16 (0 = 16-byte aligned).
void* emu_guarded_alloc(size_t bytes, int tight_end, int start_mod16) {
  static bool handler = false;
  if (!handler) {
    signal(SIGSEGV, on_segv);
    handler = true;
  }
  const size_t page = 4096;
  const size_t span = (bytes + 16 + page - 1) / page * page;
  unsigned char* region = (unsigned char*)mmap(nullptr, span + 2 * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (region == MAP_FAILED) return nullptr;
  mprotect(region, page, PROT_NONE);
  mprotect(region + page + span, page, PROT_NONE);
  unsigned char* lo = region + page;
  unsigned char* p;
  if (tight_end) {
    p = lo + span - bytes;
    p -= ((uintptr_t)p - (uintptr_t)start_mod16) & 15;  // largest address <= with the requested phase
  } else {
    p = lo + (start_mod16 & 15);
  }
  std::memset(lo, 0xFF, span);
  return p;
}
// test hook: order in which the threads of a CTA (mode 0 ascending, 1 descending, 2 random with `seed`) and the CTAs
// of a grid (reversed != 0: last block first) are run
void emu_set_schedule(int mode, unsigned seed, int blocks_reversed) {
  emu::g_sched_mode = mode;
  emu::g_sched_state = seed ? seed : 1u;
  emu::g_blocks_reversed = blocks_reversed != 0;
  emu::g_perm.clear();
}
// test hook: 0 = asynchronous copies deliver at issue (default), 1 = at the consumer's wait
void emu_set_async_mode(int lazy) { emu::g_lazy_async = lazy != 0; }
// self-test of the alignment check: dereference a float4 at base + offset_floats
float emu_selftest_align(const float* base, int offset_floats) {
  const float4 v = *reinterpret_cast<const float4*>(base + offset_floats);
  return v.x + v.w;
}
// self-test of the lazy mode: one thread bulk-loads 16 bytes and reads them with or without the mbarrier wait
void emu_selftest_async(const float* src16, float* out, int with_wait) {
  const std::function<void()> body = [=]() {
    unsigned char* sm = (unsigned char*)emu::dyn_smem();
    unsigned long long* bar = (unsigned long long*)sm;
    float* dst = (float*)(sm + 16);
    emu::mbar_init(bar, 1);
    emu::mbar_expect_tx(bar, 16);
    emu::async_load(dst, src16, 16, bar);
    if (with_wait) emu::mbar_wait(bar, 0);
    out[0] = dst[2];
  };
  emu::launch((const void*)&emu_selftest_async, dim3(1), dim3(1), 64, body);
}
// self-test of the schedule hook: every thread publishes a value in shared memory and reads its neighbour's, with or
// without the barrier in between - without it the output depends on the order the threads run in
void emu_selftest_race(float* out, int n, int with_barrier) {
  const std::function<void()> body = [=]() {
    float* sm = (float*)emu::dyn_smem();
    const int t = (int)emu::g_threadIdx.x;
    sm[t] = (float)(t + 1);
    if (with_barrier) emu::sync_threads();
    out[t] = sm[(t + 1) % n];
  };
  emu::launch((const void*)&emu_selftest_race, dim3(1), dim3(n), (size_t)n * sizeof(float), body);
}
// test hook: number of one-sided cluster stores (st.async) issued since the last call
long emu_peer_stores(void) {
  const long n = emu::g_peer_stores;
  emu::g_peer_stores = 0;
  return n;
}
// test hook: number of bulk (TMA) copies issued since the last call
long emu_bulk_copies(void) {
  const long n = emu::g_bulk_copies;
  emu::g_bulk_copies = 0;
  return n;
}
cudaError_t cudaFuncSetAttribute(const void* func, cudaFuncAttribute attr, int value) {
  if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize) {
    if (value > 227 * 1024) return cudaErrorInvalidValue;
    emu::g_max_dyn_smem[func] = value;
  }
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) {
  const int e = emu::g_last_error;
  emu::g_last_error = 0;
  return (cudaError_t)e;
}
cudaError_t cudaPeekAtLastError(void) { return (cudaError_t)emu::g_last_error; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda-emu error"; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  std::memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
}
