// simt_emu.h -- TEST INFRASTRUCTURE ONLY.  A SIMT emulator that lets g++ compile
// glia_b200/csrc/*.cu(h) (with -DGLIA_SIMT_EMU -Itests/emu) so that kernel index logic and the
// host-side PCG / time-stepping drivers can be exercised on a machine with no GPU.  It is never
// loaded by the glia_b200 package.
//
// Every CUDA thread of a CTA runs as a FIBER (its own 64 KB stack, cooperative switch at barriers) on
// the OS thread that executes the CTA; a launch hands its CTAs, in index order, to a small pool of OS
// threads.  __syncthreads / __syncwarp / named half barriers are generation counters: a fiber that
// arrives early yields to the next fiber of its CTA.  (The first version ran one OS thread per CUDA
// thread with std::barrier: correct, but a 512-thread CTA spent its time in futex calls -- the CPU
// suite took 10 minutes.)  On targets other than x86-64 the OS-thread form below is used.
#pragma once
#include <atomic>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifndef __grid_constant__
#define __grid_constant__
#endif
#include <barrier>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __launch_bounds__(...)
// "shared" statics: one copy per OS thread = per CTA in flight (CTAs of a launch run on several OS threads)
#define __shared__ static thread_local
#define GLIA_UNROLL

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef int cudaStream_t;

#if defined(__x86_64__) && !defined(GLIA_EMU_OS_THREADS)
// =============================================================== fibers ====
namespace emu {
struct Bar {  // generation barrier among fibers of one OS thread: no atomics needed
  int n = 1, count = 0;
  unsigned gen = 0;
};
struct Ctx {
  dim3 tid, bid, bdim, gdim;
  unsigned char* smem = nullptr;
  Bar* cta_bar = nullptr;
  Bar* half_bar[2] = {nullptr, nullptr};  // bar.sync 1 / 2 over the lower / upper half of the CTA's threads
  Bar* warp_bar = nullptr;
  unsigned char* warp_slots = nullptr;  // 32 x 16 bytes scratch for shuffles
  int lane = 0;
  // fiber state
  void* sp = nullptr;
  bool finished = false;
};
struct Worker {  // one per OS thread that executes CTAs
  std::vector<Ctx> fib;
  std::vector<void*> stacks;  // reused across launches
  int nthr = 0, cur = 0, live = 0;
  void* main_sp = nullptr;
  const std::function<void()>* body = nullptr;
  ~Worker() {
    for (void* st : stacks) munmap(st, kStack);
  }
  static constexpr size_t kStack = 64 * 1024;
};
inline thread_local Worker worker;
inline thread_local Ctx* cur = nullptr;

// save the callee-saved registers on the current stack, publish its stack pointer, continue on `to`
__attribute__((naked, noinline)) static void fiber_switch(void** /*from_sp*/, void* /*to_sp*/) {
  asm volatile(
      "pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
      "movq %rsp, (%rdi)\n"
      "movq %rsi, %rsp\n"
      "popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n"
      "ret\n");
}
inline void switch_to(int j) {
  Worker& w = worker;
  Ctx* from = &w.fib[w.cur];
  w.cur = j;
  cur = &w.fib[j];
  fiber_switch(&from->sp, w.fib[j].sp);
}
// run somebody else of this CTA; returns when this fiber is resumed
inline void yield() {
  Worker& w = worker;
  int j = w.cur;
  for (int k = 1; k < w.nthr; ++k) {
    if (++j == w.nthr) j = 0;
    if (!w.fib[j].finished) { switch_to(j); return; }
  }
}
inline void bar_wait(Bar* b) {
  const unsigned g = b->gen;
  if (++b->count == b->n) {
    b->count = 0;
    ++b->gen;
  } else {
    while (b->gen == g) yield();
  }
}
[[noreturn]] static void fiber_main() {
  Worker& w = worker;
  (*w.body)();
  Ctx* me = &w.fib[w.cur];
  me->finished = true;
  if (--w.live == 0) {
    cur = nullptr;
    fiber_switch(&me->sp, w.main_sp);
  } else {
    int j = w.cur;
    for (;;) {
      if (++j == w.nthr) j = 0;
      if (!w.fib[j].finished) break;
    }
    w.cur = j;
    cur = &w.fib[j];
    fiber_switch(&me->sp, w.fib[j].sp);
  }
  std::abort();  // a finished fiber is never resumed
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F f) {
  const int nthr = (int)(block.x * block.y * block.z);
  const int nwarp = (nthr + 31) / 32;
  const long ncta = (long)grid.x * grid.y * grid.z;
  const std::function<void()> body = f;
  std::atomic<long> next{0};
  auto run = [&]() {
    Worker& w = worker;
    w.nthr = nthr;
    w.body = &body;
    if ((int)w.fib.size() < nthr) w.fib.resize(nthr);
    while ((int)w.stacks.size() < nthr) {
      void* st = mmap(nullptr, Worker::kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (st == MAP_FAILED) std::abort();
      w.stacks.push_back(st);
    }
    std::vector<unsigned char> smem(smem_bytes + 64);
    Bar cta_bar, half_lo, half_hi;
    std::vector<Bar> wbars(nwarp);
    std::vector<unsigned char> wslots((size_t)nwarp * 32 * 16);
    for (;;) {
      const long c = next.fetch_add(1);  // CTAs start in index order
      if (c >= ncta) break;
      cta_bar = Bar{nthr};
      half_lo = Bar{nthr / 2 > 0 ? nthr / 2 : 1};
      half_hi = Bar{nthr - nthr / 2 > 0 ? nthr - nthr / 2 : 1};
      for (int wi = 0; wi < nwarp; ++wi) wbars[wi] = Bar{(wi == nwarp - 1) ? nthr - 32 * wi : 32};
      const dim3 bid((unsigned)(c % grid.x), (unsigned)((c / grid.x) % grid.y), (unsigned)(c / ((long)grid.x * grid.y)));
      for (int i = 0; i < nthr; ++i) {
        Ctx& x = w.fib[i];
        x.bdim = block;
        x.gdim = grid;
        x.bid = bid;
        x.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
        x.smem = smem.data();
        x.cta_bar = &cta_bar;
        x.half_bar[0] = &half_lo;
        x.half_bar[1] = &half_hi;
        x.warp_bar = &wbars[i / 32];
        x.warp_slots = wslots.data() + (size_t)(i / 32) * 32 * 16;
        x.lane = i % 32;
        x.finished = false;
        // initial frame: six zeroed callee-saved registers, the entry point, one slot so that the entry sees the
        // stack as after a call (rsp = 16 n + 8)
        void** top = reinterpret_cast<void**>(static_cast<unsigned char*>(w.stacks[i]) + Worker::kStack);
        top[-1] = nullptr;
        top[-2] = reinterpret_cast<void*>(&fiber_main);
        for (int k = 3; k <= 8; ++k) top[-k] = nullptr;
        x.sp = top - 8;
      }
      w.live = nthr;
      w.cur = 0;
      cur = &w.fib[0];
      fiber_switch(&w.main_sp, w.fib[0].sp);  // returns when the last fiber of the CTA has finished
    }
  };
  unsigned nworker = std::thread::hardware_concurrency();
  if (const char* e = std::getenv("GLIA_EMU_WORKERS")) nworker = (unsigned)std::atoi(e);
  if (nworker < 1) nworker = 1;
  if (nworker > 8) nworker = 8;
  if ((long)nworker > ncta) nworker = (unsigned)ncta;
  if (nworker <= 1) {
    run();
  } else {
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nworker; ++i) th.emplace_back(run);
    for (auto& t : th) t.join();
  }
}
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)
#define GLIA_EMU_CTX (*emu::cur)

inline void __syncthreads() { emu::bar_wait(emu::cur->cta_bar); }
inline void emu_half_barrier(int h) { emu::bar_wait(emu::cur->half_bar[h]); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::bar_wait(emu::cur->warp_bar); }
inline void emu_warp_barrier() { emu::bar_wait(emu::cur->warp_bar); }

#else
// ============================================================ OS threads ====
namespace emu {
struct Ctx {
  dim3 tid, bid, bdim, gdim;
  unsigned char* smem = nullptr;
  std::barrier<>* cta_bar = nullptr;
  std::barrier<>* half_bar[2] = {nullptr, nullptr};  // bar.sync 1 / 2 over the lower / upper half of the CTA's threads
  std::barrier<>* warp_bar = nullptr;
  unsigned char* warp_slots = nullptr;  // 32 x 16 bytes scratch for shuffles
  int lane = 0;
};
inline thread_local Ctx ctx;

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F f) {
  const int nthr = block.x * block.y * block.z;
  const int nwarp = (nthr + 31) / 32;
  std::vector<unsigned char> smem(smem_bytes + 64);
  std::barrier<> cta_bar(nthr);
  std::barrier<> half_lo(nthr / 2 > 0 ? nthr / 2 : 1), half_hi(nthr - nthr / 2 > 0 ? nthr - nthr / 2 : 1);
  std::vector<std::unique_ptr<std::barrier<>>> wbars;
  std::vector<std::vector<unsigned char>> wslots(nwarp, std::vector<unsigned char>(32 * 16));
  for (int w = 0; w < nwarp; ++w) {
    int cnt = (w == nwarp - 1) ? nthr - 32 * w : 32;
    wbars.emplace_back(new std::barrier<>(cnt));
  }
  std::vector<std::thread> th;
  th.reserve(nthr);
  for (int i = 0; i < nthr; ++i) {
    th.emplace_back([&, i]() {
      Ctx& c = ctx;
      c.bdim = block;
      c.gdim = grid;
      c.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
      c.smem = smem.data();
      c.cta_bar = &cta_bar;
      c.half_bar[0] = &half_lo;
      c.half_bar[1] = &half_hi;
      c.warp_bar = wbars[i / 32].get();
      c.warp_slots = wslots[i / 32].data();
      c.lane = i % 32;
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            c.bid = dim3(bx, by, bz);
            f();
            cta_bar.arrive_and_wait();
          }
    });
  }
  for (auto& t : th) t.join();
}
}  // namespace emu

#undef __shared__
#define __shared__ static
#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)
#define GLIA_EMU_CTX (emu::ctx)

inline void __syncthreads() { emu::ctx.cta_bar->arrive_and_wait(); }
inline void emu_half_barrier(int h) { emu::ctx.half_bar[h]->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx.warp_bar->arrive_and_wait(); }
inline void emu_warp_barrier() { emu::ctx.warp_bar->arrive_and_wait(); }
#endif

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  auto& c = GLIA_EMU_CTX;
  std::memcpy(c.warp_slots + 16 * c.lane, &v, sizeof(T));
  emu_warp_barrier();
  T r;
  std::memcpy(&r, c.warp_slots + 16 * (src & 31), sizeof(T));
  emu_warp_barrier();
  return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int mask) { return __shfl_sync(m, v, GLIA_EMU_CTX.lane ^ mask); }
template <class T>
inline T __shfl_down_sync(unsigned m, T v, int d) {
  int s = GLIA_EMU_CTX.lane + d;
  return __shfl_sync(m, v, s > 31 ? GLIA_EMU_CTX.lane : s);
}

inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

#define GLIA_DYN_SMEM(name) unsigned char* name = GLIA_EMU_CTX.smem

namespace simt {
template <class... KA, class... A>
inline void launch(void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t, A... args) {
  emu::launch(grid, block, smem, [=]() { k(args...); });
}
template <class... KA, class... A>
inline void launch_streaming(const void*, size_t, void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             A... args) {
  launch(k, grid, block, smem, st, args...);
}
inline const char* last_error() { return nullptr; }
}  // namespace simt


namespace glia {
namespace rt {
inline int dev_malloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 1; }
inline void dev_free(void* p) { std::free(p); }
inline int host_malloc(void** p, size_t n) { return dev_malloc(p, n); }
inline void host_free(void* p) { std::free(p); }
inline int copy(void* d, const void* s, size_t n, cudaStream_t) { std::memmove(d, s, n); return 0; }
inline int h2d(void* d, const void* s, size_t n, cudaStream_t st) { return copy(d, s, n, st); }
inline int d2h(void* d, const void* s, size_t n, cudaStream_t st) { return copy(d, s, n, st); }
inline int zero(void* d, size_t n, cudaStream_t) { std::memset(d, 0, n); return 0; }
inline int sync(cudaStream_t) { return 0; }
inline int set_device(int) { return 0; }
inline bool is_pinned(const void*) { return true; }
// "IPC" of the emulator: POSIX shared memory, so that two emulator PROCESSES (the gloo
// world_size-2 CPU tests) can map each other's arenas exactly like two GPU ranks do.
inline int ipc_alloc(void** p, size_t n, unsigned char handle[64]) {
  static int counter = 0;
  std::memset(handle, 0, 64);
  std::snprintf((char*)handle, 64, "/glia_emu_%d_%d", (int)getpid(), counter++);
  int fd = shm_open((const char*)handle, O_CREAT | O_RDWR, 0600);
  if (fd < 0) return 1;
  if (ftruncate(fd, (off_t)n) != 0) { close(fd); return 1; }
  void* m = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return 1;
  *p = m;
  return 0;
}
inline int ipc_open(void** p, const unsigned char handle[64], size_t n) {
  int fd = shm_open((const char*)handle, O_RDWR, 0600);
  if (fd < 0) return 1;
  void* m = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return 1;
  *p = m;
  return 0;
}
inline void ipc_close(void* p, size_t n) { if (p) munmap(p, n); }
inline void ipc_free(void* p, size_t n, const unsigned char* handle) {
  if (p) munmap(p, n);
  if (handle && handle[0]) shm_unlink((const char*)handle);
}
inline int sm_count(int) { return 3; }
inline size_t max_policy_window(int) { return 0; }
inline int stream_create(cudaStream_t* s) { *s = 0; return 0; }
// fork / join of a side stream: the emulator runs every launch to completion in program order
struct Fork {
  cudaStream_t side = 0;
  int create() { return 0; }
  void destroy() {}
  void begin(cudaStream_t) {}
  void end(cudaStream_t) {}
};
inline void stream_destroy(cudaStream_t) {}
inline int stream_wait_stream(cudaStream_t, cudaStream_t) { return 0; }
inline const char* err_string(int) { return "emu"; }
struct Profiler {
  bool on = false;
  int before(const char*, cudaStream_t) { return -1; }
  void after(int, cudaStream_t) {}
  void begin() {}
  std::string end(cudaStream_t) { return std::string(); }
  void destroy() {}
};
struct Timer { void create() {} void destroy() {} void start(cudaStream_t) {} double stop_ms(cudaStream_t) { return 0; } };
}  // namespace rt
}  // namespace glia
