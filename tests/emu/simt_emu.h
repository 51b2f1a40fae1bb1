// simt_emu.h -- TEST INFRASTRUCTURE ONLY.  A thread-per-CUDA-thread SIMT emulator that lets
// g++ compile glia_b200/csrc/*.cu(h) (with -DGLIA_SIMT_EMU -Itests/emu) so that kernel index
// logic and the host-side PCG / time-stepping drivers can be exercised on a machine with no
// GPU.  Every CUDA thread of a CTA runs as an OS thread with real barriers; CTAs run one
// after another.  It is never loaded by the glia_b200 package.
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
#define __shared__ static
#define GLIA_UNROLL

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef int cudaStream_t;

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

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)

inline void __syncthreads() { emu::ctx.cta_bar->arrive_and_wait(); }
inline void emu_half_barrier(int h) { emu::ctx.half_bar[h]->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx.warp_bar->arrive_and_wait(); }

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  auto& c = emu::ctx;
  std::memcpy(c.warp_slots + 16 * c.lane, &v, sizeof(T));
  c.warp_bar->arrive_and_wait();
  T r;
  std::memcpy(&r, c.warp_slots + 16 * (src & 31), sizeof(T));
  c.warp_bar->arrive_and_wait();
  return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int mask) { return __shfl_sync(m, v, emu::ctx.lane ^ mask); }
template <class T>
inline T __shfl_down_sync(unsigned m, T v, int d) {
  int s = emu::ctx.lane + d;
  return __shfl_sync(m, v, s > 31 ? emu::ctx.lane : s);
}

inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

#define GLIA_DYN_SMEM(name) unsigned char* name = emu::ctx.smem

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
