// simt.h -- thin portability seam between the real sm_100a build (nvcc) and the
// test-only SIMT emulator (g++ -DGLIA_SIMT_EMU, tests/emu/).
//
// The emulator exists so that kernel index logic (FFT pass maps, Hermitian
// packing, tile addressing) can be debugged on the GPU-less build container.
// It runs every CUDA thread of a CTA as an OS thread with real barriers.  It is
// NOT a product path: the glia_b200 package only ever loads the nvcc-built
// libglia_rd.so and fails loudly when it is missing.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <utility>

#if defined(GLIA_SIMT_EMU)
// ============================ emulator ====================================
#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
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
inline const char* last_error() { return nullptr; }
}  // namespace simt

#else
// ============================ real CUDA ====================================
#include <cuda_runtime.h>
#include <mutex>
#include <unordered_map>

#define GLIA_UNROLL _Pragma("unroll")
#define GLIA_DYN_SMEM(name) \
  extern __shared__ __align__(16) unsigned char _glia_dyn_smem[]; \
  unsigned char* name = _glia_dyn_smem

namespace simt {
inline std::unordered_map<const void*, size_t>& smem_attr_cache() {
  static std::unordered_map<const void*, size_t> m;
  return m;
}
template <class... KA, class... A>
inline void launch(void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  if (smem > 48 * 1024) {
    auto& m = smem_attr_cache();
    auto it = m.find((const void*)k);
    if (it == m.end() || it->second < smem) {
      cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      m[(const void*)k] = smem;
    }
  }
  k<<<grid, block, smem, st>>>(static_cast<KA>(args)...);
}
inline const char* last_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
}  // namespace simt
#endif
