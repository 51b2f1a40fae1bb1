// simt.h -- launch helper of the sm_100a build, and the one seam through which the
// test-only SIMT emulator (tests/emu/simt_emu.h, g++ -DGLIA_SIMT_EMU) can compile the
// same kernel sources on a GPU-less machine to debug index logic (FFT pass maps,
// Hermitian packing, tile addressing).  The emulator is NOT part of the product:
// nothing under glia_b200/ contains it, the package only ever loads the nvcc-built
// libglia_rd.so and fails loudly when that (or a CUDA device) is missing.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <utility>

#if defined(GLIA_SIMT_EMU)
#include "simt_emu.h"  // tests/emu/ (test infrastructure only)
#else
// ============================ real CUDA ====================================
#include <cuda_runtime.h>
#include <mutex>
#include <unordered_map>

#define GLIA_UNROLL _Pragma("unroll")
// (128-byte aligned: bulk tensor copies -- sweeps_tma.cuh -- need their shared-memory tiles on 128-byte boundaries)
#define GLIA_DYN_SMEM(name) \
  extern __shared__ __align__(128) unsigned char _glia_dyn_smem[]; \
  unsigned char* name = _glia_dyn_smem

namespace simt {
// Kernels that want more than the 48 KB default of dynamic shared memory must opt in with
// cudaFuncSetAttribute, once per (device, function): the attribute lives in the device's context.
// One mutex and one map for every launch path -- handles on different host threads (ensemble
// members) and on different devices launch concurrently.
inline void opt_in_smem(const void* k, size_t smem) {
  // opt in above 32 KB already: static __shared__ (reduction scratch) counts against the 48 KB default too
  if (smem <= 32 * 1024) return;
  struct Key {
    int dev;
    const void* fn;
    bool operator==(const Key& o) const { return dev == o.dev && fn == o.fn; }
  };
  struct Hash {
    size_t operator()(const Key& k) const { return std::hash<const void*>()(k.fn) * 31u + (size_t)k.dev; }
  };
  static std::mutex mu;
  static std::unordered_map<Key, size_t, Hash> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(Key{dev, k});
  if (it == cache.end() || it->second < smem) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cache[Key{dev, k}] = smem;
  }
}
template <class... KA, class... A>
inline void launch(void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  opt_in_smem((const void*)k, smem);
  k<<<grid, block, smem, st>>>(static_cast<KA>(args)...);
}
// programmatic dependent launch: the kernel may become resident (and run everything above its
// pdl_wait()) while the previous kernel of the stream drains.  Only for kernels that call pdl_wait()
// before touching anything an earlier kernel wrote, and before writing anything one reads.
template <class... KA, class... A>
inline void launch_pdl(void (*k)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  opt_in_smem((const void*)k, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, static_cast<KA>(args)...);
}
// the same launch with an access-policy window: every access of this kernel into [win, win + bytes)
// is treated as streaming (evict-first) by L2.  Used for operands that LDGSTS stages (cp.async
// carries no usable per-instruction policy on sm_100a with this toolchain).
template <class... KA, class... A>
inline void launch_streaming(const void* win, size_t bytes, void (*k)(KA...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t st, A... args) {
  if (!win || !bytes) return launch(k, grid, block, smem, st, args...);
  opt_in_smem((const void*)k, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeAccessPolicyWindow;
  at[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(win);
  at[0].val.accessPolicyWindow.num_bytes = bytes;
  at[0].val.accessPolicyWindow.hitRatio = 1.0f;
  at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyStreaming;
  at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, static_cast<KA>(args)...);
}
inline const char* last_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
}  // namespace simt
#endif
