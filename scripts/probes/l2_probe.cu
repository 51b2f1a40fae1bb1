// l2_probe.cu -- measurement only (not part of the library): can a 64 MiB intermediate field
// (acc / shat of the PCG iteration at 256^3 f32) be kept resident in B200's 126 MB L2 between
// consecutive sweep kernels, and by which mechanism?
//   variant 0  plain loads / stores
//   variant 1  per-access L2 policies (createpolicy + ld/st.L2::cache_hint): the chained field is
//              stored evict_last, read evict_last, its last read is evict_first; the operands
//              that are only streamed (x, k) are read evict_first
//   variant 2  as 1 with the persisting-L2 set-aside raised to the device maximum
//   variant 3  cudaLaunchAttributeAccessPolicyWindow (persisting) on the chained field, set-aside = max
//   variant 4  only the streaming hints (evict_first on x, k), nothing pinned
// The kernel chain mimics one D-apply: K1 acc = x*k; K2 acc += x*k; K3 w = x + acc + x*k,
// traversed in the same or in a permuted block order (the y and x sweeps visit tiles in a
// different order than the z sweep).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_probe.bin l2_probe.cu && ./l2_probe.bin
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

enum { P_NORMAL = 0, P_LAST = 1, P_FIRST = 2 };
__device__ __forceinline__ uint64_t make_pol(int p) {
  uint64_t r;
  if (p == P_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(r));
  else if (p == P_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(r));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(r));
  return r;
}
__device__ __forceinline__ float4 ldh(const float4* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void sth(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// one block handles a contiguous 64 KiB piece; `perm` permutes the piece order
// out = a*b (+ c);  HINT: use cache-hinted accesses with the given policy classes
template <bool HINT, bool HAS_C>
__global__ void __launch_bounds__(256) chain(const float4* a, const float4* b, const float4* c, float4* out, long npiece,
                                             int perm, int pa, int pc, int po) {
  uint64_t qa = 0, qc = 0, qo = 0;
  if (HINT) { qa = make_pol(pa); qc = make_pol(pc); qo = make_pol(po); }
  for (long piece = blockIdx.x; piece < npiece; piece += gridDim.x) {
    // permuted order: a stride-257 walk over the pieces (npiece is a power of two)
    const long pp = perm ? (piece * 257) & (npiece - 1) : piece;
    const long base = pp * 4096;  // float4 units: 64 KiB
    for (int i = threadIdx.x; i < 4096; i += 1024) {
      float4 va[4], vb[4], vc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long o = base + i + u * 256;
        if (HINT) { va[u] = ldh(a + o, qa); vb[u] = ldh(b + o, qa); if (HAS_C) vc[u] = ldh(c + o, qc); }
        else { va[u] = a[o]; vb[u] = b[o]; if (HAS_C) vc[u] = c[o]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long o = base + i + u * 256;
        float4 r = {va[u].x * vb[u].x, va[u].y * vb[u].y, va[u].z * vb[u].z, va[u].w * vb[u].w};
        if (HAS_C) { r.x += vc[u].x; r.y += vc[u].y; r.z += vc[u].z; r.w += vc[u].w; }
        if (HINT) sth(out + o, r, qo); else out[o] = r;
      }
    }
  }
}

template <bool HINT, bool HAS_C>
static void launch(cudaStream_t st, int grid, const float* a, const float* b, const float* c, float* out, long n, int perm,
                   int pa, int pc, int po, const void* win, size_t wbytes) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (win) {
    at[0].id = cudaLaunchAttributeAccessPolicyWindow;
    at[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(win);
    at[0].val.accessPolicyWindow.num_bytes = wbytes;
    at[0].val.accessPolicyWindow.hitRatio = 1.0f;
    at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = at; cfg.numAttrs = 1;
  }
  CK(cudaLaunchKernelEx(&cfg, chain<HINT, HAS_C>, (const float4*)a, (const float4*)b, (const float4*)c, (float4*)out,
                        n / 4 / 4096, perm, pa, pc, po));
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : (1L << 24);  // floats per field: 256^3
  const size_t fb = n * sizeof(float);
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("device %s  L2 %.1f MB  persistingL2CacheMaxSize %.1f MB  accessPolicyMaxWindowSize %.1f MB  SMs %d\n", pr.name,
         pr.l2CacheSize / 1e6, pr.persistingL2CacheMaxSize / 1e6, pr.accessPolicyMaxWindowSize / 1e6, pr.multiProcessorCount);
  float *x, *k, *acc, *w, *flush;
  CK(cudaMalloc(&x, fb)); CK(cudaMalloc(&k, fb)); CK(cudaMalloc(&acc, fb)); CK(cudaMalloc(&w, fb));
  CK(cudaMalloc(&flush, 512u << 20));
  CK(cudaMemset(x, 0, fb)); CK(cudaMemset(k, 0, fb)); CK(cudaMemset(acc, 0, fb)); CK(cudaMemset(w, 0, fb));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t ev[4]; for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&ev[i]));
  const int grid = pr.multiProcessorCount * 8;
  const int reps = 20;
  for (int perm = 0; perm < 2; ++perm) {
    for (int variant = 0; variant < 5; ++variant) {
      size_t aside = (variant == 2 || variant == 3) ? (size_t)pr.persistingL2CacheMaxSize : 0;
      CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, aside));
      CK(cudaCtxResetPersistingL2Cache());
      double t[3] = {0, 0, 0};
      for (int r = -3; r < reps; ++r) {
        const bool hint = variant == 1 || variant == 2 || variant == 4;
        const int keep = variant == 4 ? P_NORMAL : P_LAST;
        const void* w1 = variant == 3 ? acc : nullptr;
        const void* w3 = variant == 3 ? w : nullptr;
        CK(cudaEventRecord(ev[0], st));
        if (hint) launch<true, false>(st, grid, x, k, nullptr, acc, n, 0, P_FIRST, P_NORMAL, keep, nullptr, 0);
        else launch<false, false>(st, grid, x, k, nullptr, acc, n, 0, 0, 0, 0, w1, fb);
        CK(cudaEventRecord(ev[1], st));
        if (hint) launch<true, true>(st, grid, x, k, acc, acc, n, perm, P_FIRST, keep, keep, nullptr, 0);
        else launch<false, true>(st, grid, x, k, acc, acc, n, perm, 0, 0, 0, w1, fb);
        CK(cudaEventRecord(ev[2], st));
        if (hint) launch<true, true>(st, grid, x, k, acc, w, n, perm, P_FIRST, variant == 4 ? P_NORMAL : P_FIRST, keep, nullptr, 0);
        else launch<false, true>(st, grid, x, k, acc, w, n, perm, 0, 0, 0, w3, fb);
        CK(cudaEventRecord(ev[3], st));
        CK(cudaStreamSynchronize(st));
        if (r >= 0) for (int i = 0; i < 3; ++i) { float ms; CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); t[i] += ms; }
      }
      const double F = fb / 1e6;  // MB
      printf("perm %d variant %d  K1(3F) %.1f us %.0f GB/s | K2(4F) %.1f us %.0f GB/s | K3(4F) %.1f us %.0f GB/s | chain %.1f us\n",
             perm, variant, 1e3 * t[0] / reps, 3 * F / (t[0] / reps), 1e3 * t[1] / reps, 4 * F / (t[1] / reps),
             1e3 * t[2] / reps, 4 * F / (t[2] / reps), 1e3 * (t[0] + t[1] + t[2]) / reps);
    }
  }
  CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
  return 0;
}
