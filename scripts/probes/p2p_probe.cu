// p2p_probe.cu -- measurement only (not part of the library): what NVLink gives SM-issued
// peer reads / writes for the access shapes the slab x sweeps use (128-byte row pieces, 1 MiB
// apart) versus wider pieces and a contiguous copy.  One process, two GPUs, peer access enabled.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_probe p2p_probe.cu && ./p2p_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// copy `nrows` row pieces of `piece` bytes; piece p of row r sits at base + r*stride + col*piece
template <int UNROLL>
__global__ void copy_pieces(char* __restrict__ dst, const char* __restrict__ src, long nrows, int piece, long stride, int ncols) {
  const int cpp = piece / 16;                       // 16-byte chunks per piece
  const long total = nrows * ncols * cpp;           // all chunks
  const long nthr = (long)gridDim.x * blockDim.x;
  for (long c0 = (long)blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += nthr * UNROLL) {
    int4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long c = c0 + u * nthr;
      if (c < total) {
        const long k = c % cpp, rc = c / cpp, r = rc % nrows, col = rc / nrows;
        v[u] = *reinterpret_cast<const int4*>(src + r * stride + col * piece + k * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long c = c0 + u * nthr;
      if (c < total) {
        const long k = c % cpp, rc = c / cpp, r = rc % nrows, col = rc / nrows;
        *reinterpret_cast<int4*>(dst + r * stride + col * piece + k * 16) = v[u];
      }
    }
  }
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  const long stride = 1 << 20;           // 1 MiB between rows (512^3 f32: n1*n2c*8)
  const long nrows = 256;                // rows per slab
  const size_t bytes = (size_t)nrows * stride;   // 256 MiB = one slab field
  char* buf[2][2];
  cudaStream_t st[2];
  cudaEvent_t e0[2], e1[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&buf[d][i], bytes)); CK(cudaMemset(buf[d][i], d + 1, bytes)); }
    CK(cudaStreamCreate(&st[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
  }
  const int pieces[] = {128, 256, 512, 2048, 1 << 20};
  const char* modes[] = {"pull (peer -> local)", "push (local -> peer)", "local -> local", "pull both GPUs at once", "push both GPUs at once",
                         "pull + push on one GPU (2 streams)"};
  for (int mode = 0; mode < 6; ++mode) {
    printf("%s\n", modes[mode]);
    for (int pi = 0; pi < 5; ++pi) {
      const int piece = pieces[pi];
      const int ncols = (int)(stride / piece);
      for (int grid_mult = 1; grid_mult <= 8; grid_mult *= 4) {
        const int blocks = 148 * grid_mult, threads = 512;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          const bool both = (mode == 3 || mode == 4);
          for (int d = 0; d < (both ? 2 : 1); ++d) {
            CK(cudaSetDevice(d));
            CK(cudaEventRecord(e0[d], st[d]));
            char* dst; const char* src;
            if (mode == 0 || mode == 3) { dst = buf[d][0]; src = buf[1 - d][1]; }
            else if (mode == 1 || mode == 4) { dst = buf[1 - d][0]; src = buf[d][1]; }
            else if (mode == 2) { dst = buf[d][0]; src = buf[d][1]; }
            else { dst = buf[d][0]; src = buf[1 - d][1]; }
            copy_pieces<4><<<blocks, threads, 0, st[d]>>>(dst, src, nrows, piece, stride, ncols);
            if (mode == 5) {  // concurrent push from a second stream (default stream of device 0)
              copy_pieces<4><<<blocks, threads, 0, 0>>>(buf[1][0], buf[0][1], nrows, piece, stride, ncols);
            }
            CK(cudaEventRecord(e1[d], st[d]));
          }
          float worst = 0;
          for (int d = 0; d < (both ? 2 : 1); ++d) {
            CK(cudaSetDevice(d));
            CK(cudaEventSynchronize(e1[d]));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            if (ms > worst) worst = ms;
          }
          if (worst < best) best = worst;
        }
        printf("  piece %7d B  grid %4d x %d: %8.1f us  %7.1f GB/s per direction per GPU\n", piece, blocks, threads, best * 1e3,
               bytes / (best * 1e-3) / 1e9);
      }
    }
  }
  // copy engine
  CK(cudaSetDevice(0));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0[0], st[0]));
    CK(cudaMemcpyPeerAsync(buf[0][0], 0, buf[1][1], 1, bytes, st[0]));
    CK(cudaEventRecord(e1[0], st[0]));
    CK(cudaEventSynchronize(e1[0]));
    float ms; CK(cudaEventElapsedTime(&ms, e0[0], e1[0]));
    if (rep == 2) printf("cudaMemcpyPeerAsync 256 MiB: %.1f us  %.1f GB/s\n", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
  }
  return 0;
}
