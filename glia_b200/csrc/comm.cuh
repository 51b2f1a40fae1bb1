// comm.cuh -- device-side peer communication of the slab-decomposed (multi-GPU) path.
//
// The reference distributes the grid with AccFFT: every 3-D FFT crosses ranks through MPI
// all-to-all transposes and every PETSc VecDot/VecNorm is an MPI_Allreduce
// (src/grad/SpectralOperators.cpp:80,96,169,253,398-421; KSPCG inside
// src/pde/DiffusionSolver.cpp:241).  Here one process drives one GPU, every rank maps the
// other ranks' field arenas (CUDA IPC over NVLink / NVSwitch), and
//   * the x-axis sweeps read and write their rows directly in the owners' HBM
//     (sweeps_dist.cuh) -- transform and exchange are one kernel;
//   * ranks order themselves with flag barriers living in those arenas;
//   * the PCG scalars are all-reduced by the scalar kernels themselves, summed in rank
//     order so that every rank takes bit-identical control decisions.
#pragma once
#include "simt.h"

namespace glia {

static constexpr int MAX_RANKS = 8;
static constexpr int RED_NV = 4;  // doubles per reduction slot

// view of the communication block of every rank's arena
struct Comm {
  int G = 1, rank = 0;
  unsigned* flags[MAX_RANKS] = {};  // flags[q][s]: last epoch rank s announced to rank q
  double* red[MAX_RANKS] = {};      // red[q][(parity*MAX_RANKS + s)*RED_NV + i]
  int* err = nullptr;               // local: set when a wait timed out
  int* done = nullptr;              // local: the PCG's `done` flag, raised on a time-out so that later kernels do no work
  unsigned long long timeout_ns = 120ull * 1000000000ull;  // wall clock (GLIA_RD_PEER_TIMEOUT_S)
};

#if defined(GLIA_SIMT_EMU)
__device__ inline void st_release_sys(unsigned* p, unsigned v) { std::atomic_ref<unsigned>(*p).store(v, std::memory_order_release); }
__device__ inline unsigned ld_acquire_sys(const unsigned* p) {
  return std::atomic_ref<unsigned>(*const_cast<unsigned*>(p)).load(std::memory_order_acquire);
}
__device__ inline void st_sys(double* p, double v) { std::atomic_ref<double>(*p).store(v, std::memory_order_relaxed); }
__device__ inline double ld_sys(const double* p) {
  return std::atomic_ref<double>(*const_cast<double*>(p)).load(std::memory_order_relaxed);
}
__device__ inline void fence_sys() { std::atomic_thread_fence(std::memory_order_seq_cst); }
__device__ inline void spin_pause() { std::this_thread::yield(); }
__device__ inline unsigned long long now_ns() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
#else
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_sys() { __threadfence_system(); }
__device__ __forceinline__ void spin_pause() { __nanosleep(64); }
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#endif

// thread t < G: announce `epoch` to rank t, then wait until rank t has announced it to us.
// A rank that never shows up (crashed peer) ends the wait after `timeout_ns` of WALL CLOCK (host skew
// between ranks -- per-rank file reads, first-call module loads -- is legitimate and can be seconds), raises
// the local error flag, which Engine::sync() turns into an error return of the call that was running, and
// raises the PCG's done flag so that the kernels already enqueued behind it do not compute on partial data.
__device__ inline void peer_signal_wait(const Comm& c, unsigned epoch, int t) {
  fence_sys();
  st_release_sys(c.flags[t] + c.rank, epoch);
  const unsigned* mine = c.flags[c.rank] + t;
  unsigned long long t0 = 0;
  unsigned polls = 0;
  while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
    spin_pause();
    if ((++polls & 1023u) == 0) {  // look at the clock every ~1000 polls
      const unsigned long long now = now_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns) {
        *c.err = 1;
        if (c.done) *c.done = 1;
        break;
      }
    }
  }
}

// Rank gate folded into a kernel (the slab x sweeps and the kernels that consume their results), so that the hot
// loop needs no stand-alone k_peer_barrier launches (13 us each, 4 per PCG iteration in round 1):
//   entry  -- CTA 0 announces `signal_in` to every peer (everything this rank enqueued before the kernel is complete:
//             the kernel is launched without the programmatic-launch attribute, or has passed its pdl_wait);
//             every CTA then waits until every peer has announced `wait_epoch`;
//   exit   -- the last CTA of the grid to finish (ticket counter) announces `signal_out`: all of this rank's writes
//             into the peers' memory are complete and visible.
// All three are optional (epoch 0 = none); G <= 1 makes the gate a no-op.
struct PeerGate {
  Comm comm;
  unsigned signal_in = 0, wait_epoch = 0, signal_out = 0;
  unsigned* ticket = nullptr;  // device counter, zero between uses
};
// all threads of the CTA call this once, before touching peer-written or peer-read data
__device__ inline void gate_enter(const PeerGate& g) {
  if (g.comm.G <= 1) return;
  const int t = threadIdx.x;
  if (g.signal_in && blockIdx.x == 0 && t < g.comm.G) {
    fence_sys();
    st_release_sys(g.comm.flags[t] + g.comm.rank, g.signal_in);
  }
  if (g.wait_epoch && t < g.comm.G) {
    const unsigned* mine = g.comm.flags[g.comm.rank] + t;
    unsigned long long t0 = 0;
    unsigned polls = 0;
    while ((int)(ld_acquire_sys(mine) - g.wait_epoch) < 0) {
      spin_pause();
      if ((++polls & 1023u) == 0) {
        const unsigned long long now = now_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > g.comm.timeout_ns) {
          *g.comm.err = 1;
          if (g.comm.done) *g.comm.done = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
}
// all threads of the CTA call this once, after their last store
__device__ inline void gate_exit(const PeerGate& g) {
  if (g.comm.G <= 1 || !g.signal_out) return;
  __syncthreads();  // every thread's stores precede thread 0's fence (the grid-sync pattern)
  if (threadIdx.x == 0) {
    fence_sys();
#if defined(GLIA_SIMT_EMU)
    const unsigned prev = std::atomic_ref<unsigned>(*g.ticket).fetch_add(1u, std::memory_order_acq_rel);
#else
    const unsigned prev = atomicAdd(g.ticket, 1u);
#endif
    if (prev == gridDim.x - 1) {
      fence_sys();
      *g.ticket = 0;
      fence_sys();
      for (int q = 0; q < g.comm.G; ++q) st_release_sys(g.comm.flags[q] + g.comm.rank, g.signal_out);
    }
  }
}

// stand-alone consumer side of a gate, for the kernels that carry none (one-tile-per-CTA fallbacks)
static __global__ void k_gate_wait(PeerGate g) { gate_enter(g); }

// all ranks: everything enqueued before the barrier on every rank is complete and visible
static __global__ void k_peer_barrier(Comm c, unsigned epoch) {
  const int t = threadIdx.x;
  if (t < c.G) peer_signal_wait(c, epoch, t);
}

// v <- sum over ranks of v, in rank order, identical on every rank.  Called by all threads of a
// CTA (>= MAX_RANKS threads); `v` must hold the same values in every thread on entry.
template <int NV>
__device__ inline void peer_allreduce(const Comm& c, unsigned epoch, unsigned seq, double (&v)[NV]) {
  static_assert(NV <= RED_NV, "slot size");
  if (c.G <= 1) return;
  __shared__ double sh_red[RED_NV];
  const int t = threadIdx.x;
  const int slot = (int)(seq & 1u) * MAX_RANKS;
  if (t < c.G) {
    double* dst = c.red[t] + (size_t)(slot + c.rank) * RED_NV;
    GLIA_UNROLL
    for (int i = 0; i < NV; ++i) st_sys(dst + i, v[i]);
    peer_signal_wait(c, epoch, t);
  }
  __syncthreads();
  if (t == 0) {
    GLIA_UNROLL
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int q = 0; q < c.G; ++q) s += ld_sys(c.red[c.rank] + (size_t)(slot + q) * RED_NV + i);
      sh_red[i] = s;
    }
  }
  __syncthreads();
  GLIA_UNROLL
  for (int i = 0; i < NV; ++i) v[i] = sh_red[i];
  __syncthreads();
}

}  // namespace glia
