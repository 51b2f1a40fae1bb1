// engine.cuh -- host side of libglia_rd: buffers, sweep launch geometry, the PETSc-
// semantics PCG loop, Strang time stepping with histories, gradient integrals.
// Mirrors (does not copy) the control flow of the reference classes:
//   DiffusionSolver   src/pde/DiffusionSolver.cpp:5-250
//   PdeOperatorsRD    src/pde/PdeOperators.cpp:140-420
//   DerivativeOperators::gradDiffusion/gradReaction  src/grad/DerivativeOperators.cpp:189-321
#pragma once
#include <string>
#include <type_traits>
#include <vector>

#include "engine_base.h"
#include "netcdf_io.h"
#include "pointwise.cuh"
#include "phi.cuh"
#include "sweeps.cuh"
#include "sweeps_dist.cuh"
#include "sweeps_pipe.cuh"
#include "sweeps_zpipe.cuh"
#include "sweeps_tma.cuh"

namespace glia {

// ------------------------------------------------------------------ runtime ----
namespace rt {
#if !defined(GLIA_SIMT_EMU)  // the emulator build supplies rt:: from tests/emu/simt_emu.h
inline int dev_malloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline int host_malloc(void** p, size_t n) { return (int)cudaMallocHost(p, n ? n : 1); }
inline void host_free(void* p) { if (p) cudaFreeHost(p); }
inline int copy(void* d, const void* s, size_t n, cudaStream_t st) { return (int)cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st); }
inline int h2d(void* d, const void* s, size_t n, cudaStream_t st) { return (int)cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st); }
inline int d2h(void* d, const void* s, size_t n, cudaStream_t st) { return (int)cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st); }
inline int zero(void* d, size_t n, cudaStream_t st) { return (int)cudaMemsetAsync(d, 0, n, st); }
inline int sync(cudaStream_t st) { return (int)cudaStreamSynchronize(st); }
inline int set_device(int d) { return (int)cudaSetDevice(d); }
inline bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
// inter-process shareable device memory (CUDA IPC): the slab ranks map each other's arenas
inline int ipc_alloc(void** p, size_t n, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  int e = (int)cudaMalloc(p, n ? n : 1);
  if (e) return e;
  return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle), *p);
}
inline int ipc_open(void** p, const unsigned char handle[64], size_t) {
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, 64);
  return (int)cudaIpcOpenMemHandle(p, hd, cudaIpcMemLazyEnablePeerAccess);
}
inline void ipc_close(void* p, size_t) { if (p) cudaIpcCloseMemHandle(p); }
inline void ipc_free(void* p, size_t, const unsigned char*) { if (p) cudaFree(p); }
inline int sm_count(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
  return v;
}
// the handle's stream gets the highest priority so that, when a side stream (below) has work in flight,
// CTAs of the main stream's kernels are placed first on SMs as they free up
inline int stream_create(cudaStream_t* s) {
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); least = greatest = 0; }
  return (int)cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest);
}
// fork / join of a lowest-priority side stream by events (no host synchronisation)
struct Fork {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int create() {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); least = 0; }
    int e = (int)cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, least);
    if (e) return e;
    if ((e = (int)cudaEventCreateWithFlags(&fork, cudaEventDisableTiming))) return e;
    return (int)cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
  }
  void destroy() {
    if (fork) cudaEventDestroy(fork);
    if (join) cudaEventDestroy(join);
    if (side) cudaStreamDestroy(side);
    fork = join = nullptr; side = nullptr;
  }
  void begin(cudaStream_t main) { cudaEventRecord(fork, main); cudaStreamWaitEvent(side, fork, 0); }
  void end(cudaStream_t main) { cudaEventRecord(join, side); cudaStreamWaitEvent(main, join, 0); }
};
inline void stream_destroy(cudaStream_t s) { cudaStreamDestroy(s); }
// `consumer` waits for everything enqueued so far on `producer` (an event, no host synchronisation)
inline int stream_wait_stream(cudaStream_t consumer, cudaStream_t producer) {
  cudaEvent_t e;
  int rc = (int)cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  if (rc) return rc;
  rc = (int)cudaEventRecord(e, producer);
  if (!rc) rc = (int)cudaStreamWaitEvent(consumer, e, 0);
  cudaEventDestroy(e);  // released once the recorded work has completed
  return rc;
}
inline const char* err_string(int e) { return cudaGetErrorString((cudaError_t)e); }
// per-launch CUDA-event profiler (off by default: zero overhead besides one branch)
struct Profiler {
  struct Rec { const char* tag; cudaEvent_t a, b; };
  bool on = false;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
  int before(const char* tag, cudaStream_t s) {
    if (!on) return -1;
    Rec r{tag, get(), get()};
    cudaEventRecord(r.a, s);
    recs.push_back(r);
    return (int)recs.size() - 1;
  }
  void after(int slot, cudaStream_t s) { if (slot >= 0) cudaEventRecord(recs[slot].b, s); }
  void begin() { recs.clear(); used = 0; on = true; }
  // aggregate by tag: "tag count total_ms\n" lines
  std::string end(cudaStream_t s) {
    on = false;
    cudaStreamSynchronize(s);
    std::vector<std::string> names; std::vector<double> tot; std::vector<long> cnt;
    for (auto& r : recs) {
      float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
      size_t i = 0;
      for (; i < names.size(); ++i) if (names[i] == r.tag) break;
      if (i == names.size()) { names.push_back(r.tag); tot.push_back(0); cnt.push_back(0); }
      tot[i] += ms; cnt[i]++;
    }
    std::string out;
    char buf[256];
    for (size_t i = 0; i < names.size(); ++i) {
      std::snprintf(buf, sizeof buf, "%s %ld %.6f\n", names[i].c_str(), cnt[i], tot[i]);
      out += buf;
    }
    recs.clear(); used = 0;
    return out;
  }
  void destroy() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};
struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  void create() { cudaEventCreate(&a); cudaEventCreate(&b); }
  void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); a = b = nullptr; }
  void start(cudaStream_t st) { cudaEventRecord(a, st); }
  double stop_ms(cudaStream_t st) {
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
};
#endif
}  // namespace rt


#define GLIA_CHECK(expr)                                                                       \
  do {                                                                                         \
    int _e = (expr);                                                                           \
    if (_e != 0) throw EngineError{std::string(#expr) + ": " + rt::err_string(_e)};            \
  } while (0)

#define GLIA_DISPATCH_N(nval, ...)                                         \
  switch (nval) {                                                          \
    case 32:  { constexpr int N = 32;  __VA_ARGS__; } break;               \
    case 64:  { constexpr int N = 64;  __VA_ARGS__; } break;               \
    case 128: { constexpr int N = 128; __VA_ARGS__; } break;               \
    case 256: { constexpr int N = 256; __VA_ARGS__; } break;               \
    case 512: { constexpr int N = 512; __VA_ARGS__; } break;               \
    default: throw EngineError{"unsupported line length " + std::to_string(nval)}; \
  }


template <typename T> class Engine;
template <typename T> void fft3d_r2c(Engine<T>& E, const T* f, cplx<T>* fhat);
template <typename T> void fft3d_c2r(Engine<T>& E, const cplx<T>* fhat, T* f);

template <typename T>
class Engine : public EngineBase {
 public:
  using C = cplx<T>;
  int n[3];
  // slab decomposition along x (G = 1: the whole grid): this rank owns x-planes
  // [rank*n0l, (rank+1)*n0l) and sweeps the x lines with y in [rank*n1l, (rank+1)*n1l)
  int G = 1, rank = 0, n0l, n1l;
  int device = 0;     // CUDA ordinal this handle lives on; every entry point makes it current (make_current)
  long nreal, ncplx;  // LOCAL element counts of the whole handle (all ensemble members)
  int nb = 1;         // ensemble members carried by this handle, fields laid out [nb][n0][n1][n2] (fft_core.cuh)
  long nreal_m;       // elements of one member
  int n2c;            // n2/2: complex columns of the pair view
  cudaStream_t st = 0;
  rt::Timer timer;
  rt::Profiler prof;

  // one arena holds every field an x sweep of another rank may touch (G > 1: IPC-shared)
  char* arena = nullptr;
  size_t arena_bytes = 0;
  unsigned char arena_handle[64] = {};
  char* peer_arena[MAX_RANKS] = {};
  bool connected = false;
  char* hist_arena = nullptr;
  size_t hist_bytes = 0;
  unsigned char hist_handle[64] = {};
  char* peer_hist[MAX_RANKS] = {};
  bool hist_connected = false;
  Comm comm;
  unsigned epoch = 0, rseq = 0;
  unsigned* ticket = nullptr;   // device counter of the in-kernel rank gates (PeerGate)
  unsigned gate_pending = 0;    // epoch the last gated x sweep announces on exit; its consumer waits for it
  int nsm = 148;         // SMs of this device: grid size of the persistent (pipelined) sweeps
  // slab D-apply: run the rank-local z sweep on a side stream (into acc2) WHILE the peer x sweep, which is
  // NVLink-bound and leaves SM time unused, runs on a share of the SMs; the y sweep then takes acc + acc2
  // (the same two addends the serial order sums, so the result is bit-identical).  GLIA_RD_XZ=0/1.
  // Measured at 2 GPUs, 512^3 (gpurun_out/r1q_*): 9.17 time-steps/s serial, 9.48 with the x sweep on 55 % of the
  // SMs, 9.60 on 74 %.
  bool use_xz = true;
  int xz_ctas = 0;       // CTAs of the x sweep while overlapped (GLIA_RD_XZ_CTAS; 0 = 75 % of the persistent grid)
  rt::Fork fork;
  T* acc2 = nullptr;
  T* z_out_override = nullptr;
  bool z_side = false, pdl_hold = false;
  int dist_debug = 0;    // GLIA_RD_DIST_DEBUG: timing experiments only (results are WRONG): 1 = x sweeps read
                         // local rows instead of peer rows, 2 = write local rows instead of peer rows

  // coefficients (kT, ktilT: pencil copies for the distributed x sweeps)
  T *kf = nullptr, *ktil = nullptr, *rho = nullptr, *kT = nullptr, *ktilT = nullptr;
  T kavg[3] = {0, 0, 0};
  T k_scale = (T)1e-2;
  T dt_ctx;
  PcSym<T> sym;                  // member 0 (the only one unless this is an ensemble handle)
  std::vector<T> kavg_m;         // ensemble: k-bar of every member (3 each); used when kavg_per_member
  bool kavg_per_member = false;
  PcSym<T>* syms_d = nullptr;    // device copy of every member's frozen symbol
  std::vector<int> its_m, its_acc;  // per-member iteration counts: last solve / accumulated by the time loops
  // KSP settings (DiffusionSolver.cpp:27)
  double rtol = 1e-6, abstol = 1e-50, dtol = 1e4;
  int maxit = 5000;
  // PCG work
  T *b = nullptr, *r = nullptr, *z = nullptr, *p = nullptr, *w = nullptr, *acc = nullptr;
  C* shat = nullptr;
  C* tw[3] = {nullptr, nullptr, nullptr};
  double *partial = nullptr, *scal = nullptr;
  int* iscal = nullptr;
  int* h_iscal = nullptr;     // pinned (I_NISCAL ints, then one more: the comm error flag read by sync())
  double* h_out = nullptr;    // pinned
  long npart = 0;             // doubles per partial region
  int its_guess = 2;
  // time stepping
  int nt = 0;
  T dt = 0;
  int order = 2;               // params->tu_->order_: 2 = Strang, 1 = diffusion(dt) then reaction
  T *d0 = nullptr, *obs0 = nullptr;  // two_time_points_: data and observation mask at t = 0 (owned copies)
  bool two_snap = false, has_obs0 = false;
  T *c_hist = nullptr, *p_hist = nullptr, *chalf_hist = nullptr;
  T *c_t = nullptr, *p_0 = nullptr, *work11 = nullptr, *Tk = nullptr, *Tr = nullptr;
  T *stage = nullptr, *TkX = nullptr;  // G > 1 only
  // host staging for the host-buffer entry point
  T *hs_in = nullptr, *hs_out = nullptr;
  // Weierstrass smoother symbols (one real table per axis) and the Phi basis
  T* symtab[3] = {nullptr, nullptr, nullptr};
  double sym_sigma = -1.0;
  T* phi_filter = nullptr;
  bool phi_has_filter = false;
  std::vector<double> phi_centers;
  int phi_np = 0;
  T phi_sigma = (T)0;
  double phi_sigma_smooth = 0.0;
  double* phi_dots = nullptr;
  int phi_dots_cap = 0;

  int precision() const override { return (int)sizeof(T); }

  Engine(const int nn[3], int device_, double dt_ctx_, int rank_ = 0, int nranks_ = 1, int nbatch_ = 1) {
    for (int i = 0; i < 3; ++i) {
      n[i] = nn[i];
      if (!(n[i] == 32 || n[i] == 64 || n[i] == 128 || n[i] == 256 || n[i] == 512))
        throw EngineError{"grid sizes must be powers of two in [32, 512]"};
    }
    G = nranks_;
    rank = rank_;
    if (!(G == 1 || G == 2 || G == 4 || G == 8) || rank < 0 || rank >= G)
      throw EngineError{"slab ranks: nranks must be 1, 2, 4 or 8 and 0 <= rank < nranks"};
    n0l = n[0] / G;
    n1l = n[1] / G;
    if (n0l < 1 || n1l < 1 || (long)n0l * n[1] < 2) throw EngineError{"grid too small for this many slabs"};
    device = device_;
    GLIA_CHECK(rt::set_device(device));
    GLIA_CHECK(rt::stream_create(&st));
    nsm = rt::sm_count(device);
    if (const char* e = std::getenv("GLIA_RD_PDL")) use_pdl = std::atoi(e) != 0;
#if !defined(GLIA_SIMT_EMU)
    if (const char* e = std::getenv("GLIA_RD_TMA")) use_tma = std::atoi(e) != 0 && nranks_ == 1;
#endif
    if (const char* e = std::getenv("GLIA_RD_PDL_SLAB")) use_pdl_slab = std::atoi(e) != 0;
    if (const char* e = std::getenv("GLIA_RD_DIST_DEBUG")) dist_debug = std::atoi(e);
    if (const char* e = std::getenv("GLIA_RD_XZ")) use_xz = std::atoi(e) != 0;
    if (const char* e = std::getenv("GLIA_RD_XZ_CTAS")) xz_ctas = std::atoi(e);
    if (G == 1 || (sizeof(T) == 8 && (n[0] > 256 || n[1] > 256))) use_xz = false;
    if (use_xz) GLIA_CHECK(fork.create());
    timer.create();
    nb = nbatch_;
    if (nb < 1 || nb > 4096) throw EngineError{"ensemble size must be in [1, 4096]"};
    if (nb > 1 && G > 1) throw EngineError{"an ensemble handle cannot be slab-decomposed (members are independent: spread them over the GPUs)"};
    nreal_m = (long)n0l * n[1] * n[2];
    nreal = nreal_m * nb;
    ncplx = nreal / 2;
    n2c = n[2] / 2;
    dt_ctx = (T)dt_ctx_;
    std::vector<T**> fields = {&kf, &ktil, &rho, &b, &r, &z, &p, &w, &acc, &c_t, &p_0, &work11, &Tk, &Tr};
    if (G > 1) { fields.push_back(&kT); fields.push_back(&ktilT); fields.push_back(&stage); fields.push_back(&TkX); }
    if (use_xz) fields.push_back(&acc2);  // same switch on every rank: the arenas keep one layout
    const size_t fbytes = sizeof(T) * nreal;  // a multiple of 256 bytes for every admissible grid
    const size_t comm_bytes = 4096;
    arena_bytes = fbytes * (fields.size() + 1) + comm_bytes;
    if (G > 1) GLIA_CHECK(rt::ipc_alloc((void**)&arena, arena_bytes, arena_handle));
    else GLIA_CHECK(rt::dev_malloc((void**)&arena, arena_bytes));
    GLIA_CHECK(rt::zero(arena, arena_bytes, st));
    size_t off = 0;
    for (T** f : fields) { *f = reinterpret_cast<T*>(arena + off); off += fbytes; }
    shat = reinterpret_cast<C*>(arena + off); off += fbytes;
    // communication block: flags[MAX_RANKS] then the reduction slots
    comm.G = G;
    comm.rank = rank;
    peer_arena[rank] = arena;
    bind_comm(off);
    for (int a = 0; a < 3; ++a) {
      std::vector<C> tab(n[a]);
      for (int j = 0; j < n[a]; ++j) {
        const double ang = -2.0 * M_PI * (double)j / (double)n[a];
        tab[j] = {(T)std::cos(ang), (T)std::sin(ang)};
      }
      GLIA_CHECK(rt::dev_malloc((void**)&tw[a], sizeof(C) * n[a]));
      GLIA_CHECK(rt::h2d(tw[a], tab.data(), sizeof(C) * n[a], st));
      GLIA_CHECK(rt::sync(st));
    }
    // partial-sum regions: enough for the largest grid of any reducing kernel
    npart = 4 * (nreal / 256 + 1024);
    GLIA_CHECK(rt::dev_malloc((void**)&partial, sizeof(double) * npart * 3));
    GLIA_CHECK(rt::dev_malloc((void**)&scal, sizeof(double) * S_NSCAL * nb));
    GLIA_CHECK(rt::dev_malloc((void**)&iscal, sizeof(int) * I_NISCAL * nb));
    GLIA_CHECK(rt::dev_malloc((void**)&syms_d, sizeof(PcSym<T>) * nb));
    GLIA_CHECK(rt::dev_malloc((void**)&ticket, sizeof(unsigned) * 4));
    GLIA_CHECK(rt::zero(ticket, sizeof(unsigned) * 4, st));
    GLIA_CHECK(rt::zero(scal, sizeof(double) * S_NSCAL * nb, st));
    GLIA_CHECK(rt::zero(iscal, sizeof(int) * I_NISCAL * nb, st));
    comm.err = iscal + I_COMM_ERR;
    comm.done = iscal + I_DONE;
    if (const char* e = std::getenv("GLIA_RD_PEER_TIMEOUT_S")) {
      const double sec = std::atof(e);
      if (sec > 0) comm.timeout_ns = (unsigned long long)(sec * 1e9);
    }
    GLIA_CHECK(rt::host_malloc((void**)&h_iscal, sizeof(int) * (I_NISCAL * nb + 1)));
    h_iscal[I_NISCAL * nb] = 0;
    kavg_m.assign(3 * (size_t)nb, (T)0);
    its_m.assign(nb, 0);
    its_acc.assign(nb, 0);
    GLIA_CHECK(rt::host_malloc((void**)&h_out, sizeof(double) * 16));
    sym = PcSym<T>{dt_ctx, (T)0, (T)0, (T)0, (T)(1.0 / ((double)n[0] * n[1] * n[2]))};
    upload_syms();
    GLIA_CHECK(rt::sync(st));
  }
  ~Engine() override {
    rt::set_device(device);
    rt::sync(st);
    for (int q = 0; q < G; ++q) {
      if (q == rank) continue;
      if (peer_arena[q]) rt::ipc_close(peer_arena[q], arena_bytes);
      if (peer_hist[q]) rt::ipc_close(peer_hist[q], hist_bytes);
    }
    if (G > 1) { rt::ipc_free(arena, arena_bytes, arena_handle); rt::ipc_free(hist_arena, hist_bytes, hist_handle); }
    else { rt::dev_free(arena); rt::dev_free(hist_arena); }
    for (int a = 0; a < 3; ++a) rt::dev_free(tw[a]);
    rt::dev_free(partial); rt::dev_free(scal); rt::dev_free(iscal); rt::dev_free(ticket); rt::dev_free(syms_d);
    rt::host_free(h_iscal); rt::host_free(h_out);
    rt::host_free(hs_in); rt::host_free(hs_out);
    for (int a = 0; a < 3; ++a) rt::dev_free(symtab[a]);
    rt::dev_free(phi_filter); rt::dev_free(phi_dots);
    rt::dev_free(d0); rt::dev_free(obs0);
    timer.destroy();
    prof.destroy();
    fork.destroy();
    rt::stream_destroy(st);
  }

  // ------------------------------------------------------- slab plumbing ----
  size_t comm_off = 0;
  void bind_comm(size_t off) {
    comm_off = off;
    for (int q = 0; q < G; ++q) {
      if (!peer_arena[q]) continue;
      comm.flags[q] = reinterpret_cast<unsigned*>(peer_arena[q] + off);
      comm.red[q] = reinterpret_cast<double*>(peer_arena[q] + off + 256);
    }
  }
  void ipc_export(int which, unsigned char out[64]) {
    if (G <= 1) throw EngineError{"ipc_export: not a slab handle"};
    if (which == 1 && !hist_arena) throw EngineError{"ipc_export: resize_history() first"};
    std::memcpy(out, which == 0 ? arena_handle : hist_handle, 64);
  }
  // handles: nranks * 64 bytes, rank order (every rank passes the same array)
  void ipc_connect(int which, const unsigned char* handles) {
    if (G <= 1) throw EngineError{"ipc_connect: not a slab handle"};
    sync();
    char** peers = which == 0 ? peer_arena : peer_hist;
    const size_t bytes = which == 0 ? arena_bytes : hist_bytes;
    for (int q = 0; q < G; ++q) {
      if (q == rank) continue;
      if (peers[q]) { rt::ipc_close(peers[q], bytes); peers[q] = nullptr; }
      void* m = nullptr;
      GLIA_CHECK(rt::ipc_open(&m, handles + 64 * (size_t)q, bytes));
      peers[q] = (char*)m;
    }
    if (which == 0) { bind_comm(comm_off); connected = true; }
    else hist_connected = true;
  }
  // Close this rank's mappings of the peers' arenas (which: 0 work arena, 1 histories, -1 both).  CUDA leaves
  // freeing an exported allocation that another process still has open undefined, so teardown (and every
  // re-allocation of the histories) is: every rank disconnects -> caller-side barrier -> free.
  void ipc_disconnect(int which) {
    if (G <= 1) return;
    sync();
    for (int q = 0; q < G; ++q) {
      if (q == rank) continue;
      if ((which == 0 || which < 0) && peer_arena[q]) { rt::ipc_close(peer_arena[q], arena_bytes); peer_arena[q] = nullptr; }
      if ((which == 1 || which < 0) && peer_hist[q]) { rt::ipc_close(peer_hist[q], hist_bytes); peer_hist[q] = nullptr; }
    }
    if (which == 0 || which < 0) connected = false;
    if (which == 1 || which < 0) hist_connected = false;
  }
  bool in_arena(const void* ptr) const {
    return (const char*)ptr >= arena && (const char*)ptr < arena + arena_bytes;
  }
  bool in_hist(const void* ptr) const {
    return hist_arena && (const char*)ptr >= hist_arena && (const char*)ptr < hist_arena + hist_bytes;
  }
  // the same field in every rank's arena (or history arena)
  PeerRows<T> rows(const T* f, int dbg_bit = 0) const {
    PeerRows<T> pr{};
    if (dbg_bit && (dist_debug & dbg_bit)) {
      for (int q = 0; q < G; ++q) pr.base[q] = reinterpret_cast<C*>(const_cast<T*>(f));
      return pr;
    }
    if (in_arena(f)) {
      if (!connected) throw EngineError{"slab handle is not connected (glia_rd_ipc_connect)"};
      const size_t off = (const char*)f - arena;
      for (int q = 0; q < G; ++q) pr.base[q] = reinterpret_cast<C*>(peer_arena[q] + off);
    } else if (in_hist(f)) {
      if (!hist_connected) throw EngineError{"slab histories are not connected (glia_rd_ipc_connect)"};
      const size_t off = (const char*)f - hist_arena;
      for (int q = 0; q < G; ++q) pr.base[q] = reinterpret_cast<C*>((q == rank ? hist_arena : peer_hist[q]) + off);
    } else {
      throw EngineError{"internal: x sweep on a field outside the shared arenas"};
    }
    return pr;
  }
  TileX tile_xd() const {
    int sh = 0;
    while ((1 << sh) < n0l) ++sh;
    return TileX{(long)n[1] * n2c, (long)n2c, (long)n1l * n2c, n2c / SL, ilog2(n2c / SL), n1l, rank * n1l, sh, n0l - 1};
  }
  static dim3 grid_xd(const TileX& g) { return dim3(g.nchunk * g.n_outer); }
  unsigned next_epoch() { return ++epoch; }
  // rank gates folded into the slab x sweeps (comm.cuh): the x sweep announces itself on entry, waits for every peer,
  // and its last CTA announces completion; the kernel that consumes what the peers wrote waits for that.
  PeerGate gate_none() const { PeerGate g; g.comm.G = 1; return g; }
  PeerGate gate_xsweep() {
    PeerGate g;
    g.comm = comm;
    g.signal_in = g.wait_epoch = next_epoch();
    g.signal_out = gate_pending = next_epoch();
    g.ticket = ticket;
    return g;
  }
  PeerGate gate_consumer() {
    PeerGate g;
    g.comm = comm;
    g.wait_epoch = gate_pending;
    gate_pending = 0;
    return g;
  }
  // for consumers without a gate of their own
  void gate_wait_kernel() {
    if (G > 1 && gate_pending) L("k_gate_wait", k_gate_wait, dim3(1), dim3(32), 0, st, gate_consumer());
  }
  void barrier() {
    if (G > 1) L("k_peer_barrier", k_peer_barrier, dim3(1), dim3(32), 0, st, comm, next_epoch());
  }
  void build_pencil(const T* slab_field, T* pencil) {
    barrier();
    L("k_slab_to_pencil", k_slab_to_pencil<T>, grid_pw(ncplx), dim3(256), 0, st, tile_xd(), n[0], rows(slab_field),
      (C*)pencil);
    barrier();
  }

  // every kernel launch of the engine goes through here: counted, and (when profiling is on)
  // bracketed by CUDA events on the engine's stream so that bench.py can report each kernel's
  // average duration live.
  template <class... KA, class... A>
  void L(const char* tag, void (*k)(KA...), dim3 g, dim3 b, size_t smem, cudaStream_t s, A... args) {
    const int slot = prof.before(tag, s);
    simt::launch(k, g, b, smem, s, args...);
    prof.after(slot, s);
    ++launches;
  }
  // programmatic dependent launch for the kernels that call pdl_wait() (single-GPU handles only: the
  // slab path orders its x sweeps with k_peer_barrier launches, which stay fully serialised).
  // Off while profiling: the bracketing events would serialise the launches anyway.
#if !defined(GLIA_SIMT_EMU)
  // GLIA_RD_TMA=1: the preconditioner's y sweeps through the TMA / mbarrier kernel (sweeps_tma.cuh) -- the A/B
  // experiment of round 2; single-GPU handles only.  The tensor map covers shat, whose address never changes.
  bool use_tma = false;
  bool tmap_ready = false;
  CUtensorMap tmap_shat;
  bool tma_map() {
    if (!tmap_ready) {
      const long words = (long)n[2] * (long)sizeof(T) / 4;
      const int box_words = (int)(sizeof(C) * SL / 4);
      int rows = n[1] > 256 ? 256 : n[1];
      if (!make_tile_map_y(&tmap_shat, (void*)shat, n0l, n[1], words, box_words, rows))
        throw EngineError{"GLIA_RD_TMA=1: cuTensorMapEncodeTiled is unavailable or rejected the tile map"};
      tmap_ready = true;
    }
    return true;
  }
#endif
  bool use_pdl = true;  // GLIA_RD_PDL=0 turns it off
  // slab handles: only the rank-local chains (z, y sweeps, scalar kernels, vector update) overlap; the
  // peer x sweeps and the k_peer_barrier launches around them go through L() and stay fully serialised,
  // and a kernel launched without the attribute waits for the complete predecessor whatever it triggered.
  bool use_pdl_slab = true;  // GLIA_RD_PDL_SLAB=0 (measured at 2 GPUs, 512^3: 9.09 -> 9.19 time-steps/s)
  template <class... KA, class... A>
  void LP(const char* tag, void (*k)(KA...), dim3 g, dim3 b, size_t smem, cudaStream_t s, A... args) {
#if defined(GLIA_SIMT_EMU)
    L(tag, k, g, b, smem, s, args...);
#else
    if (!use_pdl || (G > 1 && !use_pdl_slab) || prof.on || pdl_hold) return L(tag, k, g, b, smem, s, args...);
    simt::launch_pdl(k, g, b, smem, s, args...);
    ++launches;
#endif
  }
  void check_launch() {
    const char* e = simt::last_error();
    if (e) throw EngineError{std::string("kernel launch: ") + e};
  }
  // End of every entry point.  Slab handles also look at the peer-wait error flag here, so that a collective
  // that synchronises without solving (set_diffusion, gradient, ...) cannot return success after a rank
  // barrier timed out; the flag is cleared so that the handle stays usable once the peers are back.
  void sync() {
    if (G > 1) GLIA_CHECK(rt::d2h(h_iscal + I_NISCAL * nb, iscal + I_COMM_ERR, sizeof(int), st));
    GLIA_CHECK(rt::sync(st));
    check_launch();
    if (G > 1 && h_iscal[I_NISCAL * nb]) {
      h_iscal[I_NISCAL * nb] = 0;
      rt::zero(iscal + I_COMM_ERR, sizeof(int), st);
      rt::sync(st);
      throw EngineError{"slab peer did not arrive at a rank barrier (timed out after GLIA_RD_PEER_TIMEOUT_S)"};
    }
  }
  void make_current() { GLIA_CHECK(rt::set_device(device)); }

  // ------------------------------------------------------------ geometry ----
  TileS tile_y() const { return TileS{(long)n2c, (long)n[1] * n2c, n2c / SL, n0l, 0}; }
  TileS tile_x() const { return TileS{(long)n[1] * n2c, (long)n2c, n2c / SL, n[1], 0}; }
  LinesZ lines_z() const { return LinesZ{(long)nb * n0l * n[1] / 2}; }          // every member's line pairs
  LinesZ lines_z_member() const { return LinesZ{(long)n0l * n[1] / 2}; }
  template <int N> static size_t smem_s() { return sizeof(C) * N * SL; }
  template <int N> static size_t smem_s2() { return 2 * sizeof(C) * N * SL; }  // + the kept x tile
  template <int N> static size_t smem_z() { return sizeof(C) * zlines<N>() * zpad<N>(); }
  template <int N> static dim3 block_s() { return dim3(SL * (N / FftPlan<N>::E)); }
  static dim3 grid_s(const TileS& g) { return dim3(g.nchunk * g.n_outer); }
  // Z kernels, one group of line pairs per CTA: CTAs per member (bpm) x members; a CTA never straddles two members
  template <int N> int bpm_z() const {
    const long np = lines_z_member().npairs;
    return (int)((np + zlines<N>() - 1) / zlines<N>());
  }
  // persistent Z kernels with two resident CTAs per SM (kz_r2c_pipe, kz_c2r_pipe): groups of zlines<N>() line pairs
  // of one member, CTAs per member
  template <int N> int zgroups() const {
    const long np = lines_z_member().npairs;
    return (int)((np + zlines<N>() - 1) / zlines<N>());
  }
  template <int N> int cpm_zpipe2() const {
    int cap = nsm * 2 / (nb > 2 ? 2 : nb);
    if (cap < 1) cap = 1;
    const int g = zgroups<N>();
    return g < cap ? g : cap;
  }
  template <int N> dim3 grid_z() const {
    if (nb > 1 && lines_z_member().npairs % zlines<N>() != 0) throw EngineError{"ensemble handle: grid too small for the z sweeps"};
    return dim3((unsigned)(bpm_z<N>() * nb));
  }
  static dim3 grid_pw(long nelem) {
    long g = (nelem + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return dim3((unsigned)g);
  }
  double* part(int region) { return partial + (long)region * npart; }

  // ------------------------------------------------------------ sweeps ----
  // acc = Dz(k Dz x); acc += Dy(k Dy x); then the x sweep with epilogue EPI
  // z sweep of the D-apply: acc (+)= D_z(k D_z x)
  template <int ADD>
  void sweep_deriv2_z(const char* tag, const T* x, const T* kfield, const int* done) {
    T* const zo = z_out_override ? z_out_override : acc;
    const cudaStream_t zs = z_side ? fork.side : st;
    GLIA_DISPATCH_N(n[2], {
      if constexpr (zpipe_fits<T, N>()) {
        // persistent warp-private pipelined form (sweeps_zpipe.cuh; measured: 49.7 -> 45.8 us at 256^3,
        // 603 -> 526 us at 512^3, profiles/r2d_zpipe_twldg_ab.txt)
        const long np = lines_z_member().npairs;
        const int ngroups = (int)((np + zlines<N>() - 1) / zlines<N>());   // of one member
        int cap = nsm * zpipe_ctas<T, N>() / (nb > 2 ? 2 : nb);
        if (cap < 1) cap = 1;
        const int cpm = ngroups < cap ? ngroups : cap;
        LP(tag, kz_deriv2_pipe<T, N, ADD>, dim3((unsigned)(cpm * nb)), dim3(zthreads<N>()), zpipe_smem<T, N>(), zs,
           lines_z_member(), ngroups, x, kfield, zo, (const C*)tw[2], done, cpm);
      } else {
        if (nb > 1) throw EngineError{"ensemble handle: this z line length needs the pipelined sweep"};
        LP(tag, kz_deriv2<T, N, ADD, 1>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), zs, lines_z(), x, kfield, zo,
           (const C*)tw[2], done);
      }
    });
  }
  static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
  RowsS<T> rows_s(const TileS& g, const void* ptr) const {
    return RowsS<T>{(C*)const_cast<void*>(ptr), g.row_stride, g.outer_stride, ilog2(g.nchunk), nreal_m / 2};
  }
  // persistent S kernels: CTAs per member (ntiles = tiles of ONE member); the grid is cpm x members
  // (an ensemble's members converge at different iterations and the CTAs of a converged member leave at once: every
  // member gets CTAs for at least half of the machine, so that the last few unconverged members still fill it)
  template <int N> int cpm_pipe(int ntiles) const {
    int g = nsm * pipe_ctas<T, N>() / (nb > 2 ? 2 : nb);
    if (g < 1) g = 1;
    return ntiles < g ? ntiles : g;
  }
  template <int N> dim3 grid_pipe(int ntiles) const { return dim3((unsigned)(cpm_pipe<N>(ntiles) * nb)); }
  void upload_syms() {
    std::vector<PcSym<T>> h(nb, sym);
    if (kavg_per_member)
      for (int m = 0; m < nb; ++m) { h[m].kxx = kavg_m[3 * m]; h[m].kyy = kavg_m[3 * m + 1]; h[m].kzz = kavg_m[3 * m + 2]; }
    GLIA_CHECK(rt::h2d(syms_d, h.data(), sizeof(PcSym<T>) * nb, st));
    GLIA_CHECK(rt::sync(st));
  }
  // one S-geometry D(k D x) sweep on local rows; returns the number of partial-sum blocks
  template <int EPI>
  int sweep_deriv2_local(int nline, const char* tag, const TileS& g, const T* x, const T* kfield, T alpha, T* out1,
                         T* out2, double* pp, const int* done, const T* acc_extra = nullptr) {
    constexpr bool keep_x = (EPI == EPI_MATVEC || EPI == EPI_RHS);
    int nblk = 0;
    GLIA_DISPATCH_N(nline, {
      if constexpr (pipe_fits<T, N>()) {
        const int ntiles = g.nchunk * g.n_outer;   // of one member
        const dim3 gr = grid_pipe<N>(ntiles);
        const int cpm = cpm_pipe<N>(ntiles);
        nblk = cpm;                                // partial sums per member
        bool launched = false;
        // slab handles: this sweep consumes what the peers' x sweeps wrote into acc (gate left pending by dapply_dist)
        const PeerGate gate = (G > 1 && gate_pending) ? gate_consumer() : gate_none();
        if constexpr (EPI != EPI_ADD && EPI != EPI_SET) {
          if (acc_extra) {
            const RowsS2<T> a2{(C*)acc, (C*)const_cast<T*>(acc_extra), g.row_stride, g.outer_stride, ilog2(g.nchunk),
                               nreal_m / 2};
            LP(tag, ks_deriv2_pipe<T, N, EPI, RowsS<T>, RowsS<T>, RowsS2<T>, RowsS<T>>, gr, block_s<N>(), pipe_smem<T, N>(),
               st, ntiles, rows_s(g, x), rows_s(g, kfield), a2, rows_s(g, out1), rows_s(g, out2),
               (const C*)tw_for(nline, g), alpha, pp, done, gate, cpm);
            launched = true;
          }
        }
        if (!launched)
          LP(tag, ks_deriv2_pipe<T, N, EPI, RowsS<T>, RowsS<T>, RowsS<T>, RowsS<T>>, gr, block_s<N>(), pipe_smem<T, N>(), st,
             ntiles, rows_s(g, x), rows_s(g, kfield), rows_s(g, acc), rows_s(g, out1), rows_s(g, out2),
             (const C*)tw_for(nline, g), alpha, pp, done, gate, cpm);
      } else {  // three tiles exceed shared memory (double precision at 512 points): one tile per CTA
        gate_wait_kernel();
        if (nb > 1) throw EngineError{"ensemble handle: this line length needs the pipelined sweeps"};
        if (acc_extra) throw EngineError{"internal: two-field accumulator needs the pipelined sweep"};
        if (keep_x && smem_s2<N>() > 227 * 1024)
          throw EngineError{"512-point lines in double precision: the operatorA sweep does not fit shared memory"};
        nblk = (int)grid_s(g).x;
        L(tag, ks_deriv2<T, N, EPI>, grid_s(g), block_s<N>(), keep_x ? smem_s2<N>() : smem_s<N>(), st, g, (const C*)x,
          (const C*)kfield, (const C*)acc, (const C*)tw_for(nline, g), alpha, (C*)out1, (C*)out2, pp, done);
      }
    });
    return nblk;
  }
  // twiddle table of the axis a TileS sweeps along (y tiles step rows by n2c, x tiles by n1*n2c)
  const C* tw_for(int, const TileS& g) const { return g.row_stride == (long)n2c ? tw[1] : tw[0]; }

  // slab-decomposed form: the x sweep runs first, straight on the owners' memory, between two
  // rank barriers; the z sweep adds to it and the y sweep carries the epilogue.
  template <int EPI>
  int dapply_dist(const T* x, const T* kfield, T alpha, T* out1, T* out2, double* pp, const int* done) {
    const T* kpen = (kfield == kf) ? kT : ktilT;
    const TileX txd = tile_xd();
    const TileS ty = tile_y();
    const PeerRows<T> xr = rows(x, 1), ar = rows(acc, 2);
    const bool overlapped = use_xz && !prof.on;
    if (overlapped) {
      // acc2 = D_z(k D_z x) on the side stream, concurrently with the barriers and the peer x sweep
      fork.begin(st);
      z_side = true; z_out_override = acc2; pdl_hold = true;
      sweep_deriv2_z<0>("kz_deriv2.side", x, kfield, done);
      z_side = false; z_out_override = nullptr; pdl_hold = false;
    }
    // Rank ordering around the x sweep: the pipelined kernel carries its own gates (entry: announce + wait for every
    // peer; exit: the last CTA announces), and the y sweep below waits for the peers' exit announcements in its
    // prologue -- no stand-alone barrier launches in the hot loop.  The serial order (profiling pass, GLIA_RD_XZ=0)
    // hands acc to the z sweep, which has no gate: it keeps the trailing barrier kernel.
    GLIA_DISPATCH_N(n[0], {
      if constexpr (pipe_fits<T, N>()) {
        const int ntiles = txd.nchunk * txd.n_outer;
        const RowsX<T> rx{xr, txd}, ra{ar, txd};
        dim3 gx = grid_pipe<N>(ntiles);
        if (overlapped) {  // leave SMs to the z sweep
          const int cap = xz_ctas > 0 ? xz_ctas : (int)(0.75 * nsm * pipe_ctas<T, N>());
          if ((int)gx.x > cap) gx.x = (unsigned)(cap < 1 ? 1 : cap);
        }
        PeerGate gate = gate_xsweep();
        if (!overlapped) { gate.signal_out = 0; gate_pending = 0; }
        L("kx_deriv2_dist", ks_deriv2_pipe<T, N, EPI_SET, RowsX<T>, RowsPen<T>, RowsX<T>, RowsX<T>>, gx,
          block_s<N>(), pipe_smem<T, N>(), st, ntiles, rx, RowsPen<T>{(C*)const_cast<T*>(kpen), txd}, ra, ra, ra,
          (const C*)tw[0], (T)0, (double*)nullptr, done, gate, (int)gx.x);
        if (!overlapped) barrier();
      } else {
        barrier();
        L("kx_deriv2_dist", kx_deriv2_dist<T, N>, grid_xd(txd), block_s<N>(), smem_s<N>(), st, txd, xr, (const C*)kpen, ar,
          (const C*)tw[0], done);
        barrier();
      }
    });
    if (overlapped) fork.end(st);
    else sweep_deriv2_z<1>("kz_deriv2.add", x, kfield, done);
    const char* ytag = EPI == EPI_MATVEC ? "ks_deriv2.y.matvec" : (EPI == EPI_RHS ? "ks_deriv2.y.rhs" : "ks_deriv2.y.epi");
    return sweep_deriv2_local<EPI>(n[1], ytag, ty, x, kfield, alpha, out1, out2, pp, done, overlapped ? acc2 : nullptr);
  }

  template <int EPI>
  int dapply(const T* x, const T* kfield, T alpha, T* out1, T* out2, double* pp, const int* done) {
    if (G > 1) return dapply_dist<EPI>(x, kfield, alpha, out1, out2, pp, done);
    sweep_deriv2_z<0>("kz_deriv2", x, kfield, done);
    const TileS ty = tile_y(), tx = tile_x();
    sweep_deriv2_local<EPI_ADD>(n[1], "ks_deriv2.y", ty, x, kfield, (T)0, acc, nullptr, nullptr, done);
    const char* xtag = EPI == EPI_MATVEC ? "ks_deriv2.x.matvec" : (EPI == EPI_RHS ? "ks_deriv2.x.rhs" : "ks_deriv2.x");
    return sweep_deriv2_local<EPI>(n[0], xtag, tx, x, kfield, alpha, out1, out2, pp, done);
  }

  // z = M^-1 r with optional prologue r -= a w, optional store, partial {<z,z>,<r,z>}
  int pc_apply(T* rin, const T* wv, T* zout, bool want_rz, double* pp, const int* done) {
    const TileS ty = tile_y(), tx = tile_x();
    int nblk = 0;
    GLIA_DISPATCH_N(n[2], {
      if constexpr (r2cpipe_fits<T, N>()) {  // persistent warp-private pipelined form (sweeps_zpipe.cuh)
        const int ngroups = zgroups<N>(), cpm = cpm_zpipe2<N>();
        // (the r <- r - a w form at 512-point lines keeps the one-group-per-CTA kernel: it runs at the copy rate there,
        // 328 us against 350 us pipelined, profiles/r2za_r2cpipe_ab.txt; an ensemble handle needs the member-aware form)
        if (wv && (N < 512 || nb > 1))
          LP("kz_r2c.axpy", kz_r2c_pipe<T, N, 1>, dim3((unsigned)(cpm * nb)), dim3(zthreads<N>()), r2cpipe_smem<T, N>(), st,
             lines_z_member(), ngroups, rin, wv, (const double*)(scal + S_A), shat, (const C*)tw[2], done, cpm);
        else if (wv)
          LP("kz_r2c.axpy", kz_r2c<T, N, 1>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st, lines_z(), rin, wv,
             (const double*)(scal + S_A), shat, (const C*)tw[2], done, bpm_z<N>());
        else
          LP("kz_r2c", kz_r2c_pipe<T, N, 0>, dim3((unsigned)(cpm * nb)), dim3(zthreads<N>()), r2cpipe_smem<T, N>(), st,
             lines_z_member(), ngroups, rin, (const T*)nullptr, (const double*)nullptr, shat, (const C*)tw[2], done, cpm);
      } else if (wv) {
        LP("kz_r2c.axpy", kz_r2c<T, N, 1>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st, lines_z(), rin, wv,
           (const double*)(scal + S_A), shat, (const C*)tw[2], done, bpm_z<N>());
      } else {
        LP("kz_r2c", kz_r2c<T, N, 0>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st, lines_z(), rin,
           (const T*)nullptr, (const double*)nullptr, shat, (const C*)tw[2], done, bpm_z<N>());
      }
    });
    GLIA_DISPATCH_N(n[1], {
      if constexpr (pipe_fits<T, N>()) {
        const int ntiles = ty.nchunk * ty.n_outer;
#if !defined(GLIA_SIMT_EMU)
        if (use_tma && nb == 1 && tma_map()) {
          LP("ks_c2c.y", ks_c2c_tma<T, N, -1>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles, ty.nchunk,
             tmap_shat, (const C*)tw[1], done);
        } else
#endif
        LP("ks_c2c.y", ks_c2c_pipe<T, N, -1, RowsS<T>>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles,
           rows_s(ty, shat), rows_s(ty, shat), (const C*)tw[1], done, gate_none(), cpm_pipe<N>(ntiles));
      } else {
        LP("ks_c2c.y", ks_c2c<T, N, -1>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty, (const C*)shat, shat,
           (const C*)tw[1], done);
      }
    });
    if (G > 1) {
      const TileX txd = tile_xd();
      const PeerRows<T> sr = rows((const T*)shat, 1), sw = rows((const T*)shat, 2);
      GLIA_DISPATCH_N(n[0], {
        if constexpr (pipe_fits<T, N>()) {  // gated like the D-apply's x sweep; the inverse y sweep below is the consumer
          const int ntiles = txd.nchunk * txd.n_outer;
          L("kx_pc_dist", ks_pc_pipe<T, N, RowsX<T>>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles,
            RowsX<T>{sr, txd}, RowsX<T>{sw, txd}, (const C*)tw[0], (const PcSym<T>*)syms_d, n[1], done, gate_xsweep(),
            cpm_pipe<N>(ntiles));
        } else {
          barrier();
          L("kx_pc_dist", kx_pc_dist<T, N>, grid_xd(txd), block_s<N>(), smem_s<N>(), st, txd, sr, (const C*)tw[0], sym, n[1],
            done);
          barrier();
        }
      });
    } else {
      GLIA_DISPATCH_N(n[0], {
        if constexpr (pipe_fits<T, N>()) {
          const int ntiles = tx.nchunk * tx.n_outer;
          LP("ks_pc", ks_pc_pipe<T, N, RowsS<T>>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles,
            rows_s(tx, shat), rows_s(tx, shat), (const C*)tw[0], (const PcSym<T>*)syms_d, n[1], done, gate_none(),
            cpm_pipe<N>(ntiles));
        } else {
          if (nb > 1) throw EngineError{"ensemble handle: this line length needs the pipelined sweeps"};
          L("ks_pc", ks_pc<T, N>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx, shat, (const C*)tw[0], sym, n[1], done);
        }
      });
    }
    GLIA_DISPATCH_N(n[1], {
      if constexpr (pipe_fits<T, N>()) {
        const int ntiles = ty.nchunk * ty.n_outer;
#if !defined(GLIA_SIMT_EMU)
        if (use_tma && nb == 1 && tma_map()) {
          LP("ks_c2c.y", ks_c2c_tma<T, N, +1>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles, ty.nchunk,
             tmap_shat, (const C*)tw[1], done);
        } else
#endif
        LP("ks_c2c.y", ks_c2c_pipe<T, N, +1, RowsS<T>>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles,
           rows_s(ty, shat), rows_s(ty, shat), (const C*)tw[1], done, (G > 1 && gate_pending) ? gate_consumer() : gate_none(),
           cpm_pipe<N>(ntiles));
      } else {
        gate_wait_kernel();
        LP("ks_c2c.y", ks_c2c<T, N, +1>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty, (const C*)shat, shat,
           (const C*)tw[1], done);
      }
    });
    GLIA_DISPATCH_N(n[2], {
      nblk = bpm_z<N>();   // partial sums per member
      if constexpr (c2rpipe_fits<T, N>()) {  // persistent warp-private pipelined form (sweeps_zpipe.cuh)
        const int ngroups = zgroups<N>(), cpm = cpm_zpipe2<N>();
        nblk = cpm;
        if (zout && want_rz)
          LP("kz_c2r.rz", kz_c2r_pipe<T, N, 2>, dim3((unsigned)(cpm * nb)), dim3(zthreads<N>()), c2rpipe_smem<T, N>(), st,
             lines_z_member(), ngroups, (const C*)shat, zout, (const T*)rin, pp, (const C*)tw[2], done, cpm);
        else
          LP(zout ? "kz_c2r" : "kz_c2r.norm", kz_c2r_pipe<T, N, 1>, dim3((unsigned)(cpm * nb)), dim3(zthreads<N>()),
             c2rpipe_smem<T, N>(), st, lines_z_member(), ngroups, (const C*)shat, zout, (const T*)nullptr, pp,
             (const C*)tw[2], done, cpm);
      } else if (zout && want_rz) {
        LP("kz_c2r.rz", kz_c2r<T, N, 2>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>() + smem_z_rstage<N, T>(), st,
           lines_z(), (const C*)shat, zout, (const T*)rin, pp, (const C*)tw[2], done, bpm_z<N>());
      } else {
        LP(zout ? "kz_c2r" : "kz_c2r.norm", kz_c2r<T, N, 1>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st, lines_z(),
           (const C*)shat, zout, (const T*)nullptr, pp, (const C*)tw[2], done, bpm_z<N>());
      }
    });
    return nblk;
  }

  // ------------------------------------------------------------ L0 API ----
  void gradient(T* gx, T* gy, T* gz, const T* x, int mask) {
    need_one_member("glia_rd_gradient");
    const TileS ty = tile_y(), tx = tile_x();
    if ((mask & 4) && gz) {
      GLIA_DISPATCH_N(n[2], L("kz_deriv1", kz_deriv1<T, N, 0>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st,
                                         lines_z(), x, gz, (const C*)tw[2]));
    }
    if ((mask & 2) && gy) {
      GLIA_DISPATCH_N(n[1], L("ks_deriv1.y", ks_deriv1<T, N, 0>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty,
                                         (const C*)x, (C*)gy, (const C*)tw[1]));
    }
    if (G > 1) {  // collective: every rank takes part in the x exchange, whatever its mask
      const TileX txd = tile_xd();
      GLIA_CHECK(rt::copy(stage, x, sizeof(T) * nreal, st));
      barrier();
      GLIA_DISPATCH_N(n[0], L("kx_deriv1_dist", kx_deriv1_dist<T, N, 0>, grid_xd(txd), block_s<N>(), smem_s<N>(), st, txd,
                                         rows(stage), rows(acc), (const C*)tw[0]));
      barrier();
      if ((mask & 1) && gx) GLIA_CHECK(rt::copy(gx, acc, sizeof(T) * nreal, st));
    } else if ((mask & 1) && gx) {
      GLIA_DISPATCH_N(n[0], L("ks_deriv1.x", ks_deriv1<T, N, 0>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx,
                                         (const C*)x, (C*)gx, (const C*)tw[0]));
    }
    sync();
  }
  void divergence(T* div, const T* dx, const T* dy, const T* dz) {
    need_one_member("glia_rd_divergence");
    const TileS ty = tile_y(), tx = tile_x();
    T* out = G > 1 ? acc : div;
    GLIA_DISPATCH_N(n[2], L("kz_deriv1", kz_deriv1<T, N, 0>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st,
                                       lines_z(), dz, out, (const C*)tw[2]));
    GLIA_DISPATCH_N(n[1], L("ks_deriv1.y", ks_deriv1<T, N, 1>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty,
                                       (const C*)dy, (C*)out, (const C*)tw[1]));
    if (G > 1) {
      const TileX txd = tile_xd();
      GLIA_CHECK(rt::copy(stage, dx, sizeof(T) * nreal, st));
      barrier();
      GLIA_DISPATCH_N(n[0], L("kx_deriv1_dist", kx_deriv1_dist<T, N, 1>, grid_xd(txd), block_s<N>(), smem_s<N>(), st, txd,
                                         rows(stage), rows(acc), (const C*)tw[0]));
      barrier();
      GLIA_CHECK(rt::copy(div, acc, sizeof(T) * nreal, st));
    } else {
      GLIA_DISPATCH_N(n[0], L("ks_deriv1.x", ks_deriv1<T, N, 1>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx,
                                         (const C*)dx, (C*)div, (const C*)tw[0]));
    }
    sync();
  }

  // ------------------------------------------------------------ L1 API ----
  void set_diffusion(const T* k, const double ka[3], double kscale) {
    GLIA_CHECK(rt::copy(kf, k, sizeof(T) * nreal, st));
    for (int i = 0; i < 3; ++i) kavg[i] = (T)ka[i];
    kavg_per_member = false;
    k_scale = (T)kscale;
    if (G > 1) build_pencil(kf, kT);
    sync();
  }
  void field_sum4(const T* t, const T* m0, const T* m1, const T* m2, double out[4], long nelem = -1) {
    if (nelem < 0) nelem = nreal;
    const dim3 g = grid_pw(nelem);
    L("k_dot3", k_dot3<T>, g, dim3(256), 0, st, nelem, t, m0, m1, m2, part(0));
    L("k_sum4", k_sum4, dim3(1), dim3(256), 0, st, (const double*)part(0), (int)g.x, scal + 8, comm,
      G > 1 ? next_epoch() : 0u, rseq++);
    GLIA_CHECK(rt::d2h(h_out, scal + 8, sizeof(double) * 4, st));
    sync();
    for (int i = 0; i < 4; ++i) out[i] = h_out[i];
  }
  void set_diffusion_tissue(const T* wm, const T* gm, const T* csf, double kscale, double kgm, double kglm,
                            double filter_sum) {
    // DiffCoef::setValues (src/mat/DiffCoef.cpp:77-131)
    k_scale = (T)kscale;
    T dk_gm = (T)kscale * (T)kgm, dk_wm = (T)kscale, dk_glm = (T)kscale * (T)kglm;
    if (dk_gm <= 0) dk_gm = 0;
    if (dk_glm <= 0) dk_glm = 0;
    const dim3 g = grid_pw(nreal);
    // kxx = 0; kxx += dk_gm*gm; kxx += dk_wm*wm; kxx += dk_glm*csf
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, kf, dk_gm, gm, (T)0, (const T*)nullptr);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, kf, dk_wm, wm, (T)1, (const T*)kf);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, kf, dk_glm, csf, (T)1, (const T*)kf);
    double s[4];
    field_sum4(kf, nullptr, nullptr, nullptr, s);
    const T ksum = (T)s[3];
    const T favg = (T)filter_sum;
    const T kav = ksum * ((T)1.0 / favg);
    kavg[0] = kavg[1] = kavg[2] = kav;
    kavg_per_member = false;
    if (G > 1) { build_pencil(kf, kT); sync(); }
  }
  void set_reaction_tissue(const T* wm, const T* gm, const T* csf, double rs, double rgm, double rglm) {
    T dr_gm = (T)rs * (T)rgm, dr_wm = (T)rs, dr_glm = (T)rs * (T)rglm;
    if (dr_gm <= 0) dr_gm = 0;
    if (dr_glm <= 0) dr_glm = 0;
    const dim3 g = grid_pw(nreal);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, rho, dr_gm, gm, (T)0, (const T*)nullptr);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, rho, dr_wm, wm, (T)1, (const T*)rho);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, rho, dr_glm, csf, (T)1, (const T*)rho);
    sync();
  }
  // Ensemble handles (BASELINE config 5): one set of tissue maps (ONE member's fields), a (kappa, rho) pair per
  // member.  DiffCoef::setValues / ReacCoef::setValues per member (src/mat/DiffCoef.cpp:77-131, src/mat/ReacCoef.cpp:
  // 13-38), k-bar per member for the preconditioner symbol.
  void set_coefficients_batch(const T* wm, const T* gm, const T* csf, const double* ks, double kgm, double kglm,
                              double filter_sum, const double* rs, double rgm, double rglm) {
    const dim3 g = grid_pw(nreal_m);
    for (int m = 0; m < nb; ++m) {
      T* km = kf + (size_t)m * nreal_m;
      T* rm = rho + (size_t)m * nreal_m;
      T dk_gm = (T)ks[m] * (T)kgm, dk_wm = (T)ks[m], dk_glm = (T)ks[m] * (T)kglm;
      if (dk_gm <= 0) dk_gm = 0;
      if (dk_glm <= 0) dk_glm = 0;
      L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal_m, km, dk_gm, gm, (T)0, (const T*)nullptr);
      L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal_m, km, dk_wm, wm, (T)1, (const T*)km);
      L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal_m, km, dk_glm, csf, (T)1, (const T*)km);
      double s4[4];
      field_sum4(km, nullptr, nullptr, nullptr, s4, nreal_m);
      const T kav = (T)s4[3] * ((T)1.0 / (T)filter_sum);
      kavg_m[3 * m] = kavg_m[3 * m + 1] = kavg_m[3 * m + 2] = kav;
      T dr_gm = (T)rs[m] * (T)rgm, dr_wm = (T)rs[m], dr_glm = (T)rs[m] * (T)rglm;
      if (dr_gm <= 0) dr_gm = 0;
      if (dr_glm <= 0) dr_glm = 0;
      L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal_m, rm, dr_gm, gm, (T)0, (const T*)nullptr);
      L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal_m, rm, dr_wm, wm, (T)1, (const T*)rm);
      L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal_m, rm, dr_glm, csf, (T)1, (const T*)rm);
    }
    k_scale = (T)ks[0];
    kavg[0] = kavg[1] = kavg[2] = kavg_m[0];
    kavg_per_member = true;
    sync();
  }
  // mass-effect style refresh of rho(x) and k(x) between time steps.  Like the reference it leaves the
  // averages k-bar (and with them the next prec_factor() symbol) and k_scale untouched:
  // updateReacAndDiffCoefficients writes kxx_ only, kxx_avg_ keeps the value of the last setValues.
  void update_reac_diff(const T* bg, const T* gm, const T* vt, const T* csf, double rho_s, double k_s, double gm_r,
                        double gm_k) {
    L("k_update_reac_diff", k_update_reac_diff<T>, grid_pw(nreal), dim3(256), 0, st, nreal, rho, kf, bg, gm, vt, csf,
      (T)rho_s, (T)k_s, (T)gm_r, (T)gm_k);
    if (G > 1) build_pencil(kf, kT);
    sync();
  }
  void apply_D(T* dc, const T* c, bool secondary) {
    const T* src = c;   // (an ensemble handle applies every member's own k to its member of c)
    if (dc == c || (G > 1 && !in_arena(c))) {  // the reference allows aliasing (PdeOperators.cpp:210)
      GLIA_CHECK(rt::copy(work11, c, sizeof(T) * nreal, st));
      src = work11;
    }
    dapply<EPI_PLAIN>(src, secondary ? ktil : kf, (T)0, dc, nullptr, nullptr, nullptr);
    sync();
  }

  // ------------------------------------------------------------ L2 API ----
  void prec_factor() {
    sym.dt = dt_ctx;
    sym.kxx = kavg[0];
    sym.kyy = kavg[1];
    sym.kzz = kavg[2];
    upload_syms();
  }

  void fetch_iscal() {
    GLIA_CHECK(rt::d2h(h_iscal, iscal, sizeof(int) * I_NISCAL * nb, st));
    sync();  // (throws if a peer wait timed out)
  }

  // one PCG iteration, enqueued without synchronisation; `it` is 1-based
  void enqueue_iteration(T* x, const T* xin, T dt_solve, int it) {
    const int* done = iscal + I_DONE;
    const T alph = (T)(-1.0 / 2.0 * (double)dt_solve);
    const int nb1 = dapply<EPI_MATVEC>(p, kf, alph, w, nullptr, part(0), done);
    LP("k_pcg_alpha", k_pcg_alpha<T>, dim3(nb), dim3(256), 0, st, (const double*)part(0), nb1, scal, iscal, comm,
      G > 1 ? next_epoch() : 0u, rseq++);
    const int nb2 = pc_apply(r, w, z, true, part(1), done);
    LP("k_pcg_beta", k_pcg_beta<T>, dim3(nb), dim3(256), 0, st, (const double*)part(1), nb2, scal, iscal, maxit, dtol,
      comm, G > 1 ? next_epoch() : 0u, rseq++);
    LP("k_cg_update", k_cg_update<T>, dim3(grid_pw(nreal_m).x, nb), dim3(256), 0, st, nreal_m,
       (const T*)(it == 1 ? xin : x), x, p, (const T*)z, (const double*)scal, (const int*)iscal, it);
  }

  // DiffusionSolver::solve.  Asynchronous up to the convergence read-back.  With `xin` the solve starts from that
  // field and leaves the result in x (out of place: xin is only read); without, in place like the reference.
  int diffusion_solve(T* x, double dt_in, const T* xin = nullptr) {
    if (xin == x) xin = nullptr;
    if (G > 1 && (!(in_arena(x) || in_hist(x)) || (xin && !(in_arena(xin) || in_hist(xin))))) {
      // the rhs x sweep reads its operand through the peer mappings
      GLIA_CHECK(rt::copy(stage, xin ? xin : x, sizeof(T) * nreal, st));
      const int its = diffusion_solve(stage, dt_in);
      GLIA_CHECK(rt::copy(x, stage, sizeof(T) * nreal, st));
      return its;
    }
    const T dts = (T)dt_in;
    dt_ctx = dts;  // side effect on later prec_factor() calls (trap T2)
    const T* src = xin ? xin : x;
    if (k_scale == (T)0) {
      if (xin) GLIA_CHECK(rt::copy(x, xin, sizeof(T) * nreal, st));
      return 0;
    }
    const T alph = (T)(1.0 / 2.0 * (double)dts);
    dapply<EPI_RHS>(src, kf, alph, b, r, nullptr, nullptr);
    const int nb0 = pc_apply(b, nullptr, nullptr, false, part(2), nullptr);
    const int nb1 = pc_apply(r, nullptr, p, true, part(1), nullptr);
    L("k_pcg_init", k_pcg_init, dim3(nb), dim3(256), 0, st, (const double*)part(2), nb0, (const double*)part(1), nb1,
                 scal, iscal, rtol, abstol, comm, G > 1 ? next_epoch() : 0u, rseq++);
    int it = 0;
    // speculate: enqueue as many iterations as the previous solve needed, then look.  Ensemble handles iterate until
    // EVERY member has converged; the kernels of a converged member's CTAs return at once.
    int burst = its_guess < 1 ? 1 : its_guess;
    for (;;) {
      for (int j = 0; j < burst; ++j) enqueue_iteration(x, src, dts, ++it);
      fetch_iscal();
      bool all_done = true;
      for (int m = 0; m < nb; ++m) all_done = all_done && h_iscal[m * I_NISCAL + I_DONE];
      if (all_done) break;
      burst = 1;
      if (it > maxit + 1) break;
    }
    int its_max = 0, its_sum = 0;
    for (int m = 0; m < nb; ++m) {
      const int* hi = h_iscal + m * I_NISCAL;
      if (hi[I_REASON] < 0 && hi[I_REASON] != KSP_DIVERGED_ITS)
        throw EngineError{"KSP diverged, reason " + std::to_string(hi[I_REASON]) + " (member " + std::to_string(m) + ")"};
      its_m[m] = hi[I_ITS];
      its_acc[m] += hi[I_ITS];
      its_sum += hi[I_ITS];
      if (hi[I_ITS] > its_max) its_max = hi[I_ITS];
      // a member that converged at the iteration-0 test never ran the update that moves xin into x
      if (hi[I_ITS] == 0 && xin)
        GLIA_CHECK(rt::copy(x + (size_t)m * nreal_m, xin + (size_t)m * nreal_m, sizeof(T) * nreal_m, st));
    }
    its_guess = its_max;
    return its_sum;   // one member: ksp_itr_; an ensemble: the sum over its members (per member: batch_iterations())
  }

  // ------------------------------------------------------------ L2a API ----
  void resize_history(int nt_, double dt_) {
    sync();
    for (int q = 0; q < G; ++q)
      if (q != rank && peer_hist[q]) { rt::ipc_close(peer_hist[q], hist_bytes); peer_hist[q] = nullptr; }
    hist_connected = false;
    if (G > 1) rt::ipc_free(hist_arena, hist_bytes, hist_handle);
    else rt::dev_free(hist_arena);
    hist_arena = nullptr;
    c_hist = p_hist = chalf_hist = nullptr;
    nt = nt_;
    dt = (T)dt_;
    const size_t fb = sizeof(T) * nreal;
    hist_bytes = fb * ((size_t)(nt + 1) * 2 + (size_t)(nt > 0 ? nt : 1));
    if (G > 1) GLIA_CHECK(rt::ipc_alloc((void**)&hist_arena, hist_bytes, hist_handle));
    else GLIA_CHECK(rt::dev_malloc((void**)&hist_arena, hist_bytes));
    c_hist = reinterpret_cast<T*>(hist_arena);
    p_hist = reinterpret_cast<T*>(hist_arena + fb * (size_t)(nt + 1));
    chalf_hist = reinterpret_cast<T*>(hist_arena + fb * (size_t)(nt + 1) * 2);
    GLIA_CHECK(rt::zero(hist_arena, hist_bytes, st));
    sync();
  }
  T* hist(int which, int i) {
    if (which == 0 && i >= 0 && i <= nt) return c_hist + (long)i * nreal;
    if (which == 1 && i >= 0 && i <= nt) return p_hist + (long)i * nreal;
    if (which == 2 && i >= 0 && i < nt) return chalf_hist + (long)i * nreal;
    throw EngineError{"history index out of range"};
  }
  // ct <- R(cin) (cin == nullptr: in place), nonlinear or linearised about clin
  void reaction(T* ct, const T* clin, T dtr, T* chalf_out, const T* cin = nullptr) {
    const T* in = cin ? cin : ct;
    if (clin)
      L("k_reaction_lin", k_reaction_lin<T>, grid_pw(nreal), dim3(256), 0, st, nreal, in, ct, (const T*)rho, clin, dtr);
    else
      L("k_reaction", k_reaction<T>, grid_pw(nreal), dim3(256), 0, st, nreal, in, ct, (const T*)rho, dtr, chalf_out);
  }
  // PdeOperatorsRD::solveIncremental
  void solve_incremental(T* ctil, int i, int mode, T dth) {
    L("k_incr_avg", k_incr_avg<T>, grid_pw(nreal), dim3(256), 0, st, nreal, work11, (const T*)hist(0, i),
                 (const T*)hist(0, i + 1), mode == 1 ? 1 : 0);
    // c_tilde += dt/2 * D~ temp   (caller passes dt/2, the update uses dt/2 of that)
    dapply<EPI_AXPY>(work11, ktil, (T)(dth / 2), ctil, nullptr, nullptr, nullptr);
  }
  int solve_state(const T* c0, T* cT, int linearized) {
    if (nt <= 0 || !c_hist) throw EngineError{"resize_history() first"};
    if (linearized == 2) need_one_member("solve_state(2)");
    std::fill(its_acc.begin(), its_acc.end(), 0);
    int total = 0;
    const double dth = (double)dt / 2.0;
    if (linearized == 0) {
      // The state walks through the history slots themselves: c_[i] --solve--> c_half_[i] --reaction--> c_[i+1]
      // --solve in place--> c_[i+1].  The stores of PdeOperators.cpp:252, 277, 300 cost nothing.
      GLIA_CHECK(rt::copy(hist(0, 0), c0, sizeof(T) * nreal, st));
      for (int i = 0; i < nt; ++i) {
        total += diffusion_solve(hist(2, i), order == 2 ? dth : (double)dt, hist(0, i));
        reaction(hist(0, i + 1), nullptr, dt, nullptr, hist(2, i));
        if (order == 2) total += diffusion_solve(hist(0, i + 1), dth);
      }
      GLIA_CHECK(rt::copy(c_t, hist(0, nt), sizeof(T) * nreal, st));
    } else {
      GLIA_CHECK(rt::copy(c_t, c0, sizeof(T) * nreal, st));
      for (int i = 0; i < nt; ++i) {
        if (linearized == 2) solve_incremental(c_t, i, 1, (T)dth);
        total += diffusion_solve(c_t, order == 2 ? dth : (double)dt);
        reaction(c_t, hist(0, i), dt, nullptr);
        if (order == 2) {
          total += diffusion_solve(c_t, dth);
          if (linearized == 2) solve_incremental(c_t, i, 2, (T)dth);
        }
      }
    }
    if (cT && cT != c_t) GLIA_CHECK(rt::copy(cT, c_t, sizeof(T) * nreal, st));
    sync();
    return total;
  }
  int solve_adjoint(const T* pT, T* p0out, int linearized, int adjoint_store) {
    if (nt <= 0 || !c_hist) throw EngineError{"resize_history() first"};
    std::fill(its_acc.begin(), its_acc.end(), 0);
    GLIA_CHECK(rt::copy(p_0, pT, sizeof(T) * nreal, st));
    if (linearized == 1) GLIA_CHECK(rt::copy(hist(1, nt), p_0, sizeof(T) * nreal, st));
    int total = 0;
    const double dth = (double)dt / 2.0;
    for (int i = 0; i < nt; ++i) {
      const int it = nt - i - 1;
      // the adjoint walks through p_[.]: p_[it+1] --solve--> p_0 (work) --reaction--> p_[it] --solve in place--> p_[it]
      // (p_[nt] itself is only a valid source when this call wrote it: trap T4)
      total += diffusion_solve(p_0, order == 2 ? dth : (double)dt, i == 0 ? nullptr : (const T*)hist(1, it + 1));
      const T* clin = hist(2, it);
      if (!adjoint_store) {  // reactionAdjoint re-diffuses c_[it] by dt/2 (PdeOperators.cpp:330-340), whatever the order
        total += diffusion_solve(work11, dth, hist(0, it));
        clin = work11;
      }
      reaction(hist(1, it), clin, dt, nullptr, p_0);
      if (order == 2) total += diffusion_solve(hist(1, it), dth);
    }
    GLIA_CHECK(rt::copy(p_0, hist(1, 0), sizeof(T) * nreal, st));
    if (p0out && p0out != p_0) GLIA_CHECK(rt::copy(p0out, p_0, sizeof(T) * nreal, st));
    sync();
    return total;
  }

  // ------------------------------------------------------------ L2b API ----
  void grad_kappa_rho(const T* wm, const T* gm, const T* csf, double out[6]) {
    need_one_member("glia_rd_grad_kappa_rho");
    if (nt <= 0 || !c_hist) throw EngineError{"resize_history() first"};
    GLIA_CHECK(rt::zero(Tk, sizeof(T) * nreal, st));
    GLIA_CHECK(rt::zero(Tr, sizeof(T) * nreal, st));
    if (G > 1) { GLIA_CHECK(rt::zero(TkX, sizeof(T) * nreal, st)); barrier(); }  // peers' histories complete
    const TileS ty = tile_y(), tx = tile_x();
    for (int i = 0; i <= nt; ++i) {
      const T wgt = (i == 0 || i == nt) ? (T)0.5 : (T)1.0;
      const T coef = dt * wgt;
      const T* ci = hist(0, i);
      const T* pi = hist(1, i);
      GLIA_DISPATCH_N(n[2], L("kz_gradprod", kz_gradprod<T, N>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st,
                                         lines_z(), ci, pi, Tk, Tr, coef, (const C*)tw[2]));
      GLIA_DISPATCH_N(n[1], L("ks_gradprod.y", ks_gradprod<T, N>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty,
                                         (const C*)ci, (const C*)pi, (C*)Tk, coef, (const C*)tw[1]));
      if (G > 1) {  // x part: rows fetched from the owners' histories, accumulated in the pencil layout
        const TileX txd = tile_xd();
        GLIA_DISPATCH_N(n[0], L("kx_gradprod_dist", kx_gradprod_dist<T, N>, grid_xd(txd), block_s<N>(), smem_s<N>(), st,
                                           txd, rows(ci), rows(pi), (C*)TkX, coef, (const C*)tw[0]));
      } else {
        GLIA_DISPATCH_N(n[0], L("ks_gradprod.x", ks_gradprod<T, N>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx,
                                           (const C*)ci, (const C*)pi, (C*)Tk, coef, (const C*)tw[0]));
      }
    }
    if (G > 1) {
      barrier();
      L("k_pencil_add_to_slab", k_pencil_add_to_slab<T>, grid_pw(ncplx), dim3(256), 0, st, tile_xd(), n[0], (const C*)TkX,
        rows(Tk));
      barrier();
    }
    const double leb = (2.0 * M_PI / n[0]) * (2.0 * M_PI / n[1]) * (2.0 * M_PI / n[2]);
    double s[4];
    field_sum4(Tk, wm, gm, csf, s);
    out[0] = leb * s[0]; out[1] = leb * s[1]; out[2] = leb * s[2];
    field_sum4(Tr, wm, gm, csf, s);
    out[3] = leb * s[0]; out[4] = leb * s[1]; out[5] = leb * s[2];
  }

  // ------------------------------------ objective / gradient / Hessian ----
  // DiffCoef::setSecondaryCoefficients (src/mat/DiffCoef.cpp:44-59): k~ = k1 wm + k2 gm + k3 csf
  void set_secondary_tissue(const T* wm, const T* gm, const T* csf, double k1, double k2, double k3) {
    const dim3 g = grid_pw(nreal);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, ktil, (T)k1, wm, (T)0, (const T*)nullptr);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, ktil, (T)k2, gm, (T)1, (const T*)ktil);
    L("k_axpby", k_axpby<T>, g, dim3(256), 0, st, nreal, ktil, (T)k3, csf, (T)1, (const T*)ktil);
    if (G > 1) build_pencil(ktil, ktilT);
    sync();
  }
  // t = O c - d1, pT = -O^T t into Tr; returns { <t,t>, <c0,c0> } summed over all ranks
  void terminal_condition(const T* c, const T* d1, const T* obs, const T* c0, double sums[2], T* pT_out) {
    const dim3 g = grid_pw(nreal);
    L("k_obs_mismatch", k_obs_mismatch<T>, g, dim3(256), 0, st, nreal, c, d1, obs, c0, pT_out, part(0));
    L("k_sum4", k_sum4, dim3(1), dim3(256), 0, st, (const double*)part(0), (int)g.x, scal + 8, comm,
      G > 1 ? next_epoch() : 0u, rseq++);
    GLIA_CHECK(rt::d2h(h_out, scal + 8, sizeof(double) * 4, st));
    sync();
    sums[0] = h_out[0];
    sums[1] = h_out[1];
  }
  double lebesgue() const { return (2.0 * M_PI / n[0]) * (2.0 * M_PI / n[1]) * (2.0 * M_PI / n[2]); }
  // DerivativeOperatorsRD::evaluateObjectiveAndGradient in field space
  // (src/grad/DerivativeOperatorsRD.cpp:130-226): J[4] = {J, mismatch term, regularisation, t = 0 mismatch},
  // g_c0 = -h^3 (alpha(0) - beta c0)  (g_p = Phi^T g_c0), g[6] as grad_kappa_rho.
  void objective_gradient(const T* c0, const T* d1, const T* obs, double beta, const T* wm, const T* gm, const T* csf,
                          double J[4], T* g_c0, double g[6], int ksp[2]) {
    const double leb = lebesgue();
    double sums[2], m0 = 0.0;
    if (two_snap) {  // ||O0 c(0) - d0||^2 ; Tk <- -O0^T(O0 c0 - d0), kept for the gradient term below (:149-153)
      terminal_condition(c0, d0, has_obs0 ? obs0 : nullptr, nullptr, sums, Tk);
      m0 = sums[0];
    }
    ksp[0] = solve_state(c0, nullptr, 0);
    terminal_condition(c_t, d1, obs, c0, sums, Tr);
    J[1] = leb * 0.5 * sums[0];
    J[2] = 0.5 * beta * sums[1] * leb;
    J[3] = leb * 0.5 * m0;
    J[0] = leb * 0.5 * (sums[0] + m0) + J[2];
    ksp[1] = solve_adjoint(Tr, nullptr, 1, 1);
    if (g_c0) {  // p0 - beta c0, then scaled by -h^3 (VecAXPY, VecScale)
      L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, g_c0, (T)1, (const T*)p_0, (T)(-beta), c0);
      L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, g_c0, (T)(-leb), (const T*)g_c0, (T)0,
        (const T*)nullptr);
      // + h^3 O0^T(O0 c0 - d0)  (DerivativeOperatorsRD.cpp:216-222; Tk holds its negative)
      if (two_snap)
        L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, g_c0, (T)1, (const T*)g_c0, (T)(-leb),
          (const T*)Tk);
    }
    grad_kappa_rho(wm, gm, csf, g);
  }
  // DerivativeOperatorsRD::evaluateHessian in field space (src/grad/DerivativeOperatorsRD.cpp:229-438).
  // y_c0 = h^3 (beta c0~ - alpha~(0)) [- h^3 alpha~_k(0)]; hk = h^3 <wm|gm|csf, T_kp>, <wm|gm|csf, T_kk>.
  void hessian_matvec(const T* c0t, const T* obs, double beta, int diffusivity_inversion, const T* wm, const T* gm,
                      const T* csf, T* y_c0, double hk[6], int ksp[4]) {
    if (two_snap) throw EngineError{"Hessian currently not implemented for two-snapshot scenario"};  // DerivativeOperatorsRD.cpp:234
    const double leb = lebesgue();
    double sums[2], gtmp[6];
    for (int i = 0; i < 6; ++i) hk[i] = 0;
    for (int i = 0; i < 4; ++i) ksp[i] = 0;
    ksp[0] = solve_state(c0t, nullptr, 1);
    terminal_condition(c_t, nullptr, obs, nullptr, sums, Tr);
    ksp[1] = solve_adjoint(Tr, nullptr, 2, 1);
    // y = beta c0~ - p0 ; y *= h^3
    L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, y_c0, (T)beta, c0t, (T)-1, (const T*)p_0);
    L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, y_c0, (T)leb, (const T*)y_c0, (T)0, (const T*)nullptr);
    if (!diffusivity_inversion) { sync(); return; }
    grad_kappa_rho(wm, gm, csf, gtmp);
    for (int i = 0; i < 3; ++i) hk[i] = gtmp[i];
    GLIA_CHECK(rt::zero(Tr, sizeof(T) * nreal, st));
    ksp[2] = solve_state(Tr, nullptr, 2);
    terminal_condition(c_t, nullptr, obs, nullptr, sums, Tr);
    ksp[3] = solve_adjoint(Tr, nullptr, 2, 1);
    L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, y_c0, (T)(-leb), (const T*)p_0, (T)1, (const T*)y_c0);
    grad_kappa_rho(wm, gm, csf, gtmp);
    for (int i = 0; i < 3; ++i) hk[3 + i] = gtmp[i];
    sync();
  }

  // measurement probe (slab handles): `reps` back-to-back x sweeps of the preconditioner (what = 0)
  // or of the D-apply (what = 1) on whatever the work buffers hold, with the peer reads and / or
  // writes redirected to local memory by `local_mask` (1 = reads, 2 = writes).  Timing only.
  double probe_xsweep(int what, int local_mask, int reps) {
    if (G <= 1) throw EngineError{"probe_xsweep: slab handles only"};
    const int saved = dist_debug;
    dist_debug = local_mask;
    const TileX txd = tile_xd();
    const int ntiles = txd.nchunk * txd.n_outer;
    GLIA_CHECK(rt::zero(shat, sizeof(T) * nreal, st));
    GLIA_CHECK(rt::zero(p, sizeof(T) * nreal, st));
    barrier();
    sync();
    timer.start(st);
    for (int i = 0; i < reps; ++i) {
      if (what == 0) {
        const PeerRows<T> sr = rows((const T*)shat, 1), sw = rows((const T*)shat, 2);
        GLIA_DISPATCH_N(n[0], {
          if constexpr (pipe_fits<T, N>())
            L("probe", ks_pc_pipe<T, N, RowsX<T>>, grid_pipe<N>(ntiles), block_s<N>(), pipe_smem<T, N>(), st, ntiles,
              RowsX<T>{sr, txd}, RowsX<T>{sw, txd}, (const C*)tw[0], (const PcSym<T>*)syms_d, n[1], (const int*)nullptr,
              gate_none(), cpm_pipe<N>(ntiles));
        });
      } else {
        const PeerRows<T> xr = rows(p, 1), ar = rows(acc, 2);
        const RowsX<T> rx{xr, txd}, ra{ar, txd};
        GLIA_DISPATCH_N(n[0], {
          if constexpr (pipe_fits<T, N>())
            L("probe", ks_deriv2_pipe<T, N, EPI_SET, RowsX<T>, RowsPen<T>, RowsX<T>, RowsX<T>>, grid_pipe<N>(ntiles),
              block_s<N>(), pipe_smem<T, N>(), st, ntiles, rx, RowsPen<T>{(C*)kT, txd}, ra, ra, ra, (const C*)tw[0], (T)0,
              (double*)nullptr, (const int*)nullptr, gate_none(), cpm_pipe<N>(ntiles));
        });
      }
    }
    const double ms = timer.stop_ms(st);
    barrier();
    sync();
    dist_debug = saved;
    return ms / reps;
  }

  // ------------------------------------------- smoother / Phi / MatProp ----
  // symbol of the periodised, normalised Gaussian along every axis: s_d[k] = DFT(g_d)[k] / (sum(g_d) n_d),
  // g_d(X) = exp(-X^2/2s^2) + exp(-(X-2pi)^2/2s^2)  (SpectralOperators.cpp:318-352 factorised).  The
  // grid coordinates and sigma are rounded to ScalarType first, like the reference's loop does.
  void build_symbol(double sigma) {
    if (sigma == sym_sigma && symtab[0]) return;
    const T twopi = (T)(2.0 * M_PI);
    const T sg = (T)sigma;
    for (int a = 0; a < 3; ++a) {
      const int nn = n[a];
      std::vector<double> g(nn);
      std::vector<T> tab(nn);
      if (sigma == 0.0) {  // identity (the reference returns early, SpectralOperators.cpp:272-274)
        for (int k = 0; k < nn; ++k) tab[k] = (T)(1.0 / nn);
      } else {
        const T hh = twopi / (T)nn;
        double S = 0;
        for (int j = 0; j < nn; ++j) {
          const T X = hh * (T)j;
          const T Xp = X - twopi;
          const double s2 = 2.0 * (double)sg * (double)sg;
          g[j] = std::exp(-(double)X * (double)X / s2) + std::exp(-(double)Xp * (double)Xp / s2);
          S += g[j];
        }
        for (int k = 0; k < nn; ++k) {
          double re = 0;
          for (int j = 0; j < nn; ++j) re += g[j] * std::cos(2.0 * M_PI * (double)((long)j * k % nn) / nn);
          tab[k] = (T)(re / (S * nn));
        }
      }
      if (!symtab[a]) GLIA_CHECK(rt::dev_malloc((void**)&symtab[a], sizeof(T) * nn));
      GLIA_CHECK(rt::h2d(symtab[a], tab.data(), sizeof(T) * nn, st));
      GLIA_CHECK(rt::sync(st));
    }
    sym_sigma = sigma;
  }
  void need_single(const char* what) const {
    if (G > 1) throw EngineError{std::string(what) + ": single-GPU handles only (slab handles take c(0) as a field)"};
    if (nb > 1) throw EngineError{std::string(what) + ": not available on ensemble handles"};
  }
  void need_one_member(const char* what) const {
    if (nb > 1) throw EngineError{std::string(what) + ": not available on ensemble handles (nbatch > 1)"};
  }
  // SpectralOperators::weierstrassSmoother (src/grad/SpectralOperators.cpp:263-381); out may alias in
  void smooth(T* out, const T* in, double sigma) {
    need_single("glia_rd_smooth");
    if (sigma == 0.0) {
      if (out != in) GLIA_CHECK(rt::copy(out, in, sizeof(T) * nreal, st));
      sync();
      return;
    }
    build_symbol(sigma);
    const TileS ty = tile_y(), tx = tile_x();
    const GaussPhi<T> gp{};
    GLIA_DISPATCH_N(n[2], L("kz_filter", kz_filter<T, N, 0>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st, lines_z(),
                                       in, out, (const T*)symtab[2], (const C*)tw[2], gp, n[1]));
    GLIA_DISPATCH_N(n[1], L("ks_filter.y", ks_filter<T, N, 0>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty, (const C*)out,
                                       (C*)out, (const T*)symtab[1], (const C*)tw[1], gp, (C*)nullptr, (T)0,
                                       (double*)nullptr, (double*)nullptr));
    GLIA_DISPATCH_N(n[0], L("ks_filter.x", ks_filter<T, N, 0>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx, (const C*)out,
                                       (C*)out, (const T*)symtab[0], (const C*)tw[0], gp, (C*)nullptr, (T)0,
                                       (double*)nullptr, (double*)nullptr));
    sync();
  }
  // MatProp::setValuesCustom (src/mat/MatProp.cpp:135-201); returns sum(filter)
  double mat_prop(T* gm, T* wm, T* vt, T* csf, T* bg, T* filter) {
    const dim3 g = grid_pw(nreal);
    L("k_mat_prop", k_mat_prop<T>, g, dim3(256), 0, st, nreal, gm, wm, vt, csf, bg, filter, part(0));
    L("k_sum4", k_sum4, dim3(1), dim3(256), 0, st, (const double*)part(0), (int)g.x, scal + 8, comm,
      G > 1 ? next_epoch() : 0u, rseq++);
    GLIA_CHECK(rt::d2h(h_out, scal + 8, sizeof(double) * 4, st));
    sync();
    return h_out[3];
  }
  // Phi::setGaussians / setValues (src/mat/Phi.cpp:24-120): centres, sigma, the MatProp filter
  // (copied; null = no filter) and the smoothing width sigma_smooth = smoothing_factor 2pi/n0
  void phi_set(int np, const double* centers, double sigma_phi, const T* filter, double sigma_smooth) {
    need_single("glia_rd_phi_set");
    if (np < 0 || (np > 0 && !centers)) throw EngineError{"phi_set: bad arguments"};
    if (!(sigma_phi > 0)) throw EngineError{"phi_set: sigma must be positive"};
    phi_np = np;
    phi_centers.assign(centers, centers + 3 * (size_t)np);
    phi_sigma = (T)sigma_phi;
    phi_sigma_smooth = sigma_smooth;
    phi_has_filter = filter != nullptr;
    if (filter) {
      if (!phi_filter) GLIA_CHECK(rt::dev_malloc((void**)&phi_filter, sizeof(T) * nreal));
      GLIA_CHECK(rt::copy(phi_filter, filter, sizeof(T) * nreal, st));
    }
    if (np > phi_dots_cap) {
      rt::dev_free(phi_dots);
      phi_dots = nullptr;
      GLIA_CHECK(rt::dev_malloc((void**)&phi_dots, sizeof(double) * np));
      phi_dots_cap = np;
    }
    build_symbol(sigma_smooth);
    sync();
  }
  GaussPhi<T> gauss(int i) const {
    const T twopi = (T)(2.0 * M_PI);
    GaussPhi<T> gp;
    gp.xc = (T)phi_centers[3 * i]; gp.yc = (T)phi_centers[3 * i + 1]; gp.zc = (T)phi_centers[3 * i + 2];
    gp.hx = twopi / (T)n[0]; gp.hy = twopi / (T)n[1]; gp.hz = twopi / (T)n[2];
    gp.R = (T)(std::sqrt(2.) * (double)phi_sigma);
    gp.sigma = phi_sigma;
    return gp;
  }
  // basis function i into Tk up to (not including) the x sweep; returns the x-sweep grid size
  void phi_zy(const GaussPhi<T>& gp) {
    const TileS ty = tile_y();
    GLIA_DISPATCH_N(n[2], L("kz_filter.gauss", kz_filter<T, N, 1>, grid_z<N>(), dim3(zthreads<N>()), smem_z<N>(), st,
                                       lines_z(), (const T*)(phi_has_filter ? phi_filter : nullptr), Tk,
                                       (const T*)symtab[2], (const C*)tw[2], gp, n[1]));
    GLIA_DISPATCH_N(n[1], L("ks_filter.y", ks_filter<T, N, 0>, grid_s(ty), block_s<N>(), smem_s<N>(), st, ty, (const C*)Tk,
                                       (C*)Tk, (const T*)symtab[1], (const C*)tw[1], gp, (C*)nullptr, (T)0,
                                       (double*)nullptr, (double*)nullptr));
  }
  // Phi::apply, on-the-fly mode (src/mat/Phi.cpp:324-383): out = sum_i p_i phi_i / max_i max(phi_i)
  void phi_apply(T* out, const double* p) {
    need_single("glia_rd_phi_apply");
    if (phi_np <= 0) throw EngineError{"phi_apply: glia_rd_phi_set first"};
    build_symbol(phi_sigma_smooth);
    double* run_max = scal + 12;
    GLIA_CHECK(rt::zero(out, sizeof(T) * nreal, st));
    GLIA_CHECK(rt::zero(run_max, sizeof(double), st));
    const TileS tx = tile_x();
    const int nblk = (int)grid_s(tx).x;
    bool nnz = false;
    for (int i = 0; i < phi_np; ++i) {
      const T pi = (T)p[i];
      if (pi == (T)0) continue;  // adds nothing to sum_i phi_i p_i (Phi.cpp:343)
      nnz = true;
      const GaussPhi<T> gp = gauss(i);
      phi_zy(gp);
      GLIA_DISPATCH_N(n[0], L("ks_filter.x.phi", ks_filter<T, N, 1>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx,
                                         (const C*)Tk, (C*)Tk, (const T*)symtab[0], (const C*)tw[0], gp, (C*)out, pi, part(0),
                                         (double*)nullptr));
      L("k_phi_reduce", k_phi_reduce, dim3(1), dim3(256), 0, st, (const double*)part(0), (const double*)nullptr, nblk, run_max,
        (double*)nullptr);
    }
    if (!nnz) {  // phi_max = 1 (Phi.cpp:377)
      h_out[8] = 1.0;
      GLIA_CHECK(rt::h2d(run_max, h_out + 8, sizeof(double), st));
    }
    L("k_scale_inv", k_scale_inv<T>, grid_pw(nreal), dim3(256), 0, st, nreal, out, (const double*)run_max);
    sync();
  }
  // Phi::applyTranspose, on-the-fly mode (Phi.cpp:385-434): pout_i = <phi_i, in> / max_i max(phi_i)
  void phi_apply_transpose(double* pout, const T* in) {
    need_single("glia_rd_phi_apply_transpose");
    if (phi_np <= 0) throw EngineError{"phi_apply_transpose: glia_rd_phi_set first"};
    build_symbol(phi_sigma_smooth);
    double* run_max = scal + 12;
    GLIA_CHECK(rt::zero(run_max, sizeof(double), st));
    const TileS tx = tile_x();
    const int nblk = (int)grid_s(tx).x;
    for (int i = 0; i < phi_np; ++i) {
      const GaussPhi<T> gp = gauss(i);
      phi_zy(gp);
      GLIA_DISPATCH_N(n[0], L("ks_filter.x.phiT", ks_filter<T, N, 2>, grid_s(tx), block_s<N>(), smem_s<N>(), st, tx,
                                         (const C*)Tk, (C*)Tk, (const T*)symtab[0], (const C*)tw[0], gp,
                                         (C*)const_cast<T*>(in), (T)0, part(0), part(1)));
      L("k_phi_reduce", k_phi_reduce, dim3(1), dim3(256), 0, st, (const double*)part(0), (const double*)part(1), nblk, run_max,
        phi_dots + i);
    }
    std::vector<double> dots(phi_np + 1);
    GLIA_CHECK(rt::d2h(h_out + 8, run_max, sizeof(double), st));
    sync();
    const T pm = (T)h_out[8];
    // the dots come back through a pageable vector: small, once per call
    GLIA_CHECK(rt::d2h(dots.data(), phi_dots, sizeof(double) * phi_np, st));
    sync();
    const T alpha = (T)(1.0 / (double)pm);
    for (int i = 0; i < phi_np; ++i) pout[i] = (double)((T)dots[i] * alpha);
  }

  // ------------------------------------------------ data in / out, labels ----
  // dataIn (src/utils/IO.cpp:511-538): variable "data" of a NetCDF classic file -> this rank's block
  void data_in(const char* path, T* field) {
    std::vector<T> host((size_t)nreal);
    try { nc::read_block<T>(path, n, rank * n0l, n0l, host.data()); }
    catch (const nc::Error& e) { throw EngineError{e.msg}; }
    GLIA_CHECK(rt::h2d(field, host.data(), sizeof(T) * nreal, st));
    sync();
  }
  // dataOut (IO.cpp:540-612): CDF-2 file, dims x y z, NC_FLOAT / NC_DOUBLE; every rank writes its rows
  void data_out(const char* path, const T* field) {
    std::vector<T> host((size_t)nreal);
    GLIA_CHECK(rt::d2h(host.data(), field, sizeof(T) * nreal, st));
    sync();
    try { nc::write_block<T>(path, n, rank * n0l, n0l, host.data()); }
    catch (const nc::Error& e) { throw EngineError{e.msg}; }
  }
  void split_segmentation(const T* seg, const int labels[4], T* wm, T* gm, T* vt, T* csf) {
    L("k_split_seg", k_split_seg<T>, grid_pw(nreal), dim3(256), 0, st, nreal, seg, labels[0], labels[1], labels[2], labels[3],
      wm, gm, vt, csf);
    sync();
  }

  // ------------------------------------------ forward + adjoint entries ----
  // solveState(0), p_T = -(c(T) - d1) (O = I; DerivativeOperatorsRD.cpp:156-161), solveAdjoint(1)
  void forward_adjoint(const T* c0, const T* d1, T* cT, T* p0out, int* ks, int* ka) {
    *ks = solve_state(c0, cT, 0);
    L("k_axpby", k_axpby<T>, grid_pw(nreal), dim3(256), 0, st, nreal, Tr, (T)1, d1, (T)-1, (const T*)c_t);
    *ka = solve_adjoint(Tr, p0out, 1, 1);
  }
  // same with HOST buffers: H2D of c0 and d1, D2H of c(T) and p(0) inside the call.  Pinned
  // (page-locked) caller buffers are DMA'd directly; pageable ones go through pinned staging.
  void forward_adjoint_host(const T* c0h, const T* d1h, T* cTh, T* p0h, int* ks, int* ka) {
    const size_t bytes = sizeof(T) * nreal;
    const bool pin_in = rt::is_pinned(c0h) && rt::is_pinned(d1h);
    const bool pin_out = rt::is_pinned(cTh) && rt::is_pinned(p0h);
    if ((!pin_in || !pin_out) && !hs_in) {
      GLIA_CHECK(rt::host_malloc((void**)&hs_in, bytes * 2));
      GLIA_CHECK(rt::host_malloc((void**)&hs_out, bytes * 2));
    }
    const T *src0 = c0h, *src1 = d1h;
    if (!pin_in) {
      std::memcpy(hs_in, c0h, bytes);
      std::memcpy(hs_in + nreal, d1h, bytes);
      src0 = hs_in; src1 = hs_in + nreal;
    }
    GLIA_CHECK(rt::h2d(Tk, src0, bytes, st));
    // d1 must survive the forward solve: it goes to work11, which only solveIncremental and
    // the store-less adjoint use (neither runs here)
    GLIA_CHECK(rt::h2d(work11, src1, bytes, st));
    forward_adjoint(Tk, work11, Tk, p_0, ks, ka);
    T *dst0 = pin_out ? cTh : hs_out, *dst1 = pin_out ? p0h : hs_out + nreal;
    GLIA_CHECK(rt::d2h(dst0, Tk, bytes, st));
    GLIA_CHECK(rt::d2h(dst1, p_0, bytes, st));
    sync();
    if (!pin_out) {
      std::memcpy(cTh, hs_out, bytes);
      std::memcpy(p0h, hs_out + nreal, bytes);
    }
  }

  // ------------------------------------------------- type-erased face ----
  void* stream_handle() override { return (void*)(intptr_t)st; }
  void v_make_current() override { make_current(); }
  void v_fft_r2c(const void* f, void* fhat) override {
    if (G > 1) throw EngineError{"glia_rd_fft_r2c: the stand-alone 3-D FFT is single-GPU; slab handles transform inside the sweeps"};
    need_one_member("glia_rd_fft_r2c");
    fft3d_r2c(*this, (const T*)f, (C*)fhat);
  }
  void v_fft_c2r(const void* fhat, void* f) override {
    if (G > 1) throw EngineError{"glia_rd_fft_c2r: the stand-alone 3-D FFT is single-GPU; slab handles transform inside the sweeps"};
    need_one_member("glia_rd_fft_c2r");
    fft3d_c2r(*this, (const C*)fhat, (T*)f);
  }
  void v_ipc_export(int which, unsigned char* out) override { ipc_export(which, out); }
  void v_ipc_connect(int which, const unsigned char* handles) override { ipc_connect(which, handles); }
  void v_ipc_disconnect(int which) override { ipc_disconnect(which); }
  void v_wait_stream(void* producer) override { GLIA_CHECK(rt::stream_wait_stream(st, (cudaStream_t)(intptr_t)producer)); }
  void v_set_order(int o) override { order = o; }
  int v_nbatch() const override { return nb; }
  void v_batch_iterations(int* out, int accumulated) override {
    for (int m = 0; m < nb; ++m) out[m] = accumulated ? its_acc[m] : its_m[m];
  }
  void v_set_coefficients_batch(const void* wm, const void* gm, const void* csf, const double* ks, double kgm, double kglm,
                                double fsum, const double* rs, double rgm, double rglm) override {
    set_coefficients_batch((const T*)wm, (const T*)gm, (const T*)csf, ks, kgm, kglm, fsum, rs, rgm, rglm);
  }
  void v_set_two_snapshot(const void* d0_, const void* obs0_) override {
    two_snap = d0_ != nullptr;
    has_obs0 = two_snap && obs0_ != nullptr;
    if (!two_snap) return;
    if (!d0) GLIA_CHECK(rt::dev_malloc((void**)&d0, sizeof(T) * nreal));
    GLIA_CHECK(rt::copy(d0, d0_, sizeof(T) * nreal, st));
    if (has_obs0) {
      if (!obs0) GLIA_CHECK(rt::dev_malloc((void**)&obs0, sizeof(T) * nreal));
      GLIA_CHECK(rt::copy(obs0, obs0_, sizeof(T) * nreal, st));
    }
    sync();
  }
  void v_gradient(void* gx, void* gy, void* gz, const void* x, int m) override {
    gradient((T*)gx, (T*)gy, (T*)gz, (const T*)x, m);
  }
  void v_divergence(void* div, const void* dx, const void* dy, const void* dz) override {
    divergence((T*)div, (const T*)dx, (const T*)dy, (const T*)dz);
  }
  void v_set_diffusion(const void* k, const double ka[3], double ks) override { set_diffusion((const T*)k, ka, ks); }
  void v_set_diffusion_tissue(const void* wm, const void* gm, const void* csf, double ks, double kgm, double kglm,
                              double fsum) override {
    set_diffusion_tissue((const T*)wm, (const T*)gm, (const T*)csf, ks, kgm, kglm, fsum);
  }
  void v_set_secondary_k(const void* kt) override {
    GLIA_CHECK(rt::copy(ktil, kt, sizeof(T) * nreal, st));
    if (G > 1) build_pencil(ktil, ktilT);
    sync();
  }
  void v_set_reaction(const void* rh) override {
    GLIA_CHECK(rt::copy(rho, rh, sizeof(T) * nreal, st));
    sync();
  }
  void v_set_reaction_tissue(const void* wm, const void* gm, const void* csf, double rs, double rgm,
                             double rglm) override {
    set_reaction_tissue((const T*)wm, (const T*)gm, (const T*)csf, rs, rgm, rglm);
  }
  void v_update_reac_diff(const void* bg, const void* gm, const void* vt, const void* csf, double rho_s, double k_s,
                          double gm_r, double gm_k) override {
    update_reac_diff((const T*)bg, (const T*)gm, (const T*)vt, (const T*)csf, rho_s, k_s, gm_r, gm_k);
  }
  void v_apply_D(void* dc, const void* c, int secondary) override { apply_D((T*)dc, (const T*)c, secondary != 0); }
  void v_prec_factor() override { prec_factor(); }
  int v_diffusion_solve(void* c, double dts) override {
    const int k = diffusion_solve((T*)c, dts);
    sync();
    return k;
  }
  void v_set_ksp_tolerances(double rt_, double at_, double dt_, int mi) override {
    rtol = rt_; abstol = at_; dtol = dt_; maxit = mi;
  }
  void v_resize_history(int nt_, double dt_) override { resize_history(nt_, dt_); }
  void* v_history(int which, int i) override { return (void*)hist(which, i); }
  void v_reaction(void* ct, const void* clin, double dtr) override {
    reaction((T*)ct, (const T*)clin, (T)dtr, nullptr);
    sync();
  }
  int v_solve_state(const void* c0, void* cT, int lin) override { return solve_state((const T*)c0, (T*)cT, lin); }
  int v_solve_adjoint(const void* pT, void* p0o, int lin, int store) override {
    return solve_adjoint((const T*)pT, (T*)p0o, lin, store);
  }
  void v_grad_kappa_rho(const void* wm, const void* gm, const void* csf, double out[6]) override {
    grad_kappa_rho((const T*)wm, (const T*)gm, (const T*)csf, out);
  }
  void v_set_secondary_tissue(const void* wm, const void* gm, const void* csf, double k1, double k2, double k3) override {
    set_secondary_tissue((const T*)wm, (const T*)gm, (const T*)csf, k1, k2, k3);
  }
  void v_objective_gradient(const void* c0, const void* d1, const void* obs, double beta, const void* wm, const void* gm,
                            const void* csf, double J[4], void* g_c0, double g[6], int ksp[2]) override {
    objective_gradient((const T*)c0, (const T*)d1, (const T*)obs, beta, (const T*)wm, (const T*)gm, (const T*)csf, J,
                       (T*)g_c0, g, ksp);
  }
  void v_hessian_matvec(const void* c0t, const void* obs, double beta, int dinv, const void* wm, const void* gm,
                        const void* csf, void* y_c0, double hk[6], int ksp[4]) override {
    hessian_matvec((const T*)c0t, (const T*)obs, beta, dinv, (const T*)wm, (const T*)gm, (const T*)csf, (T*)y_c0, hk, ksp);
  }
  void v_smooth(void* out, const void* in, double sigma) override { smooth((T*)out, (const T*)in, sigma); }
  double v_mat_prop(void* gm, void* wm, void* vt, void* csf, void* bg, void* filter) override {
    return mat_prop((T*)gm, (T*)wm, (T*)vt, (T*)csf, (T*)bg, (T*)filter);
  }
  void v_phi_set(int np, const double* centers, double sigma_phi, const void* filter, double sigma_smooth) override {
    phi_set(np, centers, sigma_phi, (const T*)filter, sigma_smooth);
  }
  void v_phi_apply(void* out, const double* p) override { phi_apply((T*)out, p); }
  void v_phi_apply_transpose(double* pout, const void* in) override { phi_apply_transpose(pout, (const T*)in); }
  void v_data_in(const char* path, void* field) override { data_in(path, (T*)field); }
  void v_data_out(const char* path, const void* field) override { data_out(path, (const T*)field); }
  void v_split_segmentation(const void* seg, const int labels[4], void* wm, void* gm, void* vt, void* csf) override {
    split_segmentation((const T*)seg, labels, (T*)wm, (T*)gm, (T*)vt, (T*)csf);
  }
  double v_probe(int what, int mask, int reps) override { return probe_xsweep(what, mask, reps); }
  void v_profile_begin() override { sync(); prof.begin(); }
  std::string v_profile_end() override { return prof.end(st); }
  void v_timer_start() override { timer.start(st); }
  double v_timer_stop_ms() override { return timer.stop_ms(st); }
  void v_forward_adjoint(const void* c0, const void* d1, void* cT, void* p0o, int* ks, int* ka) override {
    forward_adjoint((const T*)c0, (const T*)d1, (T*)cT, (T*)p0o, ks, ka);
  }
  void v_forward_adjoint_host(const void* c0, const void* d1, void* cT, void* p0o, int* ks, int* ka) override {
    forward_adjoint_host((const T*)c0, (const T*)d1, (T*)cT, (T*)p0o, ks, ka);
  }
};

}  // namespace glia
