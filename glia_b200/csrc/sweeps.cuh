// sweeps.cuh -- axis-sweep kernels of the RD hot path.
//
// Every spectral operator of the reference (computeGradient / computeDivergence
// src/grad/SpectralOperators.cpp:100-261; applyD src/mat/DiffCoef.cpp:249-271;
// applyPC src/pde/DiffusionSolver.cpp:182-215) couples grid points only along
// grid lines, so each is evaluated here as 1-D line transforms along one axis per
// kernel ("sweep"), with the pointwise work of the neighbouring PETSc/cuBLAS
// calls fused into the sweep's prologue / epilogue.
//
// Two geometries:
//   Z  contiguous lines along z; two real lines (y, y+1) ride as re/im of one
//      complex line; TPL threads of one warp own a line, exchanges are
//      warp-synchronous.
//   S  strided lines along y or x; the real field [n0][n1][n2] is viewed as
//      complex [n0][n1][n2/2] (adjacent z pairs), a CTA owns a tile of N rows x
//      16 complex columns (128 B rows in single precision), lanes run along z so
//      every global access is a full 128-byte segment.
#pragma once
#include "fft_core.cuh"

namespace glia {

// ------------------------------------------------------------ helpers ----
// complex columns (lanes) per S tile: 16 = 128-byte rows in single precision.  (Measured in round 1: 8 columns --
// 64-byte rows, half-size CTAs -- made the x sweep 70 -> 90 us.)
static constexpr int SL = 16;

// occupancy hint of the S kernels: two resident CTAs per SM for single-precision tiles of <= 256 threads
template <typename T, int N>
__host__ __device__ constexpr int s_min_ctas() {
  return (SL * (N / FftPlan<N>::E) <= 256 && sizeof(T) == 4) ? 2 : 1;
}

struct TileS {        // S geometry: tile -> (outer, chunk)
  long row_stride;    // complex units between rows along the sweep axis
  long outer_stride;  // complex units between consecutive outer indices
  int nchunk;         // column chunks (of SL complex) per outer index
  int n_outer;
  long batch_stride;  // complex units between batch members (0 if unused)
};

struct LinesZ {
  long npairs;  // number of (y, y+1) line pairs = n0*n1/2 (times batch)
};

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&val)[NV], double* __restrict__ partial) {
  __shared__ double red[32 * NV];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  GLIA_UNROLL
  for (int i = 0; i < NV; ++i) {
    double v = val[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid * NV + i] = v;
  }
  __syncthreads();
  if (tid < NV) {
    double s = 0;
    for (int w = 0; w < nwarp; ++w) s += red[w * NV + tid];
    partial[(size_t)blockIdx.x * NV + tid] = s;
  }
  __syncthreads();
}

// CTA-wide barrier of the S geometry; half(h): the named barrier of thread half h (threads [h B/2, (h+1) B/2) of a
// B-thread CTA -- with threadIdx = t * SL + l these are the threads of line half h, see LineFft::SPLIT)
struct SyncCta {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
  __device__ __forceinline__ void half(int h) const {
#if defined(GLIA_SIMT_EMU)
    emu_half_barrier(h);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "r"((int)blockDim.x / 2) : "memory");
#endif
  }
};
struct SyncWarp {
  __device__ __forceinline__ void operator()() const { __syncwarp(); }
  __device__ __forceinline__ void half(int) const { __syncwarp(); }
};
// (Measured in round 2 and removed: running the single resident 512-thread CTA of a 512-point S sweep as two column
// halves of 256 threads with their own named barriers -- lanes along 8 columns, so that a warp touches 64-byte row
// pieces.  The halves do drift apart, but the half-line global requests cost far more: x.matvec 665 -> 1073 us,
// y 543 -> 700, ks_c2c.y 169 -> 233 at 512^3 on one box, profiles/r2n_split512_ab.txt.  Same lesson as round 1's
// 8-column tiles: keep a warp on whole 128-byte rows.)

template <int TPL> struct ZSync { using type = SyncWarp; };
template <> struct ZSync<64> { using type = SyncCta; };
template <> struct ZSync<128> { using type = SyncCta; };

struct AmS {  // S geometry smem map: loc*SL + lane
  int l;
  __device__ __forceinline__ int operator()(int loc) const { return loc * SL + l; }
};
struct AmZ {  // Z geometry smem map: padded line region
  int base;
  __device__ __forceinline__ int operator()(int loc) const { return base + loc + (loc >> 4); }
};
template <int N> __host__ __device__ constexpr int zpad() { return N + N / 16 + 8; }
template <int N> __host__ __device__ constexpr int zlines() { return 256 / (N / FftPlan<N>::E) > 0 ? 256 / (N / FftPlan<N>::E) : 1; }
template <int N> __host__ __device__ constexpr int zthreads() { return zlines<N>() * (N / FftPlan<N>::E); }

// v <- D_axis(v): forward, i*w/N, inverse
template <typename T, int N, int V = 0, class AM, class SY>
__device__ __forceinline__ void deriv_inplace(cplx<T> (&v)[FftPlan<N>::E], const typename LineFft<T, N, V>::Tw& tw,
                                              cplx<T>* sm, AM am, SY sy, int t) {
  using F = LineFft<T, N, V>;
  F::forward(v, tw, sm, am, sy, t);
  F::mult_iw(v, t);
  F::inverse(v, tw, sm, am, sy, t);
}

// ================================================================ S ====
enum { EPI_SET = 0, EPI_ADD = 1, EPI_PLAIN = 2, EPI_MATVEC = 3, EPI_RHS = 4, EPI_AXPY = 5 };

// S-geometry second-derivative sweep: s = acc + D(k . D x) along the tile axis.
//   EPI_SET    out1 = D(k D x)
//   EPI_ADD    out1 = s                                   (the caller passes out1 = acc)
//   EPI_PLAIN  out1 = s                                   (applyD result)
//   EPI_MATVEC out1 = x + alpha*s ; partial <x, out1>     (operatorA, alpha = -dt/2)
//   EPI_RHS    out1 = x + alpha*s ; out2 = out1 - (x - alpha*s)   (rhs and r0 = b - A x0)
//   EPI_AXPY   out1 += alpha*s                            (solveIncremental)
template <typename T, int N, int EPI>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E), s_min_ctas<T, N>())
ks_deriv2(TileS geo, const cplx<T>* __restrict__ x, const cplx<T>* __restrict__ kf, const cplx<T>* acc,
          const cplx<T>* __restrict__ twt, T alpha, cplx<T>* out1, cplx<T>* out2, double* partial,
          const int* __restrict__ done) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  constexpr bool KEEP_X = (EPI == EPI_MATVEC || EPI == EPI_RHS);
  if (done && *done) return;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  cplx<T>* smx = sm + N * SL;  // x tile kept for the epilogue (KEEP_X only)
  const int l = threadIdx.x & (SL - 1), t = threadIdx.x / SL;
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  const int tile = blockIdx.x;
  const int outer = tile / geo.nchunk, chunk = tile % geo.nchunk;
  const long base = (long)blockIdx.y * geo.batch_stride + (long)outer * geo.outer_stride + (long)chunk * SL + l;
  AmS am{l};
  SyncCta sy;

  cplx<T> v[E], kk[E];
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const long off = base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride;
    v[e] = x[off];
    kk[e] = kf[off];
  }
  if (KEEP_X) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) smx[am(F::template loc<0>(t, e / F::R(0), e % F::R(0)))] = v[e];
  }
  deriv_inplace<T, N>(v, tw, sm, am, sy, t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) { v[e].x *= kk[e].x; v[e].y *= kk[e].y; }
  // the accumulator tile is fetched now, so that its latency hides behind the second derivative
  cplx<T> ac[E];
  if (EPI != EPI_SET) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e)
      ac[e] = acc[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride];
  }
  deriv_inplace<T, N>(v, tw, sm, am, sy, t);

  if (EPI == EPI_AXPY) {  // out1 += alpha * (acc + D..): loads batched ahead of the stores
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const long off = base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride;
      const cplx<T> o = out1[off];
      v[e] = {o.x + alpha * (v[e].x + ac[e].x), o.y + alpha * (v[e].y + ac[e].y)};
    }
    GLIA_UNROLL
    for (int e = 0; e < E; ++e)
      out1[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride] = v[e];
    return;
  }
  double dsum[1] = {0.0};
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const int lc = F::template loc<0>(t, e / F::R(0), e % F::R(0));
    const long off = base + (long)lc * geo.row_stride;
    cplx<T> s = v[e];
    if (EPI != EPI_SET) { s.x += ac[e].x; s.y += ac[e].y; }
    if (EPI == EPI_SET || EPI == EPI_ADD || EPI == EPI_PLAIN) {
      out1[off] = s;
    } else if (EPI == EPI_MATVEC) {
      const cplx<T> xv = smx[am(lc)];
      cplx<T> w = {xv.x + alpha * s.x, xv.y + alpha * s.y};
      out1[off] = w;
      dsum[0] += (double)xv.x * (double)w.x + (double)xv.y * (double)w.y;
    } else if (EPI == EPI_RHS) {
      const cplx<T> xv = smx[am(lc)];
      const T ds0 = alpha * s.x, ds1 = alpha * s.y;
      cplx<T> b = {xv.x + ds0, xv.y + ds1};
      cplx<T> ax = {xv.x - ds0, xv.y - ds1};
      out1[off] = b;
      out2[off] = {b.x - ax.x, b.y - ax.y};
    }
  }
  if (EPI == EPI_MATVEC) block_reduce_store<1>(dsum, partial + (size_t)blockIdx.y * gridDim.x);
}

// S-geometry first derivative: out (+)= D(in)   (computeGradient / computeDivergence)
template <typename T, int N, int ADD>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
ks_deriv1(TileS geo, const cplx<T>* __restrict__ in, cplx<T>* out, const cplx<T>* __restrict__ twt) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  const int l = threadIdx.x & (SL - 1), t = threadIdx.x / SL;
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  const int outer = blockIdx.x / geo.nchunk, chunk = blockIdx.x % geo.nchunk;
  const long base = (long)blockIdx.y * geo.batch_stride + (long)outer * geo.outer_stride + (long)chunk * SL + l;
  cplx<T> v[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a)
      v[g * F::R(0) + a] = in[base + (long)F::template loc<0>(t, g, a) * geo.row_stride];
  cplx<T> o[E];
  if (ADD) {  // fetched before the transform so the latency hides behind it
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) o[e] = out[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride];
  }
  deriv_inplace<T, N>(v, tw, sm, AmS{l}, SyncCta{}, t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    cplx<T> s = v[e];
    if (ADD) { s.x += o[e].x; s.y += o[e].y; }
    out[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride] = s;
  }
}

// S-geometry gradient-product sweep: Tk += coef * D(c) . D(p)
// (one axis of the time integrals in src/grad/DerivativeOperators.cpp:206-228)
template <typename T, int N>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
ks_gradprod(TileS geo, const cplx<T>* __restrict__ c, const cplx<T>* __restrict__ p, cplx<T>* Tk, T coef,
            const cplx<T>* __restrict__ twt) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  const int l = threadIdx.x & (SL - 1), t = threadIdx.x / SL;
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  const int outer = blockIdx.x / geo.nchunk, chunk = blockIdx.x % geo.nchunk;
  const long base = (long)outer * geo.outer_stride + (long)chunk * SL + l;
  cplx<T> v[E], u[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) {
      const long off = base + (long)F::template loc<0>(t, g, a) * geo.row_stride;
      v[g * F::R(0) + a] = c[off];
      u[g * F::R(0) + a] = p[off];
    }
  deriv_inplace<T, N>(v, tw, sm, AmS{l}, SyncCta{}, t);
  deriv_inplace<T, N>(u, tw, sm, AmS{l}, SyncCta{}, t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {  // all loads of Tk first, then all stores
    const cplx<T> o = Tk[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride];
    v[e] = {o.x + coef * (v[e].x * u[e].x), o.y + coef * (v[e].y * u[e].y)};
  }
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) Tk[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride] = v[e];
}

// S-geometry complex transform along the tile axis, in place capable.
// DIR = -1: natural rows -> frequency rows (natural frequency order); +1: inverse.
template <typename T, int N, int DIR>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
ks_c2c(TileS geo, const cplx<T>* in, cplx<T>* out, const cplx<T>* __restrict__ twt, const int* __restrict__ done) {
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  const int l = threadIdx.x & (SL - 1), t = threadIdx.x / SL;
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  const int outer = blockIdx.x / geo.nchunk, chunk = blockIdx.x % geo.nchunk;
  const long base = (long)blockIdx.y * geo.batch_stride + (long)outer * geo.outer_stride + (long)chunk * SL + l;
  cplx<T> v[E];
  if (DIR < 0) {
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(0); ++g)
      GLIA_UNROLL
      for (int a = 0; a < F::R(0); ++a)
        v[g * F::R(0) + a] = in[base + (long)F::template loc<0>(t, g, a) * geo.row_stride];
    F::forward(v, tw, sm, AmS{l}, SyncCta{}, t);
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(F::P - 1); ++g) {
      const int kb = F::kbase(t, g);
      GLIA_UNROLL
      for (int cc = 0; cc < F::RL; ++cc) out[base + (long)(kb + F::KSTEP * cc) * geo.row_stride] = v[g * F::RL + cc];
    }
  } else {
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(F::P - 1); ++g) {
      const int kb = F::kbase(t, g);
      GLIA_UNROLL
      for (int cc = 0; cc < F::RL; ++cc) v[g * F::RL + cc] = in[base + (long)(kb + F::KSTEP * cc) * geo.row_stride];
    }
    F::inverse(v, tw, sm, AmS{l}, SyncCta{}, t);
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(0); ++g)
      GLIA_UNROLL
      for (int a = 0; a < F::R(0); ++a)
        out[base + (long)F::template loc<0>(t, g, a) * geo.row_stride] = v[g * F::R(0) + a];
  }
}

// Preconditioner symbol, reference semantics (src/pde/DiffusionSolver.cpp:143-172,
// src/cuda/DiffCoef.cu:11-63): ScalarType products k*w*w, double sum and
// 1 + 0.25*dt*(...), rounded to ScalarType, then factor / that (0-guarded).
template <typename T>
struct PcSym {
  T dt, kxx, kyy, kzz, factor;
};
template <typename T>
__device__ __forceinline__ T pc_symbol(const PcSym<T>& s, int wx, double syz) {
  const T txx = (s.kxx * (T)wx) * (T)wx;
  const double sum = (double)txx + syz;
  const T pf = (T)(1.0 + 0.25 * (double)s.dt * sum);
  return (pf == (T)0) ? s.factor : s.factor / pf;
}

// S-geometry x sweep of the preconditioner on the packed half spectrum
// [n0][n1][n2/2] (y and z already transformed): forward_x . P_hat . inverse_x.
// Column 0 of z packs the DC and Nyquist planes; both have wz = 0 (trap T1), so one
// real symbol applies to the packed complex value.
template <typename T, int N>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
ks_pc(TileS geo, cplx<T>* shat, const cplx<T>* __restrict__ twt, PcSym<T> sym, int n1, const int* __restrict__ done) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  if (done && *done) return;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  const int l = threadIdx.x & (SL - 1), t = threadIdx.x / SL;
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  const int ky = blockIdx.x / geo.nchunk, chunk = blockIdx.x % geo.nchunk;
  const long base = (long)blockIdx.y * geo.batch_stride + (long)ky * geo.outer_stride + (long)chunk * SL + l;
  const int kz = chunk * SL + l;  // 0 .. n2/2-1 ; slot 0 = DC + Nyquist, wz = 0 for both
  const int wy = wavenumber(ky, n1), wz = kz;
  const T tyy = (sym.kyy * (T)wy) * (T)wy, tzz = (sym.kzz * (T)wz) * (T)wz;
  cplx<T> v[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a)
      v[g * F::R(0) + a] = shat[base + (long)F::template loc<0>(t, g, a) * geo.row_stride];
  F::forward(v, tw, sm, AmS{l}, SyncCta{}, t);
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(F::P - 1); ++g) {
    const int kb = F::kbase(t, g);
    GLIA_UNROLL
    for (int cc = 0; cc < F::RL; ++cc) {
      const int wx = wavenumber(kb + F::KSTEP * cc, N);
      // reference order: ((txx + 0 + 0 + 0) + tyy) + tzz in double
      const T txx = (sym.kxx * (T)wx) * (T)wx;
      const double sum = ((double)txx + (double)tyy) + (double)tzz;
      const T pf = (T)(1.0 + 0.25 * (double)sym.dt * sum);
      const T pw = (pf == (T)0) ? sym.factor : sym.factor / pf;
      v[g * F::RL + cc].x *= pw;
      v[g * F::RL + cc].y *= pw;
    }
  }
  F::inverse(v, tw, sm, AmS{l}, SyncCta{}, t);
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a)
      shat[base + (long)F::template loc<0>(t, g, a) * geo.row_stride] = v[g * F::R(0) + a];
}

// ================================================================ Z ====
// common Z-geometry prologue
template <typename T, int N>
struct ZCtx {
  using F = LineFft<T, N, zplan<N>()>;
  static constexpr int TPL = F::TPL, LPC = zlines<N>();
  int t, lp;
  long pair;
  bool active;
  __device__ __forceinline__ ZCtx(const LinesZ& ln) {
    t = threadIdx.x % TPL;
    lp = threadIdx.x / TPL;
    pair = (long)blockIdx.x * LPC + lp;
    active = pair < ln.npairs;
    if (!active) pair = ln.npairs - 1;  // keep addresses valid; stores are predicated
  }
  __device__ __forceinline__ AmZ am() const { return AmZ{lp * zpad<N>()}; }
};

// Z-geometry second derivative: acc (+)= D_z(k D_z x)  (first sweep of applyD; ADD: second
// sweep of the slab-decomposed applyD, whose x sweep runs first)
template <typename T, int N, int ADD = 0, int MINB = 1>
__global__ void __launch_bounds__(zthreads<N>(), MINB)
kz_deriv2(LinesZ ln, const T* __restrict__ x, const T* __restrict__ kf, T* acc, const cplx<T>* __restrict__ twt,
          const int* __restrict__ done) {
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  ZCtx<T, N> z(ln);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, z.t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  typename ZSync<F::TPL>::type sy;
  const long la = z.pair * 2 * N, lb = la + N;
  cplx<T> v[E], kk[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) {
      const int pos = F::template loc<0>(z.t, g, a);
      v[g * F::R(0) + a] = {ld_stream(x + la + pos), ld_stream(x + lb + pos)};
      kk[g * F::R(0) + a] = {ld_stream(kf + la + pos), ld_stream(kf + lb + pos)};
    }
  deriv_inplace<T, N, zplan<N>()>(v, tw, sm, z.am(), sy, z.t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) { v[e].x *= kk[e].x; v[e].y *= kk[e].y; }
  if (ADD) {  // the accumulator is fetched now so that its latency hides behind the second derivative
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      kk[e] = {acc[la + pos], acc[lb + pos]};
    }
  }
  deriv_inplace<T, N, zplan<N>()>(v, tw, sm, z.am(), sy, z.t);
  if (z.active) {
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(0); ++g)
      GLIA_UNROLL
      for (int a = 0; a < F::R(0); ++a) {
        const int pos = F::template loc<0>(z.t, g, a);
        cplx<T> s = v[g * F::R(0) + a];
        if (ADD) { s.x += kk[g * F::R(0) + a].x; s.y += kk[g * F::R(0) + a].y; }
        acc[la + pos] = s.x;
        acc[lb + pos] = s.y;
      }
  }
}

// Z-geometry first derivative: out (+)= D_z(in)
template <typename T, int N, int ADD>
__global__ void __launch_bounds__(zthreads<N>())
kz_deriv1(LinesZ ln, const T* __restrict__ in, T* out, const cplx<T>* __restrict__ twt) {
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  ZCtx<T, N> z(ln);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, z.t);
  typename ZSync<F::TPL>::type sy;
  const long la = z.pair * 2 * N, lb = la + N;
  cplx<T> v[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) {
      const int pos = F::template loc<0>(z.t, g, a);
      v[g * F::R(0) + a] = {in[la + pos], in[lb + pos]};
    }
  cplx<T> o[E];
  if (ADD) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      o[e] = {out[la + pos], out[lb + pos]};
    }
  }
  deriv_inplace<T, N, zplan<N>()>(v, tw, sm, z.am(), sy, z.t);
  if (z.active) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      T ox = v[e].x, oy = v[e].y;
      if (ADD) { ox += o[e].x; oy += o[e].y; }
      out[la + pos] = ox;
      out[lb + pos] = oy;
    }
  }
}

// Z-geometry gradient-product sweep: Tk += coef * D_z c . D_z p ;
// Tr += coef * p * (c*c - c)   (src/grad/DerivativeOperators.cpp:206-228, 275-291)
template <typename T, int N>
__global__ void __launch_bounds__(zthreads<N>())
kz_gradprod(LinesZ ln, const T* __restrict__ c, const T* __restrict__ p, T* Tk, T* Tr, T coef,
            const cplx<T>* __restrict__ twt) {
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  ZCtx<T, N> z(ln);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, z.t);
  typename ZSync<F::TPL>::type sy;
  const long la = z.pair * 2 * N, lb = la + N;
  cplx<T> v[E], u[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) {
      const int pos = F::template loc<0>(z.t, g, a);
      v[g * F::R(0) + a] = {c[la + pos], c[lb + pos]};
      u[g * F::R(0) + a] = {p[la + pos], p[lb + pos]};
    }
  if (Tr && z.active) {
    cplx<T> o[E];
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      o[e] = {Tr[la + pos], Tr[lb + pos]};
    }
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      // work = c*c ; work -= c ; work = p*work ; temp += dt*w*work
      T wa = v[e].x * v[e].x; wa = wa - v[e].x; wa = u[e].x * wa;
      T wb = v[e].y * v[e].y; wb = wb - v[e].y; wb = u[e].y * wb;
      Tr[la + pos] = o[e].x + coef * wa;
      Tr[lb + pos] = o[e].y + coef * wb;
    }
  }
  deriv_inplace<T, N, zplan<N>()>(v, tw, sm, z.am(), sy, z.t);
  deriv_inplace<T, N, zplan<N>()>(u, tw, sm, z.am(), sy, z.t);
  if (z.active) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      const T ox = Tk[la + pos], oy = Tk[lb + pos];
      v[e] = {ox + coef * (v[e].x * u[e].x), oy + coef * (v[e].y * u[e].y)};
    }
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      Tk[la + pos] = v[e].x;
      Tk[lb + pos] = v[e].y;
    }
  }
}

// Z-geometry real-to-complex: two real lines -> two packed half spectra
// (N/2 complex each; slot 0 = {DC, Nyquist}).  Optional fused PCG prologue
// r <- r - a*w (VecAXPY(R,-a,W) of KSPSolve_CG), with r written back.
template <typename T, int N, int PRO>
__global__ void __launch_bounds__(zthreads<N>())
kz_r2c(LinesZ ln, T* r, const T* __restrict__ w, const double* __restrict__ scal_a, cplx<T>* shat,
       const cplx<T>* __restrict__ twt, const int* __restrict__ done, int bpm) {
  const int member = blockIdx.x / bpm;  // bpm = CTAs per ensemble member
  done = member_done(done, member);
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  ZCtx<T, N> z(ln);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, z.t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  typename ZSync<F::TPL>::type sy;
  const long la = z.pair * 2 * N, lb = la + N;
  T aa = (T)0;
  if (PRO) aa = (T)(scal_a[(size_t)member * SCAL_STRIDE]);
  cplx<T> v[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) {
      const int pos = F::template loc<0>(z.t, g, a);
      // r and w are streamed: neither is re-read before most of L2 has turned over
      cplx<T> rv = {ld_stream(r + la + pos), ld_stream(r + lb + pos)};
      if (PRO) {
        rv.x = rv.x - aa * ld_stream(w + la + pos);
        rv.y = rv.y - aa * ld_stream(w + lb + pos);
        if (z.active) { r[la + pos] = rv.x; r[lb + pos] = rv.y; }
      }
      v[g * F::R(0) + a] = rv;
    }
  F::forward(v, tw, sm, z.am(), sy, z.t);
  // scatter by frequency, then untangle the two real spectra
  const int sb = z.lp * zpad<N>();
  sy();
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(F::P - 1); ++g) {
    const int kb = F::kbase(z.t, g);
    GLIA_UNROLL
    for (int cc = 0; cc < F::RL; ++cc) {
      const int k = kb + F::KSTEP * cc;
      sm[sb + k + (k >> 4)] = v[g * F::RL + cc];
    }
  }
  sy();
  const long oa = z.pair * 2 * (N / 2), ob = oa + N / 2;
  GLIA_UNROLL
  for (int j = 0; j < E / 2; ++j) {
    const int k = z.t + F::TPL * j;
    const int kn = (N - k) & (N - 1);
    const cplx<T> zk = sm[sb + k + (k >> 4)], zn = sm[sb + kn + (kn >> 4)];
    cplx<T> A = {(T)0.5 * (zk.x + zn.x), (T)0.5 * (zk.y - zn.y)};
    cplx<T> B = {(T)0.5 * (zk.y + zn.y), (T)-0.5 * (zk.x - zn.x)};
    if (k == 0) {
      const cplx<T> zh = sm[sb + N / 2 + ((N / 2) >> 4)];
      A = {zk.x, zh.x};
      B = {zk.y, zh.y};
    }
    if (z.active) { shat[oa + k] = A; shat[ob + k] = B; }
  }
}

// Z-geometry complex-to-real of the packed half spectrum, with the PCG epilogue:
// partial sums {<z,z>, <r,z>} (VecNorm(Z), VecXDot(Z,R) of KSPSolve_CG).
//   zout may be null (only the norm is wanted, e.g. rnorm0 = ||M^-1 b||).
//   EPI 0: no sums; 1: <z,z> (and <r,z> with plain loads if r); 2: <z,z>, <r,z> with the two r lines of a pair
//   staged through shared memory by cp.async while the inverse transform runs (the epilogue's r loads were this
//   kernel's top stall: 56 % long-scoreboard, profiles/r1c_ncu_source_summary.txt)
template <int N, typename T>
__host__ __device__ constexpr size_t smem_z_rstage() { return (size_t)zlines<N>() * 2 * N * sizeof(T); }
template <typename T, int N, int EPI>
__global__ void __launch_bounds__(zthreads<N>())
kz_c2r(LinesZ ln, const cplx<T>* __restrict__ shat, T* zout, const T* __restrict__ r, double* partial,
       const cplx<T>* __restrict__ twt, const int* __restrict__ done, int bpm) {
  done = member_done(done, blockIdx.x / bpm);  // bpm = CTAs per ensemble member; partial[blockIdx.x] is the member's
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  ZCtx<T, N> z(ln);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, z.t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  typename ZSync<F::TPL>::type sy;
  const long la = z.pair * 2 * N, lb = la + N;
  const long oa = z.pair * 2 * (N / 2), ob = oa + N / 2;
  AmZ am = z.am();
  // 512-point lines: the lines of the CTA that follows in this slot go towards L2 now (this kernel's top stall is the
  // latency of its first loads: long-scoreboard 28 % of the samples, profiles/r2l_ncu_source_summary.txt).  Measured
  // (profiles/r2t_div_prefetch_ab.txt): kz_c2r.rz 360 -> 332 us at 512^3, but 45.4 -> 49.4 at 256^3, and the same
  // prefetch in kz_r2c.axpy (which writes r back) tripled that kernel: N = 512 here only.
  if constexpr (N >= 512) {
    prefetch_next_cta(shat, (long)zlines<N>() * N * (long)sizeof(cplx<T>), ln.npairs * N * (long)sizeof(cplx<T>));
    if (EPI && r) prefetch_next_cta(r, (long)zlines<N>() * 2 * N * (long)sizeof(T), ln.npairs * 2 * N * (long)sizeof(T));
  }
  [[maybe_unused]] T* rst = nullptr;
  if constexpr (EPI == 2) {
    rst = reinterpret_cast<T*>(smraw + sizeof(cplx<T>) * zlines<N>() * zpad<N>()) + (size_t)z.lp * 2 * N;
    constexpr int PER = 16 / (int)sizeof(T), CH = 2 * N / PER;
    for (int i = z.t; i < CH; i += F::TPL) cp_async16(rst + i * PER, r + la + i * PER);
    cp_async_commit();
  } else if (EPI && r) {
    constexpr int PER = 128 / (int)sizeof(T);
    for (int i = z.t; i < 2 * N / PER; i += F::TPL) prefetch_l1(r + la + i * PER);
  }
  sy();
  GLIA_UNROLL
  for (int j = 0; j < E / 2; ++j) {
    const int k = z.t + F::TPL * j;
    const cplx<T> A = ld_stream(shat + oa + k), B = ld_stream(shat + ob + k);  // last read of shat
    if (k == 0) {
      sm[am(F::loc_of_freq(0))] = {A.x, B.x};
      sm[am(F::loc_of_freq(N / 2))] = {A.y, B.y};
    } else {
      sm[am(F::loc_of_freq(k))] = {A.x - B.y, A.y + B.x};
      sm[am(F::loc_of_freq(N - k))] = {A.x + B.y, B.x - A.y};
    }
  }
  sy();
  cplx<T> v[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(F::P - 1); ++g)
    GLIA_UNROLL
    for (int cc = 0; cc < F::RL; ++cc) v[g * F::RL + cc] = sm[am(F::template loc<F::P - 1>(z.t, g, cc))];
  F::inverse(v, tw, sm, am, sy, z.t);
  if constexpr (EPI == 2) {
    cp_async_wait<0>();
    sy();  // the r pieces fetched by the other threads of this line pair
  }
  // Dot products.  Double precision: every product and sum in double.  Single precision: the 2 E
  // products of ONE thread are summed in float (two FFMA chains per sum), everything across threads, CTAs and ranks
  // in double.  The all-double form cost this kernel 64 F2F + 119 DADD / DMUL / DFMA per thread next to 412 FP32
  // instructions (conversion and FP64 pipes at a fraction of the FP32 rate: short-scoreboard 17 % of its stall
  // samples, profiles/r2l_ncu_source_summary.txt).  A thread's 32-term float sum carries a relative error of a few
  // 1e-7 with random sign; over the 5e5 threads of a 256^3 field the dot product moves by ~1e-9 relative -- the
  // reference itself (PETSc VecDot on float Vecs) accumulates everything in float.
  double acc[2] = {0.0, 0.0};
  [[maybe_unused]] T fa[4] = {(T)0, (T)0, (T)0, (T)0};
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) {
      const int pos = F::template loc<0>(z.t, g, a);
      const cplx<T> zv = v[g * F::R(0) + a];
      if (z.active) {
        if (zout) { zout[la + pos] = zv.x; zout[lb + pos] = zv.y; }
        if (EPI) {
          if constexpr (sizeof(T) == 8) {
            acc[0] += (double)zv.x * (double)zv.x + (double)zv.y * (double)zv.y;
            if constexpr (EPI == 2) acc[1] += (double)rst[pos] * (double)zv.x + (double)rst[N + pos] * (double)zv.y;
            else if (r) acc[1] += (double)ld_stream(r + la + pos) * (double)zv.x + (double)ld_stream(r + lb + pos) * (double)zv.y;
          } else {
            fa[0] = zv.x * zv.x + fa[0];
            fa[1] = zv.y * zv.y + fa[1];
            if constexpr (EPI == 2) {
              fa[2] = rst[pos] * zv.x + fa[2];
              fa[3] = rst[N + pos] * zv.y + fa[3];
            } else if (r) {
              fa[2] = ld_stream(r + la + pos) * zv.x + fa[2];
              fa[3] = ld_stream(r + lb + pos) * zv.y + fa[3];
            }
          }
        }
      }
    }
  if constexpr (sizeof(T) != 8) {
    acc[0] = (double)fa[0] + (double)fa[1];
    acc[1] = (double)fa[2] + (double)fa[3];
  }
  if (EPI) block_reduce_store<2>(acc, partial);
}

}  // namespace glia
