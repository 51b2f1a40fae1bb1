// sweeps_dist.cuh -- x-axis sweeps of the slab-decomposed (multi-GPU) path.
//
// The grid is cut into G slabs along x (rank q owns x-planes [q*n0/G, (q+1)*n0/G), the layout
// AccFFT produces for c_dims = {G, 1}: src/grad/SpectralOperators.cpp:398-421, include/
// Parameters.h:374-451).  y and z lines are rank-local; an x line crosses every slab.  Where
// the reference transposes the whole field with MPI all-to-all around each x transform, the
// kernels here fetch each 128-byte row of their tile straight from the HBM of the rank that
// owns it and store the result rows back the same way (NVLink / NVSwitch peer access through
// CUDA IPC mappings) -- transform + exchange in one kernel, nothing packed or staged.
//
// Work split: rank q sweeps the tiles with y in [q*n1/G, (q+1)*n1/G), all z chunks.
// Fields read by an x sweep but never exchanged (the diffusion coefficient) are kept in a
// second, rank-local "pencil" copy [n0][n1/G][n2/2] built once per coefficient update.
#pragma once
#include "comm.cuh"
#include "sweeps.cuh"

namespace glia {

template <typename T>
struct PeerRows {  // base address of one field in every rank's arena (slab layout)
  cplx<T>* base[MAX_RANKS];
};

struct TileX {
  long slab_row_stride;    // n1 * n2c       (complex units between x rows inside a slab)
  long slab_outer_stride;  // n2c            (between y)
  long pen_row_stride;     // (n1/G) * n2c   (pencil copy: between x rows)
  int nchunk;              // n2c / SL (a power of two)
  int cshift;              // log2(nchunk)
  int n_outer;             // n1 / G   tiles per chunk on this rank
  int y0;                  // rank * n1 / G
  int shift, mask;         // x row -> (owner = row >> shift, local row = row & mask)
};

template <typename T, int N>
struct XCtx {
  using F = LineFft<T, N>;
  int l, t, yl, chunk;
  long slab_off, pen_off;
  __device__ __forceinline__ XCtx(const TileX& g) {
    l = threadIdx.x & (SL - 1);
    t = threadIdx.x / SL;
    yl = blockIdx.x / g.nchunk;
    chunk = blockIdx.x % g.nchunk;
    slab_off = (long)(g.y0 + yl) * g.slab_outer_stride + (long)chunk * SL + l;
    pen_off = (long)yl * g.slab_outer_stride + (long)chunk * SL + l;
  }
  template <class P>
  __device__ __forceinline__ auto* slab(const P& p, const TileX& g, int row) const {
    return p.base[row >> g.shift] + (long)(row & g.mask) * g.slab_row_stride + slab_off;
  }
  __device__ __forceinline__ long pen(const TileX& g, int row) const { return (long)row * g.pen_row_stride + pen_off; }
};

// acc(slabs) = D_x(kT . D_x x(slabs))        first sweep of the distributed applyD
template <typename T, int N>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E), s_min_ctas<T, N>())
kx_deriv2_dist(TileX geo, PeerRows<T> x, const cplx<T>* __restrict__ kT, PeerRows<T> acc,
               const cplx<T>* __restrict__ twt, const int* __restrict__ done) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  if (done && *done) return;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  XCtx<T, N> c(geo);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, c.t);
  AmS am{c.l};
  SyncCta sy;
  cplx<T> v[E], kk[E];
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const int row = F::template loc<0>(c.t, e / F::R(0), e % F::R(0));
    v[e] = *c.slab(x, geo, row);
    kk[e] = kT[c.pen(geo, row)];
  }
  deriv_inplace<T, N>(v, tw, sm, am, sy, c.t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) { v[e].x *= kk[e].x; v[e].y *= kk[e].y; }
  deriv_inplace<T, N>(v, tw, sm, am, sy, c.t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const int row = F::template loc<0>(c.t, e / F::R(0), e % F::R(0));
    *c.slab(acc, geo, row) = v[e];
  }
}

// out(slabs) (+)= D_x(in(slabs))             x component of computeGradient / computeDivergence
template <typename T, int N, int ADD>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
kx_deriv1_dist(TileX geo, PeerRows<T> in, PeerRows<T> out, const cplx<T>* __restrict__ twt) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  XCtx<T, N> c(geo);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, c.t);
  cplx<T> v[E], o[E];
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) v[e] = *c.slab(in, geo, F::template loc<0>(c.t, e / F::R(0), e % F::R(0)));
  if (ADD) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) o[e] = *c.slab(out, geo, F::template loc<0>(c.t, e / F::R(0), e % F::R(0)));
  }
  deriv_inplace<T, N>(v, tw, sm, AmS{c.l}, SyncCta{}, c.t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    cplx<T> s = v[e];
    if (ADD) { s.x += o[e].x; s.y += o[e].y; }
    *c.slab(out, geo, F::template loc<0>(c.t, e / F::R(0), e % F::R(0))) = s;
  }
}

// x sweep of the preconditioner on the packed half spectrum held slab-wise by all ranks:
// forward_x . P_hat . inverse_x, in place in the owners' memory.
template <typename T, int N>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
kx_pc_dist(TileX geo, PeerRows<T> shat, const cplx<T>* __restrict__ twt, PcSym<T> sym, int n1,
           const int* __restrict__ done) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  if (done && *done) return;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  XCtx<T, N> c(geo);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, c.t);
  const int ky = geo.y0 + c.yl;
  const int kz = c.chunk * SL + c.l;
  const int wy = wavenumber(ky, n1), wz = kz;
  const T tyy = (sym.kyy * (T)wy) * (T)wy, tzz = (sym.kzz * (T)wz) * (T)wz;
  cplx<T> v[E];
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) v[g * F::R(0) + a] = *c.slab(shat, geo, F::template loc<0>(c.t, g, a));
  F::forward(v, tw, sm, AmS{c.l}, SyncCta{}, c.t);
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(F::P - 1); ++g) {
    const int kb = F::kbase(c.t, g);
    GLIA_UNROLL
    for (int cc = 0; cc < F::RL; ++cc) {
      const int wx = wavenumber(kb + F::KSTEP * cc, N);
      const T txx = (sym.kxx * (T)wx) * (T)wx;
      const double sum = ((double)txx + (double)tyy) + (double)tzz;
      const T pf = (T)(1.0 + 0.25 * (double)sym.dt * sum);
      const T pw = (pf == (T)0) ? sym.factor : sym.factor / pf;
      v[g * F::RL + cc].x *= pw;
      v[g * F::RL + cc].y *= pw;
    }
  }
  F::inverse(v, tw, sm, AmS{c.l}, SyncCta{}, c.t);
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(0); ++g)
    GLIA_UNROLL
    for (int a = 0; a < F::R(0); ++a) *c.slab(shat, geo, F::template loc<0>(c.t, g, a)) = v[g * F::R(0) + a];
}

// TkX(pencil) += coef * D_x c(slabs) . D_x p(slabs)      x part of the gradient time integrals
template <typename T, int N>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
kx_gradprod_dist(TileX geo, PeerRows<T> cf, PeerRows<T> pf, cplx<T>* TkX, T coef, const cplx<T>* __restrict__ twt) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  XCtx<T, N> c(geo);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, c.t);
  cplx<T> v[E], u[E];
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const int row = F::template loc<0>(c.t, e / F::R(0), e % F::R(0));
    v[e] = *c.slab(cf, geo, row);
    u[e] = *c.slab(pf, geo, row);
  }
  deriv_inplace<T, N>(v, tw, sm, AmS{c.l}, SyncCta{}, c.t);
  deriv_inplace<T, N>(u, tw, sm, AmS{c.l}, SyncCta{}, c.t);
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const cplx<T> o = TkX[c.pen(geo, F::template loc<0>(c.t, e / F::R(0), e % F::R(0)))];
    v[e] = {o.x + coef * (v[e].x * u[e].x), o.y + coef * (v[e].y * u[e].y)};
  }
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) TkX[c.pen(geo, F::template loc<0>(c.t, e / F::R(0), e % F::R(0)))] = v[e];
}

// pencil[row][yl][z] = slab field of the owner of `row`      (coefficient copy for the x sweeps)
template <typename T>
__global__ void k_slab_to_pencil(TileX geo, int n0, PeerRows<T> src, cplx<T>* pencil) {
  const long per_row = geo.pen_row_stride;  // (n1/G) * n2c
  const long total = (long)n0 * per_row;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int row = (int)(i / per_row);
    const long rem = i % per_row;
    const long yl = rem / geo.slab_outer_stride, zc = rem % geo.slab_outer_stride;
    pencil[i] = src.base[row >> geo.shift][(long)(row & geo.mask) * geo.slab_row_stride +
                                           (geo.y0 + yl) * geo.slab_outer_stride + zc];
  }
}

// dst(slabs) += pencil       (returns the x part of an accumulator to the slab layout)
template <typename T>
__global__ void k_pencil_add_to_slab(TileX geo, int n0, const cplx<T>* __restrict__ pencil, PeerRows<T> dst) {
  const long per_row = geo.pen_row_stride;
  const long total = (long)n0 * per_row;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int row = (int)(i / per_row);
    const long rem = i % per_row;
    const long yl = rem / geo.slab_outer_stride, zc = rem % geo.slab_outer_stride;
    cplx<T>* d = dst.base[row >> geo.shift] + (long)(row & geo.mask) * geo.slab_row_stride +
                 (geo.y0 + yl) * geo.slab_outer_stride + zc;
    const cplx<T> a = *d, b = pencil[i];
    *d = {a.x + b.x, a.y + b.y};
  }
}

}  // namespace glia
