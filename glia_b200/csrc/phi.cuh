// phi.cuh -- the callers either side of the RD path: the Weierstrass smoother,
// the Gaussian basis Phi (c(0) = Phi p, g_p = Phi^T g) and the MatProp filter.
//
// The reference smooths with two 3-D FFTs and a spectral product against the FFT of a
// periodised Gaussian (SpectralOperators::weierstrassSmoother, src/grad/SpectralOperators.cpp:
// 295-381).  That Gaussian -- the sum of 8 images exp(-(x^2+y^2+z^2)/2s^2) -- is exactly the
// tensor product g_x(x) g_y(y) g_z(z), g_d(X) = exp(-X^2/2s^2) + exp(-(X-2pi)^2/2s^2), so its
// normalised convolution is three 1-D circular convolutions.  Each is one axis sweep
// "line FFT . real even symbol . inverse line FFT" on the same sweep engine as the derivative
// operators; g_d is even on the periodic grid, its DFT is real, and two real lines ride through
// one complex transform unmixed.
//
// Phi::apply / applyTranspose in on-the-fly mode (src/mat/Phi.cpp:324-434) build every basis
// function as  truncate_{5 sigma}( W_sigma( Gaussian_i . filter ) ): here the Gaussian is
// generated inside the z sweep's load (Phi::initialize, Phi.cpp:262-320), and the truncation
// (Phi::truncate, Phi.cpp:237-260), the running maximum (vecMax) and either out += p_i phi_i
// (VecAXPY) or <phi_i, in> (VecDot) ride in the x sweep's epilogue.
#pragma once
#include "sweeps.cuh"

namespace glia {

__device__ __forceinline__ float g_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double g_sqrt(double x) { return sqrt(x); }

template <typename T>
struct GaussPhi {  // Phi::initialize / truncate parameters, all in ScalarType like the reference
  T xc, yc, zc;    // centre
  T hx, hy, hz;    // twopi / n
  T R;             // sqrt(2) sigma
  T sigma;
};
template <typename T>
__device__ __forceinline__ T phi_radius(const GaussPhi<T>& gp, int X, int Y, int Z) {
  const T dx = gp.hx * (T)X - gp.xc, dy = gp.hy * (T)Y - gp.yc, dz = gp.hz * (T)Z - gp.zc;
  return g_sqrt(dx * dx + dy * dy + dz * dz);
}
template <typename T>
__device__ __forceinline__ T phi_gauss(const GaussPhi<T>& gp, int X, int Y, int Z) {
  const T ratio = phi_radius(gp, X, Y, Z) / gp.R;
  return g_exp(-ratio * ratio);
}

// v <- s(k) v on the frequency placement forward() leaves
template <typename T, int N>
__device__ __forceinline__ void mult_symbol(cplx<T> (&v)[FftPlan<N>::E], const T* __restrict__ symtab, int t) {
  using F = LineFft<T, N>;
  GLIA_UNROLL
  for (int g = 0; g < F::Gp(F::P - 1); ++g) {
    const int kb = F::kbase(t, g);
    GLIA_UNROLL
    for (int c = 0; c < F::RL; ++c) {
      const T s = symtab[kb + F::KSTEP * c];
      v[g * F::RL + c].x *= s;
      v[g * F::RL + c].y *= s;
    }
  }
}

// Z-geometry filter sweep: out = S_z(in).  GAUSS: the input is generated, in = Gaussian . filter
// (`in` = filter field, null = no filter) -- Phi::initialize + VecPointwiseMult(filter).
template <typename T, int N, int GAUSS>
__global__ void __launch_bounds__(zthreads<N>())
kz_filter(LinesZ ln, const T* __restrict__ in, T* out, const T* __restrict__ symtab, const cplx<T>* __restrict__ twt,
          GaussPhi<T> gp, int n1) {
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  ZCtx<T, N> z(ln);
  typename F::Tw tw;
  F::load_twiddles(tw, twt, z.t);
  typename ZSync<F::TPL>::type sy;
  const long la = z.pair * 2 * N, lb = la + N;
  const long line = z.pair * 2;
  const int X = (int)(line / n1), Y = (int)(line % n1);
  cplx<T> v[E];
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
    if (GAUSS) {
      cplx<T> gv = {phi_gauss(gp, X, Y, pos), phi_gauss(gp, X, Y + 1, pos)};
      if (in) { gv.x = in[la + pos] * gv.x; gv.y = in[lb + pos] * gv.y; }
      v[e] = gv;
    } else {
      v[e] = {in[la + pos], in[lb + pos]};
    }
  }
  F::forward(v, tw, sm, z.am(), sy, z.t);
  mult_symbol<T, N>(v, symtab, z.t);
  F::inverse(v, tw, sm, z.am(), sy, z.t);
  if (z.active) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(z.t, e / F::R(0), e % F::R(0));
      out[la + pos] = v[e].x;
      out[lb + pos] = v[e].y;
    }
  }
}

// S-geometry filter sweep along y or x (in place capable), with the Phi epilogues on the last axis:
//   MODE 0  out = S(in)
//   MODE 1  phi = truncate(S(in)); block max -> pmax; acc += coef * phi          (Phi::apply)
//   MODE 2  phi = truncate(S(in)); block max -> pmax; block <phi, acc> -> pdot   (Phi::applyTranspose)
// The tile is an x sweep in modes 1, 2: rows = x, outer = y, columns = z pairs.
template <typename T, int N, int MODE>
__global__ void __launch_bounds__(SL* (N / FftPlan<N>::E))
ks_filter(TileS geo, const cplx<T>* in, cplx<T>* out, const T* __restrict__ symtab, const cplx<T>* __restrict__ twt,
          GaussPhi<T> gp, cplx<T>* acc, T coef, double* pmax, double* pdot) {
  using F = LineFft<T, N>;
  constexpr int E = F::E;
  GLIA_DYN_SMEM(smraw);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(smraw);
  const int l = threadIdx.x & (SL - 1), t = threadIdx.x / SL;
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  const int outer = blockIdx.x / geo.nchunk, chunk = blockIdx.x % geo.nchunk;
  const long base = (long)outer * geo.outer_stride + (long)chunk * SL + l;
  cplx<T> v[E];
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) v[e] = in[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride];
  cplx<T> a[E];
  if (MODE != 0) {  // fetched before the transform so the latency hides behind it
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) a[e] = acc[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride];
  }
  F::forward(v, tw, sm, AmS{l}, SyncCta{}, t);
  mult_symbol<T, N>(v, symtab, t);
  F::inverse(v, tw, sm, AmS{l}, SyncCta{}, t);
  if (MODE == 0) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) out[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride] = v[e];
    return;
  }
  const int Z0 = 2 * (chunk * SL + l);
  double red[2] = {0.0, 0.0};  // {max, dot}; phi >= 0 up to rounding, the reference's max starts at 0 too
  GLIA_UNROLL
  for (int e = 0; e < E; ++e) {
    const int X = F::template loc<0>(t, e / F::R(0), e % F::R(0));
    // truncate to zero after radius 5*sigma (Phi.cpp:256)
    cplx<T> ph = v[e];
    if (!(phi_radius(gp, X, outer, Z0) / gp.sigma <= (T)5)) ph.x = (T)0;
    if (!(phi_radius(gp, X, outer, Z0 + 1) / gp.sigma <= (T)5)) ph.y = (T)0;
    red[0] = fmax(red[0], fmax((double)ph.x, (double)ph.y));
    if (MODE == 1) {
      a[e] = {a[e].x + coef * ph.x, a[e].y + coef * ph.y};
    } else {
      red[1] += (double)ph.x * (double)a[e].x + (double)ph.y * (double)a[e].y;
    }
  }
  if (MODE == 1) {
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) acc[base + (long)F::template loc<0>(t, e / F::R(0), e % F::R(0)) * geo.row_stride] = a[e];
  }
  // block reduction: max and sum
  __shared__ double sred[32 * 2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  double m = red[0], s = red[1];
  for (int o = 16; o > 0; o >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if (lane == 0) { sred[wid * 2] = m; sred[wid * 2 + 1] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mm = 0.0, ss = 0.0;
    for (int w = 0; w < nwarp; ++w) { mm = fmax(mm, sred[w * 2]); ss += sred[w * 2 + 1]; }
    pmax[blockIdx.x] = mm;
    if (MODE == 2) pdot[blockIdx.x] = ss;
  }
}

// one CTA: running maximum and (optionally) the dot product of one basis function
static __global__ void k_phi_reduce(const double* pmax, const double* pdot, int n, double* run_max, double* dot_out) {
  __shared__ double sh[2 * 256];
  const int tid = threadIdx.x;
  double m = 0.0, s = 0.0;
  for (int j = tid; j < n; j += 256) {
    m = fmax(m, pmax[j]);
    if (pdot) s += pdot[j];
  }
  sh[tid] = m;
  sh[256 + tid] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (tid < st) {
      sh[tid] = fmax(sh[tid], sh[tid + st]);
      sh[256 + tid] += sh[256 + tid + st];
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (sh[0] > *run_max) *run_max = sh[0];
    if (dot_out) *dot_out = sh[256];
  }
}

// out *= (T)(1.0 / phi_max)   (VecScale(out, 1.0 / phi_max), Phi.cpp:378)
template <typename T>
__global__ void k_scale_inv(long n, T* out, const double* phi_max) {
  const T pm = (T)(*phi_max);
  const T alpha = (T)(1.0 / (double)pm);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = out[i] * alpha;
}

// MatProp::setValuesCustom (src/mat/MatProp.cpp:135-201): clip the tissue maps at 0 in place,
// bg = 1 - (gm + wm + vt + csf), filter = (wm > 0.1 || gm > 0.1) && vt < 0.8.  Absent maps are
// null (treated as zero).  partial: per-block sum of the filter (DiffCoef's k-bar divisor, trap T5).
template <typename T>
__global__ void k_mat_prop(long n, T* gm, T* wm, T* vt, T* csf, T* bg, T* filter, double* partial) {
  double fs = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T g = gm ? gm[i] : (T)0, w = wm ? wm[i] : (T)0, v = vt ? vt[i] : (T)0, c = csf ? csf[i] : (T)0;
    g = (g <= (T)0) ? (T)0 : g;
    w = (w <= (T)0) ? (T)0 : w;
    v = (v <= (T)0) ? (T)0 : v;
    c = (c <= (T)0) ? (T)0 : c;
    if (gm) gm[i] = g;
    if (wm) wm[i] = w;
    if (vt) vt[i] = v;
    if (csf) csf[i] = c;
    T b = g + w;
    b = b + v;
    b = b + c;
    if (bg) bg[i] = -(b - (T)1.0);
    const T f = ((w > (T)0.1 || g > (T)0.1) && v < (T)0.8) ? (T)1 : (T)0;
    if (filter) filter[i] = f;
    fs += (double)f;
  }
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
  if (lane == 0) red[wid] = fs;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < nwarp; ++w) s += red[w];
    partial[(size_t)blockIdx.x * 4 + 0] = 0.0;
    partial[(size_t)blockIdx.x * 4 + 1] = 0.0;
    partial[(size_t)blockIdx.x * 4 + 2] = 0.0;
    partial[(size_t)blockIdx.x * 4 + 3] = s;
  }
}

// splitSegmentation, atlas form (src/utils/Utils.cpp:592-657 with tu == ed == nullptr): one-hot
// tissue maps from a label image; a label <= 0 leaves its map zero; null outputs are skipped.
template <typename T>
__global__ void k_split_seg(long n, const T* __restrict__ seg, int wm_l, int gm_l, int vt_l, int csf_l, T* wm, T* gm, T* vt,
                            T* csf) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T s = seg[i];
    if (wm) wm[i] = (wm_l > 0 && s == (T)wm_l) ? (T)1 : (T)0;
    if (gm) gm[i] = (gm_l > 0 && s == (T)gm_l) ? (T)1 : (T)0;
    if (vt) vt[i] = (vt_l > 0 && s == (T)vt_l) ? (T)1 : (T)0;
    if (csf) csf[i] = (csf_l > 0 && s == (T)csf_l) ? (T)1 : (T)0;
  }
}

}  // namespace glia
