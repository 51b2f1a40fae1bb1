// sweeps_v2.cuh -- single-precision sweeps on the packed FP32x2 line FFT (fft_core_v2.cuh).
//
// Same tiles, same persistent cp.async pipeline and the same epilogues as sweeps_pipe.cuh, but a
// thread owns TWO adjacent complex columns of the pair view -- one 16-byte piece (z0 z1 z2 z3) of
// a 128-byte row -- and transforms them as A = z0 + i z2, B = z1 + i z3 in lock step, so that the
// butterflies issue as FADD2 / FMUL2 / FFMA2.  Valid for every sweep that applies a real linear
// operator per real z column (D(k D .) and its epilogues); the result per real value is the same
// operator as the scalar kernels', with different rounding partners.
//   tile    N rows x 128 bytes; 8 lanes along z, TPL = N / 8 threads along the sweep axis
//   smem    2 staged x tiles | exchange tile | duplicated twiddle table (N x 32 bytes)
#pragma once
#include "fft_core_v2.cuh"
#include "sweeps_pipe.cuh"

namespace glia {

static constexpr int SL2 = 8;  // 16-byte lanes per 128-byte row

struct AmS2 {  // exchange / stage addressing in C2 units: row * 8 + lane
  int l;
  __device__ __forceinline__ int operator()(int loc) const { return loc * SL2 + l; }
};

template <int N>
__host__ __device__ constexpr size_t v2_smem() { return 3 * (size_t)N * 128 + (size_t)N * sizeof(TwDup); }
template <int N>
__host__ __device__ constexpr int v2_ctas() { return (2 * v2_smem<N>() <= 224 * 1024 && N <= 256) ? 2 : 1;
}

__device__ __forceinline__ C2 c2_mul(C2 a, C2 b) { return {vmul(a.x, b.x), vmul(a.y, b.y)}; }  // elementwise (real) product
__device__ __forceinline__ C2 c2_add(C2 a, C2 b) { return {vadd(a.x, b.x), vadd(a.y, b.y)}; }

// s = acc + D(k . D x) along the tile axis, epilogues as ks_deriv2 (sweeps.cuh); float fields.
template <int N, int EPI, class RX, class RK, class RA, class RO>
__global__ void __launch_bounds__(N, v2_ctas<N>())
ks2_deriv2_pipe(int ntiles, RX x, RK kf, RA acc, RO out1, RO out2, const cplx<float>* __restrict__ twt, float alpha,
                double* partial, const int* __restrict__ done) {
  using F = LineFft2<N>;
  constexpr int E = F::E;
  constexpr bool KEEP_X = (EPI == EPI_MATVEC || EPI == EPI_RHS);
  if (done && *done) return;
  GLIA_DYN_SMEM(smraw);
  cplx<float>* stage0 = reinterpret_cast<cplx<float>*>(smraw);          // 2 tiles of N x 16 cplx<float>
  C2* sm = reinterpret_cast<C2*>(smraw + 2 * (size_t)N * 128);            // exchange tile
  TwDup* twd = reinterpret_cast<TwDup*>(smraw + 3 * (size_t)N * 128);     // duplicated twiddles
  const int l = threadIdx.x & (SL2 - 1), t = threadIdx.x / SL2;
  F::fill_table(twd, twt, threadIdx.x, N);
  typename F::Tw tw;
  F::init(tw, twd, t);
  AmS2 am{l};
  SyncCta sy;
  double dsum[1] = {0.0};
  const V2 al = vdup(alpha);

  int tile = blockIdx.x, s = 0;
  if (tile < ntiles) tile_prefetch<float, N, N>(stage0, x, tile);
  cp_async_commit();
  __syncthreads();  // twiddle table visible
  for (; tile < ntiles; tile += gridDim.x, s ^= 1) {
    const C2* st = reinterpret_cast<const C2*>(stage0 + (size_t)s * N * SL);
    const int next = tile + gridDim.x;
    if (next < ntiles) tile_prefetch<float, N, N>(stage0 + (size_t)(s ^ 1) * N * SL, x, next);
    cp_async_commit();
    const long kb = kf.tile_base(tile);
    C2 v[E], kk[E];
    GLIA_UNROLL
    for (int e = 0; e < E; ++e)
      kk[e] = reinterpret_cast<const C2*>(kf.row(kb, F::template loc<0>(t, e / F::R(0), e % F::R(0))))[l];
    cp_async_wait<1>();
    __syncthreads();
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) v[e] = st[am(F::template loc<0>(t, e / F::R(0), e % F::R(0)))];
    deriv_inplace2<N>(v, tw, sm, am, sy, t);
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) v[e] = c2_mul(v[e], kk[e]);
    C2 ac[E];
    if (EPI != EPI_SET) {
      const long ab = acc.tile_base(tile);
      GLIA_UNROLL
      for (int e = 0; e < E; ++e)
        ac[e] = reinterpret_cast<const C2*>(acc.row(ab, F::template loc<0>(t, e / F::R(0), e % F::R(0))))[l];
    }
    deriv_inplace2<N>(v, tw, sm, am, sy, t);
    const long ob = out1.tile_base(tile);
    if (EPI == EPI_AXPY) {  // out1 += alpha * (acc + D..)
      GLIA_UNROLL
      for (int e = 0; e < E; ++e) {
        const C2 o = reinterpret_cast<const C2*>(out1.row(ob, F::template loc<0>(t, e / F::R(0), e % F::R(0))))[l];
        const C2 sv = c2_add(v[e], ac[e]);
        v[e] = {vfma(al, sv.x, o.x), vfma(al, sv.y, o.y)};
      }
      GLIA_UNROLL
      for (int e = 0; e < E; ++e)
        reinterpret_cast<C2*>(out1.row(ob, F::template loc<0>(t, e / F::R(0), e % F::R(0))))[l] = v[e];
    } else {
      GLIA_UNROLL
      for (int e = 0; e < E; ++e) {
        const int lc = F::template loc<0>(t, e / F::R(0), e % F::R(0));
        C2 sv = v[e];
        if (EPI != EPI_SET) sv = c2_add(sv, ac[e]);
        C2* o1 = reinterpret_cast<C2*>(out1.row(ob, lc)) + l;
        if (EPI == EPI_SET || EPI == EPI_ADD || EPI == EPI_PLAIN) {
          *o1 = sv;
        } else if (EPI == EPI_MATVEC) {
          const C2 xv = st[am(lc)];
          const C2 w = {vfma(al, sv.x, xv.x), vfma(al, sv.y, xv.y)};
          *o1 = w;
          // <x, w>: products and the 4-term sum in FP32 (FFMA2), accumulated in FP64
          const V2 pr = vfma(xv.y, w.y, vmul(xv.x, w.x));
          dsum[0] += (double)pr.a + (double)pr.b;
        } else if (EPI == EPI_RHS) {
          const C2 xv = st[am(lc)];
          const C2 ds = {vmul(al, sv.x), vmul(al, sv.y)};
          const C2 b = {vadd(xv.x, ds.x), vadd(xv.y, ds.y)};
          const C2 ax = {vsub(xv.x, ds.x), vsub(xv.y, ds.y)};
          *o1 = b;
          reinterpret_cast<C2*>(out2.row(ob, lc))[l] = {vsub(b.x, ax.x), vsub(b.y, ax.y)};
        }
      }
    }
    if (KEEP_X) __syncthreads();  // the staged tile is reused by the prefetch of the next round
  }
  cp_async_wait<0>();
  if (EPI == EPI_MATVEC) block_reduce_store<1>(dsum, partial);
}

}  // namespace glia
