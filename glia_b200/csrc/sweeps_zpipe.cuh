// sweeps_zpipe.cuh -- persistent, warp-private, LDGSTS-pipelined form of the z second-derivative
// sweep; the default wherever a line pair fits one warp (kz_deriv2 of sweeps.cuh serves the rest).
//
// kz_deriv2 puts the 64 scalar loads of a line pair's x and k in flight at once; at 512-point lines
// that costs 168 registers = ONE 256-thread CTA per SM.  Here the x lines of the NEXT line-pair group
// ride into shared memory on cp.async while the current group is transformed, k (and the accumulator)
// follow through the same stage during the transforms, and nothing but the 16 complex values of the
// transform lives in registers across it.  A line pair belongs to N/E <= 32 threads of one warp, which fetch exactly
// the bytes they later read, so the whole pipeline needs __syncwarp only: no CTA barrier, warps drift.
// Measured on the B200 (profiles/r2d_zpipe_twldg_ab.txt): 49.7 -> 45.8 us at 256^3, 603 -> 526 us at
// 512^3 (f32).
//
//   smem per line pair = stage[0] | stage[1] (2N reals each) | exchange (zpad complex)
#pragma once
#include "sweeps_pipe.cuh"

namespace glia {

template <typename T, int N>
__host__ __device__ constexpr size_t zpipe_pair_bytes() {
  return 2 * (size_t)(2 * N) * sizeof(T) + (size_t)zpad<N>() * sizeof(cplx<T>);
}
template <typename T, int N>
__host__ __device__ constexpr size_t zpipe_smem() { return zlines<N>() * zpipe_pair_bytes<T, N>(); }
template <typename T, int N>
__host__ __device__ constexpr bool zpipe_fits() {
  return (N / FftPlan<N>::E) <= 32 && zpipe_smem<T, N>() <= 200 * 1024;
}
template <typename T, int N>
__host__ __device__ constexpr int zpipe_ctas() {
  const int c = (int)((224 * 1024) / zpipe_smem<T, N>());
  return c < 1 ? 1 : (c > 2 ? 2 : c);
}

// acc (+)= D_z(k D_z x) over `ngroups` groups of zlines<N>() line pairs; CTA b takes groups b, b + gridDim.x, ...
template <typename T, int N, int ADD>
__global__ void __launch_bounds__(zthreads<N>(), zpipe_ctas<T, N>())
kz_deriv2_pipe(LinesZ ln, int ngroups, const T* __restrict__ x, const T* __restrict__ kf, T* acc,
               const cplx<T>* __restrict__ twt, const int* __restrict__ done, int cpm) {
  // ngroups = line-pair groups of ONE ensemble member, cpm = CTAs per member (fft_core.cuh: ensemble batching)
  const int member = blockIdx.x / cpm, cl = blockIdx.x % cpm;
  done = member_done(done, member);
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E, TPL = F::TPL, LPC = zlines<N>();
  static_assert(TPL <= 32, "a line pair must live inside one warp");
  GLIA_DYN_SMEM(smraw);
  const int t = threadIdx.x % TPL, lp = threadIdx.x / TPL;
  // this pair's private shared memory: two stages of 2N reals, then the exchange region
  unsigned char* mine = smraw + (size_t)lp * zpipe_pair_bytes<T, N>();
  T* stage0 = reinterpret_cast<T*>(mine);
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(mine + 2 * (size_t)(2 * N) * sizeof(T));
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  SyncWarp sy;
  const AmZ am{0};  // `sm` already points at this pair's region
  constexpr int CH = (2 * N * (int)sizeof(T)) / 16;  // 16-byte chunks of one pair's two lines
  static_assert(CH % TPL == 0, "chunks per thread");

  const long pair0 = (long)member * ln.npairs;  // ln.npairs = line pairs of one member
  auto pair_of = [&](int group) -> long {
    long p = (long)group * LPC + lp;
    return pair0 + (p < ln.npairs ? p : ln.npairs - 1);  // ragged tail: re-do the last pair, stores predicated
  };
  auto prefetch = [&](T* stage, const T* field, int group) {
    const char* src = reinterpret_cast<const char*>(field + pair_of(group) * 2 * N);
    GLIA_UNROLL
    for (int i = 0; i < CH / TPL; ++i) {
      const int c = t + i * TPL;
      cp_async16(reinterpret_cast<char*>(stage) + 16 * c, src + 16 * c);
    }
  };

  int group = cl, s = 0;
  if (group < ngroups) prefetch(stage0, x, group);
  cp_async_commit();
  for (; group < ngroups; group += cpm, s ^= 1) {
    T* st = stage0 + (size_t)s * (2 * N);
    const int next = group + cpm;
    if (next < ngroups) prefetch(stage0 + (size_t)(s ^ 1) * (2 * N), x, next);
    cp_async_commit();
    const long pair = pair_of(group);
    const bool active = (long)group * LPC + lp < ln.npairs;
    const long la = pair * 2 * N, lb = la + N;
    cp_async_wait<1>();
    sy();  // the copies of the other lanes of this pair are visible too
    cplx<T> v[E];
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(t, e / F::R(0), e % F::R(0));
      v[e] = {st[pos], st[N + pos]};
    }
    // the stage of this round now carries, in turn, k (in flight during the first derivative) and -- ADD form --
    // the accumulator (in flight during the second): nothing but the transform's 16 complex values lives in
    // registers across it (same scheme as the S sweeps, sweeps_pipe.cuh)
    sy();
    prefetch(st, kf, group);
    cp_async_commit();
    deriv_inplace<T, N, zplan<N>()>(v, tw, sm, am, sy, t);
    cp_async_wait<0>();
    sy();
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(t, e / F::R(0), e % F::R(0));
      v[e].x *= st[pos];
      v[e].y *= st[N + pos];
    }
    if (ADD) {
      sy();
      prefetch(st, (const T*)acc, group);
    }
    cp_async_commit();
    deriv_inplace<T, N, zplan<N>()>(v, tw, sm, am, sy, t);
    cp_async_wait<0>();
    sy();
    if (active) {
      GLIA_UNROLL
      for (int e = 0; e < E; ++e) {
        const int pos = F::template loc<0>(t, e / F::R(0), e % F::R(0));
        cplx<T> o = v[e];
        if (ADD) { o.x += st[pos]; o.y += st[N + pos]; }
        acc[la + pos] = o.x;
        acc[lb + pos] = o.y;
      }
    }
    sy();  // the next round refills this stage two rounds later; its other stage is refilled at the top
  }
  cp_async_wait<0>();
}

}  // namespace glia
