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

// ---------------------------------------------------------------------------------------------------------------
// kz_c2r_pipe: the preconditioner's last sweep (packed half spectra -> two real z lines, z = M^-1 r, partial sums
// {<z,z>, <r,z>}) in the same persistent, warp-private form.  The one-group-per-CTA kz_c2r ran at 35 % issue
// activity and 47 % of the DRAM rate with 36 % of the warp slots filled (ncu, profiles/r2w_ncu_full_summary.csv):
// every CTA starts with the full latency of its spectrum loads.  Here the NEXT group's spectrum lines ride into the
// pair's stage as soon as the current group has been unpacked out of it, and the next group's r lines as soon as
// the epilogue has read the current ones; both buffers are single:
//   smem per line pair = spectrum stage (N complex) | r stage (2N reals) | exchange (zpad complex)
//   cp.async groups in flight, in commit order: [spectrum(g), r(g)] at the top of group g, [r(g), spectrum(g+)] after
//   the unpack -- wait_group<1> serves both waits.
template <typename T, int N>
__host__ __device__ constexpr size_t c2rpipe_pair_bytes() {
  return (size_t)N * sizeof(cplx<T>) + (size_t)(2 * N) * sizeof(T) + (size_t)zpad<N>() * sizeof(cplx<T>);
}
template <typename T, int N>
__host__ __device__ constexpr size_t c2rpipe_smem() { return zlines<N>() * c2rpipe_pair_bytes<T, N>(); }
template <typename T, int N>
__host__ __device__ constexpr bool c2rpipe_fits() {
  return (N / FftPlan<N>::E) <= 32 && c2rpipe_smem<T, N>() <= 110 * 1024;  // two CTAs per SM
}
//   EPI 2: <z,z>, <r,z> (z and r given); EPI 1: <z,z> only (r unused; zout may be null: rnorm0 = ||M^-1 b||)
template <typename T, int N, int EPI>
__global__ void __launch_bounds__(zthreads<N>(), 2)
kz_c2r_pipe(LinesZ ln, int ngroups, const cplx<T>* __restrict__ shat, T* zout, const T* __restrict__ r, double* partial,
            const cplx<T>* __restrict__ twt, const int* __restrict__ done, int cpm) {
  // ln.npairs, ngroups: line pairs / groups of zlines<N>() pairs of ONE ensemble member; cpm = CTAs per member
  const int member = blockIdx.x / cpm, cl = blockIdx.x % cpm;
  done = member_done(done, member);
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E, TPL = F::TPL, LPC = zlines<N>();
  static_assert(TPL <= 32, "a line pair must live inside one warp");
  GLIA_DYN_SMEM(smraw);
  const int t = threadIdx.x % TPL, lp = threadIdx.x / TPL;
  unsigned char* mine = smraw + (size_t)lp * c2rpipe_pair_bytes<T, N>();
  cplx<T>* sst = reinterpret_cast<cplx<T>*>(mine);                             // [0, N/2): line a, [N/2, N): line b
  T* rst = reinterpret_cast<T*>(mine + (size_t)N * sizeof(cplx<T>));           // [0, N): line a, [N, 2N): line b
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(mine + (size_t)N * sizeof(cplx<T>) + (size_t)(2 * N) * sizeof(T));
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  SyncWarp sy;
  const AmZ am{0};
  constexpr int CH = (N * (int)sizeof(cplx<T>)) / 16;  // 16-byte chunks of one pair's spectra = of its two r lines
  static_assert(CH % TPL == 0, "chunks per thread");
  const long pair0 = (long)member * ln.npairs;
  auto pair_of = [&](int group) -> long {
    long p = (long)group * LPC + lp;
    return pair0 + (p < ln.npairs ? p : ln.npairs - 1);  // ragged tail: re-do the last pair, stores predicated
  };
  auto fetch = [&](void* stage, const void* line0) {
    GLIA_UNROLL
    for (int i = 0; i < CH / TPL; ++i) {
      const int c = t + i * TPL;
      cp_async16(reinterpret_cast<char*>(stage) + 16 * c, reinterpret_cast<const char*>(line0) + 16 * c);
    }
  };
  double acc[2] = {0.0, 0.0};
  int group = cl;
  if (group < ngroups) fetch(sst, shat + pair_of(group) * N);
  cp_async_commit();
  if (EPI == 2 && group < ngroups) fetch(rst, r + pair_of(group) * 2 * N);
  cp_async_commit();
  for (; group < ngroups; group += cpm) {
    const long pair = pair_of(group);
    const bool active = (long)group * LPC + lp < ln.npairs;
    const long la = pair * 2 * N, lb = la + N;
    cp_async_wait<1>();  // this group's spectra (its r lines may still be in flight)
    sy();
    // untangle the two Hermitian half spectra into one complex line, placed where the inverse's first pass reads
    GLIA_UNROLL
    for (int j = 0; j < E / 2; ++j) {
      const int k = t + TPL * j;
      const cplx<T> A = sst[k], B = sst[N / 2 + k];
      if (k == 0) {
        sm[am(F::loc_of_freq(0))] = {A.x, B.x};
        sm[am(F::loc_of_freq(N / 2))] = {A.y, B.y};
      } else {
        sm[am(F::loc_of_freq(k))] = {A.x - B.y, A.y + B.x};
        sm[am(F::loc_of_freq(N - k))] = {A.x + B.y, B.x - A.y};
      }
    }
    sy();  // the pair's spectrum stage has been read by all of its lanes: refill it with the next group's
    const int next = group + cpm;
    if (next < ngroups) fetch(sst, shat + pair_of(next) * N);
    cp_async_commit();
    cplx<T> v[E];
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(F::P - 1); ++g)
      GLIA_UNROLL
      for (int cc = 0; cc < F::RL; ++cc) v[g * F::RL + cc] = sm[am(F::template loc<F::P - 1>(t, g, cc))];
    F::inverse(v, tw, sm, am, sy, t);
    cp_async_wait<1>();  // this group's r lines (the next group's spectra stay in flight)
    sy();
    [[maybe_unused]] T fa[4] = {(T)0, (T)0, (T)0, (T)0};
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(0); ++g)
      GLIA_UNROLL
      for (int a = 0; a < F::R(0); ++a) {
        const int pos = F::template loc<0>(t, g, a);
        const cplx<T> zv = v[g * F::R(0) + a];
        if (active) {
          if (EPI == 2 || zout) {
            zout[la + pos] = zv.x;
            zout[lb + pos] = zv.y;
          }
          if constexpr (sizeof(T) == 8) {
            acc[0] += (double)zv.x * (double)zv.x + (double)zv.y * (double)zv.y;
            if constexpr (EPI == 2) acc[1] += (double)rst[pos] * (double)zv.x + (double)rst[N + pos] * (double)zv.y;
          } else {  // per-thread float partials (see kz_c2r, sweeps.cuh)
            fa[0] = zv.x * zv.x + fa[0];
            fa[1] = zv.y * zv.y + fa[1];
            if constexpr (EPI == 2) {
              fa[2] = rst[pos] * zv.x + fa[2];
              fa[3] = rst[N + pos] * zv.y + fa[3];
            }
          }
        }
      }
    if constexpr (sizeof(T) != 8) {
      acc[0] += (double)fa[0] + (double)fa[1];
      acc[1] += (double)fa[2] + (double)fa[3];
    }
    sy();  // the pair's r stage has been read: the next group's lines may land
    if (EPI == 2 && next < ngroups) fetch(rst, r + pair_of(next) * 2 * N);
    cp_async_commit();
  }
  cp_async_wait<0>();
  block_reduce_store<2>(acc, partial);
}

// ---------------------------------------------------------------------------------------------------------------
// kz_r2c_pipe: the preconditioner's first sweep (two real z lines -> two packed half spectra, optional PCG prologue
// r <- r - a w with r written back) in the same form: the pair's r (and w) lines are read out of their stage into
// registers at the top of a group and the next group's lines ride in behind them while this group is transformed.
//   smem per line pair = r stage (2N reals) | w stage (2N reals) | exchange (zpad complex)
template <typename T, int N>
__host__ __device__ constexpr size_t r2cpipe_pair_bytes() {
  return 2 * (size_t)(2 * N) * sizeof(T) + (size_t)zpad<N>() * sizeof(cplx<T>);
}
template <typename T, int N>
__host__ __device__ constexpr size_t r2cpipe_smem() { return zlines<N>() * r2cpipe_pair_bytes<T, N>(); }
template <typename T, int N>
__host__ __device__ constexpr bool r2cpipe_fits() {
  return (N / FftPlan<N>::E) <= 32 && r2cpipe_smem<T, N>() <= 110 * 1024;  // two CTAs per SM
}
template <typename T, int N, int PRO>
__global__ void __launch_bounds__(zthreads<N>(), 2)
kz_r2c_pipe(LinesZ ln, int ngroups, T* r, const T* __restrict__ w, const double* __restrict__ scal_a, cplx<T>* shat,
            const cplx<T>* __restrict__ twt, const int* __restrict__ done, int cpm) {
  const int member = blockIdx.x / cpm, cl = blockIdx.x % cpm;
  done = member_done(done, member);
  GLIA_PDL_ENTRY_EARLY(done);
  using F = LineFft<T, N, zplan<N>()>;
  constexpr int E = F::E, TPL = F::TPL, LPC = zlines<N>();
  static_assert(TPL <= 32, "a line pair must live inside one warp");
  GLIA_DYN_SMEM(smraw);
  const int t = threadIdx.x % TPL, lp = threadIdx.x / TPL;
  unsigned char* mine = smraw + (size_t)lp * r2cpipe_pair_bytes<T, N>();
  T* rst = reinterpret_cast<T*>(mine);
  T* wst = rst + 2 * N;
  cplx<T>* sm = reinterpret_cast<cplx<T>*>(mine + 2 * (size_t)(2 * N) * sizeof(T));
  typename F::Tw tw;
  F::load_twiddles(tw, twt, t);
  GLIA_PDL_ENTRY_LATE(done);  // everything above is independent of earlier kernels
  SyncWarp sy;
  const AmZ am{0};
  constexpr int CH = (2 * N * (int)sizeof(T)) / 16;  // 16-byte chunks of one pair's two lines
  static_assert(CH % TPL == 0, "chunks per thread");
  const long pair0 = (long)member * ln.npairs;
  auto pair_of = [&](int group) -> long {
    long p = (long)group * LPC + lp;
    return pair0 + (p < ln.npairs ? p : ln.npairs - 1);
  };
  auto fetch = [&](void* stage, const void* line0) {
    GLIA_UNROLL
    for (int i = 0; i < CH / TPL; ++i) {
      const int c = t + i * TPL;
      cp_async16(reinterpret_cast<char*>(stage) + 16 * c, reinterpret_cast<const char*>(line0) + 16 * c);
    }
  };
  T aa = (T)0;
  if (PRO) aa = (T)(scal_a[(size_t)member * SCAL_STRIDE]);
  int group = cl;
  if (group < ngroups) {
    fetch(rst, r + pair_of(group) * 2 * N);
    if (PRO) fetch(wst, w + pair_of(group) * 2 * N);
  }
  cp_async_commit();
  for (; group < ngroups; group += cpm) {
    const long pair = pair_of(group);
    const bool active = (long)group * LPC + lp < ln.npairs;
    const long la = pair * 2 * N, lb = la + N;
    cp_async_wait<0>();
    sy();
    cplx<T> v[E];
    GLIA_UNROLL
    for (int e = 0; e < E; ++e) {
      const int pos = F::template loc<0>(t, e / F::R(0), e % F::R(0));
      cplx<T> rv = {rst[pos], rst[N + pos]};
      if (PRO) {
        rv.x = rv.x - aa * wst[pos];
        rv.y = rv.y - aa * wst[N + pos];
      }
      v[e] = rv;
    }
    sy();  // the stages have been read by every lane of the pair: the next group's lines may land
    const int next = group + cpm;
    if (next < ngroups) {
      fetch(rst, r + pair_of(next) * 2 * N);
      if (PRO) fetch(wst, w + pair_of(next) * 2 * N);
    }
    cp_async_commit();
    if (PRO && active) {
      GLIA_UNROLL
      for (int e = 0; e < E; ++e) {
        const int pos = F::template loc<0>(t, e / F::R(0), e % F::R(0));
        r[la + pos] = v[e].x;
        r[lb + pos] = v[e].y;
      }
    }
    F::forward(v, tw, sm, am, sy, t);
    // scatter by frequency, then untangle the two real spectra
    sy();
    GLIA_UNROLL
    for (int g = 0; g < F::Gp(F::P - 1); ++g) {
      const int kb = F::kbase(t, g);
      GLIA_UNROLL
      for (int cc = 0; cc < F::RL; ++cc) {
        const int k = kb + F::KSTEP * cc;
        sm[k + (k >> 4)] = v[g * F::RL + cc];
      }
    }
    sy();
    const long oa = pair * 2 * (N / 2), ob = oa + N / 2;
    GLIA_UNROLL
    for (int j = 0; j < E / 2; ++j) {
      const int k = t + TPL * j;
      const int kn = (N - k) & (N - 1);
      const cplx<T> zk = sm[k + (k >> 4)], zn = sm[kn + (kn >> 4)];
      cplx<T> A = {(T)0.5 * (zk.x + zn.x), (T)0.5 * (zk.y - zn.y)};
      cplx<T> B = {(T)0.5 * (zk.y + zn.y), (T)-0.5 * (zk.x - zn.x)};
      if (k == 0) {
        const cplx<T> zh = sm[N / 2 + ((N / 2) >> 4)];
        A = {zk.x, zh.x};
        B = {zk.y, zh.y};
      }
      if (active) { shat[oa + k] = A; shat[ob + k] = B; }
    }
    sy();  // the exchange region is rewritten by the next group's forward transform
  }
  cp_async_wait<0>();
}

}  // namespace glia
