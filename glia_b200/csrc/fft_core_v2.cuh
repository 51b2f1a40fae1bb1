// fft_core_v2.cuh -- the line FFT on Blackwell's packed FP32x2 pipe (FADD2 / FMUL2 / FFMA2,
// PTX add/mul/fma.rn.f32x2, new in sm_100).
//
// ncu on the scalar sweeps (profiles/r1c_*): 70-83 % of the executed warp instructions are
// FADD / FMUL / FFMA of the butterflies, the kernels issue 2.0-2.3 of 4 instructions per cycle
// per SM while DRAM sits at 40-45 % -- they are issue bound, not memory bound.  The packed
// instructions do two FP32 operations per issue slot, so here every thread carries TWO
// independent complex sequences A and B in lock step, stored structure-of-arrays in 64-bit
// register pairs:   re = (re_A, re_B),  im = (im_A, im_B).
// All butterfly and twiddle arithmetic is then one packed instruction per two scalar ones, with
// no lane shuffles (a twiddle multiplies both sequences by the same scalar).
//
// Which two sequences ride together is free wherever the sweep applies a REAL linear operator per
// real line (the derivative sweeps on the pair view): a 16-byte load of four adjacent real z
// values (z0 z1 z2 z3) is read as A = z0 + i z2, B = z1 + i z3, i.e. re = (z0, z1) and
// im = (z2, z3) are already aligned register pairs.
//
// Plan: E = 8 points per thread and sequence (so that x, k and acc of both sequences fit the
// 128-register budget), radices (8,8,4) for N = 256, TPL = N / 8 threads per line.
#pragma once
#include "fft_core.cuh"

namespace glia {

struct alignas(8) V2 {
  float a, b;
};
using C2 = cplx<V2>;  // 16 bytes: re.a re.b im.a im.b

#if defined(GLIA_SIMT_EMU)
__device__ inline V2 vadd(V2 x, V2 y) { return {x.a + y.a, x.b + y.b}; }
__device__ inline V2 vsub(V2 x, V2 y) { return {x.a - y.a, x.b - y.b}; }
__device__ inline V2 vmul(V2 x, V2 y) { return {x.a * y.a, x.b * y.b}; }
__device__ inline V2 vfma(V2 x, V2 y, V2 z) { return {std::fma(x.a, y.a, z.a), std::fma(x.b, y.b, z.b)}; }
#else
__device__ __forceinline__ unsigned long long v2_bits(V2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ V2 v2_from(unsigned long long u) { return *reinterpret_cast<V2*>(&u); }
__device__ __forceinline__ V2 vadd(V2 x, V2 y) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(v2_bits(x)), "l"(v2_bits(y)));
  return v2_from(r);
}
__device__ __forceinline__ V2 vsub(V2 x, V2 y) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(v2_bits(x)), "l"(v2_bits(y)));
  return v2_from(r);
}
__device__ __forceinline__ V2 vmul(V2 x, V2 y) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(v2_bits(x)), "l"(v2_bits(y)));
  return v2_from(r);
}
__device__ __forceinline__ V2 vfma(V2 x, V2 y, V2 z) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(v2_bits(x)), "l"(v2_bits(y)), "l"(v2_bits(z)));
  return v2_from(r);
}
#endif
__host__ __device__ __forceinline__ constexpr V2 vdup(float s) { return V2{s, s}; }

// twiddle table entry, duplicated for both sequences: exp(-2 pi i j / N) = c + i s
struct alignas(32) TwDup {
  V2 c, s, ns, pad;
};
// v * w
__device__ __forceinline__ C2 cmul_tw(C2 v, const TwDup& w) {
  return {vfma(v.y, w.ns, vmul(v.x, w.c)), vfma(v.y, w.c, vmul(v.x, w.s))};
}
// v * conj(w)
__device__ __forceinline__ C2 cmulc_tw(C2 v, const TwDup& w) {
  return {vfma(v.y, w.s, vmul(v.x, w.c)), vfma(v.x, w.ns, vmul(v.y, w.c))};
}

// ---- in-register radix-R DIF on packed pairs ------------------------------------------------
// v[J] = a + b ; v[J+H] = (a - b) * exp(SIGN 2 pi i J / R).  Multiplications by +-i are operand
// swaps of the subtraction (no negation instruction exists for the packed type).
template <int R, int SIGN, int J>
__device__ __forceinline__ void dif_bf2(C2* v) {
  constexpr int H = R / 2;
  constexpr int j32 = (J % R) * (32 / R);
  const C2 a = v[J], b = v[J + H];
  v[J] = {vadd(a.x, b.x), vadd(a.y, b.y)};
  if constexpr (j32 == 0) {
    v[J + H] = {vsub(a.x, b.x), vsub(a.y, b.y)};
  } else if constexpr (j32 == 8) {  // (tx, ty) * (SIGN i) = SIGN (-ty, tx)
    if constexpr (SIGN > 0) v[J + H] = {vsub(b.y, a.y), vsub(a.x, b.x)};
    else v[J + H] = {vsub(a.y, b.y), vsub(b.x, a.x)};
  } else {
    static_assert(j32 < 16, "DIF twiddles stay in the first half turn");
    constexpr float c = (float)cos32(j32);
    constexpr float s = (float)(SIGN * sin32(j32));
    const C2 d = {vsub(a.x, b.x), vsub(a.y, b.y)};
    v[J + H] = {vfma(d.y, vdup(-s), vmul(d.x, vdup(c))), vfma(d.y, vdup(c), vmul(d.x, vdup(s)))};
  }
}
template <int R, int SIGN, int... J>
__device__ __forceinline__ void dif_level2(C2* v, std::integer_sequence<int, J...>) {
  (dif_bf2<R, SIGN, J>(v), ...);
}
template <int R, int SIGN>
struct Dif2 {
  static __device__ __forceinline__ void run(C2* v) {
    dif_level2<R, SIGN>(v, std::make_integer_sequence<int, R / 2>{});
    Dif2<R / 2, SIGN>::run(v);
    Dif2<R / 2, SIGN>::run(v + R / 2);
  }
};
template <int SIGN>
struct Dif2<1, SIGN> {
  static __device__ __forceinline__ void run(C2*) {}
};
template <int R, int SIGN>
__device__ __forceinline__ void dft_reg2(C2* v) {
  Dif2<R, SIGN>::run(v);
  brev_all<V2, R>(v, std::make_integer_sequence<int, R>{});
}

// ---- plans --------------------------------------------------------------------------------
template <int N> struct FftPlanV;
template <> struct FftPlanV<32>  { static constexpr int E = 8, P = 2, R0 = 8, R1 = 4, R2 = 1; };
template <> struct FftPlanV<64>  { static constexpr int E = 8, P = 2, R0 = 8, R1 = 8, R2 = 1; };
template <> struct FftPlanV<128> { static constexpr int E = 8, P = 3, R0 = 8, R1 = 4, R2 = 4; };
template <> struct FftPlanV<256> { static constexpr int E = 8, P = 3, R0 = 8, R1 = 8, R2 = 4; };
template <> struct FftPlanV<512> { static constexpr int E = 8, P = 3, R0 = 8, R1 = 8, R2 = 8; };

template <int N>
struct LineFft2 {
  using PL = FftPlanV<N>;
  static constexpr int E = PL::E, P = PL::P, TPL = N / E;
  __host__ __device__ static constexpr int R(int p) { return p == 0 ? PL::R0 : (p == 1 ? PL::R1 : PL::R2); }
  __host__ __device__ static constexpr int Np(int p) {
    int n = N;
    for (int i = 0; i < p; ++i) n /= R(i);
    return n;
  }
  __host__ __device__ static constexpr int Mp(int p) { return Np(p) / R(p); }
  __host__ __device__ static constexpr int Gp(int p) { return E / R(p); }
  static constexpr int RL = R(P - 1);
  static constexpr int KSTEP = N / RL;
  // (pass, group) slots that carry inter-pass twiddles
  __host__ __device__ static constexpr int slot(int p, int g) {
    int n = 0;
    for (int i = 0; i < p; ++i) n += Gp(i);
    return n + g;
  }
  static constexpr int NSLOT = slot(P - 1, 0) > 0 ? slot(P - 1, 0) : 1;
  struct Tw {
    const TwDup* tab;  // shared memory, N entries
    int bq[NSLOT];     // (q mod M_p) * (N / N_p) per (pass, group)
  };
  __device__ static __forceinline__ void init(Tw& tw, const TwDup* tab, int t) {
    tw.tab = tab;
    GLIA_UNROLL
    for (int p = 0; p + 1 < P; ++p) {
      GLIA_UNROLL
      for (int g = 0; g < Gp(p); ++g) tw.bq[slot(p, g)] = ((t + TPL * g) % Mp(p)) * (N / Np(p));
    }
  }
  // all threads of the CTA: expand the per-axis table exp(-2 pi i j / N) into shared memory
  __device__ static __forceinline__ void fill_table(TwDup* tab, const cplx<float>* __restrict__ table, int tid, int nthr) {
    for (int j = tid; j < N; j += nthr) {
      const cplx<float> w = table[j];
      tab[j] = TwDup{vdup(w.x), vdup(w.y), vdup(-w.y), vdup(0.f)};
    }
  }

  template <int p>
  __device__ static __forceinline__ int loc(int t, int g, int a) {
    constexpr int M = Mp(p), NP = Np(p);
    const int q = t + TPL * g;
    return (q / M) * NP + a * M + (q % M);
  }
  __device__ static __forceinline__ int kbase(int t, int g) {
    const int s = t + TPL * g;
    if constexpr (P == 2) return s;
    else return s / PL::R1 + PL::R0 * (s % PL::R1);
  }
  __device__ static __forceinline__ int loc_of_freq(int k) {
    if constexpr (P == 2) return (k % PL::R0) * PL::R1 + k / PL::R0;
    else return (k % PL::R0) * (PL::R1 * PL::R2) + ((k / PL::R0) % PL::R1) * PL::R2 + k / (PL::R0 * PL::R1);
  }

  template <int p, int SIGN>
  __device__ static __forceinline__ void butterflies(C2 (&v)[E]) {
    GLIA_UNROLL
    for (int g = 0; g < Gp(p); ++g) dft_reg2<R(p), SIGN>(&v[g * R(p)]);
  }
  template <int p, bool CONJ>
  __device__ static __forceinline__ void twiddle(C2 (&v)[E], const Tw& tw) {
    if constexpr (p + 1 < P) {
      GLIA_UNROLL
      for (int g = 0; g < Gp(p); ++g) {
        GLIA_UNROLL
        for (int c = 1; c < R(p); ++c) {
          const TwDup w = tw.tab[(tw.bq[slot(p, g)] * c) & (N - 1)];
          v[g * R(p) + c] = CONJ ? cmulc_tw(v[g * R(p) + c], w) : cmul_tw(v[g * R(p) + c], w);
        }
      }
    }
  }
  template <int pw, int pr, class AM, class SY>
  __device__ static __forceinline__ void exchange(C2 (&v)[E], C2* sm, AM am, SY sync, int t) {
    sync();
    GLIA_UNROLL
    for (int g = 0; g < Gp(pw); ++g) {
      GLIA_UNROLL
      for (int a = 0; a < R(pw); ++a) sm[am(loc<pw>(t, g, a))] = v[g * R(pw) + a];
    }
    sync();
    GLIA_UNROLL
    for (int g = 0; g < Gp(pr); ++g) {
      GLIA_UNROLL
      for (int a = 0; a < R(pr); ++a) v[g * R(pr) + a] = sm[am(loc<pr>(t, g, a))];
    }
  }
  template <class AM, class SY>
  __device__ static __forceinline__ void forward(C2 (&v)[E], const Tw& tw, C2* sm, AM am, SY sync, int t) {
    butterflies<0, -1>(v);
    twiddle<0, false>(v, tw);
    if constexpr (P >= 2) {
      exchange<0, 1>(v, sm, am, sync, t);
      butterflies<1, -1>(v);
      twiddle<1, false>(v, tw);
    }
    if constexpr (P >= 3) {
      exchange<1, 2>(v, sm, am, sync, t);
      butterflies<2, -1>(v);
    }
  }
  template <class AM, class SY>
  __device__ static __forceinline__ void inverse(C2 (&v)[E], const Tw& tw, C2* sm, AM am, SY sync, int t) {
    if constexpr (P >= 3) {
      butterflies<2, +1>(v);
      exchange<2, 1>(v, sm, am, sync, t);
    }
    if constexpr (P >= 2) {
      twiddle<1, true>(v, tw);
      butterflies<1, +1>(v);
      exchange<1, 0>(v, sm, am, sync, t);
    }
    twiddle<0, true>(v, tw);
    butterflies<0, +1>(v);
  }

  // v <- (i w(k) / N) v, Nyquist wavenumber zeroed (trap T1), both sequences at once:
  // (re, im) -> (-w im, w re) with w and -w built by packed adds from the thread's base frequency.
  __device__ static __forceinline__ void mult_iw(C2 (&v)[E], int t) {
    GLIA_UNROLL
    for (int g = 0; g < Gp(P - 1); ++g) {
      const int kb = kbase(t, g);
      const V2 kbn = vdup((float)kb * (float)(1.0 / N));
      GLIA_UNROLL
      for (int c = 0; c < RL; ++c) {
        constexpr double one = 1.0;
        const float off = (float)((c < RL / 2) ? (double)c / RL : (double)c / RL - one);
        V2 w = vadd(kbn, vdup(off));
        V2 nw = vsub(vdup(-off), kbn);
        if (c == RL / 2 && kb == 0) { w = vdup(0.f); nw = vdup(0.f); }
        const C2 z = v[g * RL + c];
        v[g * RL + c] = {vmul(nw, z.y), vmul(w, z.x)};
      }
    }
  }
};

// v <- D_axis(v) on both sequences
template <int N, class AM, class SY>
__device__ __forceinline__ void deriv_inplace2(C2 (&v)[8], const typename LineFft2<N>::Tw& tw, C2* sm, AM am, SY sy, int t) {
  using F = LineFft2<N>;
  F::forward(v, tw, sm, am, sy, t);
  F::mult_iw(v, t);
  F::inverse(v, tw, sm, am, sy, t);
}

}  // namespace glia
