// fft_core.cuh -- register-resident multi-pass line FFT for the sweep kernels.
//
// One line of N complex points is owned by TPL = N/E threads, E points per
// thread.  The transform is a decimation-in-frequency Cooley-Tukey over P passes
// of radix R_p; between passes the points move through shared memory, inside a
// pass everything stays in registers.  forward() leaves every thread holding
// frequencies in digit-reversed placement, inverse() takes exactly that
// placement back to natural positions, so `forward -> pointwise multiplier ->
// inverse` (the shape of every spectral operator on the RD path:
// src/grad/SpectralOperators.cpp:100-261, src/pde/DiffusionSolver.cpp:182-215
// in the reference) needs no reordering pass at all.
#pragma once
#include "simt.h"

namespace glia {

template <typename T>
struct alignas(2 * sizeof(T)) cplx {
  T x, y;
};

template <typename T>
__device__ __forceinline__ cplx<T> cadd(cplx<T> a, cplx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__device__ __forceinline__ cplx<T> csub(cplx<T> a, cplx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T>
__device__ __forceinline__ cplx<T> cmul(cplx<T> a, cplx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__device__ __forceinline__ cplx<T> cmulc(cplx<T> a, cplx<T> b) {  // a * conj(b)
  return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}

// ---------------------------------------------- packed FP32x2 arithmetic ----
// Blackwell (sm_100) issues two FP32 operations per instruction on a 64-bit register pair
// (add / mul / fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2).  A cplx<float> IS such a pair, so complex
// additions, subtractions and multiplications by a real scalar take one issue slot instead of two.
// ncu (profiles/r1c_*): the single-precision sweeps are issue bound -- 70-83 % of their executed
// instructions are the butterflies' FADD / FMUL / FFMA, 2.0-2.3 of 4 issue slots per cycle are
// used, DRAM sits at 40-45 % -- so instruction count is what buys time.
// MEASURED (profiles/r1d_f32x2_ab.txt): packing the butterflies' additions removed 17 % of the
// D-sweeps' instructions (992 FADD -> 456 FADD2 + 144 FADD per tile and thread) and changed no
// kernel's time by more than noise, the y sweep got 8 % slower: FADD2 holds the FMA pipe for two
// issue cycles, and the sweeps are bound by their load -> transform -> store phase structure at
// 16 resident warps per SM, not by issue slots.  Build option (-DGLIA_USE_F32X2), default off.
#if !defined(GLIA_SIMT_EMU) && defined(GLIA_USE_F32X2)
#define GLIA_F32X2 1
__device__ __forceinline__ unsigned long long c_bits(cplx<float> v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ cplx<float> c_from(unsigned long long u) { return *reinterpret_cast<cplx<float>*>(&u); }
__device__ __forceinline__ cplx<float> cadd(cplx<float> a, cplx<float> b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c_bits(a)), "l"(c_bits(b)));
  return c_from(r);
}
__device__ __forceinline__ cplx<float> csub(cplx<float> a, cplx<float> b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c_bits(a)), "l"(c_bits(b)));
  return c_from(r);
}
// a * (s, s)
__device__ __forceinline__ cplx<float> cscale(cplx<float> a, float s) {
  unsigned long long r;
  const cplx<float> ss = {s, s};
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c_bits(a)), "l"(c_bits(ss)));
  return c_from(r);
}
#else
#define GLIA_F32X2 0
#endif
template <typename T>
__device__ __forceinline__ cplx<T> cscale(cplx<T> a, T s) { return {a.x * s, a.y * s}; }

// ------------------------------------------------- L2 residency hints ----
// At 256^3 single precision one field is 67 MB and B200's L2 holds 126 MB: of the fields a PCG
// iteration touches, exactly one fits next to the streams.  With GLIA_L2_HINTS=1 the sweeps read
// every operand that is consumed once per kernel (x, k, r, w, the last read of acc / shat) with the
// evict-first policy (ld.global.cs), so that the field handed from one kernel to the next
// (acc -> w -> shat -> z) survives in L2 until its consumer runs.
// MEASURED (profiles/r1b_l2_probe.txt, profiles/r1c_l2_hints_ab.txt): a DRAM-bound 4F streaming
// kernel gains 25-30% from this (7.5 TB/s effective against 5.5), but the sweep kernels do not --
// they are issue / latency bound, not DRAM bound, at this size (65 us for 4F where DRAM needs 41) --
// and evict-first STORES of sub-sector pieces (the z sweeps write 64-byte runs) tripled kz_r2c's
// time.  Default OFF; kept as a build option for grids whose sweeps become DRAM bound.
#ifndef GLIA_L2_HINTS
#define GLIA_L2_HINTS 0
#endif
#if defined(GLIA_SIMT_EMU) || !GLIA_L2_HINTS
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) { return *p; }
template <typename V> __device__ __forceinline__ void st_stream(V* p, V v) { *p = v; }
#else
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ cplx<float> ld_stream(const cplx<float>* p) {
  const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
  return {v.x, v.y};
}
__device__ __forceinline__ cplx<double> ld_stream(const cplx<double>* p) {
  const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(cplx<float>* p, cplx<float> v) {
  __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
}
__device__ __forceinline__ void st_stream(cplx<double>* p, cplx<double> v) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
}
#endif

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may
// be scheduled while its predecessor's last CTAs are still draining; it does its set-up (indices,
// twiddle registers from the constant per-axis table) and then waits here until the predecessor grid
// has completed and its memory is visible.  Nothing produced by an earlier kernel may be touched
// before this call.
#ifndef GLIA_PDL_MODE
// 0: no PDL instructions (probe builds), 1: wait only, 2: explicit early trigger + wait.  Measured: the explicit
// griddepcontrol.launch_dependents on entry slows the multi-wave z sweeps (kz_deriv2 600 -> 629 us, kz_r2c.axpy
// 473 -> 530 us at 512^3) and gains nothing over the implicit trigger at CTA exit (90.0 vs 88.5 time-steps/s at
// 256^3), so the default is the wait alone.
#define GLIA_PDL_MODE 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if !defined(GLIA_SIMT_EMU) && GLIA_PDL_MODE >= 1
#if GLIA_PDL_MODE >= 2
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the NEXT kernel's CTAs fill our tail
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Where a sweep kernel waits for its predecessor and tests the PCG's `done` flag (written by an earlier
// kernel, so only readable after the wait).  Measured on the B200 (gpurun_out/r1l_*, profiles/r1l_pdl_matrix.txt):
// waiting AFTER the twiddle registers are loaded is the faster place for lines up to 256 points
// (kz_r2c.axpy 45 vs 51 us at 256^3), but at 512 points it costs the z sweeps up to 45 % (kz_r2c.axpy
// 473 vs 325 us at 512^3), so those wait first thing.  GLIA_PDL_TOP = 0 / 1 forces one place for probe builds.
#ifndef GLIA_PDL_TOP
#define GLIA_PDL_TOP (N >= 512)
#endif
#define GLIA_PDL_ENTRY_EARLY(done) \
  if constexpr (GLIA_PDL_TOP) { pdl_wait(); if ((done) && *(done)) return; }
#define GLIA_PDL_ENTRY_LATE(done) \
  if constexpr (!(GLIA_PDL_TOP)) { pdl_wait(); if ((done) && *(done)) return; }

// software prefetch of one 128-byte line into L1 (for operands an epilogue reads long after the
// kernel starts: zero registers held across the transform)
template <typename V>
__device__ __forceinline__ void prefetch_l1(const V* p) {
#if !defined(GLIA_SIMT_EMU)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

#ifndef GLIA_TW_LDG
#define GLIA_TW_LDG 0
#endif
// read-only load of a twiddle-table entry; volatile, because a plain __ldg is hoisted out of the transforms by
// ptxas and the values sit in registers again
template <typename T>
__device__ __forceinline__ cplx<T> ld_table(const cplx<T>* p) {
#if defined(GLIA_SIMT_EMU)
  return *p;
#else
  if constexpr (sizeof(T) == 4) {
    float x, y;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "l"(p));
    return {(T)x, (T)y};
  } else {
    double x, y;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
    return {(T)x, (T)y};
  }
#endif
}

// ---------------------------------------------------------------- plans ----
template <int N> struct FftPlan;
template <> struct FftPlan<32>  { static constexpr int E = 8,  P = 2, R0 = 8,  R1 = 4,  R2 = 1; };
template <> struct FftPlan<64>  { static constexpr int E = 8,  P = 2, R0 = 8,  R1 = 8,  R2 = 1; };
template <> struct FftPlan<128> { static constexpr int E = 16, P = 2, R0 = 16, R1 = 8,  R2 = 1; };
template <> struct FftPlan<256> { static constexpr int E = 16, P = 2, R0 = 16, R1 = 16, R2 = 1; };
template <> struct FftPlan<512> { static constexpr int E = 16, P = 3, R0 = 8,  R1 = 8,  R2 = 8; };

// cos(2 pi j / 32), j = 0..8 -- enough for every radix <= 32 by symmetry
__host__ __device__ constexpr double cos32(int j) {
  j = ((j % 32) + 32) % 32;
  if (j > 16) j = 32 - j;
  bool neg = false;
  if (j > 8) { j = 16 - j; neg = true; }
  double v = 0;
  switch (j) {
    case 0: v = 1.0; break;
    case 1: v = 0.98078528040323044913; break;
    case 2: v = 0.92387953251128675613; break;
    case 3: v = 0.83146961230254523708; break;
    case 4: v = 0.70710678118654752440; break;
    case 5: v = 0.55557023301960222474; break;
    case 6: v = 0.38268343236508977173; break;
    case 7: v = 0.19509032201612826785; break;
    default: v = 0.0; break;
  }
  return neg ? -v : v;
}
__host__ __device__ constexpr double sin32(int j) { return cos32(j - 8); }

// d * exp(SIGN * 2 pi i J / R), J and R compile-time
template <typename T, int R, int J, int SIGN>
__device__ __forceinline__ cplx<T> twmul_const(cplx<T> d) {
  static_assert(32 % R == 0, "radix");
  constexpr int j32 = (J % R) * (32 / R);
  if constexpr (j32 == 0) {
    return d;
  } else if constexpr (j32 == 8) {  // * (SIGN i)
    if constexpr (SIGN > 0) return {-d.y, d.x};
    else return {d.y, -d.x};
  } else if constexpr (j32 == 16) {
    return {-d.x, -d.y};
  } else if constexpr (j32 == 24) {
    if constexpr (SIGN > 0) return {d.y, -d.x};
    else return {-d.y, d.x};
  } else {
    constexpr T c = (T)cos32(j32);
    constexpr T s = (T)(SIGN * sin32(j32));
    return {d.x * c - d.y * s, d.x * s + d.y * c};
  }
}

// --------------------------------------------- in-register radix-R DFT ----
#if GLIA_F32X2
template <typename T, int R, int SIGN, int J>
__device__ __forceinline__ void dif_bf(cplx<T>* v) {
  constexpr int H = R / 2;
  constexpr int j32 = (J % R) * (32 / R);
  cplx<T> a = v[J], b = v[J + H];
  v[J] = cadd(a, b);
  if constexpr (j32 == 8) {
    // (a - b) * (SIGN i) written as two operand-swapped subtractions, so that no negation is needed
    // (the packed add / sub results feed other packed operations, which take no sign modifiers in PTX)
    if constexpr (SIGN > 0) v[J + H] = {b.y - a.y, a.x - b.x};
    else v[J + H] = {a.y - b.y, b.x - a.x};
  } else if constexpr (j32 == 0 || j32 == 16 || j32 == 24) {
    v[J + H] = twmul_const<T, R, J, SIGN>(csub(a, b));
  } else {
    // d (c + i s) = (d.x c, d.y c) + (-s d.y, s d.x): one scaling of the pair + two FMAs
    constexpr T c = (T)cos32(j32);
    constexpr T s = (T)(SIGN * sin32(j32));
    const cplx<T> d = csub(a, b);
    const cplx<T> dc = cscale(d, c);
    v[J + H] = {dc.x - s * d.y, dc.y + s * d.x};
  }
}
#else
template <typename T, int R, int SIGN, int J>
__device__ __forceinline__ void dif_bf(cplx<T>* v) {
  constexpr int H = R / 2;
  cplx<T> a = v[J], b = v[J + H];
  v[J] = cadd(a, b);
  v[J + H] = twmul_const<T, R, J, SIGN>(csub(a, b));
}
#endif
template <typename T, int R, int SIGN, int... J>
__device__ __forceinline__ void dif_level(cplx<T>* v, std::integer_sequence<int, J...>) {
  (dif_bf<T, R, SIGN, J>(v), ...);
}
// in-place DIF, output in bit-reversed order
template <typename T, int R, int SIGN>
struct Dif {
  static __device__ __forceinline__ void run(cplx<T>* v) {
    dif_level<T, R, SIGN>(v, std::make_integer_sequence<int, R / 2>{});
    Dif<T, R / 2, SIGN>::run(v);
    Dif<T, R / 2, SIGN>::run(v + R / 2);
  }
};
template <typename T, int SIGN>
struct Dif<T, 1, SIGN> {
  static __device__ __forceinline__ void run(cplx<T>*) {}
};

__host__ __device__ constexpr int brev(int i, int R) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (i & 1); i >>= 1; }
  return r;
}
template <typename T, int R, int I>
__device__ __forceinline__ void brev_swap(cplx<T>* v) {
  constexpr int J = brev(I, R);
  if constexpr (I < J) { cplx<T> t = v[I]; v[I] = v[J]; v[J] = t; }
}
template <typename T, int R, int... I>
__device__ __forceinline__ void brev_all(cplx<T>* v, std::integer_sequence<int, I...>) {
  (brev_swap<T, R, I>(v), ...);
}
// natural order in -> natural order out, unnormalised, sign SIGN in the exponent
template <typename T, int R, int SIGN>
__device__ __forceinline__ void dft_reg(cplx<T>* v) {
  Dif<T, R, SIGN>::run(v);
  brev_all<T, R>(v, std::make_integer_sequence<int, R>{});
}

// ------------------------------------------------------- the line FFT ----
template <typename T, int N>
struct LineFft {
  using PL = FftPlan<N>;
  static constexpr int E = PL::E, P = PL::P, TPL = N / E;
  __host__ __device__ static constexpr int R(int p) { return p == 0 ? PL::R0 : (p == 1 ? PL::R1 : PL::R2); }
  __host__ __device__ static constexpr int Np(int p) {
    int n = N;
    for (int i = 0; i < p; ++i) n /= R(i);
    return n;
  }
  __host__ __device__ static constexpr int Mp(int p) { return Np(p) / R(p); }
  __host__ __device__ static constexpr int Gp(int p) { return E / R(p); }
  static constexpr int RL = R(P - 1);          // radix of the last pass
  static constexpr int KSTEP = N / RL;         // frequency step between a thread's last-pass registers
  // twiddle registers: passes 0..P-2, (R_p - 1) per butterfly
  __host__ __device__ static constexpr int ntw() {
    int n = 0;
    for (int p = 0; p + 1 < P; ++p) n += Gp(p) * (R(p) - 1);
    return n;
  }
  static constexpr int NTW = ntw() > 0 ? ntw() : 1;
  // GLIA_TW_LDG (probe builds, default 0): three-pass plans (512-point lines) hold 28 complex inter-pass
  // twiddles per thread.  Level 1 reads the pass-0 twiddles from the 4 KB per-axis table at the point of use
  // (read-only path, L1-resident) instead of keeping them in registers; level 2 does so for every pass.
  // Same table entries either way, so results are bit-identical.  ptxas evidence in DESIGN.md 6 (lever 1).
  static constexpr int TWL = (P == 3) ? GLIA_TW_LDG : 0;
  __host__ __device__ static constexpr bool tw_in_regs(int p) { return TWL == 0 || (TWL == 1 && p >= 1); }
  struct Tw {
    cplx<T> w[NTW];
    const cplx<T>* table;
    int t;
  };

  // location (natural position index at pass 0) of register (g, a) of pass p for thread t
  template <int p>
  __device__ static __forceinline__ int loc(int t, int g, int a) {
    constexpr int M = Mp(p), NP = Np(p);
    const int q = t + TPL * g;
    return (q / M) * NP + a * M + (q % M);
  }
  // frequency held in last-pass register (g, c) after forward(): kbase(t,g) + KSTEP*c
  __device__ static __forceinline__ int kbase(int t, int g) {
    const int s = t + TPL * g;
    if constexpr (P == 2) return s;
    else return s / PL::R1 + PL::R0 * (s % PL::R1);
  }
  // inverse map: last-pass location holding frequency k
  __device__ static __forceinline__ int loc_of_freq(int k) {
    if constexpr (P == 2) return (k % PL::R0) * PL::R1 + k / PL::R0;
    else return (k % PL::R0) * (PL::R1 * PL::R2) + ((k / PL::R0) % PL::R1) * PL::R2 + k / (PL::R0 * PL::R1);
  }

  // table[j] = exp(-2 pi i j / N), j < N (built on the host in double)
  __device__ static __forceinline__ void load_twiddles(Tw& tw, const cplx<T>* __restrict__ table, int t) {
    tw.table = table;
    tw.t = t;
    int idx = 0;
    GLIA_UNROLL
    for (int p = 0; p + 1 < P; ++p) {
      GLIA_UNROLL
      for (int g = 0; g < Gp(p); ++g) {
        const int b = (t + TPL * g) % Mp(p);
        GLIA_UNROLL
        for (int c = 1; c < R(p); ++c) {
          if (tw_in_regs(p)) tw.w[idx] = table[(b * c * (N / Np(p))) % N];
          ++idx;
        }
      }
    }
  }

  template <int p, int SIGN>
  __device__ static __forceinline__ void butterflies(cplx<T> (&v)[E]) {
    GLIA_UNROLL
    for (int g = 0; g < Gp(p); ++g) dft_reg<T, R(p), SIGN>(&v[g * R(p)]);
  }
  __host__ __device__ static constexpr int twoff(int p) {
    int n = 0;
    for (int i = 0; i < p; ++i) n += Gp(i) * (R(i) - 1);
    return n;
  }
  template <int p, bool CONJ>
  __device__ static __forceinline__ void twiddle(cplx<T> (&v)[E], const Tw& tw) {
    if constexpr (p + 1 < P) {
      GLIA_UNROLL
      for (int g = 0; g < Gp(p); ++g) {
        GLIA_UNROLL
        for (int c = 1; c < R(p); ++c) {
          cplx<T> w;
          if constexpr (tw_in_regs(p)) {
            w = tw.w[twoff(p) + g * (R(p) - 1) + (c - 1)];
          } else {
            const int b = (tw.t + TPL * g) % Mp(p);
            w = ld_table(tw.table + (b * c * (N / Np(p))) % N);
          }
          v[g * R(p) + c] = CONJ ? cmulc(v[g * R(p) + c], w) : cmul(v[g * R(p) + c], w);
        }
      }
    }
  }
  // move registers from the pass-`pw` placement to the pass-`pr` placement
  template <int pw, int pr, class AM, class SY>
  __device__ static __forceinline__ void exchange(cplx<T> (&v)[E], cplx<T>* sm, AM am, SY sync, int t) {
    sync();  // previous readers of sm are done
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

  // natural positions (pass-0 placement) -> frequencies (last-pass placement)
  template <class AM, class SY>
  __device__ static __forceinline__ void forward(cplx<T> (&v)[E], const Tw& tw, cplx<T>* sm, AM am, SY sync, int t) {
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
  // frequencies (last-pass placement) -> natural positions, unnormalised
  template <class AM, class SY>
  __device__ static __forceinline__ void inverse(cplx<T> (&v)[E], const Tw& tw, cplx<T>* sm, AM am, SY sync, int t) {
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

  // v <- (i * w(k) / N) v with the Nyquist wavenumber zeroed (trap T1;
  // reference: src/cuda/SpectralOperators.cu:54-127).  w/N = kbase/N + c/RL (- 1).
  __device__ static __forceinline__ void mult_iw(cplx<T> (&v)[E], int t) {
    GLIA_UNROLL
    for (int g = 0; g < Gp(P - 1); ++g) {
      const int kb = kbase(t, g);
      const T kbn = (T)kb * (T)(1.0 / N);
      GLIA_UNROLL
      for (int c = 0; c < RL; ++c) {
        T w = kbn + (T)((c < RL / 2) ? (double)c / RL : (double)c / RL - 1.0);
        if (c == RL / 2 && kb == 0) w = (T)0;
        const cplx<T> z = v[g * RL + c];
        v[g * RL + c] = {-w * z.y, w * z.x};
      }
    }
  }
};

// wavenumber with Nyquist zeroed, generic helper
__host__ __device__ inline int wavenumber(int k, int n) {
  return (2 * k < n) ? k : ((2 * k == n) ? 0 : k - n);
}

}  // namespace glia
