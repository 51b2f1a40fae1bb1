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

// (Measured in round 1 and removed: Blackwell's packed FP32x2 adds for the butterflies' complex additions took 17 % of
// the D-sweeps' instructions away and changed no kernel's time -- FADD2 holds the FMA pipe for two issue cycles,
// profiles/r1d_f32x2_ab.txt.)
template <typename T>
__device__ __forceinline__ cplx<T> cscale(cplx<T> a, T s) { return {a.x * s, a.y * s}; }

// Loads of operands that are consumed once per kernel.  (Measured in round 1 and removed: evict-first loads /
// stores -- ld.global.cs, st.global.cs -- and an access-policy window on the staged tiles gained nothing on the
// sweeps, which are not DRAM bound at 256^3, and evict-first stores of 64-byte runs tripled kz_r2c's time;
// profiles/r1b_l2_probe.txt, r1c_l2_hints_ab.txt.)
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) { return *p; }
template <typename V> __device__ __forceinline__ void st_stream(V* p, V v) { *p = v; }

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may
// be scheduled while its predecessor's last CTAs are still draining; it does its set-up (indices,
// twiddle registers from the constant per-axis table) and then waits here until the predecessor grid
// has completed and its memory is visible.  Nothing produced by an earlier kernel may be touched
// before this call.
// (Measured in round 1: an explicit griddepcontrol.launch_dependents on entry slows the multi-wave z sweeps by 5-12 %
// at 512^3 and gains nothing over the implicit trigger at CTA exit, profiles/r1l_pdl_matrix.txt -- the wait alone.)
__device__ __forceinline__ void pdl_wait() {
#if !defined(GLIA_SIMT_EMU)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Ensemble batching (BASELINE config 5: independent members, "batched one or more per GPU").  A handle may carry NB
// members, every field laid out [NB][n0][n1][n2].  Lines along z and y never leave an x-plane, so the Z sweeps and the
// y sweeps see a batch as a taller grid; only the x sweeps address a member explicitly.  A CTA always works for ONE
// member (member = blockIdx.x / cpm, cpm = CTAs per member): its partial sums belong to that member, and the CTAs of
// a member whose PCG has converged leave at once.  Per-member PCG state sits DONE_STRIDE ints / SCAL_STRIDE doubles
// apart (pointwise.cuh: I_NISCAL, S_NSCAL).
static constexpr int DONE_STRIDE = 8, SCAL_STRIDE = 16;
__device__ __forceinline__ const int* member_done(const int* done, int member) {
  return done ? done + (size_t)member * DONE_STRIDE : done;
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

// ---- asynchronous global -> shared copies (LDGSTS): no registers, no scoreboard stall --------------
#if defined(GLIA_SIMT_EMU)
__device__ inline void cp_async16(void* smem, const void* g) { std::memcpy(smem, g, 16); }
__device__ inline void cp_async8(void* smem, const void* g) { std::memcpy(smem, g, 8); }
__device__ inline void cp_async_commit() {}
template <int K> __device__ inline void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }
#endif

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

template <typename V>
__device__ __forceinline__ void prefetch_l2(const V* p) {
#if !defined(GLIA_SIMT_EMU)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// One-CTA-per-group Z kernels: the CTA that will run in this CTA's slot `ahead` groups later finds its operand lines
// in L2 instead of HBM.  `bytes_per_cta` contiguous bytes per CTA starting at base + blockIdx.x * bytes_per_cta; the
// CTA's threads request one 128-byte line each (32 KB per field in single precision: one instruction per thread).
#ifndef GLIA_Z_AHEAD
#define GLIA_Z_AHEAD 444  // 148 SMs x 3 resident CTAs
#endif
__device__ __forceinline__ void prefetch_next_cta(const void* base, long bytes_per_cta, long total_bytes) {
  if (GLIA_Z_AHEAD <= 0) return;
  const long b0 = ((long)blockIdx.x + GLIA_Z_AHEAD) * bytes_per_cta;
  for (long i = (long)threadIdx.x * 128; i < bytes_per_cta; i += (long)blockDim.x * 128)
    if (b0 + i < total_bytes) prefetch_l2(reinterpret_cast<const char*>(base) + b0 + i);
}

// ---------------------------------------------------------------- plans ----
// FftPlan<N, V>: V = 0 the default plan; V = 1 the plan of the Z geometry (differs at 512 points only, below)
template <int N, int V = 0> struct FftPlan;
template <> struct FftPlan<32, 0>  { static constexpr int E = 8,  P = 2, R0 = 8,  R1 = 4,  R2 = 1; };
template <> struct FftPlan<64, 0>  { static constexpr int E = 8,  P = 2, R0 = 8,  R1 = 8,  R2 = 1; };
template <> struct FftPlan<128, 0> { static constexpr int E = 16, P = 2, R0 = 16, R1 = 8,  R2 = 1; };
template <> struct FftPlan<256, 0> { static constexpr int E = 16, P = 2, R0 = 16, R1 = 16, R2 = 1; };
template <> struct FftPlan<512, 0> { static constexpr int E = 16, P = 3, R0 = 8,  R1 = 8,  R2 = 8; };
template <int N> struct FftPlan<N, 1> : FftPlan<N, 0> {};
// 512 points need three passes at 16 values per thread.  With radices 2 x 16 x 16 the line is, after the radix-2
// first pass, two independent 256-point transforms, each owned by one half of the line's 32 threads (the second
// exchange only needs that half to meet: LineFft::exchange with a half barrier), and a thread holds 23 instead of 28
// inter-pass twiddles.  Measured on one B200 at 512^3 (profiles/r2h_plan512_ab.txt): the warp-synchronous Z sweeps
// gain 9-11 % (kz_deriv2 512 -> 466 us, kz_c2r.rz 405 -> 361 us, kz_r2c 254 -> 234 us) -- fewer live registers,
// no spill -- while the S sweeps LOSE 5-15 % (x.matvec 624 -> 714 us: the radix-16 butterflies spill more under the
// 128-register cap of a 512-thread CTA than the half barriers give back).  So: Z geometry only.
template <> struct FftPlan<512, 1> { static constexpr int E = 16, P = 3, R0 = 2,  R1 = 16, R2 = 16; };
template <int N> __host__ __device__ constexpr int zplan() { return 1; }

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
template <typename T, int R, int SIGN, int J>
__device__ __forceinline__ void dif_bf(cplx<T>* v) {
  constexpr int H = R / 2;
  cplx<T> a = v[J], b = v[J + H];
  v[J] = cadd(a, b);
  v[J + H] = twmul_const<T, R, J, SIGN>(csub(a, b));
}
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
template <typename T, int N, int V = 0>
struct LineFft {
  using PL = FftPlan<N, V>;
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
  // (Measured in round 2 and removed: reading the inter-pass twiddles of the three-pass 512-point plan from the per-axis
  // table at the point of use instead of holding 28 complex registers -- no kernel gained, the 512^3 step lost 5 %,
  // profiles/r2d_zpipe_twldg_ab.txt.)
  struct Tw {
    cplx<T> w[NTW];
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
    int idx = 0;
    GLIA_UNROLL
    for (int p = 0; p + 1 < P; ++p) {
      GLIA_UNROLL
      for (int g = 0; g < Gp(p); ++g) {
        const int b = (t + TPL * g) % Mp(p);
        GLIA_UNROLL
        for (int c = 1; c < R(p); ++c) {
          tw.w[idx++] = table[(b * c * (N / Np(p))) % N];
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
          const cplx<T> w = tw.w[twoff(p) + g * (R(p) - 1) + (c - 1)];
          v[g * R(p) + c] = CONJ ? cmulc(v[g * R(p) + c], w) : cmul(v[g * R(p) + c], w);
        }
      }
    }
  }
  // After a radix-2 first pass the two halves of the line never exchange data again: passes >= 1 touch only the
  // half [h N/2, (h+1) N/2) with h = t / (TPL/2), in both placements.
  static constexpr bool SPLIT = (P == 3 && PL::R0 == 2);
  // move registers from the pass-`pw` placement to the pass-`pr` placement
  template <int pw, int pr, class AM, class SY>
  __device__ static __forceinline__ void exchange(cplx<T> (&v)[E], cplx<T>* sm, AM am, SY sync, int t) {
    constexpr bool HALF = SPLIT && pw >= 1 && pr >= 1;
    const int h = t / (TPL / 2);
    if constexpr (HALF) sync.half(h); else sync();  // previous readers of sm are done
    GLIA_UNROLL
    for (int g = 0; g < Gp(pw); ++g) {
      GLIA_UNROLL
      for (int a = 0; a < R(pw); ++a) sm[am(loc<pw>(t, g, a))] = v[g * R(pw) + a];
    }
    if constexpr (HALF) sync.half(h); else sync();
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
  // frequencies (last-pass placement) -> natural positions, unnormalised.  Two halves, so that a kernel can put work
  // between the last exchange (after which a thread's own natural positions of `sm` are read by nobody else) and
  // the register-only final pass.
  template <class AM, class SY>
  __device__ static __forceinline__ void inverse_head(cplx<T> (&v)[E], const Tw& tw, cplx<T>* sm, AM am, SY sync, int t) {
    if constexpr (P >= 3) {
      butterflies<2, +1>(v);
      exchange<2, 1>(v, sm, am, sync, t);
    }
    if constexpr (P >= 2) {
      twiddle<1, true>(v, tw);
      butterflies<1, +1>(v);
      exchange<1, 0>(v, sm, am, sync, t);
    }
  }
  __device__ static __forceinline__ void inverse_tail(cplx<T> (&v)[E], const Tw& tw) {
    twiddle<0, true>(v, tw);
    butterflies<0, +1>(v);
  }
  template <class AM, class SY>
  __device__ static __forceinline__ void inverse(cplx<T> (&v)[E], const Tw& tw, cplx<T>* sm, AM am, SY sync, int t) {
    inverse_head(v, tw, sm, am, sync, t);
    inverse_tail(v, tw);
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
