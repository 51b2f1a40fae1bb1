// pointwise.cuh -- streaming kernels of the RD path: logistic reaction, PCG vector
// updates, device-resident PCG scalars, reductions.
#pragma once
#include "comm.cuh"
#include "fft_core.cuh"

namespace glia {

__device__ __forceinline__ float g_exp(float x) { return expf(x); }
__device__ __forceinline__ double g_exp(double x) { return exp(x); }
template <typename T>
__device__ __forceinline__ bool g_isinf(T x) { return x == x && (x - x) != (x - x); }

// ---- device-resident PCG state (one block per solver handle) --------------
enum { S_BETA = 0, S_BETAOLD, S_A, S_B, S_DP, S_RNORM0, S_TTOL, S_DPI, S_NSCAL = 16 };
enum { I_ITS = 0, I_DONE, I_TOTAL, I_REASON, I_COMM_ERR, I_NISCAL = 8 };
static_assert(I_NISCAL == DONE_STRIDE && S_NSCAL == SCAL_STRIDE, "per-member strides of the PCG state (fft_core.cuh)");
// reasons follow PETSc's KSPConvergedReason values
enum { KSP_CONVERGED_RTOL = 2, KSP_CONVERGED_ATOL = 3, KSP_DIVERGED_ITS = -3, KSP_DIVERGED_DTOL = -4,
       KSP_DIVERGED_NANORINF = -9, KSP_DIVERGED_INDEFINITE_MAT = -10 };

template <int NV>
__device__ __forceinline__ void sum_partials(const double* __restrict__ partial, int n, double (&out)[NV]) {
  // Fixed-order (deterministic) sum by one CTA of 256 threads: thread t adds entries t, t + 256, ... in that order,
  // a warp adds its lanes in a fixed butterfly, warp 0 adds the warps' sums in a fixed butterfly.  The loads of a
  // thread are independent and issued eight at a time (one L2 round trip per batch, not per entry: this kernel sits
  // on the critical path of every PCG iteration).
  __shared__ double sh[8 * NV];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nth = 256;
  double acc[NV];
  GLIA_UNROLL
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  if (tid < nth) {
    int j = tid;
    for (; j + 7 * nth < n; j += 8 * nth) {
      double v[8][NV];
      GLIA_UNROLL
      for (int u = 0; u < 8; ++u)
        GLIA_UNROLL
        for (int i = 0; i < NV; ++i) v[u][i] = partial[(size_t)(j + u * nth) * NV + i];
      GLIA_UNROLL
      for (int u = 0; u < 8; ++u)
        GLIA_UNROLL
        for (int i = 0; i < NV; ++i) acc[i] += v[u][i];
    }
    for (; j < n; j += nth)
      GLIA_UNROLL
      for (int i = 0; i < NV; ++i) acc[i] += partial[(size_t)j * NV + i];
  }
  GLIA_UNROLL
  for (int i = 0; i < NV; ++i) {
    double v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && wid < 8) sh[wid * NV + i] = v;
  }
  __syncthreads();
  if (wid == 0) {
    GLIA_UNROLL
    for (int i = 0; i < NV; ++i) {
      double v = lane < 8 ? sh[lane * NV + i] : 0.0;
      for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sh[i] = v;
    }
  }
  __syncthreads();
  GLIA_UNROLL
  for (int i = 0; i < NV; ++i) out[i] = sh[i];
  __syncthreads();
}

template <typename T>
__device__ __forceinline__ void pcg_alpha_finish(double dpi, double* scal, int* iscal) {
  scal[S_DPI] = dpi;
  scal[S_BETAOLD] = scal[S_BETA];
  scal[S_A] = (double)(T)(scal[S_BETA] / dpi);
  if (!(dpi > 0.0)) { iscal[I_REASON] = KSP_DIVERGED_INDEFINITE_MAT; iscal[I_TOTAL] += iscal[I_ITS]; iscal[I_DONE] = 1; }
}
template <typename T>
__device__ __forceinline__ void pcg_beta_finish(const double (&rz)[2], double* scal, int* iscal, int maxit, double dtol) {
  const double dp = sqrt(rz[0]);
  scal[S_DP] = dp;
  const int its = iscal[I_ITS] + 1;
  iscal[I_ITS] = its;
  int done = 0, reason = 0;
  if (dp != dp) { done = 1; reason = KSP_DIVERGED_NANORINF; }
  else if (dp <= scal[S_TTOL]) { done = 1; reason = KSP_CONVERGED_RTOL; }
  else if (dp >= dtol * scal[S_RNORM0]) { done = 1; reason = KSP_DIVERGED_DTOL; }
  else if (its >= maxit) { done = 1; reason = KSP_DIVERGED_ITS; }
  const double beta = rz[1];
  scal[S_B] = (double)(T)(beta / scal[S_BETAOLD]);
  scal[S_BETA] = beta;
  iscal[I_REASON] = reason;
  if (done) { iscal[I_DONE] = 1; iscal[I_TOTAL] += its; }
}
// KSPConvergedDefault at iteration 0 with a non-zero initial guess:
//   rnorm0 = ||M^-1 b|| (or dp if that is 0), ttol = max(rtol*rnorm0, abstol), test dp <= ttol.
// (one CTA per ensemble member: blockIdx.x selects the member's partial sums and state block)
static __global__ void k_pcg_init(const double* pb, int nb, const double* prz, int nrz, double* scal, int* iscal,
                           double rtol, double abstol, Comm comm, unsigned epoch, unsigned seq) {
  pb += (size_t)blockIdx.x * nb * 2;
  prz += (size_t)blockIdx.x * nrz * 2;
  scal += (size_t)blockIdx.x * S_NSCAL;
  iscal += (size_t)blockIdx.x * I_NISCAL;
  double b[2], rz[2];
  sum_partials<2>(pb, nb, b);
  sum_partials<2>(prz, nrz, rz);
  {
    double all[4] = {b[0], b[1], rz[0], rz[1]};
    peer_allreduce<4>(comm, epoch, seq, all);
    b[0] = all[0]; b[1] = all[1]; rz[0] = all[2]; rz[1] = all[3];
  }
  if (threadIdx.x == 0) {
    const double dp = sqrt(rz[0]);
    double rnorm0 = sqrt(b[0]);
    if (rnorm0 == 0.0) rnorm0 = dp;
    const double ttol = fmax(rtol * rnorm0, abstol);
    scal[S_DP] = dp; scal[S_RNORM0] = rnorm0; scal[S_TTOL] = ttol;
    scal[S_BETA] = rz[1]; scal[S_BETAOLD] = 1.0; scal[S_A] = 0.0; scal[S_B] = 0.0;
    iscal[I_ITS] = 0;
    int done = 0, reason = 0;
    if (dp != dp) { done = 1; reason = KSP_DIVERGED_NANORINF; }
    else if (dp <= ttol) { done = 1; reason = dp < abstol ? KSP_CONVERGED_ATOL : KSP_CONVERGED_RTOL; }
    else if (rz[1] == 0.0) { done = 1; reason = KSP_CONVERGED_ATOL; }
    iscal[I_DONE] = done; iscal[I_REASON] = reason;
  }
}

// a = beta / <p, A p>
template <typename T>
__global__ void k_pcg_alpha(const double* ppw, int n, double* scal, int* iscal, Comm comm, unsigned epoch,
                            unsigned seq) {
  pdl_wait();
  ppw += (size_t)blockIdx.x * n;
  scal += (size_t)blockIdx.x * S_NSCAL;
  iscal += (size_t)blockIdx.x * I_NISCAL;
  if (iscal[I_DONE]) return;
  double d[1];
  sum_partials<1>(ppw, n, d);
  peer_allreduce<1>(comm, epoch, seq, d);
  if (threadIdx.x == 0) pcg_alpha_finish<T>(d[0], scal, iscal);
}

// after z = M^-1 r: dp = ||z||, its++, convergence test, beta = <r,z>, b = beta/betaold
template <typename T>
__global__ void k_pcg_beta(const double* prz, int n, double* scal, int* iscal, int maxit, double dtol, Comm comm,
                           unsigned epoch, unsigned seq) {
  pdl_wait();
  prz += (size_t)blockIdx.x * n * 2;
  scal += (size_t)blockIdx.x * S_NSCAL;
  iscal += (size_t)blockIdx.x * I_NISCAL;
  if (iscal[I_DONE]) return;
  double rz[2];
  sum_partials<2>(prz, n, rz);
  peer_allreduce<2>(comm, epoch, seq, rz);
  if (threadIdx.x == 0) pcg_beta_finish<T>(rz, scal, iscal, maxit, dtol);
}

// iteration `it` (1-based): x += a p ; if not converged p = z + b p.
// Runs iff iteration `it` really executed (it <= I_ITS): the converged iteration still
// owes x its update (VecAXPY(X,a,P) precedes the test in KSPSolve_CG), a speculative
// launch past convergence must do nothing.
// `xin` is where the iterate lives before this update: x itself, or -- first iteration of an out-of-place solve --
// the field the solve started from (the time loops solve from one history slot into the next, PdeOperators.cpp:
// 284-300, so that no c_[i+1] / c_half_[i] / p_[i] copy is left).
template <typename T>
__global__ void k_cg_update(long n, const T* xin, T* x, T* p, const T* __restrict__ z, const double* scal,
                            const int* iscal, int it) {
  pdl_wait();
  // blockIdx.y = ensemble member; n = elements of ONE member
  scal += (size_t)blockIdx.y * S_NSCAL;
  iscal += (size_t)blockIdx.y * I_NISCAL;
  if (it > iscal[I_ITS]) return;
  const size_t mo = (size_t)blockIdx.y * (size_t)n;
  xin += mo; x += mo; p += mo; z += mo;
  const T a = (T)scal[S_A], b = (T)scal[S_B];
  const int done = iscal[I_DONE];
  const long stride = (long)gridDim.x * blockDim.x;
  // 16-byte accesses where the member's fields allow it (n a multiple of the vector width, 16-byte aligned bases)
  constexpr int V = 16 / (int)sizeof(T);
  struct alignas(16) Vec { T e[V]; };
  const bool vec = (n % V == 0) && ((((size_t)xin | (size_t)x | (size_t)p | (size_t)z) & 15) == 0);
  if (vec) {
    const long nv = n / V;
    const Vec* xin4 = reinterpret_cast<const Vec*>(xin);
    const Vec* z4 = reinterpret_cast<const Vec*>(z);
    Vec* x4 = reinterpret_cast<Vec*>(x);
    Vec* p4 = reinterpret_cast<Vec*>(p);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
      const Vec pv = p4[i], xv = xin4[i];
      Vec xo;
      GLIA_UNROLL
      for (int k = 0; k < V; ++k) xo.e[k] = xv.e[k] + a * pv.e[k];
      x4[i] = xo;
      if (!done) {
        const Vec zv = z4[i];
        Vec po;
        GLIA_UNROLL
        for (int k = 0; k < V; ++k) po.e[k] = zv.e[k] + b * pv.e[k];
        p4[i] = po;
      }
    }
    return;
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T pv = p[i];
    x[i] = xin[i] + a * pv;
    if (!done) p[i] = z[i] + b * pv;
  }
}

// ---- logistic reaction (src/pde/PdeOperators.cpp:140-190, 318-370) -----------
// nonlinear: a = c/(1-c); c <- a f/(a f + 1), f = exp(rho dt); c <- 1 if a is inf.
// `1.0 - c` and `a*f + 1.0` are double expressions in the reference (trap T6).
// (cin may be c: in place; the time loop reads c_half_[i] and writes c_[i+1])
template <typename T>
__global__ void k_reaction(long n, const T* cin, T* c, const T* __restrict__ rho, T dt, T* c_half_out) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T cv = cin[i];
    if (c_half_out) c_half_out[i] = cv;
    const T factor = g_exp((T)(rho[i] * dt));
    const T alph = (T)((double)cv / (1.0 - (double)cv));
    T o;
    if (g_isinf(alph)) o = (T)1.0;
    else {
      const T af = alph * factor;
      o = (T)((double)af / ((double)af + 1.0));
    }
    c[i] = o;
  }
}
// linearised / adjoint: u <- u f / (c f + 1 - c)^2
template <typename T>
__global__ void k_reaction_lin(long n, const T* uin, T* u, const T* __restrict__ rho, const T* __restrict__ clin, T dt) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T cv = clin[i];
    const T factor = g_exp((T)(rho[i] * dt));
    const T cf = cv * factor;
    const T alph = (T)(((double)cf + 1.0) - (double)cv);
    const T uf = uin[i] * factor;
    u[i] = uf / (alph * alph);
  }
}

// t = 0.5*(ci + cj) + (first ? ci : cj)   (solveIncremental, src/pde/PdeOperators.cpp:199-226)
template <typename T>
__global__ void k_incr_avg(long n, T* t, const T* __restrict__ ci, const T* __restrict__ cj, int first) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T v = ci[i] + cj[i];
    v = v * (T)0.5;
    v = v + (first ? ci[i] : cj[i]);
    t[i] = v;
  }
}

// PdeOperatorsMassEffect::updateReacAndDiffCoefficients (src/pde/PdeOperatorsMassEffect.cpp:98-138):
// the per-time-step coefficient refresh of the mass-effect models, one pass for both fields.
//   rho = rho_s * max(0, 1 - (bg + gm_r*gm + vt + csf)),  k = k_s * max(0, 1 - (bg + gm_k*gm + vt + csf))
// (sum order as written there; kyy, kzz are copies of kxx, i.e. the same isotropic field here).
template <typename T>
__global__ void k_update_reac_diff(long n, T* rho, T* k, const T* __restrict__ bg, const T* __restrict__ gm,
                                   const T* __restrict__ vt, const T* __restrict__ csf, T rho_s, T k_s, T gm_r, T gm_k) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T b = bg[i], g = gm[i], v = vt[i], c = csf[i];
    T t = (T)1 - (((b + gm_r * g) + v) + c);
    t = (t < (T)0) ? (T)0 : t;
    rho[i] = t * rho_s;
    t = (T)1 - (((b + gm_k * g) + v) + c);
    t = (t < (T)0) ? (T)0 : t;
    k[i] = t * k_s;
  }
}

// out = a*x + b*y (y may be null)
template <typename T>
__global__ void k_axpby(long n, T* out, T a, const T* x, T b, const T* y) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}

// mismatch and adjoint terminal condition of one objective evaluation
// (src/grad/DerivativeOperatorsRD.cpp:24-29, 78-84; Obs::apply/applyT = mask product,
// src/mat/Obs.cpp:75-140):   t = O c - d1 (d1 may be null) ;  pT = -(O t) ;
// partial sums { <t,t>, <c0,c0> } (c0 may be null).
template <typename T>
__global__ void k_obs_mismatch(long n, const T* __restrict__ c, const T* __restrict__ d1, const T* __restrict__ obs,
                               const T* __restrict__ c0, T* pT, double* partial) {
  double acc[4] = {0, 0, 0, 0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T t = c[i];
    if (obs) t = t * obs[i];
    if (d1) t = t - d1[i];
    T q = t;
    if (obs) q = q * obs[i];
    pT[i] = q * (T)-1.0;
    acc[0] += (double)t * (double)t;
    if (c0) acc[1] += (double)c0[i] * (double)c0[i];
  }
  __shared__ double red[32 * 4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  GLIA_UNROLL
  for (int j = 0; j < 4; ++j) {
    double v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid * 4 + j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double sum = 0;
    for (int w = 0; w < nwarp; ++w) sum += red[w * 4 + threadIdx.x];
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] = sum;
  }
}

// partial sums of up to 3 dot products <m_j, t> plus sum(t)
template <typename T>
__global__ void k_dot3(long n, const T* __restrict__ t, const T* __restrict__ m0, const T* __restrict__ m1,
                       const T* __restrict__ m2, double* partial) {
  double acc[4] = {0, 0, 0, 0};
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double tv = (double)t[i];
    if (m0) acc[0] += (double)m0[i] * tv;
    if (m1) acc[1] += (double)m1[i] * tv;
    if (m2) acc[2] += (double)m2[i] * tv;
    acc[3] += tv;
  }
  // block reduce
  __shared__ double red[32 * 4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  GLIA_UNROLL
  for (int j = 0; j < 4; ++j) {
    double v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid * 4 + j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0;
    for (int w = 0; w < nwarp; ++w) s += red[w * 4 + threadIdx.x];
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}
static __global__ void k_sum4(const double* partial, int n, double* out, Comm comm, unsigned epoch, unsigned seq) {
  double o[4];
  sum_partials<4>(partial, n, o);
  peer_allreduce<4>(comm, epoch, seq, o);
  if (threadIdx.x == 0) { out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3]; }
}

}  // namespace glia
