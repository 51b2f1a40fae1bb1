// spectral3d.cuh -- stand-alone 3-D R2C / C2R in the reference's layout
// (SpectralOperators::executeFFTR2C / executeFFTC2R, src/grad/SpectralOperators.cpp:68-98):
// real [n0][n1][n2]  <->  complex [n0][n1][n2/2+1], unnormalised.
//
// Internally the sweeps work on the packed half spectrum [n0][n1][n2/2] whose z slot 0
// carries the DC plane in the real part and the Nyquist plane in the imaginary part of
// the *z transform*; after the (complex) y and x transforms the two planes are untangled
// with the Hermitian symmetry of their 2-D spectra.
#pragma once
#include "engine.cuh"

namespace glia {

template <typename T>
__global__ void k_unpack_half(int n0, int n1, int n2c, const cplx<T>* __restrict__ packed, cplx<T>* __restrict__ full) {
  const long total = (long)n0 * n1 * (n2c + 1);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int kz = (int)(i % (n2c + 1));
    const long xy = i / (n2c + 1);
    const int ky = (int)(xy % n1), kx = (int)(xy / n1);
    if (kz >= 1 && kz < n2c) {
      full[i] = packed[xy * n2c + kz];
    } else {
      const cplx<T> g = packed[xy * n2c];
      const long xym = (long)((n0 - kx) % n0) * n1 + ((n1 - ky) % n1);
      const cplx<T> gm = packed[xym * n2c];
      if (kz == 0) full[i] = {(T)0.5 * (g.x + gm.x), (T)0.5 * (g.y - gm.y)};
      else full[i] = {(T)0.5 * (g.y + gm.y), (T)-0.5 * (g.x - gm.x)};
    }
  }
}

template <typename T>
__global__ void k_pack_half(int n0, int n1, int n2c, const cplx<T>* __restrict__ full, cplx<T>* __restrict__ packed) {
  const long total = (long)n0 * n1 * n2c;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int kz = (int)(i % n2c);
    const long xy = i / n2c;
    if (kz >= 1) {
      packed[i] = full[xy * (n2c + 1) + kz];
    } else {
      const cplx<T> f0 = full[xy * (n2c + 1)], fn = full[xy * (n2c + 1) + n2c];
      packed[i] = {f0.x - fn.y, f0.y + fn.x};
    }
  }
}

template <typename T>
void fft3d_r2c(Engine<T>& E, const T* f, cplx<T>* fhat) {
  using C = cplx<T>;
  const TileS ty = E.tile_y(), tx = E.tile_x();
  GLIA_DISPATCH_N(E.n[2], E.L("kz_r2c", kz_r2c<T, N, 0>, E.template grid_z<N>(), dim3(zthreads<N>()),
                                       Engine<T>::template smem_z<N>(), E.st, E.lines_z(), const_cast<T*>(f),
                                       (const T*)nullptr, (const double*)nullptr, E.shat, (const C*)E.tw[2],
                                       (const int*)nullptr, E.template bpm_z<N>()));
  GLIA_DISPATCH_N(E.n[1], E.L("ks_c2c", ks_c2c<T, N, -1>, Engine<T>::grid_s(ty), Engine<T>::template block_s<N>(),
                                       Engine<T>::template smem_s<N>(), E.st, ty, (const C*)E.shat, E.shat,
                                       (const C*)E.tw[1], (const int*)nullptr));
  GLIA_DISPATCH_N(E.n[0], E.L("ks_c2c", ks_c2c<T, N, -1>, Engine<T>::grid_s(tx), Engine<T>::template block_s<N>(),
                                       Engine<T>::template smem_s<N>(), E.st, tx, (const C*)E.shat, E.shat,
                                       (const C*)E.tw[0], (const int*)nullptr));
  E.L("k_unpack_half", k_unpack_half<T>, Engine<T>::grid_pw(E.ncplx), dim3(256), 0, E.st, E.n[0], E.n[1], E.n2c,
               (const C*)E.shat, fhat);
  E.sync();
}

template <typename T>
void fft3d_c2r(Engine<T>& E, const cplx<T>* fhat, T* f) {
  using C = cplx<T>;
  const TileS ty = E.tile_y(), tx = E.tile_x();
  E.L("k_pack_half", k_pack_half<T>, Engine<T>::grid_pw(E.ncplx), dim3(256), 0, E.st, E.n[0], E.n[1], E.n2c, fhat, E.shat);
  GLIA_DISPATCH_N(E.n[0], E.L("ks_c2c", ks_c2c<T, N, +1>, Engine<T>::grid_s(tx), Engine<T>::template block_s<N>(),
                                       Engine<T>::template smem_s<N>(), E.st, tx, (const C*)E.shat, E.shat,
                                       (const C*)E.tw[0], (const int*)nullptr));
  GLIA_DISPATCH_N(E.n[1], E.L("ks_c2c", ks_c2c<T, N, +1>, Engine<T>::grid_s(ty), Engine<T>::template block_s<N>(),
                                       Engine<T>::template smem_s<N>(), E.st, ty, (const C*)E.shat, E.shat,
                                       (const C*)E.tw[1], (const int*)nullptr));
  GLIA_DISPATCH_N(E.n[2], E.L("kz_c2r", kz_c2r<T, N, 0>, E.template grid_z<N>(), dim3(zthreads<N>()),
                                       Engine<T>::template smem_z<N>(), E.st, E.lines_z(), (const C*)E.shat, f,
                                       (const T*)nullptr, (double*)nullptr, (const C*)E.tw[2], (const int*)nullptr,
                                       E.template bpm_z<N>()));
  E.sync();
}

}  // namespace glia
