// c_api.cu -- extern "C" surface of libglia_rd (see include/glia_rd.h).
#include "../../include/glia_rd.h"

#include "engine.cuh"
#include "spectral3d.cuh"

using namespace glia;

struct glia_rd {
  EngineBase* eng = nullptr;
  std::string err;
};

namespace {
template <class F>
int guarded(glia_rd_t* h, F f) {
  if (!h || !h->eng) return 2;
  try {
    f();
    return 0;
  } catch (const EngineError& e) {
    h->err = e.msg;
    return 1;
  } catch (const std::exception& e) {
    h->err = e.what();
    return 1;
  }
}
#define WITH_ENGINE(h, body)                                                  \
  guarded(h, [&]() {                                                          \
    if (h->eng->precision() == 4) { auto& E = *static_cast<Engine<float>*>(h->eng); using T = float; (void)sizeof(T); body; } \
    else { auto& E = *static_cast<Engine<double>*>(h->eng); using T = double; (void)sizeof(T); body; }                      \
  })
}  // namespace

extern "C" {

int glia_rd_abi_version(void) { return 1; }
const char* glia_rd_build_info(void) {
#if defined(GLIA_SIMT_EMU)
  return "simt-emulator (test only)";
#else
  return "cuda-sm_100a";
#endif
}

int glia_rd_create(glia_rd_t** out, const int n[3], int precision, int device, double dt_ctx) {
  if (!out) return 2;
  *out = nullptr;
  glia_rd_t* h = new glia_rd();
  try {
#if !defined(GLIA_SIMT_EMU)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
      cudaGetLastError();
      throw EngineError{"no CUDA device: libglia_rd has no CPU fallback"};
    }
#endif
    if (precision == GLIA_RD_F32) h->eng = new Engine<float>(n, device, dt_ctx);
    else if (precision == GLIA_RD_F64) h->eng = new Engine<double>(n, device, dt_ctx);
    else throw EngineError{"precision must be 4 or 8"};
  } catch (const EngineError& e) {
    h->err = e.msg;
    *out = h;  // so the caller can read the message
    return 1;
  }
  *out = h;
  return 0;
}
int glia_rd_destroy(glia_rd_t* h) {
  if (!h) return 0;
  delete h->eng;
  delete h;
  return 0;
}
const char* glia_rd_last_error(const glia_rd_t* h) { return h ? h->err.c_str() : "null handle"; }
void* glia_rd_stream(glia_rd_t* h) {
  if (!h || !h->eng) return nullptr;
  if (h->eng->precision() == 4) return (void*)(intptr_t) static_cast<Engine<float>*>(h->eng)->st;
  return (void*)(intptr_t) static_cast<Engine<double>*>(h->eng)->st;
}
long long glia_rd_launch_count(const glia_rd_t* h) { return (h && h->eng) ? h->eng->launches : 0; }

int glia_rd_fft_r2c(glia_rd_t* h, const void* f, void* fhat) {
  return WITH_ENGINE(h, fft3d_r2c(E, (const T*)f, (cplx<T>*)fhat));
}
int glia_rd_fft_c2r(glia_rd_t* h, const void* fhat, void* f) {
  return WITH_ENGINE(h, fft3d_c2r(E, (const cplx<T>*)fhat, (T*)f));
}
int glia_rd_gradient(glia_rd_t* h, void* gx, void* gy, void* gz, const void* x, int m) {
  return WITH_ENGINE(h, E.gradient((T*)gx, (T*)gy, (T*)gz, (const T*)x, m));
}
int glia_rd_divergence(glia_rd_t* h, void* div, const void* dx, const void* dy, const void* dz) {
  return WITH_ENGINE(h, E.divergence((T*)div, (const T*)dx, (const T*)dy, (const T*)dz));
}
int glia_rd_set_diffusion(glia_rd_t* h, const void* k, const double kavg[3], double k_scale) {
  return WITH_ENGINE(h, E.set_diffusion((const T*)k, kavg, k_scale));
}
int glia_rd_set_diffusion_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double ks,
                                 double kgm, double kglm, double fsum) {
  return WITH_ENGINE(h, E.set_diffusion_tissue((const T*)wm, (const T*)gm, (const T*)csf, ks, kgm, kglm, fsum));
}
int glia_rd_set_secondary_k(glia_rd_t* h, const void* kt) {
  return WITH_ENGINE(h, { GLIA_CHECK(rt::copy(E.ktil, kt, sizeof(T) * E.nreal, E.st)); E.sync(); });
}
int glia_rd_set_reaction(glia_rd_t* h, const void* rho) {
  return WITH_ENGINE(h, { GLIA_CHECK(rt::copy(E.rho, rho, sizeof(T) * E.nreal, E.st)); E.sync(); });
}
int glia_rd_set_reaction_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double rs, double rgm,
                                double rglm) {
  return WITH_ENGINE(h, E.set_reaction_tissue((const T*)wm, (const T*)gm, (const T*)csf, rs, rgm, rglm));
}
int glia_rd_apply_D(glia_rd_t* h, void* dc, const void* c, int secondary) {
  return WITH_ENGINE(h, E.apply_D((T*)dc, (const T*)c, secondary != 0));
}
int glia_rd_prec_factor(glia_rd_t* h) { return WITH_ENGINE(h, E.prec_factor()); }
int glia_rd_diffusion_solve(glia_rd_t* h, void* c, double dt, int* its) {
  return WITH_ENGINE(h, {
    int k = E.diffusion_solve((T*)c, dt);
    E.sync();
    if (its) *its = k;
  });
}
int glia_rd_set_ksp_tolerances(glia_rd_t* h, double rtol, double abstol, double dtol, int maxit) {
  return WITH_ENGINE(h, { E.rtol = rtol; E.abstol = abstol; E.dtol = dtol; E.maxit = maxit; });
}
int glia_rd_resize_history(glia_rd_t* h, int nt, double dt) { return WITH_ENGINE(h, E.resize_history(nt, dt)); }
int glia_rd_history(glia_rd_t* h, int which, int i, void** p) {
  return WITH_ENGINE(h, { *p = (void*)E.hist(which, i); });
}
int glia_rd_reaction(glia_rd_t* h, void* ct, const void* clin, double dt) {
  return WITH_ENGINE(h, { E.reaction((T*)ct, (const T*)clin, (T)dt, nullptr); E.sync(); });
}
int glia_rd_solve_state(glia_rd_t* h, const void* c0, void* cT, int lin, int* its) {
  return WITH_ENGINE(h, {
    int k = E.solve_state((const T*)c0, (T*)cT, lin);
    if (its) *its = k;
  });
}
int glia_rd_solve_adjoint(glia_rd_t* h, const void* pT, void* p0, int lin, int store, int* its) {
  return WITH_ENGINE(h, {
    int k = E.solve_adjoint((const T*)pT, (T*)p0, lin, store);
    if (its) *its = k;
  });
}
int glia_rd_grad_kappa_rho(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double out[6]) {
  return WITH_ENGINE(h, E.grad_kappa_rho((const T*)wm, (const T*)gm, (const T*)csf, out));
}
int glia_rd_timer_start(glia_rd_t* h) { return WITH_ENGINE(h, E.timer.start(E.st)); }
int glia_rd_timer_stop_ms(glia_rd_t* h, double* ms) { return WITH_ENGINE(h, { *ms = E.timer.stop_ms(E.st); }); }
int glia_rd_forward_adjoint_host(glia_rd_t* h, const void* c0, const void* d1, void* cT, void* p0, int* ks, int* ka) {
  return WITH_ENGINE(h, E.forward_adjoint_host((const T*)c0, (const T*)d1, (T*)cT, (T*)p0, ks, ka));
}

}  // extern "C"
