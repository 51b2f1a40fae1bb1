// c_api.cu -- extern "C" surface of libglia_rd (see include/glia_rd.h).  Everything here is
// argument checking and error translation; the work is in Engine<T> (engine.cuh).
#include "../../include/glia_rd.h"

#include "engine_base.h"
#include <cstring>
#if !defined(GLIA_SIMT_EMU)
#include <cuda_runtime.h>
#endif

using namespace glia;

struct glia_rd {
  EngineBase* eng = nullptr;
  std::string err;
};

namespace {
template <class F>
int guarded(glia_rd_t* h, F f) {
  if (!h) return 2;
  if (!h->eng) { h->err = "handle has no engine (creation failed)"; return 2; }
  try {
    h->eng->v_make_current();
    f(*h->eng);
    return 0;
  } catch (const EngineError& e) {
    h->err = e.msg;
    return 1;
  } catch (const std::exception& e) {
    h->err = e.what();
    return 1;
  }
}
}  // namespace

extern "C" {

int glia_rd_abi_version(void) { return 2; }
const char* glia_rd_build_info(void) {
#if defined(GLIA_SIMT_EMU)
  return "simt-emulator (test only)";
#else
  return "cuda-sm_100a";
#endif
}

static int create_any(glia_rd_t** out, const int n[3], int precision, int device, double dt_ctx, int rank, int nranks,
                      int nbatch) {
  if (!out) return 2;
  *out = nullptr;
  glia_rd_t* h = new glia_rd();
  *out = h;  // also on failure, so that the caller can read the message
  try {
#if !defined(GLIA_SIMT_EMU)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
      cudaGetLastError();
      throw EngineError{"no CUDA device: libglia_rd has no CPU fallback"};
    }
    if (device < 0 || device >= ndev) throw EngineError{"device ordinal out of range"};
#endif
    if (!n) throw EngineError{"n is null"};
    if (precision == GLIA_RD_F32) h->eng = make_engine_f32(n, device, dt_ctx, rank, nranks, nbatch);
    else if (precision == GLIA_RD_F64) h->eng = make_engine_f64(n, device, dt_ctx, rank, nranks, nbatch);
    else throw EngineError{"precision must be 4 or 8"};
  } catch (const EngineError& e) {
    h->err = e.msg;
    return 1;
  } catch (const std::exception& e) {
    h->err = e.what();
    return 1;
  }
  return 0;
}
int glia_rd_create_slab(glia_rd_t** out, const int n[3], int precision, int device, double dt_ctx, int rank,
                        int nranks) {
  return create_any(out, n, precision, device, dt_ctx, rank, nranks, 1);
}
int glia_rd_create(glia_rd_t** out, const int n[3], int precision, int device, double dt_ctx) {
  return create_any(out, n, precision, device, dt_ctx, 0, 1, 1);
}
int glia_rd_create_batch(glia_rd_t** out, const int n[3], int precision, int device, double dt_ctx, int nbatch) {
  return create_any(out, n, precision, device, dt_ctx, 0, 1, nbatch);
}
int glia_rd_batch_size(glia_rd_t* h, int* nbatch) {
  return guarded(h, [&](EngineBase& E) {
    if (!nbatch) throw EngineError{"null output pointer"};
    *nbatch = E.v_nbatch();
  });
}
int glia_rd_batch_iterations(glia_rd_t* h, int* its_per_member, int accumulated) {
  return guarded(h, [&](EngineBase& E) {
    if (!its_per_member) throw EngineError{"null output pointer"};
    E.v_batch_iterations(its_per_member, accumulated);
  });
}
int glia_rd_set_coefficients_batch(glia_rd_t* h, const void* wm, const void* gm, const void* csf, const double* k_scale,
                                   double k_gm_wm, double k_glm_wm, double filter_sum, const double* rho_scale,
                                   double r_gm_wm, double r_glm_wm) {
  return guarded(h, [&](EngineBase& E) {
    if (!wm || !gm || !csf || !k_scale || !rho_scale) throw EngineError{"set_coefficients_batch: null argument"};
    E.v_set_coefficients_batch(wm, gm, csf, k_scale, k_gm_wm, k_glm_wm, filter_sum, rho_scale, r_gm_wm, r_glm_wm);
  });
}
int glia_rd_ipc_export(glia_rd_t* h, int which, void* handle64) {
  return guarded(h, [&](EngineBase& E) {
    if (!handle64) throw EngineError{"null handle buffer"};
    E.v_ipc_export(which, (unsigned char*)handle64);
  });
}
int glia_rd_ipc_connect(glia_rd_t* h, int which, const void* handles) {
  return guarded(h, [&](EngineBase& E) {
    if (!handles) throw EngineError{"null handle array"};
    E.v_ipc_connect(which, (const unsigned char*)handles);
  });
}
int glia_rd_ipc_disconnect(glia_rd_t* h, int which) {
  return guarded(h, [&](EngineBase& E) { E.v_ipc_disconnect(which); });
}
int glia_rd_wait_stream(glia_rd_t* h, void* producer_stream) {
  return guarded(h, [&](EngineBase& E) { E.v_wait_stream(producer_stream); });
}
int glia_rd_set_splitting_order(glia_rd_t* h, int order) {
  return guarded(h, [&](EngineBase& E) {
    if (order != 1 && order != 2) throw EngineError{"splitting order must be 1 or 2"};
    E.v_set_order(order);
  });
}
int glia_rd_set_two_snapshot(glia_rd_t* h, const void* d0, const void* obs0) {
  return guarded(h, [&](EngineBase& E) { E.v_set_two_snapshot(d0, obs0); });
}
int glia_rd_destroy(glia_rd_t* h) {
  if (!h) return 0;
  if (h->eng) {
    try { h->eng->v_make_current(); } catch (...) {}
  }
  delete h->eng;
  delete h;
  return 0;
}
const char* glia_rd_last_error(const glia_rd_t* h) { return h ? h->err.c_str() : "null handle"; }
void* glia_rd_stream(glia_rd_t* h) { return (h && h->eng) ? h->eng->stream_handle() : nullptr; }
long long glia_rd_launch_count(const glia_rd_t* h) { return (h && h->eng) ? h->eng->launches : 0; }

int glia_rd_fft_r2c(glia_rd_t* h, const void* f, void* fhat) {
  return guarded(h, [&](EngineBase& E) { E.v_fft_r2c(f, fhat); });
}
int glia_rd_fft_c2r(glia_rd_t* h, const void* fhat, void* f) {
  return guarded(h, [&](EngineBase& E) { E.v_fft_c2r(fhat, f); });
}
int glia_rd_gradient(glia_rd_t* h, void* gx, void* gy, void* gz, const void* x, int m) {
  return guarded(h, [&](EngineBase& E) { E.v_gradient(gx, gy, gz, x, m); });
}
int glia_rd_divergence(glia_rd_t* h, void* div, const void* dx, const void* dy, const void* dz) {
  return guarded(h, [&](EngineBase& E) { E.v_divergence(div, dx, dy, dz); });
}
int glia_rd_set_diffusion(glia_rd_t* h, const void* k, const double kavg[3], double k_scale) {
  return guarded(h, [&](EngineBase& E) { E.v_set_diffusion(k, kavg, k_scale); });
}
int glia_rd_set_diffusion_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double ks,
                                 double kgm, double kglm, double fsum) {
  return guarded(h, [&](EngineBase& E) { E.v_set_diffusion_tissue(wm, gm, csf, ks, kgm, kglm, fsum); });
}
int glia_rd_set_secondary_k(glia_rd_t* h, const void* kt) {
  return guarded(h, [&](EngineBase& E) { E.v_set_secondary_k(kt); });
}
int glia_rd_set_reaction(glia_rd_t* h, const void* rho) {
  return guarded(h, [&](EngineBase& E) { E.v_set_reaction(rho); });
}
int glia_rd_set_reaction_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double rs, double rgm,
                                double rglm) {
  return guarded(h, [&](EngineBase& E) { E.v_set_reaction_tissue(wm, gm, csf, rs, rgm, rglm); });
}
int glia_rd_update_reac_diff(glia_rd_t* h, const void* bg, const void* gm, const void* vt, const void* csf,
                             double rho_scale, double k_scale, double gm_r_scale, double gm_k_scale) {
  return guarded(h, [&](EngineBase& E) { E.v_update_reac_diff(bg, gm, vt, csf, rho_scale, k_scale, gm_r_scale, gm_k_scale); });
}
int glia_rd_apply_D(glia_rd_t* h, void* dc, const void* c, int secondary) {
  return guarded(h, [&](EngineBase& E) { E.v_apply_D(dc, c, secondary); });
}
int glia_rd_prec_factor(glia_rd_t* h) {
  return guarded(h, [&](EngineBase& E) { E.v_prec_factor(); });
}
int glia_rd_diffusion_solve(glia_rd_t* h, void* c, double dt, int* its) {
  return guarded(h, [&](EngineBase& E) {
    const int k = E.v_diffusion_solve(c, dt);
    if (its) *its = k;
  });
}
int glia_rd_set_ksp_tolerances(glia_rd_t* h, double rtol, double abstol, double dtol, int maxit) {
  return guarded(h, [&](EngineBase& E) { E.v_set_ksp_tolerances(rtol, abstol, dtol, maxit); });
}
int glia_rd_resize_history(glia_rd_t* h, int nt, double dt) {
  return guarded(h, [&](EngineBase& E) { E.v_resize_history(nt, dt); });
}
int glia_rd_history(glia_rd_t* h, int which, int i, void** p) {
  return guarded(h, [&](EngineBase& E) {
    if (!p) throw EngineError{"null output pointer"};
    *p = E.v_history(which, i);
  });
}
int glia_rd_reaction(glia_rd_t* h, void* ct, const void* clin, double dt) {
  return guarded(h, [&](EngineBase& E) { E.v_reaction(ct, clin, dt); });
}
int glia_rd_solve_state(glia_rd_t* h, const void* c0, void* cT, int lin, int* its) {
  return guarded(h, [&](EngineBase& E) {
    const int k = E.v_solve_state(c0, cT, lin);
    if (its) *its = k;
  });
}
int glia_rd_solve_adjoint(glia_rd_t* h, const void* pT, void* p0, int lin, int store, int* its) {
  return guarded(h, [&](EngineBase& E) {
    const int k = E.v_solve_adjoint(pT, p0, lin, store);
    if (its) *its = k;
  });
}
int glia_rd_grad_kappa_rho(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double out[6]) {
  return guarded(h, [&](EngineBase& E) { E.v_grad_kappa_rho(wm, gm, csf, out); });
}
int glia_rd_set_secondary_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double k1, double k2,
                                 double k3) {
  return guarded(h, [&](EngineBase& E) { E.v_set_secondary_tissue(wm, gm, csf, k1, k2, k3); });
}
int glia_rd_objective_gradient(glia_rd_t* h, const void* c0, const void* d1, const void* obs, double beta,
                               const void* wm, const void* gm, const void* csf, double J[4], void* g_c0, double g[6],
                               int ksp_its[2]) {
  return guarded(h, [&](EngineBase& E) {
    if (!c0 || !d1 || !J || !g) throw EngineError{"objective_gradient: null argument"};
    int k[2] = {0, 0};
    E.v_objective_gradient(c0, d1, obs, beta, wm, gm, csf, J, g_c0, g, k);
    if (ksp_its) { ksp_its[0] = k[0]; ksp_its[1] = k[1]; }
  });
}
int glia_rd_hessian_matvec(glia_rd_t* h, const void* c0_tilde, const void* obs, double beta, int diffusivity_inversion,
                           const void* wm, const void* gm, const void* csf, void* y_c0, double hk[6], int ksp_its[4]) {
  return guarded(h, [&](EngineBase& E) {
    if (!c0_tilde || !y_c0 || !hk) throw EngineError{"hessian_matvec: null argument"};
    int k[4] = {0, 0, 0, 0};
    E.v_hessian_matvec(c0_tilde, obs, beta, diffusivity_inversion, wm, gm, csf, y_c0, hk, k);
    if (ksp_its) for (int i = 0; i < 4; ++i) ksp_its[i] = k[i];
  });
}
int glia_rd_smooth(glia_rd_t* h, void* out, const void* in, double sigma) {
  return guarded(h, [&](EngineBase& E) {
    if (!out || !in) throw EngineError{"smooth: null field"};
    if (sigma < 0) throw EngineError{"smooth: negative sigma"};
    E.v_smooth(out, in, sigma);
  });
}
int glia_rd_mat_prop(glia_rd_t* h, void* gm, void* wm, void* vt, void* csf, void* bg, void* filter, double* filter_sum) {
  return guarded(h, [&](EngineBase& E) {
    const double s = E.v_mat_prop(gm, wm, vt, csf, bg, filter);
    if (filter_sum) *filter_sum = s;
  });
}
int glia_rd_phi_set(glia_rd_t* h, int np, const double* centers, double sigma_phi, const void* filter,
                    double sigma_smooth) {
  return guarded(h, [&](EngineBase& E) { E.v_phi_set(np, centers, sigma_phi, filter, sigma_smooth); });
}
int glia_rd_phi_apply(glia_rd_t* h, void* out, const double* p) {
  return guarded(h, [&](EngineBase& E) {
    if (!out || !p) throw EngineError{"phi_apply: null argument"};
    E.v_phi_apply(out, p);
  });
}
int glia_rd_phi_apply_transpose(glia_rd_t* h, double* pout, const void* in) {
  return guarded(h, [&](EngineBase& E) {
    if (!pout || !in) throw EngineError{"phi_apply_transpose: null argument"};
    E.v_phi_apply_transpose(pout, in);
  });
}
int glia_rd_data_in(glia_rd_t* h, const char* path, void* field) {
  return guarded(h, [&](EngineBase& E) {
    if (!path || !field) throw EngineError{"data_in: null argument"};
    E.v_data_in(path, field);
  });
}
int glia_rd_data_out(glia_rd_t* h, const char* path, const void* field) {
  return guarded(h, [&](EngineBase& E) {
    if (!path || !field) throw EngineError{"data_out: null argument"};
    E.v_data_out(path, field);
  });
}
int glia_rd_split_segmentation(glia_rd_t* h, const void* seg, const int labels[4], void* wm, void* gm, void* vt, void* csf) {
  return guarded(h, [&](EngineBase& E) {
    if (!seg || !labels) throw EngineError{"split_segmentation: null argument"};
    E.v_split_segmentation(seg, labels, wm, gm, vt, csf);
  });
}
int glia_rd_probe_xsweep(glia_rd_t* h, int what, int local_mask, int reps, double* ms_per_sweep) {
  return guarded(h, [&](EngineBase& E) {
    const double t = E.v_probe(what, local_mask, reps);
    if (ms_per_sweep) *ms_per_sweep = t;
  });
}
int glia_rd_profile_begin(glia_rd_t* h) {
  return guarded(h, [&](EngineBase& E) { E.v_profile_begin(); });
}
int glia_rd_profile_end(glia_rd_t* h, char* buf, int buflen) {
  return guarded(h, [&](EngineBase& E) {
    const std::string s = E.v_profile_end();
    if (buf && buflen > 0) {
      std::strncpy(buf, s.c_str(), (size_t)buflen - 1);
      buf[buflen - 1] = 0;
    }
  });
}
int glia_rd_timer_start(glia_rd_t* h) {
  return guarded(h, [&](EngineBase& E) { E.v_timer_start(); });
}
int glia_rd_timer_stop_ms(glia_rd_t* h, double* ms) {
  return guarded(h, [&](EngineBase& E) {
    const double t = E.v_timer_stop_ms();
    if (ms) *ms = t;
  });
}
int glia_rd_forward_adjoint(glia_rd_t* h, const void* c0, const void* d1, void* cT, void* p0, int* ks, int* ka) {
  return guarded(h, [&](EngineBase& E) {
    int a = 0, b = 0;
    E.v_forward_adjoint(c0, d1, cT, p0, &a, &b);
    if (ks) *ks = a;
    if (ka) *ka = b;
  });
}
int glia_rd_forward_adjoint_host(glia_rd_t* h, const void* c0, const void* d1, void* cT, void* p0, int* ks, int* ka) {
  return guarded(h, [&](EngineBase& E) {
    int a = 0, b = 0;
    E.v_forward_adjoint_host(c0, d1, cT, p0, &a, &b);
    if (ks) *ks = a;
    if (ka) *ka = b;
  });
}

}  // extern "C"
