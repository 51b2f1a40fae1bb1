// engine_base.h -- the precision-erased interface between the C ABI (c_api.cu) and the
// templated engine (engine.cuh, instantiated in engine_f32.cu / engine_f64.cu).
#pragma once
#include <string>

namespace glia {

struct EngineError {
  std::string msg;
};

// Type-erased face of Engine<T> used by the C ABI (c_api.cu); `void*` arguments are device
// pointers to T unless a name ends in _host.
class EngineBase {
 public:
  virtual ~EngineBase() {}
  long long launches = 0;
  virtual int precision() const = 0;
  virtual void* stream_handle() = 0;
  // the CUDA device is per-host-thread state: every C-ABI entry makes the handle's device current first
  virtual void v_make_current() = 0;
  virtual void v_ipc_export(int which, unsigned char* out64) = 0;
  virtual void v_ipc_connect(int which, const unsigned char* handles) = 0;
  virtual void v_ipc_disconnect(int which) = 0;
  virtual void v_wait_stream(void* producer_stream) = 0;
  virtual void v_set_order(int order) = 0;
  virtual int v_nbatch() const = 0;
  virtual void v_batch_iterations(int* out, int accumulated) = 0;
  virtual void v_set_coefficients_batch(const void* wm, const void* gm, const void* csf, const double* k_scale, double kgm,
                                        double kglm, double filter_sum, const double* rho_scale, double rgm, double rglm) = 0;
  virtual void v_set_two_snapshot(const void* d0, const void* obs0) = 0;
  virtual void v_fft_r2c(const void* f, void* fhat) = 0;
  virtual void v_fft_c2r(const void* fhat, void* f) = 0;
  virtual void v_gradient(void* gx, void* gy, void* gz, const void* x, int mask) = 0;
  virtual void v_divergence(void* div, const void* dx, const void* dy, const void* dz) = 0;
  virtual void v_set_diffusion(const void* k, const double kavg[3], double k_scale) = 0;
  virtual void v_set_diffusion_tissue(const void* wm, const void* gm, const void* csf, double ks, double kgm,
                                      double kglm, double filter_sum) = 0;
  virtual void v_set_secondary_k(const void* kt) = 0;
  virtual void v_set_reaction(const void* rho) = 0;
  virtual void v_set_reaction_tissue(const void* wm, const void* gm, const void* csf, double rs, double rgm,
                                     double rglm) = 0;
  virtual void v_update_reac_diff(const void* bg, const void* gm, const void* vt, const void* csf, double rho_s,
                                  double k_s, double gm_r, double gm_k) = 0;
  virtual void v_apply_D(void* dc, const void* c, int secondary) = 0;
  virtual void v_prec_factor() = 0;
  virtual int v_diffusion_solve(void* c, double dt) = 0;
  virtual void v_set_ksp_tolerances(double rtol, double abstol, double dtol, int maxit) = 0;
  virtual void v_resize_history(int nt, double dt) = 0;
  virtual void* v_history(int which, int i) = 0;
  virtual void v_reaction(void* ct, const void* clin, double dt) = 0;
  virtual int v_solve_state(const void* c0, void* cT, int linearized) = 0;
  virtual int v_solve_adjoint(const void* pT, void* p0, int linearized, int adjoint_store) = 0;
  virtual void v_grad_kappa_rho(const void* wm, const void* gm, const void* csf, double out[6]) = 0;
  virtual void v_set_secondary_tissue(const void* wm, const void* gm, const void* csf, double k1, double k2, double k3) = 0;
  virtual void v_objective_gradient(const void* c0, const void* d1, const void* obs, double beta, const void* wm,
                                    const void* gm, const void* csf, double J[4], void* g_c0, double g[6], int ksp[2]) = 0;
  virtual void v_hessian_matvec(const void* c0t, const void* obs, double beta, int diffusivity_inversion, const void* wm,
                                const void* gm, const void* csf, void* y_c0, double hk[6], int ksp[4]) = 0;
  virtual void v_smooth(void* out, const void* in, double sigma) = 0;
  virtual double v_mat_prop(void* gm, void* wm, void* vt, void* csf, void* bg, void* filter) = 0;
  virtual void v_phi_set(int np, const double* centers, double sigma_phi, const void* filter, double sigma_smooth) = 0;
  virtual void v_phi_apply(void* out, const double* p) = 0;
  virtual void v_phi_apply_transpose(double* pout, const void* in) = 0;
  virtual void v_data_in(const char* path, void* field) = 0;
  virtual void v_data_out(const char* path, const void* field) = 0;
  virtual void v_split_segmentation(const void* seg, const int labels[4], void* wm, void* gm, void* vt, void* csf) = 0;
  virtual double v_probe(int what, int local_mask, int reps) = 0;
  virtual void v_profile_begin() = 0;
  virtual std::string v_profile_end() = 0;
  virtual void v_timer_start() = 0;
  virtual double v_timer_stop_ms() = 0;
  virtual void v_forward_adjoint(const void* c0, const void* d1, void* cT, void* p0, int* ks, int* ka) = 0;
  virtual void v_forward_adjoint_host(const void* c0, const void* d1, void* cT, void* p0, int* ks, int* ka) = 0;
};
EngineBase* make_engine_f32(const int n[3], int device, double dt_ctx, int rank, int nranks, int nbatch);
EngineBase* make_engine_f64(const int n[3], int device, double dt_ctx, int rank, int nranks, int nbatch);

}  // namespace glia
