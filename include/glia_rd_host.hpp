// glia_rd_host.hpp -- C++ host layer over the C ABI of libglia_rd (include/glia_rd.h).
//
// The reference keeps its RD hot path behind five C++ classes; a GLIA build that links
// libglia_rd keeps those classes and forwards their methods.  PETSc, MPI and
// AccFFT are not available where this library is developed, so this header carries a minimal
// stand-in for that host side -- same class and method names, same argument meaning, same
// PetscErrorCode-style return (0 = success) -- over a small device `Vec` shim in the PETSc
// layout.  It is what the C++ parity tests (tests/cpp/) are written against, so that they read
// like the reference's own Catch2 tests (src/test/pdesolver.cpp, src/test/grad.cpp).
//
//   reference class (file)                                    here
//   SpectralOperators  include/grad/SpectralOperators.h:6-53     glia::host::SpectralOperators
//   DiffCoef           include/mat/DiffCoef.h:16-65              glia::host::DiffCoef
//   ReacCoef           include/mat/ReacCoef.h                    glia::host::ReacCoef
//   DiffusionSolver    include/pde/DiffusionSolver.h:7-42        glia::host::DiffusionSolver
//   PdeOperatorsRD     include/pde/PdeOperators.h:10-77          glia::host::PdeOperatorsRD
//   DerivativeOperatorsRD include/grad/DerivativeOperators.h:9-81 glia::host::DerivativeOperatorsRD
//   Phi                include/mat/Phi.h (on-the-fly mode)       glia::host::Phi
//   MatProp            include/mat/MatProp.h                     glia::host::MatProp
//
// Header-only; needs the CUDA runtime (device memory of the Vec shim) and -lglia_rd.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "glia_rd.h"

namespace glia {
namespace host {

typedef int ErrorCode;  // PetscErrorCode: 0 = success

template <typename Real>
struct Precision;
template <> struct Precision<float> { static constexpr int code = GLIA_RD_F32; };
template <> struct Precision<double> { static constexpr int code = GLIA_RD_F64; };

// ---- Vec shim: device array of nl local entries out of ng global ones (VecCreate / VecSetSizes
// / setupVec, src/mat/Tumor.cpp:11-16, src/utils/Utils.cpp:512-526) -----------------------------
template <typename Real>
class Vec {
 public:
  Vec(long nl, long ng) : nl_(nl), ng_(ng) {
    if (cudaMalloc((void**)&d_, sizeof(Real) * (size_t)nl) != cudaSuccess) throw std::runtime_error("Vec: cudaMalloc failed");
    set(0);
  }
  ~Vec() { cudaFree(d_); }
  Vec(const Vec&) = delete;
  Vec& operator=(const Vec&) = delete;
  Real* array() { return d_; }  // vecGetArray (src/utils/Utils.cpp:69-93): the device pointer
  const Real* array() const { return d_; }
  long local_size() const { return nl_; }
  long global_size() const { return ng_; }
  void set(Real v) {  // VecSet
    std::vector<Real> h((size_t)nl_, v);
    from_host(h.data());
  }
  void from_host(const Real* h) { cudaMemcpy(d_, h, sizeof(Real) * (size_t)nl_, cudaMemcpyHostToDevice); }
  void to_host(Real* h) const { cudaMemcpy(h, d_, sizeof(Real) * (size_t)nl_, cudaMemcpyDeviceToHost); }
  void copy_from(const Vec& o) { cudaMemcpy(d_, o.d_, sizeof(Real) * (size_t)nl_, cudaMemcpyDeviceToDevice); }  // VecCopy
  double norm2() const {  // VecNorm(NORM_2) of the local part (single-rank tests)
    std::vector<Real> h((size_t)nl_);
    to_host(h.data());
    double s = 0;
    for (Real v : h) s += (double)v * (double)v;
    return std::sqrt(s);
  }

 private:
  Real* d_ = nullptr;
  long nl_, ng_;
};

// ---- Parameters / Grid subset the path reads (include/Parameters.h:150-230, 374-451) ----------
struct Parameters {
  int n[3] = {64, 64, 64};
  double dt = 0.5;  // tu_->dt_, default 0.5 (Parameters.h:166)
  int nt = 1;
  double k = 0, k_gm_wm_ratio = 0, k_glm_wm_ratio = 0;
  double rho = 0, r_gm_wm_ratio = 0, r_glm_wm_ratio = 0;
  double beta = 0;
  int nk = 1, nr = 1;
  bool diffusivity_inversion = false;
  long nl() const { return (long)n[0] * n[1] * n[2]; }
  double lebesgue_measure() const { return (2 * M_PI / n[0]) * (2 * M_PI / n[1]) * (2 * M_PI / n[2]); }
};

// owns the glia_rd_t handle (the role SpectralOperators::setup + initializeGrid play)
template <typename Real>
class SpectralOperators {
 public:
  // initializeGrid(n, params, spec_ops)  (src/grad/SpectralOperators.cpp:424-446)
  explicit SpectralOperators(std::shared_ptr<Parameters> params, int device = 0) : params_(params) {
    if (glia_rd_create(&h_, params->n, Precision<Real>::code, device, params->dt) != 0) {
      std::string msg = glia_rd_last_error(h_);
      glia_rd_destroy(h_);
      throw std::runtime_error("glia_rd_create: " + msg);
    }
  }
  ~SpectralOperators() { glia_rd_destroy(h_); }
  glia_rd_t* handle() { return h_; }
  ErrorCode executeFFTR2C(const Real* f, void* f_hat) { return glia_rd_fft_r2c(h_, f, f_hat); }
  ErrorCode executeFFTC2R(const void* f_hat, Real* f) { return glia_rd_fft_c2r(h_, f_hat, f); }
  // computeGradient(grad_x, grad_y, grad_z, x, pXYZ)  (SpectralOperators.cpp:100-177)
  ErrorCode computeGradient(Vec<Real>& gx, Vec<Real>& gy, Vec<Real>& gz, const Vec<Real>& x, int xyz = 7) {
    return glia_rd_gradient(h_, gx.array(), gy.array(), gz.array(), x.array(), xyz);
  }
  ErrorCode computeDivergence(Vec<Real>& div, const Vec<Real>& dx, const Vec<Real>& dy, const Vec<Real>& dz) {
    return glia_rd_divergence(h_, div.array(), dx.array(), dy.array(), dz.array());
  }
  // weierstrassSmoother(wc, c, params, sigma)  (SpectralOperators.cpp:263-381); wc may alias c
  ErrorCode weierstrassSmoother(Vec<Real>& wc, const Vec<Real>& c, double sigma) {
    return glia_rd_smooth(h_, wc.array(), c.array(), sigma);
  }
  std::string lastError() const { return glia_rd_last_error(h_); }

 private:
  std::shared_ptr<Parameters> params_;
  glia_rd_t* h_ = nullptr;
};

// tissue maps (MatProp fields the coefficients are built from, src/mat/MatProp.cpp)
template <typename Real>
struct MatProp {
  std::shared_ptr<Vec<Real>> wm_, gm_, csf_, vt_, bg_, filter_;
  double filter_sum = 0;  // sum of the brain mask (MatProp.cpp:180-185)
  // setValuesCustom(gm, wm, csf, vt, bg, params)  (MatProp.cpp:135-201): clip, bg, filter
  ErrorCode setValuesCustom(SpectralOperators<Real>& spec_ops, long nl) {
    if (!bg_) bg_ = std::make_shared<Vec<Real>>(nl, nl);
    if (!filter_) filter_ = std::make_shared<Vec<Real>>(nl, nl);
    return glia_rd_mat_prop(spec_ops.handle(), gm_ ? gm_->array() : nullptr, wm_ ? wm_->array() : nullptr,
                            vt_ ? vt_->array() : nullptr, csf_ ? csf_->array() : nullptr, bg_->array(), filter_->array(),
                            &filter_sum);
  }
};

// dataIn / dataOut (src/utils/IO.cpp:511-640) and splitSegmentation, atlas form (src/utils/Utils.cpp:592-657)
template <typename Real>
ErrorCode dataIn(Vec<Real>& A, SpectralOperators<Real>& spec_ops, const std::string& fname) {
  return glia_rd_data_in(spec_ops.handle(), fname.c_str(), A.array());
}
template <typename Real>
ErrorCode dataOut(const Vec<Real>& A, SpectralOperators<Real>& spec_ops, const std::string& fname) {
  return glia_rd_data_out(spec_ops.handle(), fname.c_str(), A.array());
}
template <typename Real>
ErrorCode splitSegmentation(const Vec<Real>& seg, Vec<Real>* wm, Vec<Real>* gm, Vec<Real>* vt, Vec<Real>* csf,
                            SpectralOperators<Real>& spec_ops, const std::vector<int>& labels) {
  const int lab[4] = {labels[0], labels[1], labels[2], labels.size() > 3 ? labels[3] : 0};
  return glia_rd_split_segmentation(spec_ops.handle(), seg.array(), lab, wm ? wm->array() : nullptr, gm ? gm->array() : nullptr,
                                    vt ? vt->array() : nullptr, csf ? csf->array() : nullptr);
}

// Phi in on-the-fly mode (include/mat/Phi.h, src/mat/Phi.cpp:24-120, 324-434)
template <typename Real>
class Phi {
 public:
  Phi(std::shared_ptr<Parameters> params, std::shared_ptr<SpectralOperators<Real>> spec_ops)
      : params_(params), spec_ops_(spec_ops) {}
  // setGaussians / setValues: centres (radians), sigma, the MatProp filter, smoothing_factor
  ErrorCode setValues(const std::vector<double>& centers, double sigma, const MatProp<Real>* mat_prop,
                      double smoothing_factor = 1.0) {
    np_ = (int)(centers.size() / 3);
    sigma_ = sigma;
    const double sigma_smooth = smoothing_factor * 2.0 * M_PI / params_->n[0];  // Phi.cpp:338
    return glia_rd_phi_set(spec_ops_->handle(), np_, centers.data(), sigma,
                           (mat_prop && mat_prop->filter_) ? mat_prop->filter_->array() : nullptr, sigma_smooth);
  }
  ErrorCode apply(Vec<Real>& out, const std::vector<double>& p) {  // Phi::apply(out, p)
    return glia_rd_phi_apply(spec_ops_->handle(), out.array(), p.data());
  }
  ErrorCode applyTranspose(std::vector<double>& pout, const Vec<Real>& in) {  // Phi::applyTranspose(pout, in)
    pout.assign((size_t)np_, 0.0);
    return glia_rd_phi_apply_transpose(spec_ops_->handle(), pout.data(), in.array());
  }
  int np_ = 0;
  double sigma_ = 0;

 private:
  std::shared_ptr<Parameters> params_;
  std::shared_ptr<SpectralOperators<Real>> spec_ops_;
};

template <typename Real>
class DiffCoef {
 public:
  DiffCoef(std::shared_ptr<Parameters> params, std::shared_ptr<SpectralOperators<Real>> spec_ops)
      : params_(params), spec_ops_(spec_ops), kxx_(params->nl(), params->nl()) {}
  // setValuesSinusoidal(params, scale)  (src/mat/DiffCoef.cpp:134-177)
  ErrorCode setValuesSinusoidal(double scale) {
    const int* n = params_->n;
    std::vector<Real> h((size_t)params_->nl());
    const double freq = 4.0;
    double sum = 0;
    for (int x = 0; x < n[0]; ++x)
      for (int y = 0; y < n[1]; ++y)
        for (int z = 0; z < n[2]; ++z) {
          const double X = 2.0 * M_PI / n[0] * x, Y = 2.0 * M_PI / n[1] * y, Z = 2.0 * M_PI / n[2] * z;
          const Real v = (Real)((double)(Real)scale * (0.5 + 0.5 * std::sin(freq * X) * std::sin(freq * Y) * std::sin(freq * Z)));
          h[((size_t)x * n[1] + y) * n[2] + z] = v;
          sum += (double)v;
        }
    kxx_.from_host(h.data());
    k_scale_ = scale;
    const Real avg = (Real)sum * ((Real)1.0 / (Real)params_->nl());
    kxx_avg_ = kyy_avg_ = kzz_avg_ = avg;
    const double ka[3] = {(double)avg, (double)avg, (double)avg};
    return glia_rd_set_diffusion(spec_ops_->handle(), kxx_.array(), ka, scale);
  }
  // setValues(k_scale, k_gm_wm_ratio, k_glm_wm_ratio, mat_prop, params)  (DiffCoef.cpp:77-131)
  ErrorCode setValues(double k_scale, double k_gm_wm_ratio, double k_glm_wm_ratio, const MatProp<Real>& m) {
    k_scale_ = k_scale;
    return glia_rd_set_diffusion_tissue(spec_ops_->handle(), m.wm_->array(), m.gm_->array(), m.csf_->array(), k_scale,
                                        k_gm_wm_ratio, k_glm_wm_ratio, m.filter_sum);
  }
  // setSecondaryCoefficients(k1, k2, k3, mat_prop, params)  (DiffCoef.cpp:44-59)
  ErrorCode setSecondaryCoefficients(double k1, double k2, double k3, const MatProp<Real>& m) {
    if (params_->nk == 1) { k2 = params_->k_gm_wm_ratio * k1; k3 = params_->k_glm_wm_ratio * k1; }
    return glia_rd_set_secondary_tissue(spec_ops_->handle(), m.wm_->array(), m.gm_->array(), m.csf_->array(), k1, k2, k3);
  }
  ErrorCode applyD(Vec<Real>& dc, const Vec<Real>& c) { return glia_rd_apply_D(spec_ops_->handle(), dc.array(), c.array(), 0); }
  ErrorCode applyDWithSecondaryCoeffs(Vec<Real>& dc, const Vec<Real>& c) {
    return glia_rd_apply_D(spec_ops_->handle(), dc.array(), c.array(), 1);
  }
  Vec<Real> kxx_;
  double k_scale_ = 0;
  Real kxx_avg_ = 0, kyy_avg_ = 0, kzz_avg_ = 0;

 private:
  std::shared_ptr<Parameters> params_;
  std::shared_ptr<SpectralOperators<Real>> spec_ops_;
};

template <typename Real>
class ReacCoef {
 public:
  ReacCoef(std::shared_ptr<Parameters> params, std::shared_ptr<SpectralOperators<Real>> spec_ops)
      : params_(params), spec_ops_(spec_ops) {}
  // setValues(rho_scale, r_gm_wm_ratio, r_glm_wm_ratio, mat_prop, params)  (src/mat/ReacCoef.cpp:13-38)
  ErrorCode setValues(double rho_scale, double r_gm_wm_ratio, double r_glm_wm_ratio, const MatProp<Real>& m) {
    return glia_rd_set_reaction_tissue(spec_ops_->handle(), m.wm_->array(), m.gm_->array(), m.csf_->array(), rho_scale,
                                       r_gm_wm_ratio, r_glm_wm_ratio);
  }
  ErrorCode setValues(const Vec<Real>& rho_vec) { return glia_rd_set_reaction(spec_ops_->handle(), rho_vec.array()); }

 private:
  std::shared_ptr<Parameters> params_;
  std::shared_ptr<SpectralOperators<Real>> spec_ops_;
};

template <typename Real>
class DiffusionSolver {
 public:
  DiffusionSolver(std::shared_ptr<Parameters> params, std::shared_ptr<SpectralOperators<Real>> spec_ops,
                  std::shared_ptr<DiffCoef<Real>> k)
      : params_(params), spec_ops_(spec_ops), k_(k) {}
  int ksp_itr_ = 0;
  ErrorCode precFactor() { return glia_rd_prec_factor(spec_ops_->handle()); }             // DiffusionSolver.cpp:119-180
  ErrorCode solve(Vec<Real>& c, double dt) {                                               // DiffusionSolver.cpp:217-250
    return glia_rd_diffusion_solve(spec_ops_->handle(), c.array(), dt, &ksp_itr_);
  }

 private:
  std::shared_ptr<Parameters> params_;
  std::shared_ptr<SpectralOperators<Real>> spec_ops_;
  std::shared_ptr<DiffCoef<Real>> k_;
};

// state container subset (src/mat/Tumor.cpp): c_0, c_t, p_0, p_t
template <typename Real>
struct Tumor {
  explicit Tumor(std::shared_ptr<Parameters> p) : c_0_(p->nl(), p->nl()), c_t_(p->nl(), p->nl()), p_0_(p->nl(), p->nl()), p_t_(p->nl(), p->nl()) {}
  Vec<Real> c_0_, c_t_, p_0_, p_t_;
};

template <typename Real>
class PdeOperatorsRD {
 public:
  PdeOperatorsRD(std::shared_ptr<Tumor<Real>> tumor, std::shared_ptr<Parameters> params,
                 std::shared_ptr<SpectralOperators<Real>> spec_ops)
      : tumor_(tumor), params_(params), spec_ops_(spec_ops) {
    if (glia_rd_resize_history(spec_ops_->handle(), params->nt, params->dt) != 0)       // PdeOperators.cpp:17-103
      throw std::runtime_error("resize_history: " + spec_ops_->lastError());
  }
  int diff_ksp_itr_state_ = 0, diff_ksp_itr_adj_ = 0;
  ErrorCode solveState(int linearized) {                                                   // PdeOperators.cpp:235-316
    return glia_rd_solve_state(spec_ops_->handle(), tumor_->c_0_.array(), tumor_->c_t_.array(), linearized, &diff_ksp_itr_state_);
  }
  ErrorCode solveAdjoint(int linearized, int adjoint_store = 1) {                          // PdeOperators.cpp:372-420
    return glia_rd_solve_adjoint(spec_ops_->handle(), tumor_->p_t_.array(), tumor_->p_0_.array(), linearized, adjoint_store,
                                 &diff_ksp_itr_adj_);
  }
  ErrorCode reaction(int linearized, int iter) {                                           // PdeOperators.cpp:140-190
    const Real* c_lin = nullptr;
    if (linearized) c_lin = c_(iter);
    return glia_rd_reaction(spec_ops_->handle(), tumor_->c_t_.array(), c_lin, params_->dt);
  }
  // PdeOperatorsMassEffect::updateReacAndDiffCoefficients(seg, tumor)  (PdeOperatorsMassEffect.cpp:98-138):
  // the per-step coefficient refresh of models 4/5; bg/gm/vt/csf are the CURRENT (advected) maps.
  ErrorCode updateReacAndDiffCoefficients(const Vec<Real>& bg, const Vec<Real>& gm, const Vec<Real>& vt,
                                          const Vec<Real>& csf) {
    return glia_rd_update_reac_diff(spec_ops_->handle(), bg.array(), gm.array(), vt.array(), csf.array(), params_->rho,
                                    params_->k, 1.0 - params_->r_gm_wm_ratio, 1.0 - params_->k_gm_wm_ratio);
  }
  const Real* c_(int i) { return hist(GLIA_HIST_C, i); }
  const Real* p_(int i) { return hist(GLIA_HIST_P, i); }
  const Real* c_half_(int i) { return hist(GLIA_HIST_C_HALF, i); }

 private:
  const Real* hist(int which, int i) {
    void* p = nullptr;
    if (glia_rd_history(spec_ops_->handle(), which, i, &p) != 0) throw std::runtime_error(spec_ops_->lastError());
    return (const Real*)p;
  }
  std::shared_ptr<Tumor<Real>> tumor_;
  std::shared_ptr<Parameters> params_;
  std::shared_ptr<SpectralOperators<Real>> spec_ops_;
};

// kappa / rho blocks of the gradient as the reference assembles them from the six dot products
// (src/grad/DerivativeOperators.cpp:231-249, 293-313)
inline void assemble_kappa_rho(const double g6[6], const Parameters& p, double* g_kappa /*nk*/, double* g_rho /*nr*/) {
  g_kappa[0] = g6[0] + (p.nk == 1 ? p.k_gm_wm_ratio * g6[1] : 0.0);
  if (p.nk > 1) g_kappa[1] = g6[1];
  if (p.nk > 2) g_kappa[2] = g6[2];
  g_rho[0] = g6[3] + (p.nr == 1 ? p.r_gm_wm_ratio * g6[4] : 0.0);
  if (p.nr > 1) g_rho[1] = g6[4];
  if (p.nr > 2) g_rho[2] = g6[5];
}

template <typename Real>
class DerivativeOperatorsRD {
 public:
  DerivativeOperatorsRD(std::shared_ptr<PdeOperatorsRD<Real>> pde, std::shared_ptr<Tumor<Real>> tumor,
                        std::shared_ptr<Parameters> params, std::shared_ptr<SpectralOperators<Real>> spec_ops,
                        std::shared_ptr<MatProp<Real>> mat_prop, std::shared_ptr<Vec<Real>> obs_filter = nullptr)
      : pde_(pde), tumor_(tumor), params_(params), spec_ops_(spec_ops), mat_prop_(mat_prop), obs_(obs_filter) {}
  // evaluateObjectiveAndGradient(J, dJ, x, data) in field space (DerivativeOperatorsRD.cpp:130-226):
  // c(0) = tumor_->c_0_ (= Phi p), data d1; dJ_field = -h^3 (alpha(0) - beta c0) (g_p = Phi^T dJ_field);
  // g_kappa[nk], g_rho[nr] as gradDiffusion / gradReaction assemble them.
  ErrorCode evaluateObjectiveAndGradient(double* J, Vec<Real>& dJ_field, double* g_kappa, double* g_rho, const Vec<Real>& d1) {
    double Jv[4], g6[6];   // J, D(c1), S(c0), D(c0) -- the last one 0 unless setTwoSnapshot()
    int its[2];
    ErrorCode e = glia_rd_objective_gradient(spec_ops_->handle(), tumor_->c_0_.array(), d1.array(), obs_ ? obs_->array() : nullptr,
                                             params_->beta, mat_prop_->wm_->array(), mat_prop_->gm_->array(),
                                             mat_prop_->csf_->array(), Jv, dJ_field.array(), g6, its);
    if (e) return e;
    *J = Jv[0];
    assemble_kappa_rho(g6, *params_, g_kappa, g_rho);
    pde_->diff_ksp_itr_state_ = its[0];
    pde_->diff_ksp_itr_adj_ = its[1];
    return 0;
  }
  // params_->tu_->two_time_points_ with Data::dt0() and Obs::filter_0_ (DerivativeOperatorsRD.cpp:30-34, 149-153,
  // 216-222); obs0 == nullptr is O0 = I, d0 == nullptr switches the terms off
  ErrorCode setTwoSnapshot(const Vec<Real>* d0, const Vec<Real>* obs0 = nullptr) {
    return glia_rd_set_two_snapshot(spec_ops_->handle(), d0 ? d0->array() : nullptr, obs0 ? obs0->array() : nullptr);
  }
  // evaluateHessian(y, x) in field space (DerivativeOperatorsRD.cpp:229-438): x = (c0~ = Phi p~, k~);
  // y_field as above, y_kappa[nk] = Hkp p~ + Hkk k~ when diffusivity_inversion is set.
  ErrorCode evaluateHessian(Vec<Real>& y_field, double* y_kappa, const Vec<Real>& c0_tilde, const double* k_tilde) {
    if (params_->diffusivity_inversion) {
      const double k1 = k_tilde[0], k2 = params_->nk > 1 ? k_tilde[1] : 0.0, k3 = params_->nk > 2 ? k_tilde[2] : 0.0;
      double kk2 = k2, kk3 = k3;
      if (params_->nk == 1) { kk2 = params_->k_gm_wm_ratio * k1; kk3 = params_->k_glm_wm_ratio * k1; }
      ErrorCode e = glia_rd_set_secondary_tissue(spec_ops_->handle(), mat_prop_->wm_->array(), mat_prop_->gm_->array(),
                                                 mat_prop_->csf_->array(), k1, kk2, kk3);
      if (e) return e;
    }
    double hk[6];
    int its[4];
    ErrorCode e = glia_rd_hessian_matvec(spec_ops_->handle(), c0_tilde.array(), obs_ ? obs_->array() : nullptr, params_->beta,
                                         params_->diffusivity_inversion ? 1 : 0, mat_prop_->wm_->array(),
                                         mat_prop_->gm_->array(), mat_prop_->csf_->array(), y_field.array(), hk, its);
    if (e) return e;
    if (params_->diffusivity_inversion && y_kappa) {
      for (int b = 0; b < 2; ++b) {
        const double* g = hk + 3 * b;
        const double v0 = g[0] + (params_->nk == 1 ? params_->k_gm_wm_ratio * g[1] : 0.0);
        if (b == 0) { y_kappa[0] = v0; if (params_->nk > 1) y_kappa[1] = g[1]; if (params_->nk > 2) y_kappa[2] = g[2]; }
        else { y_kappa[0] += v0; if (params_->nk > 1) y_kappa[1] += g[1]; if (params_->nk > 2) y_kappa[2] += g[2]; }
      }
    }
    return 0;
  }

 private:
  std::shared_ptr<PdeOperatorsRD<Real>> pde_;
  std::shared_ptr<Tumor<Real>> tumor_;
  std::shared_ptr<Parameters> params_;
  std::shared_ptr<SpectralOperators<Real>> spec_ops_;
  std::shared_ptr<MatProp<Real>> mat_prop_;
  std::shared_ptr<Vec<Real>> obs_;
};

// createTestFunction(c, params): exp(-r^2/R^2), R = sqrt(2) 2 pi / 64, centred at (pi, pi, pi)
// (src/test/helper.cpp:19-49)
template <typename Real>
inline void createTestFunction(Vec<Real>& c, const Parameters& p) {
  const int* n = p.n;
  std::vector<Real> h((size_t)p.nl());
  const Real R = (Real)(std::sqrt(2.) * (2 * M_PI) / 64);
  const Real hx = (Real)(2.0 * M_PI / n[0]), hy = (Real)(2.0 * M_PI / n[1]), hz = (Real)(2.0 * M_PI / n[2]);
  for (int x = 0; x < n[0]; ++x)
    for (int y = 0; y < n[1]; ++y)
      for (int z = 0; z < n[2]; ++z) {
        // ScalarType h * int64, minus the double M_PI, rounded back to ScalarType -- as the reference does
        const Real dx = (Real)((double)(hx * (Real)x) - M_PI), dy = (Real)((double)(hy * (Real)y) - M_PI),
                   dz = (Real)((double)(hz * (Real)z) - M_PI);
        const Real r = std::sqrt(dx * dx + dy * dy + dz * dz);
        const Real ratio = r / R;
        h[((size_t)x * n[1] + y) * n[2] + z] = std::exp(-ratio * ratio);
      }
  c.from_host(h.data());
}

}  // namespace host
}  // namespace glia
