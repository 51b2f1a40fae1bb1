// C++ parity tests over include/glia_rd_host.hpp, written the way the reference writes its own
// Catch2 tests (no Catch2 here: plain REQUIRE macro, exit code = number of failures).
//
//   "Running diffusion solver"            src/test/pdesolver.cpp:7-56   (K1: ||c|| = 2.0487, 5 its)
//   "Evaluating objective function"       src/test/grad.cpp:7-80        (K4: J(p*; d = c(T)) < 1e-7, J(p*; 0) > 0)
//
// Build + run: tests/test_gpu_cpp_host.py (nvcc/g++, links libglia_rd.so), on the GPU box.
#include <cstdio>
#include <cstdlib>

#include "glia_rd_host.hpp"

using namespace glia::host;

static int failures = 0;
#define REQUIRE(cond)                                                        \
  do {                                                                       \
    if (!(cond)) { std::printf("REQUIRE failed: %s  (%s:%d)\n", #cond, __FILE__, __LINE__); ++failures; } \
  } while (0)
static bool approx(double a, double b, double rel = 100.0 * 1.1920929e-7) {  // Catch2 Approx default epsilon
  return std::fabs(a - b) <= rel * (1.0 + std::fabs(b));
}

template <typename Real>
void running_diffusion_solver() {
  auto params = std::make_shared<Parameters>();
  params->n[0] = params->n[1] = params->n[2] = 64;  // initializeGrid(64, ...)
  auto spec_ops = std::make_shared<SpectralOperators<Real>>(params);
  auto k = std::make_shared<DiffCoef<Real>>(params, spec_ops);
  REQUIRE(k->setValuesSinusoidal(1E-2) == 0);
  auto diff_solver = std::make_shared<DiffusionSolver<Real>>(params, spec_ops, k);
  Vec<Real> c(params->nl(), params->nl());
  createTestFunction(c, *params);
  REQUIRE(diff_solver->precFactor() == 0);
  for (int i = 0; i < 10; i++) REQUIRE(diff_solver->solve(c, 0.02) == 0);
  const double nrm = c.norm2();
  std::printf("  [pdesolver/%s] ||c|| = %.7f  ksp_itr_ = %d\n", sizeof(Real) == 4 ? "f32" : "f64", nrm, diff_solver->ksp_itr_);
  REQUIRE(approx(nrm, 2.0487));
  REQUIRE(diff_solver->ksp_itr_ == 5);
}

template <typename Real>
void evaluating_objective_function() {
  // grad.cpp: 64^3, dt = .01, nt = 100 (shortened to 10 here), sinusoidal k and rho, self-generated data
  auto params = std::make_shared<Parameters>();
  params->n[0] = params->n[1] = params->n[2] = 64;
  params->dt = 0.01;
  params->nt = 10;
  auto spec_ops = std::make_shared<SpectralOperators<Real>>(params);
  auto k = std::make_shared<DiffCoef<Real>>(params, spec_ops);
  REQUIRE(k->setValuesSinusoidal(1E-2) == 0);
  auto rho = std::make_shared<ReacCoef<Real>>(params, spec_ops);
  REQUIRE(rho->setValues(k->kxx_) == 0);  // the test assigns rho = k (grad.cpp:49)
  auto diff_solver = std::make_shared<DiffusionSolver<Real>>(params, spec_ops, k);
  REQUIRE(diff_solver->precFactor() == 0);
  auto tumor = std::make_shared<Tumor<Real>>(params);
  createTestFunction(tumor->c_0_, *params);
  auto pde = std::make_shared<PdeOperatorsRD<Real>>(tumor, params, spec_ops);
  REQUIRE(pde->solveState(0) == 0);
  Vec<Real> data(params->nl(), params->nl()), zero(params->nl(), params->nl());
  data.copy_from(tumor->c_t_);
  auto mat = std::make_shared<MatProp<Real>>();
  mat->wm_ = std::make_shared<Vec<Real>>(params->nl(), params->nl());
  mat->wm_->copy_from(k->kxx_);
  mat->gm_ = std::make_shared<Vec<Real>>(params->nl(), params->nl());
  mat->csf_ = std::make_shared<Vec<Real>>(params->nl(), params->nl());
  DerivativeOperatorsRD<Real> derivs(pde, tumor, params, spec_ops, mat);
  Vec<Real> dJ(params->nl(), params->nl());
  double J = -1, gk[3], gr[3];
  REQUIRE(derivs.evaluateObjectiveAndGradient(&J, dJ, gk, gr, data) == 0);
  std::printf("  [grad/%s] J(p*; d = c(T)) = %.3e, ksp its state/adj = %d/%d\n", sizeof(Real) == 4 ? "f32" : "f64", J,
              pde->diff_ksp_itr_state_, pde->diff_ksp_itr_adj_);
  REQUIRE(J < 1e-7);
  REQUIRE(derivs.evaluateObjectiveAndGradient(&J, dJ, gk, gr, zero) == 0);
  std::printf("  [grad/%s] J(p*; d = 0)    = %.6e\n", sizeof(Real) == 4 ? "f32" : "f64", J);
  REQUIRE(J > 0);
  // the solver is deterministic: the same call gives the same bits
  double J2 = 0;
  REQUIRE(derivs.evaluateObjectiveAndGradient(&J2, dJ, gk, gr, zero) == 0);
  REQUIRE(J2 == J);
}

template <typename Real>
void applying_phi() {
  // Phi in on-the-fly mode through the host classes: c(0) = Phi p is bounded by max(p) (every basis
  // function is scaled by the common maximum), and <Phi p, f> == <p, Phi^T f>.
  auto params = std::make_shared<Parameters>();
  params->n[0] = params->n[1] = params->n[2] = 64;
  auto spec_ops = std::make_shared<SpectralOperators<Real>>(params);
  const long nl = params->nl();
  MatProp<Real> mat;
  mat.wm_ = std::make_shared<Vec<Real>>(nl, nl);
  mat.wm_->set((Real)0.5);  // white matter everywhere: filter = 1
  REQUIRE(mat.setValuesCustom(*spec_ops, nl) == 0);
  REQUIRE(mat.filter_sum == (double)nl);
  Phi<Real> phi(params, spec_ops);
  const std::vector<double> centers = {3.0, 3.2, 2.9, 2.5, 3.0, 3.4};
  REQUIRE(phi.setValues(centers, 2 * M_PI / 32, &mat) == 0);
  Vec<Real> c0(nl, nl), f(nl, nl), wf(nl, nl);
  const std::vector<double> p = {0.7, -0.4};
  REQUIRE(phi.apply(c0, p) == 0);
  createTestFunction(f, *params);
  REQUIRE(spec_ops->weierstrassSmoother(wf, f, 2 * M_PI / 64) == 0);
  std::vector<double> pt;
  REQUIRE(phi.applyTranspose(pt, wf) == 0);
  std::vector<Real> hc((size_t)nl), hf((size_t)nl);
  c0.to_host(hc.data());
  wf.to_host(hf.data());
  double lhs = 0, cmax = 0;
  for (long i = 0; i < nl; ++i) { lhs += (double)hc[i] * (double)hf[i]; cmax = std::fmax(cmax, std::fabs((double)hc[i])); }
  const double rhs = p[0] * pt[0] + p[1] * pt[1];
  std::printf("  [phi/%s] <Phi p, f> = %.8e  <p, Phi^T f> = %.8e  max|c0| = %.6f\n", sizeof(Real) == 4 ? "f32" : "f64", lhs, rhs, cmax);
  REQUIRE(std::fabs(lhs - rhs) <= (sizeof(Real) == 4 ? 1e-4 : 1e-10) * std::fabs(lhs));
  REQUIRE(cmax <= 0.7 * (1 + 1e-5) && cmax > 0.3);
}

// models 4/5 time-loop body (src/pde/PdeOperatorsMassEffect.cpp:578-631) as a maintainer would write it over the
// mirror: coefficient refresh from the current maps, precFactor, full-dt diffusion, full-dt reaction.  Compiled
// here (so the mirror's signature is checked on every build); the parity case itself is
// tests/test_gpu_parity.py::test_mass_effect_style_steps through the same C entry point.
template <typename Real>
int mass_effect_rd_step(PdeOperatorsRD<Real>& pde, DiffusionSolver<Real>& diff_solver, Tumor<Real>& tumor,
                        const Vec<Real>& bg, const Vec<Real>& gm, const Vec<Real>& vt, const Vec<Real>& csf,
                        const Parameters& params, int i) {
  int ierr = pde.updateReacAndDiffCoefficients(bg, gm, vt, csf);
  if (ierr) return ierr;
  if ((ierr = diff_solver.precFactor())) return ierr;
  if ((ierr = diff_solver.solve(tumor.c_t_, params.dt))) return ierr;
  return pde.reaction(0, i);
}
template int mass_effect_rd_step<float>(PdeOperatorsRD<float>&, DiffusionSolver<float>&, Tumor<float>&, const Vec<float>&,
                                        const Vec<float>&, const Vec<float>&, const Vec<float>&, const Parameters&, int);
template int mass_effect_rd_step<double>(PdeOperatorsRD<double>&, DiffusionSolver<double>&, Tumor<double>&,
                                         const Vec<double>&, const Vec<double>&, const Vec<double>&, const Vec<double>&,
                                         const Parameters&, int);

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::printf("no CUDA device\n"); return 77; }
  std::printf("Running diffusion solver\n");
  running_diffusion_solver<float>();
  running_diffusion_solver<double>();
  std::printf("Evaluating objective function\n");
  evaluating_objective_function<float>();
  evaluating_objective_function<double>();
  std::printf("Applying Phi\n");
  applying_phi<float>();
  applying_phi<double>();
  std::printf("%s (%d failure(s))\n", failures ? "FAILED" : "ALL PASSED", failures);
  return failures;
}
