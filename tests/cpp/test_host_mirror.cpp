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

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::printf("no CUDA device\n"); return 77; }
  std::printf("Running diffusion solver\n");
  running_diffusion_solver<float>();
  running_diffusion_solver<double>();
  std::printf("Evaluating objective function\n");
  evaluating_objective_function<float>();
  evaluating_objective_function<double>();
  std::printf("%s (%d failure(s))\n", failures ? "FAILED" : "ALL PASSED", failures);
  return failures;
}
