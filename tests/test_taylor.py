"""Taylor test of the gradient and symmetry / finite-difference checks of the Gauss-Newton
Hessian -- the reference's own consistency gates, restated:

  DerivativeOperators::checkGradient   src/grad/DerivativeOperators.cpp:369-450
      |J(x + h dx) - J(x) - h <g, dx>|  for h = 10^0 .. 10^-5
  DerivativeOperators::checkHessian    src/grad/DerivativeOperators.cpp:452-542
  DerivativeOperators::computeFDHessian  :546-631 (finite differences of the gradient)

The quantities no reference unit test pins (the kappa / rho gradient, the c(0) gradient and
the Hessian product) are validated here against the OBJECTIVE ITSELF, first on the CPU oracle
(so the oracle is no longer its own judge), then through the C ABI on the GPU.

What the test can demand follows from how the reference derives its gradient:
  * c(0) block: the adjoint recursion (half diffusion, reaction linearised about c_half_, half
    diffusion) is the exact transpose of the discrete Strang step, so the Taylor remainder falls
    with slope 2 until rounding -- demanded over >= 3 decades;
  * kappa / rho blocks: optimise-then-discretise (trapezoid time integrals of grad c . grad alpha
    and alpha c (c - 1), src/grad/DerivativeOperators.cpp:189-321), consistent with the discrete
    objective to O(dt^2): slope 2 while h is above that floor, agreement with a central difference
    to ~1e-5 relative at nt = 4, and a 4x smaller gap at half the time step.
"""
import numpy as np
import pytest

import _cases as Cs
from oracle import rd_oracle as O

N = 32
T_END = 0.16
K0, R0, KGM, RGM = 0.05, 8.0, 0.2, 0.2
BETA = 1e-3
KSP_RTOL = 1e-13


def _problem(dtype=np.float64):
    P = Cs.make_problem(N, dtype)
    obs = (P["wm"] > 0.2).astype(dtype)
    d1 = (0.7 * P["c0"]).astype(dtype)
    dc = (Cs.smooth_field((N, N, N), dtype, 9, -1.0, 1.0) * (0.3 * float(P["c0"].max()))).astype(dtype)
    return P, obs, d1, dc


# ------------------------------------------------------------------ oracle side ----
class OracleModel:
    """J(c0, kappa, rho) and its gradient by the CPU restatement."""

    def __init__(self, P, obs, d1, nt):
        self.P, self.obs, self.d1, self.nt, self.dt = P, obs, d1, nt, T_END / nt

    def _ops(self, ks, rs):
        P = self.P
        dtype = P["c0"].dtype
        k = O.DiffCoef((N, N, N), dtype)
        k.set_values(ks, KGM, 0.0, P["wm"], P["gm"], P["csf"], P["filt"])
        rho = O.reac_coef(rs, RGM, 0.0, P["wm"], P["gm"], P["csf"])
        pde = O.PdeOperatorsRD(k, rho, self.nt, self.dt, dt_ctx=self.dt)
        pde.diff.RTOL = KSP_RTOL
        pde.diff.prec_factor()
        return pde, O.DerivativeOperatorsRD(pde, P["wm"], P["gm"], P["csf"], obs=self.obs, beta=BETA)

    def J(self, c0, ks=K0, rs=R0):
        pde, D = self._ops(ks, rs)
        cT = pde.solve_state(c0, 0)
        temp, _ = D._terminal(cT, self.d1)
        d = O.DiffusionSolver._dot
        return D.leb * 0.5 * d(temp, temp) + 0.5 * BETA * d(c0, c0) * D.leb

    def grad(self, c0, ks=K0, rs=R0):
        pde, D = self._ops(ks, rs)
        r = D.evaluate_objective_and_gradient(c0, self.d1)
        g = r["g6"]
        # nk = nr = 1: the gm ratio folds into the single scale (DerivativeOperators.cpp:236-243, 298-305)
        return r["J"], r["g_c0"], g[0] + KGM * g[1], g[3] + RGM * g[4], (pde, D)


# ------------------------------------------------------------------ C-ABI side ----
class AbiModel:
    """The same through libglia_rd (any backend of tests/_cases.py)."""

    def __init__(self, B, P, obs, d1, nt):
        self.B, self.P, self.nt, self.dt = B, P, nt, T_END / nt
        dtype = P["c0"].dtype
        self.h = B.handle(N, dtype, dt_ctx=self.dt)
        self.dev = {k: B.put(P[k]) for k in ("wm", "gm", "csf")}
        self.obs_d, self.d1_d = B.put(obs), B.put(d1)
        self.fsum = float(P["filt"].sum(dtype=np.float64))
        self.h.set_ksp_tolerances(rtol=KSP_RTOL)
        self.h.resize_history(nt, self.dt)
        self.gc0 = B.empty((N, N, N), dtype)

    def _coef(self, ks, rs):
        d = self.dev
        self.h.set_diffusion_tissue(d["wm"], d["gm"], d["csf"], ks, KGM, 0.0, self.fsum)
        self.h.set_reaction_tissue(d["wm"], d["gm"], d["csf"], rs, RGM, 0.0)
        self.h.prec_factor()

    def grad(self, c0, ks=K0, rs=R0):
        self._coef(ks, rs)
        d = self.dev
        r = self.h.objective_gradient(self.B.put(c0), self.d1_d, d["wm"], d["gm"], d["csf"], obs=self.obs_d, beta=BETA,
                                      g_c0=self.gc0)
        g = r["g6"]
        return r["J"], self.B.get(self.gc0).copy(), g[0] + KGM * g[1], g[3] + RGM * g[4], None

    def J(self, c0, ks=K0, rs=R0):
        return self.grad(c0, ks, rs)[0]

    def close(self):
        self.h.close()


def _taylor(J, J0, slope_dot, hs):
    return [abs(J(h) - J0 - h * slope_dot) for h in hs]


def _check_gradient(M, P, dc, nt_pair_model=None):
    """-> dict of findings; asserts the slope-2 / consistency statements of the module docstring."""
    c0 = P["c0"]
    J0, g_c0, g_k, g_r, _ = M.grad(c0)
    out = {"J0": J0}
    # ---- c(0) direction: exact discrete adjoint, slope 2 over >= 3 decades
    hs = [1e-1, 1e-2, 1e-3, 1e-4]
    dot = float(np.sum(g_c0.astype(np.float64) * dc.astype(np.float64)))
    r = _taylor(lambda h: M.J((c0 + h * dc).astype(c0.dtype)), J0, dot, hs)
    out["taylor_c0"] = r
    for a, b in zip(r[:-1], r[1:]):
        assert 80.0 < a / b < 125.0, ("c0 Taylor remainder is not second order", r)
    # ---- kappa, rho: slope 2 above the O(dt^2) consistency floor, central difference agreement
    for name, g, fn, scale in (("kappa", g_k, lambda h: M.J(c0, ks=K0 * (1 + h)), K0),
                               ("rho", g_r, lambda h: M.J(c0, rs=R0 * (1 + h)), R0)):
        r = _taylor(fn, J0, g * scale, [1e-1, 1e-2, 1e-3])
        out["taylor_" + name] = r
        for a, b in zip(r[:-1], r[1:]):
            assert 60.0 < a / b < 140.0, (name + " Taylor remainder is not second order", r)
        e = 1e-4
        fd = (fn(e) - fn(-e)) / (2 * e) / scale
        out["fd_" + name] = (g, fd)
        assert abs(fd - g) <= 1e-4 * abs(fd), (name, g, fd)
    return out


# ================================================================= CPU: the oracle ====
def test_oracle_gradient_taylor():
    P, obs, d1, dc = _problem()
    M = OracleModel(P, obs, d1, nt=4)
    out = _check_gradient(M, P, dc)
    # halving dt shrinks the optimise-then-discretise gap of the kappa / rho gradient ~4x
    M2 = OracleModel(P, obs, d1, nt=8)
    _, _, gk2, gr2, _ = M2.grad(P["c0"])
    e = 1e-4
    fdk2 = (M2.J(P["c0"], ks=K0 * (1 + e)) - M2.J(P["c0"], ks=K0 * (1 - e))) / (2 * e) / K0
    fdr2 = (M2.J(P["c0"], rs=R0 * (1 + e)) - M2.J(P["c0"], rs=R0 * (1 - e))) / (2 * e) / R0
    gk, fdk = out["fd_kappa"]
    gr, fdr = out["fd_rho"]
    gap1 = (abs(gk - fdk) / abs(fdk), abs(gr - fdr) / abs(fdr))
    gap2 = (abs(gk2 - fdk2) / abs(fdk2), abs(gr2 - fdr2) / abs(fdr2))
    assert gap2[0] < gap1[0] / 2.5 and gap2[1] < gap1[1] / 2.5, (gap1, gap2)


def _hessian_checks(grad, hess, P, dc, dtype, tol_sym, tol_fd):
    """grad(c0) -> g_c0 field; hess(x) -> H x field (p-block, no diffusivity inversion).  The data
    are d1 = O c(T; c0), so the residual vanishes and Gauss-Newton IS the Hessian of the continuous
    problem: central differences of the gradient reproduce H x (computeFDHessian) and
    <Hx,y> = <x,Hy>, both up to the O(dt) linearisation-point gap described in the module
    docstring.  -> (relative asymmetry, relative FD error)."""
    c0 = P["c0"]
    x = dc
    y = (Cs.smooth_field((N, N, N), dtype, 21, -1.0, 1.0) * (0.2 * float(c0.max()))).astype(dtype)
    Hx, Hy = hess(x), hess(y)
    a = float(np.sum(Hx.astype(np.float64) * y))
    b = float(np.sum(Hy.astype(np.float64) * x))
    assert abs(a - b) <= tol_sym * max(abs(a), abs(b)), ("Gauss-Newton Hessian is not symmetric", a, b)
    e = 1e-3
    fd = (grad((c0 + e * x).astype(dtype)).astype(np.float64) - grad((c0 - e * x).astype(dtype))) / (2 * e)
    err = Cs.rel(Hx, fd)
    assert err < tol_fd, ("H x differs from the central difference of the gradient", err)
    xHx = float(np.sum(Hx.astype(np.float64) * x))
    assert xHx > 0, "Gauss-Newton Hessian must be positive"
    return abs(a - b) / max(abs(a), abs(b)), err


def _oracle_hessian_gaps(nt):
    dtype = np.float64
    P, obs, _, dc = _problem(dtype)
    M = OracleModel(P, obs, None, nt=nt)
    pde, D = M._ops(K0, R0)
    cT = pde.solve_state(P["c0"], 0)
    d1 = D._O(cT)                      # zero residual at c0
    ref = D.evaluate_objective_and_gradient(P["c0"], d1)   # histories + stale p_[nt] for the Hessian
    assert ref["mismatch"] < 1e-25

    def grad(c0):
        return D.evaluate_objective_and_gradient(c0, d1)["g_c0"]

    def hess(x):
        D.evaluate_objective_and_gradient(P["c0"], d1)     # re-linearise about c0
        return D.evaluate_hessian(x, False)[0]

    return _hessian_checks(grad, hess, P, dc, dtype, tol_sym=2e-4, tol_fd=2e-4)


def test_oracle_hessian_symmetry_and_fd():
    a4, f4 = _oracle_hessian_gaps(4)
    a8, f8 = _oracle_hessian_gaps(8)
    # first order in dt: both gaps halve with the time step
    assert 1.6 < a4 / a8 < 2.4 and 1.6 < f4 / f8 < 2.4, (a4, a8, f4, f8)


# ================================================================= GPU: C ABI ====
@pytest.mark.gpu
def test_gpu_gradient_taylor(cuda_lib, torch_cuda):
    P, obs, d1, dc = _problem()
    B = Cs.TorchBackend(cuda_lib)
    M = AbiModel(B, P, obs, d1, nt=4)
    try:
        out = _check_gradient(M, P, dc)
        # and it is the oracle's gradient
        Mo = OracleModel(P, obs, d1, nt=4)
        J0, g_c0, gk, gr, _ = Mo.grad(P["c0"])
        Jg, g_c0g, gkg, grg, _ = M.grad(P["c0"])
        assert abs(Jg - J0) < 1e-10 * abs(J0)
        assert Cs.rel(g_c0g, g_c0) < 1e-9
        assert abs(gkg - gk) < 1e-8 * abs(gk) and abs(grg - gr) < 1e-8 * abs(gr)
    finally:
        M.close()


@pytest.mark.gpu
def test_gpu_hessian_symmetry_and_fd(cuda_lib, torch_cuda):
    dtype = np.float64
    P, obs, _, dc = _problem(dtype)
    B = Cs.TorchBackend(cuda_lib)
    nt = 4
    M = AbiModel(B, P, obs, np.zeros((N, N, N), dtype), nt=nt)
    try:
        h, d = M.h, M.dev
        M._coef(K0, R0)
        cT = B.empty((N, N, N), dtype)
        h.solve_state(B.put(P["c0"]), cT, 0)
        d1 = (B.get(cT) * obs).astype(dtype)
        M.d1_d = B.put(d1)
        y = B.empty((N, N, N), dtype)

        def grad(c0):
            return M.grad(c0)[1]

        def hess(x):
            r = h.objective_gradient(B.put(P["c0"]), M.d1_d, d["wm"], d["gm"], d["csf"], obs=M.obs_d, beta=BETA)
            assert r["mismatch"] < 1e-25
            h.hessian_matvec(B.put(x), y, d["wm"], d["gm"], d["csf"], obs=M.obs_d, beta=BETA, diffusivity_inversion=False)
            return B.get(y).copy()

        asym, fd = _hessian_checks(grad, hess, P, dc, dtype, tol_sym=2e-4, tol_fd=2e-4)
        # the oracle's numbers for the same problem (tests/test_taylor.py::test_oracle_hessian_symmetry_and_fd)
        a_ref, f_ref = _oracle_hessian_gaps(nt)
        assert abs(asym - a_ref) < 1e-3 * a_ref and abs(fd - f_ref) < 1e-3 * f_ref, (asym, a_ref, fd, f_ref)
    finally:
        M.close()
