"""CPU oracle for the GLIA reaction-diffusion forward/adjoint hot path.

TEST INFRASTRUCTURE ONLY.  This module restates, in NumPy/SciPy, the arithmetic
of the reference's *CPU* path (PETSc KSPCG + AccFFT/FFTW conventions) in the
reference's own operation order (3-D FFT gradient / divergence, 12 3-D FFTs per
PCG iteration).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product (``glia_b200``) never does.

Pinning: the reference cannot be compiled in this image
(no MPI / PETSc / AccFFT / FFTW / PnetCDF), so the oracle is pinned against the
reference's own known-answer tests instead:
  K1  src/test/pdesolver.cpp:7-56    ||c||_2 == 2.0487, ksp_itr_ == 5
  K2  src/test/simulator.cpp:94-95   ||c0||_2 == 4.09351   (atlas.nc)
  K3  src/test/simulator.cpp:41-42   ||c0||_2 == 22.0161   (sinusoid.nc)
(tests/test_oracle_pins.py).  A pure model-1 forward c(T), the adjoint alpha(t)
and the kappa/rho gradient have no reference pin that runs without TAO:
for those three quantities parity is "unpinned" and the oracle itself, anchored
by K1-K3, is the reference.

Third-party arithmetic restated here (not vendored under /root/reference):
  * AccFFT accfft_grad / accfft_divergence / accfft_execute_r2c,c2r -- convention
    taken from the only in-repo statement, src/cuda/SpectralOperators.cu:54-127
    and src/grad/SpectralOperators.cpp:100-261.
  * PETSc (3.7..3.11, doc/install.md:11) KSPCG with left preconditioning, the
    preconditioned residual norm and KSPConvergedDefault with a non-zero initial
    guess -- call sites src/pde/DiffusionSolver.cpp:19-44,241-243.

All citations are relative to /root/reference.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.fft as sfft

_WORKERS = int(os.environ.get("GLIA_ORACLE_WORKERS", os.cpu_count() or 1))


def set_workers(n: int) -> None:
    global _WORKERS
    _WORKERS = int(n)


def get_workers() -> int:
    return _WORKERS


# --------------------------------------------------------------------------
# L0  SpectralOperators
# --------------------------------------------------------------------------
def wavenumbers(n: int) -> np.ndarray:
    """Signed integer wavenumbers with the Nyquist entry zeroed (trap T1).

    src/cuda/SpectralOperators.cu:62-75, src/pde/DiffusionSolver.cpp:151-162.
    """
    w = np.arange(n, dtype=np.int64)
    w[w > n // 2] -= n
    w[n // 2] = 0
    return w


def fft_r2c(f: np.ndarray) -> np.ndarray:
    """Unnormalised 3-D real-to-complex FFT, half spectrum along z.
    src/grad/SpectralOperators.cpp:68-83."""
    return sfft.rfftn(f, workers=_WORKERS)


def fft_c2r(fhat: np.ndarray, shape) -> np.ndarray:
    """Unnormalised inverse (forward . inverse = N . id).
    src/grad/SpectralOperators.cpp:85-98."""
    return sfft.irfftn(fhat, s=shape, norm="forward", workers=_WORKERS)


def _mult_wave(fhat: np.ndarray, axis: int, shape) -> np.ndarray:
    """w_f = (i * w_d / N) * f_hat, out of place.
    src/cuda/SpectralOperators.cu:54-127 (factor = 1/(n0 n1 n2))."""
    rdt = fhat.real.dtype
    n = shape[axis]
    factor = rdt.type(1.0 / (shape[0] * shape[1] * shape[2]))
    w = wavenumbers(n)
    if axis == 2:
        w = w[: n // 2 + 1]
    fw = (factor * w.astype(rdt)).astype(rdt)          # factor * wx, ScalarType
    sh = [1, 1, 1]
    sh[axis] = fw.shape[0]
    fw = fw.reshape(sh)
    out = np.empty_like(fhat)
    out.real = -fw * fhat.imag
    out.imag = fw * fhat.real
    return out


def gradient(f: np.ndarray, xyz=(1, 1, 1)):
    """SpectralOperators::computeGradient -- 1 R2C + one multiply + C2R per
    requested component.  src/grad/SpectralOperators.cpp:100-177."""
    shape = f.shape
    fhat = fft_r2c(f)
    out = []
    for d in range(3):
        if xyz[d]:
            out.append(fft_c2r(_mult_wave(fhat, d, shape), shape))
        else:
            out.append(None)
    return out


def divergence(dx: np.ndarray, dy: np.ndarray, dz: np.ndarray) -> np.ndarray:
    """SpectralOperators::computeDivergence.  The real-space sum is taken in the
    reference's order: x, then z, then y.  src/grad/SpectralOperators.cpp:179-261."""
    shape = dx.shape
    div = fft_c2r(_mult_wave(fft_r2c(dx), 0, shape), shape)
    div = div + fft_c2r(_mult_wave(fft_r2c(dz), 2, shape), shape)
    div = div + fft_c2r(_mult_wave(fft_r2c(dy), 1, shape), shape)
    return div


def weierstrass_smoother(c: np.ndarray, sigma: float) -> np.ndarray:
    """SpectralOperators::weierstrassSmoother: periodic Gaussian built as a sum of
    8 images, normalised by sum(f)*h^3, applied as F^-1[F(f) F(c)] h^3 / N.
    src/grad/SpectralOperators.cpp:295-381 (no-op for sigma == 0, :272-274)."""
    if sigma == 0:
        return c.copy()
    dt = c.dtype
    n0, n1, n2 = c.shape
    twopi = dt.type(2.0 * np.pi)
    h = [dt.type(twopi / n) for n in (n0, n1, n2)]
    s2 = dt.type(sigma)

    def g1(n, hh):
        X = (np.arange(n).astype(dt) * hh).astype(dt)
        Xp = (X - twopi).astype(dt)
        return X, Xp

    X, Xp = g1(n0, h[0])
    Y, Yp = g1(n1, h[1])
    Z, Zp = g1(n2, h[2])

    def e(a, b, cc):
        A = (-a * a)[:, None, None]
        B = (b * b)[None, :, None]
        C = (cc * cc)[None, None, :]
        with np.errstate(over="ignore", invalid="ignore"):
            return np.exp(((A - B - C) / s2 / s2 / dt.type(2.0)).astype(dt)).astype(dt)

    f = e(X, Y, Z) + e(Xp, Yp, Zp)
    f = f + (e(Xp, Y, Z) + e(X, Yp, Z))
    f = f + (e(X, Y, Zp) + e(Xp, Yp, Z))
    f = f + (e(Xp, Y, Zp) + e(X, Yp, Zp))
    f = f.astype(dt)
    f[f != f] = 0
    sum_f = dt.type(f.sum(dtype=np.float64))
    norm = dt.type(1.0) / (sum_f * h[0] * h[1] * h[2])
    f = (f * norm).astype(dt)
    factor = dt.type(1.0 / (n0 * n1 * n2))
    fh = fft_r2c(f)
    ch = fft_r2c(c)
    fh = fh * (ch * (factor * h[0] * h[1] * h[2]))
    return fft_c2r(fh.astype(ch.dtype), c.shape).astype(dt)


# --------------------------------------------------------------------------
# L1  DiffCoef / ReacCoef
# --------------------------------------------------------------------------
class DiffCoef:
    """Isotropic k(x) with kxx=kyy=kzz and kxy=kxz=kyz=0 (src/mat/DiffCoef.cpp:77-131)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.kxx = np.zeros(shape, dtype)
        self.k_scale = self.dtype.type(1e-2)
        self.kxx_avg = self.kyy_avg = self.kzz_avg = self.dtype.type(0)
        self.kxy_avg = self.kxz_avg = self.kyz_avg = self.dtype.type(0)
        self.ktilde = None          # temp_[7], setSecondaryCoefficients

    def _avg(self, filter_sum):
        t = self.dtype.type
        s = t(self.kxx.sum(dtype=np.float64))
        fa = t(filter_sum)
        self.kxx_avg = self.kyy_avg = self.kzz_avg = t(s * (t(1.0) / fa))
        self.kxy_avg = self.kxz_avg = self.kyz_avg = t(0)

    def set_values(self, k_scale, k_gm_wm, k_glm_wm, wm, gm, csf, filt):
        """DiffCoef::setValues (src/mat/DiffCoef.cpp:77-131): negative ratios clamp
        to 0; averages normalised by the brain-mask voxel count (trap T5)."""
        t = self.dtype.type
        self.k_scale = t(k_scale)
        dk_gm = t(k_scale) * t(k_gm_wm)
        dk_wm = t(k_scale)
        dk_glm = t(k_scale) * t(k_glm_wm)
        dk_gm = t(0) if dk_gm <= 0 else dk_gm
        dk_glm = t(0) if dk_glm <= 0 else dk_glm
        k = np.zeros(self.shape, self.dtype)
        k = k + dk_gm * gm
        k = k + dk_wm * wm
        k = k + dk_glm * csf
        self.kxx = k.astype(self.dtype)
        self._avg(filt.sum(dtype=np.float64))

    def set_values_sinusoidal(self, scale):
        """DiffCoef::setValuesSinusoidal (src/mat/DiffCoef.cpp:134-189)."""
        t = self.dtype.type
        n0, n1, n2 = self.shape
        self.k_scale = t(scale)
        freq = 4.0
        X = np.sin(freq * 2.0 * np.pi / n0 * np.arange(n0))[:, None, None]
        Y = np.sin(freq * 2.0 * np.pi / n1 * np.arange(n1))[None, :, None]
        Z = np.sin(freq * 2.0 * np.pi / n2 * np.arange(n2))[None, None, :]
        self.kxx = (float(t(scale)) * (0.5 + 0.5 * X * Y * Z)).astype(self.dtype)
        self._avg(n0 * n1 * n2)

    def set_secondary(self, k1, k2, k3, wm, gm, csf, nk=1, k_gm_wm=0.0, k_glm_wm=0.0):
        """DiffCoef::setSecondaryCoefficients (src/mat/DiffCoef.cpp:44-59)."""
        t = self.dtype.type
        k1 = t(k1)
        k2 = t(k_gm_wm) * k1 if nk == 1 else t(k2)
        k3 = t(k_glm_wm) * k1 if nk == 1 else t(k3)
        kt = (wm * k1).astype(self.dtype)
        kt = kt + k2 * gm
        kt = kt + k3 * csf
        self.ktilde = kt.astype(self.dtype)

    def apply_D(self, c, secondary=False):
        """DiffCoef::applyD / applyDWithSecondaryCoeffs: grad -> K. -> div.
        src/mat/DiffCoef.cpp:249-300 (off-diagonals identically 0, :94-100)."""
        k = self.ktilde if secondary else self.kxx
        gx, gy, gz = gradient(c)
        return divergence(k * gx, k * gy, k * gz)


def reac_coef(rho_scale, r_gm_wm, r_glm_wm, wm, gm, csf):
    """ReacCoef::setValues (src/mat/ReacCoef.cpp:13-38)."""
    t = wm.dtype.type
    dr_gm = t(rho_scale) * t(r_gm_wm)
    dr_wm = t(rho_scale)
    dr_glm = t(rho_scale) * t(r_glm_wm)
    dr_gm = t(0) if dr_gm <= 0 else dr_gm
    dr_glm = t(0) if dr_glm <= 0 else dr_glm
    rho = np.zeros_like(wm)
    rho = rho + dr_gm * gm
    rho = rho + dr_wm * wm
    rho = rho + dr_glm * csf
    return rho.astype(wm.dtype)


def update_reac_diff_mass_effect(bg, gm, vt, csf, rho_scale, k_scale, gm_r_scale, gm_k_scale):
    """PdeOperatorsMassEffect::updateReacAndDiffCoefficients, CPU branch
    (src/pde/PdeOperatorsMassEffect.cpp:116-123): returns (rho, kxx).  The averages
    kxx_avg_ are NOT touched there -- the caller keeps its DiffCoef's old ones."""
    t = bg.dtype.type
    tmp = (t(1) - (((bg + t(gm_r_scale) * gm) + vt) + csf)).astype(bg.dtype)
    rho = (np.where(tmp < 0, t(0), tmp) * t(rho_scale)).astype(bg.dtype)
    tmp = (t(1) - (((bg + t(gm_k_scale) * gm) + vt) + csf)).astype(bg.dtype)
    kxx = (np.where(tmp < 0, t(0), tmp) * t(k_scale)).astype(bg.dtype)
    return rho, kxx


def mass_effect_rd_steps(c0, tissue_seq, k: "DiffCoef", solver: "DiffusionSolver", rho_scale, k_scale,
                         gm_r_scale, gm_k_scale, dt):
    """The reaction-diffusion part of PdeOperatorsMassEffect::solveState's time loop
    (src/pde/PdeOperatorsMassEffect.cpp:578-631): per step refresh rho, k from the
    current tissue maps; precFactor(); [advection -- not restated: `tissue_seq[i]` is
    the (bg, gm, vt, csf) the step sees]; diff_solver_->solve(c, dt) with the FULL dt;
    reaction(0, i) with the full dt (order-1 splitting).  Returns (c, [ksp its])."""
    c = c0.copy()
    its = []
    for bg, gm, vt, csf in tissue_seq:
        rho, kxx = update_reac_diff_mass_effect(bg, gm, vt, csf, rho_scale, k_scale, gm_r_scale, gm_k_scale)
        k.kxx = kxx
        solver.prec_factor()
        c = solver.solve(c, dt)
        its.append(solver.ksp_itr)
        c = reaction_nonlinear(c, rho, dt)
    return c, its


# --------------------------------------------------------------------------
# L2  DiffusionSolver  (Crank-Nicolson, PETSc-CG semantics)
# --------------------------------------------------------------------------
class DiffusionSolver:
    """src/pde/DiffusionSolver.cpp:5-250.

    ``dt_ctx`` is *state* (trap T2): initialised from params->tu_->dt_
    (default 0.5, include/Parameters.h:166), overwritten by every solve();
    precFactor() uses whatever value it holds at the time it is called.
    """

    RTOL = 1e-6
    ABSTOL = 1e-50
    DTOL = 1e4
    MAXIT = 5000

    def __init__(self, k: DiffCoef, dt_ctx=0.5):
        self.k = k
        self.dtype = k.dtype
        self.dt_ctx = self.dtype.type(dt_ctx)
        self.ksp_itr = 0
        self.precfactor = None
        self.prec_factor()

    def prec_factor(self):
        """DiffusionSolver::precFactor, CPU branch (:143-172).  The 0.25 / 2.0 / 1
        literals promote the expression to double before it is stored as
        ScalarType; the squares kxx_avg*wx*wx are ScalarType products."""
        t = self.dtype.type
        n0, n1, n2 = self.k.shape
        factor = t(1.0 / (n0 * n1 * n2))
        wx = wavenumbers(n0).astype(self.dtype)[:, None, None]
        wy = wavenumbers(n1).astype(self.dtype)[None, :, None]
        wz = wavenumbers(n2)[: n2 // 2 + 1].astype(self.dtype)[None, None, :]
        k = self.k
        txx = ((k.kxx_avg * wx).astype(self.dtype) * wx).astype(self.dtype).astype(np.float64)
        tyy = ((k.kyy_avg * wy).astype(self.dtype) * wy).astype(self.dtype).astype(np.float64)
        tzz = ((k.kzz_avg * wz).astype(self.dtype) * wz).astype(self.dtype).astype(np.float64)
        txy = 2.0 * float(k.kxy_avg) * wx.astype(np.float64) * wy
        txz = 2.0 * float(k.kxz_avg) * wx.astype(np.float64) * wz
        tyz = 2.0 * float(k.kyz_avg) * wy.astype(np.float64) * wz
        s = ((((txx + txy) + txz) + tyz) + tyy) + tzz
        pf = (1 + 0.25 * float(self.dt_ctx) * s).astype(self.dtype)
        out = np.empty_like(pf)
        z = pf == 0
        out[z] = factor
        out[~z] = factor / pf[~z]
        self.precfactor = out.astype(self.dtype)

    def operator_A(self, x):
        """y = x - (dt/2) D x  (:97-117); alph = -1/2*dt in ScalarType."""
        alph = self.dtype.type(-1.0 / 2.0 * float(self.dt_ctx))
        return (x + alph * self.k.apply_D(x)).astype(self.dtype)

    def apply_pc(self, x):
        """y = F^-1[ P_hat . F x ]  (:182-215)."""
        xh = fft_r2c(x)
        xh = xh * self.precfactor
        return fft_c2r(xh, x.shape).astype(self.dtype)

    @staticmethod
    def _dot(a, b):
        return float(np.dot(a.ravel().astype(np.float64), b.ravel().astype(np.float64)))

    def solve(self, c, dt):
        """DiffusionSolver::solve (:217-250) + KSPSolve_CG / KSPConvergedDefault
        (PETSc 3.11 src/ksp/ksp/impls/cg/cg.c, src/ksp/ksp/interface/iterativ.c).
        Returns the new c; sets ksp_itr."""
        t = self.dtype.type
        self.dt_ctx = t(dt)
        if self.k.k_scale == 0:
            self.ksp_itr = 0
            return c
        alph = t(1.0 / 2.0 * float(self.dt_ctx))
        b = (c + alph * self.k.apply_D(c)).astype(self.dtype)

        x = c.copy()
        r = (b - self.operator_A(x)).astype(self.dtype)
        z = self.apply_pc(r)
        dp = np.sqrt(self._dot(z, z))
        # KSPConvergedDefault at n == 0 with a non-zero initial guess
        snorm = np.sqrt(self._dot(*(lambda zb: (zb, zb))(self.apply_pc(b))))
        if snorm == 0:
            snorm = dp
        rnorm0 = snorm
        ttol = max(self.RTOL * rnorm0, self.ABSTOL)
        its = 0
        if dp <= ttol:
            self.ksp_itr = 0
            return x
        beta = 0.0
        betaold = 1.0
        p = None
        while its < self.MAXIT:
            if its == 0:
                beta = self._dot(z, r)
                if beta == 0.0:
                    break
                p = z.copy()
            else:
                bb = t(beta / betaold)
                p = (z + bb * p).astype(self.dtype)
            w = self.operator_A(p)
            dpi = self._dot(p, w)
            betaold = beta
            a = t(beta / dpi)
            x = (x + a * p).astype(self.dtype)
            r = (r - a * w).astype(self.dtype)
            z = self.apply_pc(r)
            dp = np.sqrt(self._dot(z, z))
            its += 1
            if dp <= ttol:
                break
            if dp >= self.DTOL * rnorm0:
                raise RuntimeError("KSP_DIVERGED_DTOL")
            beta = self._dot(z, r)
        self.ksp_itr = its
        return x


# --------------------------------------------------------------------------
# L2a  PdeOperatorsRD
# --------------------------------------------------------------------------
def reaction_nonlinear(c, rho, dt):
    """PdeOperatorsRD::reaction, linearized == 0 (src/pde/PdeOperators.cpp:161-169).
    `1.0 - c` and `alph*factor + 1.0` are evaluated in double (trap T6)."""
    t = c.dtype
    factor = np.exp((rho * t.type(dt)).astype(t)).astype(t)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        alph = (c.astype(np.float64) / (1.0 - c.astype(np.float64))).astype(t)
        af = (alph * factor).astype(t)
        out = (af.astype(np.float64) / (af.astype(np.float64) + 1.0)).astype(t)
    out[np.isinf(alph)] = 1.0
    return out


def reaction_linearized(u, c_lin, rho, dt):
    """u <- u e^{rho dt} / (c e^{rho dt} + 1 - c)^2
    (src/pde/PdeOperators.cpp:170-176 and :353-357)."""
    t = u.dtype
    factor = np.exp((rho * t.type(dt)).astype(t)).astype(t)
    cf = (c_lin * factor).astype(t)
    alph = ((cf.astype(np.float64) + 1.0) - c_lin.astype(np.float64)).astype(t)
    uf = (u * factor).astype(t)
    return (uf / (alph * alph).astype(t)).astype(t)


class PdeOperatorsRD:
    """Strang-split state / adjoint / incremental drivers with time histories.
    src/pde/PdeOperators.cpp:140-420."""

    def __init__(self, k: DiffCoef, rho: np.ndarray, nt: int, dt: float,
                 dt_ctx=None, adjoint_store=True, order=2):
        self.order = int(order)     # params->tu_->order_ (PdeOperators.cpp:271-290, 389-398)
        self.k = k
        self.rho = rho
        self.nt = int(nt)
        self.dt = float(dt)
        self.dtype = k.dtype
        self.adjoint_store = adjoint_store
        self.diff = DiffusionSolver(k, self.dt if dt_ctx is None else dt_ctx)
        shape = k.shape
        self.c_ = [np.zeros(shape, self.dtype) for _ in range(nt + 1)]
        self.p_ = [np.zeros(shape, self.dtype) for _ in range(nt + 1)]
        self.c_half_ = [np.zeros(shape, self.dtype) for _ in range(nt)]
        self.ksp_state = 0
        self.ksp_adj = 0
        self.ksp_trace = []

    def solve_incremental(self, c_tilde, i, mode, dt_half):
        """src/pde/PdeOperators.cpp:192-233 (weights 1.5/0.5 exactly as coded)."""
        t = self.dtype.type
        tmp = (self.c_[i] + self.c_[i + 1]).astype(self.dtype)
        tmp = (tmp * t(0.5)).astype(self.dtype)
        tmp = (tmp + (self.c_[i] if mode == 1 else self.c_[i + 1])).astype(self.dtype)
        tmp = self.k.apply_D(tmp, secondary=True).astype(self.dtype)
        return (c_tilde + t(dt_half / 2) * tmp).astype(self.dtype)

    def solve_state(self, c0, linearized=0):
        """src/pde/PdeOperators.cpp:235-316."""
        dt, nt = self.dt, self.nt
        c = c0.astype(self.dtype).copy()
        if linearized == 0:
            self.c_[0] = c.copy()
        self.ksp_state = 0
        for i in range(nt):
            if linearized == 2:
                c = self.solve_incremental(c, i, 1, dt / 2)
            c = self.diff.solve(c, dt / 2.0 if self.order == 2 else dt)
            self.ksp_state += self.diff.ksp_itr
            self.ksp_trace.append(self.diff.ksp_itr)
            if linearized == 0 and self.adjoint_store:
                self.c_half_[i] = c.copy()
            if linearized == 0:
                c = reaction_nonlinear(c, self.rho, dt)
            else:
                c = reaction_linearized(c, self.c_[i], self.rho, dt)
            if self.order == 2:
                c = self.diff.solve(c, dt / 2.0)
                self.ksp_state += self.diff.ksp_itr
                self.ksp_trace.append(self.diff.ksp_itr)
                if linearized == 2:
                    c = self.solve_incremental(c, i, 2, dt / 2)
            if linearized == 0:
                self.c_[i + 1] = c.copy()
        return c

    def solve_adjoint(self, pT, linearized=1):
        """src/pde/PdeOperators.cpp:318-420.  p_[nt] is only written when
        linearized == 1 (trap T4)."""
        dt, nt = self.dt, self.nt
        p = pT.astype(self.dtype).copy()
        if linearized == 1:
            self.p_[nt] = p.copy()
        self.ksp_adj = 0
        for i in range(nt):
            p = self.diff.solve(p, dt / 2.0 if self.order == 2 else dt)
            self.ksp_adj += self.diff.ksp_itr
            self.ksp_trace.append(self.diff.ksp_itr)
            it = nt - i - 1
            if self.adjoint_store:
                c_lin = self.c_half_[it]
            else:
                c_lin = self.diff.solve(self.c_[it].copy(), dt / 2.0)
                self.ksp_adj += self.diff.ksp_itr
            p = reaction_linearized(p, c_lin, self.rho, dt)
            if self.order == 2:
                p = self.diff.solve(p, dt / 2.0)
                self.ksp_adj += self.diff.ksp_itr
                self.ksp_trace.append(self.diff.ksp_itr)
            self.p_[it] = p.copy()
        return p


# --------------------------------------------------------------------------
# L2b  gradient time integrals
# --------------------------------------------------------------------------
def grad_integrals(pde: PdeOperatorsRD):
    """The two trapezoid time integrals of DerivativeOperators::gradDiffusion /
    gradReaction (src/grad/DerivativeOperators.cpp:189-321), accumulated in
    ScalarType exactly in the reference's order.  Returns (T_kappa, T_rho)."""
    t = pde.dtype.type
    nt = pde.nt
    Tk = np.zeros(pde.k.shape, pde.dtype)
    Tr = np.zeros(pde.k.shape, pde.dtype)
    for i in range(nt + 1):
        wgt = 0.5 if (i == 0 or i == nt) else 1.0
        cx, cy, cz = gradient(pde.c_[i])
        px, py, pz = gradient(pde.p_[i])
        w0 = (cx * px).astype(pde.dtype)
        w0 = (w0 + (cy * py).astype(pde.dtype)).astype(pde.dtype)
        w0 = (w0 + (cz * pz).astype(pde.dtype)).astype(pde.dtype)
        Tk = (Tk + t(pde.dt * wgt) * w0).astype(pde.dtype)
        r0 = (pde.c_[i] * pde.c_[i]).astype(pde.dtype)
        r0 = (r0 - pde.c_[i]).astype(pde.dtype)
        r0 = (pde.p_[i] * r0).astype(pde.dtype)
        Tr = (Tr + t(pde.dt * wgt) * r0).astype(pde.dtype)
    return Tk, Tr


def grad_kappa_rho(pde: PdeOperatorsRD, wm, gm, csf):
    """Returns the six scalars h^3 <m, T_kappa>, h^3 <m, T_rho> for m = wm, gm, csf
    (the caller combines them per nk / nr; src/grad/DerivativeOperators.cpp:231-249,
    293-313)."""
    n0, n1, n2 = pde.k.shape
    leb = (2 * np.pi / n0) * (2 * np.pi / n1) * (2 * np.pi / n2)
    Tk, Tr = grad_integrals(pde)
    d = DiffusionSolver._dot
    return np.array([leb * d(wm, Tk), leb * d(gm, Tk), leb * d(csf, Tk),
                     leb * d(wm, Tr), leb * d(gm, Tr), leb * d(csf, Tr)])


class DerivativeOperatorsRD:
    """Field-space restatement of DerivativeOperatorsRD::{evaluateObjective,
    evaluateObjectiveAndGradient, evaluateHessian}
    (src/grad/DerivativeOperatorsRD.cpp:7-438) with the Phi basis factored out: the
    initial condition c(0) (resp. c~(0) = Phi p~) is an input field and the p-block of the
    gradient / Hessian product is returned as the field g with g_p = Phi^T g.
    Obs::apply / applyT are pointwise products with the observation mask
    (src/mat/Obs.cpp:75-140); obs = None is O = I.  L2 regularisation
    (:36-39).  The kappa / rho blocks are the six scalars of grad_kappa_rho."""

    def __init__(self, pde: PdeOperatorsRD, wm, gm, csf, obs=None, beta=0.0, d0=None, obs0=None):
        self.pde, self.wm, self.gm, self.csf = pde, wm, gm, csf
        self.obs = obs
        # two_time_points_ (DerivativeOperatorsRD.cpp:30-34, 149-153, 216-222): data and mask at t = 0
        self.d0, self.obs0 = d0, obs0
        self.beta = float(beta)
        n0, n1, n2 = pde.k.shape
        self.leb = (2 * np.pi / n0) * (2 * np.pi / n1) * (2 * np.pi / n2)
        self.dtype = pde.dtype

    def _O(self, x):
        return x if self.obs is None else (x * self.obs).astype(self.dtype)

    def _terminal(self, cT, d1):
        """temp = O c(1) - d1 ; p_t = -O^T temp   (:27, :81-84)."""
        temp = self._O(cT)
        if d1 is not None:
            temp = (temp - d1).astype(self.dtype)
        pT = (self._O(temp) * self.dtype.type(-1.0)).astype(self.dtype)
        return temp, pT

    def evaluate_objective_and_gradient(self, c0, d1):
        """-> dict(J, mismatch, reg, cT, p0, g_c0, g6); J = h^3/2 ||O c(1) - d1||^2 + beta/2 h^3 ||c0||^2,
        g_c0 = -h^3 (alpha(0) - beta c0), g6 = grad_kappa_rho (:130-226)."""
        d = DiffusionSolver._dot
        pde = self.pde
        m0, t0 = 0.0, None
        if self.d0 is not None:        # Oc(0) - d0 with the t = 0 mask (:149-153)
            t0 = c0 if self.obs0 is None else (c0 * self.obs0).astype(self.dtype)
            t0 = (t0 - self.d0).astype(self.dtype)
            m0 = d(t0, t0)
        cT = pde.solve_state(c0, 0)
        temp, pT = self._terminal(cT, d1)
        m1 = d(temp, temp)
        reg = 0.5 * self.beta * d(c0, c0) * self.leb
        J = self.leb * 0.5 * (m1 + m0) + reg
        p0 = pde.solve_adjoint(pT, 1)
        t = self.dtype.type
        g_c0 = ((p0 - t(self.beta) * c0).astype(self.dtype) * t(-self.leb)).astype(self.dtype)
        if t0 is not None:             # + h^3 O0^T(O0 c0 - d0)  (:216-222, Phi^T factored out)
            q = t0 if self.obs0 is None else (t0 * self.obs0).astype(self.dtype)
            g_c0 = (g_c0 + t(self.leb) * q).astype(self.dtype)
        g6 = grad_kappa_rho(pde, self.wm, self.gm, self.csf)
        return dict(J=J, mismatch=self.leb * 0.5 * m1, mismatch0=self.leb * 0.5 * m0, reg=reg, cT=cT, p0=p0,
                    g_c0=g_c0, g6=g6, its=(pde.ksp_state, pde.ksp_adj))

    def evaluate_hessian(self, c0_tilde, diffusivity_inversion=False):
        """Gauss-Newton Hessian product (:229-438).  Needs the state history of a previous
        evaluate_objective_and_gradient (c_, c_half_; p_[nt] is the stale gradient adjoint, trap
        T4).  With diffusivity_inversion the secondary coefficients (k.set_secondary) carry
        k~.  -> (y_c0 field, hk[6] = h^3 <wm|gm|csf, T_kp>, h^3 <wm|gm|csf, T_kk>)."""
        if self.d0 is not None:
            raise NotImplementedError("Hessian currently not implemented for two-snapshot scenario")  # :234
        pde = self.pde
        t = self.dtype.type
        d = DiffusionSolver._dot
        its = []
        cTt = pde.solve_state(c0_tilde, 1)
        its.append(pde.ksp_state)
        _, pT = self._terminal(cTt, None)
        p0 = pde.solve_adjoint(pT, 2)
        its.append(pde.ksp_adj)
        y = ((t(self.beta) * c0_tilde - p0).astype(self.dtype) * t(self.leb)).astype(self.dtype)
        hk = np.zeros(6)
        if diffusivity_inversion:
            Tk, _ = grad_integrals(pde)
            hk[0:3] = [self.leb * d(m, Tk) for m in (self.wm, self.gm, self.csf)]
            cTt = pde.solve_state(np.zeros_like(c0_tilde), 2)
            its.append(pde.ksp_state)
            _, pT = self._terminal(cTt, None)
            p0 = pde.solve_adjoint(pT, 2)
            its.append(pde.ksp_adj)
            y = (y + t(-self.leb) * p0).astype(self.dtype)
            Tk, _ = grad_integrals(pde)
            hk[3:6] = [self.leb * d(m, Tk) for m in (self.wm, self.gm, self.csf)]
        return y, hk, its


# --------------------------------------------------------------------------
# fixtures shared by tests / bench
# --------------------------------------------------------------------------
def test_gaussian(n: int, dtype) -> np.ndarray:
    """createTestFunction: exp(-r^2/R^2), R = sqrt(2) 2pi/64, centre (pi,pi,pi).
    src/test/helper.cpp:19-49."""
    dt = np.dtype(dtype)
    R = dt.type(np.sqrt(2.0) * (2 * np.pi) / 64)
    h = dt.type(2 * np.pi / n)
    # dx = h*X - M_PI : ScalarType * int64 -> ScalarType, then minus a double
    d = (h * np.arange(n).astype(dt)).astype(dt).astype(np.float64) - np.pi
    d = d.astype(dt)
    r = np.sqrt((d[:, None, None] ** 2 + d[None, :, None] ** 2 + d[None, None, :] ** 2).astype(dt)).astype(dt)
    ratio = (r / R).astype(dt)
    return np.exp(-(ratio * ratio)).astype(dt)


# --------------------------------------------------------------------------
# atlas / material properties / Phi  (inputs of the path: c(0) = Phi p)
# --------------------------------------------------------------------------
def split_segmentation(seg, labels=(6, 5, 7, 8), dtype=np.float64):
    """splitSegmentation (src/utils/Utils.cpp:592-657): one-hot maps for
    labels = [wm, gm, vt, csf] (scripts/forward.py:17 uses wm=6, gm=5, vt=7, csf=8)."""
    wm_l, gm_l, vt_l, csf_l = labels
    one = lambda l: (seg == l).astype(dtype) if l > 0 else np.zeros(seg.shape, dtype)
    return {"wm": one(wm_l), "gm": one(gm_l), "vt": one(vt_l), "csf": one(csf_l)}


def read_atlas(maps, n, smoothing_factor=1.0, smoothing_factor_atlas=1.0):
    """SolverInterface::readAtlas (src/SolverInterface.cpp:502-570): every present tissue
    map is smoothed with sigma = smoothing_factor * 2 pi / n (order gm, wm, vt, csf)."""
    out = {}
    for key in ("gm", "wm", "vt", "csf"):
        v = maps.get(key)
        if v is None:
            out[key] = None
            continue
        dt = v.dtype.type
        sigma = dt(smoothing_factor * 2 * np.pi / n)
        out[key] = weierstrass_smoother(v, sigma) if smoothing_factor_atlas > 0 else v.copy()
    return out


def mat_prop(atlas, shape, dtype):
    """MatProp::setValuesCustom (src/mat/MatProp.cpp:135-201): absent maps are zero, clip at
    0, bg = 1 - sum, filter = (wm > 0.1 or gm > 0.1) and vt < 0.8."""
    z = lambda k: (np.zeros(shape, dtype) if atlas.get(k) is None else atlas[k].astype(dtype))
    m = {k: np.where(z(k) <= 0, 0, z(k)).astype(dtype) for k in ("gm", "wm", "vt", "csf")}
    bg = (m["gm"] + m["wm"]).astype(dtype)
    bg = (bg + m["vt"]).astype(dtype)
    bg = (bg + m["csf"]).astype(dtype)
    m["bg"] = (-(bg - dtype(1.0))).astype(dtype)
    m["filter"] = (((m["wm"] > 0.1) | (m["gm"] > 0.1)) & (m["vt"] < 0.8)).astype(dtype)
    return m


def phi_apply(p, centers, sigma_phi, filt, smoothing_factor=1.0):
    """Phi::apply in on-the-fly mode (src/mat/Phi.cpp:264-374): for every non-zero p_i,
    phi_i = truncate_{5 sigma}( W_sigma( Gaussian_i * filter ) ); out = sum_i p_i phi_i / max_i max(phi_i)."""
    dt = filt.dtype
    t = dt.type
    n0, n1, n2 = filt.shape
    twopi = t(2.0 * np.pi)
    hx, hy, hz = t(twopi / n0), t(twopi / n1), t(twopi / n2)
    sigma_smooth = t(smoothing_factor * 2.0 * np.pi / n0)
    sig = t(sigma_phi)
    R = t(np.sqrt(2.0) * float(sig))
    out = np.zeros(filt.shape, dt)
    phi_max = t(0)
    nnz = False
    X = (hx * np.arange(n0).astype(dt)).astype(dt)
    Y = (hy * np.arange(n1).astype(dt)).astype(dt)
    Z = (hz * np.arange(n2).astype(dt)).astype(dt)
    for pi, ctr in zip(p, centers):
        if pi == 0:
            continue
        nnz = True
        dx = (X - t(ctr[0]))[:, None, None]
        dy = (Y - t(ctr[1]))[None, :, None]
        dz = (Z - t(ctr[2]))[None, None, :]
        r = np.sqrt((dx * dx + dy * dy + dz * dz).astype(dt)).astype(dt)
        ratio = (r / R).astype(dt)
        phi = np.exp(-(ratio * ratio)).astype(dt)
        phi = (filt * phi).astype(dt)
        phi = weierstrass_smoother(phi, sigma_smooth)
        phi = np.where((r / sig) <= 5, phi, 0).astype(dt)
        phi_max = max(phi_max, t(phi.max()))
        out = (out + t(pi) * phi).astype(dt)
    if not nnz:
        phi_max = t(1)
    return (out * (t(1.0) / phi_max)).astype(dt)


def _phi_i(ctr, X, Y, Z, sig, R, filt, sigma_smooth):
    """One basis function of the on-the-fly mode: truncate(W(Gaussian . filter)).
    Phi::initialize (src/mat/Phi.cpp:262-320) + VecPointwiseMult(filter) + weierstrassSmoother +
    Phi::truncate (Phi.cpp:237-260)."""
    dt = filt.dtype
    t = dt.type
    dx = (X - t(ctr[0]))[:, None, None]
    dy = (Y - t(ctr[1]))[None, :, None]
    dz = (Z - t(ctr[2]))[None, None, :]
    r = np.sqrt((dx * dx + dy * dy + dz * dz).astype(dt)).astype(dt)
    ratio = (r / R).astype(dt)
    phi = np.exp(-(ratio * ratio)).astype(dt)
    phi = (filt * phi).astype(dt)
    phi = weierstrass_smoother(phi, sigma_smooth)
    return np.where((r / sig) <= 5, phi, 0).astype(dt)


def phi_apply_transpose(field, centers, sigma_phi, filt, smoothing_factor=1.0):
    """Phi::applyTranspose in on-the-fly mode (src/mat/Phi.cpp:385-434): pout_i = <phi_i, in>,
    every i (no skipping), then pout *= 1 / max_i max(phi_i).  VecDot accumulates in the working
    precision in PETSc; here it is accumulated in float64 and rounded once."""
    dt = filt.dtype
    t = dt.type
    n0, n1, n2 = filt.shape
    twopi = t(2.0 * np.pi)
    hx, hy, hz = t(twopi / n0), t(twopi / n1), t(twopi / n2)
    sigma_smooth = t(smoothing_factor * 2.0 * np.pi / n0)
    sig = t(sigma_phi)
    R = t(np.sqrt(2.0) * float(sig))
    X = (hx * np.arange(n0).astype(dt)).astype(dt)
    Y = (hy * np.arange(n1).astype(dt)).astype(dt)
    Z = (hz * np.arange(n2).astype(dt)).astype(dt)
    out = np.zeros(len(centers), dt)
    phi_max = t(0)
    for i, ctr in enumerate(centers):
        phi = _phi_i(ctr, X, Y, Z, sig, R, filt, sigma_smooth)
        phi_max = max(phi_max, t(phi.max()))
        out[i] = t(np.sum(phi.astype(np.float64) * field.astype(np.float64)))
    return (out * (t(1.0) / phi_max)).astype(dt)
