"""Multi-threaded CPU restatement of the reference's RD path, used ONLY as the timed CPU
baseline of bench.py (`cpu_baseline`, `--impl reference`).

TEST / BENCH INFRASTRUCTURE ONLY -- same rules as rd_oracle.py: the product never imports it.

It is rd_oracle.py's algorithm (reference operation order: 3-D FFT gradient and divergence,
10 3-D FFTs per applyD, 12 per PCG iteration, PETSc-CG semantics; file:line citations there)
expressed with torch CPU tensors so that the FFTs (MKL / pocketfft) and the field-sized
elementwise sweeps use every host core, which is what the reference's MPI + AccFFT build does.
tests/test_oracle_torch.py checks it against rd_oracle.py; parity claims are made against
rd_oracle.py only.
"""
from __future__ import annotations

import math

import torch


def wavenumbers(n, dtype):
    w = torch.arange(n, dtype=torch.int64)
    w = torch.where(w > n // 2, w - n, w)
    w[n // 2] = 0                      # trap T1 (src/cuda/SpectralOperators.cu:69-72)
    return w.to(dtype)


class SpectralOps:
    """src/grad/SpectralOperators.cpp:68-261."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = dtype
        n0, n1, n2 = self.shape
        f = 1.0 / (n0 * n1 * n2)
        self.iw = [
            (f * wavenumbers(n0, dtype)).reshape(n0, 1, 1),
            (f * wavenumbers(n1, dtype)).reshape(1, n1, 1),
            (f * wavenumbers(n2, dtype))[: n2 // 2 + 1].reshape(1, 1, n2 // 2 + 1),
        ]

    def r2c(self, f):
        return torch.fft.rfftn(f)

    def c2r(self, fh):
        return torch.fft.irfftn(fh, s=self.shape, norm="forward")

    def _mult(self, fh, d):
        return torch.complex(-self.iw[d] * fh.imag, self.iw[d] * fh.real)

    def gradient(self, f):
        fh = self.r2c(f)
        return [self.c2r(self._mult(fh, d)) for d in range(3)]

    def divergence(self, dx, dy, dz):
        div = self.c2r(self._mult(self.r2c(dx), 0))
        div = div + self.c2r(self._mult(self.r2c(dz), 2))
        div = div + self.c2r(self._mult(self.r2c(dy), 1))
        return div


class DiffusionSolver:
    """src/pde/DiffusionSolver.cpp:5-250 with PETSc KSPCG semantics (see rd_oracle.py)."""
    RTOL, ABSTOL, DTOL, MAXIT = 1e-6, 1e-50, 1e4, 5000

    def __init__(self, k, kavg, k_scale, dt_ctx, spec: SpectralOps):
        self.k, self.kavg, self.k_scale, self.spec = k, float(kavg), float(k_scale), spec
        self.dtype = k.dtype
        self.dt_ctx = float(dt_ctx)
        self.ksp_itr = 0
        self.prec_factor()

    def prec_factor(self):
        n0, n1, n2 = self.spec.shape
        dt = self.dtype
        wx = wavenumbers(n0, dt).reshape(n0, 1, 1)
        wy = wavenumbers(n1, dt).reshape(1, n1, 1)
        wz = wavenumbers(n2, dt)[: n2 // 2 + 1].reshape(1, 1, -1)
        ka = torch.tensor(self.kavg, dtype=dt)
        s = ((ka * wx) * wx).double() + ((ka * wy) * wy).double() + ((ka * wz) * wz).double()
        pf = (1.0 + 0.25 * self.dt_ctx * s).to(dt)
        factor = torch.tensor(1.0 / (n0 * n1 * n2), dtype=dt)
        self.precfactor = torch.where(pf == 0, factor, factor / pf)

    def apply_D(self, c):
        gx, gy, gz = self.spec.gradient(c)
        return self.spec.divergence(self.k * gx, self.k * gy, self.k * gz)

    def operator_A(self, x):
        return x + (-0.5 * self.dt_ctx) * self.apply_D(x)

    def apply_pc(self, x):
        return self.spec.c2r(self.spec.r2c(x) * self.precfactor)

    @staticmethod
    def _dot(a, b):
        return float(torch.dot(a.reshape(-1).double(), b.reshape(-1).double()))

    def solve(self, c, dt):
        self.dt_ctx = float(torch.tensor(dt, dtype=self.dtype))
        if self.k_scale == 0:
            self.ksp_itr = 0
            return c
        b = c + (0.5 * self.dt_ctx) * self.apply_D(c)
        x = c.clone()
        r = b - self.operator_A(x)
        z = self.apply_pc(r)
        dp = math.sqrt(self._dot(z, z))
        zb = self.apply_pc(b)
        rnorm0 = math.sqrt(self._dot(zb, zb)) or dp
        ttol = max(self.RTOL * rnorm0, self.ABSTOL)
        its = 0
        if dp <= ttol:
            self.ksp_itr = 0
            return x
        beta, betaold, p = 0.0, 1.0, None
        while its < self.MAXIT:
            if its == 0:
                beta = self._dot(z, r)
                if beta == 0.0:
                    break
                p = z.clone()
            else:
                p = z + float(torch.tensor(beta / betaold, dtype=self.dtype)) * p
            w = self.operator_A(p)
            dpi = self._dot(p, w)
            betaold = beta
            a = float(torch.tensor(beta / dpi, dtype=self.dtype))
            x = x + a * p
            r = r - a * w
            z = self.apply_pc(r)
            dp = math.sqrt(self._dot(z, z))
            its += 1
            if dp <= ttol:
                break
            if dp >= self.DTOL * rnorm0:
                raise RuntimeError("KSP_DIVERGED_DTOL")
            beta = self._dot(z, r)
        self.ksp_itr = its
        return x


def reaction_nonlinear(c, rho, dt):
    factor = torch.exp(rho * dt)
    alph = (c.double() / (1.0 - c.double())).to(c.dtype)
    af = alph * factor
    out = (af.double() / (af.double() + 1.0)).to(c.dtype)
    return torch.where(torch.isinf(alph), torch.ones_like(out), out)


def reaction_linearized(u, c_lin, rho, dt):
    factor = torch.exp(rho * dt)
    alph = ((c_lin * factor).double() + 1.0 - c_lin.double()).to(u.dtype)
    return (u * factor) / (alph * alph)


class PdeOperatorsRD:
    """src/pde/PdeOperators.cpp:235-420 (state + adjoint with c_half_ storage)."""

    def __init__(self, k, kavg, k_scale, rho, nt, dt, dt_ctx=None):
        self.spec = SpectralOps(k.shape, k.dtype)
        # dt_ctx: the solver context's time step when precFactor() ran (trap T2; default = dt as at construction)
        self.diff = DiffusionSolver(k, kavg, k_scale, dt if dt_ctx is None else dt_ctx, self.spec)
        self.rho, self.nt, self.dt = rho, int(nt), float(dt)
        self.c_ = [None] * (nt + 1)
        self.p_ = [None] * (nt + 1)
        self.c_half_ = [None] * nt
        self.ksp_state = self.ksp_adj = 0

    def solve_state(self, c0):
        c = c0.clone()
        self.c_[0] = c.clone()
        self.ksp_state = 0
        for i in range(self.nt):
            c = self.diff.solve(c, self.dt / 2.0)
            self.ksp_state += self.diff.ksp_itr
            self.c_half_[i] = c.clone()
            c = reaction_nonlinear(c, self.rho, self.dt)
            c = self.diff.solve(c, self.dt / 2.0)
            self.ksp_state += self.diff.ksp_itr
            self.c_[i + 1] = c.clone()
        return c

    def solve_adjoint(self, pT):
        p = pT.clone()
        self.p_[self.nt] = p.clone()
        self.ksp_adj = 0
        for i in range(self.nt):
            p = self.diff.solve(p, self.dt / 2.0)
            self.ksp_adj += self.diff.ksp_itr
            it = self.nt - i - 1
            p = reaction_linearized(p, self.c_half_[it], self.rho, self.dt)
            p = self.diff.solve(p, self.dt / 2.0)
            self.ksp_adj += self.diff.ksp_itr
            self.p_[it] = p.clone()
        return p


def forward_adjoint(k, kavg, k_scale, rho, c0, d1, nt, dt, dt_ctx=None):
    """One solveState(0) + terminal condition -(c(T) - d1) + solveAdjoint(1)."""
    pde = PdeOperatorsRD(k, kavg, k_scale, rho, nt, dt, dt_ctx)
    cT = pde.solve_state(c0)
    p0 = pde.solve_adjoint(-(cT - d1))
    return cT, p0, pde
