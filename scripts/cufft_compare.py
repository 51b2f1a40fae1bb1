"""cuFFT comparison point: the call sequence of the reference's CUDA backend --
cufftPlan3d R2C / C2R for every gradient, divergence and preconditioner apply
(src/grad/SpectralOperators.cpp:100-261, src/pde/DiffusionSolver.cpp:182-215), unfused
elementwise kernels between them, one host read-back per dot product -- timed on the same
synthetic 256^3 problem as bench.py.  cuFFT is reached through torch.fft on the GPU.

COMPARISON ONLY: nothing in glia_b200 calls cuFFT, and this script is not part of bench.py's
contract.  It prints one JSON line: time-steps/s of a forward+adjoint solve and the share of
the time spent inside cuFFT itself (timed separately on the same shapes).

  python scripts/cufft_compare.py [--n 256] [--nt 2]
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glia_b200 import synthetic as S  # noqa: E402  (inputs only)


def wavenumbers(n, dtype, dev):
    w = torch.arange(n, dtype=torch.int64)
    w = torch.where(w > n // 2, w - n, w)
    w[n // 2] = 0
    return w.to(dtype).to(dev)


class CufftPath:
    def __init__(self, k, kavg, rho, dt, dev):
        self.k, self.rho, self.dt, self.dev = k, rho, float(dt), dev
        self.shape = tuple(k.shape)
        n0, n1, n2 = self.shape
        dtp = k.dtype
        f = 1.0 / (n0 * n1 * n2)
        self.iw = [(f * wavenumbers(n0, dtp, dev)).reshape(n0, 1, 1), (f * wavenumbers(n1, dtp, dev)).reshape(1, n1, 1),
                   (f * wavenumbers(n2, dtp, dev))[: n2 // 2 + 1].reshape(1, 1, -1)]
        wx, wy, wz = (wavenumbers(n0, dtp, dev).reshape(n0, 1, 1), wavenumbers(n1, dtp, dev).reshape(1, n1, 1),
                      wavenumbers(n2, dtp, dev)[: n2 // 2 + 1].reshape(1, 1, -1))
        s = kavg * (wx * wx + wy * wy + wz * wz)
        self.pc = f / (1.0 + 0.25 * (self.dt / 2) * s)   # matched-dt symbol (the stale-dt trap is irrelevant for timing)
        self.nfft = 0
        self.its = 0

    def r2c(self, x):
        self.nfft += 1
        return torch.fft.rfftn(x)

    def c2r(self, xh):
        self.nfft += 1
        return torch.fft.irfftn(xh, s=self.shape, norm="forward")

    def mult(self, fh, d):
        return torch.complex(-self.iw[d] * fh.imag, self.iw[d] * fh.real)

    def apply_D(self, c):  # 10 3-D FFTs (DiffCoef::applyD, src/mat/DiffCoef.cpp:249-271)
        fh = self.r2c(c)
        g = [self.c2r(self.mult(fh, d)) for d in range(3)]
        t = [self.k * gi for gi in g]
        div = self.c2r(self.mult(self.r2c(t[0]), 0))
        div = div + self.c2r(self.mult(self.r2c(t[2]), 2))
        return div + self.c2r(self.mult(self.r2c(t[1]), 1))

    def apply_pc(self, r):
        return self.c2r(self.r2c(r) * self.pc)

    def solve(self, c, dth):
        dot = lambda a, b: float(torch.dot(a.reshape(-1), b.reshape(-1)))  # host read-back, as the reference's VecDot
        b = c + (0.5 * dth) * self.apply_D(c)
        x = c.clone()
        r = b - (x - (0.5 * dth) * self.apply_D(x))
        z = self.apply_pc(r)
        dp = math.sqrt(dot(z, z))
        zb = self.apply_pc(b)
        ttol = max(1e-6 * math.sqrt(dot(zb, zb)), 1e-50)
        beta, betaold, p = 0.0, 1.0, None
        its = 0
        while dp > ttol and its < 5000:
            beta = dot(z, r)
            p = z.clone() if its == 0 else z + (beta / betaold) * p
            w = p - (0.5 * dth) * self.apply_D(p)
            a = beta / dot(p, w)
            betaold = beta
            x = x + a * p
            r = r - a * w
            z = self.apply_pc(r)
            dp = math.sqrt(dot(z, z))
            its += 1
        self.its += its
        return x

    def forward_adjoint(self, c0, d1, nt):
        c = c0.clone()
        half = []
        for _ in range(nt):
            c = self.solve(c, self.dt / 2)
            half.append(c.clone())
            f = torch.exp(self.rho * self.dt)
            a = c / (1 - c)
            c = a * f / (a * f + 1)
            c = self.solve(c, self.dt / 2)
        p = -(c - d1)
        for i in range(nt):
            p = self.solve(p, self.dt / 2)
            ch = half[nt - 1 - i]
            f = torch.exp(self.rho * self.dt)
            p = p * f / (ch * f + 1 - ch) ** 2
            p = self.solve(p, self.dt / 2)
        return c, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--nt", type=int, default=2)
    ap.add_argument("--dt", type=float, default=0.04)
    a = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device"
    dev = torch.device("cuda:0")
    atlas = S.make_atlas(a.n, 0, np.float32)
    c0 = torch.from_numpy(S.make_initial_condition(atlas, 0, dtype=np.float32)).to(dev)
    wm, gm = torch.from_numpy(atlas["wm"]).to(dev), torch.from_numpy(atlas["gm"]).to(dev)
    k, rho = 0.01 * wm, 8.0 * wm
    kavg = float(k.sum(dtype=torch.float64)) / float(atlas["filter"].sum(dtype=np.float64))
    d1 = 0.5 * c0
    path = CufftPath(k, kavg, rho, a.dt, dev)
    path.forward_adjoint(c0, d1, 1)  # warm-up: cuFFT plans, allocator
    path.nfft = path.its = 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    path.forward_adjoint(c0, d1, a.nt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # cuFFT alone on the same shapes
    x = torch.randn(a.n, a.n, a.n, device=dev)
    xh = torch.fft.rfftn(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        xh = torch.fft.rfftn(x)
        x = torch.fft.irfftn(xh, s=x.shape, norm="forward")
    e1.record()
    torch.cuda.synchronize()
    fft_ms = e0.elapsed_time(e1) / 40
    F = 4.0 * a.n ** 3
    print(json.dumps({
        "what": "cuFFT comparison point: reference CUDA-backend call sequence (3-D cuFFT R2C/C2R + unfused elementwise, host dot read-backs)",
        "n": a.n, "nt": a.nt, "time_steps_per_s": a.nt / (ms / 1e3), "ms_per_time_step": ms / a.nt,
        "pcg_iterations": path.its, "mean_its_per_solve": path.its / (4 * a.nt), "ffts": path.nfft,
        "cufft_ms_per_3d_fft": fft_ms, "cufft_share": path.nfft * fft_ms / ms,
        "cufft_GBs_at_2F_per_fft": 2 * F / (fft_ms * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
