"""The multi-threaded torch-CPU baseline port agrees with the NumPy oracle."""
import numpy as np
import torch

import _cases as Cs
from oracle import rd_oracle as O
from oracle import rd_oracle_torch as OT


def test_torch_port_matches_numpy_oracle():
    n, nt, dt = 32, 2, 0.05
    for dtype, tol in ((np.float64, 1e-11), (np.float32, 2e-5)):
        P = Cs.make_problem(n, dtype)
        pde = O.PdeOperatorsRD(P["k"], P["rho"], nt, dt, dt_ctx=dt)
        cT = pde.solve_state(P["c0"], 0)
        d1 = (0.9 * cT).astype(dtype)
        p0 = pde.solve_adjoint((-(cT - d1)).astype(dtype), 1)
        t = torch.from_numpy
        cT_t, p0_t, pt = OT.forward_adjoint(t(P["k"].kxx), float(P["k"].kxx_avg), P["k_scale"], t(P["rho"]),
                                            t(P["c0"]), t(d1), nt, dt)
        assert pt.ksp_state == pde.ksp_state and pt.ksp_adj == pde.ksp_adj
        assert Cs.rel(cT_t.numpy(), cT) < tol
        assert Cs.rel(p0_t.numpy(), p0) < tol
