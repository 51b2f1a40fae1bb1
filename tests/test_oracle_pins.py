"""Pins of the CPU oracle against the reference's own known-answer tests."""
import numpy as np
import pytest

from oracle import rd_oracle as O


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_K1_diffusion_solver(dtype):
    """src/test/pdesolver.cpp:7-56: 64^3, sinusoidal k (1e-2), test Gaussian, precFactor with
    the construction-time dt (0.5), 10 x solve(c, 0.02): ||c||_2 == Approx(2.0487), ksp_itr_ == 5."""
    n = 64
    k = O.DiffCoef((n, n, n), dtype)
    k.set_values_sinusoidal(1e-2)
    ds = O.DiffusionSolver(k, 0.5)
    c = O.test_gaussian(n, dtype)
    for _ in range(10):
        c = ds.solve(c, 0.02)
    nrm = float(np.sqrt(np.sum(c.astype(np.float64) ** 2)))
    # Catch2 Approx: |a - b| < eps * (1 + |b|), eps = 100 * FLT_EPSILON
    assert abs(nrm - 2.0487) < 100 * np.finfo(np.float32).eps * (1 + 2.0487) + 5e-5
    assert ds.ksp_itr == 5


def test_stale_dt_trap_matters():
    """With a matched preconditioner dt the same solve needs fewer iterations (trap T2):
    the unit test's 5 only comes out with dt_ctx = 0.5."""
    n = 64
    k = O.DiffCoef((n, n, n), np.float64)
    k.set_values_sinusoidal(1e-2)
    ds = O.DiffusionSolver(k, 0.02)
    ds.solve(O.test_gaussian(n, np.float64), 0.02)
    assert ds.ksp_itr < 5


def test_derivative_matches_analytic():
    n = 32
    x = 2 * np.pi * np.arange(n) / n
    f = np.sin(3 * x)[:, None, None] * np.cos(2 * x)[None, :, None] * np.sin(x)[None, None, :]
    gx, gy, gz = O.gradient(f)
    ex = 3 * np.cos(3 * x)[:, None, None] * np.cos(2 * x)[None, :, None] * np.sin(x)[None, None, :]
    assert np.abs(gx - ex).max() < 1e-12
    # Nyquist mode is annihilated (trap T1)
    ny = np.cos(n // 2 * x)[:, None, None] * np.ones((1, n, n))
    assert np.abs(O.gradient(ny)[0]).max() < 1e-12


# ---- K2 / K3: c(0) = Phi p on the reference's shipped test data (committed as golden fixtures) ----
@pytest.mark.parametrize("dtype,tol", [(np.float64, 3e-6), (np.float32, 1.2e-5 * 5.1)])
def test_K2_brain_c0_norm(dtype, tol):
    """src/test/simulator.cpp:94-95: ||c_0||_2 == Approx(4.09351f) for test_forward_config.txt
    (atlas.nc split with labels wm 6, gm 5, vt 7, csf 8)."""
    from golden import fixtures as FX
    P = FX.brain_problem(dtype)
    nrm = float(np.sqrt(np.sum(P["c0"].astype(np.float64) ** 2)))
    assert abs(nrm - 4.09351) < tol + 1.2e-5 * (1 + 4.09351)
    assert float(P["m"]["filter"].sum()) > 1000


def test_K3_sinusoid_c0_norm():
    """src/test/simulator.cpp:24,41-42: ||c_0||_2 == Approx(22.0161f), wm = sinusoid.nc,
    smoothing_factor_atlas_ = 0, sigma_factor 4."""
    from golden import fixtures as FX
    for dtype in (np.float64, np.float32):
        c0 = FX.sinusoid_c0(dtype)
        nrm = float(np.sqrt(np.sum(c0.astype(np.float64) ** 2)))
        assert abs(nrm - 22.0161) < 1.2e-5 * (1 + 22.0161) + 5e-5


def test_forward_fixture_is_reproducible():
    """The committed config-1 fixture (tests/golden/rd_forward_64.npz) is what the oracle
    produces today (float32 leg only, to keep the CPU suite short)."""
    from golden import fixtures as FX
    z = np.load(FX.FWD)
    P = FX.brain_problem(np.float32)
    assert abs(float(np.sqrt(np.sum(P["c0"].astype(np.float64) ** 2))) - float(z["f32_c0_norm"])) < 1e-6
    pde = O.PdeOperatorsRD(P["k"], P["rho"], 3, P["dt"], dt_ctx=P["dt"])
    pde.solve_state(P["c0"], 0)
    assert pde.ksp_state > 0
