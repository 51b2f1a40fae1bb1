"""CPU tests of the product kernel sources and host drivers through the test-only SIMT
emulator build (tests/emu): same C ABI, same code, fibers instead of CUDA threads.
Small grids only; the parity tests proper are the gpu ones."""
import numpy as np
import pytest

import _cases as Cs

DT = [np.float32, np.float64]
EPS = {np.dtype(np.float32): 2e-6, np.dtype(np.float64): 1e-13}


@pytest.fixture(scope="module")
def B(emu_lib):
    return Cs.NumpyBackend(emu_lib)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n", [32, (32, 64, 32)])
def test_fft(B, n, dtype):
    for e in Cs.case_fft(B, n, dtype):
        assert e < EPS[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", DT)
def test_grad_div(B, dtype):
    for e in Cs.case_grad_div(B, (32, 32, 64), dtype):
        assert e < EPS[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("sinusoidal", [True, False])
def test_apply_D(B, dtype, sinusoidal):
    e1, e2, budget = Cs.case_apply_D(B, 32, dtype, sinusoidal)
    assert e1 < budget and e2 < budget


@pytest.mark.parametrize("dtype", DT)
def test_pcg_matches_oracle_iterations(B, dtype):
    nrm, its_g, its_o, err = Cs.case_K1(B, dtype, n=32, nsolves=2)
    assert its_g == its_o
    assert err < Cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", DT)
def test_forward_adjoint_gradient(B, dtype):
    r = Cs.case_forward_adjoint(B, 32, dtype, nt=2, dt=0.05)
    assert r["its_state"][0] == r["its_state"][1]
    assert r["its_adj"][0] == r["its_adj"][1]
    tol = Cs.TOL[np.dtype(dtype)]
    assert r["cT"] < tol and r["p0"] < tol
    assert r["grad"] < 10 * tol, r["grad_vals"]


@pytest.mark.parametrize("dtype", DT)
def test_pcg_edge_cases(B, dtype):
    r = Cs.case_pcg_edges(B, 32, dtype)
    assert r["kscale0_its"] == 0 and r["kscale0_unchanged"], r
    assert r["zero_its"] == 0 and r["zero_out"] == 0.0, r
    assert r["maxit"][0] == r["maxit"][1] == r["maxit"][2], r
    assert r["maxit_err"] < Cs.TOL[np.dtype(dtype)], r
    assert r["loose"][0] == r["loose"][1] and r["loose"][0] <= r["loose"][2], r
    assert r["loose_err"] < Cs.TOL[np.dtype(dtype)], r


@pytest.mark.parametrize("dtype", DT)
def test_mass_effect_style_steps(B, dtype):
    r = Cs.case_mass_effect_steps(B, 32, dtype, nsteps=2)
    assert r["its"][0] == r["its"][1], r
    assert r["c"] < Cs.TOL[np.dtype(dtype)], r
    assert r["moved"] > 1e-2


@pytest.mark.parametrize("dtype", [np.float64])   # single precision of the same case runs on the GPU
def test_objective_gradient_hessian(B, dtype):
    r = Cs.case_objective_hessian(B, 32, dtype, nt=1)
    tol = Cs.TOL[np.dtype(dtype)]
    assert r["its"][0] == r["its"][1] and r["h_its"][0] == r["h_its"][1], r
    assert r["J"] < 10 * tol and r["g_c0"] < 10 * tol and r["g6"] < 20 * tol, r
    assert r["h_y"] < 10 * tol and r["h_y_ponly"] < 10 * tol and r["h_k"] < 50 * tol, r


@pytest.mark.parametrize("dtype", DT)
def test_smoother(B, dtype):
    e1, e2, e0 = Cs.case_smooth(B, (32, 32, 64), dtype)
    assert e1 < 5 * EPS[np.dtype(dtype)] and e2 < 5 * EPS[np.dtype(dtype)] and e0 == 0.0


def test_mat_prop(B):
    ok, fs, fs_ref = Cs.case_mat_prop(B, 32, np.float32)
    assert ok and fs == fs_ref


@pytest.mark.parametrize("dtype", DT)
def test_phi_apply_and_transpose(B, dtype):
    e_apply, e_t, e_adj, zero_ok = Cs.case_phi(B, 32, dtype)
    tol = Cs.TOL[np.dtype(dtype)]
    assert e_apply < tol and e_t < tol and e_adj < 10 * tol and zero_ok


@pytest.mark.parametrize("dtype", DT)
def test_z_sweep_three_pass_lines(B, dtype):
    """512-point z lines (three-pass plan) through the pipelined z second-derivative sweep."""
    n = (32, 32, 512) if np.dtype(dtype) == np.float32 else (32, 64, 128)
    e1, e2, budget = Cs.case_apply_D(B, n, dtype, sinusoidal=False)
    assert e1 < budget and e2 < budget


@pytest.mark.parametrize("n,dtype", [((32, 32, 128), np.float32), ((32, 32, 256), np.float32),
                                     ((32, 32, 512), np.float32), ((32, 32, 128), np.float64)])
def test_pcg_long_z_lines(B, n, dtype):
    """128-, 256- and 512-point z lines through the whole PCG loop: the persistent warp-private r2c / c2r sweeps of the
    preconditioner (two-pass and three-pass plans; double precision at 128 points takes the one-group-per-CTA forms)."""
    r = Cs.case_forward_adjoint(B, n, dtype, nt=1, dt=0.05, with_grad=False)
    assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], r
    tol = Cs.TOL[np.dtype(dtype)]
    assert r["cT"] < tol and r["p0"] < tol, r


def test_first_order_splitting(B):
    r = Cs.case_forward_adjoint(B, 32, np.float64, nt=2, dt=0.05, order=1)
    assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], r
    assert r["cT"] < 1e-10 and r["p0"] < 1e-10 and r["grad"] < 1e-9, r


def test_two_snapshot_objective(B):
    r = Cs.case_two_snapshot(B, 32, np.float64, nt=1)
    assert r["m0_share"] > 1e-3, r            # the term matters in this case
    assert r["J"] < 1e-10 and r["m0"] < 1e-10 and r["g_c0"] < 1e-9, r
    assert r["hessian_refused"] is True, r
    assert r["J_off"] < 1e-10 and r["m0_off"] == 0.0, r


@pytest.mark.parametrize("dtype", DT)
def test_ensemble_batch_handle(B, dtype):
    r = Cs.case_ensemble_batch(B, 32, dtype, nt=1)
    tol = Cs.TOL[np.dtype(dtype)]
    assert all(a == b for a, b in r["its_state"]) and all(a == b for a, b in r["its_adj"]), r
    assert len({a for a, _ in r["its_state"]}) > 1, r     # the members really converge at different iterations
    assert max(r["cT"]) < tol and max(r["p0"]) < tol, r
    assert r["sum_ok"] and r["guard"], r


def test_s_sweeps_512_point_lines(B):
    """512-point x lines (three-pass plan, 512-thread CTAs) through the pipelined S kernels: D-apply, operatorA,
    preconditioner and PCG."""
    n = (512, 32, 32)
    e1, e2, budget = Cs.case_apply_D(B, n, np.float32, sinusoidal=False)
    assert e1 < budget and e2 < budget
    r = Cs.case_forward_adjoint(B, n, np.float32, nt=1, dt=0.05, with_grad=False)
    assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], r
    assert r["cT"] < 1e-5 and r["p0"] < 1e-5, r
