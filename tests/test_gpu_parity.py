"""Parity of the CUDA path (libglia_rd.so through the C ABI, torch CUDA tensors as device
memory) against the CPU oracle on the same seeded inputs, plus size-independent properties
at the BASELINE.json sizes.  Tolerances are north_star's: relative L2 1e-5 (f32), 1e-10 (f64)."""
import numpy as np
import pytest

import _cases as Cs
from oracle import rd_oracle as O

pytestmark = pytest.mark.gpu
DT = [np.float32, np.float64]
EPS = {np.dtype(np.float32): 2e-6, np.dtype(np.float64): 1e-13}


@pytest.fixture(scope="module")
def B(cuda_lib, torch_cuda):
    return Cs.TorchBackend(cuda_lib)


def test_library_is_the_cuda_build(B):
    h = B.handle(32, np.float32)
    assert h.lib.glia_rd_build_info() == b"cuda-sm_100a"
    h.close()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n", [32, 64, 128, 256, (64, 32, 128)])
def test_fft(B, n, dtype):
    for e in Cs.case_fft(B, n, dtype):
        assert e < EPS[np.dtype(dtype)]


def test_fft_512_f32(B):
    for e in Cs.case_fft(B, 512, np.float32):
        assert e < 3e-6


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n", [64, 256, (128, 64, 32)])
def test_grad_div(B, n, dtype):
    for e in Cs.case_grad_div(B, n, dtype):
        assert e < EPS[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n,sinusoidal", [(64, True), (64, False), (128, False), (256, False)])
def test_apply_D(B, n, dtype, sinusoidal):
    e1, e2, budget = Cs.case_apply_D(B, n, dtype, sinusoidal)
    assert e1 < budget and e2 < budget


def test_apply_D_512_f32(B):
    e1, e2, budget = Cs.case_apply_D(B, 512, np.float32, False)
    assert e1 < budget and e2 < budget


@pytest.mark.parametrize("dtype", DT)
def test_K1_reference_unit_test(B, dtype):
    """src/test/pdesolver.cpp:7-56 -- ||c||_2 == Approx(2.0487), ksp_itr_ == 5."""
    nrm, its_g, its_o, err = Cs.case_K1(B, dtype)
    assert abs(nrm - 2.0487) < 100 * np.finfo(np.float32).eps * (1 + 2.0487) + 5e-5
    assert its_g[-1] == 5
    assert its_g == its_o
    assert err < Cs.TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n,nt", [(64, 6), (128, 2)])
def test_forward_adjoint_gradient(B, n, nt, dtype):
    r = Cs.case_forward_adjoint(B, n, dtype, nt=nt, dt=0.04)
    assert r["its_state"][0] == r["its_state"][1]
    assert r["its_adj"][0] == r["its_adj"][1]
    tol = Cs.TOL[np.dtype(dtype)]
    assert r["cT"] < tol and r["p0"] < tol
    assert r["grad"] < 10 * tol, r["grad_vals"]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n,nt", [(64, 4), (128, 2)])
def test_objective_gradient_hessian(B, n, nt, dtype):
    """config 3 of BASELINE.json in miniature: one evaluateObjectiveAndGradient and Gauss-Newton
    Hessian products (incremental forward / adjoint, diffusivity inversion on and off)."""
    r = Cs.case_objective_hessian(B, n, dtype, nt=nt)
    tol = Cs.TOL[np.dtype(dtype)]
    assert r["its"][0] == r["its"][1] and r["h_its"][0] == r["h_its"][1], r
    assert r["J"] < 10 * tol and r["g_c0"] < 10 * tol and r["g6"] < 20 * tol, r
    assert r["h_y"] < 10 * tol and r["h_y_ponly"] < 10 * tol and r["h_k"] < 50 * tol, r


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n", [64, 128])
def test_mass_effect_style_steps(B, n, dtype):
    """k(x), rho(x) refreshed before every step, precFactor per step, order-1 splitting."""
    r = Cs.case_mass_effect_steps(B, n, dtype, nsteps=3)
    assert r["its"][0] == r["its"][1], r
    assert r["c"] < Cs.TOL[np.dtype(dtype)], r
    assert r["moved"] > 1e-2


def test_adjoint_without_store(B):
    r = Cs.case_forward_adjoint(B, 64, np.float64, nt=2, dt=0.04, adjoint_store=False, with_grad=False)
    assert r["its_adj"][0] == r["its_adj"][1]
    assert r["p0"] < 1e-10


# ---- size-independent properties at the headline size -------------------------------------
@pytest.mark.parametrize("n", [256])
def test_properties_at_full_size(B, n):
    torch = B.torch
    dtype = np.float32
    sh = (n, n, n)
    P = Cs.make_problem(n, dtype)
    h, dev = Cs.setup_handle(B, P, n, dtype, nt=2, dt=0.04)
    g = torch.Generator(device=B.dev).manual_seed(0)
    u = torch.randn(sh, device=B.dev, dtype=torch.float32, generator=g)
    v = torch.randn(sh, device=B.dev, dtype=torch.float32, generator=g)
    torch.cuda.synchronize()
    Du, Dv, Duv = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
    h.apply_D(Du, u)
    h.apply_D(Dv, v)
    # linearity
    w = (2.0 * u - 3.0 * v).contiguous()
    torch.cuda.synchronize()
    h.apply_D(Duv, w)
    lin = (Duv - (2.0 * Du - 3.0 * Dv)).double().norm() / Duv.double().norm()
    assert float(lin) < 1e-5
    # self-adjointness <u, D v> == <D u, v>, and D annihilates constants, and sum(D u) == 0
    a = float((u.double() * Dv.double()).sum())
    b = float((Du.double() * v.double()).sum())
    assert abs(a - b) < 1e-5 * max(abs(a), abs(b))
    ones = torch.ones_like(u)
    torch.cuda.synchronize()
    h.apply_D(Duv, ones)
    assert float(Duv.abs().max()) < 1e-6
    assert abs(float(Du.double().sum())) < 1e-4 * float(Du.double().abs().sum())
    # Crank-Nicolson solve: residual of (I - dt/4 D) x = (I + dt/4 D) c, checked with apply_D
    c = dev["wm"].clone()
    x = c.clone()
    torch.cuda.synchronize()
    its = h.diffusion_solve(x, 0.02)
    Dc, Dx = torch.empty_like(c), torch.empty_like(c)
    h.apply_D(Dc, c)
    h.apply_D(Dx, x)
    res = (x - 0.01 * Dx) - (c + 0.01 * Dc)
    assert 0 < its < 20
    assert float(res.double().norm() / c.double().norm()) < 1e-5
    # mass conservation of the diffusion step (D has zero mean)
    assert abs(float(x.double().sum() - c.double().sum())) < 1e-6 * float(c.double().sum())
    # FFT round trip
    fh = torch.empty((n, n, n // 2 + 1), dtype=torch.complex64, device=B.dev)
    y = torch.empty_like(u)
    torch.cuda.synchronize()
    h.fft_r2c(u, fh)
    h.fft_c2r(fh, y)
    assert float((y / n ** 3 - u).double().norm() / u.double().norm()) < 2e-6
    # Parseval with the half spectrum
    e_real = float((u.double() ** 2).sum())
    wgt = torch.full((n // 2 + 1,), 2.0, device=B.dev, dtype=torch.float64)
    wgt[0] = 1.0
    wgt[-1] = 1.0
    e_spec = float(((fh.real.double() ** 2 + fh.imag.double() ** 2) * wgt).sum()) / n ** 3
    assert abs(e_real - e_spec) < 1e-5 * e_real
    h.close()


def test_forward_is_deterministic_and_bounded(B):
    """Same inputs twice -> bitwise identical c(T) (fixed-order reductions); logistic growth
    keeps c within [min - eps, 1]."""
    n, nt, dt = 128, 3, 0.04
    P = Cs.make_problem(n, np.float32)
    h, dev = Cs.setup_handle(B, P, n, np.float32, nt, dt)
    c0 = B.put(P["c0"])
    a, b = B.empty((n, n, n), np.float32), B.empty((n, n, n), np.float32)
    i1 = h.solve_state(c0, a, 0)
    i2 = h.solve_state(c0, b, 0)
    assert i1 == i2
    assert B.torch.equal(a, b)
    assert float(a.max()) <= 1.0 + 1e-6
    h.close()


def test_reaction_kernels(B):
    n = 64
    for dtype in DT:
        P = Cs.make_problem(n, dtype)
        h, dev = Cs.setup_handle(B, P, n, dtype, 1, 0.04)
        c = Cs.smooth_field((n, n, n), dtype, 9, 0.0, 0.999)
        c[0, 0, 0] = 1.0  # a = c/(1-c) = inf branch
        cd = B.put(c)
        h.reaction(cd, None, 0.04)
        ref = O.reaction_nonlinear(c, P["rho"], 0.04)
        assert Cs.rel(B.get(cd), ref) < 5 * EPS[np.dtype(dtype)]
        u = Cs.smooth_field((n, n, n), dtype, 10, -1.0, 1.0)
        ud = B.put(u)
        h.reaction(ud, B.put(c), 0.04)
        assert Cs.rel(B.get(ud), O.reaction_linearized(u, c, P["rho"], 0.04)) < 5 * EPS[np.dtype(dtype)]
        h.close()


def test_error_paths(B):
    from glia_b200 import GliaRdError
    with pytest.raises(GliaRdError, match="powers of two"):
        B.handle(48, np.float32)
    h = B.handle(32, np.float32)
    with pytest.raises(GliaRdError, match="resize_history"):
        h.solve_state(B.empty((32, 32, 32), np.float32), None, 0)
    h.resize_history(2, 0.1)
    with pytest.raises(GliaRdError, match="out of range"):
        h.history_ptr(0, 7)
    h.close()


# ---- config 1 of BASELINE.json: test_forward_config.txt inputs (model 1) vs committed golden ----
@pytest.mark.parametrize("name,dtype", [("f32", np.float32), ("f64", np.float64)])
def test_config1_brain_forward_adjoint_vs_golden(B, name, dtype):
    """atlas.nc split + smoothed, one Gaussian at (137,169,96), rho 8, kappa 0.01, nt 25,
    dt 0.04; expected values generated by tests/golden/make_golden.py from the oracle in the
    build container (where K2 pins the inputs: ||c0|| = 4.09351)."""
    from golden import fixtures as FX
    z = np.load(FX.FWD)
    P = FX.brain_problem(dtype)
    n, nt, dt = P["n"], P["nt"], P["dt"]
    m = P["m"]
    h = B.handle(n, dtype, dt_ctx=dt)
    wm, gm, csf = B.put(m["wm"]), B.put(m["gm"]), B.put(m["csf"])
    h.set_diffusion_tissue(wm, gm, csf, 0.01, 0.0, 0.0, float(m["filter"].sum(dtype=np.float64)))
    h.set_reaction_tissue(wm, gm, csf, 8.0, 0.0, 0.0)
    h.prec_factor()
    h.resize_history(nt, dt)
    cT = B.empty((n, n, n), dtype)
    its_s = h.solve_state(B.put(P["c0"]), cT, 0)
    cTh = B.get(cT)
    tol = Cs.TOL[np.dtype(dtype)]
    nrm = lambda a: float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
    assert its_s == int(z[f"{name}_its_state"])
    assert abs(nrm(cTh) - float(z[f"{name}_cT_norm"])) < tol * float(z[f"{name}_cT_norm"])
    assert Cs.rel(cTh[34, 42, :], z[f"{name}_cT_line"]) < 10 * tol
    pT = (-(cTh - (0.5 * cTh).astype(dtype))).astype(dtype)
    p0 = B.empty((n, n, n), dtype)
    its_a = h.solve_adjoint(B.put(pT), p0, 1, True)
    p0h = B.get(p0)
    assert its_a == int(z[f"{name}_its_adj"])
    assert abs(nrm(p0h) - float(z[f"{name}_p0_norm"])) < 10 * tol * float(z[f"{name}_p0_norm"])
    # one z line in a low-amplitude region: single precision cannot resolve it better than the
    # oracle's own float32 run resolves it against its float64 run, so that distance is the budget
    line_budget = max(10 * tol, 2.0 * Cs.rel(z[f"{name}_p0_line"], z["f64_p0_line"]))
    assert Cs.rel(p0h[34, 42, :], z["f64_p0_line"]) < line_budget
    g = h.grad_kappa_rho(wm, gm, csf)
    gr = z[f"{name}_grad"]
    assert np.max(np.abs(g - gr) / np.maximum(np.abs(gr), 1e-300)) < 20 * tol
    h.close()


# ---- config 5 of BASELINE.json in miniature: independent ensemble members on one GPU ----------
def test_ensemble_members_concurrent(B):
    """inverse_ensemble-style batch: independent 128^3 forward solves with varied (kappa, rho), one
    handle (= one stream) per member, driven concurrently from host threads.  Every member must
    match the oracle and be bit-identical to the same member run alone."""
    import threading
    n, nt, dt, dtype = 128, 2, 0.04, np.float32
    sh = (n, n, n)
    P = Cs.make_problem(n, dtype)
    params = [(0.005, 4.0), (0.02, 9.0), (0.05, 15.0), (0.01, 8.0)]
    dev = {key: B.put(P[key]) for key in ("wm", "gm", "csf")}
    fsum = float(P["filt"].sum(dtype=np.float64))
    c0 = B.put(P["c0"])

    def member(kappa, rho, out):
        h = B.handle(n, dtype, dt_ctx=dt)
        h.set_diffusion_tissue(dev["wm"], dev["gm"], dev["csf"], kappa, 0.2, 0.0, fsum)
        h.set_reaction_tissue(dev["wm"], dev["gm"], dev["csf"], rho, 0.2, 0.0)
        h.prec_factor()
        h.resize_history(nt, dt)
        its = h.solve_state(c0, out, 0)
        h.close()
        return its

    alone = []
    for kappa, rho in params:
        o = B.empty(sh, dtype)
        alone.append((member(kappa, rho, o), B.get(o)))
    outs = [B.empty(sh, dtype) for _ in params]
    its = [None] * len(params)

    def run(i):
        its[i] = member(params[i][0], params[i][1], outs[i])

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(params))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, (kappa, rho) in enumerate(params):
        got = B.get(outs[i])
        assert its[i] == alone[i][0]
        assert np.array_equal(got, alone[i][1]), f"member {i} differs when run concurrently"
        k = O.DiffCoef(sh, dtype)
        k.set_values(kappa, 0.2, 0.0, P["wm"], P["gm"], P["csf"], P["filt"])
        pde = O.PdeOperatorsRD(k, O.reac_coef(rho, 0.2, 0.0, P["wm"], P["gm"], P["csf"]), nt, dt, dt_ctx=dt)
        ref = pde.solve_state(P["c0"], 0)
        assert its[i] == pde.ksp_state, (i, its[i], pde.ksp_state)
        assert Cs.rel(got, ref) < Cs.TOL[np.dtype(dtype)], (i, Cs.rel(got, ref))


# ---- callers either side of the path: smoother, MatProp, Phi ----------------
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n", [64, (64, 128, 32), 256])
def test_smoother(B, n, dtype):
    e1, e2, e0 = Cs.case_smooth(B, n, dtype)
    assert e1 < 5 * EPS[np.dtype(dtype)] and e2 < 5 * EPS[np.dtype(dtype)] and e0 == 0.0


def test_mat_prop(B):
    ok, fs, fs_ref = Cs.case_mat_prop(B, 64, np.float32)
    assert ok and fs == fs_ref


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n", [64, 128])
def test_phi_apply_and_transpose(B, n, dtype):
    e_apply, e_t, e_adj, zero_ok = Cs.case_phi(B, n, dtype)
    tol = Cs.TOL[np.dtype(dtype)]
    assert e_apply < tol and e_t < tol and e_adj < 10 * tol and zero_ok


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("which", ["K2", "K3"])
def test_reference_pins_K2_K3_on_the_gpu(B, which, dtype):
    """src/test/simulator.cpp:41-42 (sinusoid, ||c0|| = 22.0161) and :94-95 (brain, ||c0|| =
    4.09351): the reference's own known answers, computed end to end by the CUDA path from the
    committed test data -- tissue smoothing, MatProp filter, Phi::apply."""
    nrm, expected, err = Cs.case_K2_K3(B, dtype, which)
    # Catch2 Approx: |a - b| < eps (1 + |b|), eps = 100 FLT_EPSILON
    assert abs(nrm - expected) < 100 * np.finfo(np.float32).eps * (1 + expected) + 5e-5, (nrm, expected)
    assert err < Cs.TOL[np.dtype(dtype)], err


def test_first_order_splitting(B):
    """params->tu_->order_ != 2 (PdeOperators.cpp:284-290, 395-398) through solve_state / solve_adjoint."""
    for dtype in DT:
        r = Cs.case_forward_adjoint(B, 64, dtype, nt=3, dt=0.04, order=1)
        tol = Cs.TOL[np.dtype(dtype)]
        assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], r
        assert r["cT"] < tol and r["p0"] < tol and r["grad"] < 10 * tol, r


@pytest.mark.parametrize("dtype", DT)
def test_two_snapshot_objective(B, dtype):
    """two_time_points_: J and its c(0) gradient carry the t = 0 mismatch (DerivativeOperatorsRD.cpp:30-34,
    149-153, 216-222); evaluateHessian refuses, as in the reference (:234)."""
    r = Cs.case_two_snapshot(B, 64, dtype, nt=2)
    tol = Cs.TOL[np.dtype(dtype)]
    assert r["m0_share"] > 1e-3, r
    assert r["J"] < 10 * tol and r["m0"] < 10 * tol and r["g_c0"] < 10 * tol, r
    assert r["hessian_refused"] is True, r
    assert r["J_off"] < 10 * tol and r["m0_off"] == 0.0, r


def test_headline_size_one_time_step_vs_oracle(B):
    """256^3, single precision (BASELINE config 2's grid): one Strang time step forward + adjoint + the kappa / rho
    gradient against the CPU oracle, with EQUAL PCG iteration counts -- the m_i that set the roofline bytes of the
    bench are the oracle's.  ~1 min of CPU for the oracle."""
    r = Cs.case_forward_adjoint(B, 256, np.float32, nt=1, dt=0.04)
    assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], r
    assert r["cT"] < 1e-5 and r["p0"] < 1e-5 and r["grad"] < 1e-4, (r["cT"], r["p0"], r["grad"], r["grad_vals"])


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n,nt", [(64, 3), (128, 2)])
def test_ensemble_batch_handle(B, n, nt, dtype):
    """BASELINE config 5 in miniature: members with their own (kappa, rho) batched through the kernels' member index;
    per-member fields and PCG iteration counts equal the oracle's for each member alone."""
    r = Cs.case_ensemble_batch(B, n, dtype, nt=nt)
    tol = Cs.TOL[np.dtype(dtype)]
    assert all(a == b for a, b in r["its_state"]) and all(a == b for a, b in r["its_adj"]), r
    assert len({a for a, _ in r["its_state"]}) > 1, r
    assert max(r["cT"]) < tol and max(r["p0"]) < tol, r
    assert r["sum_ok"] and r["guard"], r


def test_wait_stream_orders_a_producer(B):
    """glia_rd_wait_stream: a field produced on another (torch) stream is ordered in front of the library's work
    without a device-wide synchronisation."""
    torch = B.torch
    n = 128
    h = B.handle(n, np.float32)
    side = torch.cuda.Stream()
    x = torch.zeros((n, n, n), dtype=torch.float32, device=B.dev)
    gz = torch.empty_like(x)
    torch.cuda.synchronize()
    ax = torch.arange(n, device=B.dev, dtype=torch.float32) * (2 * np.pi / n)
    with torch.cuda.stream(side):
        for _ in range(20):           # keep the producer busy for a while
            x.copy_(torch.sin(3 * ax)[None, None, :].expand(n, n, n))
    h.wait_stream(side)
    h.gradient(None, None, gz, x, 4)
    torch.cuda.synchronize()
    ref = (3 * torch.cos(3 * ax))[None, None, :].expand(n, n, n)
    assert float((gz - ref).abs().max()) < 1e-4
    h.close()


def test_netcdf_round_trip_and_label_split_on_device(B, tmp_path):
    """glia_rd_data_out / data_in with device fields (CDF-2 file readable by scipy) and the label split."""
    import scipy.io
    from golden import fixtures as FX
    seg = FX.atlas_labels().astype(np.float32)
    h = B.handle(64, np.float32)
    sd = B.put(seg)
    p = str(tmp_path / "seg.nc")
    h.data_out(p, sd)
    f = scipy.io.netcdf_file(p, "r", mmap=False)
    assert np.array_equal(f.variables["data"].data, seg)
    f.close()
    back = B.empty(seg.shape, np.float32)
    h.data_in(p, back)
    assert np.array_equal(B.get(back), seg)
    maps = {k: B.empty(seg.shape, np.float32) for k in ("wm", "gm", "vt", "csf")}
    h.split_segmentation(back, (6, 5, 7, 8), maps["wm"], maps["gm"], maps["vt"], maps["csf"])
    ref = O.split_segmentation(seg, (6, 5, 7, 8), np.float32)
    for k in maps:
        assert np.array_equal(B.get(maps[k]), ref[k])
    h.close()
