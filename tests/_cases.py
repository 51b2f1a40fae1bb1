"""Parity cases shared by the emulator (CPU) and the CUDA (gpu) test modules.

A `Backend` hides where fields live: NumPy arrays handed to the emulator build by host
pointer, or torch CUDA tensors handed to libglia_rd.so by device pointer.  Every case runs the
same inputs through the C ABI and through the CPU oracle and returns the comparison.
"""
from __future__ import annotations

import numpy as np

from glia_b200.rd import RDHandle
from oracle import rd_oracle as O

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-10}  # BASELINE.json north_star


def rel(a, b):
    wide = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = np.asarray(a, dtype=wide).ravel()
    b = np.asarray(b, dtype=wide).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class NumpyBackend:
    """emulator: 'device' memory is host memory."""
    def __init__(self, lib_path):
        self.lib_path = lib_path

    def handle(self, n, dtype, dt_ctx=0.5):
        return RDHandle(n, "f32" if np.dtype(dtype) == np.float32 else "f64", dt_ctx=dt_ctx,
                        lib_path=self.lib_path)

    def put(self, a):
        return np.ascontiguousarray(a).copy()

    def empty(self, shape, dtype):
        return np.zeros(shape, dtype)

    def get(self, a):
        return np.array(a)


class TorchBackend:
    def __init__(self, lib_path, device=0):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.device = device
        self.lib_path = lib_path

    def handle(self, n, dtype, dt_ctx=0.5):
        return RDHandle(n, "f32" if np.dtype(dtype) == np.float32 else "f64", device=self.device,
                        dt_ctx=dt_ctx, lib_path=self.lib_path)

    def put(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.torch.cuda.synchronize()
        return t

    def empty(self, shape, dtype):
        td = {np.dtype(np.float32): self.torch.float32, np.dtype(np.float64): self.torch.float64,
              np.dtype(np.complex64): self.torch.complex64, np.dtype(np.complex128): self.torch.complex128}
        t = self.torch.zeros(tuple(shape), dtype=td[np.dtype(dtype)], device=self.dev)
        self.torch.cuda.synchronize()
        return t

    def get(self, a):
        self.torch.cuda.synchronize()
        return a.cpu().numpy()


def shape3(n):
    return (n, n, n) if np.isscalar(n) else tuple(n)


def smooth_field(shape, dtype, seed, lo=0.0, hi=1.0, kmax=3):
    """A smooth random periodic field in [lo, hi]."""
    rng = np.random.default_rng(seed)
    ax = [2 * np.pi * np.arange(m) / m for m in shape]
    f = np.zeros(shape)
    for _ in range(6):
        kx, ky, kz = rng.integers(0, kmax + 1, 3)
        ph = rng.uniform(0, 2 * np.pi, 3)
        f += rng.standard_normal() * (np.cos(kx * ax[0] + ph[0])[:, None, None]
                                      * np.cos(ky * ax[1] + ph[1])[None, :, None]
                                      * np.cos(kz * ax[2] + ph[2])[None, None, :])
    f = (f - f.min()) / (f.max() - f.min())
    return (lo + (hi - lo) * f).astype(dtype)


# ------------------------------------------------------------------ L0 ----
def case_fft(B, n, dtype, seed=0):
    sh = shape3(n)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(sh).astype(dtype)
    cdt = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    h = B.handle(n, dtype)
    xd = B.put(x)
    fh = B.empty((sh[0], sh[1], sh[2] // 2 + 1), cdt)
    h.fft_r2c(xd, fh)
    ref = O.fft_r2c(x.astype(np.float64))
    e_fwd = rel(B.get(fh), ref)
    y = B.empty(sh, dtype)
    h.fft_c2r(fh, y)
    e_rt = rel(B.get(y), x.astype(np.float64) * np.prod(sh))
    # c2r of an oracle spectrum (independent of our own r2c)
    fo = B.put(ref.astype(cdt))
    h.fft_c2r(fo, y)
    e_inv = rel(B.get(y), x.astype(np.float64) * np.prod(sh))
    h.close()
    return e_fwd, e_rt, e_inv


def case_grad_div(B, n, dtype, seed=1):
    sh = shape3(n)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(sh).astype(dtype)
    h = B.handle(n, dtype)
    xd = B.put(x)
    g = [B.empty(sh, dtype) for _ in range(3)]
    h.gradient(g[0], g[1], g[2], xd, 7)
    ref = O.gradient(x)
    errs = [rel(B.get(g[i]), ref[i]) for i in range(3)]
    # masked call leaves unrequested outputs untouched
    gy2 = B.empty(sh, dtype)
    h.gradient(None, gy2, None, xd, 2)
    errs.append(rel(B.get(gy2), ref[1]))
    d = B.empty(sh, dtype)
    vx, vy, vz = [rng.standard_normal(sh).astype(dtype) for _ in range(3)]
    h.divergence(d, B.put(vx), B.put(vy), B.put(vz))
    errs.append(rel(B.get(d), O.divergence(vx, vy, vz)))
    h.close()
    return errs


# ------------------------------------------------------------------ L1 ----
def tissue(shape, dtype):
    wm = smooth_field(shape, dtype, 11, 0.0, 0.7)
    gm = smooth_field(shape, dtype, 12, 0.0, 0.3)
    csf = (np.float64(1.0) - wm - gm).clip(0, 1).astype(dtype) * np.asarray(0.5, dtype)
    filt = ((wm > 0.1) | (gm > 0.1)).astype(dtype)
    return wm, gm, csf, filt


def case_apply_D(B, n, dtype, sinusoidal=True):
    """-> (err, err_aliased, budget).  D amplifies rounding by ~(n/2)^2, so in single precision
    two correct evaluations differ by far more than machine epsilon: the budget is the distance
    of the ORACLE's own working-precision result from its float64 result (x2, plus epsilon),
    i.e. the CUDA path must be as accurate as the reference's arithmetic."""
    sh = shape3(n)
    k = O.DiffCoef(sh, dtype)
    k64 = O.DiffCoef(sh, np.float64)
    h = B.handle(n, dtype)
    if sinusoidal:
        k.set_values_sinusoidal(1e-2)
        h.set_diffusion(B.put(k.kxx), [float(k.kxx_avg)] * 3, 1e-2)
    else:
        wm, gm, csf, filt = tissue(sh, dtype)
        k.set_values(0.05, 0.2, 0.1, wm, gm, csf, filt)
        h.set_diffusion_tissue(B.put(wm), B.put(gm), B.put(csf), 0.05, 0.2, 0.1, float(filt.sum(dtype=np.float64)))
    k64.kxx = k.kxx.astype(np.float64)
    c = smooth_field(sh, dtype, 5) if not sinusoidal else O.test_gaussian(sh[0], dtype)
    if c.shape != sh:
        c = smooth_field(sh, dtype, 5)
    ref = k.apply_D(c)
    ref64 = k64.apply_D(c.astype(np.float64))
    budget = 2.0 * rel(ref, ref64) + 50 * float(np.finfo(dtype).eps)
    if np.dtype(dtype) == np.float64:  # no wider oracle: allow eps times the (n/2)^2 amplification
        budget = float(np.finfo(np.float64).eps) * max(sh) ** 2
    cd = B.put(c)
    dc = B.empty(sh, dtype)
    h.apply_D(dc, cd)
    e1 = rel(B.get(dc), ref64)
    h.apply_D(cd, cd)  # aliasing allowed (PdeOperators.cpp:210)
    e2 = rel(B.get(cd), ref64)
    h.close()
    return e1, e2, budget


# ------------------------------------------------------------------ L2 ----
def case_K1(B, dtype, n=64, nsolves=10):
    """src/test/pdesolver.cpp:7-56 through the C ABI."""
    k = O.DiffCoef((n, n, n), dtype)
    k.set_values_sinusoidal(1e-2)
    ds = O.DiffusionSolver(k, 0.5)
    h = B.handle(n, dtype, dt_ctx=0.5)
    h.set_diffusion(B.put(k.kxx), [float(k.kxx_avg)] * 3, 1e-2)
    h.prec_factor()
    c = O.test_gaussian(n, dtype)
    cd = B.put(c)
    its_g, its_o = [], []
    for _ in range(nsolves):
        its_g.append(h.diffusion_solve(cd, 0.02))
        c = ds.solve(c, 0.02)
        its_o.append(ds.ksp_itr)
    out = B.get(cd)
    h.close()
    return float(np.sqrt(np.sum(out.astype(np.float64) ** 2))), its_g, its_o, rel(out, c)


def make_problem(n, dtype, rho_scale=8.0, k_scale=0.01, seed=3):
    sh = shape3(n)
    wm, gm, csf, filt = tissue(sh, dtype)
    k = O.DiffCoef(sh, dtype)
    k.set_values(k_scale, 0.2, 0.0, wm, gm, csf, filt)
    rho = O.reac_coef(rho_scale, 0.2, 0.0, wm, gm, csf)
    # Gaussian initial condition
    ax = [2 * np.pi * np.arange(m) / m for m in sh]
    r2 = ((ax[0] - 3.0)[:, None, None] ** 2 + (ax[1] - 3.3)[None, :, None] ** 2 + (ax[2] - 2.8)[None, None, :] ** 2)
    c0 = (0.8 * np.exp(-r2 / (2 * 0.35 ** 2))).astype(dtype)
    return dict(wm=wm, gm=gm, csf=csf, filt=filt, k=k, rho=rho, c0=c0, k_scale=k_scale, rho_scale=rho_scale)


def setup_handle(B, P, n, dtype, nt, dt):
    h = B.handle(n, dtype, dt_ctx=dt)
    dev = {key: B.put(P[key]) for key in ("wm", "gm", "csf")}
    h.set_diffusion_tissue(dev["wm"], dev["gm"], dev["csf"], P["k_scale"], 0.2, 0.0,
                           float(P["filt"].sum(dtype=np.float64)))
    h.set_reaction_tissue(dev["wm"], dev["gm"], dev["csf"], P["rho_scale"], 0.2, 0.0)
    h.prec_factor()
    h.resize_history(nt, dt)
    return h, dev


def case_forward_adjoint(B, n, dtype, nt=3, dt=0.04, adjoint_store=True, with_grad=True, order=2):
    """solveState(0) -> p_T = -(c(T) - d) -> solveAdjoint(1) -> kappa/rho gradient integrals.
    order = 1: the first-order splitting branch of solveState / solveAdjoint (PdeOperators.cpp:284-290, 395-398)."""
    sh = shape3(n)
    P = make_problem(n, dtype)
    pde = O.PdeOperatorsRD(P["k"], P["rho"], nt, dt, dt_ctx=dt, adjoint_store=adjoint_store, order=order)
    cT_ref = pde.solve_state(P["c0"], 0)
    d1 = (0.9 * cT_ref + 0.05 * P["c0"]).astype(dtype)
    pT = (-(cT_ref - d1)).astype(dtype)
    p0_ref = pde.solve_adjoint(pT, 1)

    h, dev = setup_handle(B, P, n, dtype, nt, dt)
    h.set_splitting_order(order)
    cT = B.empty(sh, dtype)
    its_s = h.solve_state(B.put(P["c0"]), cT, 0)
    res = {"its_state": (its_s, pde.ksp_state), "cT": rel(B.get(cT), cT_ref)}
    p0 = B.empty(sh, dtype)
    its_a = h.solve_adjoint(B.put(pT), p0, 1, adjoint_store)
    res["its_adj"] = (its_a, pde.ksp_adj)
    res["p0"] = rel(B.get(p0), p0_ref)
    if with_grad:
        g = h.grad_kappa_rho(dev["wm"], dev["gm"], dev["csf"])
        g_ref = O.grad_kappa_rho(pde, P["wm"], P["gm"], P["csf"])
        res["grad"] = float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-300)))
        res["grad_vals"] = (g, g_ref)
    h.close()
    return res


def case_pcg_edges(B, n, dtype):
    """Edge behaviour of DiffusionSolver::solve as the reference has it: k_scale == 0 returns at once
    (DiffusionSolver.cpp:227); a zero field converges at the iteration-0 test; a capped maxit stops
    after exactly that many iterations with the iterate PETSc would hold; a loose rtol needs fewer
    iterations than the default and still matches the oracle run with the same tolerance."""
    sh = shape3(n)
    P = make_problem(n, dtype)
    res = {}
    # (1) k_scale == 0
    h = B.handle(n, dtype, dt_ctx=0.04)
    kd = B.put(P["k"].kxx)
    h.set_diffusion(kd, [float(P["k"].kxx_avg)] * 3, 0.0)
    h.prec_factor()
    c = B.put(P["c0"])
    res["kscale0_its"] = h.diffusion_solve(c, 0.02)
    res["kscale0_unchanged"] = bool(np.array_equal(B.get(c), P["c0"]))
    h.close()
    # (2) zero field, (3) maxit, (4) loose rtol
    h, dev = setup_handle(B, P, n, dtype, nt=1, dt=0.04)
    zf = B.put(np.zeros(sh, dtype))
    res["zero_its"] = h.diffusion_solve(zf, 0.02)
    res["zero_out"] = float(np.abs(B.get(zf)).max())
    solver = O.DiffusionSolver(P["k"], dt_ctx=0.04)
    solver.prec_factor()
    full = solver.solve(P["c0"], 0.02)
    full_its = solver.ksp_itr
    cap = max(1, full_its - 1)
    solver2 = O.DiffusionSolver(P["k"], dt_ctx=0.04)
    solver2.prec_factor()
    solver2.MAXIT = cap
    capped = solver2.solve(P["c0"], 0.02)
    h.set_ksp_tolerances(maxit=cap)
    c = B.put(P["c0"])
    res["maxit"] = (h.diffusion_solve(c, 0.02), solver2.ksp_itr, cap)
    res["maxit_err"] = rel(B.get(c), capped)
    solver3 = O.DiffusionSolver(P["k"], dt_ctx=0.04)
    solver3.prec_factor()
    solver3.RTOL = 1e-2
    loose = solver3.solve(P["c0"], 0.02)
    h.set_ksp_tolerances(rtol=1e-2)
    c = B.put(P["c0"])
    res["loose"] = (h.diffusion_solve(c, 0.02), solver3.ksp_itr, full_its)
    res["loose_err"] = rel(B.get(c), loose)
    h.close()
    return res


def case_mass_effect_steps(B, n, dtype, nsteps=3, dt=0.04):
    """the RD part of PdeOperatorsMassEffect::solveState's loop
    (src/pde/PdeOperatorsMassEffect.cpp:578-631) -- coefficients refreshed from moving tissue maps
    before every step, precFactor(), full-dt diffusion solve, full-dt reaction -- through the C ABI
    against the oracle.  "Advection" is a rigid periodic shift of the maps by one voxel per step."""
    sh = shape3(n)
    P = make_problem(n, dtype)
    t = np.dtype(dtype).type
    bg = (1.0 - P["filt"]).astype(dtype)
    vt = (0.3 * P["csf"]).astype(dtype)
    seq = [tuple(np.roll(f, i, axis=i % 3).copy() for f in (bg, P["gm"], vt, P["csf"])) for i in range(nsteps)]
    rho_s, k_s, gm_r, gm_k = 8.0, 0.05, 1.0 - 0.2, 1.0 - 0.1
    k = O.DiffCoef(sh, dtype)
    k.set_values(k_s, 0.1, 0.0, P["wm"], P["gm"], P["csf"], P["filt"])   # fixes k-bar once, like the ctor path
    solver = O.DiffusionSolver(k, dt_ctx=dt)
    c_ref, its_ref = O.mass_effect_rd_steps(P["c0"], seq, k, solver, rho_s, k_s, gm_r, gm_k, dt)

    h = B.handle(n, dtype, dt_ctx=dt)
    dev = {key: B.put(P[key]) for key in ("wm", "gm", "csf")}
    h.set_diffusion_tissue(dev["wm"], dev["gm"], dev["csf"], k_s, 0.1, 0.0, float(P["filt"].sum(dtype=np.float64)))
    c = B.put(P["c0"])
    its = []
    for fields in seq:
        d = [B.put(f) for f in fields]
        h.update_reac_diff(d[0], d[1], d[2], d[3], rho_s, k_s, gm_r, gm_k)
        h.prec_factor()
        its.append(h.diffusion_solve(c, dt))
        h.reaction(c, None, dt)
    out = B.get(c)
    h.close()
    return {"its": (its, its_ref), "c": rel(out, c_ref), "moved": rel(out, P["c0"])}


def case_objective_hessian(B, n, dtype, nt=2, dt=0.04, beta=1e-3):
    """evaluateObjectiveAndGradient + Gauss-Newton Hessian product (with diffusivity inversion,
    nk = 2) through the C ABI vs the oracle's DerivativeOperatorsRD, observation mask on."""
    sh = shape3(n)
    P = make_problem(n, dtype)
    pde = O.PdeOperatorsRD(P["k"], P["rho"], nt, dt, dt_ctx=dt)
    obs = (P["wm"] > 0.2).astype(dtype)
    D = O.DerivativeOperatorsRD(pde, P["wm"], P["gm"], P["csf"], obs=obs, beta=beta)
    d1 = (0.7 * P["c0"]).astype(dtype)
    ref = D.evaluate_objective_and_gradient(P["c0"], d1)
    h, dev = setup_handle(B, P, n, dtype, nt, dt)
    obs_d = B.put(obs)
    gc0 = B.empty(sh, dtype)
    out = h.objective_gradient(B.put(P["c0"]), B.put(d1), dev["wm"], dev["gm"], dev["csf"], obs=obs_d, beta=beta, g_c0=gc0)
    res = {"J": abs(out["J"] - ref["J"]) / abs(ref["J"]), "its": (out["its"], ref["its"]),
           "g_c0": rel(B.get(gc0), ref["g_c0"]),
           "g6": float(np.max(np.abs(out["g6"] - ref["g6"]) / np.maximum(np.abs(ref["g6"]), 1e-300)))}
    P["k"].set_secondary(0.3, 0.1, 0.0, P["wm"], P["gm"], P["csf"], nk=2)
    h.set_secondary_tissue(dev["wm"], dev["gm"], dev["csf"], 0.3, 0.1, 0.0)
    c0t = (0.5 * P["c0"] * P["wm"]).astype(dtype)
    y_ref, hk_ref, its_ref = D.evaluate_hessian(c0t, True)
    y = B.empty(sh, dtype)
    hk, its = h.hessian_matvec(B.put(c0t), y, dev["wm"], dev["gm"], dev["csf"], obs=obs_d, beta=beta,
                               diffusivity_inversion=True)
    res["h_its"] = (list(its), list(its_ref))
    res["h_y"] = rel(B.get(y), y_ref)
    res["h_k"] = float(np.max(np.abs(hk - hk_ref) / np.maximum(np.abs(hk_ref), 1e-300)))
    # p-only Hessian (no diffusivity inversion)
    y2_ref, _, _ = D.evaluate_hessian(c0t, False)
    h.hessian_matvec(B.put(c0t), y, dev["wm"], dev["gm"], dev["csf"], obs=obs_d, beta=beta, diffusivity_inversion=False)
    res["h_y_ponly"] = rel(B.get(y), y2_ref)
    h.close()
    return res


def case_ensemble_batch(B, n, dtype, nt=2, dt=0.04, members=((0.01, 8.0), (0.05, 4.0), (0.002, 15.0))):
    """Ensemble handle (glia_rd_create_batch): independent members with their own (kappa, rho) advance through ONE set
    of kernel launches.  Every member's c(T), alpha(0) and PCG iteration count must be what the oracle gets for that
    member alone -- members converge at different iterations (a converged member stops, the others continue)."""
    sh = shape3(n)
    nb = len(members)
    P = make_problem(n, dtype)
    fsum = float(P["filt"].sum(dtype=np.float64))
    h = RDHandle(n, "f32" if np.dtype(dtype) == np.float32 else "f64", dt_ctx=dt, lib_path=B.lib_path, nbatch=nb,
                 **({"device": B.device} if hasattr(B, "device") else {}))
    wm, gm, csf = B.put(P["wm"]), B.put(P["gm"]), B.put(P["csf"])
    h.set_coefficients_batch(wm, gm, csf, [m[0] for m in members], 0.2, 0.0, fsum, [m[1] for m in members], 0.2, 0.0)
    h.prec_factor()
    h.resize_history(nt, dt)
    c0b = np.ascontiguousarray(np.stack([P["c0"] * (1.0 - 0.1 * i) for i in range(nb)]).astype(dtype))
    cT = B.empty((nb,) + sh, dtype)
    tot_s = h.solve_state(B.put(c0b), cT, 0)
    its_s = h.batch_iterations(True)
    cTg = B.get(cT)
    refs = []
    res = {"cT": [], "p0": [], "its_state": [], "its_adj": []}
    pTb = np.zeros((nb,) + sh, dtype)
    for i, (ks, rs) in enumerate(members):
        k = O.DiffCoef(sh, dtype)
        k.set_values(ks, 0.2, 0.0, P["wm"], P["gm"], P["csf"], P["filt"])
        rho = O.reac_coef(rs, 0.2, 0.0, P["wm"], P["gm"], P["csf"])
        pde = O.PdeOperatorsRD(k, rho, nt, dt, dt_ctx=dt)
        ref = pde.solve_state(c0b[i], 0)
        res["cT"].append(rel(cTg[i], ref))
        res["its_state"].append((its_s[i], pde.ksp_state))
        pTb[i] = (-(ref - 0.8 * c0b[i])).astype(dtype)
        refs.append(pde)
    res["sum_ok"] = (tot_s == sum(its_s))
    p0 = B.empty((nb,) + sh, dtype)
    h.solve_adjoint(B.put(pTb), p0, 1, True)
    its_a = h.batch_iterations(True)
    p0g = B.get(p0)
    for i, pde in enumerate(refs):
        ref = pde.solve_adjoint(pTb[i], 1)
        res["p0"].append(rel(p0g[i], ref))
        res["its_adj"].append((its_a[i], pde.ksp_adj))
    from glia_b200._capi import GliaRdError
    try:
        h.grad_kappa_rho(wm, gm, csf)
        res["guard"] = False
    except GliaRdError:
        res["guard"] = True
    h.close()
    return res


def case_two_snapshot(B, n, dtype, nt=2, dt=0.04, beta=1e-3):
    """two_time_points_ (DerivativeOperatorsRD.cpp:30-34, 149-153, 216-222): the t = 0 mismatch term of the
    objective and its gradient contribution, with a t = 0 observation mask of its own; the Hessian refuses."""
    sh = shape3(n)
    P = make_problem(n, dtype)
    pde = O.PdeOperatorsRD(P["k"], P["rho"], nt, dt, dt_ctx=dt)
    obs1 = (P["wm"] > 0.2).astype(dtype)
    obs0 = (P["wm"] > 0.35).astype(dtype)
    d1 = (0.7 * P["c0"]).astype(dtype)
    d0 = (0.9 * P["c0"] * P["wm"]).astype(dtype)
    D = O.DerivativeOperatorsRD(pde, P["wm"], P["gm"], P["csf"], obs=obs1, beta=beta, d0=d0, obs0=obs0)
    ref = D.evaluate_objective_and_gradient(P["c0"], d1)
    h, dev = setup_handle(B, P, n, dtype, nt, dt)
    gc0 = B.empty(sh, dtype)
    h.set_two_snapshot(B.put(d0), B.put(obs0))
    out = h.objective_gradient(B.put(P["c0"]), B.put(d1), dev["wm"], dev["gm"], dev["csf"], obs=B.put(obs1), beta=beta,
                               g_c0=gc0)
    res = {"J": abs(out["J"] - ref["J"]) / abs(ref["J"]),
           "m0": abs(out["mismatch0"] - ref["mismatch0"]) / abs(ref["mismatch0"]),
           "m0_share": ref["mismatch0"] / ref["J"], "g_c0": rel(B.get(gc0), ref["g_c0"])}
    from glia_b200._capi import GliaRdError
    try:
        h.hessian_matvec(B.put(P["c0"]), gc0, dev["wm"], dev["gm"], dev["csf"], beta=beta)
        res["hessian_refused"] = False
    except GliaRdError as e:
        res["hessian_refused"] = "two-snapshot" in str(e)
    # off again: the one-snapshot objective
    h.set_two_snapshot(None)
    D1 = O.DerivativeOperatorsRD(pde, P["wm"], P["gm"], P["csf"], obs=obs1, beta=beta)
    ref1 = D1.evaluate_objective_and_gradient(P["c0"], d1)
    out1 = h.objective_gradient(B.put(P["c0"]), B.put(d1), dev["wm"], dev["gm"], dev["csf"], obs=B.put(obs1), beta=beta)
    res["J_off"] = abs(out1["J"] - ref1["J"]) / abs(ref1["J"])
    res["m0_off"] = out1["mismatch0"]
    h.close()
    return res


# ------------------------------------------- smoother / MatProp / Phi ----
def case_smooth(B, n, dtype, sigma_factor=1.0, seed=7):
    """weierstrassSmoother through the C ABI (three 1-D symbol sweeps) vs the oracle's 3-D-FFT
    restatement; also in place and sigma = 0."""
    sh = shape3(n)
    rng = np.random.default_rng(seed)
    x = rng.random(sh).astype(dtype)
    sigma = float(np.dtype(dtype).type(sigma_factor * 2 * np.pi / sh[0]))
    ref = O.weierstrass_smoother(x.astype(np.float64), sigma)
    h = B.handle(n, dtype)
    xd, out = B.put(x), B.empty(sh, dtype)
    h.smooth(out, xd, sigma)
    e1 = rel(B.get(out), ref)
    h.smooth(xd, xd, sigma)
    e2 = rel(B.get(xd), ref)
    xd = B.put(x)
    h.smooth(out, xd, 0.0)
    e0 = rel(B.get(out), x)
    h.close()
    return e1, e2, e0


def case_mat_prop(B, n, dtype, seed=8):
    sh = shape3(n)
    rng = np.random.default_rng(seed)
    maps = {k: (rng.random(sh) * 1.2 - 0.2).astype(dtype) for k in ("gm", "wm", "vt", "csf")}
    maps["vt"] = (maps["vt"] * 1.0).astype(dtype)
    ref = O.mat_prop(maps, sh, np.dtype(dtype).type)
    h = B.handle(n, dtype)
    dev = {k: B.put(v) for k, v in maps.items()}
    bg, filt = B.empty(sh, dtype), B.empty(sh, dtype)
    fs = h.mat_prop(dev["gm"], dev["wm"], dev["vt"], dev["csf"], bg, filt)
    ok = all(np.array_equal(B.get(dev[k]), ref[k]) for k in ("gm", "wm", "vt", "csf"))
    ok = ok and np.array_equal(B.get(filt), ref["filter"]) and np.array_equal(B.get(bg), ref["bg"])
    # vt absent
    gm2, wm2 = B.put(maps["gm"]), B.put(maps["wm"])
    fs2 = h.mat_prop(gm2, wm2, None, None, None, filt)
    ref2 = O.mat_prop({"gm": maps["gm"], "wm": maps["wm"]}, sh, np.dtype(dtype).type)
    ok = ok and np.array_equal(B.get(filt), ref2["filter"]) and fs2 == float(ref2["filter"].sum(dtype=np.float64))
    h.close()
    return ok, fs, float(ref["filter"].sum(dtype=np.float64))


def case_phi(B, n, dtype, sigma_factor=2.0):
    """Phi::apply / applyTranspose (on-the-fly mode) through the C ABI vs the oracle, with a
    filter, a zero coefficient (skipped basis function) and the all-zero p case."""
    sh = shape3(n)
    wm, gm, csf, filt = tissue(sh, dtype)
    sig = 2 * np.pi / sh[0] * sigma_factor
    ctr = [(3.0, 3.3, 2.8), (2.4, 3.0, 3.4), (3.6, 2.7, 3.0)]
    p = [1.0, 0.0, 0.6]
    ref = O.phi_apply(p, ctr, sig, filt, 1.0)
    h = B.handle(n, dtype)
    h.phi_set(ctr, sig, B.put(filt), 2 * np.pi / sh[0])
    out = B.empty(sh, dtype)
    h.phi_apply(out, p)
    e_apply = rel(B.get(out), ref)
    rng = np.random.default_rng(4)
    f = rng.standard_normal(sh).astype(dtype)
    pt = h.phi_apply_transpose(B.put(f))
    pt_ref = O.phi_apply_transpose(f, ctr, sig, filt, 1.0)
    e_t = float(np.max(np.abs(pt - pt_ref) / np.max(np.abs(pt_ref))))
    # adjointness <Phi p, f> = <p, Phi^T f> holds when every basis function takes part
    p2 = [0.3, -0.8, 0.5]
    h.phi_apply(out, p2)
    lhs = float(np.sum(B.get(out).astype(np.float64) * f.astype(np.float64)))
    rhs = float(np.dot(p2, pt))
    e_adj = abs(lhs - rhs) / max(abs(lhs), 1e-300)
    h.phi_apply(out, [0.0, 0.0, 0.0])
    zero_ok = not np.any(B.get(out))
    h.close()
    return e_apply, e_t, e_adj, zero_ok


def case_K2_K3(B, dtype, which):
    """The reference's own pins for c(0) = Phi p (src/test/simulator.cpp:41-42, 94-95) computed by
    the CUDA path end to end from the committed test data: smooth the tissue maps, MatProp filter,
    Phi::apply.  -> (||c0||, expected, rel. error against the oracle's c0)."""
    from golden import fixtures as FX
    n = 64
    t = np.dtype(dtype).type
    h = B.handle(n, dtype)
    sig_atlas = float(t(2 * np.pi / n))
    if which == "K2":
        maps = O.split_segmentation(FX.atlas_labels().astype(dtype), (6, 5, 7, 8), dtype)
        dev = {k: B.put(maps[k]) for k in ("gm", "wm", "vt", "csf")}
        for k in ("gm", "wm", "vt", "csf"):   # SolverInterface::readAtlas order
            h.smooth(dev[k], dev[k], sig_atlas)
        sigma_phi, expected = 2 * np.pi / 64, 4.09351
        ref = FX.brain_problem(dtype)["c0"]
    else:
        dev = {"wm": B.put(FX.sinusoid(dtype)), "gm": None, "vt": None, "csf": None}
        sigma_phi, expected = 4 * 2 * np.pi / 64, 22.0161
        ref = FX.sinusoid_c0(dtype)
    filt = B.empty((n, n, n), dtype)
    h.mat_prop(dev["gm"], dev["wm"], dev["vt"], dev["csf"], None, filt)
    h.phi_set([FX.TIL], sigma_phi, filt, sig_atlas)
    c0 = B.empty((n, n, n), dtype)
    h.phi_apply(c0, [1.0])
    got = B.get(c0)
    h.close()
    return float(np.sqrt(np.sum(got.astype(np.float64) ** 2))), expected, rel(got, ref)
