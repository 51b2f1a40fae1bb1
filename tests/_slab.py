"""World-size-G driver of the slab-decomposed (multi-GPU) path, shared by the CPU tests
(gloo + the SIMT-emulator build, 'device' memory = POSIX shared memory) and the GPU tests
(gloo for the 64-byte handle exchange, one CUDA device per rank, CUDA IPC arenas).

Every rank builds the same global problem, keeps its x-slab, runs the collective calls of the
C ABI, and compares its slab of the result with the matching slab of the CPU oracle's answer.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import socket
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def slab(a, rank, world):
    m = a.shape[0] // world
    return np.ascontiguousarray(a[rank * m:(rank + 1) * m])


def slab_err(got, ref_global, rank, world):
    """This rank's share of the GLOBAL relative L2 error: ||got - ref[slab]|| / ||ref||.  The global
    error of the distributed field is the root of the sum of squares over the ranks (`combine`).
    (A per-slab relative error is meaningless on slabs that only hold the tails of the field: at
    world 8 the outer slabs of a centred Gaussian are ~1e-10 of its peak, where rounding noise of the
    size of machine epsilon times the GLOBAL scale is 100 % of the local norm.)"""
    wide = np.float64
    d = np.asarray(got, wide) - slab(ref_global, rank, world).astype(wide)
    return float(np.linalg.norm(d.ravel()) / max(np.linalg.norm(np.asarray(ref_global, wide).ravel()), 1e-300))


def combine(res, key):
    """root-sum-square over ranks of a slab_err entry (scalar or list)."""
    v = [np.atleast_1d(np.asarray(r[key], dtype=np.float64)) for r in res]
    out = np.sqrt(np.sum(np.stack(v) ** 2, axis=0))
    return float(out[0]) if out.size == 1 else [float(x) for x in out]


def _rank_main(rank, world, port, kind, lib_path, case, kw, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)

        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out

        import _cases as Cs
        if kind == "emu":
            B = Cs.NumpyBackend(lib_path)
            device = 0
        else:
            torch.cuda.set_device(rank)
            B = Cs.TorchBackend(lib_path, device=rank)
            device = rank
        res = CASES[case](B, rank, world, device, all_gather, **kw)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", res))
    except Exception:
        q.put((rank, "error", traceback.format_exc()))


def run(world, kind, lib_path, case, timeout=600, **kw):
    """-> list of per-rank result dicts (raises if any rank failed)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, kind, lib_path, case, kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            rank, status, res = q.get(timeout=timeout)
            if status != "ok":
                raise RuntimeError(f"rank {rank} failed:\n{res}")
            out[rank] = res
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.terminate()
    return [out[r] for r in range(world)]


# ------------------------------------------------------------------ cases ----
def _handle(B, n, dtype, rank, world, device, all_gather, dt_ctx=0.5):
    from glia_b200.rd import RDHandle
    return RDHandle(n, "f32" if np.dtype(dtype) == np.float32 else "f64", device=device, dt_ctx=dt_ctx,
                    lib_path=B.lib_path, rank=rank, nranks=world, all_gather=all_gather)


def case_operators(B, rank, world, device, all_gather, n=32, dtype="float64"):
    """gradient, divergence, applyD on slabs vs the oracle's global answer."""
    import _cases as Cs
    from oracle import rd_oracle as O
    dtype = np.dtype(dtype).type
    sh = (n, n, n)
    lsh = (n // world, n, n)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(sh).astype(dtype)
    v = [rng.standard_normal(sh).astype(dtype) for _ in range(3)]
    h = _handle(B, n, dtype, rank, world, device, all_gather)
    S = lambda a: slab(a, rank, world)
    res = {}
    g = [B.empty(lsh, dtype) for _ in range(3)]
    h.gradient(g[0], g[1], g[2], B.put(S(x)), 7)
    ref = O.gradient(x)
    res["grad"] = [slab_err(B.get(g[i]), ref[i], rank, world) for i in range(3)]
    d = B.empty(lsh, dtype)
    h.divergence(d, B.put(S(v[0])), B.put(S(v[1])), B.put(S(v[2])))
    res["div"] = slab_err(B.get(d), O.divergence(*v), rank, world)
    # applyD with tissue coefficients (global sums through the rank all-reduce)
    wm, gm, csf, filt = Cs.tissue(sh, dtype)
    k = O.DiffCoef(sh, dtype)
    k.set_values(0.05, 0.2, 0.1, wm, gm, csf, filt)
    k64 = O.DiffCoef(sh, np.float64)
    k64.kxx = k.kxx.astype(np.float64)
    h.set_diffusion_tissue(B.put(S(wm)), B.put(S(gm)), B.put(S(csf)), 0.05, 0.2, 0.1, float(filt.sum(dtype=np.float64)))
    c = Cs.smooth_field(sh, dtype, 5)
    refD = k64.apply_D(c.astype(np.float64))
    dc = B.empty(lsh, dtype)
    h.apply_D(dc, B.put(S(c)))
    res["applyD"] = slab_err(B.get(dc), refD, rank, world)
    res["applyD_budget"] = (2.0 * Cs.rel(k.apply_D(c), refD) + 50 * float(np.finfo(dtype).eps)
                            if np.dtype(dtype) == np.float32 else float(np.finfo(np.float64).eps) * n ** 2)
    h.close()
    return res


def case_forward_adjoint(B, rank, world, device, all_gather, n=32, dtype="float64", nt=2, dt=0.04, with_grad=True):
    """solveState(0) -> p_T -> solveAdjoint(1) -> kappa/rho gradient on slabs vs the oracle."""
    import _cases as Cs
    from oracle import rd_oracle as O
    dtype = np.dtype(dtype).type
    sh = (n, n, n)
    lsh = (n // world, n, n)
    S = lambda a: slab(a, rank, world)
    P = Cs.make_problem(n, dtype)
    pde = O.PdeOperatorsRD(P["k"], P["rho"], nt, dt, dt_ctx=dt)
    cT_ref = pde.solve_state(P["c0"], 0)
    d1 = (0.9 * cT_ref + 0.05 * P["c0"]).astype(dtype)
    pT = (-(cT_ref - d1)).astype(dtype)
    p0_ref = pde.solve_adjoint(pT, 1)

    h = _handle(B, n, dtype, rank, world, device, all_gather, dt_ctx=dt)
    dev = {key: B.put(S(P[key])) for key in ("wm", "gm", "csf")}
    h.set_diffusion_tissue(dev["wm"], dev["gm"], dev["csf"], P["k_scale"], 0.2, 0.0, float(P["filt"].sum(dtype=np.float64)))
    h.set_reaction_tissue(dev["wm"], dev["gm"], dev["csf"], P["rho_scale"], 0.2, 0.0)
    h.prec_factor()
    h.resize_history(nt, dt)
    cT = B.empty(lsh, dtype)
    its_s = h.solve_state(B.put(S(P["c0"])), cT, 0)
    res = {"its_state": (its_s, pde.ksp_state), "cT": slab_err(B.get(cT), cT_ref, rank, world)}
    p0 = B.empty(lsh, dtype)
    its_a = h.solve_adjoint(B.put(S(pT)), p0, 1, True)
    res["its_adj"] = (its_a, pde.ksp_adj)
    res["p0"] = slab_err(B.get(p0), p0_ref, rank, world)
    if with_grad:
        g = h.grad_kappa_rho(dev["wm"], dev["gm"], dev["csf"])
        g_ref = O.grad_kappa_rho(pde, P["wm"], P["gm"], P["csf"])
        res["grad"] = float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-300)))
    # the fused device entry (what bench.py times)
    cT2, p02 = B.empty(lsh, dtype), B.empty(lsh, dtype)
    ks, ka = h.forward_adjoint(B.put(S(P["c0"])), B.put(S(d1)), cT2, p02)
    res["fa_its"] = (ks, ka)
    res["fa_cT"] = slab_err(B.get(cT2), cT_ref, rank, world)
    res["fa_p0"] = slab_err(B.get(p02), p0_ref, rank, world)
    h.close()
    return res


CASES = {"operators": case_operators, "forward_adjoint": case_forward_adjoint}


def case_debug_steps(B, rank, world, device, all_gather, n=64, dtype="float64", nt=3, dt=0.04):
    """Localises a slab-path discrepancy: one Strang step taken apart (diffusion solve, reaction,
    diffusion solve) and compared with the oracle after every piece; reports rel. L2 and the
    position of the largest deviation in this rank's slab."""
    import _cases as Cs
    from oracle import rd_oracle as O
    dtype = np.dtype(dtype).type
    lsh = (n // world, n, n)
    S = lambda a: slab(a, rank, world)
    P = Cs.make_problem(n, dtype)
    h = _handle(B, n, dtype, rank, world, device, all_gather, dt_ctx=dt)
    dev = {key: B.put(S(P[key])) for key in ("wm", "gm", "csf")}
    h.set_diffusion_tissue(dev["wm"], dev["gm"], dev["csf"], P["k_scale"], 0.2, 0.0, float(P["filt"].sum(dtype=np.float64)))
    h.set_reaction_tissue(dev["wm"], dev["gm"], dev["csf"], P["rho_scale"], 0.2, 0.0)
    h.prec_factor()
    solver = O.DiffusionSolver(P["k"], dt_ctx=dt)
    solver.prec_factor()
    res = {"steps": []}

    def cmp(tag, got, ref):
        g = B.get(got).astype(np.float64)
        r = S(ref).astype(np.float64)
        d = np.abs(g - r)
        idx = np.unravel_index(int(np.argmax(d)), d.shape)
        res["steps"].append((tag, Cs.rel(g, r), float(d.max()), tuple(int(i) for i in idx),
                             int((d > 1e-9 * np.abs(r).max()).sum())))

    # operators first
    c = P["c0"]
    dc = B.empty(lsh, dtype)
    h.apply_D(dc, B.put(S(c)))
    cmp("applyD", dc, P["k"].apply_D(c))
    cd = B.put(S(c))
    for i in range(nt):
        its = h.diffusion_solve(cd, dt / 2)
        c = solver.solve(c, dt / 2)
        cmp(f"diff{i}a its={its}/{solver.ksp_itr}", cd, c)
        h.reaction(cd, None, dt)
        c = O.reaction_nonlinear(c, P["rho"], dt)
        cmp(f"reac{i}", cd, c)
        its = h.diffusion_solve(cd, dt / 2)
        c = solver.solve(c, dt / 2)
        cmp(f"diff{i}b its={its}/{solver.ksp_itr}", cd, c)
    h.close()
    return res


CASES["debug_steps"] = case_debug_steps
