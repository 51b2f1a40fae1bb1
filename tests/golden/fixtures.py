"""Loaders of the committed golden fixtures + the config-1 problem built from them."""
import os

import numpy as np

from oracle import rd_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def atlas_labels():
    return np.load(os.path.join(HERE, "atlas_labels_64.npz"))["labels"]


def sinusoid(dtype=np.float32):
    x = 2 * np.pi * np.arange(64) / 64
    f = (0.5 + 0.5 * np.sin(4 * x)[:, None, None] * np.sin(4 * x)[None, :, None] * np.sin(4 * x)[None, None, :])
    f = f.astype(np.float32)
    z = np.load(os.path.join(HERE, "sinusoid_64.npz"))
    f[tuple(z["idx"].astype(np.int64).T)] = z["val"]
    return f.astype(dtype)


TIL = tuple(2 * np.pi / 256 * v for v in (137, 169, 96))   # user_cms of config/test_forward_config.txt


def brain_problem(dtype):
    """config/test_forward_config.txt with model = 1: atlas.nc split (wm 6, gm 5, vt 7, csf 8),
    smoothing_factor 1, one Gaussian at TIL with sigma 2 pi / 64, rho 8, kappa 0.01, nt 25, dt 0.04."""
    n = 64
    maps = O.split_segmentation(atlas_labels().astype(dtype), (6, 5, 7, 8), dtype)
    atlas = O.read_atlas(maps, n, 1.0, 1.0)
    m = O.mat_prop(atlas, (n, n, n), dtype)
    c0 = O.phi_apply([1.0], [TIL], 2 * np.pi / 64, m["filter"], 1.0)
    k = O.DiffCoef((n, n, n), dtype)
    k.set_values(0.01, 0.0, 0.0, m["wm"], m["gm"], m["csf"], m["filter"])
    rho = O.reac_coef(8.0, 0.0, 0.0, m["wm"], m["gm"], m["csf"])
    return dict(m=m, c0=c0, k=k, rho=rho, nt=25, dt=0.04, n=n)


def sinusoid_c0(dtype):
    """config/test_forward_sin_config.txt: wm = sinusoid.nc only, sigma_factor 4."""
    n = 64
    atlas = O.read_atlas({"wm": sinusoid(dtype)}, n, 1.0, 0.0)  # the test sets smoothing_factor_atlas_ = 0 (simulator.cpp:24)
    m = O.mat_prop(atlas, (n, n, n), dtype)
    return O.phi_apply([1.0], [TIL], 4 * 2 * np.pi / 64, m["filter"], 1.0)


FWD = os.path.join(HERE, "rd_forward_64.npz")


def write_forward_fixture():
    """Oracle (float64 and float32) run of config 1; stored so that the gpu tests can compare
    against numbers generated in the build container."""
    out = {}
    for name, dtype in (("f64", np.float64), ("f32", np.float32)):
        P = brain_problem(dtype)
        pde = O.PdeOperatorsRD(P["k"], P["rho"], P["nt"], P["dt"], dt_ctx=P["dt"])
        cT = pde.solve_state(P["c0"], 0)
        its_s = pde.ksp_state
        pT = (-(cT - (0.5 * cT).astype(dtype))).astype(dtype)
        p0 = pde.solve_adjoint(pT, 1)
        g = O.grad_kappa_rho(pde, P["m"]["wm"], P["m"]["gm"], P["m"]["csf"])
        nrm = lambda a: float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
        out.update({f"{name}_c0_norm": nrm(P["c0"]), f"{name}_cT_norm": nrm(cT), f"{name}_p0_norm": nrm(p0),
                    f"{name}_its_state": its_s, f"{name}_its_adj": pde.ksp_adj, f"{name}_grad": g,
                    f"{name}_cT_line": cT[34, 42, :].copy(), f"{name}_p0_line": p0[34, 42, :].copy(),
                    f"{name}_cT_max": float(cT.max()), f"{name}_cT_min": float(cT.min())})
        print(name, {k: v for k, v in out.items() if k.startswith(name) and np.ndim(v) == 0})
    np.savez_compressed(FWD, **out)
