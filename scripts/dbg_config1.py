"""debug: config-1 gradient on the GPU vs the golden fixture"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as Cs
from golden import fixtures as FX
import __graft_entry__ as g
B = Cs.TorchBackend(g.LIB)
z = np.load(FX.FWD)
for name, dtype in (("f32", np.float32), ("f64", np.float64)):
    P = FX.brain_problem(dtype)
    n, nt, dt, m = P["n"], P["nt"], P["dt"], P["m"]
    h = B.handle(n, dtype, dt_ctx=dt)
    wm, gm, csf = B.put(m["wm"]), B.put(m["gm"]), B.put(m["csf"])
    h.set_diffusion_tissue(wm, gm, csf, 0.01, 0.0, 0.0, float(m["filter"].sum(dtype=np.float64)))
    h.set_reaction_tissue(wm, gm, csf, 8.0, 0.0, 0.0)
    h.prec_factor(); h.resize_history(nt, dt)
    cT = B.empty((n, n, n), dtype)
    its_s = h.solve_state(B.put(P["c0"]), cT, 0)
    cTh = B.get(cT)
    pT = (-(cTh - (0.5 * cTh).astype(dtype))).astype(dtype)
    p0 = B.empty((n, n, n), dtype)
    its_a = h.solve_adjoint(B.put(pT), p0, 1, True)
    gg = h.grad_kappa_rho(wm, gm, csf)
    print(name, "its", its_s, its_a, int(z[f"{name}_its_state"]), int(z[f"{name}_its_adj"]))
    print(name, "gpu ", gg)
    print(name, "gold", z[f"{name}_grad"], "f64 gold", z["f64_grad"])
    print(name, "rel ", np.abs(gg - z[f"{name}_grad"]) / np.abs(z[f"{name}_grad"]))
    h.close()
