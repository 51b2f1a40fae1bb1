"""Small driver for ncu captures: one 256^3 (or --n) Crank-Nicolson diffusion solve on the
synthetic atlas (every sweep kernel of the hot path launches at least once), optionally a few
forward time steps.  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -c 40 -o gpurun_out/prof python scripts/prof_driver.py
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glia_b200 import synthetic as S  # noqa: E402
from glia_b200.rd import RDHandle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--precision", default="f32")
ap.add_argument("--solves", type=int, default=1)
ap.add_argument("--nt", type=int, default=0)
a = ap.parse_args()
dtype = np.float32 if a.precision == "f32" else np.float64
atlas = S.make_atlas(a.n, 0, dtype)
c0 = S.make_initial_condition(atlas, 0, dtype=dtype)
dev = torch.device("cuda:0")
put = lambda x: torch.from_numpy(x).to(dev)
wm, gm, csf, c = put(atlas["wm"]), put(atlas["gm"]), put(atlas["csf"]), put(c0)
torch.cuda.synchronize()
h = RDHandle(a.n, a.precision, 0, dt_ctx=0.04)
h.set_diffusion_tissue(wm, gm, csf, 0.01, 0.0, 0.0, float(atlas["filter"].sum(dtype=np.float64)))
h.set_reaction_tissue(wm, gm, csf, 8.0, 0.0, 0.0)
h.prec_factor()
for _ in range(a.solves):
    its = h.diffusion_solve(c, 0.02)
    print("ksp its", its)
if a.nt:
    h.resize_history(a.nt, 0.04)
    out = torch.empty_like(c)
    print("state its", h.solve_state(c, out, 0))
h.close()
