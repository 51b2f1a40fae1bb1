"""GPU parity of the slab-decomposed path (needs >= 2 CUDA devices on the box; skipped on a
single-GPU box, where tests/test_slab_emu.py's world_size-2 run covers the same code on the CPU).
One process per GPU, CUDA IPC arenas, the x sweeps on peer memory over NVLink."""
import numpy as np
import pytest

import _cases as Cs
import _slab

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worlds():
    """world sizes to run: every power of two the box holds, or the list in GLIA_SLAB_WORLDS
    (e.g. "8" on an 8-GPU box, to keep the call short)."""
    import os
    n = _ngpu()
    want = os.environ.get("GLIA_SLAB_WORLDS")
    cand = [int(w) for w in want.split(",")] if want else [2, 4, 8]
    return [w for w in cand if w <= n]


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_slab_operators_gpu(cuda_lib, dtype):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in _worlds():
        res = _slab.run(world, "cuda", cuda_lib, "operators", n=128, dtype=dtype)
        tol = 1e-11 if dtype == "float64" else 5e-6
        for r in res:
            assert max(r["grad"]) < tol and r["div"] < tol, (world, r)
            assert r["applyD"] <= max(r["applyD_budget"], tol), (world, r)


@pytest.mark.parametrize("n,nt,dtype", [(64, 3, "float64"), (128, 3, "float32"), (256, 2, "float32")])
def test_slab_forward_adjoint_gradient_gpu(cuda_lib, n, nt, dtype):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in _worlds():
        res = _slab.run(world, "cuda", cuda_lib, "forward_adjoint", n=n, dtype=dtype, nt=nt, timeout=900)
        tol = Cs.TOL[np.dtype(dtype)]
        for r in res:
            assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], (world, r)
            assert r["cT"] < tol and r["p0"] < tol, (world, r)
            assert r["grad"] < 20 * tol, (world, r)
            assert r["fa_cT"] < tol and r["fa_p0"] < tol, (world, r)
        assert len({tuple(r["fa_its"]) for r in res}) == 1
