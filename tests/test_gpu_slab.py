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
        # global relative L2 of the distributed fields (root-sum-square of the ranks' shares)
        assert max(_slab.combine(res, "grad")) < tol and _slab.combine(res, "div") < tol, (world, res)
        assert _slab.combine(res, "applyD") <= max(res[0]["applyD_budget"], tol), (world, res)


@pytest.mark.parametrize("n,nt,dtype", [(64, 3, "float64"), (128, 3, "float32"), (256, 2, "float32")])
def test_slab_forward_adjoint_gradient_gpu(cuda_lib, n, nt, dtype):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in _worlds():
        res = _slab.run(world, "cuda", cuda_lib, "forward_adjoint", n=n, dtype=dtype, nt=nt, timeout=900)
        tol = Cs.TOL[np.dtype(dtype)]
        for r in res:
            assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], (world, r)
            assert r["grad"] < 20 * tol, (world, r)
        for key in ("cT", "p0", "fa_cT", "fa_p0"):   # global relative L2 (north_star's measure)
            assert _slab.combine(res, key) < tol, (world, key, res)
        assert len({tuple(r["fa_its"]) for r in res}) == 1
