"""CPU coverage of the N > 1 (slab-decomposed) path: world_size 2, 4 and 8, gloo rendezvous on
127.0.0.1, the product sources compiled against the SIMT emulator, arenas in POSIX shared
memory so that the ranks really read and write each other's fields and spin on each other's
flags, exactly like the GPU ranks do over NVLink."""
import numpy as np
import pytest

import _cases as Cs
import _slab


@pytest.mark.parametrize("world,dtype", [(2, "float32"), (4, "float64"), (8, "float64")])
def test_slab_operators(emu_lib, world, dtype):
    res = _slab.run(world, "emu", emu_lib, "operators", n=32, dtype=dtype)
    tol = 1e-12 if dtype == "float64" else 5e-6
    assert max(_slab.combine(res, "grad")) < tol and _slab.combine(res, "div") < tol, res
    assert _slab.combine(res, "applyD") <= max(res[0]["applyD_budget"], tol), res


@pytest.mark.parametrize("world,dtype", [(2, "float64")])   # float32 and worlds 2/4/8 run on the GPU box
def test_slab_forward_adjoint_gradient(emu_lib, world, dtype):
    res = _slab.run(world, "emu", emu_lib, "forward_adjoint", n=32, dtype=dtype, nt=2)
    tol = Cs.TOL[np.dtype(dtype)]
    for r in res:
        assert r["its_state"][0] == r["its_state"][1] and r["its_adj"][0] == r["its_adj"][1], r
        assert r["grad"] < 20 * tol, r
        assert r["fa_its"] == (r["its_state"][0], r["its_adj"][0]), r
    for key in ("cT", "p0", "fa_cT", "fa_p0"):
        assert _slab.combine(res, key) < tol, (key, res)
    # every rank took the same control decisions
    assert len({tuple(r["fa_its"]) for r in res}) == 1
