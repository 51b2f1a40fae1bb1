"""Edge behaviour of the PCG solve on the CUDA path (same cases as tests/test_emu_kernels.py runs on the
emulator).  Sorted last on purpose: these were added after the round's last GPU window, so under
`pytest -x` a surprise here cannot mask the parity suite proper."""
import numpy as np
import pytest

import _cases as Cs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B(cuda_lib, torch_cuda):
    return Cs.TorchBackend(cuda_lib)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pcg_edge_cases(B, dtype):
    r = Cs.case_pcg_edges(B, 64, dtype)
    assert r["kscale0_its"] == 0 and r["kscale0_unchanged"], r
    assert r["zero_its"] == 0 and r["zero_out"] == 0.0, r
    assert r["maxit"][0] == r["maxit"][1] == r["maxit"][2], r
    assert r["maxit_err"] < Cs.TOL[np.dtype(dtype)], r
    assert r["loose"][0] == r["loose"][1] and r["loose"][0] <= r["loose"][2], r
    assert r["loose_err"] < Cs.TOL[np.dtype(dtype)], r
