"""pytest configuration: the `gpu` marker, import paths, and the two native builds.

* `emu_lib`  -- tests/emu/libglia_rd_emu.so, the product sources compiled by g++ against the
                test-only SIMT emulator (CPU tests of kernel index logic and host drivers);
* `cuda_lib` -- glia_b200/lib/libglia_rd.so, the nvcc sm_100a product (gpu tests; also loaded
                without a GPU by the ABI test, which makes no compute call).
"""
import os
import sys

import pytest


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    import __graft_entry__ as g
    return g.build_emulator()


@pytest.fixture(scope="session")
def cuda_lib():
    import __graft_entry__ as g
    return g.build_cuda()


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu test collected on a machine without CUDA")
    return torch
