"""The C++ host mirror (include/glia_rd_host.hpp) and its reference-style tests (tests/cpp/):
compiled everywhere, run on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.bin")


def _build(cuda_lib):
    libdir = os.path.dirname(cuda_lib)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    deps = [SRC, os.path.join(ROOT, "include", "glia_rd_host.hpp"), os.path.join(ROOT, "include", "glia_rd.h"), cuda_lib]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    cmd = [nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-I" + os.path.join(ROOT, "include"), "-o", EXE, SRC,
           "-L" + libdir, "-lglia_rd", "-Xlinker", "-rpath", "-Xlinker", libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def test_cpp_host_mirror_compiles_and_links(cuda_lib):
    assert os.path.exists(_build(cuda_lib))


@pytest.mark.gpu
def test_cpp_reference_style_tests(cuda_lib, torch_cuda):
    """K1 (src/test/pdesolver.cpp:46-48) and K4 (src/test/grad.cpp:74-77) through the C++ classes."""
    exe = _build(cuda_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout
