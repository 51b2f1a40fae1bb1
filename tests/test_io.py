"""NetCDF classic reader / writer and the segmentation split through the C ABI,
on the emulator build (host memory stands in for device memory; the code under test is host C++)."""
import os

import numpy as np
import pytest
import scipy.io

import _cases as Cs
from oracle import rd_oracle as O

REF_ATLAS = "/root/reference/testdata/atlas.nc"   # only in the build container


@pytest.fixture(scope="module")
def B(emu_lib):
    return Cs.NumpyBackend(emu_lib)


def _write_nc(path, a, version, nctype):
    f = scipy.io.netcdf_file(path, "w", version=version)
    for name, m in zip("xyz", a.shape):
        f.createDimension(name, m)
    f.history = "written by scipy"            # a global attribute the reader must skip
    v = f.createVariable("other", "i", ("x",))  # a variable before "data"
    v[:] = np.arange(a.shape[0], dtype=np.int32)
    d = f.createVariable("data", nctype, ("x", "y", "z"))
    d.units = "1"                              # a variable attribute
    d[:] = a
    f.close()


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("nctype,np_in", [("d", np.float64), ("f", np.float32), ("h", np.int16), ("i", np.int32)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_data_in_reads_what_scipy_writes(B, tmp_path, version, nctype, np_in, dtype):
    sh = (32, 64, 32)
    rng = np.random.default_rng(3)
    a = (rng.standard_normal(sh) * 40).astype(np_in)
    p = str(tmp_path / "in.nc")
    _write_nc(p, a, version, nctype)
    h = B.handle(sh, dtype)
    out = B.empty(sh, dtype)
    h.data_in(p, out)
    assert np.array_equal(B.get(out), a.astype(dtype))
    h.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_data_out_is_read_by_scipy_and_round_trips(B, tmp_path, dtype):
    sh = (32, 32, 64)
    a = np.random.default_rng(4).standard_normal(sh).astype(dtype)
    p = str(tmp_path / "out.nc")
    h = B.handle(sh, dtype)
    h.data_out(p, B.put(a))
    f = scipy.io.netcdf_file(p, "r", mmap=False)
    assert f.version_byte == 2 and dict(f.dimensions) == {"x": 32, "y": 32, "z": 64}
    assert getattr(f, "CDF-5 mode") == 0
    got = f.variables["data"].data
    assert got.dtype == np.dtype(dtype).newbyteorder(">") and np.array_equal(got, a)
    f.close()
    back = B.empty(sh, dtype)
    h.data_in(p, back)
    assert np.array_equal(B.get(back), a)
    h.close()


def test_data_in_errors(B, tmp_path):
    h = B.handle(32, np.float32)
    out = B.empty((32, 32, 32), np.float32)
    with pytest.raises(Exception, match="cannot open"):
        h.data_in(str(tmp_path / "missing.nc"), out)
    p = str(tmp_path / "wrong.nc")
    _write_nc(p, np.zeros((32, 32, 64), np.float32), 1, "f")
    with pytest.raises(Exception, match="shape"):
        h.data_in(p, out)
    (tmp_path / "junk.nc").write_bytes(b"HDF5 is not classic netcdf")
    with pytest.raises(Exception, match="classic"):
        h.data_in(str(tmp_path / "junk.nc"), out)
    h.close()


def test_split_segmentation_matches_oracle(B):
    from golden import fixtures as FX
    seg = FX.atlas_labels().astype(np.float32)
    ref = O.split_segmentation(seg, (6, 5, 7, 8), np.float32)
    h = B.handle(64, np.float32)
    out = {k: B.empty(seg.shape, np.float32) for k in ("wm", "gm", "vt", "csf")}
    h.split_segmentation(B.put(seg), (6, 5, 7, 8), out["wm"], out["gm"], out["vt"], out["csf"])
    for k in out:
        assert np.array_equal(B.get(out[k]), ref[k]), k
    # a label <= 0 leaves the map empty; null outputs are skipped
    h.split_segmentation(B.put(seg), (6, 5, 7, 0), out["wm"], None, None, out["csf"])
    assert not np.any(B.get(out["csf"])) and np.array_equal(B.get(out["wm"]), ref["wm"])
    h.close()


@pytest.mark.skipif(not os.path.exists(REF_ATLAS), reason="reference test data only exists in the build container")
def test_reference_atlas_file_reads_to_the_committed_labels(B):
    """testdata/atlas.nc (CDF-1, NC_DOUBLE) through glia_rd_data_in equals the golden label fixture
    that tests/golden/make_golden.py extracted from the same file with scipy."""
    from golden import fixtures as FX
    h = B.handle(64, np.float64)
    out = B.empty((64, 64, 64), np.float64)
    h.data_in(REF_ATLAS, out)
    assert np.array_equal(B.get(out), FX.atlas_labels().astype(np.float64))
    h.close()
