"""Generates the committed golden fixtures from the reference's own test data.  Run HERE (the
reference tree is not present on the GPU box):  python tests/golden/make_golden.py

  atlas_labels_64.npz   testdata/atlas.nc (64^3 label map, labels {0,5,6,7,8}) as uint8
  sinusoid_64.npz       testdata/sinusoid.nc == float32(0.5 + 0.5 sin4x sin4y sin4z) except in a
                        516-voxel ball; stored as the exception list (indices + values)
  rd_forward_64.npz     oracle outputs for config 1 (test_forward_config.txt with model=1):
                        norms and a z-line of c(0), c(T), alpha(0), PCG iteration counts
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/testdata"


def main():
    nc = sio.netcdf_file(os.path.join(REF, "atlas.nc"), "r", mmap=False)
    seg = np.array(nc.variables["data"].data)
    assert seg.shape == (64, 64, 64)
    assert set(np.unique(seg)) <= {0.0, 5.0, 6.0, 7.0, 8.0}
    np.savez_compressed(os.path.join(HERE, "atlas_labels_64.npz"), labels=seg.astype(np.uint8))

    nc = sio.netcdf_file(os.path.join(REF, "sinusoid.nc"), "r", mmap=False)
    d = np.array(nc.variables["data"].data, dtype=np.float32)
    x = 2 * np.pi * np.arange(64) / 64
    f = (0.5 + 0.5 * np.sin(4 * x)[:, None, None] * np.sin(4 * x)[None, :, None] * np.sin(4 * x)[None, None, :])
    f = f.astype(np.float32)
    idx = np.argwhere(f != d)
    np.savez_compressed(os.path.join(HERE, "sinusoid_64.npz"), idx=idx.astype(np.int16),
                        val=d[tuple(idx.T)].astype(np.float32))
    print("atlas + sinusoid fixtures written;", len(idx), "sinusoid exceptions")

    from tests.golden import fixtures as FX
    FX.write_forward_fixture()


if __name__ == "__main__":
    main()
