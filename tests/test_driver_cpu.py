"""CPU tests of the host driver pieces that need no GPU: the HDF5 writer/reader round trip and
the output layout of /root/reference/src/SpinED.hs:267-294,362-368,406-410."""
import os
import struct

import numpy as np
import pytest

from spin_ed_b200 import hdf5


def test_hdf5_round_trip_layout(tmp_path):
    path = str(tmp_path / "out.h5")
    reps = np.arange(16, dtype=np.uint64)
    with hdf5.File(path, "a") as f:
        for g in ("/basis", "/hamiltonian", "/observables", "/_workspace"):
            f.create_group(g)
    with hdf5.File(path, "a") as f:
        assert f.exists("/basis") and not f.exists("/basis/representatives")
        f.write_dataset("/basis/representatives", reps)
    evecs = np.asfortranarray(np.random.default_rng(0).standard_normal((16, 2)).astype(np.float32))
    with hdf5.File(path, "a") as f:
        f.write_dataset("/hamiltonian/eigenvalues", np.array([-8.0, -4.0], dtype=np.float32))
        f.write_dataset("/hamiltonian/eigenvectors", np.ascontiguousarray(evecs.T))
        f.write_dataset("/hamiltonian/residuals", np.array([1e-7, 2e-7], dtype=np.float32))
        f.write_dataset("/observables/Sz", np.array([0.5 + 0j, -0.5 + 1e-9j]))
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)  # end-of-file address
    with hdf5.File(path, "r") as f:
        assert np.array_equal(f.read_dataset("/basis/representatives"), reps)
        assert f.read_dataset("/basis/representatives").dtype == np.uint64
        ev = f.read_dataset("/hamiltonian/eigenvectors")
        assert ev.shape == (2, 16) and ev.dtype == np.float32 and np.array_equal(ev, evecs.T)
        assert f.read_dataset("/hamiltonian/eigenvalues").dtype == np.float32
        sz = f.read_dataset("/observables/Sz")
        assert sz.dtype == np.complex128 and np.allclose(sz, [0.5, -0.5 + 1e-9j])
        assert f.exists("/_workspace") and len(f.root["_workspace"]) == 0


def test_hdf5_many_datasets_and_overwrite(tmp_path):
    path = str(tmp_path / "many.h5")
    with hdf5.File(path, "a") as f:
        for i in range(40):  # more than one symbol-table node
            f.write_dataset(f"/observables/op{i:02d}", np.full(3, i, dtype=np.complex128))
    with hdf5.File(path, "a") as f:
        f.delete("/observables/op07")
        f.write_dataset("/observables/op07", np.array([7.5 + 0j]))
    with hdf5.File(path, "r") as f:
        assert sorted(f.root["observables"]) == [f"op{i:02d}" for i in range(40)]
        assert f.read_dataset("/observables/op07")[0] == 7.5
        assert np.all(f.read_dataset("/observables/op39") == 39)


def test_missing_file_and_bad_signature(tmp_path):
    with pytest.raises(FileNotFoundError):
        hdf5.File(str(tmp_path / "nope.h5"), "r")
    p = tmp_path / "bad.h5"
    p.write_bytes(b"not hdf5" * 20)
    with pytest.raises(ValueError):
        hdf5.File(str(p), "r")
