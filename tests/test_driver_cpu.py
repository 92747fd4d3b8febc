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


def _fill(path, n=1000, k=2, complex_vectors=False):
    rng = np.random.default_rng(5)
    with hdf5.File(path, "a") as f:
        for g in ("/basis", "/hamiltonian", "/observables", "/_workspace"):
            f.create_group(g)
    with hdf5.File(path, "a") as f:
        f.write_dataset("/basis/representatives", np.sort(rng.integers(0, 2**40, n)).astype(np.uint64))
    vec = rng.standard_normal((k, n)) + (1j * rng.standard_normal((k, n)) if complex_vectors else 0)
    with hdf5.File(path, "a") as f:
        f.write_dataset("/hamiltonian/eigenvalues", np.array([-8.0, -4.0][:k]))
        f.write_dataset("/hamiltonian/eigenvectors", vec.astype(np.complex64 if complex_vectors else np.float64))
        f.write_dataset("/hamiltonian/residuals", np.array([1e-9, 2e-9][:k]))
    for name in ("Sz", "Sx"):
        with hdf5.File(path, "a") as f:
            f.write_dataset(f"/observables/{name}", rng.standard_normal(k) + 0j)


@pytest.mark.parametrize("complex_vectors", [False, True])
def test_second_reader_accepts_what_the_writer_produces(tmp_path, complex_vectors):
    """An independently written spec-level reader/validator (tests/h5_spec_reader.py: superblock v0,
    symbol-table B-tree, local heap, version-1 object headers, compound complex) walks the file the
    driver's open/append/close sequence produces and returns the same arrays as the product reader."""
    import h5_spec_reader as spec

    path = str(tmp_path / "out.h5")
    _fill(path, complex_vectors=complex_vectors)
    tree = spec.read(path)
    assert sorted(tree) == ["_workspace", "basis", "hamiltonian", "observables"] and tree["_workspace"] == {}
    with hdf5.File(path, "r") as f:
        for g, names in (("basis", ["representatives"]), ("hamiltonian", ["eigenvalues", "eigenvectors", "residuals"]),
                         ("observables", ["Sx", "Sz"])):
            assert sorted(tree[g]) == names
            for nme in names:
                mine = f.read_dataset(f"/{g}/{nme}")
                assert tree[g][nme].dtype == mine.dtype and tree[g][nme].shape == mine.shape
                assert np.array_equal(tree[g][nme], mine)
    assert tree["hamiltonian"]["eigenvectors"].dtype == (np.complex64 if complex_vectors else np.float64)
    # many objects in one group (several symbol-table nodes) and an overwritten dataset
    many = str(tmp_path / "many.h5")
    with hdf5.File(many, "a") as f:
        for i in range(40):
            f.write_dataset(f"/observables/op{i:02d}", np.full(3, i, dtype=np.complex128))
    with hdf5.File(many, "a") as f:
        f.delete("/observables/op07")
        f.write_dataset("/observables/op07", np.array([7.5 + 0j]))
    t = spec.read(many)
    assert sorted(t["observables"]) == [f"op{i:02d}" for i in range(40)] and t["observables"]["op07"][0] == 7.5


def test_second_reader_rejects_damaged_files(tmp_path):
    """The validator is strict: flipping structural bytes must be noticed (so that its acceptance of
    the writer's files means something)."""
    import h5_spec_reader as spec

    path = str(tmp_path / "out.h5")
    _fill(path, n=64)
    raw = bytearray(open(path, "rb").read())
    spec.read(path)
    root_header = struct.unpack_from("<Q", raw, 64)[0]
    damage = [(8, 1, "superblock version"), (40, 0, "end-of-file address"), (root_header, 2, "object header version"),
              (raw.rindex(b"SNOD") + 6, 200, "symbol table node count"), (raw.rindex(b"TREE") + 4, 1, "B-tree node type"),
              (raw.rindex(b"HEAP") + 4, 3, "heap version")]
    for off, val, what in damage:
        bad = bytearray(raw)
        bad[off] = val
        p = tmp_path / "bad.h5"
        p.write_bytes(bytes(bad))
        with pytest.raises(spec.FormatError):
            spec.read(str(p))


def test_append_on_close_moves_no_existing_data_and_reads_lazily(tmp_path):
    """Closing a modified file appends (new raw data + fresh metadata) and patches the superblock: raw
    data already in the file keeps its address, is never rewritten, and comes back as a memory map;
    a file that was only read is not touched at all."""
    path = str(tmp_path / "big.h5")
    reps = np.arange(200000, dtype=np.uint64)
    with hdf5.File(path, "a") as f:
        f.write_dataset("/basis/representatives", reps)
    size1 = os.path.getsize(path)
    raw1 = open(path, "rb").read()
    at = raw1.index(reps[:64].tobytes())
    mtime = os.stat(path).st_mtime_ns
    with hdf5.File(path, "a") as f:  # read only: no rewrite
        r = f.read_dataset("/basis/representatives")
        assert isinstance(r, np.memmap) and np.array_equal(r[-5:], reps[-5:])
    assert os.stat(path).st_mtime_ns == mtime and os.path.getsize(path) == size1
    with hdf5.File(path, "a") as f:
        f.write_dataset("/hamiltonian/eigenvalues", np.array([-1.0]))
    raw2 = open(path, "rb").read()
    assert raw2[96:size1] == raw1[96:size1]              # everything behind the superblock is untouched
    assert raw2.index(reps[:64].tobytes()) == at         # the big dataset did not move
    assert len(raw2) - size1 < 4096                      # only metadata and 8 bytes of data were added
    with hdf5.File(path, "r") as f:
        assert np.array_equal(f.read_dataset("/basis/representatives"), reps)
        assert f.read_dataset("/hamiltonian/eigenvalues")[0] == -1.0


def test_reader_follows_object_header_continuation_blocks(tmp_path):
    """libhdf5 moves messages into continuation blocks (message 0x0010) when a header is touched; the
    resume path must still find dataspace / datatype / layout there."""
    import h5_spec_reader as spec

    path = str(tmp_path / "c.h5")
    data = np.arange(10, dtype=np.uint64)
    with hdf5.File(path, "a") as f:
        f.write_dataset("/basis/representatives", data)
    raw = bytearray(open(path, "rb").read())
    # find the dataset's object header: version 1, 3 messages, first message = dataspace (type 1)
    hdr = next(o for o in range(96, len(raw) - 24, 8)
               if raw[o] == 1 and raw[o + 1] == 0 and struct.unpack_from("<H", raw, o + 2)[0] == 3
               and struct.unpack_from("<HH", raw, o + 16)[0] == 1 and struct.unpack_from("<I", raw, o + 4)[0] == 1)
    size = struct.unpack_from("<I", raw, hdr + 8)[0]
    body = bytes(raw[hdr + 16:hdr + 16 + size])
    first_len = 8 + struct.unpack_from("<H", body, 2)[0]          # the dataspace message stays in the first block
    rest = body[first_len:]
    cont_at = (len(raw) + 7) // 8 * 8
    cont_msg = struct.pack("<HHB3x", 0x0010, 16, 0) + struct.pack("<QQ", cont_at, len(rest))
    filler = size - first_len - len(cont_msg) - 8
    assert filler >= 0
    new_body = body[:first_len] + cont_msg + struct.pack("<HHB3x", 0, filler, 0) + b"\0" * filler   # NIL message pads the block
    raw[hdr + 16:hdr + 16 + size] = new_body
    struct.pack_into("<H", raw, hdr + 2, 5)                        # dataspace, continuation, NIL, datatype, layout
    raw += b"\0" * (cont_at - len(raw)) + rest
    struct.pack_into("<Q", raw, 40, len(raw))                      # end-of-file address
    open(path, "wb").write(bytes(raw))
    assert np.array_equal(spec.read(path)["basis"]["representatives"], data)
    with hdf5.File(path, "r") as f:
        assert np.array_equal(f.read_dataset("/basis/representatives"), data)
