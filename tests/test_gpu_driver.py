"""End-to-end driver runs (-m gpu): YAML deck in, HDF5 layout out, as the reference's executable
does (/root/reference/app/Main.hs:14-27, README.md:37-95)."""
import numpy as np
import pytest

from spin_ed_b200 import config, decks, driver, hdf5

pytestmark = pytest.mark.gpu


def test_readme_example_end_to_end(tmp_path):
    cfg = decks.load("heisenberg_chain_4")
    cfg["output"] = str(tmp_path / "exact_diagonalization_result.h5")
    driver.run(config.parseConfig(cfg))
    with hdf5.File(cfg["output"], "r") as f:
        reps = f.read_dataset("/basis/representatives")
        assert reps.dtype == np.uint64 and np.array_equal(reps, np.arange(16, dtype=np.uint64))
        ev = f.read_dataset("/hamiltonian/eigenvalues")
        assert ev.dtype == np.float64 and ev.shape == (1,) and abs(ev[0] + 8.0) < 1e-10
        assert f.read_dataset("/hamiltonian/eigenvectors").shape == (1, 16)
        assert f.read_dataset("/hamiltonian/residuals").shape == (1,)
        assert f.exists("/observables") and f.exists("/_workspace")


def test_resume_from_representatives_and_float32_deck(tmp_path):
    cfg = decks.load("heisenberg_square_4x4")  # datatype float32, number_vectors 2
    cfg["output"] = str(tmp_path / "sq.h5")
    ev1, _, _ = driver.run(config.parseConfig(cfg))
    with hdf5.File(cfg["output"], "r") as f:
        assert f.read_dataset("/basis/representatives").shape == (107,)
        assert f.read_dataset("/hamiltonian/eigenvalues").dtype == np.float32
        assert f.read_dataset("/hamiltonian/eigenvectors").shape == (2, 107)
        assert f.read_dataset("/hamiltonian/eigenvectors").dtype == np.float32
    ev2, _, _ = driver.run(config.parseConfig(cfg))  # second run loads /basis/representatives (SpinED.hs:319-331)
    assert abs(ev1[0] + 44.9139328337) < 2e-3 and abs(ev2[0] - ev1[0]) < 1e-4


def test_observables_are_written(tmp_path):
    # an S^z_total observable on a fixed-magnetisation sector: expectation = hw - n/2 exactly
    cfg = decks.load("heisenberg_chain_10")
    cfg["observables"] = [{"name": "Sz", "terms": [{"matrix": [[-0.5, 0], [0, 0.5]], "sites": [[i] for i in range(10)]}]}]
    cfg["output"] = str(tmp_path / "obs.h5")
    driver.run(config.parseConfig(cfg))
    with hdf5.File(cfg["output"], "r") as f:
        sz = f.read_dataset("/observables/Sz")
        assert sz.dtype == np.complex128 and sz.shape == (1,) and abs(sz[0]) < 1e-12
