"""Pins the CPU oracle (-m "not gpu").  The upstream numerics cannot be built here,
so the oracle is checked against (a) the reference's README known answer and Spec.hs facts,
(b) golden vectors produced by an independent first-principles numpy construction
(tests/golden/make_golden.py), (c) that construction directly, element by element, and
(d) the independently computed values listed in SURVEY.md 8(c)."""
import json
import os

import numpy as np
import pytest

from helpers import cmat, extra_configs, oracle_problem
from spin_ed_b200 import decks

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "small_sectors.json")))


def _cfg(name):
    return extra_configs()[name] if name in extra_configs() else decks.load(name)


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_matches_golden_vectors(oracle, name):
    g = GOLDEN[name]
    ob, terms = oracle_problem(oracle, _cfg(name))
    ob.build()
    assert [int(v) for v in ob.states] == g["representatives"]
    assert np.allclose(ob.norms, g["norms"], rtol=0, atol=1e-15)
    op = oracle.Operator(ob, terms)
    assert op.is_real == (not g["complex"])
    x = np.array([complex(*v) for v in g["x"]])
    y = np.array([complex(*v) for v in g["y"]])
    got = op.matmat(x if g["complex"] else x.real.copy())
    assert np.linalg.norm(got - y) <= 1e-13 * np.linalg.norm(y)
    ev = np.linalg.eigvalsh(op.to_dense())[: len(g["eigenvalues"])]
    assert np.allclose(ev, g["eigenvalues"], rtol=0, atol=1e-11)
    # the bit-by-bit permutation path gives the same basis as the Benes path
    ob2, _ = oracle_problem(oracle, _cfg(name))
    ob2.use_naive(True)
    ob2.build()
    assert np.array_equal(ob2.states, ob.states) and np.array_equal(ob2.norms, ob.norms)


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "chain_8_k1_complex", "chain_8_chiral_3site"])
def test_oracle_matrix_elements_equal_dense_projection(oracle, name):
    from oracle import dense_truth as D

    cfg = _cfg(name)
    b = cfg["basis"]
    terms = [{"matrix": cmat(t["matrix"]), "sites": t["sites"]} for t in cfg["hamiltonian"]["terms"]]
    reps, norms, Ht = D.symmetric_hamiltonian(b["number_spins"], b.get("hamming_weight"), b.get("spin_inversion"),
                                              b["symmetries"], terms)
    ob, oterms = oracle_problem(oracle, cfg)
    ob.build()
    Ho = oracle.Operator(ob, oterms).to_dense()
    assert np.abs(Ho - Ht).max() < 1e-13
    assert np.abs(Ho - Ho.conj().T).max() < 1e-13


def test_readme_known_answer(oracle):
    # /root/reference/README.md:37-95
    ob, terms = oracle_problem(oracle, decks.load("heisenberg_chain_4"))
    ob.build()
    assert ob.number_states == 16
    assert abs(np.linalg.eigvalsh(oracle.Operator(ob, terms).to_dense())[0] + 8.0) < 1e-12


def test_spec_hs_structural_facts(oracle):
    # /root/reference/test/Spec.hs:38-43,72-79
    assert oracle.periodicity([3, 2, 1, 0]) == 2
    assert oracle.periodicity([4, 3, 4, 1]) == -1
    with pytest.raises(oracle.OracleError):
        oracle.Basis(5, None, None, [{"permutation": [4, 3, 2, 1, 0], "sector": 3}])
    d4 = [{"permutation": [3, 2, 1, 0], "sector": 0}, {"permutation": [1, 2, 3, 0], "sector": 0}]
    assert oracle.Basis(4, 2, None, d4).group_size == 8
    with pytest.raises(oracle.OracleError) as e:
        oracle.Basis(4, 2, None, [d4[0], {"permutation": [1, 2, 3, 0], "sector": 1}])
    assert e.value.code == 11


@pytest.mark.parametrize("name,dim", [("heisenberg_square_4x4", 107), ("heisenberg_triangular_19", 4862),
                                      ("heisenberg_square_5x5", 208012), ("heisenberg_chain_24", 2704156),
                                      ("xxz_triangular_19", 524288)])
def test_sector_dimensions_of_survey(oracle, name, dim):
    ob, _ = oracle_problem(oracle, decks.load(name))
    ob.build()
    assert ob.number_states == dim
    s = ob.states
    assert np.all(s[1:] > s[:-1])


def test_survey_energies(oracle):
    import scipy.sparse.linalg as sla

    for name, e0 in [("heisenberg_square_4x4", -44.9139328337), ("heisenberg_chain_24", -42.6800580661)]:
        ob, terms = oracle_problem(oracle, decks.load(name))
        ob.build()
        op = oracle.Operator(ob, terms)
        n = ob.number_states
        A = sla.LinearOperator((n, n), matvec=lambda v: op.matmat(np.ascontiguousarray(v, dtype=np.float64)), dtype=np.float64)
        ev = sla.eigsh(A, k=1, which="SA", tol=1e-11)[0]
        assert abs(ev[0] - e0) < 1e-8


def test_symmetric_sector_spectrum_is_subset_of_unsymmetrised(oracle):
    cfg = decks.chain(12, 6, 1, (0, 0))
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    sym = np.linalg.eigvalsh(oracle.Operator(ob, terms).to_dense())
    plain_cfg = decks.chain(12, 6)
    ob2, terms2 = oracle_problem(oracle, plain_cfg)
    ob2.build()
    full = np.linalg.eigvalsh(oracle.Operator(ob2, terms2).to_dense())
    for v in sym:
        assert np.min(np.abs(full - v)) < 1e-10


def test_expectation_and_block_layout(oracle):
    ob, terms = oracle_problem(oracle, decks.load("heisenberg_chain_10"))
    ob.build()
    op = oracle.Operator(ob, terms)
    rng = np.random.default_rng(0)
    X = np.asfortranarray(rng.standard_normal((13, 3)))
    Y = op.matmat(X)
    for c in range(3):
        assert np.allclose(Y[:, c], op.matmat(X[:, c].copy()))
    assert np.allclose(op.expectation(X), [X[:, c] @ Y[:, c] for c in range(3)])
    y32 = op.matmat(X.astype(np.float32))
    assert y32.dtype == np.float32 and np.allclose(y32, Y, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_square_4x4", "heisenberg_triangular_19", "heisenberg_kagome_12"])
def test_lazy_adoption_row_lists_and_window_enumeration(oracle, name):
    """The bounded-sample entry points bench.py uses at the 36-42 spin sizes agree with the plain
    oracle: lazily adopted representatives (norms derived per use), an explicit row list, and the
    independent enumeration of a window of candidate ranks."""
    from spin_ed_b200 import decks
    from helpers import oracle_problem, splitmix_vector

    cfg = decks.load(name)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    op = oracle.Operator(ob, terms)
    n = ob.number_states
    x = splitmix_vector(n, 7, np.float64 if op.is_real else np.complex128)
    want, total = op.matmat(x, count=True)
    lazy, _ = oracle_problem(oracle, cfg)
    lazy.adopt_lazy(ob.states)
    lop = oracle.Operator(lazy, terms)
    rows = np.arange(n - 1, -1, -3, dtype=np.uint64)  # any order, not only ascending
    got, visited = lop.matmat_list(x, rows)
    assert np.array_equal(got, want[rows.astype(np.int64)])
    assert visited <= total
    cand = ob.sector_candidates
    pieces = []
    step = max(1, cand // 5)
    for lo in range(0, cand, step):
        reps, w0, w1 = lazy.build_range(lo, min(cand, lo + step))
        s = ob.states
        inside = s[(s >= np.uint64(w0)) & ((s < np.uint64(w1)) if w1 != 2**64 - 1 else np.ones(len(s), bool))]
        assert np.array_equal(inside, reps)
        pieces.append(reps)
    assert np.array_equal(np.concatenate(pieces), ob.states)


def test_bethe_ansatz_exact_energies(oracle):
    """oracle/bethe.py -- the exact Bethe-ansatz ground-state energy of the Heisenberg ring, an answer
    that shares nothing with this repository's algorithms -- against the reference's known answers
    (README 4-ring: -8, /root/reference/README.md:56-95; SURVEY 8c: chain_10 -18.061785418, chain_24
    -42.6800580661), against dense diagonalisation of the oracle's matrices, and against the
    eigenvalues the GPU path produced at full size (committed driver-command bench records)."""
    import glob
    import json
    import os

    from oracle import bethe

    assert abs(bethe.sigma_sigma_ring_energy(4) + 8.0) < 1e-13
    assert abs(bethe.sigma_sigma_ring_energy(10) + 18.061785418) < 1e-9
    assert abs(bethe.sigma_sigma_ring_energy(24) + 42.6800580661) < 1e-9
    for n in (6, 8, 12):  # unsymmetrised zero-magnetisation sector, dense
        ob, terms = oracle_problem(oracle, decks.chain(n, n // 2))
        ob.build()
        e0 = np.linalg.eigvalsh(oracle.Operator(ob, terms).to_dense())[0]
        assert abs(e0 - bethe.sigma_sigma_ring_energy(n)) < 1e-12 * abs(e0)
    # full size: what the B200 runs of this round measured (north_star: eigenvalues to <= 1e-10 relative)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = 0
    for path in sorted(glob.glob(os.path.join(root, "profiles", "r02b_bench_*.json"))):
        with open(path) as f:
            d = json.load(f)
        found = []
        name = d["config"]["workload"]
        if name.startswith("heisenberg_chain_") and d["extra"].get("eigenvalues"):
            found.append((int(name.rsplit("_", 1)[1]), d["extra"]["eigenvalues"][0]))
        c = d["extra"].get("chain_40")
        if c and c.get("E0") is not None:
            found.append((40, c["E0"]))
        for n, e0 in found:
            exact = bethe.sigma_sigma_ring_energy(n)
            assert abs(e0 - exact) <= 1e-10 * abs(exact), (path, n, e0, exact)
            seen += 1
    assert seen >= 5  # chain_40 at 1/2/4/8 GPUs, chain_42 at 8, chain_24


XX_MATRIX = [[0, 0, 0, 0], [0, 0, 2, 0], [0, 2, 0, 0], [0, 0, 0, 0]]  # sx sx + sy sy


def xx_chain(n, spin_inversion=None, sectors=None):
    cfg = decks.chain(n, n // 2, spin_inversion, sectors)
    cfg["hamiltonian"]["terms"][0]["matrix"] = XX_MATRIX
    return cfg


def test_xx_ring_free_fermion_energies(oracle):
    """A second exact answer with a different two-site matrix (no diagonal part): the XX ring is free
    fermions.  The oracle reproduces it without symmetries and in the symmetric sector that holds the
    ground state (momentum 0 / parity + / inversion + for even n/2, momentum pi / - / - for odd n/2)."""
    from oracle import bethe

    for cfg, n in [(xx_chain(8), 8), (xx_chain(10), 10), (xx_chain(12), 12), (xx_chain(12, 1, (0, 0)), 12),
                   (xx_chain(16, 1, (0, 0)), 16), (xx_chain(10, -1, (5, 1)), 10), (xx_chain(14, -1, (7, 1)), 14)]:
        ob, terms = oracle_problem(oracle, cfg)
        ob.build()
        e0 = np.linalg.eigvalsh(oracle.Operator(ob, terms).to_dense())[0]
        assert abs(e0 - bethe.xx_ring_energy(n)) < 1e-12 * abs(e0), (n, e0)


def test_element_counts_of_unsymmetrised_decks_have_closed_forms(oracle):
    """Without symmetries the number of off-diagonal elements is plain combinatorics: every bond
    contributes once for each state in which its two spins are antiparallel.  chain_24 (hamming weight
    12): 24 * 2 * C(22, 11); xxz_triangular_19 (no hamming-weight restriction, 57 bonds): 57 * 2^18.
    The oracle's count and the count the GPU path reported at full size (committed bench records) must
    both equal them."""
    import json
    import math
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, rows, elements in (("heisenberg_chain_24", math.comb(24, 12), 24 * 2 * math.comb(22, 11)),
                                 ("xxz_triangular_19", 2 ** 19, 57 * 2 ** 18)):
        with open(os.path.join(root, "profiles", f"r02b_bench_{name}_n1.json")) as f:
            d = json.load(f)
        assert d["config"]["rows"] == rows and d["config"]["offdiag_elements"] == elements, name
    ob, terms = oracle_problem(oracle, decks.load("xxz_triangular_19"))
    ob.build()
    assert oracle.Operator(ob, terms).count_offdiag() == 57 * 2 ** 18
