"""Shared helpers for the test-suite: build the same problem in the oracle and in the product."""
import numpy as np

from spin_ed_b200 import config, decks, ffi


def cmat(m):
    return np.array([[complex(*v) if isinstance(v, (list, tuple)) else complex(v) for v in row] for row in m])


def oracle_problem(O, cfg):
    b = cfg["basis"]
    basis = O.Basis(b["number_spins"], b.get("hamming_weight"), b.get("spin_inversion"), b["symmetries"])
    terms = [{"matrix": cmat(t["matrix"]), "sites": t["sites"]} for t in cfg["hamiltonian"]["terms"]]
    return basis, terms


def product_problem(cfg):
    return config.toConfig(config.parseConfig(cfg))


def splitmix_vector(n, seed=0x5EED0001, dtype=np.float64, row0=0):
    """uniform(-1,1) from splitmix64(seed ^ global_row) -- SURVEY 8(d) input convention."""
    z = (np.arange(row0, row0 + n, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    re = (z >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 * 2.0 - 1.0
    if np.dtype(dtype).kind == "c":
        z2 = z + np.uint64(0x9E3779B97F4A7C15)
        z2 = (z2 ^ (z2 >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z2 = (z2 ^ (z2 >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z2 = z2 ^ (z2 >> np.uint64(31))
        im = (z2 >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 * 2.0 - 1.0
        return (re + 1j * im).astype(dtype)
    return re.astype(dtype)


SMALL_DECKS = ["heisenberg_chain_4", "heisenberg_chain_10", "heisenberg_kagome_12", "heisenberg_square_4x4",
               "heisenberg_triangular_19"]


def extra_configs():
    """Small configurations that exercise corners the shipped decks do not."""
    out = {}
    out["chain_12_full_sym"] = decks.chain(12, 6, 1, (0, 0))
    out["chain_12_pi"] = decks.chain(12, 6, -1, (6, 1))
    out["chain_8_k1_complex"] = {
        "basis": {"number_spins": 8, "hamming_weight": 4,
                  "symmetries": [{"permutation": [1, 2, 3, 4, 5, 6, 7, 0], "sector": 1}]},
        "hamiltonian": decks.chain(8)["hamiltonian"], "observables": []}
    out["chain_9_k2_nohw"] = {
        "basis": {"number_spins": 9, "symmetries": [{"permutation": [1, 2, 3, 4, 5, 6, 7, 8, 0], "sector": 2}]},
        "hamiltonian": decks.chain(9)["hamiltonian"], "observables": []}
    out["chain_10_inv_only"] = {
        "basis": {"number_spins": 10, "hamming_weight": 5, "spin_inversion": -1, "symmetries": []},
        "hamiltonian": decks.chain(10)["hamiltonian"], "observables": []}
    # wide words: > 48 spins take the plain 64-bit path (no packed keys), 64 spins fill the word
    out["chain_50_hw2"] = decks.chain(50, 2, None, (0, 0))
    out["chain_64_hw2"] = decks.chain(64, 2, None, (0, 0))
    out["chain_40_hw3_k"] = decks.chain(40, 3, None, (20, 1))
    # spin inversion without a hamming-weight restriction
    out["chain_10_inv_nohw"] = decks.chain(10, None, 1, (0, 0))
    # XX ring: a two-site matrix without a diagonal part (exact answer: free fermions, oracle/bethe.py)
    out["xx_chain_12_sym"] = decks.chain(12, 6, 1, (0, 0))
    out["xx_chain_12_sym"]["hamiltonian"]["terms"][0]["matrix"] = [[0, 0, 0, 0], [0, 0, 2, 0], [0, 2, 0, 0], [0, 0, 0, 0]]
    # 3-site and 1-site terms with a complex matrix: chirality-like term + field
    sx = np.array([[0, 1], [1, 0]], dtype=complex); sy = np.array([[0, -1j], [1j, 0]]); sz = np.diag([1.0 + 0j, -1.0])
    def kron3(a, b, c): return np.kron(a, np.kron(b, c))
    chir = (kron3(sx, sy, sz) + kron3(sy, sz, sx) + kron3(sz, sx, sy) - kron3(sx, sz, sy) - kron3(sz, sy, sx) - kron3(sy, sx, sz))
    n = 8
    out["chain_8_chiral_3site"] = {
        "basis": {"number_spins": n, "hamming_weight": 4,
                  "symmetries": [{"permutation": [(i + 1) % n for i in range(n)], "sector": 0}]},
        "hamiltonian": {"name": "H", "terms": [
            {"matrix": [[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]], "sites": [[i, (i + 1) % n] for i in range(n)]},
            {"matrix": [[[float(v.real), float(v.imag)] for v in row] for row in 0.3 * chir],
             "sites": [[i, (i + 1) % n, (i + 2) % n] for i in range(n)]},
        ]}, "observables": []}
    # 4-site ring exchange-like real term, no symmetries, no hamming weight
    p4 = np.zeros((16, 16))
    for a in range(16):
        bits = [(a >> (3 - j)) & 1 for j in range(4)]
        rot = bits[1:] + bits[:1]
        b = sum(v << (3 - j) for j, v in enumerate(rot))
        p4[b, a] += 1.0; p4[a, b] += 1.0
    out["ring_4site_nosym"] = {
        "basis": {"number_spins": 8, "symmetries": []},
        "hamiltonian": {"name": "H", "terms": [
            {"matrix": p4.tolist(), "sites": [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 6, 7], [6, 7, 0, 1]]},
            {"matrix": [[-0.5, 0], [0, 0.5]], "sites": [[i] for i in range(8)]},
        ]}, "observables": []}
    return out
