"""Regenerates tests/golden/small_sectors.json with the first-principles numpy construction
(oracle/dense_truth.py: full 2^n Hamiltonian, permutation operators, projector) -- independent of
both the C++ oracle and the CUDA product.  The reference itself cannot produce vectors here (its
numerics are un-vendored dependencies).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import cmat, extra_configs  # noqa: E402
from oracle import dense_truth as D  # noqa: E402
from spin_ed_b200 import decks  # noqa: E402


def main():
    cases = {"heisenberg_chain_4": decks.load("heisenberg_chain_4"), "heisenberg_chain_10": decks.load("heisenberg_chain_10"),
             "heisenberg_kagome_12": decks.load("heisenberg_kagome_12")}
    cases.update({k: v for k, v in extra_configs().items() if v["basis"]["number_spins"] <= 12})
    out = {}
    for name, cfg in cases.items():
        b = cfg["basis"]
        n, hw, inv, syms = b["number_spins"], b.get("hamming_weight"), b.get("spin_inversion"), b["symmetries"]
        terms = [{"matrix": cmat(t["matrix"]), "sites": t["sites"]} for t in cfg["hamiltonian"]["terms"]]
        reps, norms, Ht = D.symmetric_hamiltonian(n, hw, inv, syms, terms)
        ev = np.linalg.eigvalsh(Ht)
        ev2, dim, comm = D.sector_spectrum_by_projection(n, hw, inv, syms, terms)
        assert dim == len(reps) and np.allclose(ev, ev2, atol=1e-10) and comm < 1e-10
        rng = np.random.default_rng(7)
        x = rng.standard_normal(len(reps)) + (1j * rng.standard_normal(len(reps)) if np.abs(Ht.imag).max() > 1e-14 else 0)
        y = Ht @ x
        out[name] = {
            "representatives": [int(r) for r in reps], "norms": [float(v) for v in norms],
            "eigenvalues": [float(v) for v in ev[: min(6, len(ev))]],
            "complex": bool(np.abs(Ht.imag).max() > 1e-14),
            "x": [[float(v.real), float(v.imag)] for v in np.atleast_1d(x).astype(complex)],
            "y": [[float(v.real), float(v.imag)] for v in np.atleast_1d(y).astype(complex)],
        }
        print(name, len(reps), ev[:3])
    with open(os.path.join(HERE, "small_sectors.json"), "w") as f:
        json.dump(out, f)
        f.write("\n")


if __name__ == "__main__":
    main()
