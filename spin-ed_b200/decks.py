"""Input decks of the reference's benchmark configurations, in the reference's own YAML schema
(/root/reference/template.yaml; parsed by src/SpinED.hs:81-173).

Chains and L x L square lattices are generated from the lattice geometry; the irregular clusters
are carried as JSON fixtures under ``decks/`` (made by tools/make_decks.py).  Every generated deck
is checked field-by-field against the reference's example/*.yaml by tests/test_decks.py.
"""
from __future__ import annotations

import copy
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

HEISENBERG = [[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]]


def _chain(n, hamming_weight=None, spin_inversion=None, sectors=None, **extra):
    basis = {"number_spins": n}
    if hamming_weight is not None:
        basis["hamming_weight"] = hamming_weight
    if spin_inversion is not None:
        basis["spin_inversion"] = spin_inversion
    syms = []
    if sectors is not None:
        syms.append({"permutation": [(i + 1) % n for i in range(n)], "sector": sectors[0]})
        syms.append({"permutation": [n - 1 - i for i in range(n)], "sector": sectors[1]})
    basis["symmetries"] = syms
    cfg = {
        "basis": basis,
        "hamiltonian": {
            "name": "Heisenberg Hamiltonian",
            "terms": [{"matrix": copy.deepcopy(HEISENBERG), "sites": [[i, (i + 1) % n] for i in range(n)]}],
        },
        "observables": [],
    }
    cfg.update(extra)
    return cfg


def _square(L, **extra):
    n = L * L
    idx = lambda x, y: y * L + x  # noqa: E731
    tx = [idx((x + 1) % L, y) for y in range(L) for x in range(L)]
    ty = [idx(x, (y + 1) % L) for y in range(L) for x in range(L)]
    rx = [idx(L - 1 - x, y) for y in range(L) for x in range(L)]
    ry = [idx(x, L - 1 - y) for y in range(L) for x in range(L)]
    rot = [idx(L - 1 - y, x) for y in range(L) for x in range(L)]
    sites = []
    for y in range(L):
        for x in range(L):
            sites.append([idx(x, y), idx((x + 1) % L, y)])
            sites.append([idx(x, y), idx(x, (y + 1) % L)])
    cfg = {
        "basis": {
            "number_spins": n,
            "hamming_weight": n // 2,
            "spin_inversion": 1,
            "symmetries": [{"permutation": p, "sector": 0} for p in (tx, ty, rx, ry, rot)],
        },
        "hamiltonian": {"name": "Heisenberg Hamiltonian", "terms": [{"matrix": copy.deepcopy(HEISENBERG), "sites": sites}]},
        "observables": [],
    }
    cfg.update(extra)
    return cfg


_GENERATED = {
    "heisenberg_chain_4": lambda: _chain(4),
    "heisenberg_chain_10": lambda: _chain(10, 5, -1, (5, 1), output="data/heisenberg_chain_10.h5"),
    "heisenberg_chain_24": lambda: _chain(24, 12, output="data/heisenberg_chain_24.h5"),
    "heisenberg_chain_40": lambda: _chain(
        40, 20, 1, (0, 0), number_vectors=1, max_primme_basis_size=3, output="data/heisenberg_chain_40.h5"
    ),
    "heisenberg_chain_42": lambda: _chain(
        42, 21, -1, (21, 1), number_vectors=1, max_primme_basis_size=4, output="data/heisenberg_chain_42.h5"
    ),
    "heisenberg_square_4x4": lambda: _square(
        4, number_vectors=2, output="data/heisenberg_square_4x4.h5", datatype="float32",
        max_primme_block_size=4, max_primme_basis_size=20,
    ),
    "heisenberg_square_6x6": lambda: _square(
        6, number_vectors=2, output="data/heisenberg_square_6x6.h5", datatype="float32",
        max_primme_block_size=4, max_primme_basis_size=20,
    ),
}


def chain(n, hamming_weight=None, spin_inversion=None, sectors=None, **extra):
    """A periodic Heisenberg chain deck with translation + parity sectors (not a reference file)."""
    return _chain(n, hamming_weight, spin_inversion, sectors, **extra)


def names():
    fixtures = [f[:-5] for f in sorted(os.listdir(os.path.join(_HERE, "decks"))) if f.endswith(".json")]
    return sorted(set(_GENERATED) | set(fixtures))


def load(name: str) -> dict:
    """Deck by reference file stem, e.g. ``heisenberg_square_6x6`` (also accepts ``*.yaml`` paths)."""
    stem = os.path.basename(name)
    for ext in (".yaml", ".yml", ".json"):
        if stem.endswith(ext):
            stem = stem[: -len(ext)]
    if stem in _GENERATED:
        return _GENERATED[stem]()
    if stem.startswith("heisenberg_chain_") and stem[17:].isdigit() and int(stem[17:]) % 4 == 0:
        # size-scaling proxies of heisenberg_chain_40 (same sector and solver options), not reference files
        n = int(stem[17:])
        return _chain(n, n // 2, 1, (0, 0), number_vectors=1, max_primme_basis_size=3, output=f"data/{stem}.h5")
    path = os.path.join(_HERE, "decks", stem + ".json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    if os.path.exists(name):
        import yaml

        with open(name) as f:
            return yaml.safe_load(f)
    raise KeyError(f"unknown deck {name!r}; known: {names()}")
