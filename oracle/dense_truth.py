"""First-principles dense construction (numpy only) used to pin the C++ oracle.  TEST ONLY.

Independent of oracle.cpp: builds the full 2^n x 2^n Hamiltonian from the YAML terms, the
permutation operators U_g of the symmetry group, the projector P = 1/|G'| sum_g conj(chi(g)) U_g,
and the symmetry-adapted basis vectors |r~> = P|r> / n_r.  Semantics follow
/root/reference/template.yaml (:5-8 hamming weight, :30-34 sector <-> exp(2 pi i k / N),
:37-43 spin inversion, :52-70 matrices) and the conventions listed in oracle.cpp's header.
Feasible for n <= ~14.
"""
from __future__ import annotations

import itertools
from fractions import Fraction

import numpy as np


def _permute_state(p, x: int) -> int:
    y = 0
    for i, src in enumerate(p):
        y |= ((x >> src) & 1) << i
    return y


def _period(p) -> int:
    cur = list(range(len(p)))
    for k in itertools.count(1):
        cur = [cur[p[i]] for i in range(len(p))]
        if cur == list(range(len(p))):
            return k


def close_group(n: int, symmetries):
    """-> list of (perm tuple, phase Fraction in [0,1)); raises ValueError on inconsistency."""
    gens = []
    for s in symmetries:
        p = tuple(int(v) for v in s["permutation"])
        if sorted(p) != list(range(n)):
            raise ValueError("invalid permutation")
        per = _period(p)
        if not (0 <= s["sector"] < per):
            raise ValueError("invalid sector")
        gens.append((p, Fraction(s["sector"], per) % 1))
    ident = tuple(range(n))
    seen = {ident: Fraction(0)}
    queue = [ident]
    while queue:
        cur = queue.pop(0)
        for p, ph in gens:
            q = tuple(cur[p[i]] for i in range(n))
            phase = (seen[cur] + ph) % 1
            if q not in seen:
                seen[q] = phase
                queue.append(q)
            elif seen[q] != phase:
                raise ValueError("incompatible symmetries")
    return list(seen.items())


def dense_hamiltonian(n: int, terms) -> np.ndarray:
    dim = 1 << n
    H = np.zeros((dim, dim), dtype=np.complex128)
    for t in terms:
        M = np.asarray(t["matrix"], dtype=np.complex128)
        for sites in t["sites"]:
            k = len(sites)
            for x in range(dim):
                b = 0
                for j, s in enumerate(sites):
                    b |= ((x >> s) & 1) << (k - 1 - j)
                for a in range(1 << k):
                    if M[a, b] == 0:
                        continue
                    xp = x
                    for j, s in enumerate(sites):
                        bit = (a >> (k - 1 - j)) & 1
                        xp = (xp & ~(1 << s)) | (bit << s)
                    H[xp, x] += M[a, b]
    return H


def projector(n: int, symmetries, spin_inversion=None) -> np.ndarray:
    dim = 1 << n
    G = close_group(n, symmetries)
    full = dim - 1
    P = np.zeros((dim, dim), dtype=np.complex128)
    flips = [(False, 1.0)] if not spin_inversion else [(False, 1.0), (True, float(spin_inversion))]
    for p, ph in G:
        chi = np.exp(2j * np.pi * float(ph))
        for flip, sgn in flips:
            c = np.conj(chi * sgn)
            for x in range(dim):
                y = _permute_state(p, x)
                if flip:
                    y ^= full
                P[y, x] += c
    return P / (len(G) * len(flips))


def sector_states(n: int, hamming_weight=None):
    return [x for x in range(1 << n) if hamming_weight is None or bin(x).count("1") == hamming_weight]


def symmetric_basis(n, hamming_weight, spin_inversion, symmetries, tol=1e-9):
    """-> (representatives (sorted ints), norms, B) with B[:, j] = P|r_j> / n_j in the full space."""
    P = projector(n, symmetries, spin_inversion)
    G = close_group(n, symmetries)
    full = (1 << n) - 1
    reps, norms, cols = [], [], []
    for x in sector_states(n, hamming_weight):
        orbit = []
        for p, _ in G:
            y = _permute_state(p, x)
            orbit.append(y)
            if spin_inversion:
                orbit.append(y ^ full)
        if min(orbit) != x:
            continue
        n2 = P[x, x].real
        if n2 < tol:
            continue
        reps.append(x)
        norms.append(np.sqrt(n2))
        cols.append(P[:, x] / np.sqrt(n2))
    B = np.array(cols).T if cols else np.zeros((1 << n, 0), dtype=np.complex128)
    return reps, np.array(norms), B


def symmetric_hamiltonian(n, hamming_weight, spin_inversion, symmetries, terms):
    """-> (representatives, norms, H~ = B^H H B)."""
    reps, norms, B = symmetric_basis(n, hamming_weight, spin_inversion, symmetries)
    H = dense_hamiltonian(n, terms)
    Ht = B.conj().T @ H @ B
    return reps, norms, Ht


def sector_spectrum_by_projection(n, hamming_weight, spin_inversion, symmetries, terms):
    """Spectrum of H restricted to range(P) (intersected with the hamming-weight sector), computed
    without any notion of representatives: eigen-decompose P, keep the eigenvalue-1 space."""
    P = projector(n, symmetries, spin_inversion)
    H = dense_hamiltonian(n, terms)
    idx = sector_states(n, hamming_weight)
    Ps = P[np.ix_(idx, idx)]
    Hs = H[np.ix_(idx, idx)]
    w, v = np.linalg.eigh((Ps + Ps.conj().T) / 2)
    Q = v[:, w > 0.5]
    return np.linalg.eigvalsh(Q.conj().T @ Hs @ Q), Q.shape[1], np.linalg.norm(Hs @ Ps - Ps @ Hs)
