"""Exact ground-state energies of spin-1/2 rings (Heisenberg: Bethe ansatz; XX: free fermions) -- TEST INFRASTRUCTURE.

An answer that owes nothing to this repository's algorithms (no basis, no symmetry group, no sparse
product, no eigensolver): for even N the ground state of H = J sum_i S_i.S_{i+1} is the Bethe state of
M = N/2 real rapidities x_j with quantum numbers I_j = -(M-1)/2 ... (M-1)/2,

    N * 2 atan(2 x_j) = 2 pi I_j + sum_k 2 atan(x_j - x_k),        E = J (N/4 - sum_j 2 / (1 + 4 x_j^2)),

M coupled equations solved here by Newton's method in float64 (residual < 1e-11).  The reference's
Heisenberg decks (/root/reference/example/heisenberg_chain_*.yaml, matrix [[1,0,0,0],[0,-1,2,0],[0,2,-1,0],
[0,0,0,1]] = sigma.sigma = 4 S.S, and the symmetry sector that holds the ground state: momentum 0 for
even N/2, pi for odd N/2) therefore have E0 = 4 * E_Bethe(N).  Pins: the README's 4-ring (-8,
/root/reference/README.md:56-95) and SURVEY 8c's chain_10 / chain_24 values, see tests/test_oracle.py; at
full size bench.py compares the 40- and 42-spin eigenvalues with it (north_star: <= 1e-10 relative)."""
import numpy as np


def ground_state_energy(n_sites: int, coupling: float = 1.0) -> float:
    """E0 of H = coupling * sum_i S_i.S_{i+1} on a ring of n_sites (even, >= 4)."""
    if n_sites < 4 or n_sites % 2:
        raise ValueError("Bethe-ansatz ground state: even number of sites >= 4")
    m = n_sites // 2
    quantum = np.arange(m) - (m - 1) / 2.0
    x = np.tan(np.pi * quantum / n_sites) / 2.0  # free-magnon starting point

    def residual(x):
        d = x[:, None] - x[None, :]
        return n_sites * 2.0 * np.arctan(2.0 * x) - 2.0 * np.pi * quantum - (2.0 * np.arctan(d)).sum(axis=1)

    for _ in range(200):
        d = x[:, None] - x[None, :]
        k = 2.0 / (1.0 + d * d)
        jac = k.copy()
        np.fill_diagonal(jac, 0.0)
        jac[np.diag_indices(m)] = n_sites * 4.0 / (1.0 + 4.0 * x * x) - (k.sum(axis=1) - 2.0)
        step = np.linalg.solve(jac, -residual(x))
        x = x + step
        if np.abs(step).max() < 1e-15:
            break
    if np.abs(residual(x)).max() > 1e-11:
        raise RuntimeError("Bethe equations did not converge")
    return float(coupling * (n_sites / 4.0 - np.sum(2.0 / (1.0 + 4.0 * x * x))))


def sigma_sigma_ring_energy(n_sites: int) -> float:
    """E0 of sum_i sigma_i.sigma_{i+1} (the matrix of the reference's Heisenberg decks) = 4 * E0(S.S)."""
    return 4.0 * ground_state_energy(n_sites)


def xx_ring_energy(n_sites: int) -> float:
    """Exact ground-state energy of sum_i (sx_i sx_{i+1} + sy_i sy_{i+1}) (Pauli matrices; two-site
    matrix [[0,0,0,0],[0,0,2,0],[0,2,0,0],[0,0,0,0]]) on a ring of even n_sites at zero magnetisation:
    free fermions after the Jordan-Wigner transformation, eps(k) = 4 cos k, momenta (2j+1) pi / n for
    an even number of fermions (antiperiodic), 2 pi j / n for an odd one; the n/2 lowest levels filled.
    A second exact answer, for a Hamiltonian with a different matrix (no diagonal part)."""
    if n_sites < 4 or n_sites % 2:
        raise ValueError("even number of sites >= 4")
    m = n_sites // 2
    j = np.arange(n_sites)
    k = (2 * j + 1) * np.pi / n_sites if m % 2 == 0 else 2 * j * np.pi / n_sites
    return float(np.sort(4.0 * np.cos(k))[:m].sum())


if __name__ == "__main__":
    for n in (4, 10, 16, 24, 36, 40, 42):
        print(n, repr(sigma_sigma_ring_energy(n)))
