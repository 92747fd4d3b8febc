"""Chains: which share of the matrix elements needs a non-identity symmetry to reach its representative, and
how far from its row does an element gather?  (CPU study behind DESIGN section 4: why the 40-spin kernel is
DRAM-sector bound.)  Run by hand: python tests/analysis/chain_far_targets.py"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from spin_ed_b200 import decks
from helpers import oracle_problem
O.build()
for n in (28, 32):
    cfg = decks.chain(n, n // 2, 1, (0, 0))
    ob, terms = oracle_problem(O, cfg)
    ob.build()
    reps = ob.states; N = len(reps)
    sites = terms[0]["sites"]
    rng = np.random.default_rng(2)
    rows = np.sort(rng.choice(N - 4, size=1500, replace=False))
    d = []; ident = 0
    for r0 in rows:
        for r in range(r0, r0 + 2):
            x = int(reps[r])
            for (i, j) in sites:
                if ((x >> i) ^ (x >> j)) & 1:
                    y = x ^ (1 << i) ^ (1 << j)
                    rep, chi, norm = ob.state_info(y)
                    if norm > 0:
                        d.append(abs(ob.index(rep) - r)); ident += rep == y
    d = np.array(d)
    print(f"chain_{n}: N={N} elements {len(d)} identity-canonical fraction {ident/len(d):.3f}; distance < N/1000: {(d < N/1000).mean():.3f}, < N/100: {(d < N/100).mean():.3f}, < N/10: {(d < N/10).mean():.3f}")
