"""Chains: distinct 32-byte sectors per element when the lanes of a warp gather (a) their j-th element in bond order
(the kernels' layout) and (b) the element of the SAME bond (one slot per bond, empty where the bond has no transition)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from spin_ed_b200 import decks
from helpers import oracle_problem
O.build()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = decks.chain(n, n // 2, 1, (0, 0))
ob, terms = oracle_problem(O, cfg); ob.build()
reps = ob.states; N = len(reps)
sites = sorted(terms[0]["sites"], key=lambda s: -max(s))
rng = np.random.default_rng(3)
slices = np.sort(rng.choice(N // 32 - 1, size=300, replace=False))
E = 0; secA = secB = linA = linB = 0; slotsA = slotsB = 0; secB_sectors_stream = 0
for s in slices:
    rows = []
    for r in range(32 * s, 32 * s + 32):
        x = int(reps[r]); t = {}
        for b, (i, j) in enumerate(sites):
            if ((x >> i) ^ (x >> j)) & 1:
                rep, chi, norm = ob.state_info(x ^ (1 << i) ^ (1 << j))
                if norm > 0: t[b] = ob.index(rep)
        rows.append(t)
    E += sum(len(t) for t in rows)
    w = max(len(t) for t in rows); slotsA += 32 * w
    for j in range(w):
        col = np.array([list(t.values())[j] for t in rows if len(t) > j])
        secA += len(np.unique(col // 4)); linA += len(np.unique(col // 16))
    used = sorted(set().union(*[set(t) for t in rows])); slotsB += 32 * len(used)
    for b in used:
        lanes = [l for l, t in enumerate(rows) if b in t]
        col = np.array([rows[l][b] for l in lanes])
        secB += len(np.unique(col // 4)); linB += len(np.unique(col // 16))
        secB_sectors_stream += len(set(l // 8 for l in lanes))  # position-stream sectors actually fetched
print(f"chain_{n}: elements {E}")
print(f"(a) j-th element : slots/elem {slotsA/E:.2f}  gather sectors/elem {secA/E:.3f}  lines/elem {linA/E:.3f}")
print(f"(b) bond-aligned : slots/elem {slotsB/E:.2f}  gather sectors/elem {secB/E:.3f}  lines/elem {linB/E:.3f}  stream sectors fetched/elem {secB_sectors_stream*8/E/8:.3f} (x32 B)")
