"""Which order of a row's elements makes the lanes of a warp gather from the fewest 32-byte sectors
and 128-byte lines?  Samples slices (32 consecutive rows) of a deck and compares, per warp gather,
(a) the kernels' bond order (high sites first), (b) the row's elements sorted by target index.
Not a test; run by hand:   python tests/analysis/slice_order.py heisenberg_square_6x6 150
"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from spin_ed_b200 import decks
from helpers import oracle_problem
O.build()
name = sys.argv[1] if len(sys.argv) > 1 else "heisenberg_square_6x6"
cfg = decks.load(name)
ob, terms = oracle_problem(O, cfg)
t0 = time.time()
cache = f"/tmp/sped_oracle_reps_{name}.npy"
if os.path.exists(cache):
    ob.build(np.load(cache))
else:
    ob.build(); np.save(cache, ob.states)
print("build", time.time() - t0, "s N", ob.number_states, flush=True)
reps = ob.states
N = len(reps)
sites = sorted(terms[0]["sites"], key=lambda s: -max(s))  # high sites first, like pack_terms
rng = np.random.default_rng(1)
nsl = int(sys.argv[2]) if len(sys.argv) > 2 else 100
slices = np.sort(rng.choice(N // 32 - 1, size=nsl, replace=False))
def count(ell, per):
    # ell: [32, width] targets or -1; distinct (target // per) per column
    tot = 0
    for j in range(ell.shape[1]):
        col = ell[:, j]
        col = col[col >= 0]
        tot += len(np.unique(col // per))
    return tot
E = 0
res = {}
for s in slices:
    rows = []
    for r in range(32 * s, 32 * s + 32):
        x = int(reps[r])
        t = []
        for (i, j) in sites:
            if ((x >> i) ^ (x >> j)) & 1:
                y = x ^ (1 << i) ^ (1 << j)
                rep, chi, norm = ob.state_info(y)
                if norm > 0:
                    t.append(ob.index(rep))
        rows.append(t)
    w = max(len(t) for t in rows)
    E += sum(len(t) for t in rows)
    for oname, f in (("bond order", lambda t: t), ("sorted by target", sorted),
                     ("sorted by |target-row|", None)):
        ell = np.full((32, w), -1, np.int64)
        for l, t in enumerate(rows):
            if f is None:
                tt = sorted(t, key=lambda v: abs(v - (32 * s + l)))
            else:
                tt = f(t)
            ell[l, :len(tt)] = tt
        for per, pname in ((4, "sectors(f64)"), (16, "lines(f64)"), (8, "sectors(f32)"), (32, "lines(f32)")):
            res[(oname, pname)] = res.get((oname, pname), 0) + count(ell, per)
        res[(oname, "warp gathers")] = res.get((oname, "warp gathers"), 0) + w
print("elements", E, "slices", nsl)
for k, v in res.items():
    print(f"{k[0]:24s} {k[1]:14s} {v:9d}  per element {v / E:.3f}" + (f"  per warp gather {v / res[(k[0], 'warp gathers')]:.2f}" if k[1] != "warp gathers" else ""))
