"""How far from its row does a matrix element gather?  Samples rows of a deck (default 6x6) and
prints the distribution of |target index - row index| (not a test; run by hand, about two minutes
for 6x6: the oracle builds the basis, then one state_info call per sampled element).

    python tests/analysis/target_distance.py heisenberg_square_6x6 1500

Output: profiles/r01_locality_study.txt."""
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
sites = terms[0]["sites"]
rng = np.random.default_rng(1)
rows = np.sort(rng.choice(N - 32, size=int(sys.argv[2]) if len(sys.argv) > 2 else 600, replace=False))
d = []
ident = 0
for r0 in rows:
    for r in range(r0, r0 + 4):
        x = int(reps[r])
        for (i, j) in sites:
            if ((x >> i) ^ (x >> j)) & 1:
                y = x ^ (1 << i) ^ (1 << j)
                rep, chi, norm = ob.state_info(y)
                if norm > 0:
                    t = ob.index(rep)
                    d.append(abs(t - r))
                    ident += rep == y
d = np.array(d)
print("elements", len(d), "identity-canonical fraction", ident / len(d))
for t in (32, 256, 1024, 4096, 16384, 65536, 1 << 20):
    print("distance <", t, (d < t).mean())
