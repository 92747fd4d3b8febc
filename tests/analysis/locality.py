"""Gather-locality study behind two design choices (not a test; run by hand, about a minute):

    python tests/analysis/locality.py 28        # or 32: periodic Heisenberg chain, fully symmetric sector

* bond order (operator.cu, pack_terms): 32-byte sectors touched per gathered element when a warp
  of 32 consecutive rows walks its elements slot by slot, for ascending / descending bond order and
  for hypothetical bond-aligned slots;
* distribution block size (comm.cpp, make_row_dist): fraction of elements whose source entry lives
  on the same rank under the block-cyclic distribution, and the distribution of |target - row|.

Uses the oracle (test infrastructure) for the representatives and a numpy canonicalisation.
Output of the two runs: profiles/r01_locality_study.txt."""
import sys

import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from spin_ed_b200 import decks
from helpers import oracle_problem
O.build()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
cfg = decks.chain(n, n // 2, 1, (0, 0))
ob, terms = oracle_problem(O, cfg)
ob.build()
reps = ob.states.astype(np.uint64)
N = len(reps)
print("n", n, "N", N)
mask = np.uint64((1 << n) - 1)
def rot(x, k):
    k = np.uint64(k)
    return ((x << k) | (x >> (np.uint64(n) - k))) & mask if k else x
def rev(x):
    out = np.zeros_like(x)
    for i in range(n):
        out |= ((x >> np.uint64(i)) & np.uint64(1)) << np.uint64(n - 1 - i)
    return out
def canon(x):
    best = x.copy()
    isid = np.ones(len(x), bool)
    xr = rev(x)
    first = True
    for base in (x, xr):
        for k in range(n):
            y = rot(base, k)
            for z in (y, y ^ mask):
                if first:
                    first = False
                    continue
                lt = z < best
                best = np.where(lt, z, best)
                isid &= ~lt
    return best, isid
# per row, per bond: target index or -1
tgt = np.full((N, n), -1, dtype=np.int64)
ident = np.zeros((N, n), bool)
for i in range(n):
    j = (i + 1) % n
    bi = (reps >> np.uint64(i)) & np.uint64(1); bj = (reps >> np.uint64(j)) & np.uint64(1)
    anti = bi != bj
    fl = reps ^ ((np.uint64(1) << np.uint64(i)) | (np.uint64(1) << np.uint64(j)))
    s, isid = canon(fl[anti])
    idx = np.searchsorted(reps, s)
    ok = (idx < N) & (reps[np.minimum(idx, N - 1)] == s)
    t = np.where(ok, idx, -1)
    tgt[anti, i] = t
    ident[anti, i] = isid
E = (tgt >= 0).sum()
print("E", E, "per row", E / N, "identity-canonical fraction", ident[tgt >= 0].mean())
def sectors(order, per_sector=4):
    # ELL: row's elements in `order` of bonds, compacted; per warp of 32 rows and slot j count distinct sectors
    T = tgt[:, order]
    live = T >= 0
    slot = np.cumsum(live, axis=1) - 1
    W = (N + 31) // 32
    total = 0
    maxlen = live.sum(1).max()
    rows = np.arange(N)
    warp = rows // 32
    # build dense [N, maxlen] array
    ell = np.full((N, maxlen), -1, np.int64)
    r, c = np.nonzero(live)
    ell[r, slot[r, c]] = T[r, c]
    pad = W * 32 - N
    ell = np.vstack([ell, np.full((pad, maxlen), -1, np.int64)]).reshape(W, 32, maxlen)
    sec = np.where(ell >= 0, ell // per_sector, -1)
    sec = np.sort(sec, axis=1)
    distinct = ((sec[:, 1:, :] != sec[:, :-1, :]) & (sec[:, 1:, :] >= 0)).sum() + (sec[:, 0, :] >= 0).sum()
    return distinct
for name, order in (("ascending bonds", list(range(n))), ("descending bonds", list(range(n - 1, -1, -1)))):
    for ps in (4, 8):
        d = sectors(order, ps)
        print(f"{name}: {ps} entries/sector: sector requests {d}  = {d / E:.3f} per element")
# bond-aligned (slot == bond, padded): upper bound of alignment
for ps in (4, 8):
    T = tgt
    W = (N + 31) // 32
    ell = np.vstack([T, np.full((W * 32 - N, n), -1, np.int64)]).reshape(W, 32, n)
    sec = np.sort(np.where(ell >= 0, ell // ps, -1), axis=1)
    d = ((sec[:, 1:, :] != sec[:, :-1, :]) & (sec[:, 1:, :] >= 0)).sum() + (sec[:, 0, :] >= 0).sum()
    print(f"bond-aligned slots: {ps}/sector: {d / E:.3f} per element")
# fraction of elements whose source lives on the same rank under the block-cyclic distribution
rows = np.arange(N)[:, None].repeat(n, 1)
live = tgt >= 0
for lb in (8, 10, 12):
    for P in (2, 8):
        same = ((tgt >> lb) % P) == ((rows >> lb) % P)
        print(f"block 2^{lb}, P={P}: local-source fraction {same[live].mean():.3f}; same-block fraction {((tgt >> lb) == (rows >> lb))[live].mean():.3f}")
d = np.abs(tgt - rows)[live]
for t in (32, 256, 4096, 65536, 1 << 20):
    print("distance <", t, (d < t).mean())
