#!/usr/bin/env python
"""tools/eigh_probe.py DECK: consecutive sped_eigh calls on one operator, wall time around the C call
next to the solver's own statistics (where does time-to-ground-state go outside the solver?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from spin_ed_b200 import config as sconfig, decks, ffi

deck = sys.argv[1] if len(sys.argv) > 1 else "heisenberg_square_6x6"
ffi.setDevice(0)
spec = sconfig.parseConfig(decks.load(deck))
uc = sconfig.toConfig(spec)
basis, op = uc.cBasis, uc.cHamiltonian.operatorObject
ffi.buildBasis(basis)
dt = np.float64 if ffi.isOperatorReal(op) else np.complex128
for label, drop, torch_alloc in (("first", False, False), ("second (cache kept)", False, False), ("cache dropped", True, False),
                                 ("cache dropped, torch holds 1 GB", True, True), ("cache kept again", False, False)):
    if drop:
        ffi.operatorSetCache(op, -1)
    hold = torch.zeros(1 << 27, dtype=torch.float64, device="cuda") if torch_alloc else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev, _, rn = ffi.eigh(op, dt, spec.number_vectors, spec.precision, spec.max_primme_basis_size, spec.max_primme_block_size,
                         spec.min_primme_restart_size, want_vectors=False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = ffi.eighLastStats(op)
    print(f"{deck} {label}: call {t1 - t0:.3f}s (+sync {t2 - t1:.3f}s) solver total {st['seconds_total']:.3f}s matvec {st['seconds_matvec']:.3f} "
          f"ortho {st['seconds_ortho']:.3f} resid {st['seconds_residual']:.3f} proj {st['seconds_project']:.3f} restart {st['seconds_restart']:.3f} "
          f"E0 {ev[0]:.10f} matvecs {st['matvecs']} iters {st['iterations']}", flush=True)
    del hold
