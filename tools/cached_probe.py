#!/usr/bin/env python
"""tools/cached_probe.py DECK [STEPS]: builds the deck on one GPU, fills the operator cache and applies the
cached operator STEPS times to one device-resident vector (CUDA-event time per application).  A small
target for `ncu -k regex:cached_matvec_kernel` on decks whose headline bench legs would not fit
(heisenberg_chain_40 on one GPU: 82 GB of cache beside 16 GB of basis)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from spin_ed_b200 import config as sconfig, decks, ffi

deck = sys.argv[1] if len(sys.argv) > 1 else "heisenberg_chain_40"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ffi.setDevice(0)
uc = sconfig.toConfig(sconfig.parseConfig(decks.load(deck)))
basis, op = uc.cBasis, uc.cHamiltonian.operatorObject
t0 = time.perf_counter()
ffi.buildBasis(basis)
build_s = time.perf_counter() - t0
n = ffi.getNumberStates(basis)
dev = torch.device("cuda", 0)
real = ffi.isOperatorReal(op)
tdt, ndt = (torch.float64, np.float64) if real else (torch.complex128, np.complex128)
tag = ffi.DTYPE_TAGS[np.dtype(ndt)]
x = (torch.rand(n, dtype=torch.float64, device=dev) - 0.5).to(tdt)
x /= x.norm()
y = torch.zeros(n, dtype=tdt, device=dev)
s = torch.cuda.current_stream().cuda_stream
ffi.operatorSetCache(op, 1)
ffi.operatorMatmatDevice(op, tag, 1, x.data_ptr(), n, y.data_ptr(), n, s)  # fills the cache
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    ffi.operatorMatmatDevice(op, tag, 1, x.data_ptr(), n, y.data_ptr(), n, s)
b.record()
torch.cuda.synchronize()
rows, n_off = ffi.operatorCountElements(op)
print(json.dumps({"deck": deck, "rows": n, "offdiag_elements": n_off, "basis_build_s": build_s, "ms_per_application": a.elapsed_time(b) / steps,
                  "operator_cache": ffi.operatorCacheInfo(op), "xHx": float((x.conj() * y).sum().real.item())}))
