#!/usr/bin/env python
"""tools/block_bench.py [DECK]: device-resident block applications (1..4 columns, f64 and f32 storage)
of the cached operator on one GPU, CUDA-event times.  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from spin_ed_b200 import config as sconfig, decks, ffi

deck = sys.argv[1] if len(sys.argv) > 1 else "heisenberg_square_6x6"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ffi.setDevice(0)
uc = sconfig.toConfig(sconfig.parseConfig(decks.load(deck)))
basis, op = uc.cBasis, uc.cHamiltonian.operatorObject
ffi.buildBasis(basis)
n = ffi.getNumberStates(basis)
dev = torch.device("cuda", 0)
out = {"deck": deck, "rows": n}
real = ffi.isOperatorReal(op)
for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)) if real else ((torch.complex128, np.complex128),):
    tag = ffi.DTYPE_TAGS[np.dtype(ndt)]
    x = torch.rand(4, n, dtype=torch.float64, device=dev).to(tdt) - 0.5
    y = torch.zeros(4, n, dtype=tdt, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ref = None
    for nc in (1, 2, 3, 4):
        for _ in range(3):
            ffi.operatorMatmatDevice(op, tag, nc, x.data_ptr(), n, y.data_ptr(), n, s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            ffi.operatorMatmatDevice(op, tag, nc, x.data_ptr(), n, y.data_ptr(), n, s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        if nc == 1:
            ref = y[0].clone()
            err = 0.0
        else:  # column 0 of a block application equals the single-column application up to rounding
            err = float(((y[0] - ref).abs().max() / ref.abs().max()).item())
        out[f"{np.dtype(ndt).name}_cols{nc}_ms"] = round(ms, 4)
        out[f"{np.dtype(ndt).name}_cols{nc}_col0_maxrel"] = err
print(json.dumps(out))
