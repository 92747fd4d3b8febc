#!/usr/bin/env python
"""tools/sanitize_target.py DECK: one pass over every kernel family on a small deck, meant to run
under `compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck` (tools/sanitize.sh): basis
build, matrix-free matvec (interpreted and NVRTC-specialised), cache fill, streaming matvec (1 and
3 columns, f64/f32 or c128/c64), expectation, block Davidson (general and single-pair/small-basis
paths).  Works on one rank or under torchrun.  tools/sanitize.sh sets SPED_FILL_CHUNK_BYTES so that the
cache fill and the code compaction run in several row chunks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import extra_configs, product_problem, splitmix_vector
from spin_ed_b200 import decks, ffi

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
ffi.setDevice(local)
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [ffi.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ffi.commInit(world, rank, box[0])
for name in sys.argv[1:]:
    cfg = extra_configs()[name] if name in extra_configs() else decks.load(name)
    uc = product_problem(cfg)
    basis, op = uc.cBasis, uc.cHamiltonian.operatorObject
    ffi.buildBasis(basis)
    n = ffi.getNumberStates(basis)
    real = ffi.isOperatorReal(op)
    wide, narrow = (np.float64, np.float32) if real else (np.complex128, np.complex64)
    results = []
    for mode in (0, 1):  # matrix-free, then cached
        ffi.operatorSetCache(op, mode)
        for dt in (wide, narrow):
            x = np.asfortranarray(np.stack([splitmix_vector(n, 11 + c, wide) for c in range(3)], axis=1).astype(dt))
            results.append(ffi.apply(op, x))
            results.append(ffi.apply(op, np.ascontiguousarray(x[:, 0])))
    assert np.allclose(results[0], results[4], rtol=0, atol=1e-11 * np.abs(results[0]).max())
    ex = ffi.expectation(op, np.asfortranarray(np.stack([splitmix_vector(n, 3, wide)], axis=1)))
    ev, vecs, rn = ffi.eigh(op, wide, min(2, n))
    if n > 3:  # one wanted pair, 3-vector basis: the fused restart + residual path of the 40-spin decks
        ev1, _, rn1 = ffi.eigh(op, wide, 1, maxBasisSize=3)
        assert abs(ev1[0] - ev[0]) <= 1e-8 * max(1.0, abs(ev[0])), (ev1, ev)
    ffi.buildBasis(basis, np.array(ffi.basisGetStates(basis)))  # ls_build_unsafe
    print(f"SANITIZE_TARGET_OK rank {rank} {name} n={n} E0={ev[0]:.10f} rnorm={rn[0]:.1e} launches={ffi.kernelLaunches()}", flush=True)
if world > 1:
    ffi.commFinalize()
    dist.destroy_process_group()
