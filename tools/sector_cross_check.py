#!/usr/bin/env python
"""tools/sector_cross_check.py DECK [KEEP]: independent check of a ground-state energy at full size.

The deck's sector and a LARGER sector that contains it -- the same Hamiltonian with only the first
KEEP symmetry generators (default: translations only; spin inversion and hamming weight kept) --
are solved separately.  Different symmetry group, different representatives, different norms and
characters, a basis several times larger: the only thing the two runs share is the physics, so the
lowest eigenvalues must agree (the ground state of these decks lies in the fully symmetric
sector).  Prints one JSON line; exit code 1 unless |E0 - E0'| <= 1e-10 |E0|.  One rank or torchrun."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from spin_ed_b200 import config as sconfig, decks, ffi

name = sys.argv[1] if len(sys.argv) > 1 else "heisenberg_square_6x6"
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
ffi.setDevice(local)
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [ffi.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ffi.commInit(world, rank, box[0])
full = decks.load(name)
keep = int(sys.argv[2]) if len(sys.argv) > 2 else (2 if "square" in name else 1)
reduced = json.loads(json.dumps(full))
reduced["basis"]["symmetries"] = reduced["basis"]["symmetries"][:keep]
out = {"deck": name, "world": world}
for label, cfg in (("deck_sector", full), ("larger_sector", reduced)):
    spec = sconfig.parseConfig(cfg)
    uc = sconfig.toConfig(spec)
    t0 = time.perf_counter()
    ffi.buildBasis(uc.cBasis)
    t1 = time.perf_counter()
    op = uc.cHamiltonian.operatorObject
    dt = np.float64 if ffi.isOperatorReal(op) else np.complex128
    ev, _, rn = ffi.eigh(op, dt, 1, 0.0, spec.max_primme_basis_size, spec.max_primme_block_size, spec.min_primme_restart_size,
                         want_vectors=False)
    t2 = time.perf_counter()
    st = ffi.eighLastStats(op)
    out[label] = {"generators": len(cfg["basis"]["symmetries"]), "dimension": ffi.getNumberStates(uc.cBasis), "E0": float(ev[0]),
                  "rnorm": float(rn[0]), "matvecs": st["matvecs"], "build_s": t1 - t0, "solve_s": t2 - t1}
    ffi.operatorReleaseWorkspace(op)
    ffi.operatorSetCache(op, -1)
    del uc, op
a, b = out["deck_sector"]["E0"], out["larger_sector"]["E0"]
out["relative_difference"] = abs(a - b) / abs(a)
out["tolerance"] = 1e-10
out["agree"] = out["relative_difference"] <= 1e-10
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    ffi.commFinalize()
    dist.destroy_process_group()
sys.exit(0 if out["agree"] else 1)
