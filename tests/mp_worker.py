"""Multi-rank worker (one process per GPU), launched by torchrun from test_multi_gpu.py:
builds a deck on WORLD_SIZE ranks and checks representatives, matvec and eigenvalues against the
oracle and against what a single rank computes."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    from helpers import extra_configs, oracle_problem, product_problem, splitmix_vector
    from oracle import oracle as O
    from spin_ed_b200 import decks, ffi

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    ffi.setDevice(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [ffi.commUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ffi.commInit(world, rank, box[0])
    results = {}
    names = sys.argv[1:]
    for name in names:
        cfg = extra_configs()[name] if name in extra_configs() else decks.load(name)
        ob, terms = oracle_problem(O, cfg)
        ob.build()
        oop = O.Operator(ob, terms)
        n = ob.number_states
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        assert ffi.getNumberStates(uc.cBasis) == n
        assert np.array_equal(ffi.basisGetStates(uc.cBasis), ob.states), "representatives differ"
        rd = ffi.basisRowDistribution(uc.cBasis)
        ref = ffi.rowDistribution(n, world, rank)
        assert (rd.n_local, rd.chunk, rd.log2_block) == (ref.n_local, ref.chunk, ref.log2_block)
        assert len(ffi.operatorDiagonal(op := uc.cHamiltonian.operatorObject)) == rd.n_local
        dt = np.float64 if oop.is_real else np.complex128
        x = np.asfortranarray(np.stack([splitmix_vector(n, 0x5EED0001 + c, dt) for c in range(2)], axis=1))
        want = oop.matmat(x)
        for mode in (0, 1):
            ffi.operatorSetCache(op, mode)
            got = ffi.apply(op, x)  # every rank receives the full result
            err = np.linalg.norm(got - want) / np.linalg.norm(want)
            assert err < 1e-12, (name, mode, err)
        # the solver's per-matvec call: shard in, local rows out, all-gather overlapped with the
        # local-source class of the cached elements (mode 1) or plain gather + kernel (mode 0)
        mine = rd.local_rows().astype(np.int64)
        tdt = torch.float64 if oop.is_real else torch.complex128
        for mode in (0, 1):
            ffi.operatorSetCache(op, mode)
            xshard = torch.zeros(max(int(rd.chunk), 1), dtype=tdt, device="cuda")
            xshard[:len(mine)] = torch.from_numpy(np.ascontiguousarray(x[mine, 0])).cuda()
            xfull = torch.zeros(max(int(rd.chunk), 1) * world, dtype=tdt, device="cuda")
            ylocal = torch.full((max(len(mine), 1),), 7.0, dtype=tdt, device="cuda")
            for _ in range(2):  # twice: the second call reuses the events and the filled cache
                ffi.operatorMatvecSharded(op, ffi.DTYPE_TAGS[np.dtype(dt)], xshard.data_ptr(), ylocal.data_ptr(),
                                          xfull.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            if len(mine):
                diff = np.linalg.norm(ylocal[:len(mine)].cpu().numpy() - want[mine, 0])
                assert diff <= 1e-12 * np.linalg.norm(want[:, 0]), (name, mode, diff)
        # row-sharded host-pointer entry (a rank-parallel host solver): local rows in, local rows out
        ffi.operatorSetCache(op, -1)
        xl = np.asfortranarray(x[mine, :])
        yl = np.full_like(xl, 7.0)
        ffi.applyLocal(op, xl, yl)
        if len(mine):
            assert np.linalg.norm(yl - want[mine, :]) <= 1e-12 * np.linalg.norm(want), name
        rows, n_off = ffi.operatorCountElements(op)
        assert (rows, n_off) == (n, oop.count_offdiag())
        ex = ffi.expectation(op, x)
        assert np.allclose(ex, oop.expectation(x), rtol=1e-12, atol=1e-12)
        k = 2 if n > 8 else 1
        ev, vecs, rn = ffi.eigh(op, dt, k)
        if n <= 2000:
            ref = np.linalg.eigvalsh(oop.to_dense())[:k]
            assert np.all(np.abs(ev - ref) <= 1e-10 * max(1.0, abs(ref[0]))), (name, ev, ref)
        hv = oop.matmat(np.asfortranarray(vecs))
        for i in range(k):
            assert np.linalg.norm(hv[:, i] - ev[i] * vecs[:, i]) <= 1e-8 * max(1.0, abs(ev[0]))
        results[name] = {"n": n, "evals": [float(v) for v in ev]}
    # all ranks must hold identical eigenvalues (deterministic reductions + allreduce)
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    assert all(g == gathered[0] for g in gathered), gathered
    if rank == 0:
        print("MP_WORKER_OK " + json.dumps(results), flush=True)
    ffi.commFinalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
