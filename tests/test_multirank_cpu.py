"""World-size-2 tests on CPU (gloo): the host-side sharding logic of the N > 1 path -- block-cyclic
row distribution, padded [rank][local] all-gather layout, rank/offset bookkeeping -- checked with the
oracle standing in for the device kernel (test infrastructure only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ok):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import oracle_problem, splitmix_vector
    from oracle import oracle as O
    from spin_ed_b200 import decks, ffi

    # the 128-byte communicator id travels by broadcast_object_list exactly as in bench.py
    box = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert box[0] == bytes(range(128))
    for name in ["heisenberg_chain_10", "heisenberg_square_4x4", "heisenberg_kagome_12"]:
        ob, terms = oracle_problem(O, decks.load(name))
        ob.build()
        oop = O.Operator(ob, terms)
        n = ob.number_states
        rd = ffi.rowDistribution(n, world, rank)
        rows = rd.local_rows().astype(np.int64)  # block-cyclic rows of this rank, ascending
        chunk, n_loc = int(rd.chunk), int(rd.n_local)
        assert len(rows) == n_loc
        x = splitmix_vector(n)
        # every rank owns shard `rank` of the padded [rank][local] vector; an all-gather of the
        # local shards produces exactly that layout
        xfull = torch.zeros(chunk * world, dtype=torch.float64)
        shard = torch.zeros(chunk, dtype=torch.float64)
        shard[:n_loc] = torch.from_numpy(x[rows].copy())
        dist.all_gather_into_tensor(xfull, shard)
        pos = rd.global_to_position(np.arange(n, dtype=np.uint64)).astype(np.int64)
        assert np.array_equal(xfull.numpy()[pos], x)  # global order is recovered through the position map
        # local rows of y from the replicated x (oracle row subset), then the same gather for y
        x_global = xfull.numpy()[pos].copy()
        y_rows = np.zeros(n)
        for r in rows:  # oracle entry point takes (lo, hi, stride): one row at a time is fine at this size
            oop.matmat_rows(x_global, y_rows, int(r), int(r) + 1, 1)
        yshard = torch.zeros(chunk, dtype=torch.float64)
        yshard[:n_loc] = torch.from_numpy(y_rows[rows].copy())
        yfull = torch.zeros(chunk * world, dtype=torch.float64)
        dist.all_gather_into_tensor(yfull, yshard)
        assert np.array_equal(yfull.numpy()[pos], oop.matmat(x)), name  # bitwise: fixed term order per row
        # dot products: local partial sums all-reduced
        part = torch.tensor([float(np.dot(x[rows], y_rows[rows]))], dtype=torch.float64)
        dist.all_reduce(part)
        assert abs(part.item() - float(np.dot(x, oop.matmat(x)))) < 1e-9
    dist.barrier()
    dist.destroy_process_group()
    ok[rank] = 1


def test_world_size_2_sharding_logic():
    world = 2
    ctx = mp.get_context("spawn")  # never fork a process that already runs OpenMP threads (the oracle)
    ok = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, 29431, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert list(ok) == [1] * world
