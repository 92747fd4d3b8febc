"""World-size-2 tests on CPU (gloo): the host-side sharding logic of the N > 1 path -- equal-chunk
row partition, padded in-place all-gather layout, rank/offset bookkeeping -- checked with the
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
        b, e = ffi.rowPartition(n, world, rank)
        chunk = -(-n // world)
        assert b == min(n, rank * chunk) and e == min(n, b + chunk)
        x = splitmix_vector(n)
        # every rank owns chunk `rank` of the padded vector; all-gather restores the global order
        xfull = torch.zeros(chunk * world, dtype=torch.float64)
        shard = torch.zeros(chunk, dtype=torch.float64)
        shard[: e - b] = torch.from_numpy(x[b:e].copy())
        dist.all_gather_into_tensor(xfull, shard)
        assert np.array_equal(xfull[:n].numpy(), x)
        # local rows of y from the replicated x, then the same gather for y
        y_local = np.zeros(n)
        oop.matmat_rows(xfull[:n].numpy().copy(), y_local, b, e, 1)
        yshard = torch.zeros(chunk, dtype=torch.float64)
        yshard[: e - b] = torch.from_numpy(y_local[b:e].copy())
        yfull = torch.zeros(chunk * world, dtype=torch.float64)
        dist.all_gather_into_tensor(yfull, yshard)
        assert np.array_equal(yfull[:n].numpy(), oop.matmat(x)), name  # bitwise: fixed term order per row
        # dot products: local partial sums all-reduced
        part = torch.tensor([float(np.dot(x[b:e], y_local[b:e]))], dtype=torch.float64)
        dist.all_reduce(part)
        assert abs(part.item() - float(np.dot(x, oop.matmat(x)))) < 1e-9
    dist.barrier()
    dist.destroy_process_group()
    ok[rank] = 1


def test_world_size_2_sharding_logic():
    world = 2
    ok = mp.Array("i", [0] * world)
    procs = [mp.Process(target=_worker, args=(r, world, 29431, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert list(ok) == [1] * world
