"""NCCL all-gather bandwidth at Krylov-vector sizes (run under torchrun on N GPUs)."""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for total_gb in [0.126, 1.0, 6.9]:
        n = int(total_gb * 1e9 / 8 / world)
        out = torch.zeros(n * world, dtype=torch.float64, device="cuda")
        shard = torch.ones(n, dtype=torch.float64, device="cuda")
        inplace_shard = out[rank * n:(rank + 1) * n]
        for name, src in (("out-of-place", shard), ("in-place", inplace_shard)):
            for _ in range(3):
                dist.all_gather_into_tensor(out, src)
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            iters = 10
            for _ in range(iters):
                dist.all_gather_into_tensor(out, src)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            if rank == 0:
                alg = n * world * 8 / ms / 1e6
                print(f"allgather {total_gb:6.3f} GB {name:12s}: {ms:8.3f} ms  algbw {alg:7.1f} GB/s  busbw {alg * (world - 1) / world:7.1f} GB/s",
                      file=sys.stderr, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
