"""Does this torch build's gloo backend move CUDA tensors (all_gather_into_tensor / all_reduce)?  Two ranks, one GPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", 0)
    a = torch.full((6,), float(rank + 1), device=dev)
    out = torch.empty(12, device=dev)
    try:
        dist.all_gather_into_tensor(out, a)
        print(rank, "all_gather_into_tensor ok", out.tolist(), flush=True)
    except Exception as e:
        print(rank, "all_gather_into_tensor FAILED", repr(e)[:200], flush=True)
    try:
        dist.all_reduce(a)
        print(rank, "all_reduce ok", a.tolist(), flush=True)
    except Exception as e:
        print(rank, "all_reduce FAILED", repr(e)[:200], flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    mp.spawn(worker, args=(2, int(sys.argv[1]) if len(sys.argv) > 1 else 29533), nprocs=2)
