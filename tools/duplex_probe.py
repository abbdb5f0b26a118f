"""Copy-engine-only ceiling of the host-buffer step (bdl_chain_step_host): what do H2D and D2H of one ViT-L/32 vector
(1.22 GB each) reach on this box with NO kernel in between?

    python tools/duplex_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/duplex_probe.py   # all ranks at once (shared PCIe root)

Per rank: H2D alone, D2H alone, both at once on two streams (whole vector, and cut into 8 Mi-element chunks like the
pipeline of bdl_host.cu), plus the host's own memcpy rate (numpy copy of the same vector) -- the host memory system all
ranks share.  The e2e leg of bench.py moves the same bytes per step, so `duplex` is its ceiling.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 305_548_328


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    h_in = torch.empty(N, dtype=torch.float32).pin_memory()
    h_out = torch.empty(N, dtype=torch.float32).pin_memory()
    h_in.normal_()
    d_in, d_out = torch.empty(N, device=dev), torch.randn(N, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    nbytes = 4 * N
    chunk = 8 << 20

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=5):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    def both_chunked():
        for a in range(0, N, chunk):
            b = min(N, a + chunk)
            with torch.cuda.stream(s1):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)

    def both_chunked_multi(chunk_elems, lanes):
        """`lanes` streams per direction, chunks dealt round-robin: more than one copy in flight per direction."""
        k = 0
        for a in range(0, N, chunk_elems):
            b = min(N, a + chunk_elems)
            with torch.cuda.stream(up[k % lanes]):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            with torch.cuda.stream(down[k % lanes]):
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)
            k += 1

    up = [torch.cuda.Stream(dev) for _ in range(4)]
    down = [torch.cuda.Stream(dev) for _ in range(4)]
    res = {"h2d_alone_gbs": nbytes / timed(h2d) / 1e9, "d2h_alone_gbs": nbytes / timed(d2h) / 1e9}
    sweep = {}
    for mi in (2, 8, 32):
        for lanes in (1, 2, 4):
            t = timed(lambda: both_chunked_multi(mi << 20, lanes), reps=4)
            sweep[f"{mi}Mi_x{lanes}"] = round(nbytes / t / 1e9, 2)
    res["duplex_chunked_lanes_each_way_gbs"] = sweep
    t = timed(both)
    res["duplex_each_way_gbs"] = nbytes / t / 1e9
    res["duplex_ms"] = t * 1e3
    t = timed(both_chunked)
    res["duplex_chunked_each_way_gbs"] = nbytes / t / 1e9
    res["duplex_chunked_ms"] = t * 1e3
    a, b = h_in.numpy(), h_out.numpy()
    np.copyto(b, a)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        np.copyto(b, a)
    res["host_memcpy_gbs_read_plus_write"] = 3 * 2 * nbytes / (time.perf_counter() - t0) / 1e9
    # the product's host-buffer step on the same buffers
    try:
        from bayesdll_b200 import _lib, ops, shapes
        from bayesdll_b200.flat import FlatLayout
        named, readout = shapes.named_shapes("vit_l_32", 37)
        lay = FlatLayout(named, readout)
        assert lay.n_padded == N
        chain = ops.HostChain(N, _lib.SGHMC)
        sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, alpha=0.18)
        tab = lay.run_table("informative")
        h_in.mul_(0.01)
        k = [0]

        def step():
            k[0] += 1
            chain.step_host(h_in, h_out, tab, sc, ops.make_noise(seed=1, subseq=k[0]))
        t = timed(step, reps=5)
        res["bdl_chain_step_host_ms"] = t * 1e3
        res["bdl_chain_step_host_each_way_gbs"] = nbytes / t / 1e9
        chain.close()
    except Exception as e:  # noqa: BLE001
        res["bdl_chain_step_host_error"] = f"{type(e).__name__}: {e}"
    res["cores"] = len(os.sched_getaffinity(0))
    if world > 1:
        import torch.distributed as dist
        allres = [None] * world
        dist.all_gather_object(allres, res)
        if rank == 0:
            agg = {k: [r.get(k) for r in allres] for k in res}
            print(json.dumps({"world": world, "per_rank": agg}))
        dist.barrier()
        dist.destroy_process_group()
    else:
        print(json.dumps({"world": 1, **res}))


if __name__ == "__main__":
    main()
