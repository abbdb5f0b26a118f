"""GPU, two ranks: sample-sharded ``Runner.evaluate()`` / ``full_batch_likelihoods()`` (hparams ``eval_shard=1``) through
the product's kernels and one real exchange step, against the same Runner on one rank.

With >= 2 GPUs the ranks use NCCL, one GPU each; on a single-GPU box both ranks share cuda:0 and exchange through gloo
(which moves CUDA tensors) -- kernels and host logic are the same.  Required: the reference's 5-tuple
(methods/csgld.py:434-456, methods/sgld.py:317-321) with ``targets`` / ``logits_all`` / ``logits`` BIT-identical to one
rank, identical on both ranks, calibration bin indices and counts integer-identical (north star: ECE bin counts
bit-exact), ECE / MCE / NLL identical, cycle likelihoods identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import shard_util

pytestmark = pytest.mark.gpu

# (method, nst, cycles); the last one is BASELINE.json configs[4]'s ensemble size: 8 cycles x nst 5 = 40 samples, 20 per rank
CASES = [("csgld", 2, 3), ("csghmc", 3, 3), ("adam_csghmc", 2, 3), ("csgld", 0, 3), ("sghmc", 5, 3), ("sgld", 1, 3),
         ("adam_sghmc", 0, 3), ("csgld", 5, 8)]


def _worker(rank, world, port, method, nst, cycles, multi_gpu, log_dir, q, broadcast=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dev = torch.device("cuda", rank if multi_gpu else 0)
    torch.cuda.set_device(dev)
    if multi_gpu:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    runner = None
    if broadcast:
        # rank 1 starts from ANOTHER chain (other weights, BatchNorm statistics, moments, counts, likelihoods, seed, even
        # another number of cycles); dist.broadcast_posterior must leave it with rank 0's posterior
        from bayesdll_b200 import dist as bdist
        if rank == 0:
            runner = shard_util.make_runner(method, dev, os.path.join(log_dir, f"r{rank}"), nst, True, cycles=cycles)
        else:
            runner = shard_util.make_runner(method, dev, os.path.join(log_dir, f"r{rank}"), nst, True, cycles=cycles + 1,
                                            stat_seed=99, init_seed=12)
            runner.seed = 4242
            runner._eval_calls = 7
            with torch.no_grad():
                runner.net.features[1].running_mean.add_(1.0)
            if hasattr(runner, "post_theta_cnt"):
                runner.post_theta_cnt = 3
        bdist.broadcast_posterior(runner, src=0)
    out = shard_util.run_case(method, dev, os.path.join(log_dir, f"r{rank}"), nst, eval_shard=True, cycles=cycles, runner=runner)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _same(a, b, what):
    if isinstance(a, dict):
        assert sorted(a) == sorted(b), what
        for k in a:
            _same(a[k], b[k], f"{what}.{k}")
    elif isinstance(a, np.ndarray):
        assert a.dtype == b.dtype and a.shape == b.shape, f"{what}: {a.dtype}{a.shape} vs {b.dtype}{b.shape}"
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"{what}: {np.sum(a != b)} of {a.size} entries differ"
    elif isinstance(a, (tuple, list)):
        assert len(a) == len(b), what
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{what}[{i}]")
    else:
        assert a == b, f"{what}: {a!r} != {b!r}"


@pytest.mark.parametrize("method,nst,cycles", CASES, ids=[f"{m}-nst{n}-c{c}" for m, n, c in CASES])
def test_two_rank_evaluate_is_bit_identical_to_one_rank(cuda_device, tmp_path, method, nst, cycles):
    os.makedirs(tmp_path / "one", exist_ok=True)
    one = shard_util.run_case(method, cuda_device, tmp_path / "one", nst, eval_shard=False, cycles=cycles)
    # the reference's output contract
    e = one["eval0"]
    n_rows = sum(len(y) for _, y in shard_util.make_loader())
    cyclical = "likelihoods" in one
    assert e["targets"].dtype == np.int64 and e["targets"].shape == (n_rows,)
    assert e["logits"].shape == (n_rows, shard_util.K) and e["logits"].dtype == np.float32
    assert e["logits_all"].shape == ((n_rows, shard_util.K, max(1, nst), cycles) if cyclical else (n_rows, shard_util.K, max(1, nst)))
    assert not np.array_equal(one["eval0"]["logits_all"], one["eval1"]["logits_all"]) or nst == 0    # fresh draws per call
    assert int(e["sizes"].sum()) == n_rows * shard_util.K

    multi_gpu = torch.cuda.device_count() >= 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for r in range(2):
        os.makedirs(tmp_path / f"r{r}", exist_ok=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, method, nst, cycles, multi_gpu, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        _same(out, one, f"rank{rank}")


@pytest.mark.parametrize("method,nst,cycles", [("csgld", 2, 3), ("csghmc", 3, 2), ("sghmc", 4, 3)])
def test_broadcast_posterior_then_sharded_evaluate(cuda_device, tmp_path, method, nst, cycles):
    """"every rank holds the per-cycle mom1/mom2 (loaded from the ckpt or broadcast once)".  Rank 1 starts from a
    different chain; after dist.broadcast_posterior(runner, src=0) the sample-sharded evaluation of both ranks equals the
    single-rank evaluation of rank 0's posterior bit for bit."""
    os.makedirs(tmp_path / "one", exist_ok=True)
    one = shard_util.run_case(method, cuda_device, tmp_path / "one", nst, eval_shard=False, cycles=cycles)
    multi_gpu = torch.cuda.device_count() >= 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for r in range(2):
        os.makedirs(tmp_path / f"r{r}", exist_ok=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, method, nst, cycles, multi_gpu, str(tmp_path), q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        _same(out, one, f"rank{rank}")


def test_eval_shard_outside_a_process_group_is_the_single_rank_path(cuda_device, tmp_path):
    a = shard_util.run_case("csgld", cuda_device, tmp_path, 2, eval_shard=True)       # no process group: world 1
    b = shard_util.run_case("csgld", cuda_device, tmp_path, 2, eval_shard=False)
    _same(a, b, "eval_shard=1, world 1")


def test_eval_shard_refuses_a_torch_noise_stream(cuda_device, tmp_path):
    runner = shard_util.make_runner("sghmc", cuda_device, tmp_path, 2, eval_shard=True)
    runner.noise_mode = "torch"
    with pytest.raises(ValueError, match="Philox"):
        runner.evaluate(shard_util.make_loader())
