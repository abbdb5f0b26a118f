"""GPU, two ranks: the model-sharded Bayesian model average through the product's kernels (bdl_ce_err, bdl_bma_mean) and
one real exchange step.  With >= 2 GPUs the ranks use NCCL, one GPU each; on a single-GPU box both ranks share cuda:0
and exchange through gloo (which moves CUDA tensors) -- the kernels and the host logic are the same."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu


class TinyNet(torch.nn.Module):
    readout_name = "classifier"

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.layers = torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.Tanh())
        self.classifier = torch.nn.Linear(9, 5)

    def forward(self, x):
        return self.classifier(self.layers(x.reshape(x.shape[0], -1)))


def _problem(S, device):
    nets = []
    for j in range(S):
        net = TinyNet()
        with torch.no_grad():
            gen = torch.Generator().manual_seed(100 + j)
            for p in net.parameters():
                p.add_(0.3 * torch.randn(p.shape, generator=gen))
        nets.append(net.eval().to(device))
    gen = torch.Generator().manual_seed(2)
    loader = [(torch.randn(b, 12, generator=gen), torch.randint(0, 5, (b,), generator=gen)) for b in (64, 64, 9)]
    return nets, loader


def _worker(rank, world, port, S, multi_gpu, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from bayesdll_b200 import dist as bdist
    dev = torch.device("cuda", rank if multi_gpu else 0)
    torch.cuda.set_device(dev)
    if multi_gpu:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    nets, loader = _problem(S, dev)
    mine = {j: nets[j] for j in bdist.shard_models(S, rank, world)}
    r = bdist.bma_evaluate(mine, S, loader, dev, rank=rank, world=world)
    q.put((rank, r))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("S", [5, 1])
def test_sharded_bma_two_ranks_bit_identical_to_one(cuda_device, S):
    from bayesdll_b200 import dist as bdist
    nets, loader = _problem(S, cuda_device)
    one = bdist.bma_evaluate(dict(enumerate(nets)), S, loader, cuda_device)
    assert np.array_equal(one["logits"], so.bma_mean(one["logits_all"]))          # the average itself vs the oracle
    with torch.no_grad():
        la = torch.stack([torch.cat([net(x.to(cuda_device)) for x, _ in loader]) for net in nets], 2).cpu().numpy()
    assert np.array_equal(one["logits_all"], la)
    y = torch.cat([y for _, y in loader])
    want = torch.nn.functional.cross_entropy(torch.from_numpy(one["logits"]), y, reduction="sum").item()
    assert abs(one["bma_loss_sum"] - want) < 1e-4 * max(1.0, abs(want))
    assert one["bma_err_sum"] == float((torch.from_numpy(one["logits"]).argmax(1) != y).sum())

    multi_gpu = torch.cuda.device_count() >= 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, multi_gpu, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, r in res:
        for k in ("logits_all", "logits", "targets", "loss_per", "err_per"):
            assert np.array_equal(r[k], one[k]), k
        assert r["bma_loss_sum"] == one["bma_loss_sum"] and r["bma_err_sum"] == one["bma_err_sum"] and r["n"] == one["n"]
