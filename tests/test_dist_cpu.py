"""CPU, gloo, world_size 2: host logic of the sample-sharded ensemble (assignment, single all-reduce, finalisation,
row-sharded calibration).  The CUDA kernels are replaced by an oracle-backed stand-in *in this test only*; the
product default (bayesdll_b200.dist.CudaBackend) has no CPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from bayesdll_b200 import _lib
from bayesdll_b200 import dist as bdist
from bayesdll_b200.flat import FlatLayout
from oracle import c_oracle
from oracle import sampler_oracle as so


class OracleBackend:
    name = "oracle"

    def draw(self, comp, out_flat, seed, subseq, div_mode):
        n = out_flat.numel()
        eps = c_oracle.philox_normal(n, seed, _lib.STREAM_DRAW, subseq)
        mean, second = comp["mean"].numpy(), comp["second"].numpy()
        var = so.variance_from_moments(mean, second, comp["scale"])
        out_flat.copy_(torch.from_numpy(so.posterior_draw(mean, var, eps)))

    def lse_accum(self, logits, m, s):
        ls = torch.log_softmax(logits, 1)
        mn = torch.maximum(m, ls)
        keep = torch.where(torch.isinf(m), torch.zeros_like(s), s * torch.exp(m - mn))
        s.copy_(keep + torch.exp(ls - mn))
        m.copy_(mn)

    def lse_rescale(self, m_local, m_global, s):
        s.copy_(torch.where(torch.isinf(m_local), torch.zeros_like(s), s * torch.exp(m_local - m_global)))

    def lse_finalize(self, m, s, out, n_samples, weight, mode):
        comp = (torch.log(s) + m) - np.float32(np.log(n_samples))
        if mode == 0:
            out.copy_(comp)
        elif mode == 1:
            out.copy_(np.float32(weight) * comp)
        else:
            out += np.float32(weight) * comp

    def ce_err(self, logits, y):
        loss = torch.nn.functional.cross_entropy(logits, y, reduction="sum").double().reshape(1)
        err = (logits.argmax(1) != y).sum().to(torch.int32).reshape(1)
        return loss, err

    def calibrate(self, logits, labels, edges):
        lg, lb = logits.numpy(), labels.numpy()
        M = edges.numel()
        _, _, accs, confs, sizes = so.calc_bins(lb, lg, M)
        mx = lg.max(1, keepdims=True)
        nll = (np.log(np.exp(lg - mx).sum(1)) + mx[:, 0] - lg[np.arange(len(lb)), lb]).sum()
        return torch.from_numpy(np.concatenate([sizes, accs * sizes, confs * sizes, [nll, 0.0]]))


class TinyNet(torch.nn.Module):
    readout_name = "classifier"

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.layers = torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.Tanh())
        self.classifier = torch.nn.Linear(9, 5)

    def forward(self, x):
        return self.classifier(self.layers(x.reshape(x.shape[0], -1)))


def _problem():
    net = TinyNet()
    lay = FlatLayout.from_module(net)
    gen = torch.Generator().manual_seed(1)
    theta = lay.from_dense(torch.cat([p.detach().reshape(-1) for p in net.parameters()]))
    comps = []
    for c in (1, 2, 3):
        mean = theta + 0.05 * torch.randn(lay.n_padded, generator=gen)
        second = mean * mean + 1e-3 * torch.rand(lay.n_padded, generator=gen)
        comps.append(dict(cycle=c, mean=mean, second=second, var_mode=0, scale=1.25, weight=[0.5, 0.3, 0.2][c - 1]))
    loader = [(torch.randn(7, 12, generator=gen), torch.randint(0, 5, (7,), generator=gen)) for _ in range(3)]
    return net, lay, comps, loader


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net, lay, comps, loader = _problem()
    ens = bdist.ShardedEnsemble(net, lay, comps, nst=3, seed=9, mixture=True, rank=rank, world=world, backend=OracleBackend())
    loss, err, targets, logits = ens.evaluate(loader)
    ece, mce, nll = ens.calibrate(targets, logits, 10)
    q.put((rank, loss, err, targets, logits, ece, mce, nll, len(ens.mine)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_assignment_is_a_partition():
    for world in (1, 2, 3, 4, 8, 64):
        seen = sorted(sum((bdist.shard_samples(8, 5, r, world) for r in range(world)), []))
        assert seen == [(c, s) for c in range(8) for s in range(5)]
        sizes = [len(bdist.shard_samples(8, 5, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert bdist.chain_seed(42, 3) == 45


def test_sharded_ensemble_world2_equals_world1():
    net, lay, comps, loader = _problem()
    single = bdist.ShardedEnsemble(net, lay, comps, nst=3, seed=9, mixture=True, backend=OracleBackend())
    loss1, err1, targets1, logits1 = single.evaluate(loader)
    ece1, mce1, nll1 = single.calibrate(targets1, logits1, 10)
    # oracle check of the single-rank result itself
    e2, m2, n2 = so.analyze(targets1, logits1, 10)
    assert abs(ece1 - e2) < 1e-9 and abs(mce1 - m2) < 1e-9 and abs(nll1 - n2) < 1e-6

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[8] for r in res] == [5, 4]                       # 9 samples dealt round-robin
    for (_, loss, err, targets, logits, ece, mce, nll, _) in res:
        assert np.array_equal(targets, targets1)
        np.testing.assert_allclose(logits, logits1, rtol=2e-6, atol=2e-6)   # only the fp32 summation order differs
        assert abs(loss - loss1) < 1e-5 and err == err1
        assert abs(ece - ece1) < 1e-6 and abs(mce - mce1) < 1e-6 and abs(nll - nll1) < 1e-6
    # both ranks hold identical results
    np.testing.assert_array_equal(res[0][4], res[1][4])
