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


# ---- model-sharded Bayesian model average (csghmc_fs) ---------------------------------------------------------------
class OracleBmaBackend:
    name = "oracle"

    def ce_err(self, logits, y, loss_slot, err_slot):
        loss_slot += torch.nn.functional.cross_entropy(logits, y, reduction="sum").double()
        err_slot += (logits.argmax(1) != y).sum().to(torch.int32)

    def bma_mean(self, logits_all, out):
        out.copy_(torch.from_numpy(so.bma_mean(logits_all.numpy())))


def _bma_problem(S):
    nets = []
    for j in range(S):
        net = TinyNet()
        with torch.no_grad():
            gen = torch.Generator().manual_seed(100 + j)
            for p in net.parameters():
                p.add_(0.3 * torch.randn(p.shape, generator=gen))
        nets.append(net.eval())
    gen = torch.Generator().manual_seed(2)
    loader = [(torch.randn(b, 12, generator=gen), torch.randint(0, 5, (b,), generator=gen)) for b in (7, 7, 3)]
    return nets, loader


def _bma_worker(rank, world, port, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nets, loader = _bma_problem(S)
    mine = {j: nets[j] for j in bdist.shard_models(S, rank, world)}
    r = bdist.bma_evaluate(mine, S, loader, torch.device("cpu"), rank=rank, world=world, backend=OracleBmaBackend())
    q.put((rank, r))
    dist.barrier()
    dist.destroy_process_group()


def test_model_shard_assignment_is_a_partition():
    for world in (1, 2, 3, 8):
        for S in (1, 2, 5, 16):
            assert sorted(sum((bdist.shard_models(S, r, world) for r in range(world)), [])) == list(range(S))


@pytest.mark.parametrize("S", [5, 1])
def test_sharded_bma_world2_is_bit_identical_to_world1(S):
    """S = 5: ranks hold 3 and 2 models; S = 1: rank 1 holds none and learns the class count from the exchange."""
    nets, loader = _bma_problem(S)
    one = bdist.bma_evaluate(dict(enumerate(nets)), S, loader, torch.device("cpu"), backend=OracleBmaBackend())
    # the single-rank result against the reference's statements (csghmc_fs.py:349-377)
    with torch.no_grad():
        la = torch.stack([torch.cat([net(x) for x, _ in loader]) for net in nets], 2).numpy()
    y = torch.cat([y for _, y in loader])
    assert np.array_equal(one["logits_all"], la) and np.array_equal(one["logits"], so.bma_mean(la))
    assert one["n"] == 17 and np.array_equal(one["targets"], y.numpy())
    want_loss = torch.nn.functional.cross_entropy(torch.from_numpy(so.bma_mean(la)), y, reduction="sum").item()
    assert abs(one["bma_loss_sum"] - want_loss) < 1e-5

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bma_worker, args=(r, 2, port, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, r in res:
        for k in ("logits_all", "logits", "targets", "loss_per", "err_per"):
            assert np.array_equal(r[k], one[k]), k
        assert r["bma_loss_sum"] == one["bma_loss_sum"] and r["bma_err_sum"] == one["bma_err_sum"] and r["n"] == one["n"]


# ---- sample sharding behind Runner.evaluate (hparams eval_shard=1): the exchange helpers under gloo ----------------------
def _gather_worker(rank, world, port, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert bdist.process_group() == (rank, world)
    N, K = 11, 4
    full = torch.arange(N * K * S, dtype=torch.float32).reshape(N, K, S) * 0.5 - 3.0      # full[:, :, j] = sample j
    mine = bdist.my_samples(S, rank, world)
    s_max = (S + world - 1) // world
    local = torch.zeros(N, K, s_max)
    if mine:
        local[:, :, :len(mine)] = full[:, :, mine]
    got = bdist.gather_samples(local, S, world)
    bdist.agree_across_ranks(torch.arange(5), "identical data")
    disagree = None
    try:
        bdist.agree_across_ranks(torch.arange(5) + rank, "rank-dependent data")
    except RuntimeError as e:
        disagree = str(e)
    q.put((rank, torch.equal(got, full), mine, disagree))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,S", [(2, 5), (3, 8), (2, 1)])
def test_gather_samples_restores_sample_order(world, S):
    """Ragged shares (5 samples on 2 ranks), more ranks than samples (1 sample on 2 ranks): the all-gather returns the
    [N,K,S] stack in sample order on every rank; a rank-dependent data order is detected."""
    assert bdist.process_group() == (0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(sum((m for _, _, m, _ in res), [])) == list(range(S))
    for rank, ok, mine, disagree in res:
        assert ok, f"rank {rank}: gathered stack differs"
        assert mine == [j for j in range(S) if j % world == rank]
        assert disagree is not None and "disagree" in disagree


class _StubChain:
    def __init__(self, n, fill):
        self.device = torch.device("cpu")
        self.theta = torch.full((n,), float(fill))
        self.layout = type("L", (), {"n_padded": n})()


class _StubRunner:
    """The attributes dist.broadcast_posterior touches on a drop-in Runner (methods/_base.py), on CPU tensors."""

    def __init__(self, rank, cyclical, n=24):
        self._ch = _StubChain(n, 1 + rank)
        self.net = torch.nn.BatchNorm1d(3)
        with torch.no_grad():
            self.net.running_mean.fill_(0.5 + rank)
            self.net.num_batches_tracked.fill_(7 + rank)
        self.seed, self._eval_calls = 100 + rank, 3 + 4 * rank
        if cyclical:
            cycles = [1, 2] if rank == 0 else [1, 2, 3]                   # the receiver even holds another number of cycles
            self._cyc1 = {c: torch.full((n,), 10.0 * c + rank) for c in cycles}
            self._cyc2 = {c: torch.full((n,), 20.0 * c + rank) for c in cycles if rank or c != 2}   # src: cycle 2 has no 2nd moment yet
            self.samples_per_cycle = {c: 4 + c + rank for c in cycles}
            self.cycle_likelihoods = {c: [0.1 * c + rank] for c in cycles}
            self.current_cycle = cycles[-1]
        else:
            self._mom1 = torch.full((n,), 3.0) if rank == 0 else None     # the receiver has not started collecting
            self._mom2 = torch.full((n,), 9.0) if rank == 0 else None
            self.post_theta_cnt = 5 if rank == 0 else 1

    def _chain(self):
        return self._ch


def _bcast_worker(rank, world, port, cyclical, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = _StubRunner(rank, cyclical)
    bdist.broadcast_posterior(r, src=0)
    state = dict(theta=r._ch.theta.clone(), rm=r.net.running_mean.clone(), nbt=int(r.net.num_batches_tracked), seed=r.seed,
                 calls=r._eval_calls)
    if cyclical:
        state.update(c1={c: t.clone() for c, t in r._cyc1.items()}, c2={c: t.clone() for c, t in r._cyc2.items()},
                     spc=r.samples_per_cycle, lik=r.cycle_likelihoods, cur=r.current_cycle)
    else:
        state.update(m1=r._mom1.clone(), m2=r._mom2.clone(), cnt=r.post_theta_cnt)
    mismatch = None
    if not cyclical:                                                      # a rank holding the other runner family is refused
        other = _StubRunner(rank, cyclical=(rank == 1))
        try:
            bdist.broadcast_posterior(other, src=0)
        except RuntimeError as e:
            mismatch = str(e)
    q.put((rank, state, mismatch))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cyclical", [True, False])
def test_broadcast_posterior_world2(cyclical):
    """SURVEY 8e "broadcast once": after dist.broadcast_posterior(runner, src=0) rank 1 holds rank 0's theta, buffers,
    moments (exactly rank 0's cycles), counts, likelihoods, draw seed and evaluation counter."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, cyclical, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:                                   # a rank stuck in a collective must not outlive the test
            if p.is_alive():
                p.terminate()
    want = _StubRunner(0, cyclical)
    for rank, st, mismatch in res:
        assert torch.equal(st["theta"], want._ch.theta) and torch.equal(st["rm"], want.net.running_mean)
        assert st["nbt"] == 7 and st["seed"] == 100 and st["calls"] == 3
        if cyclical:
            assert sorted(st["c1"]) == [1, 2] and sorted(st["c2"]) == [1]
            assert all(torch.equal(st["c1"][c], want._cyc1[c]) for c in st["c1"]) and torch.equal(st["c2"][1], want._cyc2[1])
            assert st["spc"] == want.samples_per_cycle and st["lik"] == want.cycle_likelihoods and st["cur"] == 2
        else:
            assert torch.equal(st["m1"], want._mom1) and torch.equal(st["m2"], want._mom2) and st["cnt"] == 5
            assert mismatch is not None and "different runner families" in mismatch, f"rank {rank}: {mismatch}"
