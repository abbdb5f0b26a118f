"""Multi-GPU paths.  The reference is single-device; everything here is new surface.

* Independent chains: one process per GPU (``torchrun``), distinct seeds, **no communication** -- nothing to do
  here beyond ``chain_seed``.
* Sample-sharded evaluation behind the unchanged Runner API (hparams ``eval_shard=1``; methods/_base.py): the
  S = cycles x nst posterior samples are dealt round-robin to the ranks (``my_samples``); every rank draws its samples
  with the counter-based Philox stream keyed by (evaluation, batch, cycle, sample) -- so the draws do not depend on the
  rank count --, runs the backbone forward, and ONE all-gather of the local sample logits (``gather_samples``) rebuilds
  the reference's ``logits_all`` on every rank; all later reductions run on the full stack with the kernels one rank
  runs: results bit-identical to one rank.  ``agree_across_ranks`` makes ranks that walked different data fail loudly,
  ``broadcast_posterior`` hands every rank one rank's chain state and posterior statistics ("broadcast once").
* ``ShardedEnsemble``: the variant for callers that do not need ``logits_all`` -- a running logsumexp_s
  log_softmax(logits_s) per cycle as (max, scaled sum), ONE exchange step (all-reduce MAX then SUM of ``[C, N, K]`` fp32,
  4.3 MB each at Pets size), then log / GMM mixture / CE.
* ``calibrate_sharded``: calibration bins with the rows dealt to the ranks and one all-reduce of the 3*M+2 bin statistics.
* ``bma_evaluate``: model-sharded Bayesian model average of csghmc_fs (one all-gather).
"""
import copy

import numpy as np
import torch

from . import _lib, ops
from .flat import adopt_parameters, alloc_flat
from .graphfwd import GraphedForward


def chain_seed(base_seed, rank):
    """Seeds 42, 43, ... for chains 0, 1, ... (BASELINE.md cfg 4)."""
    return int(base_seed) + int(rank)


def shard_samples(n_components, nst, rank, world):
    """Round-robin assignment of the flat sample index j = component * nst + s."""
    return [(j // nst, j % nst) for j in range(n_components * nst) if j % world == rank]


class CudaBackend:
    """The product's kernels.  (Tests inject an oracle-backed stand-in to exercise the host logic under gloo on CPU.)"""
    name = "cuda"

    def draw(self, comp, out_flat, seed, subseq, div_mode):
        ops.draw(comp["mean"], comp["second"], out_flat, comp["var_mode"], comp["scale"],
                 ops.make_noise(seed=seed, subseq=subseq, stream_id=_lib.STREAM_DRAW), div_mode)

    def lse_accum(self, logits, m, s):
        ops.lse_accum(logits, m, s)

    def lse_rescale(self, m_local, m_global, s):
        ops.lse_rescale(m_local, m_global, s)

    def lse_finalize(self, m, s, out, n_samples, weight, mode):
        ops.lse_finalize(m, s, out, n_samples, weight, mode)

    def ce_err(self, logits, y):
        loss = torch.zeros(1, dtype=torch.float64, device=logits.device)
        err = torch.zeros(1, dtype=torch.int32, device=logits.device)
        for i in range(0, logits.shape[0], 4096):            # the kernel is a single CTA per call
            ops.ce_err(logits[i:i + 4096].contiguous(), y[i:i + 4096].contiguous(), loss, err)
        return loss, err

    def calibrate(self, logits, labels, edges):
        size, acc, conf, nll, near, _ = ops.calibrate(logits, labels, edges)
        return torch.cat([size, acc, conf, nll, near.double()])


def _pack_subseq(eval_id, batch, cycle, sample):
    return ((eval_id & 0xFFFF) << 48) | ((batch & 0xFFFFFF) << 24) | ((cycle & 0xFF) << 16) | (sample & 0xFFFF)


class ShardedEnsemble:
    """components: list of dicts {cycle, mean, second, var_mode, scale, weight} (padded flat tensors on this rank's
    device).  ``mixture=False`` (non-cyclical runners): exactly one component, logits = log mean_s softmax.
    ``mixture=True``: weighted sum of per-cycle log-mean-softmax (the reference's log-space GMM, Appendix B.5)."""

    def __init__(self, net, layout, components, nst, seed, *, mixture, eval_id=1, rank=0, world=1, group=None,
                 div_mode=_lib.DIV_RECIP, backend=None, graph=True):
        if nst < 1:
            raise ValueError("sample sharding needs nst >= 1")
        self.layout, self.components, self.nst, self.seed = layout, components, int(nst), int(seed)
        self.mixture, self.eval_id, self.rank, self.world, self.group = mixture, eval_id, rank, world, group
        self.div_mode = div_mode
        self.backend = backend or CudaBackend()
        self.net = copy.deepcopy(net).eval()
        dev = next(net.parameters()).device
        self.flat = alloc_flat(layout.n_padded, dev) if dev.type == "cuda" else torch.zeros(layout.n_padded)
        adopt_parameters(self.net, layout, self.flat)
        self.forward = GraphedForward(self.net, enabled=graph and dev.type == "cuda")   # stable pointers: replay as a CUDA graph
        self.mine = shard_samples(len(components), self.nst, rank, world)

    def _all_reduce(self, t, op="sum"):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=self.group)
        return t

    def evaluate(self, loader):
        """-> (loss, err, targets[N] int64 numpy, logits[N,K] f32 numpy); identical on every rank."""
        dev = self.flat.device
        C = len(self.components)
        ms, ss, ys = [], [], []
        with torch.no_grad():
            for b_idx, (x, y) in enumerate(loader):
                x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
                m = s = None
                for (ci, smp) in self.mine:
                    comp = self.components[ci]
                    self.backend.draw(comp, self.flat, self.seed, _pack_subseq(self.eval_id, b_idx, comp["cycle"], smp),
                                      self.div_mode)
                    out = self.forward(x).float().contiguous()
                    if m is None:
                        m = torch.full((C,) + tuple(out.shape), float("-inf"), dtype=torch.float32, device=dev)
                        s = torch.zeros_like(m)
                    self.backend.lse_accum(out, m[ci], s[ci])
                if m is None:                                 # more ranks than samples: contribute the neutral element
                    k = self.net(x).shape[1]
                    m = torch.full((C, x.shape[0], k), float("-inf"), dtype=torch.float32, device=dev)
                    s = torch.zeros_like(m)
                ms.append(m)
                ss.append(s)
                ys.append(y)
        m_loc = torch.cat(ms, dim=1).contiguous()            # [C, N, K] running max of the local samples
        s_loc = torch.cat(ss, dim=1).contiguous()
        if self.world > 1:                                   # the path's only data exchange: MAX then SUM over [C,N,K]
            m_glob = self._all_reduce(m_loc.clone(), "max")
            self.backend.lse_rescale(m_loc, m_glob, s_loc)
            self._all_reduce(s_loc, "sum")
        else:
            m_glob = m_loc
        y = torch.cat(ys)
        N, K = m_glob.shape[1], m_glob.shape[2]
        logits = torch.empty(N, K, dtype=torch.float32, device=dev)
        for ci, comp in enumerate(self.components):
            if self.mixture:
                self.backend.lse_finalize(m_glob[ci], s_loc[ci], logits, self.nst, comp["weight"], 1 if ci == 0 else 2)
            else:
                self.backend.lse_finalize(m_glob[ci], s_loc[ci], logits, self.nst, 1.0, 0)
        loss, err = self.backend.ce_err(logits, y)
        return loss.item() / N, err.item() / N, y.cpu().numpy(), logits.cpu().numpy()

    def calibrate(self, targets, logits, num_bins):
        """ECE / MCE / NLL with the rows of ``logits`` sharded over the ranks -> (ece, mce, nll)."""
        return calibrate_sharded(targets, logits, num_bins, self.flat.device, self.rank, self.world, self.group, self.backend)


def calibrate_sharded(targets, logits, num_bins, device, rank=0, world=1, group=None, backend=None):
    """calibration.analyze (calibration.py:215-249) with the rows dealt round-robin to the ranks: every rank bins its
    rows (bdl_calibrate), ONE all-reduce(SUM) of the 3*M+2 fp64 statistics (bin sizes, accuracy sums, confidence sums,
    NLL sum, near-edge count) follows.  Bin sizes / accuracy sums are integer-valued, so they are exact for any number
    of ranks.  -> (ece, mce, nll)."""
    from .calibration import bin_edges
    backend = backend or CudaBackend()
    lg = torch.as_tensor(logits, dtype=torch.float32)[rank::world].contiguous().to(device)
    lb = torch.as_tensor(targets, dtype=torch.int64)[rank::world].contiguous().to(device)
    edges = torch.from_numpy(bin_edges(num_bins)).to(device)
    M = num_bins
    if lg.shape[0] > 0:
        stats = backend.calibrate(lg, lb, edges)
    else:
        stats = torch.zeros(3 * M + 2, dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(stats, group=group)
    h = stats.cpu().numpy()
    sizes, acc_sum, conf_sum, nll_sum = h[:M], h[M:2 * M], h[2 * M:3 * M], h[3 * M]
    nz = sizes > 0
    accs, confs = np.zeros(M), np.zeros(M)
    accs[nz], confs[nz] = acc_sum[nz] / sizes[nz], conf_sum[nz] / sizes[nz]
    ece = (np.abs(accs - confs) * (sizes / sizes.sum())).sum()
    mce = np.abs(accs - confs).max()
    return ece, mce, nll_sum / len(targets)


def components_from_runner(runner):
    """Build the component list of a trained drop-in Runner (burn-in or cyclical)."""
    if hasattr(runner, "_cyc1"):
        w = runner.calculate_gmm_weights()
        comps = []
        for c in runner._cyc1:
            if w.get(c, 0.0) < 1e-10:
                continue
            second, var_mode, scale = runner._cycle_variance_spec(c)
            comps.append(dict(cycle=c, mean=runner._cyc1[c], second=second, var_mode=var_mode, scale=scale, weight=w.get(c, 0.0)))
        return comps, True
    return [dict(cycle=0, mean=runner._mom1, second=runner._mom2, var_mode=ops.VAR_FROM_MOMENTS,
                 scale=runner._variance_ratio(), weight=1.0)], False


# ------------------------------------------------------------------------------------------------------------
# Sample sharding behind Runner.evaluate() / full_batch_likelihoods()  (hparams eval_shard=1)
# ------------------------------------------------------------------------------------------------------------
def process_group():
    """(rank, world) of the default process group, (0, 1) outside one."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def my_samples(n_samples, rank, world):
    """Round-robin share of the flat sample index j = component * nst + s."""
    return [j for j in range(n_samples) if j % world == rank]


def gather_samples(local, n_samples, world, group=None):
    """The exchange step of a sample-sharded evaluation that must return the reference's ``logits_all``:
    ``local`` [N, K, s_max] holds this rank's samples in the order of ``my_samples`` (zero-padded to
    s_max = ceil(S / world)); ONE all-gather returns the full [N, K, S] stack in sample order on every rank.
    Sample j = r + world * i sits at gathered[r, :, :, i].  Every later reduction runs on the full stack with the very
    kernels the single-rank path uses, so its results do not depend on the number of ranks (bit-identical)."""
    import torch.distributed as dist
    N, K, s_max = local.shape
    assert s_max == (n_samples + world - 1) // world
    gathered = torch.empty((world, N, K, s_max), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered.view(-1), local.contiguous().view(-1), group=group)   # flat: one shape rule for nccl and gloo
    return gathered.permute(1, 2, 3, 0).reshape(N, K, s_max * world)[:, :, :n_samples].contiguous()


def agree_across_ranks(t, what, group=None):
    """Every rank must have walked the same data (same loader order): compare an order-sensitive checksum."""
    import torch.distributed as dist
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64, device=t.device)
    h = torch.stack([(t.reshape(-1).double() * w).sum(), torch.tensor(float(t.numel()), dtype=torch.float64, device=t.device)])
    lo, hi = h.clone(), h.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise RuntimeError(f"eval_shard: the ranks disagree on {what} -- every rank must iterate the same data in the same "
                           f"order (no shuffling, no DistributedSampler) and hold the same chain state")


def broadcast_posterior(runner, src=0, group=None):
    """Hand every rank of the process group rank ``src``'s chain state and posterior statistics -- what a sample-sharded
    ``evaluate()`` / ``full_batch_likelihoods()`` (hparams ``eval_shard=1``) needs the ranks to agree on ("every
    rank holds the per-cycle mom1/mom2, loaded from the ckpt or broadcast once"): theta, the network's buffers (BatchNorm
    running statistics), the running moments with their counts (burn-in runners: ``post_theta_mom1/2``, ``post_theta_cnt``;
    cyclical runners: ``cycle_theta_mom1/2``, ``samples_per_cycle``, ``cycle_likelihoods``, ``current_cycle``) and the seed
    of the draws.  One object broadcast of the small metadata, then one tensor broadcast per flat vector (padded layout,
    1.22 GB each at ViT-L/32 size: NCCL over NVLink).  Sampler state that only training uses (momentum, Adam moments,
    SGD buffers) stays local -- the ranks may go on as independent chains afterwards."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    gsrc = dist.get_global_rank(group, src) if group is not None else src
    ch = runner._chain()
    dev, n = ch.device, ch.layout.n_padded
    cyclical = hasattr(runner, "_cyc1")
    if rank == src:
        if cyclical:
            meta = dict(kind="cyclical", cyc1=sorted(runner._cyc1), cyc2=sorted(runner._cyc2),
                        samples_per_cycle=dict(runner.samples_per_cycle), cycle_likelihoods=dict(runner.cycle_likelihoods),
                        current_cycle=runner.current_cycle)
        else:
            meta = dict(kind="burnin", has=runner._mom1 is not None, post_theta_cnt=getattr(runner, "post_theta_cnt", 0))
        meta.update(seed=runner.seed, n=n, eval_calls=runner._eval_calls)
    else:
        meta = None
    box = [meta]
    dist.broadcast_object_list(box, src=gsrc, group=group)
    meta = box[0]
    # every rank must take the same decision, or the ranks that go on would wait forever for the one that raised
    ok = torch.tensor([int(meta["kind"] == ("cyclical" if cyclical else "burnin") and meta["n"] == n)], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if not int(ok.item()):
        raise RuntimeError("broadcast_posterior: the ranks hold different runner families / layouts")

    def bc(t):
        dist.broadcast(t, src=gsrc, group=group)
        return t

    def recv(t):                                             # receivers reuse a buffer of the right size, else allocate one
        if t is None or t.numel() != n:
            t = alloc_flat(n, dev)
        return bc(t)

    bc(ch.theta)
    for b in runner.net.buffers():
        bc(b.data)
    if cyclical:
        for attr, cycles in (("_cyc1", meta["cyc1"]), ("_cyc2", meta["cyc2"])):
            have = getattr(runner, attr)
            setattr(runner, attr, {c: recv(have.get(c)) for c in cycles})
        runner.samples_per_cycle = dict(meta["samples_per_cycle"])
        runner.cycle_likelihoods = dict(meta["cycle_likelihoods"])
        runner.current_cycle = meta["current_cycle"]
    elif meta["has"]:
        runner._mom1, runner._mom2 = recv(runner._mom1), recv(runner._mom2)
        runner.post_theta_cnt = meta["post_theta_cnt"]
    runner.seed = meta["seed"]
    runner._eval_calls = meta["eval_calls"]                  # the Philox sub-sequence of the next evaluation: same on every rank
    return runner


# ------------------------------------------------------------------------------------------------------------
# Model-sharded Bayesian model average (csghmc_fs)
# ------------------------------------------------------------------------------------------------------------
def shard_models(n_models, rank, world):
    """Round-robin assignment of the model index j (position in the reference's sorted file list)."""
    return [j for j in range(n_models) if j % world == rank]


class CudaBmaBackend:
    name = "cuda"

    def ce_err(self, logits, y, loss_slot, err_slot):
        ops.ce_err(logits, y, loss_slot, err_slot)

    def bma_mean(self, logits_all, out):
        ops.bma_mean(logits_all, out)


def bma_evaluate(nets, n_models, loader, device, *, rank=0, world=1, group=None, backend=None):
    """The BMA pass of ``evaluate_full_samples`` over one data set (methods/csghmc_fs.py:323-391) with the models dealt
    round-robin to the ranks.  ``nets``: {model index j: eval-mode network} for this rank's share (``shard_models``).
    Every batch crosses PCIe once per rank and is fed to all local models; per-model CE sums / error counts land in
    their own slots of an [S] vector.  One exchange step: an all-gather of the local logits ``[N, K, ceil(S/world)]``
    (543 KB per model at Pets size) and an all-reduce of the 2*S per-model statistics.  The average then runs over the
    full ``[N, K, S]`` stack in model order on every rank, so the result is **bit-identical** to the single-rank one
    (the fp32 running sum of csghmc_fs.py:349-351 is order-sensitive, which rules out reducing partial sums).
    -> dict(loss_per[S], err_per[S] (sums), bma_loss_sum, bma_err_sum, n, targets, logits [N,K], logits_all [N,K,S])."""
    backend = backend or CudaBmaBackend()
    S = int(n_models)
    mine = shard_models(S, rank, world)
    assert sorted(nets) == mine, f"rank {rank}: expected models {mine}, got {sorted(nets)}"
    s_max = (S + world - 1) // world
    loss_m = torch.zeros(S, dtype=torch.float64, device=device)      # per-model CE sums / error counts
    err_m = torch.zeros(S, dtype=torch.int32, device=device)
    ys, alls = [], []
    with torch.no_grad():
        for x, y in loader:
            x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
            outs = []
            for j in mine:
                out = nets[j](x).float().contiguous()
                backend.ce_err(out, y, loss_m[j:j + 1], err_m[j:j + 1])
                outs.append(out)
            if not outs:                                             # more ranks than models: shape from a forward-free stub
                outs_t = None
            else:
                outs_t = torch.stack(outs, 2)
            ys.append(y)
            alls.append(outs_t)
    targets = torch.cat(ys)
    N = targets.numel()
    if world > 1:
        import torch.distributed as dist
        kdim = torch.tensor([alls[0].shape[1] if alls[0] is not None else 0], dtype=torch.int64, device=device)
        dist.all_reduce(kdim, op=dist.ReduceOp.MAX, group=group)     # ranks without a model learn K
        K = int(kdim.item())
        local = torch.zeros(N, K, s_max, dtype=torch.float32, device=device)
        if mine:
            local[:, :, :len(mine)] = torch.cat(alls)
        gathered = torch.empty(world, N, K, s_max, dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(gathered.view(-1), local.contiguous().view(-1), group=group)   # flat: one shape rule for nccl and gloo
        # model j = r + world * i sits at gathered[r, :, :, i]
        la = gathered.permute(1, 2, 3, 0).reshape(N, K, s_max * world)[:, :, :S].contiguous()
        stats = torch.cat([loss_m, err_m.double()])
        dist.all_reduce(stats, group=group)
        loss_m, err_m = stats[:S], stats[S:]
    else:
        la = torch.cat(alls).contiguous()                             # [N,K,S]
    mean = torch.empty(la.shape[:2], dtype=torch.float32, device=device)
    backend.bma_mean(la, mean)
    bma_loss = torch.zeros(1, dtype=torch.float64, device=device)
    bma_err = torch.zeros(1, dtype=torch.int32, device=device)
    backend.ce_err(mean, targets, bma_loss, bma_err)                  # CE(mean logits) over the whole set (:365-366)
    host = torch.cat([loss_m.double(), err_m.double(), bma_loss, bma_err.double()]).cpu().numpy()   # one D2H
    return dict(loss_per=host[:S], err_per=host[S:2 * S], bma_loss_sum=host[2 * S], bma_err_sum=host[2 * S + 1], n=N,
                targets=targets.cpu().numpy(), logits=mean.cpu().numpy(), logits_all=la.cpu().numpy())
