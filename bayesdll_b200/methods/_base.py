"""Shared machinery of the drop-in ``Runner`` / ``Model`` pairs.

The reference has six near-identical 500-900 line files (methods/{sgld,sghmc,adam_sghmc,csgld,csghmc,
adam_csghmc}.py).  Here the per-method modules are thin: they pick a kernel variant, the optimizer momentum
and the capture scheme, and inherit everything else from the two runners below.

What is kept identical to the reference: constructor signatures, ``train`` / ``evaluate`` /
``train_one_epoch`` / ``save_logits`` / ``save_ckpt`` / ``load_ckpt`` names, arguments and return values, the
attributes other code reads (``net, net0, model, optimizer, criterion, Ninflate, nd, burnin, thin, nst,
post_theta_mom1/2/cnt, cycle_theta_mom1/2, samples_per_cycle, cycle_likelihoods, cycle_states,
model.momentum_buffer/m/v/t``), the checkpoint keys and the dense ``parameters_to_vector`` layout of every
stored vector, and the quirks listed in SURVEY.md Appendix B.

What is different: one fused kernel per step instead of ~12-30 eager kernels per tensor, no per-sample
``deepcopy(net)``, no host sync per batch in ``evaluate`` and none per step in ``train_one_epoch``.
"""
import contextlib
import copy
import os
import pickle
import time

import numpy as np
import torch
import torch.nn as nn
from tqdm import tqdm

from .. import _lib, calibration, ops
from .. import dist as bdist
from ..chain import ChainState, SampleRing
from ..flat import adopt_parameters, alloc_flat
from ..graphfwd import GraphedForward, capture_gc_guard
from ..writer import AsyncWriter, FlatBackedStateDict
from .cyclical import CyclicalSGMCMC


def reinitialize_fresh(net, logger):
    """Cold restart (methods/adam_csghmc.py:102-129, methods/csghmc_fs.py:91-117): Xavier for Linear, Kaiming (fan-in,
    ReLU) for Conv2d, unit / zero BatchNorm affine, ``reset_parameters`` otherwise -- written in place through the flat
    views, so the sampler state stays one buffer."""
    inits = ((nn.Linear, lambda m: nn.init.xavier_uniform_(m.weight)),
             (nn.Conv2d, lambda m: nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")),
             ((nn.BatchNorm2d, nn.BatchNorm1d), lambda m: m.weight is not None and nn.init.ones_(m.weight)))

    def fresh(m):
        for kinds, init_weight in inits:
            if isinstance(m, kinds):
                init_weight(m)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
                return
        if hasattr(m, "reset_parameters"):
            m.reset_parameters()
    net.apply(fresh)
    logger.info("Network parameters re-initialized with fresh random weights for cold restart.")


# ================================================================================================
# optimizer shim
# ================================================================================================
class FusedSGD(torch.optim.SGD):
    """Holds ``param_groups`` (cyclical runners write ``param_groups[i]['lr']`` every step, methods/csgld.py:234-239)
    and the ``momentum_buffer`` state for checkpoints.  The parameter update itself is applied by the fused kernel
    launched from ``Model.forward``, so ``step()`` does nothing: calling it after ``Model.forward`` -- as the
    reference's training loop does (methods/sghmc.py:229) -- is harmless."""

    def step(self, closure=None):  # noqa: D401
        return None if closure is None else closure()


def make_optimizer(net, lr, lr_head, momentum):
    """Body / head parameter groups exactly as methods/sghmc.py:53-57."""
    body = [p for n, p in net.named_parameters() if net.readout_name not in n]
    head = [p for n, p in net.named_parameters() if net.readout_name in n]
    return FusedSGD([{"params": body, "lr": lr}, {"params": head, "lr": lr_head}], momentum=momentum, weight_decay=0)


# ================================================================================================
# Model: fwd + bwd in PyTorch, then ONE fused update kernel
# ================================================================================================
class FusedModel(nn.Module):
    """Parameter-free module with the reference's ``Model`` interface (methods/sghmc.py:409-512)."""

    VARIANT = None

    def __init__(self, ND, prior_sig=1.0, bias="informative", momentum_decay=0.05, beta1=0.9, beta2=0.999,
                 epsilon=1e-8, temperature=1.0):
        super().__init__()
        self.ND = ND
        self.prior_sig = prior_sig
        self.bias = bias
        self.momentum_decay = momentum_decay
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.temperature = temperature
        self.t = 0
        self._chain = None
        self._opts = dict(sgd_momentum=0.0, seed=None, noise="philox", grad_mode="table", div_mode=_lib.DIV_RECIP,
                          optimizer=None, graph_train=False, release_grads=True)
        self._train_graphs = {}        # (net, shapes, criterion) -> eager-call count | captured graph | "eager"
        self._auto_graph = {}          # (id(net), id(criterion)) -> (weakref(net), eligible for graph_train=auto)

    def configure(self, **opts):
        unknown = set(opts) - set(self._opts)
        if unknown:
            raise TypeError(f"unknown option(s) {sorted(unknown)}")
        self._opts.update(opts)
        return self

    # ---- state exposed under the reference's attribute names (dict: param name -> tensor view) ------------
    def _views(self, which):
        if self._chain is None or getattr(self._chain, which) is None:
            raise AttributeError(which)
        return self._chain.named_views(getattr(self._chain, which))

    def _assign(self, which, mapping):
        views = self._views(which)
        with torch.no_grad():
            for k, t in mapping.items():
                views[k].copy_(t)

    momentum_buffer = property(lambda self: self._views("v"), lambda self, d: self._assign("v", d))

    @property
    def chain(self):
        return self._chain

    def _ensure_chain(self, net, net0):
        ch = self._chain
        if ch is not None and ch.params and ch.params[0] is next(iter(net.parameters())):
            ch.readopt_if_moved()                         # p.data re-pointed by user code (vector_to_parameters, ...)
            return ch
        o = self._opts
        seed = o["seed"] if o["seed"] is not None else torch.initial_seed()
        ch = ChainState(net, net0, variant=self.VARIANT, bias_mode=self.bias, mu=o["sgd_momentum"], noise=o["noise"],
                        seed=seed, grad_mode=o["grad_mode"], div_mode=o["div_mode"])
        self._chain = ch
        opt = o["optimizer"]
        if opt is not None and ch.buf is not None:       # expose the SGD momentum buffer through optimizer.state
            for p, view in zip(ch.params, ch.layout.views(ch.buf)):
                opt.state[p]["momentum_buffer"] = view
        return ch

    def _scalars(self, lrs, Ninflate, nd, should_sample):
        lr_body, lr_head = (lrs[0], lrs[0]) if len(lrs) == 1 else (lrs[0], lrs[1])
        return ops.make_scalars(self.VARIANT, lr_body=lr_body, lr_head=lr_head, ND=self.ND, Ninflate=Ninflate,
                                prior_sig=self.prior_sig, nd=nd, alpha=self.momentum_decay,
                                mu=self._opts["sgd_momentum"], beta1=self.beta1, beta2=self.beta2, eps=self.epsilon,
                                temperature=self.temperature, t=max(self.t, 1), add_noise=should_sample)

    def step_async(self, x, y, net, net0, criterion, lrs, Ninflate=1.0, nd=1.0, should_sample=False, capture=None, clip=None):
        """``forward`` without the host sync: returns (loss tensor, detached logits).  ``capture``: a callable returning
        an ops.make_capture spec, evaluated once the flat state exists; the sample capture that follows this step in
        the reference's loop then rides in the same kernel.  ``clip``: args.clip_grad -- the reference's
        ``clip_grad_norm_`` between this call and ``optimizer.step()`` (methods/csgld.py:250-251) folded into the update."""
        chain = self._ensure_chain(net, net0)
        if self.VARIANT in (_lib.ADAM_SGHMC, _lib.ADAM_CSGHMC):
            self.t += 1                                   # methods/adam_sghmc.py:494
        replayed = self._graphed_fwd_bwd(x, y, net, criterion) if self._graph_train_wanted(net, criterion) else None
        if replayed is not None:
            loss, out = replayed
        else:
            out = net(x)
            loss = criterion(out, y)
            net.zero_grad()                               # grads -> None; autograd hands us fresh tensors
            loss.backward()
        chain.update(self._scalars(lrs, Ninflate, nd, should_sample), capture=None if capture is None else capture(), clip=clip)
        if self._opts["release_grads"]:
            # The reference leaves the modified gradient in p.grad for the caller's optimizer.step(); here the update is
            # already applied, so a REAL torch optimizer stepping on p.grad would update theta a second time.  With the
            # gradients released every optimizer skips every parameter (``if p.grad is None: continue``); the kernel
            # launched above still owns their memory in stream order.
            for p in chain.params:
                p.grad = None
        return loss.detach(), out.detach()

    # ---- forward + backward as ONE CUDA-graph replay (hparams graph_train=auto|1|0) ----------------------------------
    # Module classes whose forward is known to have no host-side control flow on tensor VALUES: torch's own layers, the
    # torchvision model zoo (the reference's resnet101 / vit_l_32, networks/__init__.py:20-54), the reference's MLP
    # (networks/small_nets.py) and this package's copy of it.  ``graph_train=auto`` (the Runner default) captures only
    # networks AND criteria built from these alone; anything user-defined runs eagerly unless ``graph_train=1`` says so.
    _STATIC_MODULE_PREFIXES = ("torch.nn.", "torchvision.models.", "networks.small_nets", "bayesdll_b200.shapes")
    _MAX_TRAIN_GRAPHS = 4              # captured (network, batch shape) combinations per Model; further shapes run eagerly

    def _graph_train_wanted(self, net, criterion):
        want = self._opts["graph_train"]
        if want != "auto":
            return bool(want)
        key = (id(net), id(criterion))
        hit = self._auto_graph.get(key)
        if hit is None or hit[0]() is not net:            # the weak reference guards against a recycled id()
            import weakref
            ok = isinstance(criterion, nn.Module) and all(
                type(m).__module__.startswith(self._STATIC_MODULE_PREFIXES) for m in list(net.modules()) + list(criterion.modules()))
            hit = self._auto_graph[key] = (weakref.ref(net), ok)
        return hit[1]

    def _graphed_fwd_bwd(self, x, y, net, criterion):
        """The backbone's forward + loss + backward replayed as a CUDA graph (the fused sampler step stays a separate
        launch: its scalars and Philox counter change every step).  Same kernels as eager, so gradients, BatchNorm
        buffers and the loss are bit-identical; what disappears is the host's launch cost (ResNet-101, batch 16: ~1000
        launches per step), and the gradients live at fixed addresses, so the per-tensor run table stops changing.
        A capture freezes host-side control flow of ``net.forward`` / ``criterion``: ``graph_train=auto`` therefore captures
        only framework-provided modules (``_graph_train_wanted``), ``graph_train=1`` anything.  Per key the first two
        calls run eagerly (warm-up), the third captures; a failed capture falls back to eager with one warning.
        Returns (loss, out) clones with ``p.grad`` populated, or None when this call has to run eagerly."""
        if not (x.is_cuda and net.training):
            return None
        # the chain is part of the key: a rebuilt chain re-points every parameter at a new flat buffer
        key = (id(net), id(self._chain), tuple(x.shape), x.dtype, tuple(y.shape), y.dtype, id(criterion))
        ent = self._train_graphs.get(key, 0)
        if ent == "eager":
            return None
        params = self._chain.params
        if isinstance(ent, int):
            if ent < 2:
                self._train_graphs[key] = ent + 1
                return None
            if sum(isinstance(v, dict) for v in self._train_graphs.values()) >= self._MAX_TRAIN_GRAPHS:
                # a loader that keeps producing new batch shapes must not keep producing graphs (each owns a private pool
                # with a full set of activations): the first few shapes are replayed, the rest runs eagerly
                self._train_graphs[key] = "eager"
                return None
            try:
                net.zero_grad()                           # gradients must be allocated inside the graph's private pool
                sx, sy = x.clone(), y.clone()
                graph = torch.cuda.CUDAGraph()
                # eager steps create AccumulateGrad nodes on the caller's stream, the capture on torch's capture stream;
                # gradients are first assignments either way (no accumulate kernel), so torch's stream-mismatch
                # warning is moot for this module from here on
                quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
                if quiet is not None:
                    quiet(False)
                with capture_gc_guard(), torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = net(sx)
                    loss = criterion(out, sy)
                    loss.backward()
                ent = dict(graph=graph, x=sx, y=sy, out=out, loss=loss, grads=[p.grad for p in params], net=net,
                           criterion=criterion, chain=self._chain)   # the references keep the ids in the key unique
            except Exception as e:
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); running eagerly")
                torch.cuda.synchronize()
                net.zero_grad()
                self._train_graphs[key] = "eager"
                return None
            self._train_graphs[key] = ent
        else:
            ent["x"].copy_(x)
            ent["y"].copy_(y)
            for p, g in zip(params, ent["grads"]):        # after an eager step (e.g. a ragged last batch) in between
                if p.grad is not g:
                    p.grad = g
        ent["graph"].replay()
        return ent["loss"].detach().clone(), ent["out"].detach().clone()

    def forward(self, x, y, net, net0, criterion, lrs, Ninflate=1.0, nd=1.0, should_sample=False):
        """Same contract as the reference: returns ``(loss: float, out: detached [B,K])``; side effect: ``net`` holds the
        next sample (the SGD step the reference performs afterwards is already folded in)."""
        loss, out = self.step_async(x, y, net, net0, criterion, lrs, Ninflate, nd, should_sample)
        return loss.item(), out


class AdamStateMixin:
    """``Model.m`` / ``Model.v`` are the Adam first / second moment dicts (methods/adam_sghmc.py:486-491)."""
    m = property(lambda self: self._views("m"), lambda self, d: self._assign("m", d))
    v = property(lambda self: self._views("s"), lambda self, d: self._assign("s", d))


# ================================================================================================
# evaluation helpers shared by both runners
# ================================================================================================
class _EvalNet:
    """A copy of the live network whose parameters are views of one flat buffer, so a posterior sample is
    materialised by ONE kernel (bdl_draw) instead of ``deepcopy(net)`` + 5 eager kernels per tensor."""

    def __init__(self, net, layout, graph=True):
        self.net = copy.deepcopy(net)                     # inherits BatchNorm running stats (Appendix B.12)
        self.net.eval()
        self.flat = alloc_flat(layout.n_padded, next(net.parameters()).device)
        adopt_parameters(self.net, layout, self.flat)
        self.layout = layout
        self._xi = None
        # every pointer of the forward pass is stable (the draw overwrites ``flat`` in place): replay it as a CUDA graph
        self.forward = GraphedForward(self.net, enabled=graph)

    def refresh(self, net):
        """Re-inherit the live network's buffers (BatchNorm running statistics, Appendix B.12) IN PLACE: the captured
        graphs keep reading the same addresses.  False when the live network no longer matches this copy."""
        src, dst = list(net.buffers()), list(self.net.buffers())
        if len(src) != len(dst) or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(src, dst)):
            return False
        with torch.no_grad():
            if src:
                torch._foreach_copy_(dst, src)
        return True

    def load(self, flat_values):
        self.flat.copy_(flat_values)

    def draw(self, mean, second, var_mode, scale, noise_mode, seed, subseq, div_mode, center=None):
        if noise_mode == "philox":
            nz = ops.make_noise(seed=seed, subseq=subseq, stream_id=_lib.STREAM_DRAW)
        else:                                             # parity mode: torch.randn_like per tensor (sgld.py:294)
            if self._xi is None:
                self._xi = alloc_flat(self.layout.n_padded, self.flat.device)
                self._xi_views = self.layout.views(self._xi)
            for view in self._xi_views:
                view.copy_(torch.randn_like(view))
            nz = ops.make_noise(xi=self._xi)
        ops.draw(mean, second, self.flat, var_mode, scale, nz, div_mode, center=center)


class _Phases:
    """Device time per evaluation phase (draw | forward | reduce | exchange), measured with CUDA events on the launching
    stream when ``Runner.profile_eval`` is set (bench.py's ensemble leg); otherwise free."""

    def __init__(self, enabled):
        self.enabled = bool(enabled)
        self.spans = {}
        self.host = {}

    @contextlib.contextmanager
    def __call__(self, name):
        if not self.enabled:
            yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        try:
            yield
        finally:
            e1.record()
            self.spans.setdefault(name, []).append((e0, e1))
            # host time spent issuing the phase: a phase whose host time is close to its device time is host-bound
            self.host[f"{name}_host_ms"] = self.host.get(f"{name}_host_ms", 0.0) + (time.perf_counter() - t0) * 1e3

    def summary(self):
        """{phase: milliseconds}; synchronises."""
        if not self.enabled:
            return {}
        torch.cuda.synchronize()
        out = {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in self.spans.items()}
        out.update({f"{k}_launches": len(v) for k, v in self.spans.items()})
        out.update(self.host)
        return out


def _prefetch(loader, dev):
    """(x, y) batches on the device.  The NEXT batch's host-to-device copy runs on a side stream (copy engine) while the
    current batch computes: with 8 ranks sharing one PCIe root the 38 MB image batches would otherwise sit on the
    critical path of every rank."""
    if torch.device(dev).type != "cuda":
        for x, y in loader:
            yield x.to(dev), y.to(dev)
        return
    side = torch.cuda.Stream(dev)

    def issue(item):
        x, y = item
        if x.is_cuda and y.is_cuda:
            return x, y, None
        side.wait_stream(torch.cuda.current_stream(dev))      # the buffers being recycled by the allocator are idle
        with torch.cuda.stream(side):
            xd, yd = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(side)
        return xd, yd, evt

    it = iter(loader)
    try:
        nxt = issue(next(it))
    except StopIteration:
        return
    while nxt is not None:
        xd, yd, evt = nxt
        cur = torch.cuda.current_stream(dev)
        if evt is not None:
            cur.wait_event(evt)
            xd.record_stream(cur)
            yd.record_stream(cur)
        try:
            nxt = issue(next(it))
        except StopIteration:
            nxt = None
        yield xd, yd


class _EvalAccumulator:
    """Device-side accumulators of one ``evaluate()`` call: CE sum, error count and the per-batch outputs."""

    def __init__(self, dev):
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.err_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.nb, self.ys, self.lgs, self.lgalls = 0, [], [], []

    def add(self, y, logits, logits_all):
        self.ys.append(y)
        self.lgs.append(logits)
        self.lgalls.append(logits_all)
        self.nb += len(y)


def _pack_subseq(eval_id, batch, cycle, sample):
    """64-bit Philox sub-sequence of one posterior draw: independent of how samples are sharded over ranks."""
    return ((eval_id & 0xFFFF) << 48) | ((batch & 0xFFFFFF) << 24) | ((cycle & 0xFF) << 16) | (sample & 0xFFFF)


class _RunnerCommon:
    """Construction, evaluation plumbing, logits / checkpoint writers common to all six runners."""

    MODEL_CLS = None
    SGD_MOMENTUM_FROM_ARGS = False     # SGD(momentum=args.momentum) vs SGD(momentum=0)
    CONSULTS_CLIP_GRAD = False         # only the cyclical runners read args.clip_grad (methods/csgld.py:250)

    # ---- construction (methods/sghmc.py:18-67) ------------------------------------------------------------
    def __init__(self, net, net0, args, logger):
        self.args = args
        self.logger = logger
        # args.clip_grad: no driver of the reference defines it; the CYCLICAL runners consult it between Model.forward and
        # optimizer.step() (methods/csgld.py:250, adam_csghmc.py:319; csghmc.py:301 / csghmc_fs.py:509 clip a gradient that
        # is never used again, their Model.forward has already written p.data).  The burn-in runners never look at it.
        self.clip_grad = getattr(args, "clip_grad", None) if self.CONSULTS_CLIP_GRAD else None
        if args.pretrained is None:                       # zero prior mean
            self.net0 = copy.deepcopy(net)
            with torch.no_grad():
                for p in self.net0.parameters():
                    p.zero_()
        else:
            self.net0 = net0
        self.net0 = self.net0.to(args.device)
        self.net = net.to(args.device)

        hp = args.hparams
        self.model = self._build_model(hp).to(args.device)
        mu = float(args.momentum) if self.SGD_MOMENTUM_FROM_ARGS else 0.0
        self.optimizer = make_optimizer(self.net, args.lr, args.lr_head, mu)
        self.criterion = torch.nn.CrossEntropyLoss()

        self.Ninflate = float(hp["Ninflate"])
        self.nd = float(hp["nd"])
        self.nst = int(hp["nst"])
        self.thin = int(hp["thin"])

        # extra, optional knobs of the B200 path (absent from the reference's hparams -> defaults)
        self.noise_mode = str(hp.get("noise", "philox"))
        self.use_graph = bool(int(float(hp.get("graph", 1))))   # evaluation forwards replayed as CUDA graphs (graphfwd.py)
        self.div_mode = {"recip": _lib.DIV_RECIP, "ieee": _lib.DIV_IEEE}[str(hp.get("div", "recip"))]
        seed = int(hp["seed"]) if "seed" in hp else getattr(args, "seed", None)
        self.seed = int(seed) if seed is not None else torch.initial_seed()
        self.model.configure(sgd_momentum=mu, seed=self.seed, noise=self.noise_mode,
                             grad_mode=str(hp.get("grad", "table")), div_mode=self.div_mode, optimizer=self.optimizer,
                             graph_train=self._parse_graph_train(hp.get("graph_train", "auto")))
        self._eval_calls = 0
        # eval_shard=1: inside an initialised torch.distributed process group whose ranks hold the SAME chain state,
        # evaluate() / full_batch_likelihoods() deal the posterior samples round-robin to the ranks; results
        # are bit-identical to one rank and identical on every rank.  Off by default: independent chains (one per
        # rank) must not pool their samples.
        self.eval_shard = bool(int(float(hp.get("eval_shard", 0))))
        self.eval_cache = bool(int(float(hp.get("eval_cache", 1))))   # keep the evaluation copy + its graphs across calls
        self._eval_cached = None
        self.profile_eval = False        # True: evaluate() leaves a per-phase device-time split in self.eval_phases
        self.eval_phases = {}
        # fuse=1 (default): the moment capture that follows a sampler step runs inside the step kernel; fuse=0: separate
        # launch right after it, like the reference's statement order (bit-identical either way)
        self.fuse_capture = str(hp.get("fuse", "1")).lower() not in ("0", "false", "no")
        # checkpoint / sample writers: 'async' (default) overlaps D2H + serialisation with training, 'sync' behaves like
        # the reference's in-line torch.save (file complete when save_ckpt returns)
        self._writer = AsyncWriter(args.device, mode=str(hp.get("io", "async")))

    @staticmethod
    def _parse_graph_train(value):
        """hparams graph_train: 'auto' (default: CUDA-graph replay of forward + loss + backward for networks and criteria
        made of framework-provided modules only, eager otherwise), 1 (always try), 0 (never)."""
        if str(value).strip().lower() == "auto":
            return "auto"
        return bool(int(float(value)))

    def flush_io(self):
        """Block until every checkpoint / sample file submitted so far is on disk."""
        self._writer.flush()

    def _build_model(self, hp):
        raise NotImplementedError

    # ---- small helpers --------------------------------------------------------------------------------
    def _chain(self):
        return self.model._ensure_chain(self.net, self.net0)

    def _lrs(self):
        return [pg["lr"] for pg in self.optimizer.param_groups]

    def _dense(self, flat):
        return self._chain().layout.to_dense(flat)

    def _eval_net(self):
        """The evaluation copy of the live network (parameters = views of one flat buffer the draws overwrite).  Built
        once and kept across evaluate() / full_batch_likelihoods() calls -- the reference makes a fresh deepcopy(net) per
        sample per batch (methods/csgld.py:404-406); here even one copy per call would re-capture the forward's CUDA
        graphs every epoch -- and refreshed in place with the live network's buffers at every call, so each evaluation
        inherits the current BatchNorm statistics exactly like a fresh copy (Appendix B.12).  ``hparams eval_cache=0``
        restores one copy per call; ``release_eval_cache()`` frees the copy's memory."""
        layout = self._chain().layout
        ev = self._eval_cached
        if (ev is not None and self.eval_cache and ev.src is self.net and ev.layout is layout
                and ev.graph == self.use_graph and ev.refresh(self.net)):
            return ev
        ev = _EvalNet(self.net, layout, graph=self.use_graph)
        ev.src, ev.graph = self.net, self.use_graph
        self._eval_cached = ev if self.eval_cache else None
        return ev

    def release_eval_cache(self):
        self._eval_cached = None

    def _shard(self):
        """(rank, world) of the sample-sharded evaluation: the default process group when hparams ``eval_shard=1``,
        else (0, 1)."""
        if self.eval_shard:
            if self.noise_mode != "philox":
                raise ValueError("eval_shard=1 needs the counter-based Philox draws (noise=philox): a torch.randn_like "
                                 "stream cannot be dealt to ranks")
            return bdist.process_group()
        return 0, 1

    def _gather_eval(self, local_batches, rows, n_samples, rank, world, phases):
        """This rank's per-batch sample logits (lists of [B,K] tensors in ``my_samples`` order) -> the full [N,K,S]
        stack on every rank: the evaluation's one exchange step (an all-gather of N*K*ceil(S/world) floats per rank)."""
        import torch.distributed as dist
        dev = self.args.device
        n_mine = len(bdist.my_samples(n_samples, rank, world))
        k_local = local_batches[0][0].shape[1] if n_mine and local_batches else 0
        kk = torch.tensor([k_local, sum(rows)], dtype=torch.int64, device=dev)
        dist.all_reduce(kk, op=dist.ReduceOp.MAX)              # ranks without a sample learn K; all agree on N
        K, N = int(kk[0].item()), int(kk[1].item())
        if N != sum(rows):
            raise RuntimeError("eval_shard: the ranks iterated a different number of rows")
        s_max = (n_samples + world - 1) // world
        local = torch.zeros((N, K, s_max), dtype=torch.float32, device=dev)
        if n_mine:
            local[:, :, :n_mine] = torch.cat([torch.stack(outs, 2) for outs in local_batches]).float()
        with phases("exchange"):
            return bdist.gather_samples(local, n_samples, world)

    def _finish_eval(self, acc, phases=None):
        t0 = time.perf_counter()
        targets = torch.cat(acc.ys).cpu().numpy()
        logits = torch.cat(acc.lgs).cpu().numpy()
        logits_all = torch.cat(acc.lgalls).cpu().numpy()
        out = acc.loss_sum.item() / acc.nb, acc.err_cnt.item() / acc.nb, targets, logits, logits_all
        if phases is not None and phases.enabled:
            phases.host["d2h_and_sync_ms"] = (time.perf_counter() - t0) * 1e3
            self.eval_phases = phases.summary()
        return out

    # ---- writers (methods/sghmc.py:356-367) -----------------------------------------------------------
    def save_logits(self, targets, logits, logits_all, suffix=None):
        suffix = "" if suffix is None else f"_{suffix}"
        fname = os.path.join(self.args.log_dir, f"logits{suffix}.pkl")
        with open(fname, "wb") as ff:
            pickle.dump({"targets": targets, "logits": logits, "logits_all": logits_all}, ff,
                        protocol=pickle.HIGHEST_PROTOCOL)
        return fname

    def _state_dict_copy(self):
        """``copy.deepcopy(self.net.state_dict())`` (methods/csgld.py:311) as ONE TMA copy of the flat parameter buffer
        plus clones of the few buffers (BatchNorm statistics); entries are views of that snapshot."""
        ch = self._chain()
        return FlatBackedStateDict.snapshot(self.net, ch.layout, ch.names, ch.snapshot())

    def _dense_for_save(self, flat, mutable):
        """Dense ``parameters_to_vector``-ordered vector for a writer: a view when the padded layout only differs from
        the dense one by tail padding and the buffer will not change any more, else a copy."""
        L = self._chain().layout
        if not mutable and all(sg.begin == sg.dense_begin for sg in L.segments):
            return flat[:L.n_dense]
        return L.to_dense(flat)

    def _calibrate_and_log(self, targets_test, logits_test, val, fmt_topt):
        args, logger = self.args, self.logger
        ece, mce, nll = calibration.analyze(targets_test, logits_test, num_bins=args.ece_num_bins,
                                            plot_save_path=os.path.join(args.log_dir, "reliability_T1.png"),
                                            temperature=1)
        logger.info("[Calibration - Default T=1] ECE = %.4f, MCE = %.4f, NLL = %.4f" % (ece, mce, nll))
        if val is not None:
            targets_val, logits_val = val
            Topt, ok = calibration.find_optimal_temperature(
                targets_val, logits_val, plot_save_path=os.path.join(args.log_dir, "temp_scale_optim_curve.png"))
            if ok:
                ece, mce, nll = calibration.analyze(targets_test, logits_test, num_bins=args.ece_num_bins,
                                                    plot_save_path=os.path.join(args.log_dir, "reliability_Topt.png"),
                                                    temperature=Topt)
                logger.info("[Calibration - Temp-scaled Topt=%.4f] ECE = %.4f, MCE = %.4f, NLL = %.4f" %
                            (fmt_topt(Topt), ece, mce, nll))
            else:
                logger.info("!! Temperature scaling optimization failed !!")

    def _eval_and_report(self, ep, val_loader, test_loader, losses_val, errors_val, losses_test, errors_test):
        """Validation + test evaluation of one epoch with the reference's log lines; returns what train() needs."""
        logger = self.logger
        val = None
        if val_loader is not None:
            tic = time.time()
            losses_val[ep], errors_val[ep], t_val, l_val, la_val = self.evaluate(val_loader)
            logger.info(f"(Epoch {ep}) Validation summary: loss = {losses_val[ep]:.4f}, "
                        f"prediction error = {errors_val[ep]:.4f} (time: {time.time() - tic:.4f} seconds)")
            val = (t_val, l_val, la_val)
        tic = time.time()
        losses_test[ep], errors_test[ep], t_test, l_test, la_test = self.evaluate(test_loader)
        logger.info(f"(Epoch {ep}) Test summary: loss = {losses_test[ep]:.4f}, "
                    f"prediction error = {errors_test[ep]:.4f} (time: {time.time() - tic:.4f} seconds)")
        return val, (t_test, l_test, la_test)

    def _on_best(self, ep, loss_now, val, test, fmt_topt, save_ckpt):
        logger = self.logger
        logger.info(f"Best evaluation loss so far! @epoch {ep}: loss = {loss_now}")
        if val is not None:
            logger.info(f"Logits on val set saved at {self.save_logits(*val, suffix='val')}")
        logger.info(f"Logits on test set saved at {self.save_logits(*test, suffix='test')}")
        if save_ckpt:
            logger.info(f"Checkpoint saved at {self.save_ckpt(ep, wait=False)}")
        self._calibrate_and_log(test[0], test[1], None if val is None else val[:2], fmt_topt)


# ================================================================================================
# burn-in / thinning runners: sgld, sghmc, adam_sghmc      (methods/sghmc.py:16-406)
# ================================================================================================
class BurninRunner(_RunnerCommon):

    def __init__(self, net, net0, args, logger):
        super().__init__(net, net0, args, logger)
        self.burnin = int(args.hparams["burnin"])
        self._mom1 = self._mom2 = None

    # dense views under the reference's attribute names (checkpoint contract: parameters_to_vector order)
    @property
    def post_theta_mom1(self):
        if self._mom1 is None:
            raise AttributeError("post_theta_mom1")
        return self._dense(self._mom1)

    @post_theta_mom1.setter
    def post_theta_mom1(self, dense):
        ch = self._chain()
        self._mom1 = ch.layout.from_dense(dense.to(ch.device, torch.float32), self._mom1)

    @property
    def post_theta_mom2(self):
        if self._mom2 is None:
            raise AttributeError("post_theta_mom2")
        return self._dense(self._mom2)

    @post_theta_mom2.setter
    def post_theta_mom2(self, dense):
        ch = self._chain()
        self._mom2 = ch.layout.from_dense(dense.to(ch.device, torch.float32), self._mom2)

    # ---- training loop (methods/sghmc.py:70-193) -----------------------------------------------------------
    def train(self, train_loader, val_loader, test_loader):
        args, logger = self.args, self.logger
        logger.info("Start training...")
        losses_train, errors_train = np.zeros(args.epochs), np.zeros(args.epochs)
        losses_val = errors_val = None
        if val_loader is not None:
            losses_val, errors_val = np.zeros(args.epochs), np.zeros(args.epochs)
        losses_test, errors_test = np.zeros(args.epochs), np.zeros(args.epochs)
        best_loss = np.inf
        tic0 = time.time()
        bi = 0
        for ep in range(args.epochs):
            if ep == self.burnin:
                logger.info("(leaving burnin period) start collecting posterior samples")
                self._start_collecting()
            tic = time.time()
            losses_train[ep], errors_train[ep], bi = self.train_one_epoch(train_loader, collect=(ep >= self.burnin), bi=bi)
            logger.info("[Epoch %d/%d] Training summary: loss = %.4f, prediction error = %.4f (time: %.4f seconds)" %
                        (ep, args.epochs, losses_train[ep], errors_train[ep], time.time() - tic))
            if ep % args.test_eval_freq == 0 and ep >= self.burnin:
                val, test = self._eval_and_report(ep, val_loader, test_loader, losses_val, errors_val, losses_test,
                                                  errors_test)
                loss_now = losses_val[ep] if val_loader is not None else losses_test[ep]
                if loss_now < best_loss:
                    best_loss = loss_now
                    self._on_best(ep, loss_now, val, test, fmt_topt=lambda T: float(np.asarray(T).reshape(-1)[0]), save_ckpt=True)
        self.flush_io()
        toc0 = time.time()
        logger.info("Training done! Total time = %f (average per epoch = %f) seconds" %
                    (toc0 - tic0, (toc0 - tic0) / args.epochs))

    def _start_collecting(self):
        """mom1 = theta*1.0 ; mom2 = theta**2 ; cnt = 1   (methods/sghmc.py:96-103)."""
        ch = self._chain()
        n = ch.layout.n_padded
        self._mom1 = alloc_flat(n, ch.device, zero=False) if self._mom1 is None else self._mom1
        if self.nst > 0 and self._mom2 is None:
            self._mom2 = alloc_flat(n, ch.device, zero=False)
        ops.moments_avg(ch.theta, self._mom1, self._mom2 if self.nst > 0 else None, 0, init=True)
        self.post_theta_cnt = 1

    def train_one_epoch(self, train_loader, collect, bi):
        """One epoch of sampler steps; running moments every ``thin`` steps after burn-in (methods/sghmc.py:196-253)."""
        args, logger = self.args, self.logger
        self.net.train()
        dev = args.device
        loss_acc = torch.zeros((), dtype=torch.float64, device=dev)
        err_acc = torch.zeros((), dtype=torch.int64, device=dev)
        nb_samples = 0
        with tqdm(train_loader, unit="batch") as tepoch:
            for x, y in tepoch:
                x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
                captures = collect and (bi + 1) % self.thin == 0
                spec = None
                if captures and self.fuse_capture:
                    spec = lambda: ops.make_capture("avg", self._mom1, self._mom2 if self.nst > 0 else None,
                                                    self.post_theta_cnt)
                loss_, out = self.model.step_async(x, y, self.net, self.net0, self.criterion, self._lrs(),
                                                   self.Ninflate, self.nd, capture=spec)
                self.optimizer.step()                     # no-op: the update is part of the fused kernel
                loss_acc += loss_.double() * len(y)
                err_acc += out.argmax(dim=1).ne(y).sum()
                nb_samples += len(y)
                bi += 1
                if captures:
                    logger.info("(post-burnin) accumulate posterior samples")
                    if spec is None:
                        ch = self._chain()
                        ops.moments_avg(ch.theta, self._mom1, self._mom2 if self.nst > 0 else None, self.post_theta_cnt,
                                        div_mode=self.div_mode)
                    self.post_theta_cnt += 1
        loss, error = loss_acc.item(), err_acc.item()      # the only host sync of the epoch
        return loss / nb_samples, error / nb_samples, bi

    # ---- posterior-predictive ensemble (methods/sghmc.py:256-324) -------------------------------------------
    def _variance_ratio(self):
        return self.post_theta_cnt / (self.post_theta_cnt - 1) if self.post_theta_cnt > 1 else 1.0

    def evaluate(self, test_loader):
        """(loss, err, targets[N], logits[N,K], logits_all[N,K,S]) as methods/sghmc.py:256-324.  With ``eval_shard=1`` in
        a process group the ``nst`` draws are dealt round-robin to the ranks, one all-gather assembles ``logits_all`` and
        the ensemble / CE reductions then run over the full stack on every rank: same bits as one rank."""
        args = self.args
        dev = args.device
        ch = self._chain()
        ev = self._eval_net()
        self._eval_calls += 1
        ratio = self._variance_ratio()
        rank, world = self._shard()
        S = max(1, self.nst)
        mine = bdist.my_samples(S, rank, world)
        ph = _Phases(self.profile_eval)
        acc = _EvalAccumulator(dev)
        local, rows, ys = [], [], []
        with torch.no_grad(), tqdm(test_loader, unit="batch") as tepoch:
            for b_idx, (x, y) in enumerate(_prefetch(tepoch, dev)):
                outs = []
                for ii in mine:
                    with ph("draw"):
                        if self.nst == 0:
                            ev.load(self._mom1)
                        else:                              # fresh draw for every batch (Appendix B.6)
                            ev.draw(self._mom1, self._mom2, ops.VAR_FROM_MOMENTS, ratio, self.noise_mode, self.seed,
                                    _pack_subseq(self._eval_calls, b_idx, 0, ii), self.div_mode)
                    with ph("forward"):
                        outs.append(ev.forward(x))
                if world == 1:
                    with ph("reduce"):
                        self._reduce_batch(torch.stack(outs, 2).contiguous().float(), y, acc)
                else:
                    local.append(outs)
                    rows.append(len(y))
                    ys.append(y)
        if world > 1:
            bdist.agree_across_ranks(torch.cat(ys), "the evaluation targets")
            la = self._gather_eval(local, rows, S, rank, world, ph)
            with ph("reduce"):
                # the ensemble kernel is row-wise (one warp per row), so ONE launch over all N rows gives the bits the
                # per-batch launches of one rank give; the CE accumulation keeps the batch boundaries (fp64 summation order)
                logits_full = torch.empty(la.shape[:2], dtype=torch.float32, device=dev)
                ops.ensemble(la, logits_full, self.nst)
                for la_b, lg_b, y in zip(torch.split(la, rows), torch.split(logits_full, rows), ys):
                    ops.ce_err(lg_b, y, acc.loss_sum, acc.err_cnt)
                    acc.add(y, lg_b, la_b)
        return self._finish_eval(acc, ph)

    def _reduce_batch(self, logits_all_, y, acc):
        """log-mean-softmax over the samples, CE sum and error count of one batch (methods/sghmc.py:299-306)."""
        logits_ = torch.empty(logits_all_.shape[:2], dtype=torch.float32, device=logits_all_.device)
        ops.ensemble(logits_all_, logits_, self.nst)
        ops.ce_err(logits_, y, acc.loss_sum, acc.err_cnt)
        acc.add(y, logits_, logits_all_)

    def get_mean_vars_from_moments(self):
        """API compatibility (methods/sghmc.py:327-353): two ``net``-like modules holding mean and variance."""
        ch = self._chain()
        with torch.no_grad():
            mean = copy.deepcopy(self.net)
            nn.utils.vector_to_parameters(self.post_theta_mom1.clone(), mean.parameters())
            var = None
            if self.nst > 0:
                vec = self._variance_ratio() * (self.post_theta_mom2 - self.post_theta_mom1 ** 2)
                vec.clamp_(min=1e-12)
                var = copy.deepcopy(self.net)
                nn.utils.vector_to_parameters(vec, var.parameters())
        return mean, var

    # ---- checkpoints (methods/sghmc.py:370-406) --------------------------------------------------------------
    CKPT_HAS_LAST_THETA = True

    def _ckpt_extra(self):
        return {}

    def save_ckpt(self, epoch, wait=True):
        """Same file, keys and layout as methods/sghmc.py:370-388.  Called directly it returns once the file is complete,
        like the reference; the training loop passes ``wait=False`` and lets the writer thread finish it."""
        fname = os.path.join(self.args.log_dir, "ckpt.pt")
        ck = {}
        if self.CKPT_HAS_LAST_THETA:
            ck["last_theta"] = self._state_dict_copy()
        ck.update({
            "post_theta_mom1": self.post_theta_mom1,       # the property hands out a fresh dense copy (a snapshot)
            "post_theta_mom2": self.post_theta_mom2 if self.nst > 0 else None,
            "post_theta_cnt": self.post_theta_cnt,
            "prior_sig": self.model.prior_sig,
            "optimizer": self._optimizer_state_snapshot(),
        })
        ck.update(self._ckpt_extra())
        ck["epoch"] = epoch
        self._writer.submit(fname, ck)
        if wait:
            self._writer.flush()
        return fname

    def _optimizer_state_snapshot(self):
        """``optimizer.state_dict()`` whose momentum buffers are snapshots (the live ones are views the kernel updates)."""
        sd = self.optimizer.state_dict()
        ch = self._chain()
        if ch.buf is None:
            return sd
        snap = ch.layout.views(ch.snapshot(ch.buf))
        index = {id(p): i for i, p in enumerate(ch.params)}
        order = [index[id(p)] for g in self.optimizer.param_groups for p in g["params"]]
        state = {}
        for k, st in sd["state"].items():
            st = dict(st)
            if st.get("momentum_buffer") is not None:
                st["momentum_buffer"] = snap[order[k]]
            state[k] = st
        return {"state": state, "param_groups": sd["param_groups"]}

    def load_ckpt(self, ckpt_path):
        """Mirrors the reference incl. its quirk: theta is not restored and the sample count is overwritten with the
        epoch number (Appendix B.11)."""
        self.flush_io()
        ckpt = torch.load(ckpt_path, map_location=self.args.device, weights_only=False)
        self.post_theta_mom1 = ckpt["post_theta_mom1"]
        if ckpt["post_theta_mom2"] is not None:
            self.post_theta_mom2 = ckpt["post_theta_mom2"]
        self.post_theta_cnt = ckpt["epoch"]
        self.model.prior_sig = ckpt["prior_sig"]
        self._load_ckpt_extra(ckpt)
        self.optimizer.load_state_dict(ckpt["optimizer"])
        ch = self._chain()
        if ch.buf is not None:                           # re-attach the flat SGD momentum buffer
            for p, view in zip(ch.params, ch.layout.views(ch.buf)):
                st = self.optimizer.state[p]
                if "momentum_buffer" in st and st["momentum_buffer"] is not None:
                    view.copy_(st["momentum_buffer"])
                    ch.sgd_steps = max(ch.sgd_steps, 1)
                st["momentum_buffer"] = view
        return ckpt["epoch"]

    def _load_ckpt_extra(self, ckpt):
        pass


# ================================================================================================
# cyclical runners: csgld, csghmc, adam_csghmc       (methods/csgld.py:17-594, csghmc.py, adam_csghmc.py)
# ================================================================================================
class CyclicalRunner(_RunnerCommon):
    CONSULTS_CLIP_GRAD = True        # csgld.py:250, adam_csghmc.py:319 (csghmc.py:301: no effect on theta, see ChainState.update)
    CAPTURE = "avg"                  # 'avg' (csgld.py:282-290, adam_csghmc.py:349-357) | 'welford' (csghmc.py:333-345)
    LIKELIHOOD_MEAN = "theta"        # 'theta' (csgld.py:518) | 'cycle_mean' (csghmc.py:578, adam_csghmc.py:639)
    LAST_THETA_AS_VECTOR = False     # csghmc / adam_csghmc store a flat vector (csghmc.py:534)
    PASS_SHOULD_SAMPLE = False       # only csghmc's Model takes should_sample (csghmc.py:296-300)
    RESET_ADAM_AT_CYCLE_END = False  # adam_csghmc.py:370-378, 405
    STORE_ALL_SAMPLES = False        # csgld.py:278-279 (args.full_sample)
    TITLE = "Cyclical SGLD"

    def __init__(self, net, net0, args, logger):
        super().__init__(net, net0, args, logger)
        self.cyclical_scheduler = CyclicalSGMCMC(
            base_lr=args.lr,
            nbr_of_cycles=args.num_cycles if hasattr(args, "num_cycles") else 10,
            epochs=args.epochs,
            proportion_exploration=args.proportion_exploration if hasattr(args, "proportion_exploration") else 0.5)
        if "burnin" in args.hparams:
            self.burnin = int(args.hparams["burnin"])
        self.samples_collected = 0
        self.current_cycle = 0
        self.samples_per_cycle = {}
        self._cyc1, self._cyc2 = {}, {}      # cycle -> padded flat moment buffers
        self.cycle_likelihoods = {}
        self.cycle_states = {}
        self.all_samples = {}                # key "<epoch>_<batch>" -> dense view of a slot of the HBM sample ring
        self._ring = None
        self._batches_per_epoch = None

    # dict views under the reference's names: cycle -> dense vector
    @property
    def cycle_theta_mom1(self):
        return {c: self._dense(t) for c, t in self._cyc1.items()}

    @cycle_theta_mom1.setter
    def cycle_theta_mom1(self, d):
        ch = self._chain()
        self._cyc1 = {c: ch.layout.from_dense(t.to(ch.device, torch.float32)) for c, t in d.items()}

    @property
    def cycle_theta_mom2(self):
        return {c: self._dense(t) for c, t in self._cyc2.items()}

    @cycle_theta_mom2.setter
    def cycle_theta_mom2(self, d):
        ch = self._chain()
        self._cyc2 = {c: ch.layout.from_dense(t.to(ch.device, torch.float32)) for c, t in d.items()}

    # ---- training loop (methods/csgld.py:81-193) ---------------------------------------------------------------
    def train(self, train_loader, val_loader, test_loader):
        args, logger = self.args, self.logger
        logger.info(f"Start training with {self.TITLE}...")
        losses_train, errors_train = np.zeros(args.epochs), np.zeros(args.epochs)
        losses_val = errors_val = None
        if val_loader is not None:
            losses_val, errors_val = np.zeros(args.epochs), np.zeros(args.epochs)
        losses_test, errors_test = np.zeros(args.epochs), np.zeros(args.epochs)
        best_loss = np.inf
        tic0 = time.time()
        for ep in range(args.epochs):
            self.cyclical_scheduler.current_epoch = ep
            tic = time.time()
            losses_train[ep], errors_train[ep], cycle_updated = self.train_one_epoch(train_loader)
            logger.info(f"[Epoch {ep}/{args.epochs}] Training summary: loss = {losses_train[ep]:.4f}, "
                        f"prediction error = {errors_train[ep]:.4f} (time: {time.time() - tic:.4f} seconds)")
            self._after_epoch(ep, val_loader)
            if cycle_updated:
                self._before_cycle_eval(val_loader)
                val, test = self._eval_and_report(ep, val_loader, test_loader, losses_val, errors_val, losses_test,
                                                  errors_test)
                loss_now = losses_val[ep] if val_loader is not None else losses_test[ep]
                if loss_now < best_loss:
                    best_loss = loss_now
                    self._on_best(ep, loss_now, val, test, fmt_topt=lambda T: T[0], save_ckpt=False)
        self.flush_io()
        toc0 = time.time()
        logger.info(f"Training done! Total time = {toc0 - tic0:.4f} (average per epoch = "
                    f"{(toc0 - tic0) / args.epochs:.4f}) seconds")
        logger.info(f"Total samples collected: {self.samples_collected} across {self.current_cycle} cycles")
        return {"losses_train": losses_train, "errors_train": errors_train,
                "losses_val": losses_val if val_loader is not None else None,
                "errors_val": errors_val if val_loader is not None else None,
                "losses_test": losses_test, "errors_test": errors_test, "samples_per_cycle": self.samples_per_cycle}

    def _after_epoch(self, ep, val_loader):
        pass

    def _before_cycle_eval(self, val_loader):
        pass

    def _point_estimate(self, loader, net, fwd=None):
        """Deterministic pass of ``net`` over ``loader``: (mean CE, error rate).  ``fwd``: optional callable computing
        ``net(x)`` (the evaluation net's CUDA-graph replay)."""
        fwd = fwd or net
        dev = self.args.device
        was_training = net.training
        net.eval()
        loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        err_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        nb = 0
        with torch.no_grad():
            for x, y in _prefetch(loader, dev):
                ops.ce_err(fwd(x).float().contiguous(), y, loss_sum, err_cnt)
                nb += len(y)
        net.train(was_training)
        if nb == 0:
            return 0.0, 0.0
        return loss_sum.item() / nb, err_cnt.item() / nb

    def train_one_epoch(self, train_loader):
        """One epoch with the cyclical step size; per-cycle capture; end-of-cycle bookkeeping (methods/csgld.py:195-331)."""
        args, logger = self.args, self.logger
        sched = self.cyclical_scheduler
        self.net.train()
        dev = args.device
        loss_acc = torch.zeros((), dtype=torch.float64, device=dev)
        err_acc = torch.zeros((), dtype=torch.int64, device=dev)
        nb_samples = 0
        cycle_updated = False
        B = self._batches_per_epoch = len(train_loader)
        with tqdm(train_loader, unit="batch") as tepoch:
            for batch_idx, (x, y) in enumerate(tepoch):
                pos = dict(epoch=sched.current_epoch, batch=batch_idx, batches_per_epoch=B)
                current_lr = sched.calculate_lr(**pos)
                should_sample = sched.should_sample(**pos) and batch_idx % self.thin == 0
                last_in_cycle = sched.last_in_cycle(**pos)
                for i, group in enumerate(self.optimizer.param_groups):
                    group["lr"] = current_lr * (args.lr_head / args.lr) if i == 1 else current_lr

                x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
                kw = dict(should_sample=should_sample) if self.PASS_SHOULD_SAMPLE else {}
                fused = should_sample and self.fuse_capture and self.clip_grad is None
                if self.clip_grad is not None:
                    kw["clip"] = self.clip_grad           # torch.nn.utils.clip_grad_norm_(self.net.parameters(), args.clip_grad)
                if fused:
                    cyc = sched.get_cycle_number(**pos)
                    kw["capture"] = lambda: self._capture_spec(cyc)
                loss_, out = self.model.step_async(x, y, self.net, self.net0, self.criterion, self._lrs(),
                                                   self.Ninflate, self.nd, **kw)
                self.optimizer.step()                     # no-op (see FusedSGD)
                loss_acc += loss_.double() * len(y)
                err_acc += out.argmax(dim=1).ne(y).sum()
                nb_samples += len(y)

                if should_sample:
                    cycle_number = sched.get_cycle_number(**pos)
                    if batch_idx % 50 == 0:
                        logger.info(f"Sampling phase: collecting posterior sample at lr={current_lr:.6f}")
                    self._capture(cycle_number, sched.current_epoch, batch_idx, already_fused=fused)
                    self.samples_collected += 1
                    self.samples_per_cycle[cycle_number] = self.samples_per_cycle.get(cycle_number, 0) + 1
                elif batch_idx % 50 == 0:
                    logger.info(f"Exploration phase: lr={current_lr:.6f}")

                if last_in_cycle:
                    cycle_number = sched.get_cycle_number(**pos)
                    self.cycle_states[cycle_number] = self._state_dict_copy()
                    if self.RESET_ADAM_AT_CYCLE_END:
                        logger.info(f"Resetting momentum states for new cycle {cycle_number}")
                        self._reset_optimizer_states(log=False)
                    if cycle_number > self.current_cycle:
                        cycle_updated = True
                        self.current_cycle = cycle_number
                        logger.info(f"Completed cycle {cycle_number}")
                        likelihood = np.array(self.full_batch_likelihoods(train_loader))
                        self.cycle_likelihoods[cycle_number] = likelihood
                        logger.info(f"Cycle {cycle_number} full batch likelihood: {likelihood.mean():.6e}")
                        self.save_ckpt(epoch=sched.current_epoch, wait=False)
                        if self.STORE_ALL_SAMPLES and getattr(args, "full_sample", False):
                            self._writer.submit("all_samples_TEST.ckpt", dict(self.all_samples))   # csgld.py:328-329
                        self._after_cycle_completed(cycle_number)
        loss, error = loss_acc.item(), err_acc.item()
        return loss / nb_samples, error / nb_samples, cycle_updated

    def _after_cycle_completed(self, cycle_number):
        pass

    def _reset_optimizer_states(self, log=True):
        pass

    # ---- per-cycle capture ------------------------------------------------------------------------------------
    def _capture_spec(self, cycle):
        """ops.make_capture spec for folding the sample of this step into cycle ``cycle``'s moments; allocates the buffers
        on the first sample of a cycle.  Called exactly once per captured sample -- by the fused step, or by ``_capture``
        when the capture runs as its own launch; ``_capture`` then does the bookkeeping."""
        ch = self._chain()
        n = ch.layout.n_padded
        first = self._capture_first = cycle not in self._cyc1
        if first:
            self._cyc1[cycle] = alloc_flat(n, ch.device, zero=False)
            self._cyc2[cycle] = alloc_flat(n, ch.device, zero=False)
        if self.CAPTURE == "avg":
            cnt = 0 if first else self.samples_per_cycle.get(cycle, 0)          # cycle_count - 1 (csgld.py:286-290)
            return ops.make_capture("avg", self._cyc1[cycle], self._cyc2[cycle], cnt, init=first)
        n_w = 1 if first else self.samples_per_cycle.get(cycle, 0) + 1          # Welford's n (csghmc.py:339)
        return ops.make_capture("welford", self._cyc1[cycle], self._cyc2[cycle], n_w, init=first)

    def _capture(self, cycle, epoch, batch_idx, already_fused=False):
        """Per-cycle capture (csgld.py:276-293, csghmc.py:327-348).  ``already_fused``: the moment update was part of
        the step kernel; only the bookkeeping (and the optional raw-sample store) remains."""
        ch = self._chain()
        if not already_fused:
            spec = self._capture_spec(cycle)
            launch = ops.moments_avg if self.CAPTURE == "avg" else ops.moments_welford
            launch(ch.theta, self._cyc1[cycle], self._cyc2[cycle], int(spec.cnt), init=bool(spec.init), div_mode=self.div_mode)
        first = self._capture_first
        if self.CAPTURE == "avg":
            if self.STORE_ALL_SAMPLES and getattr(self.args, "full_sample", False):            # csgld.py:278-279
                if self._ring is None:
                    want = self._expected_samples()
                    self._ring = SampleRing(ch.layout, ch.device, want)
                    if self._ring.capacity < want:
                        self.logger.warning(
                            f"full_sample: the HBM sample ring holds {self._ring.capacity} of the {want} samples this run "
                            f"will capture (25 % of free HBM); older samples drop out of all_samples as the ring wraps "
                            f"(the reference keeps every sample in device memory until it runs out)")
                # a slot about to be reused may still be queued in the writer (all_samples_TEST.ckpt holds views)
                self._ring.capture(ch.theta, f"{epoch}_{batch_idx}", self.all_samples,
                                   before_overwrite=lambda _old: self._writer.flush())
        else:   # Welford with the reference's double-counted n (Appendix B.3): n runs 3, 5, 7, ...
            self.samples_per_cycle[cycle] = 1 if first else self.samples_per_cycle.get(cycle, 0) + 1

    def _expected_samples(self):
        """Number of captures the schedule will make over the whole run (host arithmetic, sizes the ring)."""
        sched, B = self.cyclical_scheduler, self._batches_per_epoch or 1
        return max(1, sum(1 for e in range(self.args.epochs) for b in range(B)
                          if b % self.thin == 0 and sched.should_sample(epoch=e, batch=b, batches_per_epoch=B)))

    def _cycle_variance_spec(self, cycle):
        """(second buffer, var_mode, scale) for bdl_draw; keeps the reference's evaluation order so that a cycle with
        exactly one sample raises ZeroDivisionError in the 'avg' scheme (Appendix B.7)."""
        if self.CAPTURE == "avg":
            cnt = self.samples_per_cycle.get(cycle, 0)
            ratio = cnt / (cnt - 1)
            return self._cyc2[cycle], ops.VAR_FROM_MOMENTS, (ratio if cnt > 1 else 1.0)
        cnt = self.samples_per_cycle.get(cycle, 0)
        if cnt > 1:
            return self._cyc2[cycle], ops.VAR_FROM_WELFORD, float(cnt - 1)
        return None, ops.VAR_TINY, 1.0

    # ---- GMM ensemble (methods/csgld.py:333-456) -----------------------------------------------------------------
    def evaluate(self, test_loader):
        """(loss, err, targets[N], logits[N,K], logits_all[N,K,nst,C]) as methods/csgld.py:333-456.  With ``eval_shard=1``
        in a process group the C x nst draws are dealt round-robin to the ranks (flat index cycle * nst + sample), one
        all-gather assembles ``logits_all`` and the per-cycle log-mean-softmax, the log-space GMM mixture and the CE
        reductions then run over the full stack on every rank: same bits as one rank."""
        args = self.args
        dev = args.device
        ch = self._chain()
        gmm_weights = self.calculate_gmm_weights()
        self.logger.info(f"GMM component weights: {gmm_weights}")
        ev = self._eval_net()
        self._eval_calls += 1
        cycles = [c for c in self._cyc1 if not gmm_weights.get(c, 0.0) < 1e-10]
        specs = {c: self._cycle_variance_spec(c) for c in cycles} if self.nst > 0 else {}
        weights = [gmm_weights.get(c, 0.0) for c in cycles]
        per = max(1, self.nst)
        S = len(cycles) * per
        rank, world = self._shard()
        mine = bdist.my_samples(S, rank, world)
        ph = _Phases(self.profile_eval)
        acc = _EvalAccumulator(dev)
        local, rows, ys = [], [], []
        with torch.no_grad(), tqdm(test_loader, unit="batch") as tepoch:
            for b_idx, (x, y) in enumerate(_prefetch(tepoch, dev)):
                outs = []
                for j in mine:
                    c, ii = cycles[j // per], j % per
                    with ph("draw"):
                        if self.nst == 0:
                            ev.load(self._cyc1[c])
                        else:
                            second, var_mode, scale = specs[c]
                            ev.draw(self._cyc1[c], second, var_mode, scale, self.noise_mode, self.seed,
                                    _pack_subseq(self._eval_calls, b_idx, c, ii), self.div_mode)
                    with ph("forward"):
                        outs.append(ev.forward(x))
                if world == 1:
                    with ph("reduce"):
                        flat = torch.stack(outs, 2).float() if outs else torch.zeros((x.size(0), args.num_classes, 0), device=dev)
                        self._reduce_batch(flat, y, acc, weights, per)
                else:
                    local.append(outs)
                    rows.append(len(y))
                    ys.append(y)
        if world > 1:
            bdist.agree_across_ranks(torch.cat(ys), "the evaluation targets")
            la = self._gather_eval(local, rows, S, rank, world, ph)
            with ph("reduce"):
                # row-wise kernels: reduce all N rows at once (2 launches per cycle instead of 2 per cycle per batch); only
                # the CE accumulation keeps the batch boundaries, so every number equals the single-rank one bit for bit
                wide = _EvalAccumulator(dev)
                self._reduce_batch(la, torch.cat(ys), wide, weights, per, with_ce=False)
                for la_b, lg_b, y in zip(torch.split(wide.lgalls[0], rows), torch.split(wide.lgs[0], rows), ys):
                    ops.ce_err(lg_b.contiguous(), y, acc.loss_sum, acc.err_cnt)
                    acc.add(y, lg_b, la_b)
        return self._finish_eval(acc, ph)

    def _reduce_batch(self, flat, y, acc, weights, per, with_ce=True):
        """One batch: ``flat`` [B,K,C*per] (sample index cycle * per + s) -> ``logits_all`` [B,K,per,C], per-cycle
        log-mean-softmax, weighted sum of log-probabilities (Appendix B.5; methods/csgld.py:416-439), CE / errors
        (``with_ce=False``: the caller accumulates them)."""
        B, K, C = flat.shape[0], flat.shape[1], len(weights)
        dev = flat.device
        if C == 0:
            logits_all = torch.zeros((B, self.args.num_classes, 1, 1), device=dev)
            batch_logits = None                            # as the reference: nothing to mix (its CE then fails, too)
        else:
            logits_all = flat.view(B, K, C, per).permute(0, 1, 3, 2).contiguous()      # torch.stack(comps, dim=3)
            batch_logits = torch.empty((B, K), dtype=torch.float32, device=dev)
        for ci, w in enumerate(weights):
            comp = logits_all[:, :, :, ci].contiguous()                                 # torch.stack(outs, 2) of cycle ci
            if self.nst == 0:      # raw logits of the mean parameters, no log-mean-softmax (csgld.py:419-420)
                wt = torch.tensor(w, dtype=torch.float32, device=dev)
                batch_logits = wt * comp.squeeze(2) if ci == 0 else batch_logits + wt * comp.squeeze(2)
            else:
                ops.ensemble(comp, batch_logits, self.nst, weight=w, mode=1 if ci == 0 else 2)
        if with_ce:
            ops.ce_err(batch_logits.contiguous(), y, acc.loss_sum, acc.err_cnt)
        acc.add(y, batch_logits, logits_all)

    # ---- cycle likelihoods and GMM weights (methods/csgld.py:508-594) ----------------------------------------
    def full_batch_likelihoods(self, train_loader):
        """exp(-mean CE over the training set) for max(1, nst) draws around the cycle's centre."""
        self.logger.info(f"Calculating full-batch likelihood for current cycle using {self.nst} samples...")
        ch = self._chain()
        dev = self.args.device
        c = self.current_cycle
        spec, center = None, None
        if self.LIKELIHOOD_MEAN == "theta":
            # centre = current theta, variance = the cycle's ratio*(mom2 - mom1^2)   (csgld.py:518-528, 541)
            center = ch.theta
            mean = self._cyc1.get(c)
            if c in self._cyc2 and c in self._cyc1 and self.samples_per_cycle.get(c, 0) > 1:
                spec = self._cycle_variance_spec(c)
        else:
            mean = self._cyc1[c]                          # KeyError if the cycle captured nothing, as the reference
            if self.CAPTURE == "welford":
                spec = self._cycle_variance_spec(c) if c in self._cyc2 else None
            elif c in self._cyc2 and self.samples_per_cycle.get(c, 0) > 1:
                spec = self._cycle_variance_spec(c)
            else:
                raise TypeError("cycle variance is None (reference: vector_to_parameters(None, ...), Appendix B.8)")
        ev = self._eval_net()
        self._eval_calls += 1
        n_draws = max(1, self.nst)
        rank, world = self._shard()
        # eval_shard=1: every draw is a full pass over the training set, so the draws are dealt round-robin to the ranks
        # (SURVEY 8f row 1); one all-reduce of the n_draws fp64 losses (zeros elsewhere: exact) hands every rank the same
        # list, each entry computed by the kernels one rank would have run
        losses = torch.zeros(n_draws, dtype=torch.float64, device=dev)
        for sample_idx in bdist.my_samples(n_draws, rank, world):
            if self.nst > 0 and spec is not None:
                second, var_mode, scale = spec
                ev.draw(mean, second, var_mode, scale, self.noise_mode, self.seed,
                        _pack_subseq(self._eval_calls, 0xFFFFFF, c, sample_idx), self.div_mode, center=center)
            else:
                ev.load(ch.theta)                        # net_sample = deepcopy(self.net), no perturbation
            avg_loss, _ = self._point_estimate(train_loader, ev.net, fwd=ev.forward)
            losses[sample_idx] = avg_loss
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(losses)
        likelihoods = []
        for sample_idx, avg_loss in enumerate(losses.tolist()):
            likelihood = np.exp(-avg_loss)
            likelihoods.append(likelihood)
            self.logger.info(f"Sample {sample_idx + 1} - Full batch average loss: {avg_loss:.6f}, "
                             f"likelihood: {likelihood:.6e}")
        return likelihoods

    def calculate_gmm_weights(self):
        """w_c = 1 / mean_j(1 / L_cj), normalised (host fp64; methods/csgld.py:565-594)."""
        if not self.cycle_likelihoods:
            return {0: 1.0}
        weights = {c: 1.0 / np.mean([1.0 / l for l in ls]) for c, ls in self.cycle_likelihoods.items()}
        total = sum(weights.values())
        if total > 0:
            return {c: w / total for c, w in weights.items()}
        return {c: 1.0 / len(weights) for c in weights}

    # ---- checkpoints (methods/csgld.py:470-506) ----------------------------------------------------------------
    def save_ckpt(self, epoch, wait=True):
        """Same file, keys and layout as methods/csgld.py:470-490; ``wait`` as in BurninRunner.save_ckpt."""
        fname = os.path.join(self.args.log_dir, f"{self.current_cycle}_ckpt.pt")
        last = self._dense(self._chain().theta) if self.LAST_THETA_AS_VECTOR else self._state_dict_copy()
        # moments of finished cycles never change again: only the newest cycle's buffers are snapshotted
        newest = max(self._cyc1) if self._cyc1 else None
        self._writer.submit(fname, {
            "last_theta": last,
            "cycle_theta_mom1": {c: self._dense_for_save(t, c == newest) for c, t in self._cyc1.items()},
            "cycle_theta_mom2": {c: self._dense_for_save(t, c == newest) for c, t in self._cyc2.items()},
            "cycle_likelihoods": dict(self.cycle_likelihoods),
            "cycle_states": dict(self.cycle_states),
            "epoch": epoch,
            "current_cycle": self.current_cycle,
            "samples_per_cycle": dict(self.samples_per_cycle),
        })
        if wait:
            self._writer.flush()
        return fname

    def load_ckpt(self, ckpt_path):
        self.flush_io()
        ckpt = torch.load(ckpt_path, map_location=self.args.device, weights_only=False)
        self.cycle_theta_mom1 = ckpt.get("cycle_theta_mom1", {})
        self.cycle_theta_mom2 = ckpt.get("cycle_theta_mom2", {})
        self.cycle_likelihoods = ckpt.get("cycle_likelihoods", {})
        self.current_cycle = ckpt.get("current_cycle", 0)
        self.samples_per_cycle = ckpt.get("samples_per_cycle", {})
        return ckpt["epoch"]
