"""GPU: the default-on CUDA-graph replay of the evaluation forward must never serve a stale batch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_replay_follows_in_place_refill_of_one_input_tensor(cuda_device):
    from bayesdll_b200.graphfwd import GraphedForward
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 7)).to(cuda_device).eval()
    fwd = GraphedForward(net, enabled=True)
    x = torch.empty(16, 32, device=cuda_device)          # ONE preallocated batch tensor, refilled in place
    with torch.no_grad():
        for i in range(6):
            x.copy_(torch.randn(16, 32, generator=torch.Generator().manual_seed(i)).to(cuda_device))
            got = fwd(x)
            assert torch.equal(got, net(x)), f"call {i}: replay returned another batch's logits"
    assert fwd.captures == 1 and fwd.replays >= 4
    # raw writes that bypass autograd's version counter (what a C-ABI kernel does) are seen as well
    with torch.no_grad():
        x.data.fill_(0.25)
        assert torch.equal(fwd(x), net(x))


@pytest.mark.parametrize("method", ["sghmc", "csgld"])
def test_cached_evaluation_copy_equals_a_fresh_copy_per_call(cuda_device, tmp_path, method):
    """The evaluation copy of the network (and its captured graphs) is kept across evaluate() calls and refreshed in place
    with the live network's buffers; hparams eval_cache=0 builds a fresh deepcopy per call like round 1 did.  Three calls
    with the BatchNorm statistics of the LIVE network changed in between (Appendix B.12: every evaluation inherits the
    current statistics): all outputs bit-identical, and the cached runner captures each input shape once."""
    import numpy as np
    import shard_util as su
    from bayesdll_b200.graphfwd import GraphedForward
    loader = su.make_loader()
    outs, captures = {}, {}
    for cache in (1, 0):
        d = tmp_path / f"c{cache}"
        d.mkdir()
        runner = su.make_runner(method, cuda_device, d, nst=3, eval_shard=0)
        runner.eval_cache = bool(cache)
        before = GraphedForward.total_captures
        res = []
        for call in range(3):
            with torch.no_grad():
                bn = runner.net.features[1]
                bn.running_mean.add_(0.05 * call)
                bn.running_var.mul_(1.0 + 0.1 * call)
            res.append(runner.evaluate(loader))
        if hasattr(runner, "full_batch_likelihoods"):
            res.append((0.0, 0.0) + tuple(np.asarray(runner.full_batch_likelihoods(su.make_loader(seed=9)))[None] for _ in range(3)))
        outs[cache], captures[cache] = res, GraphedForward.total_captures - before
        if cache:
            assert runner._eval_cached is not None and runner._eval_cached.src is runner.net
            runner.release_eval_cache()
            assert runner._eval_cached is None
        else:
            assert runner._eval_cached is None
        runner.flush_io()
    n_shapes = len({tuple(x.shape) for x, _ in loader})
    assert captures[1] <= n_shapes + 1 < captures[0], captures    # cached: one capture per shape (+1: the likelihood loader)
    for a, b in zip(outs[1], outs[0]):
        assert a[0] == b[0] and a[1] == b[1]
        for u, v in zip(a[2:], b[2:]):
            assert np.array_equal(np.asarray(u), np.asarray(v))
    # the calls differ from each other (the statistics really changed): the refresh is not a no-op
    assert not np.array_equal(outs[1][0][3], outs[1][1][3])


def test_capture_keeps_the_cyclic_collector_out(cuda_device):
    """A CUDAGraph finalised by Python's cyclic collector WHILE another graph is being captured invalidates that capture
    (cudaGraphExecDestroy is not permitted while a stream captures) -- and a Runner's cached evaluation graphs live until
    the collector finds the runner.  graphfwd.capture_gc_guard: collectable garbage is finalised BEFORE the capture starts
    and the collector stays off until it has ended."""
    import gc
    import warnings
    from bayesdll_b200.graphfwd import GraphedForward
    seen = {}

    class Owner:                                          # cyclic garbage that owns captured graphs, like a dropped Runner
        def __init__(self, fwd):
            self.me, self.fwd = self, fwd

        def __del__(self):
            seen["finalised_while_capturing"] = torch.cuda.is_current_stream_capturing()

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(16, 4)

        def forward(self, x):
            if torch.cuda.is_current_stream_capturing():
                seen["collector_enabled_during_capture"] = gc.isenabled()
                [[i] for i in range(5000)]               # enough container allocations to trigger generation-0 collections
            return self.lin(x)

    x = torch.randn(8, 16, device=cuda_device)
    was = gc.isenabled()
    gc.disable()                                          # the garbage survives until the guard collects it
    try:
        old = GraphedForward(Net().to(cuda_device).eval())
        with torch.no_grad():
            for _ in range(3):
                old(x)
        assert old.captures == 1
        Owner(old)
        del old
        fwd = GraphedForward(Net().to(cuda_device).eval())
        with torch.no_grad():
            fwd(x)                                        # eager
            gc.enable()
            with warnings.catch_warnings():
                warnings.simplefilter("error")            # a failed capture warns and falls back to eager
                out = fwd(x)                              # captures
            assert torch.equal(out, fwd.net(x))
        assert fwd.captures == 1
        assert seen == {"finalised_while_capturing": False, "collector_enabled_during_capture": False}
        assert gc.isenabled()                             # restored
    finally:
        gc.enable() if was else gc.disable()


def test_runner_default_captures_framework_networks_and_only_those(cuda_device, tmp_path):
    """hparams graph_train defaults to 'auto': a Runner training a network built from torch.nn modules replays forward +
    backward as a CUDA graph from its third step on, bit-identical to graph_train=0; a user-defined module class stays
    eager under the default."""
    import argparse
    import logging
    from bayesdll_b200.methods import sghmc
    import shard_util as su
    monkey = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        def make_net(seed, custom):
            torch.manual_seed(seed)
            if custom:
                return su.ConvNet()
            net = torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                                      torch.nn.AdaptiveAvgPool2d(4), torch.nn.Flatten(), torch.nn.Linear(96, su.K))
            net.readout_name = "5"
            return net

        def run(hp_extra, custom=False):
            hp = dict(prior_sig=1.0, Ninflate=10.0, nd=1.0, burnin=0, thin=1, nst=2, bias="informative", momentum_decay=0.18, seed=5,
                      **hp_extra)
            a = argparse.Namespace(device=cuda_device, ND=64, lr=1e-3, lr_head=1e-2, momentum=0.5, epochs=1, pretrained="synthetic",
                                   hparams={k: str(v) for k, v in hp.items()}, test_eval_freq=1, ece_num_bins=15, num_classes=su.K,
                                   log_dir=str(tmp_path), seed=5)
            lg = logging.getLogger("auto")
            lg.addHandler(logging.NullHandler())
            lg.propagate = False
            runner = sghmc.Runner(make_net(1, custom), make_net(2, custom), a, lg)
            runner.net.train()
            gen = torch.Generator().manual_seed(9)
            outs = []
            for _ in range(6):
                x, y = torch.randn(8, 3, 8, 8, generator=gen).to(cuda_device), torch.randint(0, su.K, (8,), generator=gen).to(cuda_device)
                loss, out = runner.model(x, y, runner.net, runner.net0, runner.criterion, [1e-3, 1e-2], runner.Ninflate, runner.nd)
                outs.append((loss, out.clone(), runner.model.chain.theta.clone(), [b.clone() for b in runner.net.buffers()]))
            captured = sum(isinstance(v, dict) for v in runner.model._train_graphs.values())
            runner.flush_io()
            return outs, captured, runner.model._opts["graph_train"]

        auto, n_auto, mode = run({})
        eager, n_eager, _ = run(dict(graph_train=0))
        assert mode == "auto" and n_auto == 1 and n_eager == 0
        for (l1, o1, t1, b1), (l0, o0, t0, b0) in zip(auto, eager):
            assert l1 == l0 and torch.equal(o1, o0) and torch.equal(t1, t0) and all(torch.equal(a, b) for a, b in zip(b1, b0))
        _, n_custom, _ = run({}, custom=True)
        assert n_custom == 0                                # su.ConvNet is a user-defined class: not captured by default
        _, n_forced, _ = run(dict(graph_train=1), custom=True)
        assert n_forced == 1
    finally:
        torch.backends.cudnn.deterministic = monkey


def test_training_graphs_are_capped_per_model(cuda_device):
    """A loader that keeps producing new batch shapes must not keep producing CUDA graphs (each owns a private pool with a
    full set of activations): FusedModel._MAX_TRAIN_GRAPHS shapes are replayed, later ones run eagerly -- same results."""
    from bayesdll_b200.methods import sgld
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).to(cuda_device).train()
    net.readout_name = "2"
    net0 = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).to(cuda_device)
    model = sgld.Model(ND=100, prior_sig=1.0, bias="informative")
    model.configure(seed=3, graph_train="auto")
    crit = torch.nn.CrossEntropyLoss()
    gen = torch.Generator().manual_seed(1)
    for rep in range(4):
        for b in range(2, 9):                                # 7 batch shapes, each seen 4 times
            x, y = torch.randn(b, 6, generator=gen).to(cuda_device), torch.randint(0, 3, (b,), generator=gen).to(cuda_device)
            loss, out = model.forward(x, y, net, net0, crit, [1e-2, 1e-2], Ninflate=1.0, nd=1.0)
            assert out.shape == (b, 3) and loss == loss
    kinds = list(model._train_graphs.values())
    assert sum(isinstance(v, dict) for v in kinds) == model._MAX_TRAIN_GRAPHS == 4
    assert kinds.count("eager") == 3
