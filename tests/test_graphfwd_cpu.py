"""CPU: host logic of the CUDA-graph wrappers.  Without a CUDA device they must be transparent pass-throughs (the
backbone forward is PyTorch plumbing, not the product path), and the opt-in training capture must decline CPU tensors."""
import torch

from bayesdll_b200.graphfwd import GraphedForward


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(4, 3)
        self.calls = 0

    def forward(self, x):
        self.calls += 1
        return self.lin(x)


def test_graphed_forward_is_a_pass_through_without_cuda():
    net = _Net().eval()
    fwd = GraphedForward(net, enabled=True)
    x = torch.randn(5, 4)
    with torch.no_grad():
        a, b, c = fwd(x), fwd(x), fwd(x)
    assert torch.equal(a, net.lin(x)) and torch.equal(a, b) and torch.equal(b, c)
    assert net.calls == 3 and fwd.captures == 0 and fwd.replays == 0
    off = GraphedForward(net, enabled=False)
    with torch.no_grad():
        assert torch.equal(off(x), a)
    assert off._entries == {}


def test_training_capture_declines_cpu_tensors_and_eval_mode():
    from bayesdll_b200.methods import sghmc
    model = sghmc.Model(ND=10)
    model.configure(graph_train=True)
    net = _Net()
    x, y = torch.randn(5, 4), torch.randint(0, 3, (5,))
    assert model._graphed_fwd_bwd(x, y, net, torch.nn.CrossEntropyLoss()) is None      # CPU tensor: eager
    assert model._train_graphs == {}
    try:
        model.configure(no_such_option=1)
    except TypeError as e:
        assert "no_such_option" in str(e)
    else:
        raise AssertionError("unknown option accepted")


def test_graph_train_auto_only_trusts_framework_modules():
    """hparams graph_train=auto (the Runner default): forward + loss + backward are captured only when the network AND the
    criterion consist of framework-provided modules (torch.nn, the torchvision model zoo, the reference's MLP), whose
    forward has no host-side control flow on tensor values; anything user-defined stays eager unless graph_train=1."""
    import torchvision
    from bayesdll_b200 import shapes
    from bayesdll_b200.methods import sghmc
    from bayesdll_b200.methods._base import _RunnerCommon
    assert _RunnerCommon._parse_graph_train("auto") == "auto" and _RunnerCommon._parse_graph_train("AUTO ") == "auto"
    assert _RunnerCommon._parse_graph_train("1") is True and _RunnerCommon._parse_graph_train("0.0") is False
    model = sghmc.Model(ND=10)
    assert model._opts["graph_train"] is False            # a bare Model keeps the explicit opt-in; the Runner configures "auto"
    model.configure(graph_train="auto")
    ce = torch.nn.CrossEntropyLoss()
    seq = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.LazyLinear(5))
    assert model._graph_train_wanted(torchvision.models.resnet18(), ce)
    assert model._graph_train_wanted(torchvision.models.vit_b_32(), ce)
    assert model._graph_train_wanted(shapes.create_backbone("mlp_mnist", 10), ce)
    assert model._graph_train_wanted(seq, ce)
    assert not model._graph_train_wanted(_Net(), ce)                                  # user-defined forward
    assert not model._graph_train_wanted(torch.nn.Sequential(torch.nn.Linear(4, 4), _Net()), ce)   # ... anywhere inside
    assert not model._graph_train_wanted(seq, lambda out, y: out.sum())                # user-defined criterion
    model.configure(graph_train=True)
    assert model._graph_train_wanted(_Net(), ce)
    model.configure(graph_train=False)
    assert not model._graph_train_wanted(seq, ce)


def test_eval_copy_refresh_inherits_live_buffers_in_place():
    """_EvalNet.refresh (the cached evaluation copy, methods/_base.py): the live network's buffers are copied IN PLACE -- the
    addresses a captured graph reads stay valid --, parameters stay views of the copy's own flat buffer, and a network whose
    buffers no longer match is refused (the caller then builds a fresh copy)."""
    from bayesdll_b200.flat import FlatLayout
    from bayesdll_b200.methods._base import _EvalNet
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.BatchNorm1d(6), torch.nn.Linear(6, 3))
    net.readout_name = "2"
    lay = FlatLayout([(n, tuple(p.shape)) for n, p in net.named_parameters()], "2")
    ev = _EvalNet(net, lay, graph=True)
    bn_live, bn_copy = net[1], ev.net[1]
    ptrs = [b.data_ptr() for b in ev.net.buffers()]
    with torch.no_grad():
        bn_live.running_mean.add_(1.5)
        bn_live.running_var.mul_(2.0)
        bn_live.num_batches_tracked.add_(3)
    assert not torch.equal(bn_copy.running_mean, bn_live.running_mean)
    assert ev.refresh(net) is True
    assert torch.equal(bn_copy.running_mean, bn_live.running_mean) and torch.equal(bn_copy.running_var, bn_live.running_var)
    assert int(bn_copy.num_batches_tracked) == int(bn_live.num_batches_tracked)
    assert [b.data_ptr() for b in ev.net.buffers()] == ptrs                  # in place
    assert all(p.data_ptr() != q.data_ptr() for p, q in zip(ev.net.parameters(), net.parameters()))
    assert not ev.net.training and ev.net[0].weight.untyped_storage().data_ptr() == ev.flat.untyped_storage().data_ptr()
    other = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    assert ev.refresh(other) is False
