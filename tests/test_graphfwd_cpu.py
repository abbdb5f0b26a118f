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
