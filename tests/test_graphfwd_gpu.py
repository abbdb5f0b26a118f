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
