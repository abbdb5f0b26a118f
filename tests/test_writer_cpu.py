"""CPU: host logic of the asynchronous writers (bayesdll_b200/writer.py) -- nesting, state_dict views, atomic files."""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from bayesdll_b200.flat import FlatLayout
from bayesdll_b200.writer import AsyncWriter, FlatBackedStateDict, to_host


class _Net(nn.Module):
    readout_name = "head"

    def __init__(self):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(1, 3, 3), nn.BatchNorm2d(3))
        self.head = nn.Linear(5, 3)


def test_flat_backed_state_dict_matches_state_dict(tmp_path):
    net = _Net()
    L = FlatLayout.from_module(net)
    flat = torch.zeros(L.n_padded)
    for v, p in zip(L.views(flat), net.parameters()):
        v.copy_(p.detach())
    names = [n for n, _ in net.named_parameters()]
    sd = FlatBackedStateDict.snapshot(net, L, names, flat)
    ref = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(sd[k], ref[k]) and sd[k].shape == ref[k].shape
    # parameter entries are views of the flat snapshot, buffers are independent clones
    assert sd["head.weight"].untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    assert sd["body.1.running_mean"].data_ptr() != net.body[1].running_mean.data_ptr()
    # written file loads as a plain state_dict into a fresh module
    w = AsyncWriter("cpu")
    path = w.submit(os.path.join(tmp_path, "sample.pth"), sd)
    w.flush()
    net2 = _Net()
    net2.load_state_dict(torch.load(path))
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    assert not [f for f in os.listdir(tmp_path) if ".tmp" in f]
    # pickles as an ordinary OrderedDict
    torch.save({"cycle_states": {1: sd}}, os.path.join(tmp_path, "c.pt"))
    back = torch.load(os.path.join(tmp_path, "c.pt"), weights_only=False)["cycle_states"][1]
    assert type(back) is OrderedDict and list(back.keys()) == list(ref.keys())


def test_to_host_preserves_structure():
    obj = {"a": torch.arange(3), "b": [torch.ones(2), 5, "x"], "c": OrderedDict(z=torch.zeros(1)), "d": None, "e": (1, 2)}
    out = to_host(obj)
    assert out["b"][1:] == [5, "x"] and out["d"] is None and out["e"] == (1, 2)
    assert type(out["c"]) is OrderedDict and torch.equal(out["a"], obj["a"])


def test_writer_rejects_bad_mode():
    import pytest
    with pytest.raises(ValueError):
        AsyncWriter("cpu", mode="later")


def test_async_writer_thread_path_on_cpu(tmp_path):
    """The threaded path (queue, atomic rename, pending set, flush) without CUDA."""
    w = AsyncWriter("cpu")
    paths = [w.submit(os.path.join(tmp_path, f"f{i}.pt"), {"i": i, "t": torch.full((1000,), float(i))}) for i in range(8)]
    w.flush()
    assert not w.pending_paths
    for i, p in enumerate(paths):
        d = torch.load(p)
        assert d["i"] == i and torch.equal(d["t"], torch.full((1000,), float(i)))
    assert w.stats["jobs"] == 8
    w.close()
    w.close()


def test_async_writer_reports_failures_and_never_hangs(tmp_path):
    import pytest

    def boom(obj, path):
        raise OSError("disk on fire")
    w = AsyncWriter("cpu")
    w.submit(os.path.join(tmp_path, "a.pt"), {"x": 1}, serializer=boom)
    with pytest.raises(RuntimeError, match="disk on fire"):
        w.flush()
    # the writer keeps working after a failed job
    p = w.submit(os.path.join(tmp_path, "b.pt"), {"x": 2})
    w.flush()
    assert torch.load(p)["x"] == 2
    # a writer thread that dies outside a job (this is what a bare torch.device('cuda') once did in set_device) must make
    # flush() raise, not wait forever
    w2 = AsyncWriter("cpu")
    w2._one = None                                   # makes the thread's loop raise TypeError on the first job
    w2.submit(os.path.join(tmp_path, "c.pt"), {"x": 3})
    with pytest.raises(RuntimeError):
        w2.flush()
    with pytest.raises(RuntimeError):
        w2.submit(os.path.join(tmp_path, "d.pt"), {"x": 4})


def test_async_writer_resolves_bare_cuda_device():
    import pytest
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA to resolve the current device")
    assert AsyncWriter(torch.device("cuda")).device.index is not None


def test_process_exits_cleanly_with_an_open_writer(tmp_path):
    """A script that never calls close() must still exit 0 with its file complete (atexit joins the writer thread)."""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent(f"""
        import os, torch
        from bayesdll_b200.writer import AsyncWriter
        w = AsyncWriter("cpu")
        w.submit(os.path.join({str(tmp_path)!r}, "late.pt"), {{"x": torch.arange(100000)}})
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    assert torch.equal(torch.load(os.path.join(tmp_path, "late.pt"))["x"], torch.arange(100000))
