"""CPU: host-side logic of the runners that needs no device -- schedules, capture bookkeeping, quirks."""
import argparse

import pytest

from bayesdll_b200.methods import _base, csghmc_fs


def _bare(cls, **attrs):
    r = cls.__new__(cls)
    for k, v in attrs.items():
        setattr(r, k, v)
    return r


@pytest.mark.parametrize("epochs,cycles", [(6, 2), (40, 8), (30, 3), (12, 4), (7, 2), (100, 10), (3, 1)])
def test_full_sample_epochs_match_reference_condition(epochs, cycles):
    """methods/csghmc_fs.py:176: ``ep % L > L - 4 and ep % L < L - 1`` with ``L = epochs // num_cycles``."""
    r = _bare(csghmc_fs.Runner, args=argparse.Namespace(epochs=epochs, num_cycles=cycles))
    L = epochs // cycles
    want = [ep for ep in range(epochs) if ep % L > L - 4 and ep % L < L - 1]
    got = [ep for ep in range(epochs) if r._stores_full_sample(ep)]
    assert got == want
    assert r._expected_full_samples() == len(want)
    if L >= 3:
        assert len(want) >= 2 * cycles          # two dumps per complete cycle


def test_full_sample_file_order_is_lexicographic(tmp_path):
    """The reference sorts file names as strings (csghmc_fs.py:270): ep10 comes before ep2.  The BMA sums logits in that
    order, so the drop-in must list them the same way."""
    class _W:
        def is_pending(self, p):
            return p.endswith("ep3.pth")
    for ep in (2, 10, 1):
        (tmp_path / f"full_samples_net_ep{ep}.pth").write_bytes(b"x")
    (tmp_path / "other.pth").write_bytes(b"x")
    r = _bare(csghmc_fs.Runner, args=argparse.Namespace(log_dir=str(tmp_path)), _writer=_W(),
              _fs_resident={"full_samples_net_ep3.pth": object(), "full_samples_net_ep9.pth": object()})
    assert r._full_sample_files() == ["full_samples_net_ep1.pth", "full_samples_net_ep10.pth", "full_samples_net_ep2.pth",
                                      "full_samples_net_ep3.pth"]     # ep3: still being written; ep9: neither on disk nor pending


def test_gmm_weights_follow_reference_formula():
    """w_c = 1 / mean_j(1 / L_cj), normalised (methods/csgld.py:565-594); no likelihoods -> {0: 1.0}."""
    r = _bare(_base.CyclicalRunner, cycle_likelihoods={})
    assert r.calculate_gmm_weights() == {0: 1.0}
    r.cycle_likelihoods = {1: [0.5, 0.25], 2: [0.1, 0.1]}
    w = r.calculate_gmm_weights()
    raw = {1: 1 / ((2 + 4) / 2), 2: 0.1}
    tot = sum(raw.values())
    assert w == pytest.approx({c: x / tot for c, x in raw.items()})


def test_cycle_variance_spec_keeps_reference_quirks():
    """'avg' scheme: ratio n/(n-1) is evaluated before the n > 1 guard -> ZeroDivisionError for a one-sample cycle
    (Appendix B.7); Welford: divisor count-1 with the double-counted count, tiny variance for count <= 1."""
    from bayesdll_b200 import ops
    avg = _bare(_base.CyclicalRunner, samples_per_cycle={1: 1, 2: 4}, _cyc2={1: "m2a", 2: "m2b"})
    with pytest.raises(ZeroDivisionError):
        avg._cycle_variance_spec(1)
    assert avg._cycle_variance_spec(2) == ("m2b", ops.VAR_FROM_MOMENTS, 4 / 3)

    class W(_base.CyclicalRunner):
        CAPTURE = "welford"
    wel = _bare(W, samples_per_cycle={1: 1, 2: 6}, _cyc2={1: "a", 2: "b"})
    assert wel._cycle_variance_spec(1) == (None, ops.VAR_TINY, 1.0)
    assert wel._cycle_variance_spec(2) == ("b", ops.VAR_FROM_WELFORD, 5.0)


@pytest.mark.parametrize("method", ["sgld", "sghmc", "csgld", "csghmc", "csghmc_fs", "adam_sghmc", "adam_csghmc"])
def test_runner_and_model_carry_the_reference_surface(method, tmp_path):
    """Every public method of the reference's Runner / Model and every attribute their __init__ sets
    (tests/golden/api_inventory.json, extracted from the reference's sources by oracle/make_golden.py attrs) exists on
    the drop-in objects right after construction.  Construction needs no device: the flat state is built on first use."""
    import importlib
    import json
    import logging
    import os
    import torch
    import golden_util as gu
    from oracle import make_golden_runner as mgr
    inv = json.load(open(os.path.join(os.path.dirname(gu.golden_path("cyclical")), "api_inventory.json")))
    hp = dict(prior_sig=1.0, Ninflate=1.0, nd=1.0, burnin=1, thin=1, nst=2, bias="informative", momentum_decay=0.1)
    args = mgr.make_args(hp, str(tmp_path), torch.device("cpu"), momentum=0.5, epochs=4, num_cycles=2)
    lg = logging.getLogger("surface")
    lg.addHandler(logging.NullHandler())
    mod = importlib.import_module(f"bayesdll_b200.methods.{method}")
    runner = mod.Runner(mgr.InjectNet(1), mgr.InjectNet(2), args, lg)
    for kind, obj in (("Runner", runner), ("Model", runner.model)):
        entry = inv[f"{method}.{kind}"]
        missing = [m for m in entry["methods"] if not callable(getattr(obj, m, None))]
        assert not missing, f"{method}.{kind} lacks methods {missing}"
        missing = [a for a in entry["init_attributes"] if not hasattr(obj, a)]
        assert not missing, f"{method}.{kind} lacks attributes {missing}"
    # the values other code reads have the reference's types
    assert isinstance(runner.Ninflate, float) and isinstance(runner.nst, int) and isinstance(runner.thin, int)
    assert [g["lr"] for g in runner.optimizer.param_groups] == [args.lr, args.lr_head]
    assert isinstance(runner.criterion, torch.nn.CrossEntropyLoss)
    # no CPU fallback: the first step on a CPU device is refused
    from bayesdll_b200 import _lib
    x, y = torch.zeros(4, 1, 4, 4), torch.zeros(4, dtype=torch.long)
    with pytest.raises(_lib.BdlError, match="CUDA devices only"):
        runner.model(x, y, runner.net, runner.net0, mgr.InjectCriterion(), [1e-3, 1e-2], 1.0, 1.0)
