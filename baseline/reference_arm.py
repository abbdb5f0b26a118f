"""Drive and TIME the unmodified reference (omarezz46/BayesDLL) for bench.py's reference arm.

BASELINE INFRASTRUCTURE: imported by bench.py's ``--impl reference`` / ``cpu_baseline`` / ``reference_eager_gpu`` /
``cfg1`` legs only, never by ``bayesdll_b200/``.  The reference tree is located by ``oracle/refshim.find_reference``
($BDL_REF, /root/reference, baseline/_ref -- the latter is what exists on the GPU box, see baseline/install_ref.py).

* ``sghmc_update_rate``  the reference's own ``methods/sghmc.Model.forward`` (:435-512) + ``torch.optim.SGD.step``
  (:229, optimizer built as :53-57) on parameter tensors of the ViT-L/32 shapes, gradients handed out by
  ``TimedInjectNet`` (no backbone cost), ``criterion = identity`` -- the time is the per-tensor update loop's.
* ``cfg1_mlp_mnist``     BASELINE.json configs[0]: the reference's ``methods/sgld.Runner`` on its own ``mlp_mnist``
  backbone, synthetic 28x28 batches of 128: ms/step, ensemble preds/s (nst=5), ``calibration.analyze`` ms.
"""
import argparse
import logging
import os
import time
from collections import OrderedDict

import numpy as np
import torch


def available():
    from oracle import refshim
    try:
        return refshim.find_reference()
    except FileNotFoundError:
        return None


def _quiet_logger():
    lg = logging.getLogger("reference_arm")
    lg.addHandler(logging.NullHandler())
    lg.propagate = False
    return lg


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def sghmc_update_rate(named_shapes, readout_name, device, *, steps, warmup, hp, seconds=None, max_params=None):
    """params/s of the reference's SGHMC ``Model.forward`` + ``optimizer.step()`` over the first tensors of ``named_shapes``
    (all of them unless ``max_params`` / ``seconds`` bound the sample).  ``seconds``: size the sample so that
    ``steps + warmup`` steps take about that long (probed on the first ~8 Mi parameters)."""
    from oracle import refshim
    sghmc = refshim.load("methods.sghmc")
    device = torch.device(device)
    if device.type == "cpu":
        torch.set_num_threads(host_cores())

    def build(shapes):
        net = refshim.TimedInjectNet(OrderedDict(shapes), readout_name, init_std=0.02, seed=1).to(device)
        net0 = refshim.TimedInjectNet(OrderedDict(shapes), readout_name, init_std=0.02, seed=2).to(device)
        gen = torch.Generator(device=device).manual_seed(3)
        net.set_grads([torch.randn(p.shape, device=device, generator=gen) * 0.01 for p in net.parameters()])
        model = sghmc.Model(ND=hp["ND"], prior_sig=hp["prior_sig"], bias="informative", momentum_decay=hp["alpha"]).to(device)
        opt = torch.optim.SGD(                                       # methods/sghmc.py:53-57
            [{"params": [p for pn, p in net.named_parameters() if readout_name not in pn], "lr": hp["lr_body"]},
             {"params": [p for pn, p in net.named_parameters() if readout_name in pn], "lr": hp["lr_head"]}],
            momentum=0, weight_decay=0)
        x = torch.zeros(1, device=device)
        y = torch.zeros(1, dtype=torch.long, device=device)

        def one():
            model(x, y, net, net0, refshim.identity_criterion, [pg["lr"] for pg in opt.param_groups], hp["Ninflate"], hp["nd"])
            opt.step()
        return one, sum(int(np.prod(s)) for _, s in shapes)

    def take(limit):
        out, tot = [], 0
        for name, shape in named_shapes:
            out.append((name, tuple(shape)))
            tot += int(np.prod(shape))
            if limit is not None and tot >= limit:
                break
        return out

    limit = max_params
    if seconds is not None:
        one, n_probe = build(take(8 << 20))
        one()
        _sync(device)
        t0 = time.perf_counter()
        one()
        _sync(device)
        rate = n_probe / (time.perf_counter() - t0)
        want = int(rate * seconds / (steps + warmup))
        limit = max(8 << 20, want if limit is None else min(limit, want))
        del one
    shapes = take(limit)
    one, n = build(shapes)
    for _ in range(warmup):
        one()
    _sync(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    _sync(device)
    dt = time.perf_counter() - t0
    total = sum(int(np.prod(s)) for _, s in named_shapes)
    return {"value": n * steps / dt, "unit": "params/s", "ms_per_step": dt / steps * 1e3, "steps": steps, "warmup": warmup,
            "n": n, "tensors": len(shapes), "cores": host_cores() if device.type == "cpu" else None, "kind": "reference",
            "sample": f"{steps} steps of the reference's own methods/sghmc.py Model.forward (per-tensor update loop, "
                      f"torch.randn_like noise) + torch.optim.SGD.step on {device.type}, over the first {len(shapes)} of "
                      f"{len(named_shapes)} ViT-L/32 parameter tensors ({n} of {total} parameters), gradients injected "
                      f"(no backbone cost)"}


def train_step_ms(backbone, device, *, batch, steps, hp, method="sghmc", num_classes=37, seed=42):
    """ms per training step of the UNMODIFIED reference on ``device``: its own ``networks.create_backbone`` (torchvision
    ResNet-101 / ViT-L/32, random init), its ``methods/<method>.Runner`` and the statements of its ``train_one_epoch``
    (methods/sghmc.py:220-236: batch to the device, ``Model.forward`` = forward + backward + per-tensor update loop,
    ``optimizer.step()``, prediction error) on synthetic 224x224 batches from pinned host memory -- the whole user-visible
    step the drop-in's ``Model.forward`` replaces."""
    import importlib
    import tempfile
    from oracle import refshim
    device = torch.device(device)
    torch.manual_seed(seed)
    with refshim.reference_imports():
        networks = importlib.import_module("networks")
        mod = importlib.import_module(f"methods.{method}")
    a = argparse.Namespace(device=device, ND=hp["ND"], lr=hp["lr_body"], lr_head=hp["lr_head"], momentum=0.5, epochs=1,
                           pretrained="synthetic", num_classes=num_classes, backbone=backbone, ece_num_bins=15, test_eval_freq=1,
                           log_dir=tempfile.mkdtemp(prefix="bdl_refstep_"), seed=seed, num_cycles=1, proportion_exploration=0.5,
                           full_sample=False,
                           hparams=dict(prior_sig=str(hp["prior_sig"]), Ninflate=str(hp["Ninflate"]), nd=str(hp["nd"]),
                                        momentum_decay=str(hp["alpha"]), burnin="5", thin="1", bias="informative", nst="5"))
    net, net0 = networks.create_backbone(a), networks.create_backbone(a)
    runner = mod.Runner(net, net0, a, _quiet_logger())
    runner.net.train()
    x_host = torch.randn(batch, 3, 224, 224)
    y_host = torch.randint(0, num_classes, (batch,))
    if device.type == "cuda":
        x_host, y_host = x_host.pin_memory(), y_host.pin_memory()

    def one():
        x, y = x_host.to(device), y_host.to(device)
        lrs = [pg["lr"] for pg in runner.optimizer.param_groups]
        if method == "csghmc":                            # methods/csghmc.py:296-304: p.data updated inside, no optimizer.step()
            loss_, out = runner.model(x, y, runner.net, runner.net0, runner.criterion, lrs, runner.Ninflate, runner.nd,
                                      should_sample=True)
        else:                                             # methods/sghmc.py:220-229
            loss_, out = runner.model(x, y, runner.net, runner.net0, runner.criterion, lrs, runner.Ninflate, runner.nd)
            runner.optimizer.step()
        pred = out.data.max(dim=1)[1]
        return loss_, pred.ne(y.data).sum().item()
    for _ in range(2):
        one()
    _sync(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    _sync(device)
    return (time.perf_counter() - t0) / steps * 1e3


def cfg1_mlp_mnist(device="cpu", batches=20, batch_size=128, test_batches=8, reference=True, seed=42, graph_train=False):
    """BASELINE.json configs[0] (README.md:83): mlp_mnist SGLD, prior_sig=1, Ninflate=1e3, nd=1, burnin=5, thin=10, nst=5,
    lr 1e-2, momentum 0.5, synthetic 28x28 batches of 128, ND = 30 000.  ``reference=True``: the unmodified reference Runner
    (CPU by contract); False: the drop-in Runner on ``device``.  -> ms/step, param updates/s, ensemble preds/s, analyze ms."""
    import tempfile
    device = torch.device(device)
    torch.manual_seed(seed)
    np.random.seed(seed)
    if device.type == "cpu":
        torch.set_num_threads(host_cores())
    hp = dict(prior_sig="1.0", Ninflate="1e3", nd="1.0", burnin="5", thin="10", bias="informative", nst="5")
    if not reference:                                     # drop-in only: the default (auto = CUDA-graph replay for this MLP) or eager
        hp["graph_train"] = "auto" if graph_train else "0"
    a = argparse.Namespace(device=device, ND=30000, lr=1e-2, lr_head=1e-2, momentum=0.5, epochs=1, pretrained=None,
                           hparams=hp, test_eval_freq=1, ece_num_bins=15, num_classes=10, backbone="mlp_mnist",
                           log_dir=tempfile.mkdtemp(prefix="bdl_cfg1_"), seed=seed)
    gen = torch.Generator().manual_seed(seed)
    train = [(torch.randn(batch_size, 1, 28, 28, generator=gen), torch.randint(0, 10, (batch_size,), generator=gen))
             for _ in range(batches)]
    test = [(torch.randn(batch_size, 1, 28, 28, generator=gen), torch.randint(0, 10, (batch_size,), generator=gen))
            for _ in range(test_batches)]
    if reference:
        from oracle import refshim
        with refshim.reference_imports():
            import importlib
            networks = importlib.import_module("networks")
            sgld = importlib.import_module("methods.sgld")
            calibration = importlib.import_module("calibration")
        net = networks.create_backbone(a)
    else:
        from bayesdll_b200 import calibration
        from bayesdll_b200.methods import sgld
        from bayesdll_b200 import shapes
        net = shapes.create_backbone("mlp_mnist", 10)      # same architecture as networks/small_nets.MLP(784, 10, 1000, 3)
    runner = sgld.Runner(net, None, a, _quiet_logger())
    # warm-up epoch fragment, then the timed epoch (collect=True: moments every `thin` steps, as after burn-in)
    runner.train_one_epoch(train[:4] if graph_train else train[:3], collect=False, bi=0)   # graph_train: 2 eager + capture + replay
    n_params = sum(p.numel() for p in runner.net.parameters())
    with torch.no_grad():
        theta = torch.nn.utils.parameters_to_vector(runner.net.parameters())
    if reference:
        runner.post_theta_mom1, runner.post_theta_mom2, runner.post_theta_cnt = theta * 1.0, theta ** 2, 1
    else:
        runner._start_collecting()
    _sync(device)
    t0 = time.perf_counter()
    runner.train_one_epoch(train, collect=True, bi=0)
    _sync(device)
    step_ms = (time.perf_counter() - t0) / len(train) * 1e3
    runner.evaluate(test[:1])
    _sync(device)
    t0 = time.perf_counter()
    loss, err, targets, logits, logits_all = runner.evaluate(test)
    _sync(device)
    eval_s = time.perf_counter() - t0
    if not reference:
        calibration.analyze(targets, logits, 15, None)    # first call: edge upload, allocator warm-up
    t0 = time.perf_counter()
    if reference:
        ece, mce, nll = calibration.analyze(targets, logits, num_bins=15, plot_save_path=os.path.join(a.log_dir, "r.png"), temperature=1)
    else:
        ece, mce, nll = calibration.analyze(targets, logits, 15, None)
    analyze_ms = (time.perf_counter() - t0) * 1e3
    rows = len(targets)
    return {"impl": "reference (unmodified methods/sgld.py Runner)" if reference else
            ("bayesdll_b200.methods.sgld.Runner (default: graph_train=auto)" if graph_train else "bayesdll_b200.methods.sgld.Runner, hparams graph_train=0"),
            "device": str(device), "cores": host_cores() if device.type == "cpu" else None, "params": n_params,
            "ms_per_step": step_ms, "param_updates_per_s": n_params / (step_ms * 1e-3), "batch": batch_size,
            "ensemble_preds_per_s": rows * 5 / eval_s, "eval_rows": rows, "nst": 5, "analyze_ms": analyze_ms,
            "ece": float(ece), "nll": float(nll)}
