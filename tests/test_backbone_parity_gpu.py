"""BASELINE.json configs[1], [2], [3] at full size with real autograd: the drop-in ``Model.forward`` on torchvision
ResNet-101 / ViT-L/32 next to the reference's statements (tests/eager_reference.py: its update loop in torch CUDA eager
ops + the real ``torch.optim.SGD``) on the SAME GPU, same batches, same seeded ``torch.randn_like`` stream.

Cross-device recordings are useless here: random-init deep networks are numerically chaotic (plain PyTorch CPU vs GPU
gradients of ResNet-101's conv1 differ by 6 % at step 1).  On one device both arms see identical inputs at every step,
so the parameters (42.6 M / 305.5 M), the momentum, the Adam moments and the BatchNorm buffers must stay **bit-identical**
step after step -- any deviation of the fused kernel, the flat layout, the per-tensor gradient-pointer table or the
division semantics would show up and be amplified by the network."""
import copy
import logging

import numpy as np
import pytest
import torch

import eager_reference as er

pytestmark = pytest.mark.gpu


def _logger():
    lg = logging.getLogger("backbone_parity")
    lg.addHandler(logging.NullHandler())
    lg.propagate = False
    return lg


@pytest.fixture
def deterministic_fp32():
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
             torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    yield
    (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
     torch.backends.cudnn.benchmark) = saved


def _assert_identical(step, runner, named, vs, ms=None, ss=None):
    ch = runner.model.chain
    for (n, q), th, v in zip(named, ch.layout.views(ch.theta), ch.layout.views(ch.v)):
        assert torch.equal(th, q.data), f"step {step}: theta[{n}] differs (max {(th - q.data).abs().max().item():.3e})"
        assert torch.equal(v, vs[n]), f"step {step}: momentum[{n}] differs"
    if ms is not None:
        for (n, _), m, s in zip(named, ch.layout.views(ch.m), ch.layout.views(ch.s)):
            assert torch.equal(m, ms[n]) and torch.equal(s, ss[n]), f"step {step}: Adam moments[{n}] differ"


@pytest.mark.parametrize("graph_train", [0, 1])
def test_cfg2_resnet101_csghmc_training_steps(cuda_device, deterministic_fp32, tmp_path, graph_train):
    """configs[1]: ResNet-101 cSGHMC, batches of 16 (methods/csghmc.py:747-778; alternating exploration / sampling).
    graph_train=1: the drop-in replays forward + backward as one CUDA graph from its third step on -- still bit-identical."""
    from oracle import make_golden_runner as mgr
    from bayesdll_b200.methods import csghmc
    dev = cuda_device
    net, net0 = mgr.cfg2_networks()
    ref = copy.deepcopy(net).to(dev)
    args = mgr.cfg2_args(str(tmp_path), dev, extra_hp=dict(noise="torch", graph_train=graph_train))
    runner = csghmc.Runner(net, net0, args, _logger())
    crit = torch.nn.CrossEntropyLoss()
    named = list(ref.named_parameters())
    vs = {n: torch.zeros_like(p) for n, p in named}
    N = args.ND * runner.Ninflate
    ref.train()
    runner.net.train()
    lrs = [args.lr, args.lr_head]
    for t, (x, y) in enumerate(mgr.cfg2_loaders()[0] * 2):
        x, y = x.to(dev), y.to(dev)
        sampling = t % 2 == 1
        torch.manual_seed(50 + t)
        loss_a, _ = runner.model(x, y, runner.net, runner.net0, crit, lrs, runner.Ninflate, runner.nd, should_sample=sampling)
        loss = crit(ref(x), y)
        ref.zero_grad()
        loss.backward()
        torch.manual_seed(50 + t)
        xis = [torch.randn_like(p) for _, p in named]
        er.csghmc(named, xis, vs, "fc", lr_body=lrs[0], lr_head=lrs[1], N=N, prior_sig=1.0, nd=runner.nd, alpha=0.18,
                  should_sample=sampling)
        assert loss_a == loss.item()
        _assert_identical(t, runner, named, vs)
        for (bn, a), (_, b) in zip(runner.net.named_buffers(), ref.named_buffers()):
            assert torch.equal(a, b), f"step {t}: buffer {bn} differs"
    assert runner.model.chain.layout.n_dense == 42575973
    captured = sum(isinstance(v, dict) for v in runner.model._train_graphs.values())
    assert captured == (1 if graph_train else 0)


@pytest.mark.parametrize("method", ["sghmc", "adam_csghmc"])
def test_cfg3_cfg4_vit_l_32_training_steps(cuda_device, deterministic_fp32, tmp_path, method):
    """configs[2] / [3]: ViT-L/32 SGHMC with the net0 prior mean (methods/sghmc.py:482-510 + SGD.step) and Adam-cSGHMC
    (methods/adam_csghmc.py:814-861), batches of 8."""
    import importlib
    from oracle import make_golden_runner as mgr
    from bayesdll_b200 import shapes
    dev = cuda_device
    torch.manual_seed(7)
    net, net0 = shapes.create_backbone("vit_l_32", 37), shapes.create_backbone("vit_l_32", 37)
    ref, ref0 = copy.deepcopy(net).to(dev), copy.deepcopy(net0).to(dev)
    hp = dict(prior_sig=1.0, Ninflate=1e3, nd=1.0, momentum_decay=0.18 if method == "sghmc" else 0.05, burnin=5, thin=1,
              bias="informative", nst=5, noise="torch", beta1=0.9, beta2=0.999, epsilon=1e-8, temperature=1.0)
    args = mgr.make_args(hp, str(tmp_path), dev, momentum=0.5, epochs=4, num_cycles=2, lr=1e-4, lr_head=1e-2, ND=1840)
    args.num_classes = 37
    runner = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner(net, net0, args, _logger())
    crit = torch.nn.CrossEntropyLoss()
    named = list(ref.named_parameters())
    p0s = [p for _, p in ref0.named_parameters()]
    head = "heads.head"
    opt = er.make_sgd([p for n, p in named if head not in n], [p for n, p in named if head in n], args.lr, args.lr_head, 0.0)
    vs = {n: torch.zeros_like(p) for n, p in named}
    ms = {n: torch.zeros_like(p) for n, p in named}
    ss = {n: torch.zeros_like(p) for n, p in named}
    N = args.ND * runner.Ninflate
    alpha = float(hp["momentum_decay"])
    ref.train()
    runner.net.train()
    gen = torch.Generator().manual_seed(3)
    for t in range(1, 4):
        x = torch.randn(8, 3, 224, 224, generator=gen).to(dev)
        y = torch.randint(0, 37, (8,), generator=gen).to(dev)
        torch.manual_seed(50 + t)
        loss_a, _ = runner.model(x, y, runner.net, runner.net0, crit, [args.lr, args.lr_head], runner.Ninflate, runner.nd)
        loss = crit(ref(x), y)
        ref.zero_grad()
        loss.backward()
        torch.manual_seed(50 + t)
        xis = [torch.randn_like(p) for _, p in named]
        kw = dict(lr_body=args.lr, lr_head=args.lr_head, N=N, prior_sig=1.0, nd=runner.nd, alpha=alpha, bias="informative")
        if method == "sghmc":
            er.sghmc(named, p0s, xis, vs, head, **kw)
        else:
            er.adam(named, p0s, xis, vs, ms, ss, head, beta1=0.9, beta2=0.999, eps=1e-8, t=t, cyclical=True, temperature=1.0, **kw)
        opt.step()
        assert loss_a == loss.item()
        _assert_identical(t, runner, named, vs, *( (ms, ss) if method == "adam_csghmc" else ()))
    assert runner.model.chain.layout.n_dense == 305548325


def test_cfg5_resnet101_csgld_ensemble_and_calibration(cuda_device, deterministic_fp32, tmp_path):
    """configs[4]: ResNet-101 cSGLD posterior-predictive ensemble (cycles x nst samples, GMM mixture) + ECE/MCE/NLL.  The
    drop-in trains two short cycles, then ``evaluate()`` (noise=torch, seeded) is compared with the reference's evaluate
    statements (methods/csgld.py:333-456) fed with the drop-in's own cycle moments on the same GPU: every sampled
    network's logits are bit-identical (same draws, same forward), the mixture within 1e-5, and the calibration metrics
    of both agree."""
    from oracle import make_golden_runner as mgr
    from bayesdll_b200 import calibration
    from bayesdll_b200.methods import csgld
    dev = cuda_device
    net, net0 = mgr.cfg2_networks()
    hp = dict(prior_sig=1.0, Ninflate=1.0, nd=0.01, thin=1, bias="informative", nst=2, noise="torch")
    args = mgr.make_args(hp, str(tmp_path), dev, momentum=0.5, epochs=4, num_cycles=2, lr=1e-4, lr_head=1e-2, ND=1840)
    args.num_classes = 37
    runner = csgld.Runner(net, net0, args, _logger())
    train, _, test = mgr.cfg2_loaders()
    gen = torch.Generator().manual_seed(9)
    test = test + [(torch.randn(16, 3, 224, 224, generator=gen), torch.randint(0, 37, (16,), generator=gen))]
    torch.manual_seed(1)
    runner.train(train, None, test)
    assert sorted(runner.cycle_theta_mom1) == [1, 2] and all(v >= 2 for v in runner.samples_per_cycle.values())
    torch.manual_seed(123)
    loss, err, targets, logits, logits_all = runner.evaluate(test)
    weights = runner.calculate_gmm_weights()
    torch.manual_seed(123)
    ref_logits, ref_all, ref_targets = er.evaluate_cyclical_avg(runner.net, runner.cycle_theta_mom1, runner.cycle_theta_mom2,
                                                                runner.samples_per_cycle, weights, runner.nst, test, dev)
    assert logits_all.shape == tuple(ref_all.shape) == (32, 37, 2, 2)
    assert torch.equal(torch.from_numpy(logits_all), ref_all.cpu()), "a sampled network's logits differ"
    assert torch.equal(torch.from_numpy(targets), ref_targets.cpu())
    torch.testing.assert_close(torch.from_numpy(logits), ref_logits.float().cpu(), atol=1e-5, rtol=1e-5)
    ref_loss = torch.nn.functional.cross_entropy(ref_logits.float(), ref_targets).item()
    assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    assert err == ref_logits.argmax(1).ne(ref_targets).float().mean().item()
    a = calibration.analyze(targets, logits, 15, None)
    b = calibration.analyze(ref_targets.cpu().numpy(), ref_logits.float().cpu().numpy(), 15, None)
    assert all(abs(x - y) <= 1e-5 for x, y in zip(a, b))


def test_cfg5_full_sample_count_8_cycles_x_5_draws(cuda_device, deterministic_fp32, tmp_path):
    """configs[4] at its stated ensemble size: 8 cycles x nst 5 = 40 posterior samples on ResNet-101 (K = 37).  The cycle
    statistics are injected through the reference's attribute names (checkpoint layout), ``evaluate()`` (noise=torch,
    seeded) runs over two batches (16 + a ragged 5) and is compared with the reference's evaluate statements
    (methods/csgld.py:333-456) on the same GPU: all 40 sampled networks' logits bit-identical in the reference's
    [N, K, nst, C] layout, mixture within 1e-5, same error count, same calibration bins."""
    from oracle import make_golden_runner as mgr
    from bayesdll_b200 import calibration
    from bayesdll_b200.methods import csgld
    dev = cuda_device
    net, net0 = mgr.cfg2_networks()
    C, nst = 8, 5
    hp = dict(prior_sig=1.0, Ninflate=1.0, nd=0.01, thin=1, bias="informative", nst=nst, noise="torch")
    args = mgr.make_args(hp, str(tmp_path), dev, momentum=0.5, epochs=8, num_cycles=C, lr=1e-4, lr_head=1e-2, ND=1840)
    args.num_classes = 37
    runner = csgld.Runner(net, net0, args, _logger())
    gen = torch.Generator(device=dev).manual_seed(4)
    with torch.no_grad():
        theta = torch.nn.utils.parameters_to_vector(runner.net.parameters())
    m1, m2 = {}, {}
    for c in range(1, C + 1):
        mean = theta + 1e-3 * torch.randn(theta.numel(), device=dev, generator=gen)
        m1[c], m2[c] = mean, mean * mean + 1e-6 * torch.rand(theta.numel(), device=dev, generator=gen)
    runner.cycle_theta_mom1, runner.cycle_theta_mom2 = m1, m2
    runner.samples_per_cycle = {c: 3 + c for c in m1}
    runner.cycle_likelihoods = {c: [0.1 + 0.02 * c, 0.12 + 0.01 * c] for c in m1}
    runner.current_cycle = C
    g2 = torch.Generator().manual_seed(9)
    test = [(torch.randn(b, 3, 224, 224, generator=g2), torch.randint(0, 37, (b,), generator=g2)) for b in (16, 5)]
    torch.manual_seed(321)
    loss, err, targets, logits, logits_all = runner.evaluate(test)
    weights = runner.calculate_gmm_weights()
    assert len(weights) == C and abs(sum(weights.values()) - 1) < 1e-12
    torch.manual_seed(321)
    ref_logits, ref_all, ref_targets = er.evaluate_cyclical_avg(runner.net, runner.cycle_theta_mom1, runner.cycle_theta_mom2,
                                                                runner.samples_per_cycle, weights, nst, test, dev)
    assert logits_all.shape == tuple(ref_all.shape) == (21, 37, nst, C)
    assert torch.equal(torch.from_numpy(logits_all), ref_all.cpu()), "a sampled network's logits differ"
    assert torch.equal(torch.from_numpy(targets), ref_targets.cpu())
    torch.testing.assert_close(torch.from_numpy(logits), ref_logits.float().cpu(), atol=1e-5, rtol=1e-5)
    assert round(err * 21) == int(ref_logits.argmax(1).ne(ref_targets).sum().item()) and abs(err * 21 - round(err * 21)) < 1e-9
    _, _, _, _, sizes_a = calibration.calc_bins(targets, logits, 15)
    a = calibration.analyze(targets, logits, 15, None)
    b = calibration.analyze(ref_targets.cpu().numpy(), ref_logits.float().cpu().numpy(), 15, None)
    assert int(np.asarray(sizes_a).sum()) == 21 * 37 and all(abs(x - y) <= 1e-5 for x, y in zip(a, b))
