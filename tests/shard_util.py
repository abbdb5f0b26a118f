"""Shared by tests/test_eval_shard_gpu.py and its spawned ranks: a small drop-in Runner with synthetic posterior
statistics injected through the reference's own attribute names (cycle_theta_mom1/2, samples_per_cycle,
cycle_likelihoods / post_theta_mom1/2, post_theta_cnt), and a deterministic loader with a ragged last batch."""
import argparse
import importlib
import logging

import numpy as np
import torch

K = 7


class ConvNet(torch.nn.Module):
    """conv + BatchNorm + linear head: parameters, buffers that are not sampled (Appendix B.12) and a readout."""
    readout_name = "classifier"

    def __init__(self):
        super().__init__()
        self.features = torch.nn.Sequential(torch.nn.Conv2d(3, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                                            torch.nn.AdaptiveAvgPool2d(4))
        self.classifier = torch.nn.Linear(6 * 16, K)

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))


def make_loader(batches=(16, 16, 5), seed=3):
    gen = torch.Generator().manual_seed(seed)
    return [(torch.randn(b, 3, 8, 8, generator=gen), torch.randint(0, K, (b,), generator=gen)) for b in batches]


def make_runner(method, device, log_dir, nst, eval_shard, cycles=3, stat_seed=5, init_seed=11):
    torch.manual_seed(init_seed)
    net, net0 = ConvNet(), ConvNet()
    with torch.no_grad():                                 # non-trivial BatchNorm statistics
        bn = net.features[1]
        bn.running_mean.copy_(torch.linspace(-0.2, 0.3, 6))
        bn.running_var.copy_(torch.linspace(0.5, 1.5, 6))
    hp = dict(prior_sig=1.0, Ninflate=10.0, nd=1.0, burnin=0, thin=1, nst=nst, bias="informative", momentum_decay=0.18,
              beta1=0.9, beta2=0.999, epsilon=1e-8, temperature=1.0, seed=77, eval_shard=int(eval_shard))
    a = argparse.Namespace(device=device, ND=64, lr=1e-3, lr_head=1e-2, momentum=0.5, epochs=4, pretrained="synthetic",
                           hparams={k: str(v) for k, v in hp.items()}, num_cycles=2, proportion_exploration=0.5,
                           full_sample=False, test_eval_freq=1, ece_num_bins=15, num_classes=K, log_dir=str(log_dir), seed=77)
    lg = logging.getLogger("shard_util")
    lg.addHandler(logging.NullHandler())
    lg.propagate = False
    runner = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner(net, net0, a, lg)
    n = sum(p.numel() for p in runner.net.parameters())
    gen = torch.Generator().manual_seed(stat_seed)
    theta = torch.cat([p.detach().reshape(-1).cpu() for p in runner.net.parameters()])
    if hasattr(runner, "cycle_theta_mom1"):
        m1, m2 = {}, {}
        for c in range(1, cycles + 1):
            mean = theta + 0.05 * torch.randn(n, generator=gen)
            m1[c] = mean
            m2[c] = mean * mean + 1e-3 * torch.rand(n, generator=gen)      # 'avg': second moment; Welford: M2
        runner.cycle_theta_mom1, runner.cycle_theta_mom2 = m1, m2
        runner.samples_per_cycle = {c: 4 + c for c in m1}
        runner.cycle_likelihoods = {c: [0.2 + 0.1 * c, 0.25 + 0.05 * c] for c in m1}
        runner.current_cycle = cycles
    else:
        mean = theta + 0.05 * torch.randn(n, generator=gen)
        runner.post_theta_mom1 = mean
        runner.post_theta_mom2 = mean * mean + 1e-3 * torch.rand(n, generator=gen)
        runner.post_theta_cnt = 6
    return runner


def run_case(method, device, log_dir, nst, eval_shard, cycles=3, runner=None):
    """evaluate() twice (the second call uses the captured CUDA graphs and another Philox sub-sequence), the calibration
    bins of the result, and -- cyclical runners -- full_batch_likelihoods()."""
    from bayesdll_b200 import calibration
    if runner is None:
        runner = make_runner(method, device, log_dir, nst, eval_shard, cycles=cycles)
    loader = make_loader()
    out = {}
    for i in range(2):
        loss, err, targets, logits, logits_all = runner.evaluate(loader)
        binned, bins, accs, confs, sizes = calibration.calc_bins(targets, logits, 15)
        out[f"eval{i}"] = dict(loss=loss, err=err, targets=targets, logits=logits, logits_all=logits_all,
                               binned=np.asarray(binned), sizes=np.asarray(sizes), accs=np.asarray(accs), confs=np.asarray(confs),
                               analyze=calibration.analyze(targets, logits, 15, None))
    if hasattr(runner, "full_batch_likelihoods"):
        out["likelihoods"] = np.asarray(runner.full_batch_likelihoods(make_loader(seed=9)))
    runner.flush_io()
    return out
