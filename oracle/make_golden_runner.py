"""Runner-level golden vectors: the reference's own ``Runner.train()`` executed end-to-end on CPU.

A tiny network (``InjectNet``) makes the run fully deterministic and device-independent:
  * in training, ``criterion(out, y) == out.sum()`` and ``out`` carries ``sum_t (p_t * G_t[step]).sum() / 16``
    broadcast over a [4,4] output, so after ``backward()`` every ``p.grad`` equals the injected ``G_t[step]``
    *exactly* (16 * 1/16 is exact in fp32);
  * in evaluation (no grad) ``out`` is a small detached MLP of all sampled parameters and the criterion is the
    ordinary cross entropy, so ``evaluate()`` / ``calibration.analyze`` see realistic logits;
  * every ``torch.randn_like`` (training noise and posterior draws) is served from a recorded tape.
The same ``InjectNet`` / criterion / tape are used by tests/test_runner_gpu.py to drive the drop-in Runner on the
GPU, which must reproduce every recorded quantity.

TEST INFRASTRUCTURE ONLY.  Run via ``python oracle/make_golden.py runner`` in the build container.
"""
import argparse
import logging
import os
import tempfile
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import refshim

K_CLASSES, BATCH, FEATS = 4, 4, 16
SHAPES = OrderedDict([
    ("layers.0.weight", (29, 11)), ("layers.0.bias", (29,)),
    ("layers.2.weight", (17, 29)), ("layers.2.bias", (17,)),
    ("norm.weight", (17,)), ("norm.bias", (17,)),
    ("classifier.weight", (K_CLASSES, 17)), ("classifier.bias", (K_CLASSES,)),
])


class InjectNet(nn.Module):
    readout_name = "classifier"

    def __init__(self, seed, G=None):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        for name, shape in SHAPES.items():
            parts = name.split(".")
            mod = self
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    mod.add_module(part, nn.Module())
                mod = getattr(mod, part)
            scale = 1.0 if name == "norm.weight" else 0.3
            mod.register_parameter(parts[-1], nn.Parameter(torch.randn(shape, generator=gen) * scale))
        self.register_buffer("G", torch.zeros(1, 1) if G is None else torch.as_tensor(G, dtype=torch.float32))
        self.train_calls = 0

    def _mlp(self, x):
        P = {n: p.detach() for n, p in self.named_parameters()}
        f = x.reshape(x.shape[0], -1)[:, :11]
        h = torch.tanh(f @ P["layers.0.weight"].t() + P["layers.0.bias"])
        h = torch.tanh(h @ P["layers.2.weight"].t() + P["layers.2.bias"])
        h = h * P["norm.weight"] + P["norm.bias"]
        return h @ P["classifier.weight"].t() + P["classifier.bias"]

    def forward(self, x):
        out = self._mlp(x)
        if self.training and torch.is_grad_enabled():
            g = self.G[self.train_calls]
            self.train_calls += 1
            tot, pos = 0.0, 0
            for p in self.parameters():
                tot = tot + (p * g[pos:pos + p.numel()].view(p.shape)).sum()
                pos += p.numel()
            assert out.shape == (BATCH, K_CLASSES)
            out = out + (tot * (1.0 / (BATCH * K_CLASSES))).expand(BATCH, K_CLASSES)
        return out


class InjectCriterion:
    """train: sum of outputs (gradient injection); eval: mean cross entropy, like torch.nn.CrossEntropyLoss()."""

    def __call__(self, out, y):
        if out.requires_grad:
            return out.sum()
        return F.cross_entropy(out, y)


class RealNet(nn.Module):
    """A real little conv net with BatchNorm (buffers are not sampled, Appendix B.12) trained through ordinary
    autograd -- no gradient injection.  Used for the end-to-end goldens ``runner_real_*``: the reference runs it on CPU,
    the drop-in on the GPU, so gradients differ at the 1e-7 level and results are compared with a tolerance."""
    readout_name = "classifier"

    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.features = nn.Sequential(nn.Conv2d(1, 6, 3, padding=1), nn.BatchNorm2d(6), nn.ReLU(), nn.MaxPool2d(2),
                                      nn.Conv2d(6, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU())
        self.classifier = nn.Linear(8 * 2 * 2, K_CLASSES)

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))


REAL_CASES = {
    "real_sghmc": ("sghmc", dict(prior_sig=1.0, Ninflate=10.0, nd=0.3, burnin=1, thin=1, nst=3, bias="informative",
                                 momentum_decay=0.18), dict(momentum=0.5, epochs=3)),
    "real_csghmc": ("csghmc", dict(prior_sig=0.05, Ninflate=5.0, nd=0.3, burnin=0, thin=1, nst=2, bias="informative",
                                   momentum_decay=0.18), dict(momentum=0.0, epochs=4, num_cycles=2)),
    "real_adam_csghmc": ("adam_csghmc", dict(prior_sig=1.0, Ninflate=10.0, nd=0.3, burnin=0, thin=1, nst=2,
                                             bias="uninformative", momentum_decay=0.1, beta1=0.9, beta2=0.99,
                                             epsilon=1e-3, temperature=1.0), dict(momentum=0.0, epochs=4, num_cycles=2)),
    # raw-sample store with BatchNorm buffers in every stored state_dict (methods/csghmc_fs.py)
    "real_csghmc_fs": ("csghmc_fs", dict(prior_sig=0.05, Ninflate=5.0, nd=0.3, burnin=0, thin=1, nst=2,
                                         bias="informative", momentum_decay=0.18), dict(momentum=0.0, epochs=6, num_cycles=2)),
}
REAL_SEEDS = {"real_adam_csghmc": 900, "real_csghmc": 901, "real_sghmc": 902, "real_csghmc_fs": 903}


def run_reference_real_case(name):
    method, hp, over = REAL_CASES[name]
    mod = refshim.load(f"methods.{method}")
    seed = REAL_SEEDS[name]
    rng = np.random.default_rng(seed)
    loaders = make_loaders(seed)
    tape = rng.standard_normal(200_000).astype(np.float32)
    net, net0 = RealNet(seed), RealNet(seed + 1)
    log_dir = tempfile.mkdtemp(prefix="bdl_golden_real_")
    args = make_args(hp, log_dir, torch.device("cpu"), lr=2e-2, lr_head=5e-2, **over)
    logger = logging.getLogger(f"golden.{name}")
    logger.addHandler(logging.NullHandler())
    logger.propagate = False
    runner = mod.Runner(net, net0, args, logger)
    evals = []
    orig_eval = runner.evaluate

    def recording_eval(loader):
        res = orig_eval(loader)
        evals.append(res)
        return res
    runner.evaluate = recording_eval
    bmas = []
    if hasattr(runner, "evaluate_full_samples"):
        orig_bma = runner.evaluate_full_samples

        def recording_bma(*a, **k):
            res = orig_bma(*a, **k)
            bmas.append((res, sorted(f for f in os.listdir(log_dir) if f.startswith("full_samples_net_ep"))))
            return res
        runner.evaluate_full_samples = recording_bma
    cwd = os.getcwd()
    os.chdir(log_dir)
    try:
        with refshim.injected_noise(tape) as tp:
            ret = runner.train(loaders[0], loaders[1], loaders[2])
            used = tp.pos
    finally:
        os.chdir(cwd)
    extra = {}
    if bmas:
        extra["n_bma"] = len(bmas)
        res, files = bmas[-1]
        extra["bma_files"] = np.array(files)
        for ds in ("train", "val", "test"):
            r = res[ds]
            for k in ("loss", "error", "num_models", "individual_avg_loss", "individual_avg_error"):
                extra[f"bma_{ds}_{k}"] = np.float64(r[k])
            extra[f"bma_{ds}_logits"], extra[f"bma_{ds}_logits_all"] = r["logits"], r["logits_all"]
        for i, f in enumerate(files):          # BatchNorm statistics differ from sample to sample
            sd = torch.load(os.path.join(log_dir, f), map_location="cpu")
            extra[f"fs{i}_bn_mean"] = sd["features.1.running_mean"].numpy()
            extra[f"fs{i}_theta"] = torch.cat([sd[n].reshape(-1) for n, _ in runner.net.named_parameters()]).numpy()
        extra["fs_state_keys"] = np.array(list(sd.keys()))
    rec = dict(tape=tape[:used], tape_used=used, n_evals=len(evals), method=np.array(method), **extra,
               theta_final=torch.cat([p.detach().reshape(-1) for p in runner.net.parameters()]).numpy(),
               bn_mean=runner.net.features[1].running_mean.numpy().copy(),
               bn_var=runner.net.features[1].running_var.numpy().copy(), **loaders_to_arrays(loaders))
    for i, (loss, err, targets, logits, logits_all) in enumerate(evals):
        rec[f"eval{i}_loss"], rec[f"eval{i}_err"] = loss, err
        rec[f"eval{i}_targets"], rec[f"eval{i}_logits"] = targets, logits
    if hasattr(runner, "post_theta_mom1"):
        rec["post_theta_mom1"] = runner.post_theta_mom1.numpy()
        rec["post_theta_mom2"] = runner.post_theta_mom2.numpy()
        rec["post_theta_cnt"] = runner.post_theta_cnt
    if hasattr(runner, "cycle_theta_mom1"):
        cyc = sorted(runner.cycle_theta_mom1)
        rec["cycles"] = np.array(cyc)
        for c in cyc:
            rec[f"cyc{c}_mom1"] = runner.cycle_theta_mom1[c].numpy()
            rec[f"cyc{c}_count"] = runner.samples_per_cycle[c]
            rec[f"cyc{c}_lik"] = np.asarray(runner.cycle_likelihoods[c], dtype=np.float64)
        rec["losses_train"] = ret["losses_train"]
    return rec


# BASELINE.json configs[0]: mlp_mnist SGLD, the reference's own CPU-runnable case (README.md:83), on synthetic 28x28 batches
# of 128.  burnin / thin / epochs are shortened (1 / 2 / 3 instead of 5 / 10 / 100) so the run takes seconds; everything
# else is the documented command line.  2.8 M parameters: the noise comes from a seeded generator (refshim.SeededTape) and
# the fixture keeps summaries (strided samples + fp64 sums) instead of whole vectors.
CFG1 = dict(hp=dict(prior_sig=1.0, Ninflate=1e3, nd=1.0, burnin=1, thin=2, bias="informative", nst=5),
            lr=1e-2, lr_head=1e-2, momentum=0.5, epochs=3, ND=30000, batch=128, n_train=4, n_eval=2, seed=4242, tape_seed=777)


def cfg1_loaders():
    rng = np.random.default_rng(CFG1["seed"])

    def mk(nb):
        return [(torch.from_numpy(rng.standard_normal((CFG1["batch"], 1, 28, 28)).astype(np.float32)),
                 torch.from_numpy(rng.integers(0, 10, CFG1["batch"]).astype(np.int64))) for _ in range(nb)]
    return mk(CFG1["n_train"]), mk(CFG1["n_eval"]), mk(CFG1["n_eval"])


def cfg1_network():
    """The reference's mlp_mnist (networks/small_nets.py:7-45: 784-1000-1000-1000-10) with a fixed CPU initialisation."""
    from bayesdll_b200 import shapes
    torch.manual_seed(CFG1["seed"])
    return shapes.create_backbone("mlp_mnist", 10)


def cfg1_args(log_dir, device, extra_hp=None):
    a = make_args(dict(CFG1["hp"], **(extra_hp or {})), log_dir, device, momentum=CFG1["momentum"], epochs=CFG1["epochs"],
                  lr=CFG1["lr"], lr_head=CFG1["lr_head"], ND=CFG1["ND"])
    a.num_classes = 10
    a.pretrained = None                        # zero prior mean, as the documented command line (no --pretrained)
    return a


def summarize(vec):
    v = np.asarray(vec, np.float32).reshape(-1)
    return dict(sample=v[::701].copy(), sum=np.float64(v.astype(np.float64).sum()),
                sumsq=np.float64((v.astype(np.float64) ** 2).sum()), n=v.size)


def run_reference_cfg1():
    mod = refshim.load("methods.sgld")
    loaders = cfg1_loaders()
    net = cfg1_network()
    log_dir = tempfile.mkdtemp(prefix="bdl_golden_cfg1_")
    args = cfg1_args(log_dir, torch.device("cpu"))
    logger = logging.getLogger("golden.cfg1")
    logger.addHandler(logging.NullHandler())
    logger.propagate = False
    runner = mod.Runner(net, None, args, logger)
    evals = []
    orig_eval = runner.evaluate

    def recording_eval(loader):
        res = orig_eval(loader)
        evals.append(res)
        return res
    runner.evaluate = recording_eval
    cwd = os.getcwd()
    os.chdir(log_dir)
    try:
        with refshim.injected_noise(CFG1["tape_seed"]) as tp:
            runner.train(*loaders)
            used, calls = tp.pos, tp.calls
    finally:
        os.chdir(cwd)
    rec = dict(tape_used=used, tape_calls=calls, n_evals=len(evals), post_theta_cnt=runner.post_theta_cnt)
    theta = torch.cat([p.detach().reshape(-1) for p in runner.net.parameters()]).numpy()
    for name, vec in (("theta", theta), ("mom1", runner.post_theta_mom1.numpy()), ("mom2", runner.post_theta_mom2.numpy())):
        for k, v in summarize(vec).items():
            rec[f"{name}_{k}"] = v
    for i, (loss, err, targets, logits, logits_all) in enumerate(evals):
        rec[f"eval{i}_loss"], rec[f"eval{i}_err"] = loss, err
        rec[f"eval{i}_targets"], rec[f"eval{i}_logits"] = targets, logits
    return rec


# BASELINE.json configs[1]: torchvision ResNet-101 (37 classes) cSGHMC with the cyclical step size, Pets-shaped synthetic
# 224x224 batches of 16, random-init net0 (pretrain_resnet101.py:127, README.md:111).  No CPU recording is kept for this
# case: a random-init ResNet-101 is numerically chaotic (plain PyTorch CPU vs GPU gradients of conv1 already differ by 6 %
# at the first step and 180 % at the second), so a cross-device trajectory comparison says nothing about the sampler.  The
# helpers below feed tests/test_backbone_parity_gpu.py, which runs the reference's statements and the drop-in side by
# side on the SAME device and requires bit-identical parameters and BatchNorm buffers.
CFG2 = dict(hp=dict(prior_sig=1.0, Ninflate=1.0, nd=0.01, burnin=0, momentum_decay=0.18, thin=1, bias="informative", nst=1),
            lr=1e-4, lr_head=1e-2, momentum=0.0, epochs=2, num_cycles=2, ND=1840, batch=16, n_train=3, n_eval=1, seed=2121,
            tape_seed=888)


def cfg2_loaders():
    rng = np.random.default_rng(CFG2["seed"])

    def mk(nb):
        return [(torch.from_numpy(rng.standard_normal((CFG2["batch"], 3, 224, 224)).astype(np.float32)),
                 torch.from_numpy(rng.integers(0, 37, CFG2["batch"]).astype(np.int64))) for _ in range(nb)]
    return mk(CFG2["n_train"]), mk(CFG2["n_eval"]), mk(CFG2["n_eval"])


def cfg2_networks():
    from bayesdll_b200 import shapes
    torch.manual_seed(CFG2["seed"])
    net = shapes.create_backbone("resnet101", 37)
    net0 = shapes.create_backbone("resnet101", 37)
    return net, net0


def cfg2_args(log_dir, device, extra_hp=None):
    a = make_args(dict(CFG2["hp"], **(extra_hp or {})), log_dir, device, momentum=CFG2["momentum"], epochs=CFG2["epochs"],
                  num_cycles=CFG2["num_cycles"], lr=CFG2["lr"], lr_head=CFG2["lr_head"], ND=CFG2["ND"])
    a.num_classes = 37
    return a


def make_loaders(seed, n_train=3, n_val=2, n_test=2):
    rng = np.random.default_rng(seed)

    def mk(nb):
        return [(torch.from_numpy(rng.standard_normal((BATCH, 1, 4, 4)).astype(np.float32)),
                 torch.from_numpy(rng.integers(0, K_CLASSES, BATCH).astype(np.int64))) for _ in range(nb)]
    return mk(n_train), mk(n_val), mk(n_test)


def loaders_to_arrays(loaders):
    out = {}
    for nm, ld in zip(("train", "val", "test"), loaders):
        out[f"x_{nm}"] = np.stack([x.numpy() for x, _ in ld])
        out[f"y_{nm}"] = np.stack([y.numpy() for _, y in ld])
    return out


def loaders_from_arrays(z, device=None):
    res = []
    for nm in ("train", "val", "test"):
        res.append([(torch.from_numpy(x), torch.from_numpy(y)) for x, y in zip(z[f"x_{nm}"], z[f"y_{nm}"])])
    return res


CASES = {
    # name: (method, hparams, args overrides)
    "sgld": ("sgld", dict(prior_sig=0.8, Ninflate=5.0, nd=0.6, burnin=1, thin=2, nst=3, bias="informative"),
             dict(momentum=0.5, epochs=3)),
    "sghmc": ("sghmc", dict(prior_sig=0.8, Ninflate=5.0, nd=0.6, burnin=1, thin=1, nst=3, bias="uninformative",
                            momentum_decay=0.18), dict(momentum=0.9, epochs=3)),
    "adam_sghmc": ("adam_sghmc", dict(prior_sig=0.8, Ninflate=5.0, nd=0.6, burnin=1, thin=2, nst=2, bias="informative",
                                      momentum_decay=0.1, beta1=0.9, beta2=0.99, epsilon=1e-6),
                   dict(momentum=0.5, epochs=3)),
    "sgld_nst0": ("sgld", dict(prior_sig=0.8, Ninflate=5.0, nd=0.6, burnin=1, thin=2, nst=0, bias="informative"),
                  dict(momentum=0.0, epochs=2)),
    "csgld": ("csgld", dict(prior_sig=0.8, Ninflate=5.0, nd=0.6, thin=1, nst=2, bias="informative"),
              dict(momentum=0.5, epochs=4, num_cycles=2)),
    "csghmc": ("csghmc", dict(prior_sig=0.05, Ninflate=5.0, nd=0.6, burnin=0, thin=1, nst=2, bias="informative",
                              momentum_decay=0.18), dict(momentum=0.0, epochs=4, num_cycles=2)),
    "adam_csghmc": ("adam_csghmc", dict(prior_sig=0.8, Ninflate=5.0, nd=0.6, burnin=0, thin=1, nst=2,
                                        bias="uninformative", momentum_decay=0.1, beta1=0.9, beta2=0.99, epsilon=1e-6,
                                        temperature=1.5), dict(momentum=0.0, epochs=4, num_cycles=2)),
}


def make_args(hparams, log_dir, device, *, momentum, epochs, num_cycles=2, lr=5e-3, lr_head=2e-2, ND=12):
    a = argparse.Namespace()
    a.device = device
    a.ND = ND
    a.lr, a.lr_head, a.momentum, a.epochs = lr, lr_head, momentum, epochs
    a.pretrained = "synthetic"
    a.hparams = {k: str(v) for k, v in hparams.items()}
    a.num_cycles = num_cycles
    a.proportion_exploration = 0.5
    a.full_sample = False
    a.test_eval_freq = 1
    a.ece_num_bins = 15
    a.num_classes = K_CLASSES
    a.log_dir = log_dir
    a.seed = 1234
    return a


def n_params():
    return sum(int(np.prod(s)) for s in SHAPES.values())


# Full-sample store + Bayesian model average (methods/csghmc_fs.py).  L = epochs // num_cycles = 3, so epochs 0, 1, 3, 4
# dump a state_dict and each dump is followed by a BMA over every file written so far (csghmc_fs.py:176-181).
FS_CASES = {
    "csghmc_fs": ("csghmc_fs", dict(prior_sig=0.05, Ninflate=5.0, nd=0.6, burnin=0, thin=1, nst=2, bias="informative",
                                    momentum_decay=0.18), dict(momentum=0.0, epochs=6, num_cycles=2)),
}
FS_SEEDS = {"csghmc_fs": 700}


def case_seed(name):
    return FS_SEEDS[name] if name in FS_SEEDS else 500 + sorted(CASES).index(name)


def case_spec(name):
    return FS_CASES[name] if name in FS_CASES else CASES[name]


def run_reference_case(name):
    method, hp, over = case_spec(name)
    mod = refshim.load(f"methods.{method}")
    seed = case_seed(name)
    rng = np.random.default_rng(seed)
    loaders = make_loaders(seed)
    steps = over["epochs"] * len(loaders[0])
    G = (rng.standard_normal((steps, n_params())) * 0.05).astype(np.float32)
    tape = rng.standard_normal(400_000).astype(np.float32)
    net = InjectNet(seed, G)
    net0 = InjectNet(seed + 1)
    theta_init = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy().copy()
    theta0 = torch.cat([p.detach().reshape(-1) for p in net0.parameters()]).numpy().copy()
    log_dir = tempfile.mkdtemp(prefix="bdl_golden_runner_")
    args = make_args(hp, log_dir, torch.device("cpu"), **over)
    logger = logging.getLogger(f"golden.{name}")
    logger.addHandler(logging.NullHandler())
    logger.propagate = False
    runner = mod.Runner(net, net0, args, logger)
    runner.criterion = InjectCriterion()

    evals = []
    orig_eval = runner.evaluate

    def recording_eval(loader):
        res = orig_eval(loader)
        evals.append(res)
        return res
    runner.evaluate = recording_eval
    bmas = []
    if hasattr(runner, "evaluate_full_samples"):
        orig_bma = runner.evaluate_full_samples

        def recording_bma(*a, **k):
            res = orig_bma(*a, **k)
            bmas.append((res, sorted(f for f in os.listdir(log_dir) if f.startswith("full_samples_net_ep"))))
            return res
        runner.evaluate_full_samples = recording_bma

    cwd = os.getcwd()
    os.chdir(log_dir)
    try:
        with refshim.injected_noise(tape) as tp:
            ret = runner.train(loaders[0], loaders[1], loaders[2])
            used = tp.pos
    finally:
        os.chdir(cwd)

    rec = dict(G=G, tape=tape[:used], tape_used=used, theta_init=theta_init, theta0=theta0,
               theta_final=torch.cat([p.detach().reshape(-1) for p in runner.net.parameters()]).numpy(),
               n_evals=len(evals), **loaders_to_arrays(loaders))
    for i, (loss, err, targets, logits, logits_all) in enumerate(evals):
        rec[f"eval{i}_loss"], rec[f"eval{i}_err"] = loss, err
        rec[f"eval{i}_targets"], rec[f"eval{i}_logits"], rec[f"eval{i}_logits_all"] = targets, logits, logits_all
    if bmas:
        rec["n_bma"] = len(bmas)
        for i, (res, files) in enumerate(bmas):
            rec[f"bma{i}_files"] = np.array(files)
            for ds in ("train", "val", "test"):
                r = res[ds]
                for k in ("loss", "error", "num_models", "individual_avg_loss", "individual_avg_error"):
                    rec[f"bma{i}_{ds}_{k}"] = np.float64(r[k])
                if i == len(bmas) - 1 or ds == "test":
                    rec[f"bma{i}_{ds}_targets"], rec[f"bma{i}_{ds}_logits"] = r["targets"], r["logits"]
                    rec[f"bma{i}_{ds}_logits_all"] = r["logits_all"]
        # on-disk contract of the sample store and of the BMA outputs
        sd = torch.load(os.path.join(log_dir, bmas[-1][1][-1]), map_location="cpu")
        rec["fs_state_keys"] = np.array(list(sd.keys()))
        rec["fs_last_sample"] = torch.cat([sd[n].reshape(-1) for n, _ in runner.net.named_parameters()]).numpy()
        import pickle
        with open(os.path.join(log_dir, "bma_evaluation_results.pkl"), "rb") as f:
            pk = pickle.load(f)
        rec["bma_pkl_keys"] = np.array(sorted(pk["test"].keys()))
        rec["bma_files_in_logdir"] = np.array(sorted(f for f in os.listdir(log_dir) if "bma" in f))
    if hasattr(runner, "post_theta_mom1"):
        rec["post_theta_mom1"] = runner.post_theta_mom1.numpy()
        if hasattr(runner, "post_theta_mom2"):
            rec["post_theta_mom2"] = runner.post_theta_mom2.numpy()
        rec["post_theta_cnt"] = runner.post_theta_cnt
    if hasattr(runner, "cycle_theta_mom1"):
        cyc = sorted(runner.cycle_theta_mom1)
        rec["cycles"] = np.array(cyc)
        for c in cyc:
            rec[f"cyc{c}_mom1"] = runner.cycle_theta_mom1[c].numpy()
            rec[f"cyc{c}_mom2"] = runner.cycle_theta_mom2[c].numpy()
            rec[f"cyc{c}_count"] = runner.samples_per_cycle[c]
            rec[f"cyc{c}_lik"] = np.asarray(runner.cycle_likelihoods[c], dtype=np.float64)
        rec["samples_collected"] = runner.samples_collected
        rec["losses_train"] = ret["losses_train"]
        rec["losses_test"] = ret["losses_test"]
    if hasattr(runner.model, "momentum_buffer"):
        names = [n for n, _ in runner.net.named_parameters()]
        rec["v_final"] = torch.cat([runner.model.momentum_buffer[n].reshape(-1) for n in names]).numpy()
    # checkpoint key inventory (on-disk contract)
    ck_files = sorted(f for f in os.listdir(log_dir) if f.endswith("ckpt.pt"))
    rec["ckpt_files"] = np.array(ck_files)
    if ck_files:
        ck = torch.load(os.path.join(log_dir, ck_files[-1]), map_location="cpu", weights_only=False)
        rec["ckpt_keys"] = np.array(sorted(ck.keys()))
        lt = ck.get("last_theta")
        rec["ckpt_last_theta_kind"] = np.array("none" if lt is None else ("vector" if torch.is_tensor(lt) else "state_dict"))
    rec["hp_keys"] = np.array(sorted(hp))
    rec["hp_vals"] = np.array([str(hp[k]) for k in sorted(hp)])
    rec["method"] = np.array(method)
    return rec


def main(save, only=None):
    torch.set_num_threads(1)
    if only == "cfg1":
        rec = run_reference_cfg1()
        save("runner_cfg1_mlp_sgld", **rec)
        print(f"  cfg1 mlp_mnist SGLD: {rec['tape_calls']} noise calls / {rec['tape_used']} draws, evaluate() calls {rec['n_evals']}")
        return
    if only is not None:
        for name in only:
            rec = run_reference_real_case(name) if name in REAL_CASES else run_reference_case(name)
            save(f"runner_{name}", **rec)
            print(f"  {name}: tape used {rec['tape_used']}, evaluate() calls {rec['n_evals']}, BMA calls {rec.get('n_bma', 0)}")
        return
    for name in CASES:
        rec = run_reference_case(name)
        save(f"runner_{name}", **rec)
        print(f"  {name}: tape used {rec['tape_used']}, evaluate() calls {rec['n_evals']}")
    for name in FS_CASES:
        rec = run_reference_case(name)
        save(f"runner_{name}", **rec)
        print(f"  {name}: tape used {rec['tape_used']}, evaluate() calls {rec['n_evals']}, BMA calls {rec['n_bma']}")
    for name in REAL_CASES:
        rec = run_reference_real_case(name)
        save(f"runner_{name}", **rec)
        print(f"  {name}: tape used {rec['tape_used']}, evaluate() calls {rec['n_evals']}")
