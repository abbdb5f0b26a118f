"""Generate golden vectors by executing the REAL reference (read-only /root/reference).

Run in the build container only:   python oracle/make_golden.py
Writes tests/golden/*.npz (committed).  Every file carries the library versions used.

What is recorded:
  step_<method>_<case>.npz   per-step trajectories of the reference's own
      ``Model.forward`` + ``optimizer.step()`` with injected gradients (GradInjectNet)
      and injected noise (NoiseTape replacing torch.randn_like).
  cyclical.npz               methods/cyclical.py schedule values on several (epochs, B, M).
  calibration.npz            calibration.calc_bins / analyze on seeded logits.
  runner_<method>.npz        whole ``Runner.train()`` runs on a tiny synthetic problem
      (moments, counts, evaluate() outputs, ECE/MCE/NLL) -- see make_runner_goldens().
"""
import argparse
import logging
import os
import sys
import tempfile
from collections import OrderedDict

import numpy as np
import scipy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import refshim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

VERSIONS = dict(torch=torch.__version__, numpy=np.__version__, scipy=scipy.__version__)

# tiny "network": odd sizes so 16-byte padding is exercised, bias-named and head-named tensors
SHAPES = OrderedDict([
    ("layers.0.weight", (29, 11)),
    ("layers.0.bias", (29,)),
    ("layers.2.weight", (17, 29)),
    ("layers.2.bias", (17,)),
    ("norm.weight", (17,)),
    ("norm.bias", (17,)),
    ("classifier.weight", (5, 17)),
    ("classifier.bias", (5,)),
])
READOUT = "classifier"


def flat(params):
    return torch.cat([p.detach().reshape(-1) for p in params]).numpy().copy()


def flat_dict(d, names):
    return torch.cat([d[n].detach().reshape(-1) for n in names]).numpy().copy()


def make_args(method_hparams, *, lr, lr_head, momentum, ND, epochs=4, pretrained="synthetic",
              num_cycles=2, proportion_exploration=0.5, log_dir=None):
    a = argparse.Namespace()
    a.device = torch.device("cpu")
    a.ND = ND
    a.lr, a.lr_head, a.momentum = lr, lr_head, momentum
    a.epochs = epochs
    a.pretrained = pretrained
    a.hparams = {k: str(v) for k, v in method_hparams.items()}
    a.num_cycles = num_cycles
    a.proportion_exploration = proportion_exploration
    a.full_sample = False
    a.test_eval_freq = 1
    a.ece_num_bins = 15
    a.num_classes = 4
    a.log_dir = log_dir or tempfile.mkdtemp(prefix="bdl_golden_")
    return a


def quiet_logger():
    lg = logging.getLogger("golden")
    lg.addHandler(logging.NullHandler())
    lg.setLevel(logging.CRITICAL)
    return lg


def sgd_buf_flat(opt, net, names):
    bufs = {}
    for (n, p) in net.named_parameters():
        st = opt.state.get(p, {})
        b = st.get("momentum_buffer", None)
        bufs[n] = b if b is not None else torch.zeros_like(p)
    return flat_dict(bufs, names)


def run_step_case(method, hparams, *, lr, lr_head, momentum, ND, T, seed, sample_pattern=None,
                  cyc_lrs=None, clip_grad=None):
    """Drive the reference Model.forward (+ optimizer.step()) exactly like train_one_epoch does
    (methods/sghmc.py:220-229, methods/csghmc.py:285-304) and record the state after every step."""
    mod = refshim.load(f"methods.{method}")
    rng = np.random.default_rng(seed)
    net = refshim.GradInjectNet(SHAPES, READOUT, init_std=0.1, seed=seed)
    net0 = refshim.GradInjectNet(SHAPES, READOUT, init_std=0.1, seed=seed + 1000)
    names = [n for n, _ in net.named_parameters()]
    n = sum(p.numel() for p in net.parameters())
    args = make_args(hparams, lr=lr, lr_head=lr_head, momentum=momentum, ND=ND)
    runner = mod.Runner(net, net0, args, quiet_logger())
    model, opt = runner.model, runner.optimizer

    rec = dict(theta_init=flat(net.parameters()), theta0=flat(runner.net0.parameters()))
    G = (rng.standard_normal((T, n)) * 0.05).astype(np.float32)
    XI = rng.standard_normal((T, n)).astype(np.float32)
    out = {k: [] for k in ("theta", "v", "m", "s", "buf", "lr_body", "lr_head", "total_norm", "pgrad")}
    for t in range(T):
        if cyc_lrs is not None:                       # cyclical runners overwrite param_group lrs each step
            cur = cyc_lrs[t]
            opt.param_groups[0]["lr"] = cur
            opt.param_groups[1]["lr"] = cur * (args.lr_head / args.lr)
        lrs = [pg["lr"] for pg in opt.param_groups]
        # split the flat injected gradient per tensor
        gs, pos = [], 0
        for p in net.parameters():
            gs.append(torch.from_numpy(G[t, pos:pos + p.numel()]).reshape(p.shape))
            pos += p.numel()
        net.set_grads(gs)
        with refshim.injected_noise(XI[t]) as tape:
            if method == "csghmc":
                model(None, None, runner.net, runner.net0, refshim.identity_criterion, lrs,
                      runner.Ninflate, runner.nd, should_sample=bool(sample_pattern[t]))
            else:
                model(None, None, runner.net, runner.net0, refshim.identity_criterion, lrs,
                      runner.Ninflate, runner.nd)
                if clip_grad is not None:             # the statement of methods/csgld.py:250-251 / adam_csghmc.py:319-320
                    out["pgrad"].append(flat([p.grad for p in net.parameters()]))
                    tn = torch.nn.utils.clip_grad_norm_(runner.net.parameters(), clip_grad)
                    out["total_norm"].append(np.float32(tn.item()))
                opt.step()
        assert tape.calls == len(names) and tape.pos == n
        out["theta"].append(flat(net.parameters()))
        out["lr_body"].append(lrs[0])
        out["lr_head"].append(lrs[1])
        if hasattr(model, "momentum_buffer"):
            out["v"].append(flat_dict(model.momentum_buffer, names))
        if hasattr(model, "m"):
            out["m"].append(flat_dict(model.m, names))
            out["s"].append(flat_dict(model.v, names))
        out["buf"].append(sgd_buf_flat(opt, net, names))
    rec.update(G=G, XI=XI)
    for k, v in out.items():
        if v:
            rec[k] = np.stack(v)
    return rec, names


def meta_arrays(names):
    sizes = [int(np.prod(SHAPES[n])) for n in names]
    is_head = np.concatenate([np.full(s, READOUT in n, dtype=bool) for n, s in zip(names, sizes)])
    is_bias = np.concatenate([np.full(s, "bias" in n, dtype=bool) for n, s in zip(names, sizes)])
    return np.array(sizes), is_head, is_bias


def save(name, **arrays):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    arrays["versions"] = np.array([f"{k}={v}" for k, v in VERSIONS.items()])
    np.savez_compressed(path, **arrays)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def make_step_goldens():
    T = 5
    base = dict(prior_sig=0.7, Ninflate=10.0, nd=0.5, burnin=1, thin=2, nst=3)
    cases = []
    for bias in ("informative", "uninformative"):
        for mu in (0.5, 0.0):
            cases.append(("sgld", f"{bias[:3]}_mu{mu}", dict(base, bias=bias), dict(momentum=mu)))
        cases.append(("sghmc", f"{bias[:3]}", dict(base, bias=bias, momentum_decay=0.18), dict(momentum=0.9)))
        for mu in (0.5, 0.0):
            cases.append(("adam_sghmc", f"{bias[:3]}_mu{mu}",
                          dict(base, bias=bias, momentum_decay=0.05, beta1=0.9, beta2=0.999, epsilon=1e-8),
                          dict(momentum=mu)))
        cases.append(("adam_csghmc", f"{bias[:3]}",
                      dict(base, bias=bias, momentum_decay=0.05, beta1=0.8, beta2=0.99, epsilon=1e-6,
                           temperature=1.7), dict(momentum=0.9)))
    cases.append(("csgld", "inf_mu0.5", dict(base, bias="informative"), dict(momentum=0.5)))
    cases.append(("csghmc", "inf", dict(base, bias="informative", momentum_decay=0.18), dict(momentum=0.0)))
    cases.append(("csghmc", "uni", dict(base, bias="uninformative", momentum_decay=0.18), dict(momentum=0.0)))

    cyc = refshim.load("methods.cyclical")
    for i, (method, tag, hp, kw) in enumerate(cases):
        lr, lr_head, ND = 1e-2, 3e-2, 500
        extra = {}
        if method in ("csgld", "csghmc", "adam_csghmc"):
            sched = cyc.CyclicalSGMCMC(base_lr=lr, nbr_of_cycles=2, epochs=2, proportion_exploration=0.5)
            extra["cyc_lrs"] = [sched.calculate_lr(epoch=0, batch=b, batches_per_epoch=5) for b in range(T)]
        if method == "csghmc":
            extra["sample_pattern"] = [0, 1, 1, 0, 1]
        rec, names = run_step_case(method, hp, lr=lr, lr_head=lr_head, ND=ND, T=T, seed=100 + i,
                                   **kw, **extra)
        sizes, is_head, is_bias = meta_arrays(names)
        hp_keys = sorted(hp)
        save(f"step_{method}_{tag}", names=np.array(names), sizes=sizes, is_head=is_head, is_bias=is_bias,
             hp_keys=np.array(hp_keys), hp_vals=np.array([str(hp[k]) for k in hp_keys]),
             lr=lr, lr_head_arg=lr_head, ND=ND, momentum=kw["momentum"],
             sample_pattern=np.array(extra.get("sample_pattern", []), dtype=np.int64), **rec)


def make_clip_goldens():
    """args.clip_grad: no driver of the reference defines it, the cyclical runners consult it (csgld.py:250, adam_csghmc.py:319).
    The clip values sit inside the range of the recorded norms, so both regimes (coef < 1, coef clamped to 1) occur."""
    T = 6
    base = dict(prior_sig=0.7, Ninflate=10.0, nd=0.5, burnin=1, thin=2, nst=3)
    cases = [("csgld", "clip_inf_mu0.5", dict(base, bias="informative"), dict(momentum=0.5), 4.6),
             ("csgld", "clip_uni_mu0.0", dict(base, bias="uninformative"), dict(momentum=0.0), 4.6),
             ("adam_csghmc", "clip_inf", dict(base, bias="informative", momentum_decay=0.05, beta1=0.8, beta2=0.99, epsilon=1e-6,
                                              temperature=1.7), dict(momentum=0.9), 1.3)]
    cyc = refshim.load("methods.cyclical")
    for i, (method, tag, hp, kw, clip) in enumerate(cases):
        lr, lr_head, ND = 1e-2, 3e-2, 500
        sched = cyc.CyclicalSGMCMC(base_lr=lr, nbr_of_cycles=2, epochs=2, proportion_exploration=0.5)
        cyc_lrs = [sched.calculate_lr(epoch=0, batch=b, batches_per_epoch=6) for b in range(T)]
        rec, names = run_step_case(method, hp, lr=lr, lr_head=lr_head, ND=ND, T=T, seed=300 + i, cyc_lrs=cyc_lrs,
                                   clip_grad=clip, **kw)
        sizes, is_head, is_bias = meta_arrays(names)
        hp_keys = sorted(hp)
        print(method, tag, "total norms", rec["total_norm"], "clip", clip)
        save(f"step_{method}_{tag}", names=np.array(names), sizes=sizes, is_head=is_head, is_bias=is_bias,
             hp_keys=np.array(hp_keys), hp_vals=np.array([str(hp[k]) for k in hp_keys]),
             lr=lr, lr_head_arg=lr_head, ND=ND, momentum=kw["momentum"], clip_grad=clip,
             sample_pattern=np.array([], dtype=np.int64), **rec)


def make_cyclical_golden():
    cyc = refshim.load("methods.cyclical")
    rows = []
    for (epochs, B, M, beta, lr0) in [(4, 5, 2, 0.5, 1e-2), (40, 115, 8, 0.5, 1e-4), (100, 115, 8, 0.8, 1e-4),
                                      (7, 13, 3, 0.25, 0.1), (10, 9, 4, 0.5, 1e-3)]:
        s = cyc.CyclicalSGMCMC(base_lr=lr0, nbr_of_cycles=M, epochs=epochs, proportion_exploration=beta)
        for ep in range(epochs):
            for b in range(B):
                if epochs * B > 2000 and (ep * B + b) % 7 not in (0, 3) and (ep * B + b + 1) % (epochs * B // M) > 2:
                    continue
                kw = dict(epoch=ep, batch=b, batches_per_epoch=B)
                rows.append((epochs, B, M, beta, lr0, ep, b, s.calculate_lr(**kw), float(s.should_sample(**kw)),
                             float(s.last_in_cycle(**kw)), s.get_cycle_number(**kw)))
    save("cyclical", rows=np.array(rows, dtype=np.float64),
         cols=np.array("epochs B M beta lr0 ep b lr should_sample last_in_cycle cycle".split()))


def make_calibration_golden():
    cal = refshim.load("calibration")
    rng = np.random.default_rng(7)
    out = {}
    for tag, (N, K, M, scale) in dict(a=(257, 10, 15, 3.0), b=(1000, 37, 15, 1.0), c=(64, 3, 10, 8.0),
                                      d=(3669, 37, 15, 2.5)).items():
        logits = (rng.standard_normal((N, K)) * scale).astype(np.float32)
        labels = rng.integers(0, K, size=N).astype(np.int64)
        # make ~60% of the labels agree with the argmax so acc/conf are not degenerate
        agree = rng.random(N) < 0.6
        labels[agree] = logits[agree].argmax(1)
        for tname, T in (("T1", 1), ("Tarr", np.array([1.37]))):
            bins, binned, accs, confs, sizes = cal.calc_bins(labels, logits, M, T)
            ece, mce, nll = cal.analyze(labels, logits, M, os.path.join(tempfile.gettempdir(), "x.png"), T)
            pre = f"{tag}_{tname}_"
            out.update({pre + "bins": bins, pre + "binned": binned, pre + "accs": accs, pre + "confs": confs,
                        pre + "sizes": sizes, pre + "ece": ece, pre + "mce": mce, pre + "nll": nll,
                        pre + "T": np.asarray(T, dtype=np.float64)})
        out[tag + "_logits"] = logits
        out[tag + "_labels"] = labels
        out[tag + "_M"] = M
    save("calibration", **out)


def make_attribute_inventory():
    """Names the reference's Runner / Model instances carry after __init__ (``self.<name> = ...`` statements) and the
    public methods of both classes, per method file: the duck-typed surface other code may read."""
    import ast
    import json
    ref = refshim.find_reference()
    inv = {}
    for m in ("sgld", "sghmc", "csgld", "csghmc", "csghmc_fs", "adam_sghmc", "adam_csghmc"):
        tree = ast.parse(open(os.path.join(ref, "methods", f"{m}.py")).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in ("Runner", "Model"):
                methods = [n.name for n in node.body if isinstance(n, ast.FunctionDef)]
                init = next(n for n in node.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
                attrs = sorted({t.attr for st in ast.walk(init) if isinstance(st, (ast.Assign, ast.AugAssign, ast.AnnAssign))
                                for t in ast.walk(st) if isinstance(t, ast.Attribute) and isinstance(t.ctx, ast.Store)
                                and isinstance(t.value, ast.Name) and t.value.id == "self"})
                inv[f"{m}.{node.name}"] = {"methods": methods, "init_attributes": attrs}
    path = os.path.join(GOLDEN_DIR, "api_inventory.json")
    with open(path, "w") as f:
        json.dump(inv, f, indent=1, sort_keys=True)
    print(f"wrote {path}")


def make_reparam_draw_goldens():
    """The per-step reparameterisation draws of the reference's own ``vi.Model.forward`` and ``mc_dropout.Model.forward``
    (eval_grad=0), with ``torch.randn_like`` / ``torch.rand_like`` served from recorded tapes: the workhorse network's
    parameters after the call are the draw."""
    import contextlib
    from oracle import make_golden_runner as mgr
    rng = np.random.default_rng(31)
    x = torch.from_numpy(rng.standard_normal((mgr.BATCH, 1, 4, 4)).astype(np.float32))
    y = torch.from_numpy(rng.integers(0, mgr.K_CLASSES, mgr.BATCH).astype(np.int64))
    crit = torch.nn.CrossEntropyLoss()
    names = [n for n, _ in mgr.InjectNet(0).named_parameters()]
    sizes = [p.numel() for _, p in mgr.InjectNet(0).named_parameters()]
    out = dict(names=np.array(names), sizes=np.array(sizes))
    flatten = lambda mod: torch.cat([p.detach().reshape(-1) for p in mod.parameters()]).numpy().copy()

    vi = refshim.load("methods.vi")
    net, net0 = mgr.InjectNet(41), mgr.InjectNet(42)
    model = vi.Model(net, ND=12)
    with torch.no_grad():
        for p in model.s_.parameters():                      # spread around the clamp: negatives, < 1e-8, ordinary values
            vals = rng.choice([-1e-3, 0.0, 3e-9, 1e-8, 2e-8, 1e-6, 1e-3, 0.05], size=tuple(p.shape)).astype(np.float32)
            p.copy_(torch.from_numpy(vals) * torch.from_numpy(rng.uniform(0.5, 1.5, tuple(p.shape)).astype(np.float32)))
    tape = rng.standard_normal(sum(sizes)).astype(np.float32)
    net.eval()
    with refshim.injected_noise(tape) as tp:
        model.forward(x, y, net, net0, crit, eval_grad=0)
        assert tp.pos == tape.size
    out.update(vi_m=flatten(model.m), vi_s=flatten(model.s_), vi_eps=tape, vi_theta=flatten(net))

    mcd = refshim.load("methods.mc_dropout")

    @contextlib.contextmanager
    def injected_uniforms(flat):
        t = refshim.NoiseTape(flat)
        orig = torch.rand_like
        torch.rand_like = t
        try:
            yield t
        finally:
            torch.rand_like = orig
    for mode in ("gaussian", "spikymix", "ignore"):
        net, net0 = mgr.InjectNet(51), mgr.InjectNet(52)
        model = mcd.Model(net, ND=12, p_drop=0.3, bias=mode)
        with torch.no_grad():
            for p in model.m.parameters():
                p.add_(torch.from_numpy(rng.standard_normal(tuple(p.shape)).astype(np.float32)))
        u = rng.random(sum(sizes)).astype(np.float32)
        u[::97] = np.float32(0.3)                            # exactly p_drop: '>' must not keep these
        net.eval()
        with injected_uniforms(u) as tp:
            model.forward(x, y, net, net0, crit, eval_grad=0)
            used = tp.pos
        out.update({f"mcd_{mode}_m": flatten(model.m), f"mcd_{mode}_theta0": flatten(net0), f"mcd_{mode}_u": u[:used],
                    f"mcd_{mode}_used": used, f"mcd_{mode}_theta": flatten(net)})
    out["p_drop"] = np.float32(0.3)
    save("reparam_draws", **out)


def make_temperature_golden():
    """calibration.find_optimal_temperature (the reference's own function, scipy BFGS) on the calibration data sets, plus
    values of its objective at fixed temperatures computed with the reference's expression (calibration.py:179-183)."""
    import scipy.special
    cal = refshim.load("calibration")
    z = np.load(os.path.join(GOLDEN_DIR, "calibration.npz"))
    out = {}
    Ts = np.array([0.25, 0.7, 1.0, 1.37, 2.5, 9.0])
    for tag in "abcd":
        logits, labels = z[tag + "_logits"], z[tag + "_labels"]
        Topt, ok = cal.find_optimal_temperature(labels, logits, os.path.join(tempfile.gettempdir(), "t.png"))
        out[tag + "_Topt"] = np.asarray(Topt, dtype=np.float64)
        out[tag + "_success"] = bool(ok)
        vals = []
        for T in Ts:
            lg = logits / np.array([T])
            vals.append(np.mean(scipy.special.logsumexp(lg, axis=1) - lg[np.arange(len(labels)), labels]))
        out[tag + "_fun"] = np.array(vals, dtype=np.float64)
        print(f"  temperature {tag}: Topt = {np.asarray(Topt).reshape(-1)[0]:.6f} success={ok}")
    out["Ts"] = Ts
    save("temperature", **out)


if __name__ == "__main__":
    torch.set_num_threads(1)
    which = sys.argv[1:] or ["step", "cyclical", "calibration", "runner"]
    if "step" in which:
        make_step_goldens()
    if "clip" in which:                       # args.clip_grad cases only (added in round 2; "step" leaves them alone)
        make_clip_goldens()
    if "cyclical" in which:
        make_cyclical_golden()
    if "calibration" in which:
        make_calibration_golden()
    if "runner" in which:
        from oracle import make_golden_runner
        make_golden_runner.main(save)
    if "runner_fs" in which:                  # only the full-sample-store cases (methods/csghmc_fs.py)
        from oracle import make_golden_runner
        make_golden_runner.main(save, only=list(make_golden_runner.FS_CASES) + ["real_csghmc_fs"])
    if "cfg1" in which:                       # BASELINE.json configs[0]: mlp_mnist SGLD end to end
        from oracle import make_golden_runner
        make_golden_runner.main(save, only="cfg1")
    if "temperature" in which:
        make_temperature_golden()
    if "attrs" in which:
        make_attribute_inventory()
    if "reparam" in which:
        make_reparam_draw_goldens()
