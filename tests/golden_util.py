"""Helpers to load golden fixtures (tests/golden/*.npz) and replay them through an implementation.

A "stepper" is any callable with the signature of ``oracle.sampler_oracle.step_*``; the GPU tests
pass thin adapters around the C-ABI instead, so the CPU-oracle test and the CUDA parity test share
this replay code.
"""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


def step_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "step_*.npz")))


def load_step_case(name):
    z = np.load(golden_path(name), allow_pickle=False)
    hp = dict(zip(z["hp_keys"].tolist(), z["hp_vals"].tolist()))
    method = name[len("step_"):]
    for m in ("adam_csghmc", "adam_sghmc", "csghmc", "csgld", "sghmc", "sgld"):
        if method.startswith(m):
            method = m
            break
    return z, hp, method


def hparams_from(hp, z, method):
    from oracle.sampler_oracle import HParams
    mu = float(z["momentum"])
    if method in ("sghmc", "csghmc", "adam_csghmc"):
        mu = 0.0                                    # these Runners build SGD(momentum=0)
    return HParams(
        ND=float(z["ND"]), Ninflate=float(hp["Ninflate"]), prior_sig=float(hp["prior_sig"]), nd=float(hp["nd"]),
        alpha=float(hp.get("momentum_decay", 0.05)), beta1=float(hp.get("beta1", 0.9)),
        beta2=float(hp.get("beta2", 0.999)), eps=float(hp.get("epsilon", 1e-8)),
        temperature=float(hp.get("temperature", 1.0)), mu=mu)


def prior_mask(hp, z):
    if hp["bias"] == "uninformative":
        return (~z["is_bias"]).astype(np.float32)
    return np.ones(z["is_bias"].shape, np.float32)


def max_rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def recorded_clip_coef(z, t):
    """clip_coef_clamped of step t as torch.nn.utils.clip_grad_norm_ forms it from the RECORDED total norm (fp32):
    max_norm / (total_norm + 1e-6) = (total_norm + 1e-6).reciprocal() * max_norm (Tensor.__rtruediv__), clamped to <= 1."""
    f32 = np.float32
    coef = (f32(1.0) / (f32(z["total_norm"][t]) + f32(1e-6))) * f32(float(z["clip_grad"]))
    return f32(min(coef, f32(1.0)))


def valid_mask(z):
    """Real elements of tensors with a gradient: everything, in the dense layout of the goldens."""
    return np.ones(z["G"].shape[1], bool)


def replay_step_case(name, impl, chained=True, div_mode="true"):
    """Replay one step golden through ``impl`` (module/object exposing the oracle's step_* API).

    chained=True : start from theta_init and feed the implementation's own outputs forward
                   (trajectory parity).
    chained=False: every step starts from the golden pre-state (per-step parity).
    Returns dict name -> list over steps of (got, want) pairs.
    """
    z, hp, method = load_step_case(name)
    H = hparams_from(hp, z, method)
    is_head = z["is_head"]
    P = prior_mask(hp, z)
    T = z["G"].shape[0]
    n = z["G"].shape[1]
    zeros = np.zeros(n, np.float32)
    theta0 = z["theta0"]
    st = dict(theta=z["theta_init"].copy(), v=zeros.copy(), m=zeros.copy(), s=zeros.copy(), buf=zeros.copy())
    pairs = {k: [] for k in ("theta", "v", "m", "s", "buf") if k in z.files}
    t_adam = 0
    for t in range(T):
        if not chained and t > 0:
            for k in st:
                if k in z.files:
                    st[k] = z[k][t - 1].copy()
        g, xi = z["G"][t], z["XI"][t]
        lrb, lrh = float(z["lr_body"][t]), float(z["lr_head"][t])
        kw = dict(is_head=is_head, lr_body=lrb, lr_head=lrh, hp=H)
        # args.clip_grad goldens: replay the reference's own coefficient so that the update itself is compared bit for
        # bit; the norm (a reduction: order-dependent) has its own tolerance test
        ck = dict(coef=recorded_clip_coef(z, t)) if "clip_grad" in z.files else {}
        if method in ("sgld", "csgld"):
            st["theta"], st["buf"] = impl.step_sgld(st["theta"], g, theta0, st["buf"], xi, P=P,
                                                    first_step=(t == 0), div_mode=div_mode, **kw, **ck)
        elif method == "sghmc":
            st["theta"], st["v"] = impl.step_sghmc(st["theta"], g, theta0, st["v"], xi, P=P,
                                                   div_mode=div_mode, **kw)
        elif method == "csghmc":
            st["theta"], st["v"] = impl.step_csghmc(st["theta"], g, st["v"], xi,
                                                    should_sample=bool(z["sample_pattern"][t]), **kw)
        elif method == "adam_sghmc":
            t_adam += 1
            st["theta"], st["v"], st["m"], st["s"], st["buf"] = impl.step_adam_sghmc(
                st["theta"], g, theta0, st["v"], st["m"], st["s"], st["buf"], xi, P=P, t=t_adam,
                first_step=(t == 0), div_mode=div_mode, **kw)
        elif method == "adam_csghmc":
            t_adam += 1
            st["theta"], st["v"], st["m"], st["s"] = impl.step_adam_csghmc(
                st["theta"], g, theta0, st["v"], st["m"], st["s"], xi, P=P, t=t_adam,
                div_mode=div_mode, **kw, **ck)
        else:
            raise ValueError(method)
        for k in pairs:
            if k == "buf" and H.mu == 0:
                continue
            pairs[k].append((np.asarray(st[k]).copy(), z[k][t]))
    return pairs


def mc_dropout_dense_inputs(z, mode):
    """(u, nodrop) over the dense parameter vector for one MC-Dropout golden: the reference's uniform tape holds values
    only for the tensors it drops, in order; tensors kept whole (bias, modes 'gaussian' / 'ignore') get u = 0."""
    names, sizes = z["names"].tolist(), z["sizes"].tolist()
    tape = z[f"mcd_{mode}_u"]
    u, nodrop, pos = [], [], 0
    for name, k in zip(names, sizes):
        keep = "bias" in name and mode != "spikymix"
        if keep:
            u.append(np.zeros(k, np.float32))
        else:
            u.append(tape[pos:pos + k])
            pos += k
        nodrop.append(np.full(k, keep))
    assert pos == int(z[f"mcd_{mode}_used"]) == tape.size
    return np.concatenate(u).astype(np.float32), np.concatenate(nodrop)
