"""CPU oracle: numpy restatement of BayesDLL's SG-MCMC sampler hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this module.
The product path (``bayesdll_b200``) never imports it and has no CPU fallback.

Parity pin: every function here is checked against golden vectors produced by
running the reference's own PyTorch code (``/root/reference``) with injected
gradients and injected noise -- see ``oracle/make_golden.py`` and
``tests/test_oracle_golden.py``.  Versions used for the pin are stored next to
the vectors (torch 2.11.0+cu128 / numpy 2.3.5 / scipy 1.18.1).

All tensors are fp32 numpy arrays; Python scalars are doubles that are rounded
to fp32 *at the op where they meet a tensor* exactly as torch does
(SURVEY.md Appendix A).  Each numpy op rounds once to fp32 (IEEE RN), which is
what the reference's eager elementwise torch ops do on CPU.  Two places differ
from "one rounding per op":
  * ``Tensor.add_(other, alpha=s)`` (used by torch.optim.SGD for
    ``p.add_(grad, alpha=-lr)``) is a fused multiply-add on CPU (vec::fmadd) and on
    CUDA (compiler contraction inside the functor); restated by ``_fma``.
  * ``div_mode='recip'`` restates torch-CUDA semantics where ``tensor / python_scalar``
    is evaluated as ``tensor * (float)(1.0 / scalar)``, the reciprocal taken in double
    (BinaryDivTrueKernel.cu).  Pinned on the GPU box against torch CUDA itself:
    tools/probe_torch_div.py and tests/test_cuda_eager_parity_gpu.py (the kernels in this
    mode are bit-identical to the reference's statements run in torch CUDA eager ops).
    ``div_mode='true'`` is IEEE division, what the reference does on CPU and what
    the golden vectors pin.

Elements are addressed through flat fp32 vectors in ``named_parameters()`` order.
Per-element metadata: ``is_head`` (readout_name in pname -> lr_head) and ``P``
(0 where 'bias' in pname and bias=='uninformative', else 1).
"""
from dataclasses import dataclass

import numpy as np

f32 = np.float32
f64 = np.float64


def _fma(a, b, c):
    """RN(a*b + c) for fp32 operands.  a*b is exact in fp64 (48-bit product); the fp64
    add rounds to 53 bits and the final cast to 24: double rounding differs from a
    true fma with probability ~2^-29 per element (documented; the C oracle uses fmaf)."""
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def _div_scalar(x, s, div_mode):
    """tensor / python_scalar.  'true': IEEE divide by fp32(s) (torch CPU);
    'recip': multiply by fp32(1.0 / s) with the reciprocal taken in DOUBLE from the Python double (torch CUDA,
    BinaryDivTrueKernel.cu; pinned by tools/probe_torch_div.py and tests/test_cuda_eager_parity_gpu.py)."""
    if div_mode == "true":
        return x / f32(s)
    if div_mode == "recip":
        return x * f32(1.0 / float(s))
    raise ValueError(div_mode)


@dataclass
class HParams:
    ND: float                 # training-set size (args.ND)
    Ninflate: float = 1.0
    prior_sig: float = 1.0
    nd: float = 1.0           # noise discount
    alpha: float = 0.05       # momentum_decay
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    temperature: float = 1.0
    mu: float = 0.0           # torch SGD momentum (args.momentum)

    @property
    def N(self):
        return self.ND * self.Ninflate


def _per_class(is_head, body_val, head_val):
    return np.where(is_head, f32(head_val), f32(body_val)).astype(f32)


def _prior_term(theta, theta0, hp, div_mode):
    # ((p - p0) / (prior_sig**2)) / N        methods/sghmc.py:497, methods/sgld.py:482
    d = theta - theta0
    d = _div_scalar(d, hp.prior_sig ** 2, div_mode)
    d = _div_scalar(d, hp.N, div_mode)
    return d


def _sgd_apply(theta, gprime, buf, lr_e, mu, first_step):
    """torch.optim.SGD single step, dampening=0, nesterov=False, weight_decay=0
    (torch/optim/sgd.py::_single_tensor_sgd).  Returns (theta', buf')."""
    if mu != 0:
        if first_step:
            buf = gprime.copy()                      # buf = clone(grad)
        else:
            buf = buf * f32(mu)                      # buf.mul_(momentum)
            buf = buf + gprime                       # .add_(grad, alpha=1-dampening) == exact add
        d = buf
    else:
        d = gprime
    theta = _fma(d, -lr_e, theta)                    # param.add_(d, alpha=-lr)
    return theta, buf


# ----------------------------------------------------------------------------------------------
# (a1) SGLD / cSGLD      methods/sgld.py:469-484 + SGD.step :226 ; methods/csgld.py:665-680 + :253
# ----------------------------------------------------------------------------------------------
def clip_coef(pgrad, valid, max_norm):
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm) (methods/csgld.py:250-251, methods/adam_csghmc.py:319-320):
    total_norm = || stack(||g_t||_2) ||_2 over the tensors that have a gradient, clip_coef = max_norm / (total_norm + 1e-6)
    -- evaluated by torch as ``(total_norm + 1e-6).reciprocal() * max_norm`` (Tensor.__rtruediv__: two roundings) --,
    clamped to <= 1, all in fp32.  ``pgrad``: what the reference holds in p.grad at that point (flat, padded layout);
    ``valid``: boolean mask of the real elements of tensors with a gradient.  The sum of squares is formed in fp64 here
    (torch: fp32, per tensor, vectorised), so total_norm agrees with torch's to ~1e-7 relative, not bit for bit.
    -> (coef fp32, total_norm fp32)."""
    q = np.asarray(pgrad, f64)[np.asarray(valid, bool)]
    total = f32(np.sqrt(np.sum(q * q)))
    coef = (f32(1.0) / (total + f32(1e-6))) * f32(max_norm)      # float / Tensor is Tensor.__rtruediv__: reciprocal() * other
    return f32(min(coef, f32(1.0))), total


def step_sgld(theta, g, theta0, buf, xi, *, is_head, P, lr_body, lr_head, hp, first_step,
              div_mode="true", clip=None, valid=None, coef=None):
    """``clip`` = args.clip_grad: the modified gradient is scaled by clip_coef(g', valid, clip) before the SGD step
    (methods/csgld.py:250-253); ``coef`` overrides the coefficient (replaying a recorded one)."""
    N = hp.N
    lr_e = _per_class(is_head, lr_body, lr_head)
    c_body = hp.nd * np.sqrt(2 / (N * lr_body))      # host fp64, methods/sgld.py:478/483
    c_head = hp.nd * np.sqrt(2 / (N * lr_head))
    noise = _per_class(is_head, c_body, c_head) * xi
    with_prior = _prior_term(theta, theta0, hp, div_mode) + noise
    add = np.where(P.astype(bool), with_prior, noise)
    gprime = g + add                                  # p.grad = p.grad + ( ... )
    if clip is not None or coef is not None:
        if coef is None:
            coef, _ = clip_coef(gprime, valid, clip)
        gprime = gprime * f32(coef)                   # g.mul_(clip_coef_clamped)
    return _sgd_apply(theta, gprime, buf, lr_e, hp.mu, first_step)


# ----------------------------------------------------------------------------------------------
# (a2) SGHMC             methods/sghmc.py:482-510 + SGD(momentum=0).step :229
# ----------------------------------------------------------------------------------------------
def step_sghmc(theta, g, theta0, v, xi, *, is_head, P, lr_body, lr_head, hp, div_mode="true"):
    N = hp.N
    lr_e = _per_class(is_head, lr_body, lr_head)
    gU = np.where(P.astype(bool), g + _prior_term(theta, theta0, hp, div_mode), g)   # :494-497
    c_body = hp.nd * np.sqrt(2 * hp.alpha / (N * lr_body))                           # :500
    c_head = hp.nd * np.sqrt(2 * hp.alpha / (N * lr_head))
    noise = _per_class(is_head, c_body, c_head) * xi                                 # :501
    v = (v * f32(1 - hp.alpha) + lr_e * gU) + noise                                  # :504
    gprime = g + v                                                                   # :510
    theta = _fma(gprime, -lr_e, theta)                                               # SGD, momentum 0
    return theta, v


# ----------------------------------------------------------------------------------------------
# (a3) cSGHMC            methods/csghmc.py:747-778 (no optimizer step, :304)
# ----------------------------------------------------------------------------------------------
def step_csghmc(theta, g, v, xi, *, is_head, lr_body, lr_head, hp, should_sample):
    N = hp.N
    lr_e = _per_class(is_head, lr_body, lr_head)
    gU = g + f32(hp.prior_sig) * theta                                               # :760/:762
    c_body = hp.nd * np.sqrt(2 * hp.alpha * lr_body) / N                             # :765
    c_head = hp.nd * np.sqrt(2 * hp.alpha * lr_head) / N
    noise = _per_class(is_head, c_body, c_head) * xi                                 # :766
    v = v * f32(1 - hp.alpha) - lr_e * gU                                            # :772
    if should_sample:
        v = v + noise                                                                # :770
    theta = theta + v                                                                # :778
    return theta, v


# ----------------------------------------------------------------------------------------------
# (a4) Adam-SGHMC        methods/adam_sghmc.py:507-553, t+=1 :494, SGD(momentum=args.momentum) :233
# (a5) Adam-cSGHMC       methods/adam_csghmc.py:814-861, SGD(momentum=0) :322
# ----------------------------------------------------------------------------------------------
def _adam_core(theta, g, theta0, v, m, s, xi, *, is_head, P, lr_body, lr_head, hp, t, cyc, div_mode):
    N = hp.N
    lr_e = _per_class(is_head, lr_body, lr_head)
    gl = _div_scalar(g, hp.temperature, div_mode) if cyc else g                      # adam_csghmc:829/831
    gU = np.where(P.astype(bool), gl + _prior_term(theta, theta0, hp, div_mode), gl)
    m = f32(hp.beta1) * m + f32(1 - hp.beta1) * gU                                   # :527 / :834
    s = f32(hp.beta2) * s + f32(1 - hp.beta2) * (gU * gU)                            # :530 / :837
    m_hat = _div_scalar(m, 1 - hp.beta1 ** t, div_mode)                              # :533 / :840
    s_hat = _div_scalar(s, 1 - hp.beta2 ** t, div_mode)                              # :534 / :841
    den = np.sqrt(s_hat) + f32(hp.eps)
    pg = m_hat / den                                                                 # :537 / :844
    pre = (f32(1.0) / den) * f32(1.0)                                                # :540 / :847 (rtruediv)
    ns = f32(hp.nd) * np.sqrt(_div_scalar(f32(2 * hp.alpha) * pre, N, div_mode))     # :541 / :848
    noise = ns * xi                                                                  # :542 / :849
    v = (v * f32(1 - hp.alpha) + lr_e * pg) + noise                                  # :545 / :852
    return lr_e, v, m, s


def step_adam_sghmc(theta, g, theta0, v, m, s, buf, xi, *, is_head, P, lr_body, lr_head, hp, t,
                    first_step, div_mode="true"):
    """``t`` is the value of Model.t *after* the ``self.t += 1`` at adam_sghmc.py:494."""
    lr_e, v, m, s = _adam_core(theta, g, theta0, v, m, s, xi, is_head=is_head, P=P, lr_body=lr_body,
                               lr_head=lr_head, hp=hp, t=t, cyc=False, div_mode=div_mode)
    gprime = g + v                                                                   # :553
    theta, buf = _sgd_apply(theta, gprime, buf, lr_e, hp.mu, first_step)
    return theta, v, m, s, buf


def step_adam_csghmc(theta, g, theta0, v, m, s, xi, *, is_head, P, lr_body, lr_head, hp, t,
                     div_mode="true", clip=None, valid=None, coef=None):
    """``clip`` / ``coef`` as in step_sgld: p.grad = v is scaled before the SGD step (adam_csghmc.py:319-322); the
    momentum buffer itself keeps the unclipped v."""
    lr_e, v, m, s = _adam_core(theta, g, theta0, v, m, s, xi, is_head=is_head, P=P, lr_body=lr_body,
                               lr_head=lr_head, hp=hp, t=t, cyc=True, div_mode=div_mode)
    pgrad = v
    if clip is not None or coef is not None:
        if coef is None:
            coef, _ = clip_coef(v, valid, clip)
        pgrad = v * f32(coef)
    theta = _fma(pgrad, -lr_e, theta)                                                # p.grad = v (:861); SGD mu=0
    return theta, v, m, s


# ----------------------------------------------------------------------------------------------
# (a6) cyclical schedule  methods/cyclical.py:29-74 (pure host scalar math, fp64)
# ----------------------------------------------------------------------------------------------
class CyclicalOracle:
    def __init__(self, base_lr, nbr_of_cycles, epochs, proportion_exploration=0.5):
        self.base_lr, self.M, self.epochs, self.beta = base_lr, nbr_of_cycles, epochs, proportion_exploration

    def calculate_lr(self, epoch, batch, B):
        K = self.epochs * B
        L = K // self.M                               # integer cycle length (:32)
        k = epoch * B + batch + 1
        pos = ((k - 1) % L) / L
        return self.base_lr * (1 + np.cos(pos * np.pi)) / 2

    def should_sample(self, epoch, batch, B):
        K = self.epochs * B
        L = K / self.M                                # float cycle length (:54)
        k = epoch * B + batch + 1
        return ((k - 1) % L) / L >= self.beta

    def last_in_cycle(self, epoch, batch, B):
        K = self.epochs * B
        L = K / self.M
        k = epoch * B + batch + 1
        return (k % L) == 0

    def get_cycle_number(self, epoch, batch, B):
        K = self.epochs * B
        L = K / self.M
        k = epoch * B + batch + 1
        return int((k - 1) // L) + 1


# ----------------------------------------------------------------------------------------------
# (a7) running moments    methods/sgld.py:95-102, 239-246
# (a8) cyclical moments   methods/csgld.py:276-293 ; Welford methods/csghmc.py:327-348
# ----------------------------------------------------------------------------------------------
def moments_init(theta):
    return theta * f32(1.0), theta * theta            # theta_vec*1.0 ; theta_vec**2


def moments_avg(theta, mom1, mom2, cnt, div_mode="true"):
    """mom <- (theta^k + cnt*mom)/(cnt+1); caller does cnt += 1.  (csgld: cnt = cycle_count-1)."""
    mom1 = _div_scalar(theta + f32(cnt) * mom1, cnt + 1, div_mode)
    mom2 = _div_scalar(theta * theta + f32(cnt) * mom2, cnt + 1, div_mode)
    return mom1, mom2


def moments_welford(theta, mean, M2, n, div_mode="true"):
    """csghmc.py:340-345 with n = samples_per_cycle + 1 supplied by the caller."""
    delta = theta - mean
    mean = mean + _div_scalar(delta, n, div_mode)
    delta2 = theta - mean
    M2 = M2 + delta * delta2
    return mean, M2


# ----------------------------------------------------------------------------------------------
# (a9) variance + posterior draw   methods/sgld.py:338-348, 292-297 ; csghmc.py:451-459
# ----------------------------------------------------------------------------------------------
def variance_from_moments(mom1, mom2, ratio):
    var = f32(ratio) * (mom2 - mom1 * mom1)
    return np.maximum(var, f32(1e-12))


def variance_from_welford(M2, n_samples, div_mode="true"):
    if n_samples > 1:
        var = _div_scalar(M2, n_samples - 1, div_mode)
    else:
        var = np.ones_like(M2) * f32(1e-12)
    return np.maximum(var, f32(1e-12))


def posterior_draw(mean, var, eps):
    return mean + np.sqrt(var) * eps                  # p_m + p_v.sqrt()*eps


# ----------------------------------------------------------------------------------------------
# (a10) ensemble average   methods/sgld.py:283-305 ; mixture methods/csgld.py:416-439, weights :565-594
# ----------------------------------------------------------------------------------------------
def _log_softmax(x, axis):
    x = np.asarray(x, f32)
    mx = x.max(axis=axis, keepdims=True)
    sh = x - mx
    lse = np.log(np.exp(sh).sum(axis=axis, keepdims=True, dtype=f32))
    return (sh - lse).astype(f32)


def _logsumexp(x, axis):
    mx = x.max(axis=axis, keepdims=True)
    out = np.log(np.exp(x - mx).sum(axis=axis, keepdims=True, dtype=f32)) + mx
    return np.squeeze(out, axis=axis).astype(f32)


def ensemble_average(logits_all, nst):
    """logits_all [B,K,S] -> [B,K] log-mean-softmax.  nst==0: single mean-parameter pass, no -log S."""
    ls = _log_softmax(logits_all, 1)
    out = _logsumexp(ls, -1)
    if nst > 0:
        out = out - f32(np.log(nst))
    return out.astype(f32)


def gmm_weights(cycle_likelihoods):
    """dict cycle -> iterable of likelihoods -> normalised weights (fp64, host)."""
    if not cycle_likelihoods:
        return {0: 1.0}
    w = {c: 1.0 / np.mean([1.0 / l for l in ls]) for c, ls in cycle_likelihoods.items()}
    tot = sum(w.values())
    if tot > 0:
        return {c: x / tot for c, x in w.items()}
    return {c: 1.0 / len(w) for c in w}


def mixture(component_logits, weights, nst):
    """component_logits: list over kept cycles of [B,K,S]; weights: matching list of python floats.
    Weighted *sum of log-probabilities* (csgld.py:428-431)."""
    out = None
    for cl, w in zip(component_logits, weights):
        comp = cl[:, :, 0] if nst == 0 else ensemble_average(cl, nst)
        term = f32(w) * comp
        out = term if out is None else out + term
    return out.astype(f32)


def cross_entropy_mean(logits, y):
    ls = _log_softmax(logits, 1)
    return float(-ls[np.arange(len(y)), y].mean(dtype=f32))


# ----------------------------------------------------------------------------------------------
# (a11) calibration        calibration.py:24-67, 215-259   (reference is numpy/scipy itself)
# ----------------------------------------------------------------------------------------------
def softmax_scipy(x, axis=1):
    """scipy.special.softmax (scipy 1.18.1 _logsumexp.py): exp(x - max) / sum, in x's dtype."""
    x_max = np.amax(x, axis=axis, keepdims=True)
    e = np.exp(x - x_max)
    return e / np.sum(e, axis=axis, keepdims=True)


def bin_edges(num_bins):
    return np.linspace(0, 1 + 1e-8, num_bins + 1)[1:]


def calc_bins(labels, logits, num_bins, temperature=1):
    K = logits.shape[1]
    onehot = np.eye(K)[labels].flatten()
    preds = softmax_scipy(logits / temperature, axis=1).flatten()
    edges = bin_edges(num_bins)
    binned = np.digitize(preds, edges)
    sizes = np.zeros(num_bins)
    accs = np.zeros(num_bins)
    confs = np.zeros(num_bins)
    for b in range(num_bins):
        sel = binned == b
        sizes[b] = sel.sum()
        if sizes[b] > 0:
            accs[b] = onehot[sel].sum() / sizes[b]
            confs[b] = preds[sel].sum() / sizes[b]
    return edges, binned, accs, confs, sizes


def analyze(labels, logits, num_bins, temperature=1):
    edges, binned, accs, confs, sizes = calc_bins(labels, logits, num_bins, temperature)
    ece = (np.abs(accs - confs) * (sizes / sizes.sum())).sum()
    mce = np.abs(accs - confs).max()
    lg = logits / temperature
    mx = lg.max(axis=1, keepdims=True)
    lse = np.log(np.exp(lg - mx).sum(axis=1)) + mx[:, 0]
    nll = np.mean(lse - lg[np.arange(len(labels)), labels])
    return ece, mce, nll


# ----------------------------------------------------------------------------------------------
# (8f row 4) temperature scaling   calibration.py:123-212 (objective :178-184, driver :196)
# ----------------------------------------------------------------------------------------------
def nll_temperature(labels, logits, T):
    """fun(T) of find_optimal_temperature: mean(logsumexp(logits/T, axis=1) - (logits/T)[i, y_i]).  ``T`` is the fp64
    ndarray scipy hands to the objective, so ``logits / T`` is fp64 even for fp32 logits."""
    z = np.asarray(logits) / np.asarray(T, dtype=np.float64)
    mx = z.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z - mx).sum(axis=1)) + mx[:, 0]
    return float(np.mean(lse - z[np.arange(len(labels)), labels]))


def find_optimal_temperature(labels, logits, max_iter=10000):
    """scipy BFGS (numerical gradient) from T = 1, as calibration.py:196.  -> (result.x, result.success)"""
    import scipy.optimize
    res = scipy.optimize.minimize(lambda T: nll_temperature(labels, logits, T), np.ones(1), options={"maxiter": max_iter})
    return res.x, res.success


# ----------------------------------------------------------------------------------------------
# (8f row 2) Bayesian model average over stored raw samples   methods/csghmc_fs.py:349-377
# ----------------------------------------------------------------------------------------------
def bma_mean(logits_all):
    """logits_all [N,K,S] fp32 (S = models in sorted file order) -> [N,K]: ``all_logits_sum += model_logits`` model by
    model in fp32, then ``/ num_models`` (numpy: fp32 array / python int -> IEEE fp32 division)."""
    la = np.asarray(logits_all, f32)
    acc = la[:, :, 0].copy()
    for m in range(1, la.shape[2]):
        acc += la[:, :, m]
    return (acc / la.shape[2]).astype(f32)


# ----------------------------------------------------------------------------------------------
# (8f row 4) per-step reparameterisation draws of the VI and MC-Dropout families
# ----------------------------------------------------------------------------------------------
def vi_sample(m, s_, eps):
    """methods/vi.py:402-406: ``p.copy_(p_m + p_s_.clamp(min=1e-8) * eps)``."""
    return (np.asarray(m, f32) + np.maximum(np.asarray(s_, f32), f32(1e-8)) * np.asarray(eps, f32)).astype(f32)


def mc_dropout_mix(m, theta0, u, p_drop, nodrop):
    """methods/mc_dropout.py:378-394: ``z = (rand_like(p) > p_drop).float()`` (ones where ``nodrop``: bias tensors in the
    'gaussian' / 'ignore' modes), ``p.copy_(z*p_m + (1-z)*p0)``.  Returns (theta, z)."""
    z = np.where(np.asarray(nodrop, bool) | (np.asarray(u, f32) > f32(p_drop)), f32(1), f32(0)).astype(f32)
    return (z * np.asarray(m, f32) + (f32(1) - z) * np.asarray(theta0, f32)).astype(f32), z
