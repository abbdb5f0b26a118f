"""The reference's update statements written with torch eager ops (TEST INFRASTRUCTURE ONLY).

On the GPU box the reference checkout does not exist, but torch does: these functions restate, statement by statement
and in the reference's operation order, the per-tensor update loops of methods/{sgld,sghmc,csghmc,adam_sghmc,
adam_csghmc}.py and then call the REAL ``torch.optim.SGD`` for the parameter step, so the CUDA kernels can be compared
with what the reference computes *on the same device* (torch CUDA evaluates ``tensor / python_scalar`` as a multiply by
the fp32 reciprocal -- the semantics of the product's default ``div=recip`` mode).  Noise is injected: ``xi`` replaces
``torch.randn_like(p)``.
"""
import numpy as np
import torch


def make_sgd(params_body, params_head, lr_body, lr_head, momentum):
    """methods/sghmc.py:53-57: two parameter groups, weight_decay 0."""
    groups = [{"params": params_body, "lr": lr_body}]
    if params_head:
        groups.append({"params": params_head, "lr": lr_head})
    return torch.optim.SGD(groups, momentum=momentum, weight_decay=0)


def _lr(name, readout, lr_body, lr_head):
    return lr_head if readout in name else lr_body


@torch.no_grad()
def sgld(named, p0s, xis, readout, *, lr_body, lr_head, N, prior_sig, nd, bias):
    """methods/sgld.py:469-484 (the SGD step with momentum follows, :226)."""
    for (name, p), p0, xi in zip(named, p0s, xis):
        lr = _lr(name, readout, lr_body, lr_head)
        if "bias" in name and bias == "uninformative":
            p.grad = p.grad + (nd * np.sqrt(2 / (N * lr)) * xi)
        else:
            p.grad = p.grad + ((p - p0) / (prior_sig ** 2) / N + nd * np.sqrt(2 / (N * lr)) * xi)


@torch.no_grad()
def sghmc(named, p0s, xis, vs, readout, *, lr_body, lr_head, N, prior_sig, nd, alpha, bias):
    """methods/sghmc.py:482-510; ``vs`` is the momentum_buffer dict."""
    for (name, p), p0, xi in zip(named, p0s, xis):
        lr = _lr(name, readout, lr_body, lr_head)
        v = vs[name]
        grad_U = p.grad if ("bias" in name and bias == "uninformative") else p.grad + (p - p0) / (prior_sig ** 2) / N
        noise = nd * np.sqrt(2 * alpha / (N * lr)) * xi
        v = v * (1 - alpha) + lr * grad_U + noise
        vs[name] = v
        p.grad = p.grad + v.clone()


@torch.no_grad()
def csghmc(named, xis, vs, readout, *, lr_body, lr_head, N, prior_sig, nd, alpha, should_sample):
    """methods/csghmc.py:747-778: no net0, no optimizer step -- p.data.add_(v)."""
    for (name, p), xi in zip(named, xis):
        lr = _lr(name, readout, lr_body, lr_head)
        v = vs[name]
        grad_U = p.grad + prior_sig * p.data
        noise = nd * np.sqrt((2 * alpha * lr)) / N * xi
        v = v * (1 - alpha) - lr * grad_U + noise if should_sample else v * (1 - alpha) - lr * grad_U
        vs[name] = v
        p.data.add_(v)


@torch.no_grad()
def adam(named, p0s, xis, vs, ms, ss, readout, *, lr_body, lr_head, N, prior_sig, nd, alpha, beta1, beta2, eps, t, bias,
         cyclical, temperature=1.0):
    """methods/adam_sghmc.py:507-553 (p.grad = p.grad + v) and methods/adam_csghmc.py:814-861 (g / T, p.grad = v)."""
    for (name, p), p0, xi in zip(named, p0s, xis):
        lr = _lr(name, readout, lr_body, lr_head)
        vm, m, s = vs[name], ms[name], ss[name]
        g = p.grad / temperature if cyclical else p.grad
        grad_U = g if ("bias" in name and bias == "uninformative") else g + (p - p0) / (prior_sig ** 2) / N
        m = beta1 * m + (1 - beta1) * grad_U
        s = beta2 * s + (1 - beta2) * (grad_U * grad_U)
        m_hat = m / (1 - beta1 ** t)
        s_hat = s / (1 - beta2 ** t)
        precond_grad = m_hat / (torch.sqrt(s_hat) + eps)
        precond_term = 1.0 / (torch.sqrt(s_hat) + eps)
        noise = nd * torch.sqrt(2 * alpha * precond_term / N) * xi
        vm = vm * (1 - alpha) + lr * precond_grad + noise
        vs[name], ms[name], ss[name] = vm, m, s
        p.grad = vm.clone() if cyclical else p.grad + vm.clone()


@torch.no_grad()
def evaluate_cyclical_avg(net, mom1, mom2, samples_per_cycle, gmm_weights, nst, loader, device):
    """methods/csgld.py:333-456 for nst > 0: per batch, per kept cycle, ``nst`` networks drawn as
    ``p_mean + p_var.sqrt() * randn_like(p)`` with ``p_var = clamp(ratio * (mom2 - mom1**2), 1e-12)``; component output
    ``logsumexp_S(log_softmax_K) - log S``; mixture = weighted sum of the component outputs.
    ``mom1`` / ``mom2``: dict cycle -> dense parameters_to_vector-ordered vectors.  Returns (logits, logits_all, targets)."""
    import copy
    import torch.nn.functional as F
    from torch.nn.utils import vector_to_parameters
    logits, logits_all, targets = [], [], []
    for x, y in loader:
        x, y = x.to(device), y.to(device)
        comps, mix = [], None
        for c in mom1:
            w = gmm_weights.get(c, 0.0)
            if w < 1e-10:
                continue
            net_c = copy.deepcopy(net)
            net_c.eval()
            means, variances = copy.deepcopy(net_c), copy.deepcopy(net_c)
            cnt = samples_per_cycle.get(c, 0)
            ratio = cnt / (cnt - 1)
            var = ratio * (mom2[c] - mom1[c] ** 2) if cnt > 1 else mom2[c] - mom1[c] ** 2
            var.clamp_(min=1e-12)
            vector_to_parameters(var, variances.parameters())
            vector_to_parameters(mom1[c], means.parameters())
            outs = []
            for _ in range(nst):
                sample = copy.deepcopy(net_c)
                for p, pm, pv in zip(sample.parameters(), means.parameters(), variances.parameters()):
                    eps = torch.randn_like(p)
                    p.copy_(pm + pv.sqrt() * eps)
                outs.append(sample(x))
            stacked = torch.stack(outs, dim=2)
            comp = F.log_softmax(stacked, dim=1).logsumexp(-1) - np.log(nst)
            comps.append(stacked)
            mix = w * comp if mix is None else mix + w * comp
        logits.append(mix)
        logits_all.append(torch.stack(comps, dim=3))
        targets.append(y)
    return torch.cat(logits), torch.cat(logits_all), torch.cat(targets)
