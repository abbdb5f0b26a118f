"""Torch-eager, per-tensor restatement of the reference's SGHMC update loop + SGD step, as the reference itself
executes it (methods/sghmc.py:482-510 then torch.optim.SGD.step, :229): ~12 elementwise ops per tensor, one
``torch.randn_like`` per tensor.  BASELINE ONLY: bench.py times it on the host cores ("what the reference's own
structure achieves on this CPU", next to the fused C port oracle/bdl_oracle.c) and on the same B200 ("the number the
fused kernel replaces").  Nothing under bayesdll_b200/ imports it.
"""
import numpy as np
import torch


def sghmc_step_eager(params, grads, params0, momentum, names, readout_name, *, lr_body, lr_head, ND, Ninflate, prior_sig,
                     nd, alpha, bias="informative"):
    N = ND * Ninflate
    with torch.no_grad():
        for i, (pname, p, g, p0) in enumerate(zip(names, params, grads, params0)):
            lr = lr_head if readout_name in pname else lr_body
            v = momentum[i]
            if "bias" in pname and bias == "uninformative":
                grad_U = g
            else:
                grad_U = g + (p - p0) / (prior_sig ** 2) / N
            noise = nd * np.sqrt(2 * alpha / (N * lr)) * torch.randn_like(p)
            v = v * (1 - alpha) + lr * grad_U + noise
            momentum[i] = v
            new_grad = g + v.clone()
            p.add_(new_grad, alpha=-lr)            # SGD(momentum=0).step()
