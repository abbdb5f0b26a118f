"""Adapter exposing the oracle's step_* API on top of the CUDA C-ABI (padded flat device buffers), so the
golden replay code in golden_util.py drives the product path exactly like it drives the oracle."""
import numpy as np
import torch

from bayesdll_b200 import _lib, ops
from bayesdll_b200.flat import FlatLayout

DIV = {"true": _lib.DIV_IEEE, "recip": _lib.DIV_RECIP}


class GpuStepper:
    def __init__(self, names, sizes, readout, bias_mode, device, per_tensor_runs=False):
        self.layout = FlatLayout([(n, (int(s),)) for n, s in zip(names, sizes)], readout)
        self.bias_mode = bias_mode
        self.dev = device
        self.per_tensor_runs = per_tensor_runs
        tab = self.layout.run_table(bias_mode) if not per_tensor_runs else \
            self.layout.run_table(bias_mode, grad_ptrs=[0] * len(self.layout.segments))
        self.runs_dev, self.nruns = ops.upload_runs(tab, device)

    def up(self, dense):
        return torch.from_numpy(self.layout.padded_numpy(np.asarray(dense, np.float32))).to(self.dev)

    def down(self, t):
        return self.layout.dense_numpy(t.cpu().numpy())

    def _run(self, variant, sc, theta, g, theta0, v, m, s, buf, xi, coef=None):
        T = {k: (None if a is None else self.up(a)) for k, a in
             dict(theta=theta, g=g, theta0=theta0, v=v, m=m, s=s, buf=buf, xi=xi).items()}
        if coef is not None:        # args.clip_grad: pass 2 of the clipped update with a given coefficient (device scalar)
            tab = self.layout.run_table(self.bias_mode, grad_ptrs=[0] * len(self.layout.segments))
            ops.step_clipped(variant, T["theta"], T["g"], T["theta0"], T["v"], T["m"], T["s"], T["buf"], tab, len(tab), sc,
                             ops.make_noise(xi=T["xi"]), torch.tensor([coef], dtype=torch.float32, device=self.dev))
            torch.cuda.synchronize()
            return {k: (None if t is None else self.down(t)) for k, t in T.items()}
        ops.step(variant, T["theta"], T["g"], T["theta0"], T["v"], T["m"], T["s"], T["buf"], self.runs_dev,
                 self.nruns, sc, ops.make_noise(xi=T["xi"]))
        torch.cuda.synchronize()
        return {k: (None if t is None else self.down(t)) for k, t in T.items()}

    @staticmethod
    def _sc(variant, hp, lr_body, lr_head, div_mode, **kw):
        return ops.make_scalars(variant, lr_body=lr_body, lr_head=lr_head, ND=hp.ND, Ninflate=hp.Ninflate,
                                prior_sig=hp.prior_sig, nd=hp.nd, alpha=hp.alpha, mu=hp.mu, beta1=hp.beta1,
                                beta2=hp.beta2, eps=hp.eps, temperature=hp.temperature, div_mode=DIV[div_mode], **kw)

    def step_sgld(self, theta, g, theta0, buf, xi, *, is_head, P, lr_body, lr_head, hp, first_step, div_mode="true", coef=None):
        sc = self._sc(_lib.SGLD, hp, lr_body, lr_head, div_mode, first_step=first_step)
        o = self._run(_lib.SGLD, sc, theta, g, theta0, None, None, None, buf if hp.mu != 0 else None, xi, coef=coef)
        return o["theta"], (o["buf"] if hp.mu != 0 else buf)

    def step_sghmc(self, theta, g, theta0, v, xi, *, is_head, P, lr_body, lr_head, hp, div_mode="true"):
        sc = self._sc(_lib.SGHMC, hp, lr_body, lr_head, div_mode)
        o = self._run(_lib.SGHMC, sc, theta, g, theta0, v, None, None, None, xi)
        return o["theta"], o["v"]

    def step_csghmc(self, theta, g, v, xi, *, is_head, lr_body, lr_head, hp, should_sample):
        sc = self._sc(_lib.CSGHMC, hp, lr_body, lr_head, "true", add_noise=should_sample)
        o = self._run(_lib.CSGHMC, sc, theta, g, None, v, None, None, None, xi)
        return o["theta"], o["v"]

    def step_adam_sghmc(self, theta, g, theta0, v, m, s, buf, xi, *, is_head, P, lr_body, lr_head, hp, t, first_step,
                        div_mode="true"):
        sc = self._sc(_lib.ADAM_SGHMC, hp, lr_body, lr_head, div_mode, t=t, first_step=first_step)
        o = self._run(_lib.ADAM_SGHMC, sc, theta, g, theta0, v, m, s, buf if hp.mu != 0 else None, xi)
        return o["theta"], o["v"], o["m"], o["s"], (o["buf"] if hp.mu != 0 else buf)

    def step_adam_csghmc(self, theta, g, theta0, v, m, s, xi, *, is_head, P, lr_body, lr_head, hp, t, div_mode="true", coef=None):
        sc = self._sc(_lib.ADAM_CSGHMC, hp, lr_body, lr_head, div_mode, t=t)
        o = self._run(_lib.ADAM_CSGHMC, sc, theta, g, theta0, v, m, s, None, xi, coef=coef)
        return o["theta"], o["v"], o["m"], o["s"]
