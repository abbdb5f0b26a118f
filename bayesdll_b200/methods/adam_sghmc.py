"""Adam-SGHMC: drop-in for methods/adam_sghmc.py (``Runner`` :16-420, ``Model`` :423-556).

update (fused kernel BDL_ADAM_SGHMC): Adam first/second moments of gU, bias correction with t, preconditioned
drift m_hat/(sqrt(s_hat)+eps), noise scale nd*sqrt(2a*pre/N), then theta <- SGD(momentum=args.momentum) on g+v.
hparams: SGHMC's plus beta1, beta2, epsilon.  Checkpoints carry momentum_buffer, m, v, t (:385-388) and -- like the
reference -- no ``last_theta``.
"""
from .. import _lib
from ._base import AdamStateMixin, BurninRunner, FusedModel


class Model(AdamStateMixin, FusedModel):
    VARIANT = _lib.ADAM_SGHMC

    def __init__(self, ND, prior_sig=1.0, bias="informative", momentum_decay=0.05, beta1=0.9, beta2=0.999,
                 epsilon=1e-8):
        super().__init__(ND, prior_sig=prior_sig, bias=bias, momentum_decay=momentum_decay, beta1=beta1, beta2=beta2,
                         epsilon=epsilon)


class Runner(BurninRunner):
    SGD_MOMENTUM_FROM_ARGS = True           # methods/adam_sghmc.py:57-61
    CKPT_HAS_LAST_THETA = False             # methods/adam_sghmc.py:379

    def _build_model(self, hp):
        return Model(ND=self.args.ND, prior_sig=float(hp["prior_sig"]), bias=str(hp["bias"]),
                     momentum_decay=float(hp["momentum_decay"]), beta1=float(hp.get("beta1", 0.9)),
                     beta2=float(hp.get("beta2", 0.999)), epsilon=float(hp.get("epsilon", 1e-8)))

    def _ckpt_extra(self):
        clone = lambda d: {k: v.clone() for k, v in d.items()}
        return {"momentum_buffer": clone(self.model.momentum_buffer), "m": clone(self.model.m),
                "v": clone(self.model.v), "t": self.model.t}

    def _load_ckpt_extra(self, ckpt):
        if "momentum_buffer" in ckpt:
            self.model.momentum_buffer = ckpt["momentum_buffer"]
        if "m" in ckpt:
            self.model.m = ckpt["m"]
        if "v" in ckpt:
            self.model.v = ckpt["v"]
        if "t" in ckpt:
            self.model.t = ckpt["t"]
