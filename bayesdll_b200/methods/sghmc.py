"""SGHMC: drop-in for methods/sghmc.py (``Runner`` :16-406, ``Model`` :409-512).

update (fused kernel BDL_SGHMC):  v <- v(1-a) + lr*(g + prior) + nd*sqrt(2a/(N*lr))*xi ; theta <- theta - lr*(g+v)
(the raw gradient enters twice, Appendix B.2).  hparams: prior_sig, Ninflate, nd, burnin, thin, nst, bias,
momentum_decay.  Checkpoints additionally carry ``momentum_buffer`` (dict name -> tensor, :382).
"""
from .. import _lib
from ._base import BurninRunner, FusedModel


class Model(FusedModel):
    VARIANT = _lib.SGHMC

    def __init__(self, ND, prior_sig=1.0, bias="informative", momentum_decay=0.05):
        super().__init__(ND, prior_sig=prior_sig, bias=bias, momentum_decay=momentum_decay)


class Runner(BurninRunner):
    SGD_MOMENTUM_FROM_ARGS = False          # SGD(momentum=0), methods/sghmc.py:53-57

    def _build_model(self, hp):
        return Model(ND=self.args.ND, prior_sig=float(hp["prior_sig"]), bias=str(hp["bias"]),
                     momentum_decay=float(hp["momentum_decay"]))

    def _ckpt_extra(self):
        return {"momentum_buffer": {k: v.clone() for k, v in self.model.momentum_buffer.items()}}

    def _load_ckpt_extra(self, ckpt):
        if "momentum_buffer" in ckpt:
            self.model.momentum_buffer = ckpt["momentum_buffer"]
