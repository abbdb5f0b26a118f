"""SGLD: drop-in for methods/sgld.py (``Runner`` :15-400, ``Model`` :401-487).

update (fused kernel BDL_SGLD):  g' = g + P*(theta-theta0)/sigma^2/N + nd*sqrt(2/(N*lr))*xi ;
torch SGD with momentum ``args.momentum`` folded in.  hparams: prior_sig, Ninflate, nd, burnin, thin, nst, bias.
"""
from .. import _lib
from ._base import BurninRunner, FusedModel


class Model(FusedModel):
    VARIANT = _lib.SGLD

    def __init__(self, ND, prior_sig=1.0, bias="informative"):
        super().__init__(ND, prior_sig=prior_sig, bias=bias)


class Runner(BurninRunner):
    SGD_MOMENTUM_FROM_ARGS = True           # methods/sgld.py:52-56

    def _build_model(self, hp):
        return Model(ND=self.args.ND, prior_sig=float(hp["prior_sig"]), bias=str(hp["bias"]))
