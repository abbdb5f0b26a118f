"""Cyclical SGLD: drop-in for methods/csgld.py (``Runner`` :17-594, ``Model`` :597-682).

SGLD update with the cyclical cosine step size, per-cycle running moments, GMM-weighted ensemble, optional raw
sample store (``args.full_sample``).  hparams: prior_sig, Ninflate, nd, thin, nst, bias.
"""
from .. import _lib
from ._base import CyclicalRunner, FusedModel


class Model(FusedModel):
    VARIANT = _lib.SGLD

    def __init__(self, ND, prior_sig=1.0, bias="informative"):
        super().__init__(ND, prior_sig=prior_sig, bias=bias)


class Runner(CyclicalRunner):
    SGD_MOMENTUM_FROM_ARGS = True           # methods/csgld.py:48-52
    CAPTURE = "avg"
    LIKELIHOOD_MEAN = "theta"               # methods/csgld.py:518
    STORE_ALL_SAMPLES = True                # methods/csgld.py:278-279, 328-329
    TITLE = "Cyclical SGLD"

    def _build_model(self, hp):
        return Model(ND=self.args.ND, prior_sig=float(hp["prior_sig"]), bias=str(hp["bias"]))
