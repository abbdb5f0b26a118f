"""Cyclical SGHMC: drop-in for methods/csghmc.py (``Runner`` :17-670, ``Model`` :673-781).

update (fused kernel BDL_CSGHMC): v <- v(1-a) - lr*(g + prior_sig*theta) [+ nd*sqrt(2a*lr)/N*xi if sampling] ;
theta <- theta + v.  ``net0`` and the ``bias`` option are ignored exactly like the reference (Appendix B.1); per-cycle
moments use Welford's update with the reference's double-counted n (Appendix B.3).
"""
import time

from .. import _lib
from ._base import CyclicalRunner, FusedModel


class Model(FusedModel):
    VARIANT = _lib.CSGHMC

    def __init__(self, ND, runner=None, prior_sig=1.0, bias="informative", momentum_decay=0.05):
        # `bias` has no effect on this sampler: both branches of the reference are identical (csghmc.py:759-762) and
        # the BDL_CSGHMC kernel ignores the prior-mask class bit.
        super().__init__(ND, prior_sig=prior_sig, bias=bias, momentum_decay=momentum_decay)
        self.runner = runner


class Runner(CyclicalRunner):
    SGD_MOMENTUM_FROM_ARGS = False          # methods/csghmc.py:54-58; and no optimizer.step() at all (:304)
    CAPTURE = "welford"                     # methods/csghmc.py:333-345
    LIKELIHOOD_MEAN = "cycle_mean"          # methods/csghmc.py:578
    LAST_THETA_AS_VECTOR = True             # methods/csghmc.py:534
    PASS_SHOULD_SAMPLE = True               # methods/csghmc.py:296-300
    TITLE = "Cyclical SGHMC"

    def _build_model(self, hp):
        return Model(ND=self.args.ND, prior_sig=float(hp["prior_sig"]), runner=self, bias=str(hp["bias"]),
                     momentum_decay=float(hp["momentum_decay"]))

    def evaluate_point_estimate(self, data_loader, net_to_evaluate, desc_prefix="Point Estimate"):
        """Deterministic evaluation of a given network (methods/csghmc.py:211-244)."""
        return self._point_estimate(data_loader, net_to_evaluate)

    def _after_epoch(self, ep, val_loader):   # methods/csghmc.py:118-128
        if val_loader is not None and (ep % 5 == 0 or ep == self.args.epochs - 1):
            tic = time.time()
            loss, err = self.evaluate_point_estimate(val_loader, self.net,
                                                     desc_prefix=f"PE Val (Cycle {self.current_cycle} Mean)")
            self.net.train()
            self.logger.info(f"(Epoch {ep}) Point Estimate Val (Cycle {self.current_cycle} Mean): loss = {loss:.4f}, "
                             f"prediction error = {err:.4f} (time: {time.time() - tic:.4f} seconds)")
