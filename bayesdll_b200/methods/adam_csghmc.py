"""Cyclical Adam-SGHMC: drop-in for methods/adam_csghmc.py (``Runner`` :17-727, ``Model`` :730-863).

update (fused kernel BDL_ADAM_CSGHMC): as Adam-SGHMC with gU = g/T + prior and theta <- theta - lr*v (p.grad = v,
SGD momentum 0).  Momentum and Adam state are zeroed at every end of cycle (:370-378, :131-143); optional cold
restarts re-initialise the network (:102-129, :408-410).  hparams add temperature, perform_cold_restarts.
"""
from .. import _lib
from ._base import AdamStateMixin, CyclicalRunner, FusedModel, reinitialize_fresh


class Model(AdamStateMixin, FusedModel):
    VARIANT = _lib.ADAM_CSGHMC

    def __init__(self, ND, prior_sig=1.0, bias="informative", momentum_decay=0.05, beta1=0.9, beta2=0.999,
                 epsilon=1e-8, temperature=1.0):
        super().__init__(ND, prior_sig=prior_sig, bias=bias, momentum_decay=momentum_decay, beta1=beta1, beta2=beta2,
                         epsilon=epsilon, temperature=temperature)


class Runner(CyclicalRunner):
    SGD_MOMENTUM_FROM_ARGS = False          # methods/adam_csghmc.py:70-75
    CAPTURE = "avg"                         # methods/adam_csghmc.py:349-357
    LIKELIHOOD_MEAN = "cycle_mean"          # methods/adam_csghmc.py:639
    LAST_THETA_AS_VECTOR = True
    RESET_ADAM_AT_CYCLE_END = True
    TITLE = "Cyclical SGHMC"

    def __init__(self, net, net0, args, logger):
        hp = args.hparams
        self.temperature = float(hp.get("temperature", 1.0))
        self.perform_cold_restarts = str(hp.get("perform_cold_restarts", False)).lower() == "true"
        logger.info("Performing cold restarts: re-initializing network parameters with fresh random weights at the "
                    "start of each cycle." if self.perform_cold_restarts else
                    "Cold restarts disabled: keeping network parameters across cycles.")
        super().__init__(net, net0, args, logger)

    def _build_model(self, hp):
        return Model(ND=self.args.ND, prior_sig=float(hp["prior_sig"]), bias=str(hp["bias"]),
                     momentum_decay=float(hp["momentum_decay"]), beta1=float(hp.get("beta1", 0.9)),
                     beta2=float(hp.get("beta2", 0.999)), epsilon=float(hp.get("epsilon", 1e-8)),
                     temperature=self.temperature)

    def _reinitialize_network_fresh(self):
        reinitialize_fresh(self.net, self.logger)

    def _reset_optimizer_states(self, log=True):
        """v, m, s <- 0 and t <- 0 (methods/adam_csghmc.py:131-143): three memsets on the flat buffers."""
        if self.model.chain is not None:
            self.model.chain.reset_momenta()
        self.model.t = 0
        if log:
            self.logger.info("All optimizer states (momentum, m, v, t) reset for new cycle.")

    def evaluate_simple(self, test_loader):
        """Deterministic pass with the live network (methods/adam_csghmc.py:544-575)."""
        out = self._point_estimate(test_loader, self.net)
        self.net.train()
        return out

    def _before_cycle_eval(self, val_loader):   # methods/adam_csghmc.py:185-188
        if val_loader is not None:
            loss, err = self.evaluate_simple(val_loader)
            self.logger.info(f"Point estimation on validation set: loss = {loss:.4f}, error = {err:.4f}")

    def _after_cycle_completed(self, cycle_number):   # methods/adam_csghmc.py:404-413
        self._reset_optimizer_states()
        if self.perform_cold_restarts and cycle_number >= 1:
            self.logger.info(f"Performing COLD RESTART: Fresh random weight initialization for cycle {cycle_number + 1}")
            self._reinitialize_network_fresh()
        else:
            self.logger.info(f"Standard cycle transition: keeping weights, optimizer states reset for cycle "
                             f"{cycle_number + 1}")
