"""Cyclical step-size schedule (drop-in for methods/cyclical.py:12-74).

Pure host scalar math in fp64; the resulting step size and the ``should_sample`` flag travel to the fused kernel as
scalar arguments.  The reference's quirks are preserved on purpose: ``calculate_lr`` uses
the *integer* cycle length ``K // M`` while the other three methods use the float ``K / M``, and the cosine keeps
decaying through the sampling phase.
"""
import numpy as np


class CyclicalSGMCMC:

    def __init__(self, base_lr, nbr_of_cycles, epochs, proportion_exploration=0.5):
        self.base_lr = base_lr
        self.number_of_cycles = nbr_of_cycles
        self.epochs = epochs
        self.proportion_exploration = proportion_exploration
        self.current_epoch = 0
        self.sample_at_bottom = True

    def _k(self, epoch, batch, batches_per_epoch):
        return epoch * batches_per_epoch + batch + 1

    def calculate_lr(self, epoch, batch, batches_per_epoch):
        cycle_length = (self.epochs * batches_per_epoch) // self.number_of_cycles
        pos = ((self._k(epoch, batch, batches_per_epoch) - 1) % cycle_length) / cycle_length
        return self.base_lr * (1 + np.cos(pos * np.pi)) / 2

    def should_sample(self, epoch, batch, batches_per_epoch):
        if not self.sample_at_bottom:
            return True
        cycle_length = (self.epochs * batches_per_epoch) / self.number_of_cycles
        pos = ((self._k(epoch, batch, batches_per_epoch) - 1) % cycle_length) / cycle_length
        return pos >= self.proportion_exploration

    def last_in_cycle(self, epoch, batch, batches_per_epoch):
        cycle_length = (self.epochs * batches_per_epoch) / self.number_of_cycles
        return (self._k(epoch, batch, batches_per_epoch) % cycle_length) == 0

    def get_cycle_number(self, epoch, batch, batches_per_epoch):
        cycle_length = (self.epochs * batches_per_epoch) / self.number_of_cycles
        return int((self._k(epoch, batch, batches_per_epoch) - 1) // cycle_length) + 1
