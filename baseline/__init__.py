"""Baseline implementations timed next to the product in bench.py (never imported by bayesdll_b200/).

``eager_port`` restates the reference's own *structure* for the sampler update -- a per-tensor Python loop of torch
eager ops followed by ``SGD.step`` (methods/sghmc.py:482-510, :229) -- so that it can be timed on the GPU box, where
the reference checkout does not exist: on the host cores ("what the reference's structure reaches on this CPU") and on
the same B200 ("the number the fused kernel replaces").  ``baseline/_ref/`` is reserved for a
driver-side install of the reference and is git-ignored.
"""
