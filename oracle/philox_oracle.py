"""Philox4x32-10 + Box-Muller N(0,1) stream, CPU restatement (thin alias over c_oracle)."""
from .c_oracle import philox4x32_10, philox_normal as normal  # noqa: F401
