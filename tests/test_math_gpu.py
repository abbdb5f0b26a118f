"""GPU: the branch-free fast paths of the correctly rounded sqrt / reciprocal / quotient (bdl_common.cuh) equal the CUDA
intrinsics bit for bit on EVERY fp32 input their range flag admits -- the arithmetic contract of the Adam update rules
(methods/adam_sghmc.py:536-541: sqrt, 1.0 / x and m_hat / x are correctly rounded in the reference)."""
import pytest

pytestmark = pytest.mark.gpu


def test_optimistic_sqrt_rcp_div_equal_the_intrinsics_for_every_float(cuda_device):
    from bayesdll_b200 import ops
    res = ops.selftest_math(cuda_device)
    assert res["mismatch"] == [0, 0, 0], res
    # the fast paths are the common case, not a corner: > 44 % of all bit patterns for sqrt (every normal positive float
    # above 2^-100), > 97 % for the reciprocal, and a sizeable share of random quotients
    total = 2 ** 32
    assert res["fast"][0] > 0.44 * total and res["fast"][1] > 0.97 * total and res["fast"][2] > 0.05 * total, res
