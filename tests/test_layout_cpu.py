"""CPU: host logic -- padded flat layout, run tables, cyclical schedule, scalar preparation."""
import ctypes

import numpy as np
import pytest
import torch

import golden_util as gu
from bayesdll_b200 import _lib, ops, shapes
from bayesdll_b200.flat import ALIGN, FlatLayout
from bayesdll_b200.methods.cyclical import CyclicalSGMCMC


def test_layout_offsets_alignment_and_roundtrip():
    lay = FlatLayout([("a.weight", (3, 5)), ("a.bias", (5,)), ("classifier.weight", (2, 7)), ("classifier.bias", (2,))],
                     "classifier")
    assert [s.begin for s in lay.segments] == [0, 16, 24, 40]
    assert all(s.begin % ALIGN == 0 and s.end % ALIGN == 0 for s in lay.segments)
    assert lay.n_dense == 15 + 5 + 14 + 2 and lay.n_padded == 44
    dense = torch.arange(lay.n_dense, dtype=torch.float32)
    flat = lay.from_dense(dense)
    assert torch.equal(lay.to_dense(flat), dense)
    assert flat[15] == 0 and flat[21:24].abs().sum() == 0            # padding stays zero
    views = lay.views(flat)
    assert [tuple(v.shape) for v in views] == [(3, 5), (5,), (2, 7), (2,)]
    assert views[2][0, 0] == dense[20]
    # numpy helpers agree with the torch ones
    assert np.array_equal(lay.dense_numpy(flat.numpy()), dense.numpy())
    assert np.array_equal(lay.padded_numpy(dense.numpy()), flat.numpy())


def test_run_tables_merge_by_class():
    lay = FlatLayout([("l.0.weight", (10,)), ("l.0.bias", (3,)), ("l.1.weight", (8,)), ("fc.weight", (6,)), ("fc.bias", (2,))], "fc")
    inf = lay.runs("informative")
    assert [(b, e, c) for b, e, _, c in inf] == [(0, 24, _lib.CLS_PRIOR), (24, 36, _lib.CLS_PRIOR | _lib.CLS_HEAD)]
    uni = lay.runs("uninformative")
    assert [c for *_, c in uni] == [_lib.CLS_PRIOR, 0, _lib.CLS_PRIOR, _lib.CLS_PRIOR | _lib.CLS_HEAD, _lib.CLS_HEAD]
    assert uni[0][0] == 0 and uni[-1][1] == lay.n_padded
    assert all(a[1] == b[0] for a, b in zip(uni, uni[1:]))            # contiguous cover
    tab = lay.run_table("informative", grad_ptrs=[16 * (i + 1) for i in range(5)])
    assert len(tab) == 5 and tab[1].g_dev == 32 and tab[1].valid_end == tab[1].begin + 3
    is_head, P = lay.per_element("uninformative")
    assert is_head.sum() == 12 and P[12:16].sum() == 0 and P[:10].all()


def test_backbone_sizes_match_survey():
    for name, K, n, ntens in (("mlp_mnist", 10, 2_797_010, 8), ("resnet101", 37, 42_575_973, 314),
                              ("vit_l_32", 37, 305_548_325, 296)):
        named, readout = shapes.named_shapes(name, K)
        lay = FlatLayout(named, readout)
        assert lay.n_dense == n and len(lay.segments) == ntens
        assert lay.n_padded - lay.n_dense < 4 * ntens


def test_cyclical_schedule_matches_reference_golden():
    z = np.load(gu.golden_path("cyclical"))
    for (epochs, B, M, beta, lr0, ep, b, lr, ss, lic, cyc) in z["rows"]:
        s = CyclicalSGMCMC(lr0, int(M), int(epochs), beta)
        kw = dict(epoch=int(ep), batch=int(b), batches_per_epoch=int(B))
        assert s.calculate_lr(**kw) == lr
        assert float(s.should_sample(**kw)) == ss and float(s.last_in_cycle(**kw)) == lic
        assert s.get_cycle_number(**kw) == cyc


def test_cyclical_quirk_integer_vs_float_cycle_length():
    """K=11500, M=8: K/M is not an integer -> last_in_cycle only fires for cycles 2,4,6,8."""
    s = CyclicalSGMCMC(1e-4, 8, 100, 0.5)
    fired = [s.get_cycle_number(e, b, 115) for e in range(100) for b in range(115) if s.last_in_cycle(e, b, 115)]
    assert fired == [2, 4, 6, 8]


def test_make_scalars_rounding_follows_reference_expressions():
    f = np.float32
    ND, Ninf, lr, lrh, a, nd = 1840, 1e3, 1e-4, 1e-2, 0.18, 0.7
    N = ND * Ninf
    sc = ops.make_scalars(_lib.SGHMC, lr_body=lr, lr_head=lrh, ND=ND, Ninflate=Ninf, prior_sig=0.9, nd=nd, alpha=a)
    assert sc.noise_scale[0] == f(nd * np.sqrt(2 * a / (N * lr))) and sc.noise_scale[1] == f(nd * np.sqrt(2 * a / (N * lrh)))
    assert sc.one_minus_alpha == f(1 - a) and sc.sig2 == f(0.9 ** 2) and sc.N == f(N) and sc.mu == 0
    sc = ops.make_scalars(_lib.SGLD, lr_body=lr, lr_head=lrh, ND=ND, Ninflate=Ninf, nd=nd, mu=0.5)
    assert sc.noise_scale[0] == f(nd * np.sqrt(2 / (N * lr))) and sc.mu == f(0.5)
    sc = ops.make_scalars(_lib.CSGHMC, lr_body=lr, lr_head=lrh, ND=ND, Ninflate=Ninf, prior_sig=0.9, nd=nd, alpha=a)
    assert sc.noise_scale[1] == f(nd * np.sqrt(2 * a * lrh) / N) and sc.sig2 == f(0.9)
    sc = ops.make_scalars(_lib.ADAM_CSGHMC, lr_body=lr, lr_head=lrh, ND=ND, beta1=0.9, beta2=0.999, t=7, alpha=a,
                          temperature=1.5)
    assert sc.bias_corr1 == f(1 - 0.9 ** 7) and sc.bias_corr2 == f(1 - 0.999 ** 7) and sc.two_alpha == f(2 * a)
    assert sc.temperature == f(1.5)
