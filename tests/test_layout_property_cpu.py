"""CPU, hypothesis: properties of the padded flat layout and run tables for arbitrary (ragged) networks."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from bayesdll_b200 import _lib
from bayesdll_b200.flat import ALIGN, FlatLayout

tensor_st = st.tuples(st.sampled_from(["weight", "bias", "scale"]), st.integers(1, 300), st.booleans())


@settings(max_examples=60, deadline=None)
@given(st.lists(tensor_st, min_size=1, max_size=40), st.sampled_from(["informative", "uninformative"]))
def test_layout_and_runs_invariants(tensors, bias_mode):
    named = [(f"{'classifier' if head else 'layers'}.{i}.{kind}", (numel,)) for i, (kind, numel, head) in enumerate(tensors)]
    lay = FlatLayout(named, "classifier")
    # segments: aligned, ordered, tight, padding < ALIGN
    pos = 0
    for s in lay.segments:
        assert s.begin == pos and s.begin % ALIGN == 0 and s.end % ALIGN == 0
        assert 0 <= s.end - s.valid_end < ALIGN
        pos = s.end
    assert lay.n_padded == pos and lay.n_dense == sum(t[1] for t in tensors)
    # dense <-> padded round trip, padding stays zero
    dense = torch.arange(1, lay.n_dense + 1, dtype=torch.float32)
    flat = lay.from_dense(dense)
    assert torch.equal(lay.to_dense(flat), dense)
    assert flat.sum() == dense.sum()
    assert np.array_equal(lay.padded_numpy(dense.numpy()), flat.numpy())
    # run tables: sorted contiguous cover, classes constant inside a run, merged == per-element classes
    for merge in (True, False):
        runs = lay.runs(bias_mode, merge=merge)
        assert runs[0][0] == 0 and runs[-1][1] == lay.n_padded
        assert all(a[1] == b[0] for a, b in zip(runs, runs[1:]))
        assert all(b < ve <= e for b, e, ve, _ in runs)
        if merge:
            assert all(a[3] != b[3] for a, b in zip(runs, runs[1:]))
    is_head, P = lay.per_element(bias_mode)
    for b, e, _, c in lay.runs(bias_mode):
        assert (is_head[b:e] == bool(c & _lib.CLS_HEAD)).all() and (P[b:e] == (1.0 if c & _lib.CLS_PRIOR else 0.0)).all()
    for s in lay.segments:
        want_prior = not (s.is_bias and bias_mode == "uninformative")
        assert bool(lay.seg_cls(s, bias_mode) & _lib.CLS_PRIOR) == want_prior
        assert bool(lay.seg_cls(s, bias_mode) & _lib.CLS_HEAD) == s.is_head


@settings(max_examples=60, deadline=None)
@given(st.lists(tensor_st, min_size=1, max_size=40), st.sampled_from(["gaussian", "spikymix", "ignore"]))
def test_dropout_run_table_invariants(tensors, bias_mode):
    """Run table of bdl_dropout_mix: sorted contiguous cover of the padded layout, neighbouring runs differ in class,
    BDL_CLS_NODROP exactly on the bias tensors unless the mode treats them like weights (mc_dropout.py:383-389)."""
    named = [(f"layers.{i}.{kind}", (numel,)) for i, (kind, numel, _) in enumerate(tensors)]
    lay = FlatLayout(named, "classifier")
    tab = lay.dropout_run_table(bias_mode)
    rows = [(r.begin, r.end, r.valid_end, r.cls, r.g_dev) for r in tab]
    assert rows[0][0] == 0 and rows[-1][1] == lay.n_padded
    assert all(a[1] == b[0] and a[3] != b[3] for a, b in zip(rows, rows[1:]))
    assert all(b < ve <= e and b % ALIGN == 0 and g in (0, None) for b, e, ve, _, g in rows)
    nodrop = np.zeros(lay.n_padded, bool)
    for b, e, _, c, _ in rows:
        assert c in (0, _lib.CLS_NODROP)
        nodrop[b:e] = bool(c & _lib.CLS_NODROP)
    for s in lay.segments:
        assert nodrop[s.begin:s.end].all() == (s.is_bias and bias_mode != "spikymix")
        assert nodrop[s.begin:s.end].any() == (s.is_bias and bias_mode != "spikymix")
    if bias_mode == "spikymix":
        assert len(rows) == 1
