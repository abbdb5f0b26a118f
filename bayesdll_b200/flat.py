"""Padded flat layout of a network's parameters in HBM (DESIGN.md "Data layout").

Tensors are laid out in ``named_parameters()`` order (the order of ``nn.utils.parameters_to_vector``,
methods/sgld.py:98), each tensor start rounded up to a multiple of ``ALIGN`` elements so every tensor
begins on a 16-byte boundary and 128-bit vector accesses never straddle two tensors.  Padding elements
are ordinary elements that carry g = 0 and theta0 = 0 and are never visible through the views.

Only host logic lives here (offsets, run tables, views); it is importable without CUDA.
"""
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import _lib

ALIGN = 4  # elements (16 bytes)


def _round_up(x, a):
    return (x + a - 1) // a * a


@dataclass(frozen=True)
class Segment:
    name: str
    shape: tuple
    numel: int
    begin: int      # padded-flat offset (multiple of ALIGN)
    end: int        # begin of the next segment (multiple of ALIGN)
    dense_begin: int
    is_head: bool
    is_bias: bool

    @property
    def valid_end(self):
        return self.begin + self.numel


class FlatLayout:
    """Offsets and element classes for one network architecture."""

    def __init__(self, named_shapes: Sequence, readout_name: str):
        segs = []
        pos = dense = 0
        for name, shape in named_shapes:
            numel = int(np.prod(shape)) if len(shape) else 1
            end = _round_up(pos + numel, ALIGN)
            segs.append(Segment(name, tuple(shape), numel, pos, end, dense,
                                is_head=(readout_name in name), is_bias=("bias" in name)))
            pos, dense = end, dense + numel
        self.segments: List[Segment] = segs
        self.n_padded = pos
        self.n_dense = dense
        self.readout_name = readout_name
        if self.n_padded // 4 >= 0xFFFFFFFF:
            raise ValueError("flat buffer too large for 32-bit group indices")

    @classmethod
    def from_module(cls, net):
        return cls([(n, tuple(p.shape)) for n, p in net.named_parameters()], net.readout_name)

    # ---- element classes --------------------------------------------------------------------
    def seg_cls(self, seg: Segment, bias_mode: str) -> int:
        """BDL_CLS_* bits.  Head: ``readout_name in pname`` (methods/sghmc.py:485-488).  Prior pull is
        dropped for ``'bias' in pname and bias == 'uninformative'`` (methods/sghmc.py:494)."""
        c = _lib.CLS_HEAD if seg.is_head else 0
        if not (seg.is_bias and bias_mode == "uninformative"):
            c |= _lib.CLS_PRIOR
        return c

    def runs(self, bias_mode: str, merge: bool = True):
        """List of (begin, end, valid_end, cls) covering [0, n_padded).  merge=True coalesces neighbouring
        segments of equal class (2 runs for bias='informative'); merge=False keeps one run per tensor
        (needed when every run carries its own gradient pointer)."""
        out = []
        for s in self.segments:
            c = self.seg_cls(s, bias_mode)
            if merge and out and out[-1][3] == c:
                b, _, _, _ = out[-1]
                out[-1] = (b, s.end, s.valid_end, c)
            else:
                out.append((s.begin, s.end, s.valid_end, c))
        if len(out) > _lib.MAX_RUNS:
            raise ValueError(f"{len(out)} runs exceed BDL_MAX_RUNS={_lib.MAX_RUNS}")
        return out

    def run_table(self, bias_mode: str, grad_ptrs=None):
        """ctypes array of bdl_run.  grad_ptrs: optional list (one per segment) of device addresses
        (0 -> read the flat gradient buffer for that tensor); implies one run per tensor."""
        runs = self.runs(bias_mode, merge=grad_ptrs is None)
        arr = (_lib.Run * len(runs))()
        for i, (b, e, ve, c) in enumerate(runs):
            arr[i].begin, arr[i].end, arr[i].valid_end, arr[i].cls = b, e, ve, c
            arr[i].g_dev = 0 if grad_ptrs is None else int(grad_ptrs[i])
        return arr

    def dropout_run_table(self, bias_mode: str):
        """Run table for ``ops.dropout_mix``: bias tensors carry BDL_CLS_NODROP in the 'gaussian' / 'ignore' bias modes
        (z = ones_like(p), methods/mc_dropout.py:383-389); 'spikymix' drops biases like weights."""
        if bias_mode not in ("gaussian", "spikymix", "ignore"):
            raise KeyError(bias_mode)
        runs = []
        for s in self.segments:
            c = _lib.CLS_NODROP if (s.is_bias and bias_mode != "spikymix") else 0
            if runs and runs[-1][3] == c:
                runs[-1] = (runs[-1][0], s.end, s.valid_end, c)
            else:
                runs.append((s.begin, s.end, s.valid_end, c))
        if len(runs) > _lib.MAX_RUNS:
            raise ValueError(f"{len(runs)} runs exceed BDL_MAX_RUNS={_lib.MAX_RUNS}")
        arr = (_lib.Run * len(runs))()
        for i, (b, e, ve, c) in enumerate(runs):
            arr[i].begin, arr[i].end, arr[i].valid_end, arr[i].cls, arr[i].g_dev = b, e, ve, c, 0
        return arr

    def per_element(self, bias_mode: str):
        """(is_head, P) fp32/bool arrays over the padded layout (padding inherits its tensor's class).
        Used by tests to drive the oracle on the same padded buffers."""
        is_head = np.zeros(self.n_padded, bool)
        P = np.zeros(self.n_padded, np.float32)
        for s in self.segments:
            c = self.seg_cls(s, bias_mode)
            is_head[s.begin:s.end] = bool(c & _lib.CLS_HEAD)
            P[s.begin:s.end] = 1.0 if c & _lib.CLS_PRIOR else 0.0
        return is_head, P

    # ---- dense <-> padded -------------------------------------------------------------------
    def views(self, flat: torch.Tensor):
        assert flat.numel() == self.n_padded
        return [flat[s.begin:s.begin + s.numel].view(s.shape) for s in self.segments]

    def flat_views(self, flat: torch.Tensor):
        return [flat[s.begin:s.begin + s.numel] for s in self.segments]

    def to_dense(self, flat: torch.Tensor) -> torch.Tensor:
        """Dense unpadded vector in parameters_to_vector order (checkpoint contract)."""
        if self.n_dense == self.n_padded:
            return flat.clone()
        return torch.cat(self.flat_views(flat))

    def from_dense(self, dense: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        assert dense.numel() == self.n_dense
        if out is None:
            out = torch.zeros(self.n_padded, dtype=dense.dtype, device=dense.device)
        dense = dense.reshape(-1)
        if self.n_dense == self.n_padded:
            out.copy_(dense)
            return out
        pieces = list(torch.split(dense, [s.numel for s in self.segments]))
        torch._foreach_copy_(self.flat_views(out), pieces)
        return out

    def dense_numpy(self, flat_np: np.ndarray) -> np.ndarray:
        return np.concatenate([flat_np[s.begin:s.begin + s.numel] for s in self.segments])

    def padded_numpy(self, dense_np: np.ndarray, fill=0.0) -> np.ndarray:
        out = np.full(self.n_padded, fill, dtype=dense_np.dtype)
        for s in self.segments:
            out[s.begin:s.begin + s.numel] = dense_np[s.dense_begin:s.dense_begin + s.numel]
        return out


def alloc_flat(n, device, zero=True):
    """fp32 flat buffer; torch's caching allocator returns >= 512-byte aligned blocks."""
    t = torch.zeros(n, dtype=torch.float32, device=device) if zero else \
        torch.empty(n, dtype=torch.float32, device=device)
    assert t.data_ptr() % 16 == 0
    return t


def adopt_parameters(net, layout: FlatLayout, flat: torch.Tensor):
    """Copy ``net``'s parameters into ``flat`` and re-point every ``p.data`` at its view, so eager code
    (forward/backward, state_dict, deepcopy, parameters_to_vector) keeps working while the kernels see one
    contiguous buffer (hard part 2)."""
    views = layout.views(flat)
    with torch.no_grad():
        params = [p for _, p in net.named_parameters()]
        torch._foreach_copy_(views, [p.data for p in params])
        for p, v in zip(params, views):
            p.data = v
    return views
