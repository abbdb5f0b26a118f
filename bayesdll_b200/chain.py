"""Flat HBM state of one SG-MCMC chain and the host logic that feeds the fused kernels.

``ChainState`` owns the padded flat buffers (theta / theta0 / v / m / s / SGD momentum / optional flat
gradient / injected-noise scratch), re-points the network's parameters at views of ``theta`` and builds
the per-step run table.  It is the host-side mirror of what the reference keeps as per-tensor Python
objects: ``net.parameters()``, ``net0.parameters()``, ``Model.momentum_buffer`` / ``.m`` / ``.v`` dicts
and the torch SGD ``momentum_buffer`` state (methods/sghmc.py:462-465, methods/adam_sghmc.py:484-491).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from .flat import FlatLayout, adopt_parameters, alloc_flat

_RUN_DTYPE = np.dtype([("begin", "<u8"), ("end", "<u8"), ("valid_end", "<u8"), ("g_dev", "<u8"), ("cls", "<u4"),
                       ("reserved", "<u4")])
assert _RUN_DTYPE.itemsize == C.sizeof(_lib.Run)
_PTABLE_ROWS = 512          # bdl_step.cu kPRows: tables up to this size ride in the kernel arguments


class ChainState:
    def __init__(self, net, net0, *, variant, bias_mode, mu=0.0, device=None, noise="philox", seed=0,
                 grad_mode="table", div_mode=_lib.DIV_RECIP):
        params = [p for _, p in net.named_parameters()]
        device = device or params[0].device
        if device.type != "cuda":
            raise _lib.BdlError("bayesdll_b200 runs on CUDA devices only (no CPU fallback); got " + str(device))
        if any(p.dtype != torch.float32 for p in params):
            raise _lib.BdlError("all parameters must be fp32 (the reference is fp32 throughout)")
        _lib.load()
        self.variant, self.bias_mode, self.mu = variant, bias_mode, float(mu)
        self.device, self.noise_mode, self.seed = device, noise, int(seed)
        self.grad_mode, self.div_mode = grad_mode, div_mode
        self.layout = L = FlatLayout.from_module(net)
        self.params = params
        self.names = [n for n, _ in net.named_parameters()]
        n = L.n_padded

        self.theta = alloc_flat(n, device)
        self._theta_views = adopt_parameters(net, L, self.theta)
        self._view_ptrs = [v.data_ptr() for v in self._theta_views]
        self._probe = sorted({0, len(params) // 2, len(params) - 1})
        self.theta0 = None
        if variant != _lib.CSGHMC:                       # cSGHMC ignores net0 (Appendix B.1)
            self.theta0 = alloc_flat(n, device)
            with torch.no_grad():
                torch._foreach_copy_(L.views(self.theta0), [p.detach().to(device) for p in net0.parameters()])
        self.v = alloc_flat(n, device) if variant != _lib.SGLD else None
        adam = variant in (_lib.ADAM_SGHMC, _lib.ADAM_CSGHMC)
        self.m = alloc_flat(n, device) if adam else None
        self.s = alloc_flat(n, device) if adam else None
        self.buf = alloc_flat(n, device) if (self.mu != 0.0 and variant in (_lib.SGLD, _lib.ADAM_SGHMC)) else None
        self.g_flat = None
        self.xi = None
        self.step_count = 0            # Philox sub-sequence = number of updates applied so far
        self.sgd_steps = 0             # torch SGD creates its momentum buffer on the first step

        # static merged run table (flat gradient buffer mode) and per-tensor template (pointer-table mode)
        self._runs_merged = ops.upload_runs(L.run_table(bias_mode), device)
        tmpl = L.run_table(bias_mode, grad_ptrs=[0] * len(L.segments))
        self._run_np = np.frombuffer(bytes(tmpl), dtype=_RUN_DTYPE).copy()
        # two pinned staging buffers + events: the host may run ahead of the stream by one step
        self._run_pinned = [torch.empty(self._run_np.nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._run_dev = [torch.empty(self._run_np.nbytes, dtype=torch.uint8, device=device) for _ in range(2)]
        self._run_evt = [None, None]
        self._run_slot = 0
        self._table_sig = None

    def readopt_if_moved(self):
        """The kernels update the flat ``theta`` buffer; the network reads it through ``p.data`` views.  Code that
        re-points ``p.data`` (``vector_to_parameters``, ``p.data = ...``) would silently detach the two: probe three
        parameters per step, and when one moved bring every moved parameter's CURRENT values into the flat buffer and
        re-point it.  A dtype / device change cannot be adopted and raises."""
        if all(self.params[i].data_ptr() == self._view_ptrs[i] for i in self._probe):
            return False
        moved = 0
        with torch.no_grad():
            for p, view, ptr in zip(self.params, self._theta_views, self._view_ptrs):
                if p.data_ptr() == ptr:
                    continue
                if p.dtype != torch.float32 or p.device != self.device or p.shape != view.shape:
                    raise _lib.BdlError(
                        f"a sampled parameter was re-pointed at a {p.dtype} tensor of shape {tuple(p.shape)} on {p.device}: "
                        f"the chain's flat state is fp32 {tuple(view.shape)} on {self.device} (net.to / .half() after the "
                        f"first step is not supported; build a new Runner)")
                view.copy_(p.data)
                p.data = view
                moved += 1
        return moved > 0

    # ---- views handed to reference-style code ------------------------------------------------
    def named_views(self, flat):
        return dict(zip(self.names, self.layout.views(flat)))

    def _ensure_g_flat(self):
        if self.g_flat is None:
            self.g_flat = alloc_flat(self.layout.n_padded, self.device)
            self._g_views = self.layout.views(self.g_flat)
        return self.g_flat

    # ---- gradients ---------------------------------------------------------------------------
    def _gradient_table(self):
        """Run table whose rows point straight at autograd's gradient tensors: the kernel reads each p.grad in
        place, so no gather / flatten pass (8 B/param) precedes the update.  Tensors whose gradient cannot be
        read in place (non-contiguous, mis-aligned, wrong dtype) are copied into the flat gradient buffer; a
        tensor without gradient is skipped like the reference does (``if p.grad is not None``)."""
        rows = self._run_np
        params = self.params
        grads = [p.grad for p in params]
        n_none = sum(g is None for g in grads)
        if n_none:
            ptrs = np.array([0 if g is None else g.data_ptr() for g in grads], dtype=np.uint64)
        else:
            ptrs = np.array([g.data_ptr() for g in grads], dtype=np.uint64)
        # dtype / contiguity / device: autograd's gradient-layout contract makes a fresh .grad match its (contiguous,
        # fp32) parameter, so the per-tensor Python checks (the bulk of this function's host time) run when any
        # gradient ADDRESS differs from the last verified set (the caching allocator hands a steady-state training
        # loop the same blocks every step) and every 32nd step; alignment is checked every step (vectorised).
        self._grad_checks = getattr(self, "_grad_checks", 0) + 1
        bad = (ptrs & np.uint64(15)) != 0
        verified = getattr(self, "_verified_ptrs", None)
        ptrs_changed = verified is None or not np.array_equal(verified, ptrs)
        if ptrs_changed or self._grad_checks % 32 == 1 or n_none != getattr(self, "_last_none", 0):
            self._verified_ptrs = ptrs.copy()
            f32, dev = torch.float32, self.device
            for i, g in enumerate(grads):
                if g is not None and not (g.dtype is f32 and g.is_contiguous() and g.device == dev):
                    bad[i] = True
            self._grad_irregular = np.flatnonzero(bad)
        elif len(self._grad_irregular):
            bad[self._grad_irregular] = True
        self._last_none = n_none
        if bad.any():
            self._ensure_g_flat()
            for i in np.flatnonzero(bad):
                if grads[i] is not None:
                    self._g_views[i].copy_(grads[i])
                    ptrs[i] = 0
        rows["g_dev"] = ptrs
        if n_none or getattr(self, "_had_skip", False):
            skip = np.fromiter((g is None for g in grads), dtype=bool, count=len(grads))
            base_cls = rows["cls"] & ~np.uint32(_lib.CLS_SKIP)
            rows["cls"] = np.where(skip, base_cls | np.uint32(_lib.CLS_SKIP), base_cls)
            self._had_skip = bool(n_none)
        if len(rows) <= _PTABLE_ROWS and self.layout.n_padded < 2 ** 32:
            # the table travels inside the kernel arguments (step_ptable_kernel): nothing to stage on the device
            return None, len(rows)
        # larger tables live in device memory.  Unchanged table (gradients at fixed addresses, e.g. a CUDA-graph-captured
        # backward): the copy on the device is still valid
        sig = rows.tobytes()
        if sig == self._table_sig:
            return self._run_dev[self._run_slot], len(rows)
        self._table_sig = sig
        k = self._run_slot = self._run_slot ^ 1
        if self._run_evt[k] is not None:
            self._run_evt[k].synchronize()               # the copy that last used this staging buffer has run
        self._run_pinned[k].numpy()[:] = rows.view(np.uint8)
        self._run_dev[k].copy_(self._run_pinned[k], non_blocking=True)
        self._run_evt[k] = torch.cuda.Event()
        self._run_evt[k].record()
        return self._run_dev[k], len(rows)

    def _gradient_flat(self):
        self._ensure_g_flat()
        grads = [p.grad for p in self.params]
        if any(g is None for g in grads):
            return self._gradient_table()
        torch._foreach_copy_(self._g_views, grads)
        return self._runs_merged

    # ---- noise -------------------------------------------------------------------------------
    def _noise(self):
        if self.noise_mode == "philox":
            return ops.make_noise(seed=self.seed, subseq=self.step_count, stream_id=_lib.STREAM_STEP)
        if self.noise_mode == "torch":
            # Parity / debugging mode: one torch.randn_like per tensor in named_parameters() order, exactly the
            # calls the reference makes (methods/sghmc.py:501), so seeding or patching torch.randn_like drives both.
            if self.xi is None:
                self.xi = alloc_flat(self.layout.n_padded, self.device)
                self._xi_views = self.layout.views(self.xi)
            for view, p in zip(self._xi_views, self.params):
                if p.grad is not None:
                    view.copy_(torch.randn_like(view))
            return ops.make_noise(xi=self.xi)
        raise ValueError(f"unknown noise mode {self.noise_mode!r} (philox | torch)")

    # ---- the fused update --------------------------------------------------------------------
    CLIPPED_VARIANTS = (_lib.SGLD, _lib.ADAM_CSGHMC)   # the runners whose clipped p.grad reaches theta (csgld, adam_csghmc)

    def update(self, scalars, capture=None, clip=None):
        """Apply one fused update using the gradients currently held in ``p.grad``.  ``capture`` (ops.make_capture)
        folds the new theta into running moments in the same launch.  ``clip`` = args.clip_grad: the reference's
        ``clip_grad_norm_(net.parameters(), clip)`` between Model.forward and optimizer.step() (methods/csgld.py:250-251),
        as two passes with the coefficient kept on the device (no host sync)."""
        if clip is not None and self.variant in self.CLIPPED_VARIANTS:
            if capture is not None:
                raise _lib.BdlError("clipping and fused capture cannot be combined (capture after the step instead)")
            _, nruns = self._gradient_table()                # fills the per-tensor host table (pointers, skip flags)
            if getattr(self, "_clip_buf", None) is None:
                self._clip_sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
                self._clip_buf = torch.zeros(2, dtype=torch.float32, device=self.device)     # [coef, total_norm]
            scalars.div_mode = self.div_mode
            if self.buf is not None:
                scalars.first_step = int(self.sgd_steps == 0)
            nz = self._noise()                                # ONE draw for both passes
            host = self._run_np.ctypes.data
            state = (self.variant, self.theta, self.g_flat, self.theta0, self.v, self.m, self.s, self.buf, host, nruns, scalars, nz)
            self._clip_sumsq.zero_()
            ops.step_gradnorm(*state, self._clip_sumsq)
            ops.clip_coef(self._clip_sumsq, clip, self._clip_buf[0:1], self._clip_buf[1:2])
            ops.step_clipped(*state, self._clip_buf[0:1])
            self.step_count += 1
            self.sgd_steps += 1
            return
        runs_host = None
        if self.grad_mode == "table":
            runs_dev, nruns = self._gradient_table()
            runs_host = self._run_np.ctypes.data            # host copy: small tables ride in the kernel arguments
            g = self.g_flat
        else:
            runs_dev, nruns = self._gradient_flat()
            g = self.g_flat
            if runs_dev is None or not hasattr(runs_dev, "_bdl_host"):
                runs_host = self._run_np.ctypes.data        # fell back to the per-tensor table (a gradient was None)
        scalars.div_mode = self.div_mode
        if self.buf is not None:
            scalars.first_step = int(self.sgd_steps == 0)
        ops.step(self.variant, self.theta, g, self.theta0, self.v, self.m, self.s, self.buf, runs_dev, nruns,
                 scalars, self._noise(), capture=capture, runs_host=runs_host)
        self.step_count += 1
        self.sgd_steps += 1

    def snapshot(self, flat=None):
        """Device-side copy of a flat buffer (default theta) taken with ONE TMA ring-copy launch (8 B/param): what
        checkpoint / cycle-state / raw-sample writers hold while training carries on."""
        src = self.theta if flat is None else flat
        out = torch.empty_like(src)
        ops.capture_ring(src, out.view(1, -1), 0)
        return out

    def reset_momenta(self):
        for t in (self.v, self.m, self.s):
            if t is not None:
                t.zero_()


class SampleRing:
    """Preallocated HBM ring of raw posterior samples (north star (b)); replaces ``theta_vec.clone()`` into a Python
    dict (methods/csgld.py:278-279).  Each capture is one TMA bulk-copy kernel (bdl_capture_ring, 8 B/param) into the
    next slot; when the ring wraps the oldest sample is overwritten and its key dropped."""

    def __init__(self, layout, device, expected_samples, mem_fraction=0.25):
        free, _ = torch.cuda.mem_get_info(device)
        per_slot = layout.n_padded * 4
        cap = max(1, min(int(expected_samples), int(free * mem_fraction) // per_slot))
        self.layout = layout
        self.buf = torch.empty((cap, layout.n_padded), dtype=torch.float32, device=device)
        self.keys = [None] * cap
        self.count = 0

    @property
    def capacity(self):
        return self.buf.shape[0]

    def capture(self, theta_flat, key, store, before_overwrite=None):
        """Copy ``theta_flat`` into the next slot; ``store[key]`` becomes the dense view of that slot.
        ``before_overwrite(old_key)`` runs before a wrapped slot is reused (e.g. wait for its spill to disk)."""
        slot = self.count % self.capacity
        old = self.keys[slot]
        if old is not None:
            if before_overwrite is not None:
                before_overwrite(old)
            store.pop(old, None)
        ops.capture_ring(theta_flat, self.buf, slot)
        self.keys[slot] = key
        self.count += 1
        row = self.buf[slot]
        store[key] = row[:self.layout.n_dense] if self._tail_only() else self.layout.to_dense(row)
        return slot

    def _tail_only(self):
        segs = self.layout.segments
        return all(s.begin == s.dense_begin for s in segs)
