"""CUDA-graph replay of an evaluation network's forward pass.

The posterior-predictive ensemble calls the same eval-mode backbone ``batches x cycles x nst`` times with nothing but the
parameter VALUES changing between calls -- and those live in one flat buffer that ``bdl_draw`` overwrites in place, so
every pointer a forward pass touches is stable.  That is exactly what a CUDA graph wants: the pass is captured once per
input shape and replayed, which removes the host's per-kernel launch cost (ResNet-101: ~350 launches; measured on B200,
``profiles/r01_probe_graph_forward.log``: 10.36 -> 9.13 ms at batch 64, 4.63 -> 3.19 ms at batch 16, outputs
bit-identical because the replay runs the very same kernels).  No tracing compiler, no second code path for the math.

Policy: the first call for an input shape runs eagerly (it is also the warm-up), the second one captures, later ones
replay.  A network whose forward cannot be captured (host-side control flow on tensor values, ...) is run eagerly from
then on, with one warning; ``hparams graph=0`` turns the mechanism off.
"""
import contextlib
import gc
import warnings

import torch


@contextlib.contextmanager
def capture_gc_guard():
    """Around every CUDA-graph capture of this package.  A ``torch.cuda.CUDAGraph`` that Python's CYCLIC collector happens to
    finalise while a stream is capturing calls ``cudaGraphExecDestroy`` in the middle of the capture, which CUDA answers with
    "operation not permitted when stream is capturing" and an INVALIDATED capture (seen in the test suite: runners of
    earlier tests, kept alive by reference cycles, own cached evaluation graphs; torch 2.11's ``torch.cuda.graph`` no
    longer runs ``gc.collect()`` on entry).  So: collect what is collectable BEFORE the capture begins, and keep the
    cyclic collector off until it has ended (reference-counted frees are unaffected)."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class GraphedForward:
    total_replays = 0               # process-wide counters (diagnostics, tests)
    total_captures = 0

    def __init__(self, net, enabled=True, pool=None):
        """``pool``: a ``torch.cuda.graph_pool_handle()`` shared with other GraphedForward objects whose replays never
        overlap in time (every replay's output is cloned before the next one runs, so they may reuse each other's
        activation memory); default: a private pool."""
        self.net = net
        self.enabled = bool(enabled) and torch.cuda.is_available()
        self._entries = {}          # (shape, dtype) -> dict(graph, x, out) | "eager" | int (eager calls so far)
        self._pool = pool
        self.replays = 0
        self.captures = 0

    def _capture(self, x):
        dev = x.device
        static_x = x.clone()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        # thread_local: the asynchronous checkpoint writer may allocate pinned memory on its own thread meanwhile
        with capture_gc_guard(), torch.cuda.graph(graph, pool=self._pool, capture_error_mode="thread_local"):
            out = self.net(static_x)
        if not isinstance(out, torch.Tensor) or out.device != dev:
            raise TypeError("forward did not return a tensor on the input's device")
        self.captures += 1
        GraphedForward.total_captures += 1
        return dict(graph=graph, x=static_x, out=out)

    def __call__(self, x):
        """net(x) -- a fresh tensor every call (replays clone the graph's static output).  The input is copied into the
        graph's static buffer on EVERY call: tensor identity says nothing about the values (a caller may refill one
        preallocated batch tensor in place), and the copy is noise next to the forward it feeds (38 MB = ~12 us on
        B200 against a 9 ms ResNet-101 pass)."""
        if not self.enabled or not x.is_cuda:
            return self.net(x)
        key = (tuple(x.shape), x.dtype, x.device.index)
        ent = self._entries.get(key, 0)
        if ent == "eager":
            return self.net(x)
        if isinstance(ent, int):
            if ent == 0:                                  # first sight of this shape: eager (and warm-up)
                self._entries[key] = 1
                return self.net(x)
            try:
                ent = self._capture(x)
            except Exception as e:                        # not capturable: eager from now on
                warnings.warn(f"CUDA-graph capture of the evaluation forward failed ({type(e).__name__}: {e}); "
                              f"running it eagerly")
                torch.cuda.synchronize()
                self._entries[key] = "eager"
                return self.net(x)
            self._entries[key] = ent
        ent["x"].copy_(x)                                 # always: never trust identity for freshness
        ent["graph"].replay()
        self.replays += 1
        GraphedForward.total_replays += 1
        return ent["out"].clone()
