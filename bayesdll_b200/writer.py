"""Asynchronous writers for the on-disk side of the sampler path.

The reference writes checkpoints, per-cycle ``state_dict`` copies and raw samples with synchronous ``torch.save``
calls in the middle of the training loop (methods/sghmc.py:370-388, methods/csgld.py:311,458-490,
methods/csghmc_fs.py:176-181): the GPU idles while 0.2-20 GB go through a device->host copy and the pickler.  Here

  * the caller takes a device-side snapshot of whatever is still mutable (theta: one TMA ring-copy launch,
    ``ChainState.snapshot``), everything else a checkpoint holds is immutable once written (finished cycles);
  * ``AsyncWriter.submit`` records an event on the compute stream and returns; a writer thread waits for that event
    on a side stream, moves the object to (pinned) host memory there and serialises it -- same keys, same dense
    ``parameters_to_vector`` layout, same file names as the reference;
  * files appear atomically (``<name>.tmp`` + ``os.replace``); ``flush()`` joins outstanding jobs and re-raises the
    first failure.  Runners flush at the end of ``train()``, before ``load_ckpt`` and before reading spilled samples.

``FlatBackedStateDict`` is a ``state_dict`` whose parameter entries are views of one padded flat buffer: it is what
``net.state_dict()`` looks like after one snapshot launch instead of ~300 per-tensor clones, and it moves to the host
with one copy.
"""
import atexit
import os
import queue
import threading
import weakref
from collections import OrderedDict

import torch

PIN_LIMIT_BYTES = 4 << 30      # stage through pinned memory up to this much per job; larger jobs use pageable copies

_LIVE_WRITERS = weakref.WeakSet()


@atexit.register
def _close_all_writers():
    """Join every writer thread before the interpreter starts tearing down: a daemon thread that wakes up during
    finalisation inside torch / CUDA code aborts the process ('terminate called without an active exception')."""
    for w in list(_LIVE_WRITERS):
        try:
            w.close()
        except Exception:
            pass


class FlatBackedStateDict(OrderedDict):
    """``state_dict`` (same keys, same order as ``net.state_dict()``) whose parameter tensors are views of ``flat``
    (padded layout) and whose buffer tensors (BatchNorm statistics ...) are independent clones."""

    flat = None
    layout = None

    @classmethod
    def snapshot(cls, net, layout, names, flat_copy):
        views = dict(zip(names, layout.views(flat_copy)))
        out = cls()
        out.flat, out.layout = flat_copy, layout
        for k, v in net.state_dict().items():
            out[k] = views[k] if k in views else v.detach().clone()
        out._param_keys = [k for k in out if k in views]
        return out

    def to_host(self, pin):
        """Plain OrderedDict of CPU tensors; the flat buffer crosses PCIe once."""
        host_flat = _host_copy(self.flat, pin)
        views = dict(zip([s.name for s in self.layout.segments], self.layout.views(host_flat)))
        pk = set(self._param_keys)
        return OrderedDict((k, views[k] if k in pk else _host_copy(v, pin)) for k, v in self.items())

    def __reduce__(self):                       # pickles as an ordinary OrderedDict (torch.save of cycle_states)
        return (OrderedDict, (list(self.items()),))


def _host_copy(t, pin):
    if not t.is_cuda:
        return t
    dst = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=pin)
    dst.copy_(t, non_blocking=pin)
    return dst


def _nbytes(obj):
    if isinstance(obj, FlatBackedStateDict):
        return obj.flat.numel() * 4
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size() if obj.is_cuda else 0
    if isinstance(obj, dict):
        return sum(_nbytes(v) for v in obj.values())
    if isinstance(obj, (list, tuple)):
        return sum(_nbytes(v) for v in obj)
    return 0


def to_host(obj, pin=False):
    """Recursively replace CUDA tensors by CPU copies (on the current stream)."""
    if isinstance(obj, FlatBackedStateDict):
        return obj.to_host(pin)
    if isinstance(obj, torch.Tensor):
        return _host_copy(obj, pin)
    if isinstance(obj, OrderedDict):
        return OrderedDict((k, to_host(v, pin)) for k, v in obj.items())
    if isinstance(obj, dict):
        return {k: to_host(v, pin) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_host(v, pin) for v in obj)
    return obj


def _atomic_save(obj, path, serializer):
    tmp = f"{path}.tmp.{os.getpid()}"
    serializer(obj, tmp)
    os.replace(tmp, path)


class AsyncWriter:
    """One writer thread (+ one side stream on CUDA devices) per runner.  ``mode='sync'`` performs the same work inline
    (the reference's behaviour: the file exists when ``save_ckpt`` returns)."""

    def __init__(self, device, mode="async"):
        if mode not in ("async", "sync"):
            raise ValueError(f"io mode must be 'async' or 'sync', got {mode!r}")
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:        # args.device is often a bare 'cuda'
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.mode = mode
        self._q = None
        self._thread = None
        self._stream = None
        self._error = None
        self.pending_paths = set()
        self._lock = threading.Lock()
        self.stats = {"jobs": 0, "bytes": 0}

    # ---- public -----------------------------------------------------------------------------------------
    def submit(self, path, obj, serializer=torch.save):
        """Serialise ``obj`` (a nest of dicts / lists / tensors) to ``path``.  Device tensors inside ``obj`` must not be
        written by the caller any more (snapshot mutable state first)."""
        if self._error is not None:
            self._raise()
        nbytes = _nbytes(obj)
        self.stats["jobs"] += 1
        self.stats["bytes"] += nbytes
        if self.mode == "sync":
            _atomic_save(to_host(obj), path, serializer)
            return path
        self._start()
        ev = None
        if self.device.type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        with self._lock:
            self.pending_paths.add(path)
        self._q.put((path, obj, serializer, ev, nbytes))
        return path

    def flush(self):
        """Wait for every submitted file; re-raise the first failure.  Never waits on a dead writer thread."""
        q = self._q
        if q is not None:
            with q.all_tasks_done:
                while q.unfinished_tasks:
                    if self._thread is None or not self._thread.is_alive():
                        break
                    q.all_tasks_done.wait(timeout=0.2)
            if q.unfinished_tasks and self._error is None:
                self._error = RuntimeError("writer thread exited with jobs outstanding")
        if self._error is not None:
            self._raise()

    def close(self):
        if self._thread is not None:
            try:
                self.flush()
            finally:
                self._q.put(None)
                self._thread.join(timeout=10)
                self._thread = self._q = None
        if self._error is not None:
            self._raise()

    def is_pending(self, path):
        with self._lock:
            return path in self.pending_paths

    # ---- internals ----------------------------------------------------------------------------------------
    def _raise(self):
        err, self._error = self._error, None
        raise RuntimeError(f"asynchronous write failed: {err!r}") from err

    def _start(self):
        if self._thread is not None and self._thread.is_alive():
            return
        if self._thread is not None:                 # a previous thread died: surface that instead of queueing forever
            self._error = self._error or RuntimeError("writer thread is not running")
            self._raise()
        self._q = queue.Queue()
        if self.device.type == "cuda":
            self._stream = torch.cuda.Stream(self.device)
        self._thread = threading.Thread(target=self._run, name="bdl-writer", daemon=True)
        self._thread.start()
        _LIVE_WRITERS.add(self)

    def _run(self):
        try:
            if self.device.type == "cuda":
                torch.cuda.set_device(self.device)
            while True:
                job = self._q.get()
                if job is None:
                    self._q.task_done()
                    return
                self._one(*job)
        except BaseException as e:                  # anything outside a job: remember it, flush() reports it
            if self._error is None:
                self._error = e

    def _one(self, path, obj, serializer, ev, nbytes):
        try:
            if self.device.type == "cuda":
                with torch.cuda.stream(self._stream):
                    self._stream.wait_event(ev)
                    host = to_host(obj, pin=nbytes <= PIN_LIMIT_BYTES)
                    self._stream.synchronize()
            else:
                host = to_host(obj)
            del obj
            _atomic_save(host, path, serializer)
        except BaseException as e:                  # surfaced by the next submit() / flush()
            if self._error is None:
                self._error = e
        finally:
            with self._lock:
                self.pending_paths.discard(path)
            self._q.task_done()

    def __del__(self):
        try:
            if self._thread is not None and self._q is not None and self._thread.is_alive():
                self._q.put(None)
                self._thread.join(timeout=5)
        except Exception:
            pass
