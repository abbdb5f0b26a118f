"""Loader + injection harness for the *real* reference (omarezz46/BayesDLL).

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under ``bayesdll_b200/`` may import this.
It is used in the build container (where ``/root/reference`` is mounted) by
``oracle/make_golden.py`` to run the reference's own PyTorch code on injected
gradients / injected noise and record golden vectors under ``tests/golden/``, and by
``baseline/reference_arm.py`` (bench.py's reference arm / cpu_baseline legs) to TIME the
unmodified reference.  ``/root/reference`` does not exist on the GPU box: there the loader
finds the install ``baseline/install_ref.py`` leaves in ``baseline/_ref/`` (git-ignored, travels
with the gpurun snapshot); nothing in ``tests -m gpu`` or ``smoke()`` needs it.

Two shims are needed to import the reference unmodified:
  * a stub ``matplotlib`` (imported at calibration.py:14-15, only used by the plots)
  * ``<ref>/src`` on sys.path so ``from bayesdll import calibration``
    (methods/cyclical.py:10) resolves.
"""
import contextlib
import importlib
import os
import sys
import types

import torch
import torch.nn as nn


_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    """$BDL_REF, the mounted checkout, then the install under baseline/_ref/  -- in that order."""
    for cand in (os.environ.get("BDL_REF"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "methods", "sghmc.py")):
            return cand
    raise FileNotFoundError("reference not found: set $BDL_REF, mount /root/reference, or run baseline/install_ref.py "
                            "where the checkout is mounted (it fills baseline/_ref/)")


def _install_matplotlib_stub():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    pyplot = types.ModuleType("matplotlib.pyplot")
    patches = types.ModuleType("matplotlib.patches")

    class _Anything:
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __getitem__(self, k):
            return _Anything()

        def __iter__(self):
            return iter(())

    def _mod_getattr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    pyplot.__getattr__ = _mod_getattr
    patches.__getattr__ = _mod_getattr
    mpl.pyplot = pyplot
    mpl.patches = patches
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = pyplot
    sys.modules["matplotlib.patches"] = patches


_REF_MODULE_NAMES = ("calibration", "methods", "networks", "bayesdll", "utils", "datasets")


@contextlib.contextmanager
def reference_imports():
    """Context manager: inside it ``import methods.sghmc`` etc. resolve to the reference.

    On exit the reference's top-level module names are removed from ``sys.modules``
    and ``sys.path`` so they cannot shadow ``bayesdll_b200``'s own modules.
    """
    ref = find_reference()
    _install_matplotlib_stub()
    added = [ref, os.path.join(ref, "src")]
    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _REF_MODULE_NAMES}
    for k in saved:
        del sys.modules[k]
    sys.path[:0] = added
    try:
        yield ref
    finally:
        for p in added:
            with contextlib.suppress(ValueError):
                sys.path.remove(p)
        for k in list(sys.modules):
            if k.split(".")[0] in _REF_MODULE_NAMES:
                del sys.modules[k]
        sys.modules.update(saved)


def load(modname):
    """Import one reference module (e.g. 'methods.sghmc', 'calibration') and return it.

    The module object stays usable after the context exits."""
    with reference_imports():
        return importlib.import_module(modname)


class NoiseTape:
    """Replacement for ``torch.randn_like`` that hands out consecutive slices of a
    pre-generated flat fp32 noise vector ('Noise injection')."""

    def __init__(self, flat_noise):
        self.flat = torch.as_tensor(flat_noise, dtype=torch.float32)
        self.pos = 0
        self.calls = 0

    def __call__(self, like, **kw):
        n = like.numel()
        out = self.flat[self.pos:self.pos + n].reshape(like.shape).clone()
        assert out.numel() == n, "noise tape exhausted"
        self.pos += n
        self.calls += 1
        return out.to(like.device)


class SeededTape:
    """Like ``NoiseTape`` for runs whose noise would not fit a fixture (millions of parameters x many draws): values come
    from ``numpy.random.default_rng(seed).standard_normal(n, float32)`` call by call, so the golden generator and the GPU
    test regenerate the same stream from the seed alone (both make the same sequence of calls)."""

    def __init__(self, seed):
        import numpy as np
        self.rng = np.random.default_rng(seed)
        self.pos = 0
        self.calls = 0

    def __call__(self, like, **kw):
        import numpy as np
        n = like.numel()
        out = torch.from_numpy(self.rng.standard_normal(n, dtype=np.float32)).reshape(like.shape)
        self.pos += n
        self.calls += 1
        return out.to(like.device)


@contextlib.contextmanager
def injected_noise(flat_noise):
    tape = SeededTape(int(flat_noise)) if isinstance(flat_noise, (int,)) else NoiseTape(flat_noise)
    orig = torch.randn_like
    torch.randn_like = tape
    try:
        yield tape
    finally:
        torch.randn_like = orig


class GradInjectNet(nn.Module):
    """A 'network' whose parameters have reference-style names and whose forward
    returns sum_t (p_t * G_t).sum(), so that after ``loss.backward()`` with
    ``criterion = lambda out, y: out`` every ``p.grad == G_t`` exactly
    ('Gradient injection without touching reference code').

    ``shapes``: ordered dict name -> shape; names may contain dots
    (``layers.0.bias``, ``classifier.weight``) – they are registered through nested
    ModuleDicts so ``named_parameters()`` yields exactly those names.
    """

    def __init__(self, shapes, readout_name, init_std=0.1, seed=0):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        self._names = list(shapes)
        for name, shape in shapes.items():
            parts = name.split(".")
            mod = self
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    mod.add_module(part, nn.Module())
                mod = getattr(mod, part)
            mod.register_parameter(parts[-1], nn.Parameter(torch.randn(shape, generator=gen) * init_std))
        self.readout_name = readout_name
        self._G = None

    def set_grads(self, grads):
        self._G = [torch.as_tensor(g, dtype=torch.float32) for g in grads]

    def forward(self, x):
        tot = 0.0
        for p, g in zip(self.parameters(), self._G):
            tot = tot + (p * g).sum()
        return tot


def identity_criterion(out, y):
    return out


class _HandOutGrads(torch.autograd.Function):
    """forward: a scalar; backward: hands every parameter its prepared gradient tensor, no arithmetic at all."""

    @staticmethod
    def forward(ctx, holder, *params):
        ctx.holder = holder
        return params[0].new_zeros(())

    @staticmethod
    def backward(ctx, grad_out):
        return (None,) + tuple(ctx.holder.grads)


class TimedInjectNet(GradInjectNet):
    """GradInjectNet for TIMING the reference's update loop: the forward / backward pair costs
    nothing but autograd's bookkeeping (one clone of each prepared gradient by AccumulateGrad, 8 B/param), so the time of
    the reference's unmodified ``Model.forward`` + ``optimizer.step()`` is the time of its per-tensor update statements."""

    def set_grads(self, grads):
        self.grads = list(grads)

    def forward(self, x):
        return _HandOutGrads.apply(self, *self.parameters())
