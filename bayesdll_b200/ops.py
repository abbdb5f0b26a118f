"""Torch-tensor front end of the C ABI (include/bdl.h).  PyTorch is plumbing here: it owns the device
memory and the stream; every numerical operation is a hand-written sm_100a kernel in libbdl.so.

Every function validates device / dtype / contiguity on the Python side ("Error
conventions"), launches asynchronously on the caller's current CUDA stream and raises ``BdlError`` on
a non-zero status.  There is no CPU implementation: CPU tensors are rejected.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ADAM_SGHMC, CSGHMC, DIV_RECIP, SGHMC, SGLD, STREAM_STEP, STREAM_USER, BdlError, Noise, Scalars

__all__ = ["make_scalars", "upload_runs", "step", "step_gradnorm", "clip_coef", "step_clipped", "make_capture", "philox_normal", "moments_avg", "moments_welford",
           "capture_ring", "draw", "ensemble", "ce_err", "lse_accum", "lse_rescale", "lse_finalize", "calibrate", "bma_mean", "dropout_mix",
           "nll_temperature",
           "set_launch_config"]


def _ptr(t, name, dtype=torch.float32, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise BdlError(f"{name}: tensor required")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise BdlError(f"{name}: expected a CUDA tensor (bayesdll_b200 has no CPU path)")
    if t.dtype != dtype:
        raise BdlError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise BdlError(f"{name}: tensor must be contiguous")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its first tensor argument current (the C ABI launches on the calling thread's
    current device and stream; torch's own ops switch devices implicitly, a ctypes call does not)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def set_launch_config(ctas_per_sm=0, unroll=0, threads=0):
    _lib.check(_lib.load().bdl_set_launch_config(ctas_per_sm, unroll, threads), "bdl_set_launch_config")


# ----------------------------------------------------------------------------------------------
# host scalar preparation: Python doubles -> the fp32 values the reference's eager ops would use
# ----------------------------------------------------------------------------------------------
def make_scalars(variant, *, lr_body, lr_head, ND, Ninflate=1.0, prior_sig=1.0, nd=1.0, alpha=0.05, mu=0.0,
                 beta1=0.9, beta2=0.999, eps=1e-8, temperature=1.0, t=1, first_step=False, add_noise=True,
                 div_mode=DIV_RECIP):
    """All arithmetic below is host fp64 exactly as written in the reference; ctypes' c_float performs the
    double -> fp32 rounding torch applies when a Python scalar meets an fp32 tensor.

    noise scale  SGLD   nd*np.sqrt(2/(N*lr))        methods/sgld.py:478,483
                 SGHMC  nd*np.sqrt(2*a/(N*lr))      methods/sghmc.py:500
                 cSGHMC nd*np.sqrt(2*a*lr)/N        methods/csghmc.py:765
    """
    N = ND * Ninflate
    sc = Scalars()
    lrs = (float(lr_body), float(lr_head))
    for h, lr in enumerate(lrs):
        sc.lr[h] = lr
        if variant == SGLD:
            c = nd * np.sqrt(2 / (N * lr))
        elif variant == SGHMC:
            c = nd * np.sqrt(2 * alpha / (N * lr))
        elif variant == CSGHMC:
            c = nd * np.sqrt((2 * alpha * lr)) / N
        else:
            c = 0.0
        sc.noise_scale[h] = c
    sc.one_minus_alpha = 1 - alpha
    sc.sig2 = prior_sig if variant == CSGHMC else prior_sig ** 2
    sc.N = N
    sc.mu = mu if variant in (SGLD, ADAM_SGHMC) else 0.0
    sc.beta1, sc.one_minus_beta1 = beta1, 1 - beta1
    sc.beta2, sc.one_minus_beta2 = beta2, 1 - beta2
    sc.bias_corr1 = 1 - beta1 ** t
    sc.bias_corr2 = 1 - beta2 ** t
    # torch CUDA evaluates `tensor / python_scalar` as tensor * fp32(1.0 / s), the reciprocal taken in double
    # (BinaryDivTrueKernel.cu; probed, tools/probe_torch_div.py): hand the kernel exactly those factors
    if variant != CSGHMC:
        sc.inv_sig2 = 1.0 / (prior_sig ** 2)
    sc.inv_N = 1.0 / N
    sc.inv_bias_corr1 = 1.0 / (1 - beta1 ** t)
    sc.inv_bias_corr2 = 1.0 / (1 - beta2 ** t)
    sc.inv_temperature = 1.0 / temperature
    sc.eps = eps
    sc.two_alpha = 2 * alpha
    sc.nd = nd
    sc.temperature = temperature
    sc.first_step = int(bool(first_step))
    sc.add_noise = int(bool(add_noise))
    sc.div_mode = div_mode
    return sc


def upload_runs(run_array, device, out=None):
    """Copy a ctypes bdl_run array to the device (uint8 tensor).  Returns (tensor, nruns)."""
    nbytes = C.sizeof(run_array)
    host = torch.frombuffer(bytearray(bytes(run_array)), dtype=torch.uint8)
    if out is None or out.numel() < nbytes:
        out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out[:nbytes].copy_(host, non_blocking=False)
    out._bdl_host = run_array          # host copy rides along: lets bdl_step inline small tables into the kernel arguments
    return out, len(run_array)


def make_noise(xi=None, seed=0, subseq=0, stream_id=STREAM_STEP):
    nz = Noise()
    nz.xi_dev = 0 if xi is None else _ptr(xi, "xi")
    nz.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    nz.subseq = int(subseq) & 0xFFFFFFFFFFFFFFFF
    nz.stream_id = int(stream_id)
    return nz


# ----------------------------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------------------------
def make_capture(kind, first, second, cnt, init=False):
    """Capture spec for the fused step (bdl_capture).  kind 'avg': running moments, ``cnt`` = samples averaged so far
    (the argument of moments_avg); kind 'welford': ``cnt`` = n of moments_welford.  ``second`` may be None for 'avg'."""
    cap = _lib.Capture()
    cap.kind = {"avg": _lib.CAPTURE_AVG, "welford": _lib.CAPTURE_WELFORD}[kind]
    cap.init = int(bool(init))
    cap.first_dev = _ptr(first, "capture first")
    cap.second_dev = _ptr(second, "capture second", allow_none=(kind == "avg")) or 0
    cap.cnt = float(cnt)
    cap.cnt_plus_1 = float(cnt + 1)
    cap._keep = (first, second)
    return cap


@_on_tensor_device
def step(variant, theta, g, theta0, v, m, s, buf, runs_dev, nruns, scalars, noise, capture=None, runs_host=None):
    """One fused sampler update (bdl_step).  Tensors are padded-flat fp32 CUDA buffers; unused state may be
    None.  ``noise`` from make_noise(); ``runs_dev`` from upload_runs().  ``capture`` (make_capture) additionally folds
    the new theta into running moments in the same pass (bdl_step_capture).  ``runs_host``: address (or ctypes array) of a
    HOST copy of the run table -- with it a table of <= 512 rows travels inside the kernel arguments and ``runs_dev`` may
    be None (default: the host copy upload_runs() attached to ``runs_dev``)."""
    n = theta.numel()
    for name, t in (("g", g), ("theta0", theta0), ("v", v), ("m", m), ("s", s), ("buf", buf)):
        if t is not None and t.numel() != n:
            raise BdlError(f"{name}: length {t.numel()} != theta length {n}")
    args = (int(variant), _ptr(theta, "theta"), _ptr(g, "g", allow_none=True), _ptr(theta0, "theta0", allow_none=True),
            _ptr(v, "v", allow_none=True), _ptr(m, "m", allow_none=True), _ptr(s, "s", allow_none=True),
            _ptr(buf, "buf", allow_none=True), n, _ptr(runs_dev, "runs", torch.uint8, allow_none=runs_host is not None), nruns,
            runs_host if runs_host is not None else getattr(runs_dev, "_bdl_host", None), C.byref(scalars), C.byref(noise))
    if capture is None:
        _lib.check(_lib.load().bdl_step(*args, _stream()), "bdl_step")
        return
    for t in capture._keep:
        if t is not None and t.numel() != n:
            raise BdlError(f"capture buffer length {t.numel()} != theta length {n}")
    _lib.check(_lib.load().bdl_step_capture(*args, C.byref(capture), _stream()), "bdl_step_capture")


def _clip_args(variant, theta, g, theta0, v, m, s, buf, runs_host, nruns, scalars, noise):
    n = theta.numel()
    for name, t in (("g", g), ("theta0", theta0), ("v", v), ("m", m), ("s", s), ("buf", buf)):
        if t is not None and t.numel() != n:
            raise BdlError(f"{name}: length {t.numel()} != theta length {n}")
    if runs_host is None:
        raise BdlError("gradient-norm clipping needs the host copy of a per-tensor run table")
    return (int(variant), _ptr(theta, "theta"), _ptr(g, "g", allow_none=True), _ptr(theta0, "theta0", allow_none=True),
            _ptr(v, "v", allow_none=True), _ptr(m, "m", allow_none=True), _ptr(s, "s", allow_none=True),
            _ptr(buf, "buf", allow_none=True), n, runs_host, nruns, C.byref(scalars), C.byref(noise))


@_on_tensor_device
def step_gradnorm(variant, theta, g, theta0, v, m, s, buf, runs_host, nruns, scalars, noise, sumsq):
    """Pass 1 of the clipped update (bdl_step_gradnorm): ``sumsq`` (fp64[1], caller-zeroed) += sum of squares of what the
    reference holds in p.grad between Model.forward and optimizer.step().  Nothing else is written."""
    args = _clip_args(variant, theta, g, theta0, v, m, s, buf, runs_host, nruns, scalars, noise)
    _lib.check(_lib.load().bdl_step_gradnorm(*args, _ptr(sumsq, "sumsq", torch.float64), _stream()), "bdl_step_gradnorm")


@_on_tensor_device
def clip_coef(sumsq, max_norm, coef, total_norm=None):
    """coef[0] = min(1, max_norm / (sqrt(sumsq[0]) + 1e-6)) in fp32 (torch.nn.utils.clip_grad_norm_), on the device."""
    _lib.check(_lib.load().bdl_clip_coef(_ptr(sumsq, "sumsq", torch.float64), float(max_norm), _ptr(coef, "coef"),
                                         _ptr(total_norm, "total_norm", allow_none=True), _stream()), "bdl_clip_coef")


@_on_tensor_device
def step_clipped(variant, theta, g, theta0, v, m, s, buf, runs_host, nruns, scalars, noise, coef):
    """Pass 2 (bdl_step_clipped): the fused update with p.grad scaled by the device scalar ``coef``."""
    args = _clip_args(variant, theta, g, theta0, v, m, s, buf, runs_host, nruns, scalars, noise)
    _lib.check(_lib.load().bdl_step_clipped(*args, _ptr(coef, "coef"), _stream()), "bdl_step_clipped")


@_on_tensor_device
def philox_normal(out, seed, stream_id=STREAM_USER, subseq=0):
    rc = _lib.load().bdl_philox_normal(_ptr(out, "out"), out.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id),
                                       int(subseq) & 0xFFFFFFFFFFFFFFFF, _stream())
    _lib.check(rc, "bdl_philox_normal")
    return out


@_on_tensor_device
def moments_avg(theta, mom1, mom2, cnt, init=False, div_mode=DIV_RECIP):
    """init: mom1 = theta*1.0, mom2 = theta**2.  else mom <- (theta^k + cnt*mom)/(cnt+1)."""
    rc = _lib.load().bdl_moments_avg(_ptr(theta, "theta"), _ptr(mom1, "mom1"), _ptr(mom2, "mom2", allow_none=True),
                                     theta.numel(), float(cnt), float(cnt + 1), int(init), div_mode, _stream())
    _lib.check(rc, "bdl_moments_avg")


@_on_tensor_device
def moments_welford(theta, mean, M2, n, init=False, div_mode=DIV_RECIP):
    rc = _lib.load().bdl_moments_welford(_ptr(theta, "theta"), _ptr(mean, "mean"), _ptr(M2, "M2"), theta.numel(),
                                         float(n), int(init), div_mode, _stream())
    _lib.check(rc, "bdl_moments_welford")


def set_ring_config(chunks_per_cta=4):
    _lib.check(_lib.load().bdl_set_ring_config(int(chunks_per_cta)), "bdl_set_ring_config")


@_on_tensor_device
def capture_ring(theta, ring, slot):
    n = theta.numel()
    if ring.dim() != 2 or ring.shape[1] != n or not (0 <= slot < ring.shape[0]):
        raise BdlError("ring must be [slots, n] and slot in range")
    rc = _lib.load().bdl_capture_ring(_ptr(theta, "theta"), _ptr(ring, "ring"), int(slot), n, _stream())
    _lib.check(rc, "bdl_capture_ring")


VAR_FROM_MOMENTS, VAR_FROM_WELFORD, VAR_TINY, VAR_GIVEN, STD_GIVEN = 0, 1, 2, 3, 4


@_on_tensor_device
def draw(mean, second, out, var_mode, scale, noise, div_mode=DIV_RECIP, center=None):
    """out = (center or mean) + sqrt(var(mean, second)) * eps   (see bdl_draw)."""
    rc = _lib.load().bdl_draw(_ptr(mean, "mean"), _ptr(second, "second", allow_none=True),
                              _ptr(center, "center", allow_none=True), _ptr(out, "out"),
                              mean.numel(), int(var_mode), float(scale), div_mode, C.byref(noise), _stream())
    _lib.check(rc, "bdl_draw")


@_on_tensor_device
def dropout_mix(m, theta0, out, p_drop, noise, runs_dev=None, nruns=0, z_out=None):
    """MC-Dropout draw: out = z*m + (1-z)*theta0, z = (u > p_drop) (see bdl_dropout_mix).  ``noise``: make_noise(seed=...)
    for in-kernel Philox uniforms or make_noise(xi=u) for injected uniforms; ``runs_dev``: optional device run table
    whose CLS_NODROP rows keep z = 1; ``z_out``: optional fp32 mask output."""
    n = m.numel()
    for name, t in (("theta0", theta0), ("out", out), ("z_out", z_out)):
        if t is not None and t.numel() != n:
            raise BdlError(f"{name}: length {t.numel()} != m length {n}")
    rc = _lib.load().bdl_dropout_mix(_ptr(m, "m"), _ptr(theta0, "theta0"), _ptr(out, "out"), _ptr(z_out, "z_out", allow_none=True),
                                     n, None if runs_dev is None else _ptr(runs_dev, "runs", torch.uint8), int(nruns),
                                     float(p_drop), C.byref(noise), _stream())
    _lib.check(rc, "bdl_dropout_mix")
    return out


@_on_tensor_device
def ensemble(logits_all, out, nst, weight=1.0, mode=0):
    """logits_all [B,K,S] -> out [B,K] (see bdl_ensemble).  nst == 0 -> no '- log S'."""
    B, K, S = logits_all.shape
    log_S = float(np.float32(np.log(nst))) if nst > 0 else 0.0
    rc = _lib.load().bdl_ensemble(_ptr(logits_all, "logits_all"), B, K, S, log_S, float(weight), int(mode),
                                  _ptr(out, "out"), _stream())
    _lib.check(rc, "bdl_ensemble")
    return out


@_on_tensor_device
def ce_err(logits, y, loss_sum, err_count):
    B, K = logits.shape
    rc = _lib.load().bdl_ce_err(_ptr(logits, "logits"), _ptr(y, "y", torch.int64), B, K,
                                _ptr(loss_sum, "loss_sum", torch.float64), _ptr(err_count, "err_count", torch.int32),
                                _stream())
    _lib.check(rc, "bdl_ce_err")


@_on_tensor_device
def lse_accum(logits, m, s):
    """Running logsumexp over samples of log_softmax(logits): (m, s) <- combine((m, s), log_softmax(logits))."""
    B, K = logits.shape
    rc = _lib.load().bdl_lse_accum(_ptr(logits, "logits"), B, K, _ptr(m, "m"), _ptr(s, "s"), _stream())
    _lib.check(rc, "bdl_lse_accum")


@_on_tensor_device
def lse_rescale(m_local, m_global, s):
    rc = _lib.load().bdl_lse_rescale(_ptr(m_local, "m_local"), _ptr(m_global, "m_global"), _ptr(s, "s"), s.numel(), _stream())
    _lib.check(rc, "bdl_lse_rescale")


@_on_tensor_device
def lse_finalize(m, s, out, n_samples, weight=1.0, mode=0):
    B, K = out.shape
    log_S = float(np.float32(np.log(n_samples))) if n_samples > 0 else 0.0
    rc = _lib.load().bdl_lse_finalize(_ptr(m, "m"), _ptr(s, "s"), B, K, log_S, float(weight), int(mode), _ptr(out, "out"),
                                      _stream())
    _lib.check(rc, "bdl_lse_finalize")
    return out


@_on_tensor_device
def calibrate(logits, labels, edges, temperature=1.0, use_f64=False, want_binned=False):
    """Returns device tensors (bin_size[M], acc_sum[M], conf_sum[M], nll_sum[1], near_edge[1], binned|None)."""
    N, K = logits.shape
    M = edges.numel()
    dev = logits.device
    stats = torch.zeros(3 * M + 1, dtype=torch.float64, device=dev)
    near = torch.zeros(1, dtype=torch.int64, device=dev)
    binned = torch.empty(N * K, dtype=torch.int32, device=dev) if want_binned else None
    base = stats.data_ptr()
    rc = _lib.load().bdl_calibrate(
        _ptr(logits, "logits"), _ptr(labels, "labels", torch.int64), N, K, float(temperature), int(use_f64),
        _ptr(edges, "edges", torch.float64), M, base, base + 8 * M, base + 16 * M, base + 24 * M,
        near.data_ptr(), None if binned is None else binned.data_ptr(), _stream())
    _lib.check(rc, "bdl_calibrate")
    return stats[:M], stats[M:2 * M], stats[2 * M:3 * M], stats[3 * M:], near, binned


@_on_tensor_device
def bma_mean(logits_all, out):
    """logits_all [B,K,S] -> out [B,K]: fp32 running sum over S in order, divided by fp32(S) (see bdl_bma_mean)."""
    B, K, S = logits_all.shape
    rc = _lib.load().bdl_bma_mean(_ptr(logits_all, "logits_all"), B, K, S, _ptr(out, "out"), _stream())
    _lib.check(rc, "bdl_bma_mean")
    return out


@_on_tensor_device
def nll_temperature(logits, labels, temperature, row_nll, out_mean):
    """out_mean[0] = mean_i(logsumexp(logits[i]/T) - logits[i,y_i]/T) in fp64 (see bdl_nll_temperature)."""
    N, K = logits.shape
    rc = _lib.load().bdl_nll_temperature(_ptr(logits, "logits"), _ptr(labels, "labels", torch.int64), N, K,
                                         float(temperature), _ptr(row_nll, "row_nll", torch.float64),
                                         _ptr(out_mean, "out_mean", torch.float64), _stream())
    _lib.check(rc, "bdl_nll_temperature")
    return out_mean


def selftest_math(device):
    """Exhaustive device self-test of the branch-free sqrt / reciprocal / quotient helpers (bdl_selftest_math):
    -> dict(mismatch=[sqrt, rcp, div], fast=[...]) over all 2^32 fp32 bit patterns."""
    out = torch.zeros(6, dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().bdl_selftest_math(out.data_ptr(), _stream()), "bdl_selftest_math")
    h = out.cpu().tolist()
    return {"mismatch": h[:3], "fast": h[3:]}


@_on_tensor_device
def probe_stream(a, b, c, d, reads, writes, threads=64):
    """Bare-traffic yardstick (bdl_probe_stream): (reads, writes) = (4, 2) / (2, 1) / (1, 1) streams over flat fp32 buffers,
    next to no arithmetic, the product kernels' launch shape.  Overwrites ``a`` (and ``b`` for two writes)."""
    rc = _lib.load().bdl_probe_stream(_ptr(a, "a"), _ptr(b, "b") if b is not None else None, _ptr(c, "c"),
                                      _ptr(d, "d") if d is not None else None, a.numel(), int(reads), int(writes), int(threads),
                                      _stream())
    _lib.check(rc, "bdl_probe_stream")


class HostChain:
    """Python handle of the host-buffer chain API (bdl_chain_*): sampler state resident in HBM, gradient in / theta
    out through pinned host tensors.  See include/bdl.h."""

    def __init__(self, n, variant, with_sgd_momentum=False, chunk_elems=0):
        self._h = C.c_void_p()
        self.n, self.variant = int(n), int(variant)
        _lib.check(_lib.load().bdl_chain_create(self.n, self.variant, int(with_sgd_momentum), int(chunk_elems),
                                                C.byref(self._h)), "bdl_chain_create")

    def close(self):
        if self._h:
            _lib.load().bdl_chain_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @staticmethod
    def _host(t, n):
        if not (isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous()
                and t.numel() == n):
            raise BdlError("expected a contiguous fp32 CPU tensor of the chain's length")
        return t.data_ptr()

    def upload(self, which, host):
        _lib.check(_lib.load().bdl_chain_upload(self._h, int(which), self._host(host, self.n)), "bdl_chain_upload")

    def download(self, which, host):
        _lib.check(_lib.load().bdl_chain_download(self._h, int(which), self._host(host, self.n)), "bdl_chain_download")
        return host

    def step_host(self, g_host, theta_out_host, run_array, scalars, noise):
        for t in (g_host, theta_out_host):
            if not t.is_pinned():
                raise BdlError("bdl_chain_step_host needs pinned host tensors (tensor.pin_memory())")
        _lib.check(_lib.load().bdl_chain_step_host(self._h, self._host(g_host, self.n), self._host(theta_out_host, self.n),
                                                   run_array, len(run_array), C.byref(scalars), C.byref(noise)),
                   "bdl_chain_step_host")
