"""ctypes front end of oracle/bdl_oracle.c (TEST INFRASTRUCTURE ONLY -- see module docstring of
oracle/sampler_oracle.py for who may import this).  Builds the shared object with `make` on first use."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libbdl_oracle.so")

_lib = None


def build(force=False):
    src = os.path.join(HERE, "bdl_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "bdl.h")
    stale = (not os.path.exists(SO)) or any(os.path.getmtime(f) > os.path.getmtime(SO) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
    return _lib


def set_threads(n=0):
    """Use ``n`` OpenMP threads (0: leave unchanged); returns the thread count the next call will use."""
    fn = lib().bdl_oracle_set_threads
    fn.restype = C.c_int
    return int(fn(C.c_int(int(n))))


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().bdl_oracle_philox4x32_10(c, k, out)
    return tuple(out)


def philox_normal(n, seed, stream_id, subseq):
    out = np.empty(n, np.float32)
    rc = lib().bdl_oracle_philox_normal(_fp(out), C.c_uint64(n), C.c_uint64(seed & (2**64 - 1)), C.c_uint32(stream_id),
                                        C.c_uint64(subseq & (2**64 - 1)))
    assert rc == 0
    return out


def step(variant, theta, g, theta0, v, m, s, buf, runs, scalars, noise):
    """In-place on the numpy arrays.  runs: ctypes array of bayesdll_b200._lib.Run with HOST g pointers;
    scalars / noise: the same ctypes structs the CUDA path receives (noise.xi_dev = host address or 0)."""
    n = theta.size
    rc = lib().bdl_oracle_step(C.c_int(variant), _fp(theta), _fp(g), _fp(theta0), _fp(v), _fp(m), _fp(s), _fp(buf),
                               C.c_uint64(n), runs, C.c_uint32(len(runs)), C.byref(scalars), C.byref(noise))
    assert rc == 0, rc


def step_clipped(variant, theta, g, theta0, v, m, s, buf, runs, scalars, noise, coef):
    """bdl_step_clipped on the host: the step with p.grad scaled by ``coef`` (args.clip_grad)."""
    rc = lib().bdl_oracle_step_clipped(C.c_int(variant), _fp(theta), _fp(g), _fp(theta0), _fp(v), _fp(m), _fp(s), _fp(buf),
                                       C.c_uint64(theta.size), runs, C.c_uint32(len(runs)), C.byref(scalars), C.byref(noise),
                                       C.c_float(coef))
    assert rc == 0, rc


def step_gradnorm(variant, theta, g, theta0, v, m, s, buf, runs, scalars, noise):
    """bdl_step_gradnorm on the host -> sum of squares (float) of what the reference holds in p.grad; ``runs``: per tensor."""
    out = C.c_double(0.0)
    rc = lib().bdl_oracle_step_gradnorm(C.c_int(variant), _fp(theta), _fp(g), _fp(theta0), _fp(v), _fp(m), _fp(s), _fp(buf),
                                        C.c_uint64(theta.size), runs, C.c_uint32(len(runs)), C.byref(scalars), C.byref(noise),
                                        C.byref(out))
    assert rc == 0, rc
    return out.value


def clip_coef(sumsq, max_norm):
    lib().bdl_oracle_clip_coef.restype = C.c_float
    tn = C.c_float(0)
    coef = lib().bdl_oracle_clip_coef(C.c_double(sumsq), C.c_float(max_norm), C.byref(tn))
    return coef, tn.value


def draw(mean, second, out, var_mode, scale, div_mode, noise, center=None):
    rc = lib().bdl_oracle_draw(_fp(mean), _fp(second), _fp(center), _fp(out), C.c_uint64(mean.size), C.c_int(var_mode),
                               C.c_float(scale), C.c_int(div_mode), C.byref(noise))
    assert rc == 0


def dropout_mix(m, theta0, out, p_drop, noise, runs=None, z_out=None):
    """``runs``: ctypes array of bdl_run (host) or None."""
    rc = lib().bdl_oracle_dropout_mix(_fp(m), _fp(theta0), _fp(out), _fp(z_out), C.c_uint64(m.size),
                                      runs if runs is not None else None, C.c_uint32(0 if runs is None else len(runs)),
                                      C.c_float(p_drop), C.byref(noise))
    assert rc == 0


def moments_avg(theta, mom1, mom2, cnt, init, div_mode):
    rc = lib().bdl_oracle_moments_avg(_fp(theta), _fp(mom1), _fp(mom2), C.c_uint64(theta.size), C.c_float(cnt),
                                      C.c_float(cnt + 1), C.c_int(init), C.c_int(div_mode))
    assert rc == 0


def moments_welford(theta, mean, M2, n, init, div_mode):
    rc = lib().bdl_oracle_moments_welford(_fp(theta), _fp(mean), _fp(M2), C.c_uint64(theta.size), C.c_float(n),
                                          C.c_int(init), C.c_int(div_mode))
    assert rc == 0
