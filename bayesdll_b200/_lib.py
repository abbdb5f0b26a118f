"""ctypes binding of libbdl.so (C ABI declared in include/bdl.h).

The library is the product: there is no CPU / eager fallback.  If the shared object is missing or
a symbol cannot be resolved this module raises immediately (``python -m bayesdll_b200.build`` or
``__graft_entry__.build()`` produces it in-tree).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BDL_LIB_PATH") or os.path.join(HERE, "libbdl.so")   # BDL_LIB_PATH: A/B builds (tools/ab_builds.py)

# ---- enums / constants (include/bdl.h) -------------------------------------------------------
BDL_ABI_VERSION = 7
SGLD, SGHMC, CSGHMC, ADAM_SGHMC, ADAM_CSGHMC = range(5)
VARIANT_NAMES = {SGLD: "sgld", SGHMC: "sghmc", CSGHMC: "csghmc", ADAM_SGHMC: "adam_sghmc",
                 ADAM_CSGHMC: "adam_csghmc"}
CLS_HEAD, CLS_PRIOR, CLS_SKIP, CLS_NODROP = 1, 2, 4, 8
DIV_IEEE, DIV_RECIP = 0, 1
STREAM_STEP, STREAM_DRAW, STREAM_USER = 0, 1, 2
BUF_THETA, BUF_THETA0, BUF_V, BUF_M, BUF_S, BUF_SGD, BUF_GRAD = range(7)
MAX_RUNS = 2048


class Run(C.Structure):
    _fields_ = [("begin", C.c_uint64), ("end", C.c_uint64), ("valid_end", C.c_uint64),
                ("g_dev", C.c_uint64), ("cls", C.c_uint32), ("reserved", C.c_uint32)]


class Scalars(C.Structure):
    _fields_ = [("lr", C.c_float * 2), ("noise_scale", C.c_float * 2),
                ("one_minus_alpha", C.c_float), ("sig2", C.c_float), ("N", C.c_float), ("mu", C.c_float),
                ("beta1", C.c_float), ("one_minus_beta1", C.c_float), ("beta2", C.c_float),
                ("one_minus_beta2", C.c_float), ("bias_corr1", C.c_float), ("bias_corr2", C.c_float),
                ("eps", C.c_float), ("two_alpha", C.c_float), ("nd", C.c_float), ("temperature", C.c_float),
                ("first_step", C.c_int32), ("add_noise", C.c_int32), ("div_mode", C.c_int32),
                ("reserved", C.c_int32),
                ("inv_sig2", C.c_float), ("inv_N", C.c_float), ("inv_bias_corr1", C.c_float),
                ("inv_bias_corr2", C.c_float), ("inv_temperature", C.c_float), ("reserved2", C.c_int32)]


class Noise(C.Structure):
    _fields_ = [("xi_dev", C.c_uint64), ("seed", C.c_uint64), ("subseq", C.c_uint64),
                ("stream_id", C.c_uint32), ("reserved", C.c_uint32)]


class Capture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("init", C.c_int32), ("first_dev", C.c_uint64), ("second_dev", C.c_uint64),
                ("cnt", C.c_float), ("cnt_plus_1", C.c_float)]


CAPTURE_NONE, CAPTURE_AVG, CAPTURE_WELFORD = 0, 1, 2

assert C.sizeof(Capture) == 32
assert C.sizeof(Run) == 40 and C.sizeof(Noise) == 32 and C.sizeof(Scalars) == 112

_P = C.c_void_p
_U64, _U32, _I32, _F, _D = C.c_uint64, C.c_uint32, C.c_int, C.c_float, C.c_double

# name -> argtypes; every function returns int except the two noted below.  This table is also what
# tests/test_abi.py checks against include/bdl.h.
SIGNATURES = {
    "bdl_set_launch_config": [_I32, _I32, _I32],
    "bdl_step": [_I32, _P, _P, _P, _P, _P, _P, _P, _U64, _P, _U32, _P, C.POINTER(Scalars), C.POINTER(Noise), _P],
    "bdl_step_capture": [_I32, _P, _P, _P, _P, _P, _P, _P, _U64, _P, _U32, _P, C.POINTER(Scalars), C.POINTER(Noise),
                         C.POINTER(Capture), _P],
    "bdl_step_gradnorm": [_I32, _P, _P, _P, _P, _P, _P, _P, _U64, _P, _U32, C.POINTER(Scalars), C.POINTER(Noise), _P, _P],
    "bdl_clip_coef": [_P, _F, _P, _P, _P],
    "bdl_step_clipped": [_I32, _P, _P, _P, _P, _P, _P, _P, _U64, _P, _U32, C.POINTER(Scalars), C.POINTER(Noise), _P, _P],
    "bdl_philox_normal": [_P, _U64, _U64, _U32, _U64, _P],
    "bdl_moments_avg": [_P, _P, _P, _U64, _F, _F, _I32, _I32, _P],
    "bdl_moments_welford": [_P, _P, _P, _U64, _F, _I32, _I32, _P],
    "bdl_capture_ring": [_P, _P, _U64, _U64, _P],
    "bdl_set_ring_config": [_I32],
    "bdl_draw": [_P, _P, _P, _P, _U64, _I32, _F, _I32, C.POINTER(Noise), _P],
    "bdl_dropout_mix": [_P, _P, _P, _P, _U64, _P, _U32, _F, C.POINTER(Noise), _P],
    "bdl_ensemble": [_P, _U32, _U32, _U32, _F, _F, _I32, _P, _P],
    "bdl_ce_err": [_P, _P, _U32, _U32, _P, _P, _P],
    "bdl_lse_accum": [_P, _U32, _U32, _P, _P, _P],
    "bdl_lse_rescale": [_P, _P, _P, _U64, _P],
    "bdl_lse_finalize": [_P, _P, _U32, _U32, _F, _F, _I32, _P, _P],
    "bdl_calibrate": [_P, _P, _U64, _U32, _D, _I32, _P, _U32, _P, _P, _P, _P, _P, _P, _P],
    "bdl_bma_mean": [_P, _U32, _U32, _U32, _P, _P],
    "bdl_nll_temperature": [_P, _P, _U64, _U32, _D, _P, _P, _P],
    "bdl_selftest_math": [_P, _P],
    "bdl_probe_stream": [_P, _P, _P, _P, _U64, _I32, _I32, _I32, _P],
    "bdl_chain_create": [_U64, _I32, _I32, _U64, C.POINTER(_P)],
    "bdl_chain_destroy": [_P],
    "bdl_chain_upload": [_P, _I32, _P],
    "bdl_chain_download": [_P, _I32, _P],
    "bdl_chain_device_ptr": [_P, _I32, C.POINTER(_P)],
    "bdl_chain_step_host": [_P, _P, _P, _P, _U32, C.POINTER(Scalars), C.POINTER(Noise)],
}

_lib = None


class BdlError(RuntimeError):
    pass


def load():
    """Return the loaded CDLL; raise loudly if it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BdlError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run `python -m bayesdll_b200.build` "
            "(needs nvcc; cross-compiles for sm_100a). bayesdll_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.bdl_abi_version.restype = C.c_int
    lib.bdl_abi_version.argtypes = []
    lib.bdl_last_error.restype = C.c_char_p
    lib.bdl_last_error.argtypes = []
    ver = lib.bdl_abi_version()
    if ver != BDL_ABI_VERSION:
        raise BdlError(f"libbdl ABI version {ver} != expected {BDL_ABI_VERSION}; rebuild the library")
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing -> loud failure
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().bdl_last_error().decode("utf-8", "replace")
        raise BdlError(f"{what} failed with status {rc}: {msg}")
