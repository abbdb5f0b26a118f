"""CPU: the C-ABI shared library loads and exports every symbol include/bdl.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "bdl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|const char\*)\s+(bdl_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    fns = _declared_functions()
    for must in ("bdl_step", "bdl_moments_avg", "bdl_moments_welford", "bdl_capture_ring", "bdl_draw", "bdl_ensemble",
                 "bdl_calibrate", "bdl_abi_version", "bdl_last_error", "bdl_philox_normal"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from bayesdll_b200 import _lib
    from bayesdll_b200 import build
    build.build()                                  # nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/bdl.h but not exported by libbdl.so"
    lib.bdl_abi_version.restype = ctypes.c_int
    assert lib.bdl_abi_version() == _lib.BDL_ABI_VERSION
    # python binding table covers the header (minus the two non-int-returning calls bound by hand)
    assert set(_lib.SIGNATURES) | {"bdl_abi_version", "bdl_last_error"} == set(_declared_functions())


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct of include/bdl.h, as the C compiler sees them, against the ctypes mirrors."""
    import shutil
    import subprocess
    from bayesdll_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    mirrors = {"bdl_run": _lib.Run, "bdl_scalars": _lib.Scalars, "bdl_noise": _lib.Noise, "bdl_capture": _lib.Capture}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bdl.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, field, val = ln.split()
        got[(cname, field)] = int(val)
    for cname, cls in mirrors.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    assert ctypes.sizeof(_lib.Scalars) == 112 and ctypes.sizeof(_lib.Run) == 40


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from bayesdll_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.BdlError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from bayesdll_b200 import _lib, ops
    with pytest.raises(_lib.BdlError, match="CUDA tensor"):
        ops.philox_normal(torch.zeros(8), seed=1)


def test_library_depends_only_on_the_cuda_runtime():
    """The drop-in boundary is a plain C ABI: libbdl.so must not link libtorch, libpython or anything beyond the CUDA
    runtime and the C/C++ runtimes."""
    import shutil
    import subprocess
    from bayesdll_b200 import _lib
    if shutil.which("readelf") is None:
        pytest.skip("readelf (binutils) not available")
    out = subprocess.check_output(["readelf", "-d", _lib.LIB_PATH], text=True)
    needed = [ln.split("[")[1].rstrip("]") for ln in out.splitlines() if "NEEDED" in ln]
    assert needed, out
    allowed = ("libcudart", "libstdc++", "libm.", "libgcc_s", "libc.", "libdl", "libpthread", "librt", "ld-linux")
    assert all(n.startswith(allowed) for n in needed), needed
    assert any(n.startswith("libcudart") for n in needed)
