"""The drop-in boundary from plain C (examples/c_abi_demo.c): compiled with gcc against include/bdl.h and linked with
libbdl.so -- no Python, no torch in the loop.  CPU: compile + link.  GPU: run, bit-exact vs its own host restatement."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA headers are not available")
    from bayesdll_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", os.path.join(ROOT, "examples", "c_abi_demo.c"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(CUDA, "include"), "-L", libdir, "-lbdl", "-L", os.path.join(CUDA, "lib64"),
                           "-lcudart", "-lm", "-ffp-contract=off", f"-Wl,-rpath,{libdir}", "-Wall", "-Werror", "-o", exe])
    return exe


def test_c_demo_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_demo_runs_bit_exact(cuda_device, tmp_path):
    exe = _build(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(CUDA, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 of 4160 elements differ" in r.stdout and r.stdout.strip().endswith("OK")
