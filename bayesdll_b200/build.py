"""Build libbdl.so (hand-written sm_100a kernels behind the C ABI in include/bdl.h) in-tree with nvcc.

    python -m bayesdll_b200.build        # or:  python bayesdll_b200/build.py

The shared object is written next to this file (bayesdll_b200/libbdl.so): it is git-ignored but travels
to the GPU box with the repo snapshot.  nvcc cross-compiles for sm_100a without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbdl.so")
STEP_INSTANCES = ["sgld", "sgld_buf", "sghmc", "csghmc", "adam_sghmc", "adam_sghmc_buf", "adam_csghmc"]
# the slowest units first: they are compiled in parallel and the Adam kernels take longest
SOURCES = [f"bdl_step_inst_{v}.cu" for v in reversed(STEP_INSTANCES)] + ["bdl_api.cu", "bdl_step.cu", "bdl_capture.cu", "bdl_draw.cu", "bdl_predict.cu", "bdl_host.cu", "bdl_selftest.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",                     # no implicit contraction: fused ops are written explicitly (__fmaf_rn)
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
] + os.environ.get("BDL_EXTRA_NVCC_FLAGS", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_hash():
    """sha256 over every kernel source, the ABI header and the nvcc flags: identifies the build an ncu capture came from
    (profiles/traffic.json)."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(ROOT, "include", "bdl.h")]:
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode() + b"\0" + f.read())
    h.update(" ".join(f for f in NVCC_FLAGS if not f.startswith(ROOT)).encode())      # the include path is checkout-specific
    return h.hexdigest()


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "bdl.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, only=None):
    """``only``: recompile just these sources and relink with the objects already in build/ (A/B builds)."""
    if not force and not only and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if only and src not in only and os.path.exists(obj):
            continue
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("BDL_NVCC_EXTRA", "").split(), "-c", path, "-o", obj]   # e.g. -DBDL_LD_MODE=1 for A/B builds
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose and out:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("libbdl build failed")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT, *objs, "-lcudart"])
    return OUT


if __name__ == "__main__":
    _only = [a for a in sys.argv[1:] if a.endswith(".cu")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, only=_only or None))
