"""Install the UNMODIFIED reference (omarezz46/BayesDLL) into ``baseline/_ref/`` for the reference arm of bench.py.

    python baseline/install_ref.py            # in the build container, where /root/reference is mounted

``baseline/_ref/`` is git-ignored (never part of the history) but NOT gpurun-ignored, so it travels to the GPU box with
the snapshot -- the only way the reference's own Python can run there.  Two steps:

1. the one sanctioned offline install: ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref``
   of a scratch copy of the checkout (the source tree is read-only).  The package it installs (``setup.py``:
   ``packages=['bayesdll']``, ``package_dir={'': 'src'}``) holds only the ORIGINAL five methods and ``calibration``;
2. the SG-MCMC family the north star names is not packaged at all -- it lives in top-level modules of the checkout
   (``methods/*.py``, ``calibration.py``, ``networks/``, ``utils.py``) that ``demo_vision.py`` imports by path.  They are
   staged next to the package byte for byte (sha256 recorded in ``INSTALL.json``), so ``oracle/refshim.find_reference``
   sees the same layout as the checkout: ``<root>/methods/sghmc.py`` plus an importable ``bayesdll``.

Nothing is edited, nothing is copied into tracked files.  Only ``bench.py --impl reference`` / the ``cpu_baseline`` and
``reference_eager_gpu`` legs (through oracle/refshim.py) ever import from here.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
STAGED = ("methods", "networks", "calibration.py", "utils.py")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def installed():
    return os.path.isfile(os.path.join(DST, "methods", "sghmc.py")) and os.path.isfile(os.path.join(DST, "INSTALL.json"))


def install(src="/root/reference", force=False):
    if not os.path.isfile(os.path.join(src, "methods", "sghmc.py")):
        return None
    if installed() and not force:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    pip_note = "ok"
    with tempfile.TemporaryDirectory(prefix="bdl_ref_") as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "figures"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:                         # the package is pure Python: fall back to placing src/bayesdll by hand
            pip_note = f"pip failed ({r.stderr.strip().splitlines()[-1] if r.stderr.strip() else r.returncode}); src/bayesdll staged"
            shutil.copytree(os.path.join(src, "src", "bayesdll"), os.path.join(DST, "bayesdll"))
    files = {}
    for item in STAGED:
        s, d = os.path.join(src, item), os.path.join(DST, item)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(s, d)
    for root, _, names in os.walk(DST):
        for nm in names:
            if nm.endswith(".py"):
                full = os.path.join(root, nm)
                files[os.path.relpath(full, DST)] = _sha(full)
    # every staged file is byte-identical to the checkout
    for rel, digest in files.items():
        for cand in (os.path.join(src, rel), os.path.join(src, "src", rel)):
            if os.path.isfile(cand):
                assert _sha(cand) == digest, f"{rel} differs from the checkout"
                break
    with open(os.path.join(DST, "INSTALL.json"), "w") as f:
        json.dump({"source": src, "pip": pip_note, "files": files}, f, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    out = install(force="--force" in sys.argv)
    print(out or "no reference checkout at /root/reference: nothing installed")
