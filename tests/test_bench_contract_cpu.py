"""CPU: the bench.py contract that does not need a GPU -- the reference arm prints ONE JSON line with the keys the driver
reads (same metric / unit / config as the GPU arm), and the GPU arm refuses to run without a CUDA device instead of
falling back to anything."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], cwd=ROOT, capture_output=True, text=True,
                          timeout=600, env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-seconds", "0.5", "--no-cfg1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "params/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("sampler step params/s") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "ViT-L/32" in d["config"]["workload"] and "configs[2]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # where a reference tree exists (this container, or baseline/_ref on the GPU box) the reference's own code is timed
    from baseline import reference_arm
    assert cb["kind"] == ("reference" if reference_arm.available() else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "params/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return                                            # on the GPU box the real bench runs (driver), nothing to check here
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
