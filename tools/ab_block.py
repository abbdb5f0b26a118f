"""A/B of the CTA size per update rule with back-to-back launches (the way bench.py times the step), interleaved and
repeated so that clock drift under the power cap hits every candidate alike.

    python tools/ab_block.py [--steps 200] [--rounds 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import _lib, ops, shapes  # noqa: E402
from bayesdll_b200.flat import FlatLayout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--only", default=None, help="regex on the case name")
    ap.add_argument("--threads", default="0,64,128,256", help="CTA sizes to sweep (0 = library default)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    named, readout = shapes.named_shapes("vit_l_32")
    lay = FlatLayout(named, readout)
    n = lay.n_padded
    gen = torch.Generator(device=dev).manual_seed(0)
    buf = {k: torch.randn(n, device=dev, generator=gen) * sc for k, sc in
           dict(theta=0.02, g=0.01, theta0=0.02, v=0.001, m=0.001, b=0.001).items()}
    buf["s"] = torch.rand(n, device=dev, generator=gen) * 1e-4 + 1e-8
    runs = {bm: ops.upload_runs(lay.run_table(bm), dev) for bm in ("informative", "uninformative")}
    cases = [("SGHMC", _lib.SGHMC, 24, 0.0, "informative"), ("SGLD mu=0.5", _lib.SGLD, 24, 0.5, "informative"),
             ("SGLD mu=0", _lib.SGLD, 16, 0.0, "informative"), ("cSGHMC", _lib.CSGHMC, 20, 0.0, "informative"),
             ("Adam-cSGHMC", _lib.ADAM_CSGHMC, 40, 0.0, "informative"),
             ("SGHMC uninformative (295 runs)", _lib.SGHMC, 24, 0.0, "uninformative"),
             ("SGHMC per-tensor grad pointers (296 runs)", _lib.SGHMC, 24, 0.0, "pointers"),
             ("Adam-cSGHMC per-tensor grad pointers", _lib.ADAM_CSGHMC, 40, 0.0, "pointers"),
             ("SGHMC pointers into the flat buffer (placement control)", _lib.SGHMC, 24, 0.0, "flatptr"),
             ("SGHMC pointers into one packed buffer (512 B pitch)", _lib.SGHMC, 24, 0.0, "packedptr")]
    # the training-loop launch: one run per tensor, each row carrying the address of that tensor's own gradient
    grads = [torch.randn(sg.numel, device=dev, generator=gen) * 1e-2 for sg in lay.segments]
    runs["pointers"] = ops.upload_runs(lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in grads]), dev)
    # controls: the same dependent gradient load, but the gradients sit (a) at their flat-layout offsets, (b) packed
    # back to back in one allocation at the caching allocator's 512 B pitch
    flat_views = lay.flat_views(buf["g"])
    runs["flatptr"] = ops.upload_runs(lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in flat_views]), dev)
    pitch = lambda k: (k * 4 + 511) // 512 * 512
    packed = torch.randn(sum(pitch(sg.numel) for sg in lay.segments) // 4, device=dev, generator=gen) * 1e-2
    offs, o = [], 0
    for sg in lay.segments:
        offs.append(packed.data_ptr() + o)
        o += pitch(sg.numel)
    runs["packedptr"] = ops.upload_runs(lay.run_table("informative", grad_ptrs=offs), dev)
    step_no = [0]
    import re
    for name, variant, bpp, mu, bias in cases:
        if a.only and not re.search(a.only, name):
            continue
        adam = variant == _lib.ADAM_CSGHMC
        sc = ops.make_scalars(variant, lr_body=1e-4, lr_head=1e-2, ND=3680, Ninflate=1e3, prior_sig=1.0, nd=1.0, alpha=0.18,
                              mu=mu, t=10)
        rd, nr = runs[bias]

        def fn():
            step_no[0] += 1
            ops.step(variant, buf["theta"], None if bias.endswith("ptr") or bias == "pointers" else buf["g"], None if variant == _lib.CSGHMC else buf["theta0"],
                     None if variant == _lib.SGLD else buf["v"], buf["m"] if adam else None, buf["s"] if adam else None,
                     buf["b"] if mu else None, rd, nr, sc, ops.make_noise(seed=42, subseq=step_no[0]))
        res = {}
        try:
            fn()
        except _lib.BdlError as e:                            # e.g. a slim A/B build without this variant
            print(f"{name:32s} skipped: {e}", flush=True)
            continue
        for _ in range(a.rounds):
            for T in [int(t) for t in a.threads.split(",")]:                      # 0 = library default (the lean kFast build where it applies)
                ops.set_launch_config(0, 0 if T == 0 else 1, T)
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res.setdefault(T, []).append(e0.elapsed_time(e1) / a.steps)
        ops.set_launch_config(0, 0, 0)
        line = "  ".join(f"T={T}: " + "/".join(f"{x:.4f}" for x in v) + f" (best {min(v):.4f} ms = "
                         f"{bpp * lay.n_dense / (min(v) * 1e-3) / 1e9:.0f} GB/s)" for T, v in res.items())
        print(f"{name:32s} {line}", flush=True)


if __name__ == "__main__":
    main()
