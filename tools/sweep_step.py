"""Launch-shape sweep and per-kernel roofline table on a B200 (run under gpurun).

    python tools/sweep_step.py [--backbone vit_l_32] [--iters 20] [--full]
Prints one line per configuration: time, params/s, algorithmic GB/s, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import _lib, ops, shapes  # noqa: E402
from bayesdll_b200.flat import FlatLayout  # noqa: E402


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="vit_l_32")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default=None, help="regex on the row tag: run only matching rows (for ncu captures)")
    ap.add_argument("--warm", type=int, default=5)
    a = ap.parse_args()
    import re
    want = (lambda tag: re.search(a.only, tag) is not None) if a.only else (lambda tag: True)
    global timeit
    _timeit = timeit
    timeit = lambda fn, iters: _timeit(fn, iters, warm=a.warm)
    dev = torch.device("cuda:0")
    named, readout = shapes.named_shapes(a.backbone)
    lay = FlatLayout(named, readout)
    n = lay.n_padded
    pk, pk_kind = peak()
    print(f"# {a.backbone}: n_dense={lay.n_dense} n_padded={n} tensors={len(lay.segments)} peak={pk} GB/s ({pk_kind})")
    gen = torch.Generator(device=dev).manual_seed(0)
    buf = {k: torch.randn(n, device=dev, generator=gen) * sc for k, sc in
           dict(theta=0.02, g=0.01, theta0=0.02, v=0.001, m=0.001, b=0.001, xi=1.0).items()}
    buf["s"] = torch.rand(n, device=dev, generator=gen) * 1e-4 + 1e-8
    runs = {bm: ops.upload_runs(lay.run_table(bm), dev) for bm in ("informative", "uninformative")}
    step_no = [0]

    def report(tag, bytes_per, med, mn):
        gbs = bytes_per * lay.n_dense / (med * 1e-3) / 1e9
        print(f"{tag:58s} med {med:7.3f} ms  min {mn:7.3f} ms  {lay.n_dense / (med * 1e-3) / 1e9:7.1f} Gparam/s  "
              f"{gbs:7.1f} GB/s  {gbs / pk:5.3f} of {pk_kind} peak", flush=True)

    def stepper(variant, bytes_per, *, philox=True, mu=0.0, bias="informative", div=_lib.DIV_RECIP):
        adam = variant in (_lib.ADAM_SGHMC, _lib.ADAM_CSGHMC)
        sc = ops.make_scalars(variant, lr_body=1e-4, lr_head=1e-2, ND=3680, Ninflate=1e3, prior_sig=1.0, nd=1.0, alpha=0.18,
                              mu=mu, t=10, div_mode=div)
        rd, nr = runs[bias]

        def fn():
            step_no[0] += 1
            nz = ops.make_noise(seed=42, subseq=step_no[0]) if philox else ops.make_noise(xi=buf["xi"])
            ops.step(variant, buf["theta"], buf["g"], None if variant == _lib.CSGHMC else buf["theta0"],
                     None if variant == _lib.SGLD else buf["v"], buf["m"] if adam else None, buf["s"] if adam else None,
                     buf["b"] if mu else None, rd, nr, sc, nz)
        return fn

    # 1. launch-shape sweep on the headline kernel (ctas_per_sm = 0: one tile per CTA, grid = #tiles)
    for threads in (64, 128, 256):
        for unroll in (1, 2):
            for per_sm in (0, 16):
                tag = f"SGHMC philox recip  T={threads} U={unroll} CTAs/SM={per_sm if per_sm else 'all'}"
                if not want(tag):
                    continue
                ops.set_launch_config(per_sm, unroll, threads)
                med, mn = timeit(stepper(_lib.SGHMC, 24), a.iters)
                report(tag, 24, med, mn)
    ops.set_launch_config(0, 0, 0)
    # 2. every variant at the default launch shape
    table = [("SGHMC philox recip", _lib.SGHMC, 24, {}),
             ("SGHMC philox ieee-div", _lib.SGHMC, 24, dict(div=_lib.DIV_IEEE)),
             ("SGHMC injected-noise", _lib.SGHMC, 28, dict(philox=False)),
             ("SGHMC philox uninformative-bias (295 runs)", _lib.SGHMC, 24, dict(bias="uninformative")),
             ("SGLD mu=0.5 philox", _lib.SGLD, 24, dict(mu=0.5)),
             ("SGLD mu=0 philox", _lib.SGLD, 16, {}),
             ("cSGHMC philox", _lib.CSGHMC, 20, {}),
             ("Adam-SGHMC mu=0.5 philox", _lib.ADAM_SGHMC, 48, dict(mu=0.5)),
             ("Adam-cSGHMC philox", _lib.ADAM_CSGHMC, 40, {}),
             ("Adam-cSGHMC philox ieee-div", _lib.ADAM_CSGHMC, 40, dict(div=_lib.DIV_IEEE))]
    for tag, variant, bpp, kw in table:
        if not want(tag):
            continue
        med, mn = timeit(stepper(variant, bpp, **kw), a.iters)
        report(tag, bpp, med, mn)
    tag = "SGHMC philox per-tensor gradient pointers (%d runs)" % len(lay.segments)
    if want(tag):
        grads = [torch.randn(sg.numel, device=dev) * 1e-2 for sg in lay.segments]
        rd, nr = ops.upload_runs(lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in grads]), dev)
        sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-4, lr_head=1e-2, ND=3680, Ninflate=1e3, prior_sig=1.0, nd=1.0, alpha=0.18)

        def fn_ptr():
            step_no[0] += 1
            ops.step(_lib.SGHMC, buf["theta"], None, buf["theta0"], buf["v"], None, None, None, rd, nr, sc,
                     ops.make_noise(seed=42, subseq=step_no[0]))
        med, mn = timeit(fn_ptr, a.iters)
        report(tag, 24, med, mn)
        del grads
    tag = "SGHMC philox + fused moment capture (bdl_step_capture)"
    if want(tag):
        rd, nr = runs["informative"]
        sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-4, lr_head=1e-2, ND=3680, Ninflate=1e3, prior_sig=1.0, nd=1.0, alpha=0.18)

        def fn_cap():
            step_no[0] += 1
            ops.step(_lib.SGHMC, buf["theta"], buf["g"], buf["theta0"], buf["v"], None, None, None, rd, nr, sc,
                     ops.make_noise(seed=42, subseq=step_no[0]), capture=ops.make_capture("avg", buf["m"], buf["s"], 7))
        med, mn = timeit(fn_cap, a.iters)
        report(tag, 40, med, mn)
    if a.full:
        for threads in (64, 128, 256):
            for unroll in (1,):
                tag = f"Adam-cSGHMC philox T={threads} U={unroll}"
                if not want(tag):
                    continue
                ops.set_launch_config(0, unroll, threads)
                med, mn = timeit(stepper(_lib.ADAM_CSGHMC, 40), a.iters)
                report(tag, 40, med, mn)
                tag = f"SGLD mu=0 philox T={threads} U={unroll}"
                med, mn = timeit(stepper(_lib.SGLD, 16), a.iters)
                report(tag, 16, med, mn)
        ops.set_launch_config(0, 0, 0)
    # 3. capture / draw kernels
    ring = buf["xi"].view(1, n)
    rows = [("moments_avg (running mean / 2nd moment)", 20, lambda: ops.moments_avg(buf["theta"], buf["m"], buf["s"], 7)),
            ("moments_welford", 20, lambda: ops.moments_welford(buf["theta"], buf["m"], buf["s"], 7)),
            ("posterior draw (philox)", 12, lambda: ops.draw(buf["theta"], buf["s"], buf["b"], ops.VAR_FROM_MOMENTS, 1.1,
                                                            ops.make_noise(seed=1, subseq=3, stream_id=1))),
            ("sample-ring TMA copy", 8, lambda: ops.capture_ring(buf["theta"], ring, 0)),
            ("sample-ring TMA copy per_cta=2", 8, lambda: (ops.set_ring_config(2), ops.capture_ring(buf["theta"], ring, 0))),
            ("sample-ring TMA copy per_cta=8", 8, lambda: (ops.set_ring_config(8), ops.capture_ring(buf["theta"], ring, 0))),
            ("sample-ring TMA copy per_cta=16", 8, lambda: (ops.set_ring_config(16), ops.capture_ring(buf["theta"], ring, 0))),
            ("sample-ring TMA copy per_cta=4 (default)", 8, lambda: (ops.set_ring_config(4), ops.capture_ring(buf["theta"], ring, 0))),
            ("torch copy_ (reference point for the peak)", 8, lambda: buf["b"].copy_(buf["theta"]))]
    for tag, bpp, fn in rows:
        if not want(tag):
            continue
        med, mn = timeit(fn, a.iters)
        report(tag, bpp, med, mn)
    if want("sample-ring TMA copy"):
        ops.capture_ring(buf["theta"], ring, 0)
        assert torch.equal(ring[0], buf["theta"]), "ring copy mismatch"
        print("ring copy verified")


if __name__ == "__main__":
    main()
