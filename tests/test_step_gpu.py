"""GPU parity: the fused CUDA step (through the C ABI) against the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu

EXACT = ("sgld", "csgld", "sghmc", "csghmc")


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def _stepper(name, device, per_tensor_runs=False):
    from gpu_impl import GpuStepper
    z, hp, method = gu.load_step_case(name)
    return GpuStepper(z["names"].tolist(), z["sizes"].tolist(), "classifier", hp["bias"], device,
                      per_tensor_runs=per_tensor_runs)


@pytest.mark.parametrize("name", gu.step_cases())
@pytest.mark.parametrize("chained", [False, True])
def test_step_matches_reference_golden(cuda_device, name, chained):
    """IEEE-division mode, injected noise: bit-exact vs the reference for SGLD/cSGLD/SGHMC/cSGHMC, within the
    north-star fp32 rel 1e-6 for the Adam variants (whose CPU reference has a 1-ulp sqrt, see test_oracle_golden)."""
    _, _, method = gu.load_step_case(name)
    pairs = gu.replay_step_case(name, _stepper(name, cuda_device), chained=chained, div_mode="true")
    for key, lst in pairs.items():
        for t, (got, want) in enumerate(lst):
            if method in EXACT or key in ("m", "s"):
                assert bits_equal(got, want), f"{name} {key} step {t}"
            else:
                assert gu.max_rel(got, want) <= 1e-6, f"{name} {key} step {t}: {gu.max_rel(got, want):.2e}"


@pytest.mark.parametrize("name", gu.step_cases())
@pytest.mark.parametrize("div_mode", ["true", "recip"])
@pytest.mark.parametrize("per_tensor_runs", [False, True])
def test_step_bit_exact_vs_oracle(cuda_device, name, div_mode, per_tensor_runs):
    """Both division semantics, merged and per-tensor run tables: CUDA == oracle bit for bit (all variants)."""
    gpu = gu.replay_step_case(name, _stepper(name, cuda_device, per_tensor_runs), chained=True, div_mode=div_mode)
    ora = gu.replay_step_case(name, so, chained=True, div_mode=div_mode)
    for key in gpu:
        for t, ((got, _), (want, _)) in enumerate(zip(gpu[key], ora[key])):
            assert bits_equal(got, want), f"{name} {key} step {t} ({div_mode})"


def _random_layout(rng, ntensors, max_numel):
    from bayesdll_b200.flat import FlatLayout
    shapes = []
    for i in range(ntensors):
        numel = int(rng.integers(1, max_numel))
        kind = ("weight", "bias")[i % 2]
        prefix = "classifier" if i >= ntensors - 2 else f"layers.{i // 2}"
        shapes.append((f"{prefix}.{kind}", (numel,)))
    return FlatLayout(shapes, "classifier")


@pytest.mark.parametrize("variant_name", ["sgld", "sghmc", "csghmc", "adam_sghmc", "adam_csghmc"])
@pytest.mark.parametrize("bias_mode", ["informative", "uninformative"])
@pytest.mark.parametrize("own_g", [False, True])
def test_step_ragged_layout_vs_oracle(cuda_device, variant_name, bias_mode, own_g):
    """Ragged tensors (numel 1..5000, many not multiples of 4), multi-tile grids, optional per-run gradient
    pointers with tail masking.  Oracle runs on the same padded buffers."""
    from bayesdll_b200 import _lib, ops
    rng = np.random.default_rng(hash((variant_name, bias_mode, own_g)) % 2**32)
    lay = _random_layout(rng, 41, 5000)
    n = lay.n_padded
    is_head, P = lay.per_element(bias_mode)
    hp = so.HParams(ND=1840, Ninflate=3.0, prior_sig=0.9, nd=0.7, alpha=0.18, beta1=0.9, beta2=0.999, eps=1e-8,
                    temperature=1.3, mu=0.5 if variant_name in ("sgld", "adam_sghmc") else 0.0)
    lrb, lrh = 1e-3, 1e-2
    f = lambda scale=1.0: (rng.standard_normal(n) * scale).astype(np.float32)
    theta, theta0, v, m, xi = f(0.1), f(0.1), f(0.01), f(0.01), f()
    s = np.abs(f(1e-3)).astype(np.float32) + np.float32(1e-6)
    buf = f(0.01)
    g_dense = [rng.standard_normal(sg.numel).astype(np.float32) * 0.05 for sg in lay.segments]
    g = np.zeros(n, np.float32)
    for sg, gd in zip(lay.segments, g_dense):
        g[sg.begin:sg.begin + sg.numel] = gd
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_sghmc=_lib.ADAM_SGHMC,
                   adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    dev = cuda_device
    T = {k: torch.from_numpy(a.copy()).to(dev) for k, a in
         dict(theta=theta, theta0=theta0, v=v, m=m, s=s, buf=buf, xi=xi, g=g).items()}
    keep = []
    if own_g:
        ptrs = []
        for gd in g_dense:
            # allocate each gradient separately, poison what follows its end to prove the tail mask works
            t = torch.full((gd.size + 8,), float("nan"), device=dev)
            t[:gd.size] = torch.from_numpy(gd).to(dev)
            keep.append(t)
            ptrs.append(t.data_ptr())
        tab = lay.run_table(bias_mode, grad_ptrs=ptrs)
        g_arg = None
    else:
        tab = lay.run_table(bias_mode)
        g_arg = T["g"]
    runs_dev, nruns = ops.upload_runs(tab, dev)
    runs_dev_only, _ = ops.upload_runs(tab, dev)
    del runs_dev_only._bdl_host        # no host copy: the table is searched in device memory (step_table_kernel / generic build)
    # with the host copy a pointer table rides in the kernel arguments (step_ptable_kernel); both must give the oracle's bits
    for (div_name, div), runs_dev in [(d, r) for d in (("true", _lib.DIV_IEEE), ("recip", _lib.DIV_RECIP))
                                      for r in (runs_dev, runs_dev_only)]:
        for k, a in dict(theta=theta, v=v, m=m, s=s, buf=buf).items():
            T[k].copy_(torch.from_numpy(a))
        sc = ops.make_scalars(variant, lr_body=lrb, lr_head=lrh, ND=hp.ND, Ninflate=hp.Ninflate, prior_sig=hp.prior_sig,
                              nd=hp.nd, alpha=hp.alpha, mu=hp.mu, beta1=hp.beta1, beta2=hp.beta2, eps=hp.eps,
                              temperature=hp.temperature, t=3, first_step=False, add_noise=True, div_mode=div)
        adam = variant_name.startswith("adam")
        ops.step(variant, T["theta"], g_arg, None if variant_name == "csghmc" else T["theta0"],
                 None if variant_name == "sgld" else T["v"], T["m"] if adam else None, T["s"] if adam else None,
                 T["buf"] if hp.mu != 0 else None, runs_dev, nruns, sc, ops.make_noise(xi=T["xi"]))
        torch.cuda.synchronize()
        kw = dict(is_head=is_head, lr_body=lrb, lr_head=lrh, hp=hp)
        if variant_name == "sgld":
            want = dict(zip(("theta", "buf"), so.step_sgld(theta, g, theta0, buf, xi, P=P, first_step=False,
                                                            div_mode=div_name, **kw)))
        elif variant_name == "sghmc":
            want = dict(zip(("theta", "v"), so.step_sghmc(theta, g, theta0, v, xi, P=P, div_mode=div_name, **kw)))
        elif variant_name == "csghmc":
            want = dict(zip(("theta", "v"), so.step_csghmc(theta, g, v, xi, should_sample=True, **kw)))
        elif variant_name == "adam_sghmc":
            want = dict(zip(("theta", "v", "m", "s", "buf"),
                            so.step_adam_sghmc(theta, g, theta0, v, m, s, buf, xi, P=P, t=3, first_step=False,
                                               div_mode=div_name, **kw)))
        else:
            want = dict(zip(("theta", "v", "m", "s"),
                            so.step_adam_csghmc(theta, g, theta0, v, m, s, xi, P=P, t=3, div_mode=div_name, **kw)))
        for k, w in want.items():
            got = T[k].cpu().numpy()
            assert not np.isnan(got).any(), f"{k}: NaN leaked from gradient padding"
            assert bits_equal(got, w), f"{variant_name} {k} ({div_name}, own_g={own_g}): " \
                                       f"{(got.view(np.uint32) != w.view(np.uint32)).sum()} mismatches"


def test_step_empty_and_argument_errors(cuda_device):
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    lay = FlatLayout([("a.weight", (8,))], "classifier")
    runs_dev, nruns = ops.upload_runs(lay.run_table("informative"), cuda_device)
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-3, ND=10)
    z = lambda n: torch.zeros(n, device=cuda_device)
    # n == 0 is a no-op
    e = torch.zeros(0, device=cuda_device)
    ops.step(_lib.SGHMC, e, e, e, e, None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=1))
    # n not a multiple of 4
    with pytest.raises(_lib.BdlError, match="multiple of 4"):
        ops.step(_lib.SGHMC, z(6), z(6), z(6), z(6), None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=1))
    # missing momentum buffer
    with pytest.raises(_lib.BdlError, match="momentum"):
        ops.step(_lib.SGHMC, z(8), z(8), z(8), None, None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=1))
    # unaligned pointer
    with pytest.raises(_lib.BdlError, match="aligned"):
        ops.step(_lib.SGHMC, z(9)[1:], z(8), z(8), z(8), None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=1))
    # CPU tensors are rejected: there is no CPU path
    with pytest.raises(_lib.BdlError, match="CUDA tensor"):
        ops.step(_lib.SGHMC, torch.zeros(8), z(8), z(8), z(8), None, None, None, runs_dev, nruns, sc,
                 ops.make_noise(seed=1))


@pytest.mark.parametrize("bias_mode", ["informative", "uninformative"])
def test_inline_run_table_equals_device_run_table(cuda_device, bias_mode):
    """Tables of <= 8 runs ride in the kernel arguments when the host copy is supplied; the device-table path (warp
    search + cursor) must give the same bits.  Also covers a head run that starts in the middle of a warp's tile."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    lay = FlatLayout([("b.0.weight", (70_001,)), ("b.0.bias", (13,)), ("b.1.weight", (333,)), ("b.1.bias", (5,)),
                      ("classifier.weight", (37, 11)), ("classifier.bias", (37,))], "classifier")
    n = lay.n_padded
    tab = lay.run_table(bias_mode)
    assert len(tab) <= 8
    gen = torch.Generator(device=cuda_device).manual_seed(0)
    init = {k: torch.randn(n, device=cuda_device, generator=gen) * sc for k, sc in
            dict(theta=0.1, g=0.05, theta0=0.1, v=0.01).items()}
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=10.0, alpha=0.18)
    outs = []
    for use_host in (True, False):
        runs_dev, nruns = ops.upload_runs(tab, cuda_device)
        if not use_host:
            del runs_dev._bdl_host
        st = {k: v.clone() for k, v in init.items()}
        ops.step(_lib.SGHMC, st["theta"], st["g"], st["theta0"], st["v"], None, None, None, runs_dev, nruns, sc,
                 ops.make_noise(seed=5, subseq=2))
        outs.append(st)
    assert torch.equal(outs[0]["theta"], outs[1]["theta"]) and torch.equal(outs[0]["v"], outs[1]["v"])


@pytest.mark.parametrize("variant_name", ["sgld", "sghmc", "csghmc", "adam_sghmc", "adam_csghmc"])
@pytest.mark.parametrize("kind", ["avg", "avg_nomom2", "welford"])
@pytest.mark.parametrize("own_g", [False, True, "device-table"])
def test_fused_capture_equals_step_then_moments(cuda_device, variant_name, kind, own_g):
    """bdl_step_capture == bdl_step followed by bdl_moments_avg / bdl_moments_welford, bit for bit: first sample (init)
    and later samples, both division modes, in-kernel Philox, ragged layout, a tensor without gradient (BDL_CLS_SKIP:
    left untouched by the step, still part of the captured sample) and -- via the oracle -- the reference arithmetic."""
    from bayesdll_b200 import _lib, ops
    rng = np.random.default_rng(abs(hash((variant_name, kind, own_g))) % 2**32)
    lay = _random_layout(rng, 23, 4000)
    n = lay.n_padded
    dev = cuda_device
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_sghmc=_lib.ADAM_SGHMC,
                   adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    mu = 0.5 if variant_name in ("sgld", "adam_sghmc") else 0.0
    adam = variant_name.startswith("adam")
    f = lambda scale=1.0: torch.from_numpy((rng.standard_normal(n) * scale).astype(np.float32)).to(dev)
    init_state = dict(theta=f(0.1), v=f(0.01), m=f(0.01), s=f(1e-3).abs() + 1e-6, buf=f(0.01))
    theta0, g = f(0.1), f(0.05)
    keep, ptrs = [], None
    if own_g:
        ptrs = []
        for sg in lay.segments:
            t = torch.full((sg.numel + 8,), float("nan"), device=dev)
            t[:sg.numel] = g[sg.begin:sg.begin + sg.numel]
            keep.append(t)
            ptrs.append(t.data_ptr())
    tab = lay.run_table("uninformative", grad_ptrs=ptrs if own_g else [0] * len(lay.segments))
    skip_idx = 5
    tab[skip_idx].cls |= _lib.CLS_SKIP                                  # this tensor has p.grad None
    runs_dev, nruns = ops.upload_runs(tab, dev)
    if own_g == "device-table":     # no host copy: step_table_kernel searches the table in device memory; with it
        del runs_dev._bdl_host      # (own_g is True) the pointer table rides in the kernel arguments (step_ptable_kernel)
    for div in (_lib.DIV_IEEE, _lib.DIV_RECIP):
        sc = ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=3.0, prior_sig=0.9, nd=0.7, alpha=0.18,
                              mu=mu, t=4, first_step=False, add_noise=True, div_mode=div)
        A = {k: t.clone() for k, t in init_state.items()}               # fused
        B = {k: t.clone() for k, t in init_state.items()}               # step, then capture
        capA = [torch.full((n,), 7.0, device=dev), torch.full((n,), 7.0, device=dev)]
        capB = [t.clone() for t in capA]
        second = None if kind == "avg_nomom2" else 1

        def run(S, cap, fused, cnt, init, subseq):
            spec = None
            if fused:
                spec = ops.make_capture("welford" if kind == "welford" else "avg", cap[0],
                                        None if second is None else cap[1], cnt, init=init)
            ops.step(variant, S["theta"], None if own_g else g, None if variant_name == "csghmc" else theta0,
                     None if variant_name == "sgld" else S["v"], S["m"] if adam else None, S["s"] if adam else None,
                     S["buf"] if mu else None, runs_dev, nruns, sc, ops.make_noise(seed=9, subseq=subseq), capture=spec)
            if not fused:
                if kind == "welford":
                    ops.moments_welford(S["theta"], cap[0], cap[1], cnt, init=init, div_mode=div)
                else:
                    ops.moments_avg(S["theta"], cap[0], None if second is None else cap[1], cnt, init=init, div_mode=div)

        seg = lay.segments[skip_idx]
        before = A["theta"][seg.begin:seg.end].clone()
        for step_no, (cnt, init) in enumerate([(1 if kind == "welford" else 0, True), (3, False), (5, False)]):
            run(A, capA, True, cnt, init, step_no)
            run(B, capB, False, cnt, init, step_no)
        torch.cuda.synchronize()
        for k in A:
            assert torch.equal(A[k], B[k]), f"{variant_name} {k}: fused step differs from the plain step (div={div})"
        assert torch.equal(capA[0], capB[0]), f"{variant_name}/{kind}: first moment differs (div={div})"
        if second is not None:
            assert torch.equal(capA[1], capB[1]), f"{variant_name}/{kind}: second moment differs (div={div})"
        else:
            assert torch.equal(capA[1], torch.full((n,), 7.0, device=dev))          # untouched
        assert torch.isfinite(capA[0]).all()
        assert torch.equal(A["theta"][seg.begin:seg.end], before)                  # skipped tensor left untouched ...
        assert not torch.equal(capA[0][seg.begin:seg.end], torch.full((seg.end - seg.begin,), 7.0, device=dev))  # ... but captured


def test_fused_capture_argument_errors(cuda_device):
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    lay = FlatLayout([("a.weight", (64,))], "classifier")
    runs_dev, nruns = ops.upload_runs(lay.run_table("informative"), cuda_device)
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-3, ND=10)
    z = lambda n: torch.zeros(n, device=cuda_device)
    with pytest.raises(_lib.BdlError, match="length"):
        ops.step(_lib.SGHMC, z(64), z(64), z(64), z(64), None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=1),
                 capture=ops.make_capture("avg", z(32), None, 1))
    with pytest.raises(_lib.BdlError, match="tensor required"):
        ops.make_capture("welford", z(64), None, 1)
    cap = ops.make_capture("avg", z(64), z(64), 1)
    cap.kind = 9
    with pytest.raises(_lib.BdlError, match="capture kind"):
        ops.step(_lib.SGHMC, z(64), z(64), z(64), z(64), None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=1), capture=cap)


@pytest.mark.parametrize("variant_name", ["sgld", "sgld_mu0", "sghmc", "csghmc", "adam_sghmc", "adam_csghmc"])
@pytest.mark.parametrize("philox", [True, False])
def test_fast_path_build_equals_generic_build(cuda_device, variant_name, philox):
    """The library's default launch for a two-run (body | head) inline table with a flat gradient is a leaner build of the
    same kernel (no tile loop, no table dispatch).  An explicit launch-shape request runs the generic build: same bits,
    both division modes, ragged tail, head boundary in the middle of a tile, and a skipped head run."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    dev = cuda_device
    base = variant_name.split("_mu")[0]
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_sghmc=_lib.ADAM_SGHMC,
                   adam_csghmc=_lib.ADAM_CSGHMC)[base]
    mu = 0.5 if variant_name in ("sgld", "adam_sghmc") else 0.0
    adam = base.startswith("adam")
    default_threads = 256 if variant_name == "sgld_mu0" else (128 if base == "csghmc" else 64)
    lay = FlatLayout([("body.weight", (1_000_003,)), ("classifier.weight", (37, 1021)), ("classifier.bias", (37,))], "classifier")
    n = lay.n_padded
    gen = torch.Generator(device=dev).manual_seed(5)
    init = {k: torch.randn(n, device=dev, generator=gen) * sc for k, sc in
            dict(theta=0.1, g=0.05, theta0=0.1, v=0.01, m=0.01, buf=0.01, xi=1.0).items()}
    init["s"] = torch.rand(n, device=dev, generator=gen) * 1e-3 + 1e-6
    for skip_head in (False, True):
        tab = lay.run_table("informative")
        assert len(tab) == 2
        if skip_head:
            tab[1].cls |= _lib.CLS_SKIP
        runs_dev, nruns = ops.upload_runs(tab, dev)
        for div in (_lib.DIV_RECIP, _lib.DIV_IEEE):
            sc = ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=10.0, prior_sig=0.9, nd=0.7, alpha=0.18,
                                  mu=mu, t=5, temperature=1.3, div_mode=div)
            res = []
            for cfg in ((0, 0, 0), (0, 1, default_threads), (0, 2, 128)):
                ops.set_launch_config(*cfg)
                st = {k: v.clone() for k, v in init.items()}
                nz = ops.make_noise(seed=11, subseq=4) if philox else ops.make_noise(xi=st["xi"])
                ops.step(variant, st["theta"], st["g"], None if base == "csghmc" else st["theta0"],
                         None if base == "sgld" else st["v"], st["m"] if adam else None, st["s"] if adam else None,
                         st["buf"] if mu else None, runs_dev, nruns, sc, nz)
                torch.cuda.synchronize()
                ops.set_launch_config(0, 0, 0)
                res.append(st)
            for k in ("theta", "v", "m", "s", "buf"):
                assert torch.equal(res[0][k], res[1][k]) and torch.equal(res[0][k], res[2][k]), (k, skip_head, div)
            assert not torch.equal(res[0]["theta"], init["theta"])
            if skip_head:
                hb = lay.segments[1].begin
                assert torch.equal(res[0]["theta"][hb:], init["theta"][hb:])


@pytest.mark.parametrize("variant_name", ["sghmc", "csghmc", "adam_csghmc", "sgld"])
@pytest.mark.parametrize("kind", ["avg", "welford"])
def test_fast_path_with_capture_equals_step_then_moments(cuda_device, variant_name, kind):
    """The lean default build also carries the fused capture (two-run inline table + flat gradient, the runners' launch
    after burn-in when the gradient is gathered): identical to the plain step followed by the stand-alone moment kernel."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    dev = cuda_device
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    mu = 0.5 if variant_name == "sgld" else 0.0
    adam = variant_name.startswith("adam")
    lay = FlatLayout([("body.weight", (500_003,)), ("classifier.weight", (37, 301)), ("classifier.bias", (37,))], "classifier")
    n = lay.n_padded
    gen = torch.Generator(device=dev).manual_seed(8)
    init = {k: torch.randn(n, device=dev, generator=gen) * sc for k, sc in
            dict(theta=0.1, g=0.05, theta0=0.1, v=0.01, m=0.01, buf=0.01).items()}
    init["s"] = torch.rand(n, device=dev, generator=gen) * 1e-3 + 1e-6
    runs_dev, nruns = ops.upload_runs(lay.run_table("informative"), dev)
    for div in (_lib.DIV_RECIP, _lib.DIV_IEEE):
        sc = ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=10.0, prior_sig=0.9, nd=0.7, alpha=0.18,
                              mu=mu, t=5, temperature=1.3, div_mode=div)
        A = {k: v.clone() for k, v in init.items()}
        B = {k: v.clone() for k, v in init.items()}
        capA = [torch.full((n,), 3.0, device=dev), torch.full((n,), 3.0, device=dev)]
        capB = [t.clone() for t in capA]
        for step_no, (cnt, first) in enumerate([(1 if kind == "welford" else 0, True), (3, False)]):
            args = lambda S: (variant, S["theta"], S["g"], None if variant_name == "csghmc" else S["theta0"],
                              None if variant_name == "sgld" else S["v"], S["m"] if adam else None, S["s"] if adam else None,
                              S["buf"] if mu else None, runs_dev, nruns, sc, ops.make_noise(seed=2, subseq=step_no))
            ops.step(*args(A), capture=ops.make_capture(kind, capA[0], capA[1], cnt, init=first))
            ops.step(*args(B))
            (ops.moments_welford if kind == "welford" else ops.moments_avg)(B["theta"], capB[0], capB[1], cnt, init=first, div_mode=div)
        torch.cuda.synchronize()
        for k in A:
            assert torch.equal(A[k], B[k]), (k, div)
        assert torch.equal(capA[0], capB[0]) and torch.equal(capA[1], capB[1]), div
