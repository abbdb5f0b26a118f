"""In-kernel Philox noise: known-answer / oracle agreement, distribution tests (KS, moments, correlations),
grid-shape independence and equivalence of in-kernel vs. injected noise (north star parity criterion ii)."""
import numpy as np
import pytest
import torch
from scipy import stats

pytestmark = pytest.mark.gpu


def _normals(dev, n, seed, stream_id=2, subseq=0):
    from bayesdll_b200 import ops
    out = torch.empty(n, device=dev)
    ops.philox_normal(out, seed, stream_id, subseq)
    torch.cuda.synchronize()
    return out


def test_philox_matches_c_oracle(cuda_device):
    from oracle import philox_oracle
    n = 1 << 16
    for seed, sid, sub in [(0, 0, 0), (42, 0, 7), (0xDEADBEEFCAFEF00D, 1, (5 << 32) | 3)]:
        got = _normals(cuda_device, n, seed, sid, sub).cpu().numpy()
        want = philox_oracle.normal(n, seed, sid, sub)
        err = np.abs(got - want)
        # radius/angle use MUFU approximations on the GPU (lg2, sin, cos): |err| ~ 1e-6, a handful of
        # elements with u within 1e-5 of 1 can reach 1e-3 (DESIGN.md "Noise").
        assert np.mean(err > 2e-5) <= 1e-4
        assert err.max() < 5e-3


def test_philox_distribution(cuda_device):
    n = 1 << 24
    x = _normals(cuda_device, n, seed=1234).double()
    mean, var = x.mean().item(), x.var().item()
    skew = ((x - mean) ** 3).mean().item() / var ** 1.5
    kurt = ((x - mean) ** 4).mean().item() / var ** 2
    assert abs(mean) < 5 / np.sqrt(n)
    assert abs(var - 1) < 5 * np.sqrt(2 / n)
    assert abs(skew) < 5 * np.sqrt(6 / n)
    assert abs(kurt - 3) < 5 * np.sqrt(24 / n)
    sub = x[:: n // (1 << 20)][: 1 << 20].cpu().numpy()
    assert stats.kstest(sub, "norm").pvalue > 1e-3
    # lag-1, lag-2 (within a Box-Muller pair / across pairs) and lag-4 (across Philox calls) correlation
    for lag in (1, 2, 4, 1024):
        c = (x[:-lag] * x[lag:]).mean().item()
        assert abs(c) < 5 / np.sqrt(n), f"lag {lag} corr {c}"
    # tails
    assert abs((x.abs() > 3).double().mean().item() - 2 * stats.norm.sf(3)) < 5 * np.sqrt(2.7e-3 / n)
    assert torch.isfinite(x).all()


def test_philox_streams_are_independent(cuda_device):
    n = 1 << 22
    a = _normals(cuda_device, n, seed=9, subseq=100).double()
    b = _normals(cuda_device, n, seed=9, subseq=101).double()      # next step
    c = _normals(cuda_device, n, seed=10, subseq=100).double()     # other chain
    d = _normals(cuda_device, n, seed=9, stream_id=1, subseq=100).double()
    for other in (b, c, d):
        assert abs((a * other).mean().item()) < 5 / np.sqrt(n)
        assert not torch.equal(a, other)
    assert torch.equal(a, _normals(cuda_device, n, seed=9, subseq=100).double())   # reproducible


@pytest.mark.parametrize("variant_name", ["sgld", "sghmc", "csghmc", "adam_sghmc", "adam_csghmc"])
def test_inkernel_noise_equals_injected_stream(cuda_device, variant_name):
    """step(philox) must be bit-identical to step(xi = bdl_philox_normal(same key)), for every launch shape."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    dev = cuda_device
    lay = FlatLayout([("body.weight", (300_001,)), ("body.bias", (1023,)), ("classifier.weight", (37 * 1024,)),
                      ("classifier.bias", (37,))], "classifier")
    n = lay.n_padded
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_sghmc=_lib.ADAM_SGHMC,
                   adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    gen = torch.Generator(device=dev).manual_seed(3)
    init = {k: torch.randn(n, device=dev, generator=gen) * sc for k, sc in
            dict(theta=0.1, g=0.05, theta0=0.1, v=0.01, m=0.01, buf=0.01).items()}
    init["s"] = torch.rand(n, device=dev, generator=gen) * 1e-3 + 1e-6
    runs_dev, nruns = ops.upload_runs(lay.run_table("uninformative"), dev)
    mu = 0.5 if variant_name in ("sgld", "adam_sghmc") else 0.0
    sc = ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=10.0, prior_sig=1.0, nd=1.0, alpha=0.18,
                          mu=mu, t=5, div_mode=_lib.DIV_RECIP)
    seed, step_no = 42, 17
    xi = torch.empty(n, device=dev)
    ops.philox_normal(xi, seed, _lib.STREAM_STEP, step_no)

    def run(noise, cfg):
        ops.set_launch_config(*cfg)
        st = {k: v.clone() for k, v in init.items()}
        adam = variant_name.startswith("adam")
        ops.step(variant, st["theta"], st["g"], None if variant_name == "csghmc" else st["theta0"],
                 None if variant_name == "sgld" else st["v"], st["m"] if adam else None, st["s"] if adam else None,
                 st["buf"] if mu else None, runs_dev, nruns, sc, noise)
        torch.cuda.synchronize()
        ops.set_launch_config(0, 0, 0)
        return st

    ref = run(ops.make_noise(xi=xi), (0, 0, 0))
    for cfg in [(0, 0, 0), (1, 1, 128), (3, 2, 256), (0, 2, 512), (7, 1, 512), (64, 2, 128)]:
        got = run(ops.make_noise(seed=seed, subseq=step_no, stream_id=_lib.STREAM_STEP), cfg)
        for k in ("theta", "v", "m", "s", "buf"):
            assert torch.equal(got[k], ref[k]), f"{variant_name} {k} differs for launch config {cfg}"
