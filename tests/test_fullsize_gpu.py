"""GPU, BASELINE.json's full sizes: one fused step over the real ViT-L/32 / ResNet-101 layouts compared element by
element (bit-exact) with the C oracle on the host, plus size-independent properties (determinism, independence of
launch shape and of host-buffer chunking, untouched padding)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {
    # BASELINE.json configs[2]: ViT-L/32 SGHMC with net0 prior mean
    "vit_l_32-sghmc": ("vit_l_32", "sghmc", dict(lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, prior_sig=1.0, nd=1.0,
                                                 alpha=0.18)),
    # configs[1]: ResNet-101 cSGHMC, cyclical lr
    "resnet101-csghmc": ("resnet101", "csghmc", dict(lr_body=7.3e-5, lr_head=7.3e-3, ND=1840, Ninflate=1.0, prior_sig=1.0,
                                                     nd=0.01, alpha=0.18)),
    # configs[3]: ViT-L/32 Adam-cSGHMC
    "vit_l_32-adam_csghmc": ("vit_l_32", "adam_csghmc", dict(lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, prior_sig=1.0,
                                                             nd=1.0, alpha=0.05, beta1=0.9, beta2=0.999, eps=1e-8,
                                                             temperature=1.0, t=3)),
}


@pytest.mark.parametrize("case", list(CASES))
def test_full_size_step_bit_exact_vs_c_oracle(cuda_device, case):
    from bayesdll_b200 import _lib, ops, shapes
    from bayesdll_b200.flat import FlatLayout
    from oracle import c_oracle
    backbone, vname, kw = CASES[case]
    variant = dict(sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_csghmc=_lib.ADAM_CSGHMC)[vname]
    named, readout = shapes.named_shapes(backbone, 37)
    lay = FlatLayout(named, readout)
    n = lay.n_padded
    dev = cuda_device
    gen = torch.Generator(device=dev).manual_seed(42)
    adam = vname.startswith("adam")
    D = dict(theta=torch.randn(n, device=dev, generator=gen) * 0.02, g=torch.randn(n, device=dev, generator=gen) * 0.01,
             theta0=torch.randn(n, device=dev, generator=gen) * 0.02, v=torch.randn(n, device=dev, generator=gen) * 1e-3)
    if adam:
        D["m"] = torch.randn(n, device=dev, generator=gen) * 1e-3
        D["s"] = torch.rand(n, device=dev, generator=gen) * 1e-5 + 1e-9
    D["theta"][lay.n_dense:] = 0            # padding
    H = {k: t.cpu().numpy() for k, t in D.items()}
    tab = lay.run_table("uninformative")
    runs_dev, nruns = ops.upload_runs(tab, dev)
    sc = ops.make_scalars(variant, div_mode=_lib.DIV_RECIP, **kw)
    seed, sub = 42, 11
    args = lambda T: (T["theta"], T["g"], None if variant == _lib.CSGHMC else T["theta0"], T["v"], T.get("m"), T.get("s"), None)

    # determinism + launch-shape independence at full size
    A = {k: t.clone() for k, t in D.items()}
    ops.set_launch_config(3, 2, 512)
    ops.step(variant, *args(A), runs_dev, nruns, sc, ops.make_noise(seed=seed, subseq=sub))
    ops.set_launch_config(0, 0, 0)
    ops.step(variant, *args(D), runs_dev, nruns, sc, ops.make_noise(seed=seed, subseq=sub))
    torch.cuda.synchronize()
    for k in ("theta", "v") + (("m", "s") if adam else ()):
        assert torch.equal(A[k], D[k]), f"{k}: result depends on the launch shape"
    del A

    # the same step on the host with the same noise stream
    xi = torch.empty(n, device=dev)
    ops.philox_normal(xi, seed, _lib.STREAM_STEP, sub)
    xi_h = xi.cpu().numpy()
    del xi
    nz = _lib.Noise()
    nz.xi_dev = xi_h.ctypes.data
    c_oracle.step(variant, H["theta"], H["g"], None if variant == _lib.CSGHMC else H["theta0"], H["v"], H.get("m"), H.get("s"),
                  None, tab, sc, nz)
    for k in ("theta", "v") + (("m", "s") if adam else ()):
        got = D[k].cpu().numpy()
        neq = int((got.view(np.uint32) != H[k].view(np.uint32)).sum())
        assert neq == 0, f"{case} {k}: {neq} of {n} elements differ from the C oracle"
    # checksum of the whole updated state (a cheap regression handle printed on failure of later rounds)
    assert np.isfinite(float(D["theta"].double().sum().item()))


def test_full_size_host_chain_equals_device_step(cuda_device):
    """ViT-L/32: bdl_chain_step_host (19 pipelined chunks) == one device-resident launch, bit for bit."""
    from bayesdll_b200 import _lib, ops, shapes
    from bayesdll_b200.flat import FlatLayout
    named, readout = shapes.named_shapes("vit_l_32", 37)
    lay = FlatLayout(named, readout)
    n = lay.n_padded
    dev = cuda_device
    gen = torch.Generator(device=dev).manual_seed(1)
    theta = torch.randn(n, device=dev, generator=gen) * 0.02
    theta0 = torch.randn(n, device=dev, generator=gen) * 0.02
    g = torch.randn(n, device=dev, generator=gen) * 0.01
    v = torch.zeros(n, device=dev)
    tab = lay.run_table("informative")
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, alpha=0.18)
    ch = ops.HostChain(n, _lib.SGHMC)
    stage = torch.empty(n).pin_memory()
    stage.copy_(theta); ch.upload(_lib.BUF_THETA, stage)
    stage.copy_(theta0); ch.upload(_lib.BUF_THETA0, stage)
    stage.copy_(g)
    out = torch.empty(n).pin_memory()
    ch.step_host(stage, out, tab, sc, ops.make_noise(seed=3, subseq=1))
    runs_dev, nruns = ops.upload_runs(tab, dev)
    ops.step(_lib.SGHMC, theta, g, theta0, v, None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=3, subseq=1))
    assert torch.equal(theta.cpu(), out)
    ch.close()
