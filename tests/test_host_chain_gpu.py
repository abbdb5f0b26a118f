"""GPU: the host-buffer chain API (bdl_chain_*) == the device-resident step, independent of the chunking."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant_name,mu", [("sghmc", 0.0), ("sgld", 0.5), ("adam_csghmc", 0.0), ("csghmc", 0.0)])
@pytest.mark.parametrize("chunk", [0, 4096, 100_000])
def test_host_chain_matches_device_step(cuda_device, variant_name, mu, chunk):
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    lay = FlatLayout([("b.weight", (250_003,)), ("b.bias", (1001,)), ("classifier.weight", (37, 64)), ("classifier.bias", (37,))],
                     "classifier")
    n = lay.n_padded
    gen = torch.Generator().manual_seed(5)
    theta = torch.randn(n, generator=gen) * 0.1
    theta0 = torch.randn(n, generator=gen) * 0.1
    tab = lay.run_table("uninformative")
    sc = lambda t, first: ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=10.0, nd=0.5, alpha=0.18,
                                           mu=mu, t=t, first_step=first, div_mode=_lib.DIV_RECIP)
    # device-resident reference trajectory
    dev = cuda_device
    adam = variant_name.startswith("adam")
    D = dict(theta=theta.to(dev), theta0=theta0.to(dev), v=torch.zeros(n, device=dev), m=torch.zeros(n, device=dev),
             s=torch.zeros(n, device=dev), buf=torch.zeros(n, device=dev))
    runs_dev, nruns = ops.upload_runs(tab, dev)
    grads = [torch.randn(n, generator=gen) * 0.05 for _ in range(3)]
    for t, g in enumerate(grads, 1):
        ops.step(variant, D["theta"], g.to(dev), None if variant == _lib.CSGHMC else D["theta0"],
                 None if variant == _lib.SGLD else D["v"], D["m"] if adam else None, D["s"] if adam else None,
                 D["buf"] if mu else None, runs_dev, nruns, sc(t, t == 1), ops.make_noise(seed=11, subseq=t))
    torch.cuda.synchronize()
    # host-buffer chain
    ch = ops.HostChain(n, variant, with_sgd_momentum=bool(mu), chunk_elems=chunk)
    ch.upload(_lib.BUF_THETA, theta)
    if variant != _lib.CSGHMC:
        ch.upload(_lib.BUF_THETA0, theta0)
    out = torch.empty(n).pin_memory()
    for t, g in enumerate(grads, 1):
        ch.step_host(g.pin_memory(), out, tab, sc(t, t == 1), ops.make_noise(seed=11, subseq=t))
    assert torch.equal(out, D["theta"].cpu())
    back = torch.empty(n)
    assert torch.equal(ch.download(_lib.BUF_THETA, back), D["theta"].cpu())
    if variant != _lib.SGLD:
        assert torch.equal(ch.download(_lib.BUF_V, back), D["v"].cpu())
    ch.close()


def test_host_chain_rejects_unpinned_and_bad_args(cuda_device):
    from bayesdll_b200 import _lib, ops
    ch = ops.HostChain(1024, _lib.SGHMC)
    tab = (_lib.Run * 1)()
    tab[0].begin, tab[0].end, tab[0].valid_end, tab[0].cls = 0, 1024, 1024, _lib.CLS_PRIOR
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-3, ND=10)
    with pytest.raises(ops.BdlError, match="pinned"):
        ch.step_host(torch.zeros(1024), torch.zeros(1024), tab, sc, ops.make_noise(seed=1))
    with pytest.raises(ops.BdlError, match="length"):
        ch.upload(_lib.BUF_THETA, torch.zeros(8))
    with pytest.raises(ops.BdlError):
        ch.upload(_lib.BUF_M, torch.zeros(1024))          # SGHMC chain has no Adam state
    with pytest.raises(ops.BdlError):
        ops.HostChain(10, _lib.SGHMC)                     # n not a multiple of 4
    ch.close()
