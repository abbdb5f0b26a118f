"""North-star trajectory criterion: 'parameter and momentum trajectories must match within fp32 rel 1e-6 per step and
1e-4 after 1k steps'.  1000 chained steps of the CUDA path (in-kernel Philox) against the C oracle fed with the same
noise stream; every update op is correctly rounded on both sides, so the trajectories are in fact bit-identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant_name,mu", [("sghmc", 0.0), ("adam_csghmc", 0.0), ("sgld", 0.5)])
def test_1k_step_trajectory(cuda_device, variant_name, mu):
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.flat import FlatLayout
    from oracle import c_oracle
    dev = cuda_device
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    lay = FlatLayout([("layers.0.weight", (300, 401)), ("layers.0.bias", (401,)), ("classifier.weight", (37, 401)),
                      ("classifier.bias", (37,))], "classifier")
    n = lay.n_padded
    rng = np.random.default_rng(1)
    f = lambda sc: (rng.standard_normal(n) * sc).astype(np.float32)
    H = dict(theta=f(0.1), theta0=f(0.1), v=np.zeros(n, np.float32), m=np.zeros(n, np.float32), s=np.zeros(n, np.float32),
             buf=np.zeros(n, np.float32))
    D = {k: torch.from_numpy(a.copy()).to(dev) for k, a in H.items()}
    tab = lay.run_table("uninformative")
    runs_dev, nruns = ops.upload_runs(tab, dev)
    adam = variant_name.startswith("adam")
    xi = torch.empty(n, device=dev)
    g_dev = torch.empty(n, device=dev)
    steps = 1000
    worst = 0.0
    for t in range(1, steps + 1):
        # well-scaled synthetic gradient that depends on the current state (a quadratic bowl + noise)
        g = (0.05 * H["theta"] + 0.01 * rng.standard_normal(n).astype(np.float32)).astype(np.float32)
        g_dev.copy_(torch.from_numpy(g))
        sc = ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=1840, Ninflate=10.0, prior_sig=1.0, nd=0.5, alpha=0.18,
                              mu=mu, t=t, first_step=(t == 1), temperature=1.2, div_mode=_lib.DIV_RECIP)
        ops.step(variant, D["theta"], g_dev, D["theta0"], None if variant == _lib.SGLD else D["v"], D["m"] if adam else None,
                 D["s"] if adam else None, D["buf"] if mu else None, runs_dev, nruns, sc,
                 ops.make_noise(seed=7, subseq=t, stream_id=_lib.STREAM_STEP))
        ops.philox_normal(xi, 7, _lib.STREAM_STEP, t)
        xi_h = xi.cpu().numpy()
        nz = _lib.Noise()
        nz.xi_dev = xi_h.ctypes.data
        c_oracle.step(variant, H["theta"], g, H["theta0"], None if variant == _lib.SGLD else H["v"], H["m"] if adam else None,
                      H["s"] if adam else None, H["buf"] if mu else None, tab, sc, nz)
        if t in (1, 10, 100, 500, 1000):
            got = D["theta"].cpu().numpy()
            assert np.isfinite(got).all()
            rel = float(np.max(np.abs(got.astype(np.float64) - H["theta"])) / np.max(np.abs(H["theta"])))
            worst = max(worst, rel)
            assert rel <= (1e-6 if t == 1 else 1e-4), f"step {t}: rel {rel:.2e}"
    for k in ("theta", "v", "m", "s", "buf"):
        assert np.array_equal(D[k].cpu().numpy().view(np.uint32), H[k].view(np.uint32)), f"{k} not bit-identical after {steps} steps"
    assert worst == 0.0
