"""GPU parity of the capture / draw / ensemble / calibration kernels against the oracle and reference goldens."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("n", [4, 1028, 100_004, 3_000_000])
@pytest.mark.parametrize("div_name", ["true", "recip"])
def test_moments_avg_and_welford_bit_exact(cuda_device, n, div_name):
    from bayesdll_b200 import ops
    div = dict(true=0, recip=1)[div_name]
    rng = np.random.default_rng(n)
    th = rng.standard_normal(n).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(cuda_device)
    m1, m2 = so.moments_init(th)
    g1, g2 = torch.empty(n, device=cuda_device), torch.empty(n, device=cuda_device)
    ops.moments_avg(d(th), g1, g2, 0, init=True)
    assert bits_equal(g1.cpu().numpy(), m1) and bits_equal(g2.cpu().numpy(), m2)
    mean, M2 = th.copy(), np.zeros(n, np.float32)
    w1, w2 = torch.empty(n, device=cuda_device), torch.empty(n, device=cuda_device)
    ops.moments_welford(d(th), w1, w2, 1, init=True)
    for cnt in range(1, 6):
        th = rng.standard_normal(n).astype(np.float32)
        m1, m2 = so.moments_avg(th, m1, m2, cnt, div_name)
        ops.moments_avg(d(th), g1, g2, cnt, div_mode=div)
        assert bits_equal(g1.cpu().numpy(), m1) and bits_equal(g2.cpu().numpy(), m2)
        nw = 2 * cnt + 1                                    # the reference's double-counted n: 3, 5, 7, ...
        mean, M2 = so.moments_welford(th, mean, M2, nw, div_name)
        ops.moments_welford(d(th), w1, w2, nw, div_mode=div)
        assert bits_equal(w1.cpu().numpy(), mean) and bits_equal(w2.cpu().numpy(), M2)
    # mom2 == None (nst == 0)
    only1 = torch.empty(n, device=cuda_device)
    ops.moments_avg(d(th), only1, None, 0, init=True)
    assert bits_equal(only1.cpu().numpy(), th * np.float32(1.0))


@pytest.mark.parametrize("n", [8, 40_004, 2_000_000])
def test_posterior_draw_bit_exact_and_philox_equivalent(cuda_device, n):
    from bayesdll_b200 import _lib, ops
    rng = np.random.default_rng(n + 1)
    mean = rng.standard_normal(n).astype(np.float32)
    mom2 = (mean * mean + np.abs(rng.standard_normal(n)).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    mom2[::7] = mean[::7] ** 2 - np.float32(1e-6)           # negative raw variance -> clamp path
    M2 = np.abs(rng.standard_normal(n)).astype(np.float32)
    eps = rng.standard_normal(n).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(cuda_device)
    out = torch.empty(n, device=cuda_device)
    ops.draw(d(mean), d(mom2), out, ops.VAR_FROM_MOMENTS, 5 / 4, ops.make_noise(xi=d(eps)))
    assert bits_equal(out.cpu().numpy(), so.posterior_draw(mean, so.variance_from_moments(mean, mom2, 5 / 4), eps))
    for div_name, div in (("true", 0), ("recip", 1)):
        ops.draw(d(mean), d(M2), out, ops.VAR_FROM_WELFORD, 6.0, ops.make_noise(xi=d(eps)), div_mode=div)
        assert bits_equal(out.cpu().numpy(), so.posterior_draw(mean, so.variance_from_welford(M2, 7, div_name), eps))
    ops.draw(d(mean), None, out, ops.VAR_TINY, 1.0, ops.make_noise(xi=d(eps)))
    assert bits_equal(out.cpu().numpy(), so.posterior_draw(mean, so.variance_from_welford(M2, 1), eps))
    # in-kernel Philox == injected Philox stream
    xi = torch.empty(n, device=cuda_device)
    ops.philox_normal(xi, 77, _lib.STREAM_DRAW, 12345)
    a, b = torch.empty(n, device=cuda_device), torch.empty(n, device=cuda_device)
    ops.draw(d(mean), d(mom2), a, ops.VAR_FROM_MOMENTS, 1.25, ops.make_noise(xi=xi))
    ops.draw(d(mean), d(mom2), b, ops.VAR_FROM_MOMENTS, 1.25, ops.make_noise(seed=77, subseq=12345, stream_id=_lib.STREAM_DRAW))
    assert torch.equal(a, b)


@pytest.mark.parametrize("n", [4, 4096, 16384 * 5 + 4, 7_000_004])
def test_ring_capture_tma_copy(cuda_device, n):
    from bayesdll_b200 import ops
    theta = torch.randn(n, device=cuda_device)
    ring = torch.full((3, n), -1.0, device=cuda_device)
    ops.capture_ring(theta, ring, 1)
    torch.cuda.synchronize()
    assert torch.equal(ring[1], theta)
    assert (ring[0] == -1).all() and (ring[2] == -1).all()    # neighbours untouched
    with pytest.raises(ops.BdlError):
        ops.capture_ring(theta, ring, 3)


@pytest.mark.parametrize("B,K,S", [(16, 37, 5), (64, 37, 40), (3, 10, 1), (128, 1000, 2), (1, 2, 7)])
def test_ensemble_matches_oracle_and_torch(cuda_device, B, K, S):
    from bayesdll_b200 import ops
    rng = np.random.default_rng(B * K * S)
    L = (rng.standard_normal((B, K, S)) * 3).astype(np.float32)
    y = rng.integers(0, K, B).astype(np.int64)
    Ld = torch.from_numpy(L).to(cuda_device)
    out = torch.empty(B, K, device=cuda_device)
    nst = S if S > 1 else 0
    ops.ensemble(Ld, out, nst)
    want = so.ensemble_average(L, nst)
    np.testing.assert_allclose(out.cpu().numpy(), want, atol=1e-5, rtol=1e-5)
    # the reference's own expression (methods/sgld.py:300), evaluated by torch on the same device
    ref = torch.log_softmax(Ld, 1).logsumexp(-1) - (np.log(nst) if nst else 0.0)
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), atol=1e-5, rtol=1e-5)
    # mixture accumulation in log space (methods/csgld.py:428-431)
    L2 = (rng.standard_normal((B, K, S)) * 3).astype(np.float32)
    mixt = torch.empty(B, K, device=cuda_device)
    ops.ensemble(Ld, mixt, nst, weight=0.3, mode=1)
    ops.ensemble(torch.from_numpy(L2).to(cuda_device), mixt, nst, weight=0.7, mode=2)
    want_mix = so.mixture([L, L2], [0.3, 0.7], nst) if nst else None
    if nst:
        np.testing.assert_allclose(mixt.cpu().numpy(), want_mix, atol=1e-5, rtol=1e-5)
    # CE + error count
    loss = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    err = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    ops.ce_err(out, torch.from_numpy(y).to(cuda_device), loss, err)
    ops.ce_err(out, torch.from_numpy(y).to(cuda_device), loss, err)      # accumulates
    ce = torch.nn.functional.cross_entropy(out, torch.from_numpy(y).to(cuda_device), reduction="sum").item()
    assert abs(loss.item() - 2 * ce) <= 2e-5 * max(1.0, abs(ce))
    assert err.item() == 2 * int((out.argmax(1).cpu().numpy() != y).sum())


def test_lse_path_equals_ensemble(cuda_device):
    """Sample-sharded formulation: running (max, sum) logsumexp over samples, split over two 'ranks', merged with
    MAX / rescale / SUM == the direct [B,K,S] reduction -- including logits whose probabilities underflow in fp32."""
    from bayesdll_b200 import ops
    B, K, S = 32, 37, 8
    L = torch.randn(B, K, S, device=cuda_device) * 2
    L[:4] *= 200.0                                            # softmax underflows to exactly 0 for most classes
    direct = torch.empty(B, K, device=cuda_device)
    ops.ensemble(L, direct, S)
    parts = []
    for samples in (range(0, S, 2), range(1, S, 2)):          # two ranks, round-robin
        m = torch.full((B, K), float("-inf"), device=cuda_device)
        s = torch.zeros(B, K, device=cuda_device)
        for i in samples:
            ops.lse_accum(L[:, :, i].contiguous(), m, s)
        parts.append((m, s))
    m_glob = torch.maximum(parts[0][0], parts[1][0])          # all-reduce(MAX)
    for m, s in parts:
        ops.lse_rescale(m, m_glob, s)
    s_glob = parts[0][1] + parts[1][1]                        # all-reduce(SUM)
    out = torch.empty(B, K, device=cuda_device)
    ops.lse_finalize(m_glob, s_glob, out, S)
    assert torch.isfinite(out).all() and torch.isfinite(direct).all()
    np.testing.assert_allclose(out.cpu().numpy(), direct.cpu().numpy(), atol=1e-4, rtol=2e-6)
    # a rank without samples contributes the neutral element
    m0 = torch.full((B, K), float("-inf"), device=cuda_device)
    s0 = torch.zeros(B, K, device=cuda_device)
    ops.lse_rescale(m0, m_glob, s0)
    assert (s0 == 0).all()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
@pytest.mark.parametrize("tname", ["T1", "Tarr"])
def test_calibration_matches_reference_golden(cuda_device, tag, tname):
    """Bin counts bit-exact, ECE/MCE/NLL within 1e-6 of the reference's calibration.analyze (north star iii)."""
    from bayesdll_b200 import calibration
    z = np.load(gu.golden_path("calibration"))
    logits, labels, M = z[f"{tag}_logits"], z[f"{tag}_labels"], int(z[f"{tag}_M"])
    T = 1 if tname == "T1" else z[f"{tag}_{tname}_T"]
    pre = f"{tag}_{tname}_"
    bins, binned, accs, confs, sizes = calibration.calc_bins(labels, logits, M, T)
    assert np.array_equal(bins, z[pre + "bins"])
    assert np.array_equal(sizes, z[pre + "sizes"]), "bin counts must be bit-exact"
    assert np.array_equal(binned, z[pre + "binned"])
    assert np.array_equal(accs, z[pre + "accs"])                      # integer count / integer count in fp64
    np.testing.assert_allclose(confs, z[pre + "confs"], rtol=1e-6)
    ece, mce, nll = calibration.analyze(labels, logits, M, None, T)
    assert calibration.last_near_edge == 0                            # certified: no probability within 16 ulp of an edge
    assert abs(ece - z[pre + "ece"]) <= 1e-6 and abs(mce - z[pre + "mce"]) <= 1e-6
    assert abs(nll - z[pre + "nll"]) <= 1e-6 * max(1.0, abs(z[pre + "nll"]))


def test_calibration_edge_cases(cuda_device):
    from bayesdll_b200 import calibration
    # one-hot certain predictions: p == 1.0 must land in the last bin thanks to the 1+1e-8 upper edge
    logits = np.full((8, 4), -200.0, np.float32)
    labels = np.arange(8) % 4
    logits[np.arange(8), labels] = 200.0
    bins, binned, accs, confs, sizes = calibration.calc_bins(labels, logits, 15)
    o = so.calc_bins(labels, logits, 15)
    assert np.array_equal(sizes, o[4]) and sizes[-1] == 8 and sizes[0] == 24
    assert np.array_equal(binned, o[1])
    # uniform predictions, K = 1 class, a single row
    for lg, lb in ((np.zeros((5, 10), np.float32), np.zeros(5, np.int64)), (np.zeros((3, 1), np.float32), np.zeros(3, np.int64)),
                   (np.array([[0.3, -1.2, 2.0]], np.float32), np.array([2]))):
        got = calibration.calc_bins(lb, lg, 15)
        want = so.calc_bins(lb, lg, 15)
        assert np.array_equal(got[4], want[4]) and np.array_equal(got[1], want[1])
        e1, m1, n1 = calibration.analyze(lb, lg, 15, None)
        e2, m2, n2 = so.analyze(lb, lg, 15)
        assert abs(e1 - e2) < 1e-6 and abs(m1 - m2) < 1e-6 and abs(n1 - n2) < 1e-6


@pytest.mark.parametrize("B,K,S", [(4, 4, 1), (16, 37, 4), (64, 37, 16), (3669, 37, 12), (1, 1, 3)])
def test_bma_mean_bit_exact(cuda_device, B, K, S):
    """bdl_bma_mean == the reference's `all_logits_sum += model_logits` ... `/ num_models` (csghmc_fs.py:349-377)."""
    from bayesdll_b200 import ops
    rng = np.random.default_rng(B * 131 + K * 7 + S)
    la = (rng.standard_normal((B, K, S)) * 20).astype(np.float32)
    out = torch.empty(B, K, dtype=torch.float32, device=cuda_device)
    ops.bma_mean(torch.from_numpy(la).to(cuda_device), out)
    want = so.bma_mean(la)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_temperature_objective_and_optimum(cuda_device, tag):
    """bdl_nll_temperature vs the reference's objective values, and the drop-in find_optimal_temperature vs the
    reference's own optimum (tests/golden/temperature.npz): NLL rel 1e-13, Topt rel 1e-5, same success flag."""
    from bayesdll_b200 import calibration, ops
    z = np.load(gu.golden_path("calibration"))
    t = np.load(gu.golden_path("temperature"))
    logits, labels = z[f"{tag}_logits"], z[f"{tag}_labels"]
    lg = torch.from_numpy(logits).to(cuda_device)
    lb = torch.from_numpy(labels).to(cuda_device)
    row = torch.empty(len(labels), dtype=torch.float64, device=cuda_device)
    out = torch.empty(1, dtype=torch.float64, device=cuda_device)
    for T, want in zip(t["Ts"], t[f"{tag}_fun"]):
        ops.nll_temperature(lg, lb, float(T), row, out)
        got = out.item()
        assert abs(got - want) <= 1e-13 * abs(want), (T, got, want)
        zz = logits.astype(np.float64) / T                     # per-row values too
        mx = zz.max(1)
        ref_rows = np.log(np.exp(zz - mx[:, None]).sum(1)) + mx - zz[np.arange(len(labels)), labels]
        np.testing.assert_allclose(row.cpu().numpy(), ref_rows, rtol=1e-12, atol=1e-13)
    # reproducible: same bits on a second evaluation (fixed reduction order)
    ops.nll_temperature(lg, lb, 1.37, row, out)
    a = out.item()
    ops.nll_temperature(lg, lb, 1.37, row, out)
    assert out.item() == a
    Topt, ok = calibration.find_optimal_temperature(labels, logits, None)
    assert bool(ok) == bool(t[f"{tag}_success"])
    assert isinstance(Topt, np.ndarray) and Topt.dtype == np.float64 and Topt.shape == (1,)
    np.testing.assert_allclose(Topt, t[f"{tag}_Topt"], rtol=1e-5)
    # and the calibrated metrics downstream agree with the reference's Topt
    e1 = calibration.analyze(labels, logits, 15, None, temperature=Topt)
    e2 = calibration.analyze(labels, logits, 15, None, temperature=t[f"{tag}_Topt"])
    assert all(abs(x - y) <= 1e-5 * max(1.0, abs(y)) for x, y in zip(e1, e2))


# ---- (8f row 4) per-step reparameterisation draws of the VI / MC-Dropout families --------------------------------
def _reparam_layout(z):
    from bayesdll_b200.flat import FlatLayout
    return FlatLayout([(nm, (int(k),)) for nm, k in zip(z["names"].tolist(), z["sizes"].tolist())], "classifier")


def test_vi_reparam_draw_matches_reference_golden(cuda_device):
    from bayesdll_b200 import ops
    z = np.load(gu.golden_path("reparam_draws"), allow_pickle=False)
    lay = _reparam_layout(z)
    d = lambda a: torch.from_numpy(lay.padded_numpy(a)).to(cuda_device)
    out = torch.empty(lay.n_padded, device=cuda_device)
    ops.draw(d(z["vi_m"]), d(z["vi_s"]), out, ops.STD_GIVEN, 1.0, ops.make_noise(xi=d(z["vi_eps"])))
    assert bits_equal(lay.dense_numpy(out.cpu().numpy()), z["vi_theta"])


@pytest.mark.parametrize("n", [4, 100_004, 3_000_000])
def test_vi_reparam_draw_bit_exact_and_philox_equivalent(cuda_device, n):
    from bayesdll_b200 import _lib, ops
    rng = np.random.default_rng(n + 5)
    m = rng.standard_normal(n).astype(np.float32)
    s_ = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    s_[::5] = rng.choice([-1.0, 0.0, 5e-9, 1e-8, 2e-8], size=s_[::5].shape).astype(np.float32)
    eps = rng.standard_normal(n).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(cuda_device)
    out = torch.empty(n, device=cuda_device)
    ops.draw(d(m), d(s_), out, ops.STD_GIVEN, 1.0, ops.make_noise(xi=d(eps)))
    assert bits_equal(out.cpu().numpy(), so.vi_sample(m, s_, eps))
    xi = torch.empty(n, device=cuda_device)
    ops.philox_normal(xi, 5, _lib.STREAM_DRAW, 99)
    a, b = torch.empty(n, device=cuda_device), torch.empty(n, device=cuda_device)
    ops.draw(d(m), d(s_), a, ops.STD_GIVEN, 1.0, ops.make_noise(xi=xi))
    ops.draw(d(m), d(s_), b, ops.STD_GIVEN, 1.0, ops.make_noise(seed=5, subseq=99, stream_id=_lib.STREAM_DRAW))
    assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["gaussian", "spikymix", "ignore"])
def test_mc_dropout_draw_matches_reference_golden(cuda_device, mode):
    from bayesdll_b200 import ops
    z = np.load(gu.golden_path("reparam_draws"), allow_pickle=False)
    lay = _reparam_layout(z)
    u_dense, _ = gu.mc_dropout_dense_inputs(z, mode)
    d = lambda a: torch.from_numpy(lay.padded_numpy(a)).to(cuda_device)
    runs_dev, nruns = ops.upload_runs(lay.dropout_run_table(mode), cuda_device)
    out, mask = torch.empty(lay.n_padded, device=cuda_device), torch.empty(lay.n_padded, device=cuda_device)
    ops.dropout_mix(d(z[f"mcd_{mode}_m"]), d(z[f"mcd_{mode}_theta0"]), out, float(z["p_drop"]), ops.make_noise(xi=d(u_dense)),
                    runs_dev, nruns, z_out=mask)
    assert bits_equal(lay.dense_numpy(out.cpu().numpy()), z[f"mcd_{mode}_theta"])
    out2 = torch.empty_like(out)                                                    # without the mask output
    ops.dropout_mix(d(z[f"mcd_{mode}_m"]), d(z[f"mcd_{mode}_theta0"]), out2, float(z["p_drop"]), ops.make_noise(xi=d(u_dense)),
                    runs_dev, nruns)
    assert torch.equal(out, out2)
    mk = mask.cpu().numpy()
    assert set(np.unique(mk).tolist()) <= {0.0, 1.0}
    for sg in lay.segments:
        if sg.is_bias and mode != "spikymix":
            assert np.all(mk[sg.begin:sg.begin + sg.numel] == 1.0)


@pytest.mark.parametrize("n", [4, 40_004, 4_000_000])
def test_mc_dropout_draw_bit_exact_and_philox_statistics(cuda_device, n):
    from bayesdll_b200 import ops
    from bayesdll_b200.flat import FlatLayout
    rng = np.random.default_rng(n + 9)
    k = n // 4
    lay = FlatLayout([("a.weight", (k,)), ("a.bias", (k,)), ("b.weight", (k,)), ("b.bias", (n - 3 * k,))], "b")
    n = lay.n_padded
    m, th0, u = (rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32),
                 rng.random(n).astype(np.float32))
    u[::11] = np.float32(0.25)
    d = lambda a: torch.from_numpy(a).to(cuda_device)
    for mode in ("gaussian", "spikymix"):
        nodrop = np.zeros(n, bool)
        for sg in lay.segments:
            nodrop[sg.begin:sg.end] = sg.is_bias and mode != "spikymix"
        runs_dev, nruns = ops.upload_runs(lay.dropout_run_table(mode), cuda_device)
        out, mask = torch.empty(n, device=cuda_device), torch.empty(n, device=cuda_device)
        ops.dropout_mix(d(m), d(th0), out, 0.25, ops.make_noise(xi=d(u)), runs_dev, nruns, z_out=mask)
        want, wz = so.mc_dropout_mix(m, th0, u, 0.25, nodrop)
        assert bits_equal(out.cpu().numpy(), want) and bits_equal(mask.cpu().numpy(), wz)
    # no run table = dropout everywhere
    ops.dropout_mix(d(m), d(th0), out, 0.25, ops.make_noise(xi=d(u)), z_out=mask)
    want, wz = so.mc_dropout_mix(m, th0, u, 0.25, np.zeros(n, bool))
    assert bits_equal(out.cpu().numpy(), want) and bits_equal(mask.cpu().numpy(), wz)
    # in-kernel Philox uniforms: deterministic per (seed, subseq), different across subseq, keep rate 1 - p_drop
    za, zb, zc = (torch.empty(n, device=cuda_device) for _ in range(3))
    ops.dropout_mix(d(m), d(th0), out, 0.25, ops.make_noise(seed=3, subseq=1), z_out=za)
    ops.dropout_mix(d(m), d(th0), out, 0.25, ops.make_noise(seed=3, subseq=1), z_out=zb)
    ops.dropout_mix(d(m), d(th0), out, 0.25, ops.make_noise(seed=3, subseq=2), z_out=zc)
    assert torch.equal(za, zb)
    assert torch.equal(out, torch.where(zc > 0, d(m) + 0.0 * d(th0), 0.0 * d(m) + d(th0)))
    if n >= 40_000:
        assert not torch.equal(za, zc)
        keep = za.double().mean().item()
        assert abs(keep - 0.75) < 5 * np.sqrt(0.75 * 0.25 / n)
        agree = (za == zc).double().mean().item()                                    # independent masks: 0.75^2 + 0.25^2
        assert abs(agree - 0.625) < 5 * np.sqrt(0.625 * 0.375 / n)
    # p_drop = 0 keeps everything (u > 0 fails only for u == 0: probability 2^-24 per element, handled like the reference)
    ops.dropout_mix(d(m), d(th0), out, 1.0, ops.make_noise(seed=3, subseq=1), z_out=za)
    assert za.sum().item() == 0.0


def test_dropout_mix_argument_errors(cuda_device):
    from bayesdll_b200 import ops
    from bayesdll_b200._lib import BdlError
    a = torch.zeros(8, device=cuda_device)
    with pytest.raises(BdlError):
        ops.dropout_mix(a, a.clone(), torch.empty(8, device=cuda_device), 1.5, ops.make_noise(seed=1))
    with pytest.raises(BdlError):
        ops.dropout_mix(a, torch.zeros(12, device=cuda_device), torch.empty(8, device=cuda_device), 0.5, ops.make_noise(seed=1))


def test_clamped_variance_sqrt_is_ieee_for_every_float(cuda_device):
    """bdl_draw (modes 0-2) clamps the variance at 1e-12 and takes an IEEE square root.  Sweep EVERY fp32 bit pattern from
    1e-12 up to and including +inf (1.41e9 values) plus a block of values below the clamp (zero, denormals, negatives)
    against torch's CUDA sqrt, which is the correctly rounded one: bit-identical.  (Guards any future replacement of the
    library square root, cf. the A/B in profiles/r01_ab_draw_sqrt.log.)"""
    from bayesdll_b200 import _lib, ops
    lo, hi = int(np.float32(1e-12).view(np.uint32)), 0x7F800000
    chunk = 1 << 26
    zeros, ones = torch.zeros(chunk, device=cuda_device), torch.ones(chunk, device=cuda_device)
    out = torch.empty(chunk, device=cuda_device)
    clamp = torch.tensor(1e-12, dtype=torch.float32, device=cuda_device)
    n_checked = 0
    for start in range(lo - 4096, hi + 1, chunk):
        stop = min(start + chunk, hi + 1)
        m = (stop - start + 3) // 4 * 4
        x = torch.arange(start, start + m, device=cuda_device, dtype=torch.int32).clamp_(max=hi).view(torch.float32)
        ops.draw(zeros[:m], x, out[:m], ops.VAR_FROM_WELFORD, 1.0, ops.make_noise(xi=ones[:m]), div_mode=_lib.DIV_IEEE)
        want = torch.sqrt(torch.maximum(x, clamp))
        assert torch.equal(out[:m], want), f"mismatch in bit range [{start:#x}, {stop:#x})"
        n_checked += stop - start
    assert n_checked == hi - lo + 1 + 4096
    low = torch.tensor([0.0, -0.0, 1e-45, 1e-38, -1.0, 9.9e-13, float("-inf"), 1e-12], device=cuda_device)
    ops.draw(zeros[:8], low, out[:8], ops.VAR_FROM_WELFORD, 1.0, ops.make_noise(xi=ones[:8]), div_mode=_lib.DIV_IEEE)
    assert torch.equal(out[:8], torch.sqrt(clamp).expand(8))
    # the moments path shares the function: ratio*(mom2 - mom1^2) with mom1 = 0, ratio = 1 is mom2 itself
    x = torch.tensor([1e-12, 3.0, 1e30, float("inf")], device=cuda_device)
    ops.draw(zeros[:4], x, out[:4], ops.VAR_FROM_MOMENTS, 1.0, ops.make_noise(xi=ones[:4]))
    assert torch.equal(out[:4], torch.sqrt(x))


@pytest.mark.parametrize("n", [64, 1_000_004])
def test_mc_dropout_in_kernel_philox_uniforms_equal_the_c_oracle(cuda_device, n):
    """The mask drawn from the in-kernel Philox stream is bit-for-bit the one the C restatement (oracle/bdl_oracle.c)
    derives from the same (seed, stream, subseq): integer arithmetic + one exact int->float conversion, no MUFU."""
    from bayesdll_b200 import _lib, ops
    from oracle import c_oracle as co
    rng = np.random.default_rng(n)
    m, th0 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(cuda_device)
    out, mask = torch.empty(n, device=cuda_device), torch.empty(n, device=cuda_device)
    for p_drop, sub in ((0.1, 5), (0.5, (1 << 40) + 9), (0.97, 0)):
        ops.dropout_mix(d(m), d(th0), out, p_drop, ops.make_noise(seed=(3 << 33) + 11, subseq=sub, stream_id=_lib.STREAM_DRAW), z_out=mask)
        nz = _lib.Noise()
        nz.xi_dev, nz.seed, nz.subseq, nz.stream_id = 0, (3 << 33) + 11, sub, _lib.STREAM_DRAW
        want, wz = np.empty(n, np.float32), np.empty(n, np.float32)
        co.dropout_mix(m, th0, want, p_drop, nz, z_out=wz)
        assert np.array_equal(mask.cpu().numpy(), wz) and bits_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("threads", [64, 128, 256])
def test_probe_stream_moves_what_it_says(cuda_device, threads):
    """bdl_probe_stream (the bare-traffic yardstick bench.py times next to the real kernels): every element is read and
    written exactly once per stream, for a length that is not a multiple of the CTA size; neighbours stay untouched."""
    from bayesdll_b200 import ops
    n = 4 * (3 * 256 + 37)
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    pad = 64
    def buf():
        return torch.randn(n + 2 * pad, device=cuda_device, generator=gen)
    A, B, Cc, D = buf(), buf(), buf(), buf()
    a, b, c, d = (t[pad:pad + n] for t in (A, B, Cc, D))
    A0, B0 = A.clone(), B.clone()
    ops.probe_stream(a, None, c, None, 1, 1, threads=threads)
    assert torch.equal(a, c) and torch.equal(A[:pad], A0[:pad]) and torch.equal(A[pad + n:], A0[pad + n:])
    ops.probe_stream(a, None, c, d, 2, 1, threads=threads)
    assert torch.equal(a, c + d)
    a_old, b_old = a.clone(), b.clone()
    ops.probe_stream(a, b, c, d, 4, 2, threads=threads)
    want_a = 0.5 * ((c + d) + a_old)
    assert torch.equal(a, want_a) and torch.equal(b, 0.5 * (want_a - b_old))
    assert torch.equal(B[:pad], B0[:pad]) and torch.equal(B[pad + n:], B0[pad + n:])
    with pytest.raises(Exception, match="reads, writes"):
        ops.probe_stream(a, b, c, d, 3, 1, threads=threads)
