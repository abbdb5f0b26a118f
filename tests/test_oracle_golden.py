"""CPU: the oracle (oracle/sampler_oracle.py) against golden vectors recorded from the real reference
(oracle/make_golden.py).  This is the pin that makes the oracle trustworthy."""
import numpy as np
import pytest

import golden_util as gu
from oracle import sampler_oracle as so

# SGLD / cSGLD / SGHMC / cSGHMC: every op of the reference is a correctly-rounded fp32 op -> bit exact.
# Adam variants: torch's CPU sqrt (AVX-512 build, 2.11.0) is not correctly rounded in ~0.7% of inputs, so the
# reference itself deviates by 1 ulp there; m and s (no sqrt) are still bit exact.
EXACT = ("sgld", "csgld", "sghmc", "csghmc")


def _bits_equal(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("name", gu.step_cases())
@pytest.mark.parametrize("chained", [False, True])
def test_step_oracle_matches_reference(name, chained):
    _, _, method = gu.load_step_case(name)
    pairs = gu.replay_step_case(name, so, chained=chained, div_mode="true")
    assert pairs["theta"], "no steps replayed"
    for key, lst in pairs.items():
        for t, (got, want) in enumerate(lst):
            if method in EXACT or key in ("m", "s"):
                assert _bits_equal(got, want), f"{name} {key} step {t}: not bit exact"
            else:
                assert gu.max_rel(got, want) <= 1e-6, f"{name} {key} step {t}: rel {gu.max_rel(got, want):.2e}"


def test_recip_mode_is_within_tolerance_of_reference():
    """CUDA-eager semantics (multiply by fp32 reciprocal) stay inside the north-star 1e-6 per-step bound."""
    for name in gu.step_cases():
        pairs = gu.replay_step_case(name, so, chained=False, div_mode="recip")
        for key, lst in pairs.items():
            for got, want in lst:
                assert gu.max_rel(got, want) <= 1e-6


CLIP_CASES = [n for n in gu.step_cases() if "_clip_" in n]


@pytest.mark.parametrize("name", CLIP_CASES)
def test_clip_norm_and_coefficient_match_reference(name):
    """args.clip_grad (methods/csgld.py:250-251, adam_csghmc.py:319-320): the oracle's total norm of the recorded p.grad
    agrees with torch.nn.utils.clip_grad_norm_'s return value to fp32 rel 1e-6 (only the summation order differs), both
    regimes occur (coef < 1 and coef clamped to 1), and the oracle recomputes that p.grad itself bit for bit."""
    z, hp, method = gu.load_step_case(name)
    assert len(CLIP_CASES) >= 3
    clip = float(z["clip_grad"])
    coefs = []
    for t in range(z["G"].shape[0]):
        coef, total = so.clip_coef(z["pgrad"][t], gu.valid_mask(z), clip)
        assert abs(float(total) - float(z["total_norm"][t])) <= 1e-6 * float(z["total_norm"][t])
        want = gu.recorded_clip_coef(z, t)
        assert abs(float(coef) - float(want)) <= 1e-6 * float(want)
        coefs.append(float(want))
    assert min(coefs) < 1.0 and max(coefs) == 1.0, coefs
    # with its OWN coefficient (not the recorded one) the oracle stays within the per-step tolerance of the reference
    H = gu.hparams_from(hp, z, method)
    P = gu.prior_mask(hp, z)
    zeros = np.zeros(z["G"].shape[1], np.float32)
    for t in range(1, z["G"].shape[0]):
        kw = dict(is_head=z["is_head"], lr_body=float(z["lr_body"][t]), lr_head=float(z["lr_head"][t]), hp=H, clip=clip,
                  valid=gu.valid_mask(z))
        if method == "csgld":
            th, _ = so.step_sgld(z["theta"][t - 1], z["G"][t], z["theta0"], z["buf"][t - 1] if H.mu else zeros, z["XI"][t], P=P,
                                 first_step=False, **kw)
        else:
            th, _, _, _ = so.step_adam_csghmc(z["theta"][t - 1], z["G"][t], z["theta0"], z["v"][t - 1], z["m"][t - 1], z["s"][t - 1],
                                              z["XI"][t], P=P, t=t + 1, **kw)
        assert gu.max_rel(th, z["theta"][t]) <= 1e-6


def test_cyclical_schedule():
    z = np.load(gu.golden_path("cyclical"))
    rows = z["rows"]
    assert len(rows) > 500
    for (epochs, B, M, beta, lr0, ep, b, lr, ss, lic, cyc) in rows:
        s = so.CyclicalOracle(lr0, int(M), int(epochs), beta)
        kw = dict(epoch=int(ep), batch=int(b), B=int(B))
        assert s.calculate_lr(**kw) == lr
        assert float(s.should_sample(**kw)) == ss
        assert float(s.last_in_cycle(**kw)) == lic
        assert s.get_cycle_number(**kw) == cyc


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
@pytest.mark.parametrize("tname", ["T1", "Tarr"])
def test_calibration_oracle(tag, tname):
    z = np.load(gu.golden_path("calibration"))
    logits, labels, M = z[f"{tag}_logits"], z[f"{tag}_labels"], int(z[f"{tag}_M"])
    T = 1 if tname == "T1" else z[f"{tag}_{tname}_T"]
    edges, binned, accs, confs, sizes = so.calc_bins(labels, logits, M, T)
    pre = f"{tag}_{tname}_"
    assert np.array_equal(edges, z[pre + "bins"])
    assert np.array_equal(binned, z[pre + "binned"])
    assert np.array_equal(sizes, z[pre + "sizes"])
    assert np.array_equal(accs, z[pre + "accs"])
    assert np.array_equal(confs, z[pre + "confs"])
    ece, mce, nll = so.analyze(labels, logits, M, T)
    assert ece == z[pre + "ece"] and mce == z[pre + "mce"]
    assert abs(nll - z[pre + "nll"]) <= 1e-6 * abs(z[pre + "nll"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_temperature_oracle(tag):
    """Oracle objective and driver vs the reference's own find_optimal_temperature (tests/golden/temperature.npz)."""
    z = np.load(gu.golden_path("calibration"))
    t = np.load(gu.golden_path("temperature"))
    logits, labels = z[f"{tag}_logits"], z[f"{tag}_labels"]
    for T, want in zip(t["Ts"], t[f"{tag}_fun"]):
        got = so.nll_temperature(labels, logits, np.array([T]))
        assert abs(got - want) <= 4e-15 * abs(want)          # scipy's logsumexp vs the explicit form: last-ulp only
    Topt, ok = so.find_optimal_temperature(labels, logits)
    assert bool(ok) == bool(t[f"{tag}_success"])
    np.testing.assert_allclose(Topt, t[f"{tag}_Topt"], rtol=1e-6)


def test_bma_mean_oracle_order():
    """fp32 running sum in model order then one division: not the same as a pairwise / fp64 mean."""
    rng = np.random.default_rng(3)
    la = (rng.standard_normal((50, 7, 9)) * 30).astype(np.float32)
    got = so.bma_mean(la)
    acc = np.zeros((50, 7), np.float32)
    for m in range(9):
        acc = (acc + la[:, :, m]).astype(np.float32)
    assert np.array_equal(got, (acc / np.float32(9)).astype(np.float32))
    assert got.dtype == np.float32


def _reparam_golden():
    return np.load(gu.golden_path("reparam_draws"), allow_pickle=False)


def test_vi_reparam_draw_oracle():
    """(8f row 4) methods/vi.py:402-406 recorded from the reference's own vi.Model.forward."""
    z = _reparam_golden()
    assert (z["vi_s"] < 1e-8).any() and (z["vi_s"] > 1e-8).any()         # the clamp is exercised on both sides
    assert _bits_equal(so.vi_sample(z["vi_m"], z["vi_s"], z["vi_eps"]), z["vi_theta"])


@pytest.mark.parametrize("mode", ["gaussian", "spikymix", "ignore"])
def test_mc_dropout_draw_oracle(mode):
    """(8f row 4) methods/mc_dropout.py:378-394 recorded from the reference's own mc_dropout.Model.forward.  The
    reference consumes uniforms only for tensors it drops (bias tensors draw none in 'gaussian' / 'ignore')."""
    z = _reparam_golden()
    u_dense, nodrop = gu.mc_dropout_dense_inputs(z, mode)
    theta, mask = so.mc_dropout_mix(z[f"mcd_{mode}_m"], z[f"mcd_{mode}_theta0"], u_dense, z["p_drop"], nodrop)
    assert _bits_equal(theta, z[f"mcd_{mode}_theta"])
    assert 0 < (mask == 0).sum() < mask.size
    assert np.all(mask[(u_dense == z["p_drop"]) & ~nodrop] == 0)         # u == p_drop is dropped ('>' is strict)
