"""CPU: the C restatement (oracle/bdl_oracle.c) against Random123 known-answer vectors, against the numpy
oracle (bit-exact) and hence -- transitively -- against the reference goldens."""
import ctypes as C

import numpy as np
import pytest

import golden_util as gu
from bayesdll_b200 import _lib as L
from bayesdll_b200.flat import FlatLayout
from oracle import c_oracle as co
from oracle import sampler_oracle as so


def test_philox_known_answers():
    # Random123 v1.14 examples/kat_vectors, "philox4x32 10"
    assert co.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert co.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert co.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_philox_normal_moments_and_counter_layout():
    n = 1 << 20
    x = co.philox_normal(n, seed=5, stream_id=0, subseq=9).astype(np.float64)
    assert abs(x.mean()) < 5 / np.sqrt(n) and abs(x.var() - 1) < 5 * np.sqrt(2 / n)
    # element 4q+k depends only on (seed, stream, subseq, q): a shorter fill is a prefix of a longer one
    assert np.array_equal(co.philox_normal(1024, 5, 0, 9), x[:1024].astype(np.float32))
    assert not np.array_equal(co.philox_normal(1024, 5, 0, 10), x[:1024].astype(np.float32))
    assert not np.array_equal(co.philox_normal(1024, 5, 1, 9), x[:1024].astype(np.float32))


class COracleStepper:
    """oracle-API adapter over bdl_oracle_step (padded flat host buffers + run table), mirrors tests/gpu_impl.py."""

    def __init__(self, names, sizes, readout, bias_mode):
        self.layout = FlatLayout([(n, (int(s),)) for n, s in zip(names, sizes)], readout)
        self.runs = self.layout.run_table(bias_mode)

    def _run(self, variant, sc, coef=None, **arrs):
        T = {k: (None if a is None else np.ascontiguousarray(self.layout.padded_numpy(np.asarray(a, np.float32))))
             for k, a in arrs.items()}
        nz = L.Noise()
        nz.xi_dev = T["xi"].ctypes.data
        if coef is not None:      # args.clip_grad goldens: the step with the recorded coefficient
            co.step_clipped(variant, T["theta"], T["g"], T["theta0"], T["v"], T["m"], T["s"], T["buf"], self.runs, sc, nz, coef)
        else:
            co.step(variant, T["theta"], T["g"], T["theta0"], T["v"], T["m"], T["s"], T["buf"], self.runs, sc, nz)
        return {k: (None if t is None else self.layout.dense_numpy(t)) for k, t in T.items()}

    @staticmethod
    def _sc(variant, hp, lrb, lrh, div_mode, **kw):
        from bayesdll_b200 import ops
        return ops.make_scalars(variant, lr_body=lrb, lr_head=lrh, ND=hp.ND, Ninflate=hp.Ninflate, prior_sig=hp.prior_sig,
                                nd=hp.nd, alpha=hp.alpha, mu=hp.mu, beta1=hp.beta1, beta2=hp.beta2, eps=hp.eps,
                                temperature=hp.temperature, div_mode={"true": 0, "recip": 1}[div_mode], **kw)

    def step_sgld(self, theta, g, theta0, buf, xi, *, is_head, P, lr_body, lr_head, hp, first_step, div_mode="true", coef=None):
        sc = self._sc(L.SGLD, hp, lr_body, lr_head, div_mode, first_step=first_step)
        o = self._run(L.SGLD, sc, coef=coef, theta=theta, g=g, theta0=theta0, v=None, m=None, s=None,
                      buf=buf if hp.mu != 0 else None, xi=xi)
        return o["theta"], (o["buf"] if hp.mu != 0 else buf)

    def step_sghmc(self, theta, g, theta0, v, xi, *, is_head, P, lr_body, lr_head, hp, div_mode="true"):
        o = self._run(L.SGHMC, self._sc(L.SGHMC, hp, lr_body, lr_head, div_mode), theta=theta, g=g, theta0=theta0, v=v,
                      m=None, s=None, buf=None, xi=xi)
        return o["theta"], o["v"]

    def step_csghmc(self, theta, g, v, xi, *, is_head, lr_body, lr_head, hp, should_sample):
        o = self._run(L.CSGHMC, self._sc(L.CSGHMC, hp, lr_body, lr_head, "true", add_noise=should_sample), theta=theta,
                      g=g, theta0=None, v=v, m=None, s=None, buf=None, xi=xi)
        return o["theta"], o["v"]

    def step_adam_sghmc(self, theta, g, theta0, v, m, s, buf, xi, *, is_head, P, lr_body, lr_head, hp, t, first_step,
                        div_mode="true"):
        sc = self._sc(L.ADAM_SGHMC, hp, lr_body, lr_head, div_mode, t=t, first_step=first_step)
        o = self._run(L.ADAM_SGHMC, sc, theta=theta, g=g, theta0=theta0, v=v, m=m, s=s,
                      buf=buf if hp.mu != 0 else None, xi=xi)
        return o["theta"], o["v"], o["m"], o["s"], (o["buf"] if hp.mu != 0 else buf)

    def step_adam_csghmc(self, theta, g, theta0, v, m, s, xi, *, is_head, P, lr_body, lr_head, hp, t, div_mode="true", coef=None):
        sc = self._sc(L.ADAM_CSGHMC, hp, lr_body, lr_head, div_mode, t=t)
        o = self._run(L.ADAM_CSGHMC, sc, coef=coef, theta=theta, g=g, theta0=theta0, v=v, m=m, s=s, buf=None, xi=xi)
        return o["theta"], o["v"], o["m"], o["s"]


@pytest.mark.parametrize("name", gu.step_cases())
@pytest.mark.parametrize("div_mode", ["true", "recip"])
def test_c_oracle_bit_exact_vs_numpy_oracle(name, div_mode):
    z, hp, _ = gu.load_step_case(name)
    cimpl = COracleStepper(z["names"].tolist(), z["sizes"].tolist(), "classifier", hp["bias"])
    a = gu.replay_step_case(name, cimpl, chained=True, div_mode=div_mode)
    b = gu.replay_step_case(name, so, chained=True, div_mode=div_mode)
    for key in a:
        for (got, _), (want, _) in zip(a[key], b[key]):
            assert np.array_equal(got.view(np.uint32), np.asarray(want, np.float32).view(np.uint32)), (name, key)


@pytest.mark.parametrize("name", [n for n in gu.step_cases() if "_clip_" in n])
def test_c_oracle_gradnorm_matches_reference_norm(name):
    """The C restatement's norm pass (bdl_oracle_step_gradnorm, per-tensor runs) reproduces the total norm the reference's
    clip_grad_norm_ returned at every recorded step to fp32 rel 1e-6, and its coefficient is the recorded one."""
    z, hp, method = gu.load_step_case(name)
    H = gu.hparams_from(hp, z, method)
    lay = FlatLayout([(n, (int(s),)) for n, s in zip(z["names"].tolist(), z["sizes"].tolist())], "classifier")
    tab = lay.run_table(hp["bias"], grad_ptrs=[0] * len(lay.segments))
    variant = L.SGLD if method == "csgld" else L.ADAM_CSGHMC
    pad = lambda a: np.ascontiguousarray(lay.padded_numpy(np.asarray(a, np.float32)))
    zeros = np.zeros(z["G"].shape[1], np.float32)
    for t in range(z["G"].shape[0]):
        prev = (lambda k: z[k][t - 1] if t > 0 and k in z.files else (z["theta_init"] if k == "theta" else zeros))
        sc = COracleStepper._sc(variant, H, float(z["lr_body"][t]), float(z["lr_head"][t]), "true", t=t + 1, first_step=(t == 0))
        nz = L.Noise()
        xi = pad(z["XI"][t])
        nz.xi_dev = xi.ctypes.data
        adam = variant == L.ADAM_CSGHMC
        sumsq = co.step_gradnorm(variant, pad(prev("theta")), pad(z["G"][t]), pad(z["theta0"]), pad(prev("v")) if adam else None,
                                 pad(prev("m")) if adam else None, pad(prev("s")) if adam else None,
                                 pad(prev("buf")) if H.mu else None, tab, sc, nz)
        coef, total = co.clip_coef(sumsq, float(z["clip_grad"]))
        assert abs(total - float(z["total_norm"][t])) <= 1e-6 * float(z["total_norm"][t]), (t, total, float(z["total_norm"][t]))
        assert abs(coef - float(gu.recorded_clip_coef(z, t))) <= 1e-6


def test_c_oracle_moments_and_draw_vs_numpy():
    rng = np.random.default_rng(0)
    n = 4096
    th = rng.standard_normal(n).astype(np.float32)
    for div_name, div in (("true", 0), ("recip", 1)):
        m1, m2 = so.moments_init(th)
        c1, c2 = np.empty(n, np.float32), np.empty(n, np.float32)
        co.moments_avg(th, c1, c2, 0, 1, div)
        assert np.array_equal(c1, m1) and np.array_equal(c2, m2)
        for cnt in range(1, 5):
            th2 = rng.standard_normal(n).astype(np.float32)
            m1, m2 = so.moments_avg(th2, m1, m2, cnt, div_name)
            co.moments_avg(th2, c1, c2, cnt, 0, div)
            assert np.array_equal(c1, m1) and np.array_equal(c2, m2)
        mean, M2 = th.copy(), np.zeros(n, np.float32)
        cm, cM = np.empty(n, np.float32), np.empty(n, np.float32)
        co.moments_welford(th, cm, cM, 1, 1, div)
        for k in (3, 5, 7):
            th2 = rng.standard_normal(n).astype(np.float32)
            mean, M2 = so.moments_welford(th2, mean, M2, k, div_name)
            co.moments_welford(th2, cm, cM, k, 0, div)
            assert np.array_equal(cm, mean) and np.array_equal(cM, M2)
        eps = rng.standard_normal(n).astype(np.float32)
        nz = L.Noise()
        nz.xi_dev = eps.ctypes.data
        out = np.empty(n, np.float32)
        co.draw(m1, m2, out, 0, 5 / 4, div, nz)
        assert np.array_equal(out, so.posterior_draw(m1, so.variance_from_moments(m1, m2, 5 / 4), eps))
        co.draw(mean, M2, out, 1, 6.0, div, nz)
        assert np.array_equal(out, so.posterior_draw(mean, so.variance_from_welford(M2, 7, div_name), eps))
        s_ = (M2 * np.float32(1e-8) - np.float32(5e-9)).astype(np.float32)          # straddles the 1e-8 clamp
        co.draw(mean, s_, out, 4, 1.0, div, nz)
        assert np.array_equal(out, so.vi_sample(mean, s_, eps))


def test_c_oracle_dropout_mix_vs_numpy_and_reference_golden():
    """bdl_oracle_dropout_mix (C) == sampler_oracle.mc_dropout_mix (numpy) on injected uniforms, incl. the recordings of
    the reference's own mc_dropout.Model.forward; its Philox uniforms lie in [0, 1) with 24-bit resolution."""
    import golden_util as gu
    from bayesdll_b200.flat import FlatLayout
    z = np.load(gu.golden_path("reparam_draws"), allow_pickle=False)
    lay = FlatLayout([(nm, (int(k),)) for nm, k in zip(z["names"].tolist(), z["sizes"].tolist())], "classifier")
    for mode in ("gaussian", "spikymix", "ignore"):
        u_dense, _ = gu.mc_dropout_dense_inputs(z, mode)
        m, th0, u = (lay.padded_numpy(z[f"mcd_{mode}_m"]), lay.padded_numpy(z[f"mcd_{mode}_theta0"]), lay.padded_numpy(u_dense))
        nz = L.Noise()
        nz.xi_dev = u.ctypes.data
        out, mask = np.empty_like(m), np.empty_like(m)
        co.dropout_mix(m, th0, out, float(z["p_drop"]), nz, runs=lay.dropout_run_table(mode), z_out=mask)
        assert np.array_equal(lay.dense_numpy(out).view(np.uint32), z[f"mcd_{mode}_theta"].view(np.uint32))
        nodrop = np.zeros(lay.n_padded, bool)
        for sg in lay.segments:
            nodrop[sg.begin:sg.end] = sg.is_bias and mode != "spikymix"
        want, wz = so.mc_dropout_mix(m, th0, u, z["p_drop"], nodrop)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)) and np.array_equal(mask, wz)
    # in-oracle Philox uniforms: reproducible, 24-bit grid in [0, 1), keep rate 1 - p_drop
    n = 1 << 18
    m, th0 = np.ones(n, np.float32), np.zeros(n, np.float32)
    nz = L.Noise()
    nz.xi_dev, nz.seed, nz.subseq, nz.stream_id = 0, 7, 3, L.STREAM_DRAW
    a, b = np.empty(n, np.float32), np.empty(n, np.float32)
    co.dropout_mix(m, th0, a, 0.25, nz)
    co.dropout_mix(m, th0, b, 0.25, nz)
    assert np.array_equal(a, b) and set(np.unique(a).tolist()) == {0.0, 1.0}
    assert abs(a.mean() - 0.75) < 5 * np.sqrt(0.75 * 0.25 / n)
    ctr = [5, L.STREAM_DRAW, 3, 0]
    r = co.philox4x32_10(ctr, [7, 0])
    co.dropout_mix(m, th0, a, 0.5, nz)
    assert a[20:24].tolist() == [1.0 if (x >> 8) * 2.0 ** -24 > 0.5 else 0.0 for x in r]
