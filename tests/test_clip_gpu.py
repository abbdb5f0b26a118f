"""GPU: args.clip_grad -- ``torch.nn.utils.clip_grad_norm_(net.parameters(), clip)`` between ``Model.forward`` and
``optimizer.step()`` (methods/csgld.py:250-251, methods/adam_csghmc.py:319-320), fused as bdl_step_gradnorm ->
bdl_clip_coef -> bdl_step_clipped.

The update itself is pinned bit for bit by the reference goldens (tests/golden/step_*_clip_*.npz, replayed with the
recorded coefficient in test_step_gpu.py); here: the norm pass against the oracle (fp64 sum of squares: rel 1e-12;
coefficient: the bits torch's statements give for that norm), NaN-poisoned gradient tails, a tensor without gradient, and
the whole three-launch sequence through ChainState.update against the reference's statements + the REAL
clip_grad_norm_ + torch.optim.SGD in torch CUDA eager (north-star tolerance fp32 rel 1e-6 per step)."""
import numpy as np
import pytest
import torch

import eager_reference as er
from oracle import sampler_oracle as so
from test_cuda_eager_parity_gpu import HP, READOUT, _Net

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant_name,mu", [("sgld", 0.5), ("sgld", 0.0), ("adam_csghmc", 0.0)])
@pytest.mark.parametrize("philox", [False, True])
def test_gradnorm_coef_and_clipped_step_vs_oracle(cuda_device, variant_name, mu, philox):
    from bayesdll_b200 import _lib, ops
    from test_step_gpu import _random_layout, bits_equal
    dev = cuda_device
    rng = np.random.default_rng(abs(hash((variant_name, mu, philox))) % 2**32)
    lay = _random_layout(rng, 31, 3000)
    n = lay.n_padded
    variant = dict(sgld=_lib.SGLD, adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    adam = variant_name == "adam_csghmc"
    hp = so.HParams(ND=1840, Ninflate=3.0, prior_sig=0.9, nd=0.7, alpha=0.18, beta1=0.9, beta2=0.999, eps=1e-8,
                    temperature=1.3, mu=mu)
    is_head, P = lay.per_element("uninformative")
    f = lambda scale=1.0: (rng.standard_normal(n) * scale).astype(np.float32)
    theta, theta0, v, m, buf = f(0.1), f(0.1), f(0.01), f(0.01), f(0.01)
    s = np.abs(f(1e-3)).astype(np.float32) + np.float32(1e-6)
    g = np.zeros(n, np.float32)
    keep, ptrs = [], []
    skip_idx = 7
    valid = np.zeros(n, bool)
    for k, sg in enumerate(lay.segments):
        gd = rng.standard_normal(sg.numel).astype(np.float32) * 0.05
        g[sg.begin:sg.begin + sg.numel] = gd
        t = torch.full((sg.numel + 8,), float("nan"), device=dev)           # poison what follows the tensor's end
        t[:sg.numel] = torch.from_numpy(gd).to(dev)
        keep.append(t)
        ptrs.append(t.data_ptr())
        if k != skip_idx:
            valid[sg.begin:sg.begin + sg.numel] = True                       # padding and gradient-less tensors: not in the norm
    tab = lay.run_table("uninformative", grad_ptrs=ptrs)
    tab[skip_idx].cls |= _lib.CLS_SKIP
    tab[skip_idx].g_dev = 0
    T = {k: torch.from_numpy(a.copy()).to(dev) for k, a in dict(theta=theta, theta0=theta0, v=v, m=m, s=s, buf=buf).items()}
    xi = torch.empty(n, device=dev)
    ops.philox_normal(xi, 5, _lib.STREAM_STEP, 3)
    xi_h = xi.cpu().numpy()
    nz = ops.make_noise(seed=5, subseq=3) if philox else ops.make_noise(xi=xi)
    sc = ops.make_scalars(variant, lr_body=1e-3, lr_head=1e-2, ND=hp.ND, Ninflate=hp.Ninflate, prior_sig=hp.prior_sig, nd=hp.nd,
                          alpha=hp.alpha, mu=mu, beta1=hp.beta1, beta2=hp.beta2, eps=hp.eps, temperature=hp.temperature, t=3,
                          first_step=False, div_mode=_lib.DIV_IEEE)
    state = (variant, T["theta"], None, T["theta0"], T["v"] if adam else None, T["m"] if adam else None, T["s"] if adam else None,
             T["buf"] if mu else None, tab, len(tab), sc, nz)
    # --- oracle: p.grad as the reference holds it before clip_grad_norm_ -------------------------------------------
    kw = dict(is_head=is_head, lr_body=1e-3, lr_head=1e-2, hp=hp)
    if adam:
        _, pgrad, _, _ = so._adam_core(theta, g, theta0, v, m, s, xi_h, P=P, t=3, cyc=True, div_mode="true", **kw)
    else:
        c = so._per_class(is_head, hp.nd * np.sqrt(2 / (hp.N * 1e-3)), hp.nd * np.sqrt(2 / (hp.N * 1e-2))) * xi_h
        pgrad = g + np.where(P.astype(bool), so._prior_term(theta, theta0, hp, "true") + c, c)
    want_sumsq = float(np.sum(np.asarray(pgrad, np.float64)[valid] ** 2))
    for clip in (0.5 * np.sqrt(want_sumsq), 2.0 * np.sqrt(want_sumsq)):      # clipping active / coefficient clamped to 1
        before = {k: t.clone() for k, t in T.items()}
        sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        out = torch.zeros(2, device=dev)
        ops.step_gradnorm(*state, sumsq)
        assert all(torch.equal(T[k], before[k]) for k in T), "the norm pass must not write the state"
        ops.clip_coef(sumsq, clip, out[0:1], out[1:2])
        got_sumsq, coef, total = sumsq.item(), np.float32(out[0].item()), np.float32(out[1].item())
        assert abs(got_sumsq - want_sumsq) <= 1e-12 * want_sumsq
        assert total == np.float32(np.sqrt(got_sumsq))
        want_coef = np.float32(min((np.float32(1.0) / (total + np.float32(1e-6))) * np.float32(clip), np.float32(1.0)))
        assert coef == want_coef and (coef < 1.0) == (clip < np.sqrt(want_sumsq))
        ops.step_clipped(*state, out[0:1])
        torch.cuda.synchronize()
        if adam:
            want = dict(zip(("theta", "v", "m", "s"), so.step_adam_csghmc(theta, g, theta0, v, m, s, xi_h, P=P, t=3, coef=coef, **kw)))
        else:
            want = dict(zip(("theta", "buf"), so.step_sgld(theta, g, theta0, buf, xi_h, P=P, first_step=False, coef=coef, **kw)))
            if not mu:
                want.pop("buf")
        seg = lay.segments[skip_idx]
        for k, w in want.items():
            got = T[k].cpu().numpy()
            assert not np.isnan(got).any(), f"{k}: NaN leaked from the gradient padding"
            w = w.copy()
            w[seg.begin:seg.end] = dict(theta=theta, v=v, m=m, s=s, buf=buf)[k][seg.begin:seg.end]    # p.grad None: untouched
            assert bits_equal(got, w), f"{variant_name} {k}: {(got.view(np.uint32) != w.view(np.uint32)).sum()} mismatches"
        for k, a in dict(theta=theta, v=v, m=m, s=s, buf=buf).items():
            T[k].copy_(torch.from_numpy(a))


@pytest.mark.parametrize("variant_name", ["sgld", "adam_csghmc"])
def test_chain_update_with_clip_equals_reference_statements(cuda_device, variant_name):
    """ChainState.update(clip=...) -- what the cyclical runners call when args.clip_grad is set -- against the reference's
    update statements + torch.nn.utils.clip_grad_norm_ + torch.optim.SGD.step in torch CUDA eager, three chained steps,
    the clip threshold chosen so that step 1 clips and a later one does not."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.chain import ChainState
    dev = cuda_device
    gen = torch.Generator().manual_seed(23)
    mu = 0.5 if variant_name == "sgld" else 0.0
    variant = dict(sgld=_lib.SGLD, adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    N = HP["ND"] * HP["Ninflate"]
    ref, ref0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net, net0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net.load_state_dict(ref.state_dict())
    net0.load_state_dict(ref0.state_dict())
    chain = ChainState(net, net0, variant=variant, bias_mode="informative", mu=mu, noise="torch", seed=0)
    named = list(ref.named_parameters())
    names = [n for n, _ in named]
    p0s = [p for _, p in ref0.named_parameters()]
    opt = er.make_sgd([p for n, p in named if READOUT not in n], [p for n, p in named if READOUT in n], HP["lr_body"], HP["lr_head"], mu)
    vs, ms, ss = ({n: torch.zeros_like(p) for n, p in named} for _ in range(3))
    coefs = []
    clip = None
    for t in range(1, 4):
        scale = 0.05 if t < 3 else 0.005                     # smaller gradients at step 3: the norm drops below the threshold
        grads = [(torch.randn(p.shape, generator=gen) * scale).to(dev) for _, p in named]
        xis = [(torch.randn(p.shape, generator=gen) * (1.0 if t < 3 else 0.1)).to(dev) for _, p in named]
        for (_, p), g in zip(named, grads):
            p.grad = g.clone()
        kw = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], N=N, prior_sig=HP["prior_sig"], nd=HP["nd"])
        if variant_name == "sgld":
            er.sgld(named, p0s, xis, READOUT, bias="informative", **kw)
        else:
            er.adam(named, p0s, xis, vs, ms, ss, READOUT, alpha=HP["alpha"], beta1=HP["beta1"], beta2=HP["beta2"], eps=HP["eps"],
                    t=t, bias="informative", cyclical=True, temperature=HP["temperature"], **kw)
        if clip is None:
            clip = 0.6 * float(torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad) for _, p in named])))
        total = torch.nn.utils.clip_grad_norm_([p for _, p in named], clip)          # the reference's statement
        coefs.append(min(1.0, clip / (float(total) + 1e-6)))
        opt.step()
        # product
        for p, g in zip(chain.params, grads):
            p.grad = g.clone()
        import bayesdll_b200.chain as chain_mod
        tape = iter(xis)
        orig = torch.randn_like
        torch.randn_like = lambda like, **k: next(tape).reshape(like.shape)
        try:
            sc = ops.make_scalars(variant, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                                  prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=HP["alpha"], mu=mu, beta1=HP["beta1"],
                                  beta2=HP["beta2"], eps=HP["eps"], temperature=HP["temperature"], t=t)
            chain.update(sc, clip=clip)
        finally:
            torch.randn_like = orig
        torch.cuda.synchronize()
        got_total = float(chain._clip_buf[1].item())
        assert abs(got_total - float(total)) <= 1e-6 * float(total)
        pairs = {"theta": (chain.layout.views(chain.theta), [p.data for _, p in named])}
        if chain.v is not None:
            pairs["v"] = (chain.layout.views(chain.v), [vs[n] for n in names])
            pairs["m"] = (chain.layout.views(chain.m), [ms[n] for n in names])
            pairs["s"] = (chain.layout.views(chain.s), [ss[n] for n in names])
        if chain.buf is not None:
            pairs["sgd_buf"] = (chain.layout.views(chain.buf), [opt.state[p]["momentum_buffer"] for _, p in named])
        for key, (got, want) in pairs.items():
            for n_, g_, w_ in zip(names, got, want):
                rel = ((g_.double() - w_.double()).abs().max() / w_.double().abs().max().clamp_min(1e-30)).item()
                assert rel <= 1e-6, f"{variant_name} step {t} {key}[{n_}]: rel {rel:.2e}"
    # SGLD: the third step's small gradients fall below the threshold (coefficient clamped to 1); Adam normalises the
    # gradient scale away, its clamped regime is covered by the oracle test above and the reference goldens
    assert min(coefs) < 1.0 and (variant_name == "adam_csghmc" or max(coefs) == 1.0), coefs


def test_runner_clip_grad_plumbing(cuda_device, tmp_path):
    """Runner level: a huge args.clip_grad (coefficient clamped to 1: x * 1.0f is exact) leaves a cSGLD run bit-identical to
    one without clipping; a small one changes it; the burn-in runners ignore the attribute like the reference's do."""
    import shard_util
    from bayesdll_b200.methods import csgld, sghmc

    def run(method, clip):
        torch.manual_seed(3)
        runner = shard_util.make_runner(method, cuda_device, tmp_path, 2, eval_shard=False)
        runner.args.clip_grad = clip
        runner.__init__(runner.net, runner.net0, runner.args, runner.logger)          # re-read args.clip_grad
        loader = shard_util.make_loader(batches=(16, 16), seed=4)
        if hasattr(runner, "cyclical_scheduler"):
            runner.cyclical_scheduler.current_epoch = 0
            runner.train_one_epoch(loader)
        else:
            runner.train_one_epoch(loader, collect=False, bi=0)
        torch.cuda.synchronize()
        return runner, runner.model.chain.theta.clone()
    r0, base = run("csgld", None)
    r1, huge = run("csgld", 1e9)
    r2, small = run("csgld", 1e-3)
    assert r0.clip_grad is None and r1.clip_grad == 1e9
    assert torch.equal(base, huge)
    assert not torch.equal(base, small) and torch.isfinite(small).all()
    r3, a = run("sghmc", None)
    r4, b = run("sghmc", 1e-3)
    assert r4.clip_grad is None and torch.equal(a, b)
