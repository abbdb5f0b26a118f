"""GPU parity against the reference's own PyTorch update on the SAME device: the update statements of the reference in
torch CUDA eager ops + the real ``torch.optim.SGD`` (tests/eager_reference.py) vs the fused kernel in its default mode
(``div=recip`` = what torch CUDA does for ``tensor / python_scalar``), identical injected noise, three chained steps.
North-star tolerance: fp32 rel 1e-6 per step; all five update rules are in fact bit-identical (theta, momentum, Adam
moments, SGD momentum buffer), which pins the reciprocal-division semantics: torch CUDA multiplies by
fp32(1.0 / s) with the reciprocal taken in double (tools/probe_torch_div.py)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import eager_reference as er

pytestmark = pytest.mark.gpu

SHAPES = [("layers.0.weight", (61, 37)), ("layers.0.bias", (61,)), ("layers.1.weight", (33, 61)), ("layers.1.bias", (33,)),
          ("norm.weight", (33,)), ("norm.bias", (33,)), ("classifier.weight", (7, 33)), ("classifier.bias", (7,))]
READOUT = "classifier"
HP = dict(lr_body=1e-3, lr_head=2e-2, ND=1840, Ninflate=3.0, prior_sig=0.9, nd=0.7, alpha=0.18, beta1=0.9, beta2=0.99, eps=1e-6,
          temperature=1.3)


class _Net(nn.Module):
    readout_name = READOUT

    def __init__(self, gen):
        super().__init__()
        for name, shape in SHAPES:
            mod, parts = self, name.split(".")
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    mod.add_module(part, nn.Module())
                mod = getattr(mod, part)
            mod.register_parameter(parts[-1], nn.Parameter(torch.randn(shape, generator=gen) * 0.1))


def _ulp_diff(a, b):
    ia = a.view(torch.int32).to(torch.int64)
    ib = b.view(torch.int32).to(torch.int64)
    return (ia - ib).abs().max().item()


@pytest.mark.parametrize("variant_name", ["sgld", "sghmc", "csghmc", "adam_sghmc", "adam_csghmc"])
@pytest.mark.parametrize("bias", ["informative", "uninformative"])
def test_fused_kernel_equals_torch_cuda_eager(cuda_device, variant_name, bias):
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.chain import ChainState
    dev = cuda_device
    gen = torch.Generator().manual_seed(11)
    mu = 0.5 if variant_name in ("sgld", "adam_sghmc") else 0.0
    variant = dict(sgld=_lib.SGLD, sghmc=_lib.SGHMC, csghmc=_lib.CSGHMC, adam_sghmc=_lib.ADAM_SGHMC,
                   adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    N = HP["ND"] * HP["Ninflate"]
    # --- reference side: plain modules on CUDA, torch eager + torch.optim.SGD -----------------------------------
    ref, ref0 = _Net(gen).to(dev), _Net(gen).to(dev)
    # --- product side: same initial values, flat state, fused kernel ----------------------------------------------
    net, net0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net.load_state_dict(ref.state_dict())
    net0.load_state_dict(ref0.state_dict())
    chain = ChainState(net, net0, variant=variant, bias_mode=bias, mu=mu, noise="torch", seed=0)
    named = list(ref.named_parameters())
    names = [n for n, _ in named]
    p0s = [p for _, p in ref0.named_parameters()]
    body = [p for n, p in named if READOUT not in n]
    head = [p for n, p in named if READOUT in n]
    opt = er.make_sgd(body, head, HP["lr_body"], HP["lr_head"], mu)
    vs = {n: torch.zeros_like(p) for n, p in named}
    ms = {n: torch.zeros_like(p) for n, p in named}
    ss = {n: torch.zeros_like(p) for n, p in named}
    worst = {}
    for t in range(1, 4):
        grads = [(torch.randn(p.shape, generator=gen) * 0.05).to(dev) for _, p in named]
        xis = [torch.randn(p.shape, generator=gen).to(dev) for _, p in named]
        # reference
        for (_, p), g in zip(named, grads):
            p.grad = g.clone()
        kw = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], N=N, prior_sig=HP["prior_sig"], nd=HP["nd"])
        if variant_name == "sgld":
            er.sgld(named, p0s, xis, READOUT, bias=bias, **kw)
            opt.step()
        elif variant_name == "sghmc":
            er.sghmc(named, p0s, xis, vs, READOUT, alpha=HP["alpha"], bias=bias, **kw)
            opt.step()
        elif variant_name == "csghmc":
            er.csghmc(named, xis, vs, READOUT, alpha=HP["alpha"], should_sample=(t != 2), **kw)
        else:
            er.adam(named, p0s, xis, vs, ms, ss, READOUT, alpha=HP["alpha"], beta1=HP["beta1"], beta2=HP["beta2"], eps=HP["eps"],
                    t=t, bias=bias, cyclical=(variant_name == "adam_csghmc"), temperature=HP["temperature"], **kw)
            opt.step()
        # product: same gradients in p.grad, same noise through the injected-noise buffer
        for p, g in zip(chain.params, grads):
            p.grad = g.clone()
        xi_flat = torch.zeros(chain.layout.n_padded, device=dev)
        for view, xi in zip(chain.layout.views(xi_flat), xis):
            view.copy_(xi)
        sc = ops.make_scalars(variant, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                              prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=HP["alpha"], mu=mu, beta1=HP["beta1"],
                              beta2=HP["beta2"], eps=HP["eps"], temperature=HP["temperature"], t=t, add_noise=(t != 2),
                              div_mode=_lib.DIV_RECIP)
        runs_dev, nruns = chain._gradient_table()
        sc.first_step = int(chain.sgd_steps == 0)
        ops.step(variant, chain.theta, None, chain.theta0, chain.v, chain.m, chain.s, chain.buf, runs_dev, nruns, sc,
                 ops.make_noise(xi=xi_flat), runs_host=chain._run_np.ctypes.data)      # the launch ChainState.update makes
        chain.sgd_steps += 1
        torch.cuda.synchronize()
        pairs = {"theta": (chain.layout.views(chain.theta), [p.data for _, p in named])}
        if chain.v is not None:
            pairs["v"] = (chain.layout.views(chain.v), [vs[n] for n in names])
        if chain.m is not None:
            pairs["m"] = (chain.layout.views(chain.m), [ms[n] for n in names])
            pairs["s"] = (chain.layout.views(chain.s), [ss[n] for n in names])
        if chain.buf is not None:
            pairs["sgd_buf"] = (chain.layout.views(chain.buf), [opt.state[p]["momentum_buffer"] for _, p in named])
        for key, (got, want) in pairs.items():
            for n_, g_, w_ in zip(names, got, want):
                assert g_.shape == w_.shape
                worst[key] = max(worst.get(key, 0), _ulp_diff(g_.contiguous(), w_.contiguous()))
                rel = ((g_.double() - w_.double()).abs().max() / w_.double().abs().max().clamp_min(1e-30)).item()
                assert rel <= 1e-6, f"{variant_name}/{bias} step {t} {key}[{n_}]: rel {rel:.2e}"
    print(f"{variant_name}/{bias}: max ulp distance to torch CUDA eager {worst}")
    assert all(u == 0 for u in worst.values()), worst       # every state vector of every update rule: bit-identical


@pytest.mark.parametrize("n", [4, 100_004, 1_000_000])
def test_capture_and_draw_equal_torch_cuda_eager(cuda_device, n):
    """Running moments (methods/sgld.py:95-102,239-246), Welford with the reference's n = 3, 5, 7 ... (methods/csghmc.py:
    327-348), variance + posterior draw (methods/sgld.py:292-297,338-348; methods/csghmc.py:451-459) written in torch CUDA
    eager ops vs the kernels in the default division mode: bit-identical."""
    from bayesdll_b200 import _lib, ops
    dev = cuda_device
    gen = torch.Generator().manual_seed(n)
    thetas = [(torch.randn(n, generator=gen) * 0.3).to(dev) for _ in range(6)]
    # --- running average ---
    mom1, mom2, cnt = thetas[0] * 1.0, thetas[0] ** 2, 1
    k1, k2 = torch.empty(n, device=dev), torch.empty(n, device=dev)
    ops.moments_avg(thetas[0], k1, k2, 0, init=True)
    for th in thetas[1:]:
        mom1 = (th + cnt * mom1) / (cnt + 1)
        mom2 = (th ** 2 + cnt * mom2) / (cnt + 1)
        ops.moments_avg(th, k1, k2, cnt)                                   # default: DIV_RECIP
        cnt += 1
    assert torch.equal(k1, mom1) and torch.equal(k2, mom2)
    # --- Welford, double-counted sample count ---
    mean, M2, count = thetas[0].clone(), torch.zeros(n, device=dev), 1
    w1, w2 = torch.empty(n, device=dev), torch.empty(n, device=dev)
    ops.moments_welford(thetas[0], w1, w2, 1, init=True)
    count += 1
    for th in thetas[1:]:
        nn_ = count + 1
        delta = th - mean
        mean = mean + delta / nn_
        delta2 = th - mean
        M2 = M2 + delta * delta2
        ops.moments_welford(th, w1, w2, nn_)
        count = nn_ + 1
    assert torch.equal(w1, mean) and torch.equal(w2, M2)
    # --- variance + draw ---
    eps = torch.randn(n, generator=gen).to(dev)
    ratio = cnt / (cnt - 1)
    var = ratio * (mom2 - mom1 ** 2)
    var.clamp_(min=1e-12)
    want = mom1 + var.sqrt() * eps
    out = torch.empty(n, device=dev)
    ops.draw(k1, k2, out, ops.VAR_FROM_MOMENTS, ratio, ops.make_noise(xi=eps))
    assert torch.equal(out, want)
    var_w = (M2 / (count - 1)).clamp(min=1e-12)
    want_w = mean + var_w.sqrt() * eps
    ops.draw(w1, w2, out, ops.VAR_FROM_WELFORD, float(count - 1), ops.make_noise(xi=eps))
    assert torch.equal(out, want_w)


@pytest.mark.parametrize("B,K,S", [(16, 37, 5), (64, 37, 40), (7, 10, 1)])
def test_ensemble_and_ce_close_to_torch_cuda_eager(cuda_device, B, K, S):
    """logsumexp_S(log_softmax_K) - log S, CE and error count (methods/sgld.py:299-306) vs torch CUDA eager: these go
    through exp / log whose last bits differ between implementations -> abs/rel 1e-5 (SURVEY 8c iii); errors exact."""
    from bayesdll_b200 import ops
    dev = cuda_device
    gen = torch.Generator().manual_seed(B * S)
    la = (torch.randn(B, K, S, generator=gen) * 3).to(dev)
    y = torch.randint(0, K, (B,), generator=gen).to(dev)
    want = la.log_softmax(dim=1).logsumexp(-1) - np.log(S)
    out = torch.empty(B, K, device=dev)
    ops.ensemble(la, out, S)
    torch.testing.assert_close(out, want.float(), atol=1e-5, rtol=1e-5)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.ce_err(out, y, loss, err)
    ref_loss = torch.nn.functional.cross_entropy(want.float(), y).item()
    assert abs(loss.item() / B - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    assert err.item() == want.argmax(1).ne(y).sum().item()


def test_noise_torch_mode_reproduces_reference_rng_stream(cuda_device):
    """hparams noise=torch: the chain draws ``torch.randn_like`` per tensor in named_parameters() order, exactly the calls
    the reference makes (methods/sghmc.py:501).  After the same ``torch.manual_seed`` both consume the same CUDA
    generator stream, so a seeded reference run on a GPU is reproduced bit for bit -- noise included."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.chain import ChainState
    dev = cuda_device
    gen = torch.Generator().manual_seed(3)
    ref, ref0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net, net0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net.load_state_dict(ref.state_dict())
    net0.load_state_dict(ref0.state_dict())
    chain = ChainState(net, net0, variant=_lib.SGHMC, bias_mode="informative", noise="torch", seed=0)
    named = list(ref.named_parameters())
    p0s = [p for _, p in ref0.named_parameters()]
    opt = er.make_sgd([p for n, p in named if READOUT not in n], [p for n, p in named if READOUT in n], HP["lr_body"],
                      HP["lr_head"], 0.0)
    vs = {n: torch.zeros_like(p) for n, p in named}
    N = HP["ND"] * HP["Ninflate"]
    for t in range(3):
        grads = [(torch.randn(p.shape, generator=gen) * 0.05).to(dev) for _, p in named]
        for (_, p), g in zip(named, grads):
            p.grad = g.clone()
        torch.manual_seed(100 + t)
        xis = [torch.randn_like(p) for _, p in named]                      # what the reference's loop would draw
        er.sghmc(named, p0s, xis, vs, READOUT, lr_body=HP["lr_body"], lr_head=HP["lr_head"], N=N, prior_sig=HP["prior_sig"],
                 nd=HP["nd"], alpha=HP["alpha"], bias="informative")
        opt.step()
        for p, g in zip(chain.params, grads):
            p.grad = g.clone()
        torch.manual_seed(100 + t)
        chain.update(ops.make_scalars(_lib.SGHMC, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"],
                                      Ninflate=HP["Ninflate"], prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=HP["alpha"]))
    torch.cuda.synchronize()
    for (n_, p), view, vview in zip(named, chain.layout.views(chain.theta), chain.layout.views(chain.v)):
        assert torch.equal(view, p.data), n_
        assert torch.equal(vview, vs[n_]), n_


@pytest.mark.parametrize("variant_name", ["sghmc", "adam_csghmc"])
def test_parameters_without_gradient_are_left_untouched(cuda_device, variant_name):
    """``if p.grad is not None`` (methods/sghmc.py:484): frozen / unused parameters keep theta and momentum; no flat
    gradient buffer exists in that case, so the kernel must not read a gradient for those runs at all."""
    from bayesdll_b200 import _lib, ops
    from bayesdll_b200.chain import ChainState
    dev = cuda_device
    gen = torch.Generator().manual_seed(21)
    variant = dict(sghmc=_lib.SGHMC, adam_csghmc=_lib.ADAM_CSGHMC)[variant_name]
    ref, ref0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net, net0 = _Net(gen).to(dev), _Net(gen).to(dev)
    net.load_state_dict(ref.state_dict())
    net0.load_state_dict(ref0.state_dict())
    chain = ChainState(net, net0, variant=variant, bias_mode="informative", noise="torch", seed=0)
    named = list(ref.named_parameters())
    names = [n for n, _ in named]
    frozen = {"layers.1.weight", "norm.bias"}
    p0s = [p for _, p in ref0.named_parameters()]
    opt = er.make_sgd([p for n, p in named if READOUT not in n], [p for n, p in named if READOUT in n], HP["lr_body"],
                      HP["lr_head"], 0.0)
    vs = {n: torch.zeros_like(p) for n, p in named}
    ms = {n: torch.zeros_like(p) for n, p in named}
    ss = {n: torch.zeros_like(p) for n, p in named}
    N = HP["ND"] * HP["Ninflate"]
    before = {n: p.detach().clone() for n, p in named if n in frozen}
    for t in range(1, 3):
        grads = [None if n in frozen else (torch.randn(p.shape, generator=gen) * 0.05).to(dev) for n, p in named]
        xis = [torch.randn(p.shape, generator=gen).to(dev) for _, p in named]
        live = [(n, p) for (n, p), g in zip(named, grads) if g is not None]
        live_p0 = [p0 for p0, g in zip(p0s, grads) if g is not None]
        live_xi = [x for x, g in zip(xis, grads) if g is not None]
        for (_, p), g in zip(named, grads):
            p.grad = None if g is None else g.clone()
        kw = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], N=N, prior_sig=HP["prior_sig"], nd=HP["nd"])
        if variant_name == "sghmc":
            er.sghmc(live, live_p0, live_xi, vs, READOUT, alpha=HP["alpha"], bias="informative", **kw)
        else:
            er.adam(live, live_p0, live_xi, vs, ms, ss, READOUT, alpha=HP["alpha"], beta1=HP["beta1"], beta2=HP["beta2"],
                    eps=HP["eps"], t=t, bias="informative", cyclical=True, temperature=HP["temperature"], **kw)
        opt.step()
        for p, g in zip(chain.params, grads):
            p.grad = None if g is None else g.clone()
        xi_flat = torch.zeros(chain.layout.n_padded, device=dev)
        for view, xi in zip(chain.layout.views(xi_flat), xis):
            view.copy_(xi)
        sc = ops.make_scalars(variant, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                              prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=HP["alpha"], beta1=HP["beta1"], beta2=HP["beta2"],
                              eps=HP["eps"], temperature=HP["temperature"], t=t)
        runs_dev, nruns = chain._gradient_table()
        assert chain.g_flat is None                               # nothing was gathered: every live gradient is read in place
        ops.step(variant, chain.theta, None, chain.theta0, chain.v, chain.m, chain.s, None, runs_dev, nruns, sc,
                 ops.make_noise(xi=xi_flat), runs_host=chain._run_np.ctypes.data)
    torch.cuda.synchronize()
    for (n_, p), th, vv in zip(named, chain.layout.views(chain.theta), chain.layout.views(chain.v)):
        assert torch.equal(th, p.data), n_
        assert torch.equal(vv, vs[n_]), n_
        if n_ in frozen:
            assert torch.equal(th, before[n_]) and not vv.any()
