"""GPU end-to-end parity: the drop-in ``Runner.train()`` against recordings of the reference's own
``Runner.train()`` (tests/golden/runner_*.npz, produced by oracle/make_golden_runner.py).

Gradients are injected through ``InjectNet`` and every ``torch.randn_like`` -- sampler noise and posterior draws --
is served from the recorded tape, so both implementations consume identical inputs in identical order.  The test
reads like the reference's own usage: build args, construct ``Runner(net, net0, args, logger)``, call ``train``.
"""
import logging
import os

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

CASES = ["sgld", "sghmc", "adam_sghmc", "sgld_nst0", "csgld", "csghmc", "adam_csghmc", "csghmc_fs"]
BIT_EXACT = {"sgld", "sghmc", "sgld_nst0", "csgld", "csghmc", "csghmc_fs"}


def _logger():
    lg = logging.getLogger("test_runner")
    lg.addHandler(logging.NullHandler())
    lg.propagate = False
    return lg


def _run(name, device, tmp_path, extra_hp=None):
    import importlib
    from oracle import make_golden_runner as mgr
    from oracle import refshim
    z = np.load(gu.golden_path(f"runner_{name}"), allow_pickle=False)
    method, hp, over = mgr.case_spec(name)
    hp = dict(hp, noise="torch", div="ieee", **(extra_hp or {}))
    seed = mgr.case_seed(name)
    net, net0 = mgr.InjectNet(seed, z["G"]), mgr.InjectNet(seed + 1)
    assert np.array_equal(torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy(), z["theta_init"])
    args = mgr.make_args(hp, str(tmp_path), device, **over)
    Runner = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner
    runner = Runner(net, net0, args, _logger())
    runner.criterion = mgr.InjectCriterion()
    evals = []
    orig = runner.evaluate

    def rec(loader):
        r = orig(loader)
        evals.append(r)
        return r
    runner.evaluate = rec
    runner._bma_calls = []
    if hasattr(runner, "evaluate_full_samples"):
        orig_bma = runner.evaluate_full_samples

        def rec_bma(*a, **k):
            r = orig_bma(*a, **k)
            runner._bma_calls.append((r, runner._full_sample_files()))
            return r
        runner.evaluate_full_samples = rec_bma
    loaders = mgr.loaders_from_arrays(z)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with refshim.injected_noise(z["tape"]) as tape:
            ret = runner.train(*loaders)
    finally:
        os.chdir(cwd)
    return z, runner, evals, ret, tape


def _close(name, got, want, what, rel=1e-6):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if name in BIT_EXACT:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
            f"{what}: {(got.view(np.uint32) != want.view(np.uint32)).sum()} of {got.size} elements differ"
    else:
        assert gu.max_rel(got, want) <= rel, f"{what}: rel {gu.max_rel(got, want):.2e}"


@pytest.mark.parametrize("name", CASES)
def test_runner_train_matches_reference(cuda_device, tmp_path, name):
    z, runner, evals, ret, tape = _run(name, cuda_device, tmp_path)
    # every recorded draw was consumed, in order, and none beyond
    assert tape.pos == int(z["tape_used"])
    if "n_bma" in z.files:
        _check_bma(z, runner, tmp_path)
    dense = lambda flat: runner._dense(flat).cpu().numpy()
    _close(name, dense(runner.model.chain.theta), z["theta_final"], "theta_final")
    if "v_final" in z.files:
        _close(name, dense(runner.model.chain.v), z["v_final"], "momentum_buffer")
    if "post_theta_mom1" in z.files:
        assert runner.post_theta_cnt == int(z["post_theta_cnt"])
        _close(name, runner.post_theta_mom1.cpu().numpy(), z["post_theta_mom1"], "post_theta_mom1")
        if "post_theta_mom2" in z.files:
            _close(name, runner.post_theta_mom2.cpu().numpy(), z["post_theta_mom2"], "post_theta_mom2")
    if "cycles" in z.files:
        assert sorted(runner.cycle_theta_mom1) == z["cycles"].tolist()
        assert runner.samples_collected == int(z["samples_collected"])
        for c in z["cycles"].tolist():
            assert runner.samples_per_cycle[c] == int(z[f"cyc{c}_count"])
            _close(name, runner.cycle_theta_mom1[c].cpu().numpy(), z[f"cyc{c}_mom1"], f"cycle {c} mom1")
            _close(name, runner.cycle_theta_mom2[c].cpu().numpy(), z[f"cyc{c}_mom2"], f"cycle {c} mom2", rel=2e-5)
            np.testing.assert_allclose(runner.cycle_likelihoods[c], z[f"cyc{c}_lik"], rtol=2e-5)
        np.testing.assert_allclose(ret["losses_train"], z["losses_train"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ret["losses_test"], z["losses_test"], rtol=2e-5, atol=1e-6)
    # evaluate(): same number of calls, same outputs (logits go through a GPU matmul -> tolerance, SURVEY 8c iii)
    assert len(evals) == int(z["n_evals"])
    for i, (loss, err, targets, logits, logits_all) in enumerate(evals):
        assert np.array_equal(targets, z[f"eval{i}_targets"]) and targets.dtype == np.int64
        assert logits.dtype == np.float32 and logits_all.dtype == np.float32
        assert logits_all.shape == z[f"eval{i}_logits_all"].shape
        np.testing.assert_allclose(logits_all, z[f"eval{i}_logits_all"], atol=1e-5, rtol=1e-5)
        np.testing.assert_allclose(logits, z[f"eval{i}_logits"], atol=1e-5, rtol=1e-5)
        assert abs(loss - float(z[f"eval{i}_loss"])) <= 1e-5 * max(1.0, abs(float(z[f"eval{i}_loss"])))
        assert err == float(z[f"eval{i}_err"])
    # on-disk contract: same checkpoint files and keys
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith("ckpt.pt"))
    assert files == z["ckpt_files"].tolist()
    if files:
        ck = torch.load(os.path.join(tmp_path, files[-1]), map_location="cpu", weights_only=False)
        assert sorted(ck.keys()) == z["ckpt_keys"].tolist()
        kind = "none" if ck.get("last_theta") is None else ("vector" if torch.is_tensor(ck["last_theta"]) else "state_dict")
        assert kind == str(z["ckpt_last_theta_kind"])
        n_dense = z["theta_init"].size
        for key in ("post_theta_mom1", "post_theta_mom2"):
            if ck.get(key) is not None:
                assert ck[key].shape == (n_dense,)          # dense parameters_to_vector order, no padding
        if kind == "vector":
            assert ck["last_theta"].shape == (n_dense,)
        assert os.path.exists(os.path.join(tmp_path, "logits_test.pkl"))


def _check_bma(z, runner, tmp_path):
    """Bayesian model average over the stored raw samples vs the reference's evaluate_full_samples recordings."""
    import pickle
    calls = runner._bma_calls
    assert len(calls) == int(z["n_bma"])
    for i, (res, files) in enumerate(calls):
        assert files == z[f"bma{i}_files"].tolist()
        for ds in ("train", "val", "test"):
            r = res[ds]
            assert r["num_models"] == int(z[f"bma{i}_{ds}_num_models"])
            assert r["error"] == float(z[f"bma{i}_{ds}_error"])
            assert r["individual_avg_error"] == float(z[f"bma{i}_{ds}_individual_avg_error"])
            for k in ("loss", "individual_avg_loss"):
                want = float(z[f"bma{i}_{ds}_{k}"])
                assert abs(r[k] - want) <= 1e-5 * max(1.0, abs(want)), (i, ds, k, r[k], want)
            if f"bma{i}_{ds}_logits" in z.files:
                assert np.array_equal(r["targets"], z[f"bma{i}_{ds}_targets"]) and r["targets"].dtype == np.int64
                assert r["logits_all"].shape == z[f"bma{i}_{ds}_logits_all"].shape and r["logits"].dtype == np.float32
                np.testing.assert_allclose(r["logits_all"], z[f"bma{i}_{ds}_logits_all"], atol=1e-5, rtol=1e-5)
                np.testing.assert_allclose(r["logits"], z[f"bma{i}_{ds}_logits"], atol=1e-5, rtol=1e-5)
                # the average itself is bit-exact given the per-model logits: fp32 running sum in file order / S
                la = r["logits_all"]
                acc = la[:, :, 0].copy()
                for m in range(1, la.shape[2]):
                    acc += la[:, :, m]
                assert np.array_equal((acc / la.shape[2]).view(np.uint32), r["logits"].view(np.uint32))
    # on-disk contract: same files, a loadable state_dict with the reference's keys, the stored sample bit-exact
    assert sorted(f for f in os.listdir(tmp_path) if "bma" in f) == z["bma_files_in_logdir"].tolist()
    last = calls[-1][1][-1]
    sd = torch.load(os.path.join(tmp_path, last), map_location="cpu")
    assert list(sd.keys()) == z["fs_state_keys"].tolist()
    names = [n for n, _ in runner.net.named_parameters()]
    got = torch.cat([sd[n].reshape(-1) for n in names]).numpy()
    assert np.array_equal(got.view(np.uint32), z["fs_last_sample"].view(np.uint32))
    with open(os.path.join(tmp_path, "bma_evaluation_results.pkl"), "rb") as f:
        assert sorted(pickle.load(f)["test"].keys()) == z["bma_pkl_keys"].tolist()


@pytest.mark.parametrize("extra", [dict(), dict(fs_ring_slots=2), dict(io="sync"), dict(graph=0)],
                         ids=["resident", "evicting-ring", "sync-io", "eager-forward"])
def test_csghmc_fs_sample_store_and_bma(cuda_device, tmp_path, extra):
    """csghmc_fs: raw samples go to the HBM ring (TMA copy) and are spilled asynchronously; the BMA runs on the resident
    slots.  A 2-slot ring forces eviction, so older samples are re-read from the spilled files: same results.  The stored
    models' forwards are replayed as CUDA graphs sharing one memory pool (graph=0: eager)."""
    from bayesdll_b200.graphfwd import GraphedForward
    replays = GraphedForward.total_replays
    z, runner, evals, ret, tape = _run("csghmc_fs", cuda_device, tmp_path, extra_hp=extra)
    assert (GraphedForward.total_replays > replays) == ("graph" not in extra)
    assert tape.pos == int(z["tape_used"])
    got = runner._dense(runner.model.chain.theta).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), z["theta_final"].view(np.uint32))
    _check_bma(z, runner, tmp_path)
    cap = runner._fs_ring.capacity
    assert cap == (2 if "fs_ring_slots" in extra else 4)
    assert len(runner._fs_resident) == cap


def test_runner_accepts_bare_cuda_device(cuda_device, tmp_path):
    """args.device = torch.device('cuda') (no index), as the reference's drivers build it (demo_vision.py:64): the writer
    thread, the flat state and the BMA must all cope.  (A bare device once killed the writer thread and hung flush().)"""
    z, runner, _, _, _ = _run("csghmc_fs", torch.device("cuda"), tmp_path)
    assert runner._writer.device.index is not None
    _check_bma(z, runner, tmp_path)


def test_async_and_sync_checkpoints_are_identical(cuda_device, tmp_path):
    """io=async (side-stream D2H + writer thread) writes byte-for-byte the tensors io=sync writes."""
    outs = {}
    for mode in ("async", "sync"):
        d = tmp_path / mode
        d.mkdir()
        _, runner, _, _, _ = _run("adam_sghmc", cuda_device, d, extra_hp=dict(io=mode))
        assert runner._writer.stats["jobs"] >= 1
        outs[mode] = torch.load(os.path.join(d, "ckpt.pt"), map_location="cpu", weights_only=False)

    def same(a, b):
        if torch.is_tensor(a):
            return torch.equal(a, b)
        if isinstance(a, dict):
            return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
        if isinstance(a, (list, tuple)):
            return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b
    assert same(outs["async"], outs["sync"])


@pytest.mark.parametrize("name", ["sghmc", "csgld", "csghmc", "sgld_nst0"])
def test_unfused_capture_is_identical(cuda_device, tmp_path, name):
    """fuse=0 launches the moment capture as its own kernel after the step (the reference's statement order); the default
    folds it into the step kernel.  Same trajectory, same moments, bit for bit."""
    z, runner, _, _, tape = _run(name, cuda_device, tmp_path, extra_hp=dict(fuse=0))
    assert not runner.fuse_capture and tape.pos == int(z["tape_used"])
    got = runner._dense(runner.model.chain.theta).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), z["theta_final"].view(np.uint32))
    if "post_theta_mom1" in z.files:
        assert np.array_equal(runner.post_theta_mom1.cpu().numpy().view(np.uint32), z["post_theta_mom1"].view(np.uint32))
        assert runner.post_theta_cnt == int(z["post_theta_cnt"])
    if "cycles" in z.files:
        for c in z["cycles"].tolist():
            assert runner.samples_per_cycle[c] == int(z[f"cyc{c}_count"])
            assert np.array_equal(runner.cycle_theta_mom1[c].cpu().numpy().view(np.uint32), z[f"cyc{c}_mom1"].view(np.uint32))


def test_runner_flat_gradient_mode_is_identical(cuda_device, tmp_path):
    """grad=flat (gather into the flat buffer) and grad=table (read p.grad in place) give the same trajectory."""
    z, runner, _, _, _ = _run("sghmc", cuda_device, tmp_path, extra_hp=dict(grad="flat"))
    got = runner._dense(runner.model.chain.theta).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), z["theta_final"].view(np.uint32))


def test_model_forward_keeps_reference_contract(cuda_device):
    """Model.forward(x, y, net, net0, criterion, lrs, Ninflate, nd) -> (float, detached [B,K]); optimizer.step() after it
    is harmless; state is reachable under the reference's attribute names."""
    from oracle import make_golden_runner as mgr
    from bayesdll_b200.methods import adam_sghmc
    z = np.load(gu.golden_path("runner_adam_sghmc"), allow_pickle=False)
    net, net0 = mgr.InjectNet(1, z["G"]).to(cuda_device), mgr.InjectNet(2).to(cuda_device)
    args = mgr.make_args(dict(mgr.CASES["adam_sghmc"][1]), "/tmp", cuda_device, momentum=0.5, epochs=1)
    runner = adam_sghmc.Runner(net, net0, args, _logger())
    x = torch.randn(mgr.BATCH, 1, 4, 4, device=cuda_device)
    y = torch.randint(0, mgr.K_CLASSES, (mgr.BATCH,), device=cuda_device)
    net.train()
    before = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    loss, out = runner.model(x, y, runner.net, runner.net0, mgr.InjectCriterion(),
                             [pg["lr"] for pg in runner.optimizer.param_groups], runner.Ninflate, runner.nd)
    after = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    runner.optimizer.step()
    assert isinstance(loss, float) and out.shape == (mgr.BATCH, mgr.K_CLASSES) and not out.requires_grad
    assert not torch.equal(before, after)
    assert torch.equal(after, torch.cat([p.detach().reshape(-1) for p in net.parameters()]))   # step() is a no-op
    names = [n for n, _ in net.named_parameters()]
    assert list(runner.model.momentum_buffer) == names and list(runner.model.m) == names and list(runner.model.v) == names
    assert runner.model.t == 1
    sd = runner.optimizer.state_dict()
    assert len(sd["state"]) == len(names) and all("momentum_buffer" in s for s in sd["state"].values())


@pytest.mark.parametrize("name", ["sghmc", "csgld", "csghmc"])
def test_sharded_ensemble_equals_runner_evaluate(cuda_device, tmp_path, name):
    """The sample-sharded formulation (prob sums + one all-reduce, here world=1) reproduces Runner.evaluate with the
    same Philox keys: sharding changes where samples are computed, not what is computed."""
    import importlib
    from oracle import make_golden_runner as mgr
    from bayesdll_b200 import dist as bdist
    z = np.load(gu.golden_path(f"runner_{name}"), allow_pickle=False)
    method, hp, over = mgr.CASES[name]
    seed = 500 + sorted(mgr.CASES).index(name)
    net, net0 = mgr.InjectNet(seed, z["G"]), mgr.InjectNet(seed + 1)
    args = mgr.make_args(dict(hp), str(tmp_path), cuda_device, **over)     # default noise = in-kernel Philox
    runner = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner(net, net0, args, _logger())
    runner.criterion = mgr.InjectCriterion()
    loaders = mgr.loaders_from_arrays(z)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        runner.train(*loaders)
    finally:
        os.chdir(cwd)
    loss, err, targets, logits, _ = runner.evaluate(loaders[2])
    comps, mixture = bdist.components_from_runner(runner)
    ens = bdist.ShardedEnsemble(runner.net, runner.model.chain.layout, comps, runner.nst, runner.seed, mixture=mixture,
                                eval_id=runner._eval_calls, div_mode=runner.div_mode)
    loss2, err2, targets2, logits2 = ens.evaluate(loaders[2])
    assert np.array_equal(targets, targets2)
    np.testing.assert_allclose(logits2, logits, rtol=2e-5, atol=2e-5)
    assert abs(loss - loss2) < 2e-5 * max(1.0, abs(loss)) and err == err2
    from bayesdll_b200 import calibration
    e1, m1, n1 = calibration.analyze(targets, logits2, 15, None)
    e2, m2, n2 = ens.calibrate(targets2, logits2, 15)
    assert abs(e1 - e2) < 1e-9 and abs(m1 - m2) < 1e-9 and abs(n1 - n2) < 1e-9


def test_csgld_full_sample_uses_hbm_ring(cuda_device, tmp_path):
    """args.full_sample: raw samples land in the preallocated HBM ring through the TMA copy kernel; ``all_samples``
    keeps the reference's "<epoch>_<batch>" keys and every stored vector equals theta at capture time."""
    import importlib
    from oracle import make_golden_runner as mgr
    z = np.load(gu.golden_path("runner_csgld"), allow_pickle=False)
    method, hp, over = mgr.CASES["csgld"]
    seed = 500 + sorted(mgr.CASES).index("csgld")
    net, net0 = mgr.InjectNet(seed, z["G"]), mgr.InjectNet(seed + 1)
    args = mgr.make_args(dict(hp, noise="torch", div="ieee"), str(tmp_path), cuda_device, **over)
    args.full_sample = True
    from oracle import refshim
    runner = importlib.import_module("bayesdll_b200.methods.csgld").Runner(net, net0, args, _logger())
    runner.criterion = mgr.InjectCriterion()
    captured = {}
    orig = runner._capture

    def spy(cycle, epoch, batch_idx, **kw):
        orig(cycle, epoch, batch_idx, **kw)
        captured[f"{epoch}_{batch_idx}"] = runner._dense(runner.model.chain.theta).clone()
    runner._capture = spy
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with refshim.injected_noise(z["tape"]):
            runner.train(*mgr.loaders_from_arrays(z))
    finally:
        os.chdir(cwd)
    assert runner._ring is not None and runner._ring.capacity == int(z["samples_collected"])
    assert sorted(runner.all_samples) == sorted(captured) and len(captured) == int(z["samples_collected"])
    for k, v in captured.items():
        assert torch.equal(runner.all_samples[k], v)
    assert os.path.exists(os.path.join(tmp_path, "all_samples_TEST.ckpt"))
    # trajectory unchanged by the capture
    got = runner._dense(runner.model.chain.theta).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), z["theta_final"].view(np.uint32))


@pytest.mark.parametrize("graph_train", [0, 1])
@pytest.mark.parametrize("name", ["real_sghmc", "real_csghmc", "real_adam_csghmc", "real_csghmc_fs"])
def test_runner_real_network_end_to_end(cuda_device, tmp_path, name, graph_train):
    """A real conv + BatchNorm network trained through ordinary autograd: the reference on CPU (recorded) vs the drop-in
    on the GPU, both fed the same noise tape.  Gradients come from different conv implementations (1e-7-level
    differences), so trajectories are compared with a tolerance instead of bit-for-bit.  graph_train=1: the whole
    ``Runner.train()`` with forward + backward replayed as a CUDA graph (fused capture, raw-sample store, BMA included)."""
    import importlib
    from oracle import make_golden_runner as mgr
    from oracle import refshim
    z = np.load(gu.golden_path(f"runner_{name}"), allow_pickle=False)
    method, hp, over = mgr.REAL_CASES[name]
    seed = mgr.REAL_SEEDS[name]
    net, net0 = mgr.RealNet(seed), mgr.RealNet(seed + 1)
    args = mgr.make_args(dict(hp, noise="torch", div="ieee", graph_train=graph_train), str(tmp_path), cuda_device, lr=2e-2,
                         lr_head=5e-2, **over)
    runner = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner(net, net0, args, _logger())
    evals = []
    orig = runner.evaluate

    def rec(loader):
        r = orig(loader)
        evals.append(r)
        return r
    runner.evaluate = rec
    bmas = []
    if hasattr(runner, "evaluate_full_samples"):
        orig_bma = runner.evaluate_full_samples

        def rec_bma(*a, **k):
            r = orig_bma(*a, **k)
            bmas.append(r)
            return r
        runner.evaluate_full_samples = rec_bma
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with refshim.injected_noise(z["tape"]) as tape:
            ret = runner.train(*mgr.loaders_from_arrays(z))
    finally:
        os.chdir(cwd)
    assert tape.pos == int(z["tape_used"])
    assert sum(isinstance(v, dict) for v in runner.model._train_graphs.values()) == (1 if graph_train else 0)
    theta = runner._dense(runner.model.chain.theta).cpu().numpy()
    assert gu.max_rel(theta, z["theta_final"]) <= 2e-4, gu.max_rel(theta, z["theta_final"])
    if "n_bma" in z.files:
        # raw-sample store on a network with BatchNorm: every stored state_dict carries the statistics of ITS epoch
        assert len(bmas) == int(z["n_bma"])
        files = z["bma_files"].tolist()
        assert runner._full_sample_files() == files
        names = [n for n, _ in runner.net.named_parameters()]
        for i, f in enumerate(files):
            sd = torch.load(os.path.join(tmp_path, f), map_location="cpu")
            assert list(sd.keys()) == z["fs_state_keys"].tolist()
            np.testing.assert_allclose(sd["features.1.running_mean"].numpy(), z[f"fs{i}_bn_mean"], rtol=1e-4, atol=1e-6)
            got = torch.cat([sd[n].reshape(-1) for n in names]).numpy()
            assert gu.max_rel(got, z[f"fs{i}_theta"]) <= 2e-4
            res = runner._fs_resident[f]           # the resident copy evaluates with the same buffers the file holds
            assert torch.equal(res.net.features[1].running_mean.cpu(), sd["features.1.running_mean"])
        assert not np.allclose(z["fs0_bn_mean"], z[f"fs{len(files) - 1}_bn_mean"])
        for ds in ("train", "val", "test"):
            r = bmas[-1][ds]
            assert r["num_models"] == int(z[f"bma_{ds}_num_models"])
            np.testing.assert_allclose(r["logits_all"], z[f"bma_{ds}_logits_all"], atol=5e-3, rtol=5e-3)
            np.testing.assert_allclose(r["logits"], z[f"bma_{ds}_logits"], atol=5e-3, rtol=5e-3)
            for k in ("loss", "individual_avg_loss"):
                assert abs(r[k] - float(z[f"bma_{ds}_{k}"])) <= 5e-3
    # BatchNorm running statistics are buffers: updated by the forward passes, never sampled (Appendix B.12)
    np.testing.assert_allclose(runner.net.features[1].running_mean.cpu().numpy(), z["bn_mean"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(runner.net.features[1].running_var.cpu().numpy(), z["bn_var"], rtol=1e-4, atol=1e-6)
    if "post_theta_mom1" in z.files:
        assert runner.post_theta_cnt == int(z["post_theta_cnt"])
        assert gu.max_rel(runner.post_theta_mom1.cpu().numpy(), z["post_theta_mom1"]) <= 2e-4
        assert gu.max_rel(runner.post_theta_mom2.cpu().numpy(), z["post_theta_mom2"]) <= 4e-4
    if "cycles" in z.files:
        assert sorted(runner.cycle_theta_mom1) == z["cycles"].tolist()
        for c in z["cycles"].tolist():
            assert runner.samples_per_cycle[c] == int(z[f"cyc{c}_count"])
            assert gu.max_rel(runner.cycle_theta_mom1[c].cpu().numpy(), z[f"cyc{c}_mom1"]) <= 2e-4
            np.testing.assert_allclose(runner.cycle_likelihoods[c], z[f"cyc{c}_lik"], rtol=2e-3)
        np.testing.assert_allclose(ret["losses_train"], z["losses_train"], rtol=1e-3, atol=1e-4)
    assert len(evals) == int(z["n_evals"])
    for i, (loss, err, targets, logits, _) in enumerate(evals):
        assert np.array_equal(targets, z[f"eval{i}_targets"])
        np.testing.assert_allclose(logits, z[f"eval{i}_logits"], atol=5e-3, rtol=5e-3)
        assert abs(loss - float(z[f"eval{i}_loss"])) <= 5e-3


@pytest.mark.parametrize("name", ["sghmc", "adam_sghmc", "sgld", "csghmc"])
def test_checkpoint_roundtrip(cuda_device, tmp_path, name):
    """save_ckpt -> load_ckpt restores what the reference restores (moments, prior_sig, sampler state dicts, optimizer
    state incl. the SGD momentum buffer living in the flat buffer) and keeps the reference's quirk
    post_theta_cnt = epoch (Appendix B.11)."""
    import importlib
    from oracle import make_golden_runner as mgr
    z, runner, _, _, _ = _run(name, cuda_device, tmp_path)
    method = mgr.CASES[name][0]
    cyclical = hasattr(runner, "_cyc1")
    ck_path = runner.save_ckpt(7)
    ck = torch.load(ck_path, map_location="cpu", weights_only=False)
    # a fresh runner of the same kind, state initialised by one step so the flat buffers exist
    seed = 500 + sorted(mgr.CASES).index(name)
    hp = dict(mgr.CASES[name][1], noise="philox")
    args = mgr.make_args(hp, str(tmp_path), cuda_device, **mgr.CASES[name][2])
    fresh = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner(mgr.InjectNet(seed, z["G"]), mgr.InjectNet(seed + 1),
                                                                              args, _logger())
    fresh.criterion = mgr.InjectCriterion()
    x, y = mgr.loaders_from_arrays(z)[0][0]
    fresh.net.train()
    fresh.model(x.to(cuda_device), y.to(cuda_device), fresh.net, fresh.net0, fresh.criterion,
                [pg["lr"] for pg in fresh.optimizer.param_groups], fresh.Ninflate, fresh.nd)
    epoch = fresh.load_ckpt(ck_path)
    assert epoch == 7
    if cyclical:
        assert sorted(fresh.cycle_theta_mom1) == sorted(runner.cycle_theta_mom1)
        for c in runner.cycle_theta_mom1:
            assert torch.equal(fresh.cycle_theta_mom1[c], runner.cycle_theta_mom1[c])
            assert torch.equal(fresh.cycle_theta_mom2[c], runner.cycle_theta_mom2[c])
        assert fresh.samples_per_cycle == runner.samples_per_cycle and fresh.current_cycle == runner.current_cycle
        return
    assert torch.equal(fresh.post_theta_mom1, runner.post_theta_mom1)
    assert torch.equal(fresh.post_theta_mom2, runner.post_theta_mom2)
    assert fresh.post_theta_cnt == 7                      # the reference's quirk: count overwritten by the epoch
    if "momentum_buffer" in ck:
        for k, v in runner.model.momentum_buffer.items():
            assert torch.equal(fresh.model.momentum_buffer[k], v)
    if "m" in ck:
        for k in runner.model.m:
            assert torch.equal(fresh.model.m[k], runner.model.m[k]) and torch.equal(fresh.model.v[k], runner.model.v[k])
        assert fresh.model.t == runner.model.t
    if runner.model.chain.buf is not None:                # SGD momentum buffer restored into the flat buffer
        assert torch.equal(fresh._dense(fresh.model.chain.buf), runner._dense(runner.model.chain.buf))   # padding is not part of a ckpt
        p0 = fresh.model.chain.params[0]
        assert fresh.optimizer.state[p0]["momentum_buffer"].data_ptr() == fresh.model.chain.layout.views(fresh.model.chain.buf)[0].data_ptr()


def test_baseline_cfg1_mlp_mnist_sgld_end_to_end(cuda_device, tmp_path):
    """BASELINE.json configs[0] -- mlp_mnist SGLD (prior_sig=1, Ninflate=1e3, nst=5, lr=1e-2, momentum=0.5, batch 128), the
    reference's own CPU-runnable case: the reference ran it on CPU (oracle/make_golden_runner.py cfg1, shortened burn-in),
    the drop-in runs it on the GPU with the same seeded noise stream.  Real autograd on both sides, so gradients differ
    at the 1e-7 level (CPU vs GPU matmul) and results are compared with a tolerance; 2.8 M parameters are compared
    through strided samples and fp64 sums."""
    from oracle import make_golden_runner as mgr
    from oracle import refshim
    from bayesdll_b200.methods import sgld
    z = np.load(gu.golden_path("runner_cfg1_mlp_sgld"), allow_pickle=False)
    net = mgr.cfg1_network()
    args = mgr.cfg1_args(str(tmp_path), cuda_device, extra_hp=dict(noise="torch", div="ieee"))
    runner = sgld.Runner(net, None, args, _logger())
    evals = []
    orig = runner.evaluate

    def rec(loader):
        r = orig(loader)
        evals.append(r)
        return r
    runner.evaluate = rec
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with refshim.injected_noise(mgr.CFG1["tape_seed"]) as tape:
            runner.train(*mgr.cfg1_loaders())
    finally:
        os.chdir(cwd)
    assert tape.pos == int(z["tape_used"]) and tape.calls == int(z["tape_calls"])
    assert runner.post_theta_cnt == int(z["post_theta_cnt"]) and len(evals) == int(z["n_evals"])
    got = {"theta": runner._dense(runner.model.chain.theta), "mom1": runner.post_theta_mom1, "mom2": runner.post_theta_mom2}
    for name, vec in got.items():
        s = mgr.summarize(vec.cpu().numpy())
        assert s["n"] == int(z[f"{name}_n"]) == 2797010
        assert gu.max_rel(s["sample"], z[f"{name}_sample"]) <= 2e-4, (name, gu.max_rel(s["sample"], z[f"{name}_sample"]))
        assert abs(s["sumsq"] - float(z[f"{name}_sumsq"])) <= 1e-5 * float(z[f"{name}_sumsq"]), name
        assert abs(s["sum"] - float(z[f"{name}_sum"])) <= 1e-4 * np.sqrt(float(z[f"{name}_sumsq"]) * s["n"]), name
    for i, (loss, err, targets, logits, _) in enumerate(evals):
        assert np.array_equal(targets, z[f"eval{i}_targets"])
        np.testing.assert_allclose(logits, z[f"eval{i}_logits"], atol=5e-3, rtol=5e-3)
        assert abs(loss - float(z[f"eval{i}_loss"])) <= 5e-3
    assert os.path.exists(os.path.join(tmp_path, "ckpt.pt")) and os.path.exists(os.path.join(tmp_path, "logits_test.pkl"))


@pytest.mark.parametrize("name", ["sghmc", "csgld", "csghmc"])
def test_graphed_evaluation_forward_is_bit_identical_to_eager(cuda_device, tmp_path, name):
    """evaluate() / full_batch_likelihoods replay the evaluation net's forward as a CUDA graph (graphfwd.py); hparams
    graph=0 runs it eagerly.  Same kernels either way: every output must be bit-identical, and the graph must really
    have been used."""
    import importlib
    from oracle import make_golden_runner as mgr
    from bayesdll_b200.graphfwd import GraphedForward
    z = np.load(gu.golden_path(f"runner_{name}"), allow_pickle=False)
    method, hp, over = mgr.CASES[name]
    seed = 500 + sorted(mgr.CASES).index(name)
    results = {}
    for graph in (1, 0):
        net, net0 = mgr.InjectNet(seed, z["G"]), mgr.InjectNet(seed + 1)
        d = tmp_path / f"g{graph}"
        d.mkdir()
        args = mgr.make_args(dict(hp, graph=graph), str(d), cuda_device, **over)
        runner = importlib.import_module(f"bayesdll_b200.methods.{method}").Runner(net, net0, args, _logger())
        runner.criterion = mgr.InjectCriterion()
        loaders = mgr.loaders_from_arrays(z)
        before = (GraphedForward.total_captures, GraphedForward.total_replays)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            runner.train(*loaders)
        finally:
            os.chdir(cwd)
        out = runner.evaluate(loaders[2])
        used = (GraphedForward.total_captures - before[0], GraphedForward.total_replays - before[1])
        results[graph] = (out, getattr(runner, "cycle_likelihoods", None), used)
    (o1, l1, used1), (o0, l0, used0) = results[1], results[0]
    assert used0 == (0, 0) and used1[0] >= 1 and used1[1] >= runner.nst
    assert o1[0] == o0[0] and o1[1] == o0[1]
    for a, b in zip(o1[2:], o0[2:]):
        assert np.array_equal(np.asarray(a), np.asarray(b)) and np.asarray(a).dtype == np.asarray(b).dtype
    if l1 is not None:                                      # cycle -> likelihoods of the nst full-batch passes
        assert sorted(l1) == sorted(l0)
        for c in l1:
            assert np.array_equal(np.asarray(l1[c], np.float64), np.asarray(l0[c], np.float64))


class _ConvBN(torch.nn.Module):
    """A real autograd network with BatchNorm buffers and dropout (host-side control flow free: capturable)."""
    readout_name = "classifier"

    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.features = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                                            torch.nn.Conv2d(8, 12, 3, stride=2, padding=1), torch.nn.BatchNorm2d(12),
                                            torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten())
        self.drop = torch.nn.Dropout(0.1)
        self.classifier = torch.nn.Linear(12, 7)

    def forward(self, x):
        return self.classifier(self.drop(self.features(x)))


@pytest.mark.parametrize("method", ["sghmc", "adam_csghmc", "sgld"])
def test_graphed_training_step_is_bit_identical_to_eager(cuda_device, method, request):
    """hparams graph_train=1: forward + loss + backward replayed as one CUDA graph, the fused sampler step launched
    after it.  Parameters, sampler state, BatchNorm buffers, loss and logits must equal the eager run bit for bit at
    every step -- including an odd-sized batch in between (runs eagerly, then the graph resumes)."""
    import importlib
    from bayesdll_b200 import _lib
    mod = importlib.import_module(f"bayesdll_b200.methods.{method}")
    gen = torch.Generator().manual_seed(5)
    batches = [(torch.randn(6 if i == 5 else 16, 3, 12, 12, generator=gen), torch.randint(0, 7, (6 if i == 5 else 16,), generator=gen))
               for i in range(9)]
    runs = {}
    # cuDNN's default backward kernels for these tiny convolutions use atomics: two EAGER runs already differ in the last
    # bits of the gradient (1e-10; Adam's normalisation then amplifies it), so the comparison pins deterministic kernels
    monkey = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    request.addfinalizer(lambda: setattr(torch.backends.cudnn, "deterministic", monkey))
    for graph in (True, False):
        net, net0 = _ConvBN(1).to(cuda_device).train(), _ConvBN(2).to(cuda_device)
        model = mod.Model(ND=500, prior_sig=1.0, bias="informative")
        model.configure(seed=11, graph_train=graph, sgd_momentum=0.5 if method == "sgld" else 0.0)
        crit = torch.nn.CrossEntropyLoss()
        torch.manual_seed(123)                                      # dropout masks come from torch's CUDA generator
        trace = []
        for x, y in batches:
            loss, out = model.forward(x.to(cuda_device), y.to(cuda_device), net, net0, crit, [1e-3, 1e-2], Ninflate=10.0, nd=1.0,
                                      **({"should_sample": True} if method == "adam_csghmc" else {}))
            ch = model.chain
            state = [ch.theta.clone()] + [t.clone() for t in (ch.v, ch.m, ch.s, ch.buf) if t is not None]
            state += [b.clone() for b in net.buffers()]
            trace.append((loss, out.clone(), state))
        runs[graph] = (trace, dict(model._train_graphs))
    (tg, used), (te, unused) = runs[True], runs[False]
    assert unused == {} and sum(isinstance(v, dict) for v in used.values()) == 1     # one captured graph (the 16-row batches)
    for (l1, o1, s1), (l0, o0, s0) in zip(tg, te):
        assert l1 == l0 and torch.equal(o1, o0)
        assert len(s1) == len(s0) and all(torch.equal(a, b) for a, b in zip(s1, s0))
