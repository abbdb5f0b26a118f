"""Error conventions and degenerate inputs of the C ABI (SURVEY.md section 8b "Error conventions"; include/bdl.h):
integer status, thread-local message, no exception across the boundary, empty inputs are no-ops, the Python front end
rejects CPU tensors / wrong dtypes before the call."""
import ctypes as C

import numpy as np
import pytest
import torch

from bayesdll_b200 import _lib, ops
from bayesdll_b200.flat import FlatLayout

pytestmark = pytest.mark.gpu

OK, INVALID, ALIGN, CUDA, UNSUPPORTED = 0, -1, -2, -3, -4


def _msg():
    return _lib.load().bdl_last_error().decode()


def _state(n, dev):
    return [torch.zeros(n, device=dev) for _ in range(4)]


def _runs(n):
    lay = FlatLayout([("w", (n,))], "head")
    return lay.run_table("informative")


def test_step_argument_errors(cuda_device):
    lib = _lib.load()
    n = 1024
    th, g, th0, v = _state(n, cuda_device)
    tab = _runs(n)
    rd, nr = ops.upload_runs(tab, cuda_device)
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=10)
    nz = ops.make_noise(seed=1)
    st = torch.cuda.current_stream().cuda_stream

    def call(variant=_lib.SGHMC, theta=th.data_ptr(), grad=g.data_ptr(), theta0=th0.data_ptr(), mom=v.data_ptr(), nn=n,
             runs=rd.data_ptr(), nruns=nr, scal=sc):
        return lib.bdl_step(variant, theta, grad, theta0, mom, None, None, None, nn, runs, nruns, None, C.byref(scal),
                            C.byref(nz), st)

    assert call() == OK
    assert call(variant=17) == INVALID and "variant" in _msg()
    assert call(theta=None) == INVALID
    assert call(theta0=None) == INVALID and "theta0" in _msg()
    assert call(mom=None) == INVALID and "momentum" in _msg()
    assert call(nn=n - 2) == INVALID and "multiple of 4" in _msg()
    assert call(theta=th.data_ptr() + 4) == ALIGN and "aligned" in _msg()
    assert call(nruns=0) == INVALID and call(nruns=_lib.MAX_RUNS + 1) == INVALID
    assert call(variant=_lib.ADAM_SGHMC) == INVALID and "Adam" in _msg()          # m, s missing
    bad = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=10)
    bad.div_mode = 7
    assert call(scal=bad) == INVALID
    # the state was not touched by any rejected call, and the library is still usable
    torch.cuda.synchronize()
    assert call() == OK
    torch.cuda.synchronize()
    assert torch.isfinite(th).all()


def test_empty_inputs_are_noops(cuda_device):
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=10)
    nz = ops.make_noise(seed=1)
    assert lib.bdl_step(_lib.SGHMC, None, None, None, None, None, None, None, 0, None, 0, None, C.byref(sc), C.byref(nz),
                        st) == OK
    e = torch.zeros(0, device=cuda_device)
    ops.moments_avg(e, e, e, 0, init=True)
    ops.moments_welford(e, e, e, 1, init=True)
    ops.philox_normal(e, seed=3)
    ops.draw(e, e, e, ops.VAR_FROM_MOMENTS, 1.0, ops.make_noise(seed=1))
    ring = torch.zeros((1, 0), device=cuda_device)
    ops.capture_ring(e, ring, 0)
    # zero-row batches
    la = torch.zeros((0, 5, 3), device=cuda_device)
    out = torch.zeros((0, 5), device=cuda_device)
    ops.ensemble(la, out, 3)
    ops.bma_mean(la, out)
    loss = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    err = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    ops.ce_err(out, torch.zeros(0, dtype=torch.int64, device=cuda_device), loss, err)
    edges = torch.linspace(0, 1, 16, dtype=torch.float64, device=cuda_device)[1:]
    size, acc, conf, nll, near, _ = ops.calibrate(out, torch.zeros(0, dtype=torch.int64, device=cuda_device), edges)
    torch.cuda.synchronize()
    assert loss.item() == 0 and err.item() == 0 and size.sum().item() == 0 and nll.item() == 0


def test_predict_argument_errors(cuda_device):
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    la = torch.zeros((2, 3, 4), device=cuda_device)
    out = torch.zeros((2, 3), device=cuda_device)
    assert lib.bdl_ensemble(la.data_ptr(), 2, 3, 4, 0.0, 1.0, 5, out.data_ptr(), st) == INVALID and "mode" in _msg()
    assert lib.bdl_ensemble(la.data_ptr(), 2, 0, 4, 0.0, 1.0, 0, out.data_ptr(), st) == INVALID
    assert lib.bdl_ensemble(None, 2, 3, 4, 0.0, 1.0, 0, out.data_ptr(), st) == INVALID
    assert lib.bdl_ensemble(la.data_ptr(), 2, 3, 100000, 0.0, 1.0, 0, out.data_ptr(), st) == UNSUPPORTED
    assert lib.bdl_bma_mean(la.data_ptr(), 2, 3, 0, out.data_ptr(), st) == INVALID
    lg = torch.zeros((4, 3), device=cuda_device)
    lb = torch.zeros(4, dtype=torch.int64, device=cuda_device)
    edges = torch.linspace(0, 1, 16, dtype=torch.float64, device=cuda_device)[1:].contiguous()
    stats = torch.zeros(64, dtype=torch.float64, device=cuda_device)
    p = stats.data_ptr()
    assert lib.bdl_calibrate(lg.data_ptr(), lb.data_ptr(), 4, 3, 0.0, 0, edges.data_ptr(), 15, p, p + 128, p + 256, p + 384,
                             None, None, st) == INVALID and "temperature" in _msg()
    assert lib.bdl_calibrate(lg.data_ptr(), lb.data_ptr(), 4, 3, 1.0, 0, edges.data_ptr(), 4096, p, p + 128, p + 256, p + 384,
                             None, None, st) == UNSUPPORTED
    row = torch.zeros(4, dtype=torch.float64, device=cuda_device)
    assert lib.bdl_nll_temperature(lg.data_ptr(), lb.data_ptr(), 4, 3, 0.0, row.data_ptr(), p, st) == INVALID
    assert lib.bdl_nll_temperature(lg.data_ptr(), lb.data_ptr(), 4, 3, float("nan"), row.data_ptr(), p, st) == INVALID
    assert lib.bdl_nll_temperature(lg.data_ptr(), lb.data_ptr(), 0, 3, 1.0, row.data_ptr(), p, st) == INVALID


def test_python_front_end_rejects_cpu_and_wrong_dtype(cuda_device):
    cpu = torch.zeros(8)
    with pytest.raises(ops.BdlError, match="CUDA tensor"):
        ops.moments_avg(cpu, cpu, cpu, 0, init=True)
    dev64 = torch.zeros(8, dtype=torch.float64, device=cuda_device)
    with pytest.raises(ops.BdlError, match="dtype"):
        ops.philox_normal(dev64, seed=1)
    nc = torch.zeros((8, 8), device=cuda_device).t()[:, :4]
    with pytest.raises(ops.BdlError, match="contiguous"):
        ops.philox_normal(nc, seed=1)
    th = torch.zeros(8, device=cuda_device)
    with pytest.raises(ops.BdlError, match="length"):
        ops.step(_lib.SGHMC, th, torch.zeros(4, device=cuda_device), th, th, None, None, None, *ops.upload_runs(_runs(8), cuda_device),
                 ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=10), ops.make_noise(seed=1))
    with pytest.raises(ops.BdlError, match="ring"):
        ops.capture_ring(th, torch.zeros((2, 4), device=cuda_device), 0)
    from bayesdll_b200.chain import ChainState
    import torch.nn as nn
    net = nn.Linear(3, 2)
    net.readout_name = "weight"
    with pytest.raises(_lib.BdlError, match="CUDA devices only"):
        ChainState(net, net, variant=_lib.SGHMC, bias_mode="informative")


def test_host_chain_argument_errors(cuda_device):
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.bdl_chain_create(6, _lib.SGHMC, 0, 0, C.byref(h)) == INVALID          # n not a multiple of 4
    assert lib.bdl_chain_create(8, 99, 0, 0, C.byref(h)) == INVALID
    ch = ops.HostChain(1024, _lib.SGHMC)
    pageable = torch.zeros(1024)
    with pytest.raises(ops.BdlError, match="pinned"):
        ch.step_host(pageable, pageable, _runs(1024), ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=10),
                     ops.make_noise(seed=1))
    with pytest.raises(ops.BdlError, match="contiguous fp32 CPU tensor"):
        ch.upload(_lib.BUF_THETA, torch.zeros(512))
    assert lib.bdl_chain_upload(h if h else None, 0, None) == INVALID
    ch.close()
    ch.close()                                                                        # idempotent


def test_ops_follow_the_tensor_device_not_the_current_one():
    """Tensors on cuda:1 while cuda:0 is current (args.device='cuda:1' without torch.cuda.set_device): every entry point
    must launch on the tensors' device.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    torch.cuda.set_device(0)
    n = 4096
    res = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        gen = torch.Generator().manual_seed(5)
        th, g, th0, v = [(torch.randn(n, generator=gen) * 0.1).to(dev) for _ in range(4)]
        rd, nr = ops.upload_runs(_runs(n), dev)
        sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-3, lr_head=1e-2, ND=10)
        m1, m2 = torch.empty(n, device=dev), torch.empty(n, device=dev)
        ops.step(_lib.SGHMC, th, g, th0, v, None, None, None, rd, nr, sc, ops.make_noise(seed=3, subseq=1),
                 capture=ops.make_capture("avg", m1, m2, 0, init=True))
        out = torch.empty(n, device=dev)
        ops.draw(m1, m2, out, ops.VAR_FROM_MOMENTS, 1.5, ops.make_noise(seed=3, subseq=2, stream_id=_lib.STREAM_DRAW))
        ring = torch.zeros((2, n), device=dev)
        ops.capture_ring(out, ring, 1)
        la = torch.randn(8, 5, 3, generator=gen).to(dev)
        ens = torch.empty(8, 5, device=dev)
        ops.ensemble(la, ens, 3)
        torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0
        res.append([t.cpu() for t in (th, v, m1, m2, ring[1], ens)])
    for a, b in zip(*res):
        assert torch.equal(a, b)
