#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native BayesDLL sampler hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): fused SGHMC update of a ViT-L/32
chain (37-class head, n = 305 548 325 fp32 parameters, net0 prior mean, in-kernel Philox noise), hyper-parameters
``prior_sig=1.0,Ninflate=1e3,nd=1.0,momentum_decay=0.18,bias=informative`` with lr 1e-4 / lr_head 1e-2, ND=1840.
A "step" is one pass of the fused update over the whole flat state (theta, g, theta0, v resident in HBM; 4.9 GB of
inputs per step >> 126 MB L2, so no L2 flush is needed between iterations).  One independent chain per GPU
(weak scaling, no collective on the data path -- SURVEY.md section 8e).

value       params/s over all GPUs, state resident in HBM, CUDA events on the launching stream, max over ranks.
e2e         the same update through the host-buffer C ABI (bdl_chain_step_host): the gradient comes from pinned HOST
            memory and the new theta returns to pinned HOST memory every step (4 B/param each way, inside the timed
            region); theta0 / momentum stay resident.  This is what a CPU-resident caller of the reference binds.
roofline    HBM-bound; achieved = 24 B/param * n / average kernel duration; peak = MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  the UNMODIFIED reference's own methods/sghmc.py Model.forward + optimizer.step() on the host cores (kind
            "reference"; baseline/reference_arm.py drives it from /root/reference or baseline/_ref), bounded sample, rank 0,
            N=1; the fused C/OpenMP port (oracle/bdl_oracle.c, kind "port") is reported next to it and is the fallback
            when no reference tree exists.
train_step  (extra) the full user call Model.forward on a real torchvision ViT-L/32, batch 64: x,y from pinned host,
            PyTorch fwd/bwd, fused update reading autograd's gradients in place, loss.item().
ensemble    BASELINE.json configs[4] through the Runner API: csgld.Runner.evaluate() (hparams eval_shard=1) on ResNet-101,
            8 cycles x nst 5 = 40 samples, N = 3 669 rows (Pets test size), batch 64, + ECE/MCE/NLL; preds/s and a
            per-phase device-time split (draw | forward | reduce | exchange | d2h | calibrate).  Also mirrored into
            e2e["ensemble_*"].
cfg1        BASELINE.json configs[0]: mlp_mnist SGLD, the unmodified reference Runner on the host cores next to the
            drop-in Runner on the GPU (ms/step, ensemble preds/s, calibration.analyze ms).

--impl reference times the reference's own CPU implementation of the path (kind "reference"; the C port when no reference
tree is found) on the same config / metric / unit (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")                  # before anything imports tqdm: the Runners' progress bars would flood stderr

import numpy as np  # noqa: E402
import torch  # noqa: E402

BYTES_PER_PARAM = 24          # SGHMC: R theta,g,theta0,v ; W v,theta   
HP = dict(prior_sig=1.0, Ninflate=1e3, nd=1.0, alpha=0.18, lr_body=1e-4, lr_head=1e-2, ND=1840)
METRIC = "sampler step params/s (ViT-L/32 fused SGHMC update)"
WORKLOAD = "ViT-L/32 (K=37, n=305548325) SGHMC fused step, net0 prior mean, in-kernel Philox; BASELINE.json configs[2]"


def bench_config(world, n_dense):
    """The `config` object of the JSON line -- identical in both arms (`--impl ours|reference`)."""
    return {"workload": WORKLOAD, "hparams": "prior_sig=1.0,Ninflate=1e3,nd=1.0,momentum_decay=0.18,bias=informative",
            "lr": HP["lr_body"], "lr_head": HP["lr_head"], "ND": HP["ND"], "params_per_chain": n_dense,
            "chains": world, "parallelism": f"{world} independent chain(s), one per GPU, no data-path collective",
            "l2": "inputs (4.9 GB per step) larger than L2 (126 MB); no flush needed",
            "noise": "in-kernel Philox4x32-10 + Box-Muller", "division": "reciprocal (reference-on-CUDA semantics)"}


def env_int(name, default):
    return int(os.environ.get(name, default))


def guarded(what, fn, *a, **k):
    """Extras must never take the headline line down with them: report the failure instead."""
    try:
        return fn(*a, **k)
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {"error": f"{what}: {type(e).__name__}: {e}"[:300]}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel from the committed ``ncu --set full``
    capture -- only when that capture was taken from THIS build of the kernels (profiles/traffic.json records a hash of
    csrc/ + include/bdl.h + the nvcc flags; tools/traffic_from_ncu.py writes it).  -> (bytes | None, note)."""
    try:
        from bayesdll_b200 import build as b
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)
        have = b.source_hash()
        if rec.get("build_hash") != have:
            return None, (f"profiles/traffic.json was captured from build {str(rec.get('build_hash'))[:12]}, the loaded "
                          f"library is build {have[:12]}: re-run tools/round_end_run.sh")
        return rec.get("sghmc_step_dram_bytes_per_launch"), f"ncu capture of build {have[:12]} ({rec.get('source', '')[:80]})"
    except Exception as e:  # noqa: BLE001
        return None, f"no usable profiles/traffic.json ({type(e).__name__})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load (B200_PROFILING.md recipe)."""
    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, cmax = float(parts[2]), float(parts[3])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                mx.append(cmax)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def build_layout():
    from bayesdll_b200 import shapes
    from bayesdll_b200.flat import FlatLayout
    named, readout = shapes.named_shapes("vit_l_32", 37)
    return FlatLayout(named, readout)


def make_scalars(variant):
    from bayesdll_b200 import _lib, ops
    return ops.make_scalars(variant, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                            prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=HP["alpha"], div_mode=_lib.DIV_RECIP)


def synth_state(n, device, seed):
    """Random-init weights of the named shapes, synthetic gradients g ~ N(0, 1e-2^2)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    theta = torch.randn(n, device=device, generator=gen) * 0.02
    theta0 = torch.randn(n, device=device, generator=gen) * 0.02
    g = torch.randn(n, device=device, generator=gen) * 0.01
    v = torch.zeros(n, device=device)
    return theta, g, theta0, v


def allmax(x, world, device):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    from bayesdll_b200 import _lib, ops
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    _lib.load()
    lay = build_layout()
    n, n_dense = lay.n_padded, lay.n_dense
    theta, g, theta0, v = synth_state(n, device, 42 + rank)
    runs_dev, nruns = ops.upload_runs(lay.run_table("informative"), device)
    sc = make_scalars(_lib.SGHMC)
    seed = 42 + rank
    K, W = args.steps, max(args.warmup, 3)

    def step(i):
        ops.step(_lib.SGHMC, theta, g, theta0, v, None, None, None, runs_dev, nruns, sc,
                 ops.make_noise(seed=seed, subseq=i, stream_id=_lib.STREAM_STEP))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(W):
        step(i)
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.time()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    t_region1 = time.time()
    barrier(world)
    ms_total = allmax(e0.elapsed_time(e1), world, device)
    ms_per_step = ms_total / K
    value = world * n_dense * K / (ms_total * 1e-3)
    # keep the sampler's region long enough for >= a few nvidia-smi samples
    t_extra = time.time()
    while rank == 0 and time.time() - t_extra < 0.4:
        step(0)
        torch.cuda.synchronize()
    t_region1b = time.time()
    assert torch.isfinite(theta[:: max(1, n // 4096)]).all(), "state diverged"

    # ---- roofline of the dominant (only) kernel -------------------------------------------------
    peak, peak_kind = measured_peak()
    kernel_ms = e0.elapsed_time(e1) / K                      # this rank's average launch duration
    achieved = BYTES_PER_PARAM * n_dense / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_note = recorded_traffic()
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_kind,
                "frac_of_nominal_8TBps": round(achieved / 8000.0, 4), "kernel": "bdl::step_kernel<SGHMC,philox,recip,U=1,T=64>, one tile per CTA",
                "algorithmic_bytes_per_launch": BYTES_PER_PARAM * n_dense, "kernel_ms": round(kernel_ms, 4)}

    # the bare-traffic yardstick: a kernel that only MOVES the step's bytes (4 reads + 2 writes per element, next to no
    # arithmetic, same launch shape; bdl_probe_stream) on scratch copies of the same size, timed right after the step
    bare = guarded("bare traffic", bare_traffic_ms, theta, g, theta0, device, reps=min(K, 50))
    if isinstance(bare, dict) and "ms" in bare:
        roofline["bare_traffic_kernel"] = {"ms": round(bare["ms"], 4), "gbs": round(BYTES_PER_PARAM * n_dense / (bare["ms"] * 1e-3) / 1e9, 1),
                                           "step_over_bare": round(kernel_ms / bare["ms"], 4),
                                           "note": "bdl_probe_stream(4 reads, 2 writes): the step's traffic without its arithmetic"}

    # ---- the other update rules on the same state (extra; BASELINE.json configs[1], [3], [4] kernels) ----------
    variants = {} if args.no_variants else guarded("variants", variant_rates, lay, theta, g, theta0, v, runs_dev, nruns, device,
                                                   world, peak, seed)

    # ---- e2e through the host-buffer C ABI ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = guarded("e2e", e2e_host_chain, lay, theta, theta0, g, sc, seed, K, args, world, device)
    t_load_end = time.time()
    if rank == 0:
        sampler.stop()
    clocks = sampler.summary(t_region0, t_region1b) if rank == 0 else None

    # ---- extras --------------------------------------------------------------------------------------------
    extras = {}
    if rank == 0 and not args.no_sample_store:
        extras["sample_store"] = guarded("sample_store", sample_store_extra, lay, theta, step, device, peak)
    if not args.no_train_step:
        del theta, g, theta0, v
        torch.cuda.empty_cache()
        extras["train_step"] = guarded("train_step", train_step_extra, device, rank, world)
        torch.cuda.empty_cache()
        extras["train_step_resnet101"] = guarded("train_step_resnet101", train_step_extra, device, rank, world, steps=10,
                                                 batch=16, backbone="resnet101")
        torch.cuda.empty_cache()
        extras["train_step_resnet101_csghmc"] = guarded("train_step_resnet101_csghmc", train_step_extra, device, rank, world,
                                                        steps=10, batch=16, backbone="resnet101", method="csghmc")   # configs[1]
    if not args.no_ensemble:
        torch.cuda.empty_cache()
        extras["ensemble"] = guarded("ensemble", ensemble_extra, device, rank, world, args)
    if rank == 0 and world == 1 and not args.no_eager_gpu:
        torch.cuda.empty_cache()
        extras["reference_eager_gpu"] = guarded("reference_eager_gpu", eager_gpu_rate, lay, device)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        port = guarded("cpu_baseline", cpu_port_rate, lay, seconds=args.cpu_seconds, with_eager=True)
        ref = guarded("cpu_baseline(reference)", reference_cpu_rate, lay, args.cpu_seconds, 3, 1)
        if ref and "error" not in ref:                       # the reference's own code is the baseline; the port is context
            cpu_baseline = dict(ref, fused_c_port=port)
        else:
            cpu_baseline = dict(port, reference_error=ref)
    if rank == 0 and world == 1 and not args.no_cfg1:
        extras["cfg1_mlp_mnist_sgld"] = {"reference_cpu": guarded("cfg1 reference", cfg1_line, True, "cpu"),
                                         "ours_b200": guarded("cfg1 ours", cfg1_line, False, device, True),
                                         "ours_b200_eager": guarded("cfg1 ours (graph_train=0)", cfg1_line, False, device, False)}

    ens = extras.get("ensemble")
    if isinstance(e2e, dict) and isinstance(ens, dict) and "value" in ens:
        # BASELINE.json's metric has a second half ("ensemble preds/s at 1/2/4/8 GPU"): the same end-to-end figure (host
        # batches in, reference 5-tuple + calibration out) rides in e2e so that the per-N record carries it
        e2e.update({"ensemble_preds_per_s": ens["value"], "ensemble_rows": ens["rows"], "ensemble_samples": ens["samples"],
                    "ensemble_seconds": ens["seconds"], "ensemble_phases_rank0_ms": ens["phases_rank0_ms"],
                    "ensemble_h2d_bytes": ens["h2d_bytes"], "ensemble_d2h_bytes": ens["d2h_bytes"]})
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "params/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (random-init weights of the named shapes, g ~ N(0,1e-2^2))",
            "config": bench_config(world, n_dense),
            "hbm_gbs": achieved * 1.0, "roofline": roofline, "e2e": e2e, "gpu_launches": K, "clocks": clocks,
            "cpu_baseline": cpu_baseline,
        }
        line["variants"] = variants
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
def bare_traffic_ms(theta, g, theta0, device, reps=50):
    """Average duration of bdl_probe_stream(4 reads, 2 writes) over buffers of the state's size (two scratch buffers stand
    in for theta and v so the chain's state is left alone)."""
    from bayesdll_b200 import ops
    a, b = torch.zeros_like(theta), torch.zeros_like(theta)
    for _ in range(3):
        ops.probe_stream(a, b, g, theta0, 4, 2, threads=64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.probe_stream(a, b, g, theta0, 4, 2, threads=64)
    e1.record()
    torch.cuda.synchronize()
    return {"ms": e0.elapsed_time(e1) / reps}


def e2e_host_chain(lay, theta, theta0, g, sc, seed, K, args, world, device):
    """The same update through the host-buffer C ABI: pinned host gradient in, pinned host theta out, every step."""
    from bayesdll_b200 import _lib, ops
    n, n_dense = lay.n_padded, lay.n_dense
    tab = lay.run_table("informative")
    chain = ops.HostChain(n, _lib.SGHMC)
    stage = torch.empty(n, dtype=torch.float32).pin_memory()
    stage.copy_(theta)
    chain.upload(_lib.BUF_THETA, stage)
    stage.copy_(theta0)
    chain.upload(_lib.BUF_THETA0, stage)
    g_host = stage                                       # pinned host gradient (synthetic)
    g_host.copy_(g)
    theta_host = torch.empty(n, dtype=torch.float32).pin_memory()
    Ke = max(3, min(K, args.e2e_steps))
    for i in range(2):
        chain.step_host(g_host, theta_host, tab, sc, ops.make_noise(seed=seed, subseq=i))
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(Ke):
        chain.step_host(g_host, theta_host, tab, sc, ops.make_noise(seed=seed, subseq=2 + i))
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0, world, device)
    assert np.isfinite(theta_host[:1024].numpy()).all()
    chain.close()
    return {"value": world * n_dense * Ke / dt, "unit": "params/s", "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n,
            "steps": Ke, "ms_per_step": dt / Ke * 1e3,
            "api": "bdl_chain_step_host (include/bdl.h): pinned host gradient in, pinned host theta out, "
                   "theta0/momentum resident in HBM, chunk-pipelined H2D | fused step | D2H"}


def variant_rates(lay, theta, g, theta0, v, runs_dev, nruns, device, world, peak, seed, steps=30):
    """params/s and roofline fraction of every other fused update rule at ViT-L/32 size (one chain per GPU)."""
    from bayesdll_b200 import _lib, ops
    n, n_dense = lay.n_padded, lay.n_dense
    m = torch.zeros(n, device=device)
    s2 = torch.full((n,), 1e-6, device=device)
    buf = torch.zeros(n, device=device)
    table = [("sgld_mu0.5", _lib.SGLD, 24, 0.5), ("sgld_mu0", _lib.SGLD, 16, 0.0), ("csghmc", _lib.CSGHMC, 20, 0.0),
             ("adam_sghmc_mu0.5", _lib.ADAM_SGHMC, 48, 0.5), ("adam_csghmc", _lib.ADAM_CSGHMC, 40, 0.0)]
    out = {}
    for name, variant, bpp, mu in table:
        adam = variant in (_lib.ADAM_SGHMC, _lib.ADAM_CSGHMC)
        sc = ops.make_scalars(variant, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                              prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=0.05 if adam else HP["alpha"], mu=mu, t=10,
                              div_mode=_lib.DIV_RECIP)

        def one(i):
            ops.step(variant, theta, g, None if variant == _lib.CSGHMC else theta0, None if variant == _lib.SGLD else v,
                     m if adam else None, s2 if adam else None, buf if mu else None, runs_dev, nruns, sc,
                     ops.make_noise(seed=seed, subseq=1000 + i))
        for i in range(3):
            one(i)
        barrier(world)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(3 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = allmax(e0.elapsed_time(e1), world, device) / steps
        gbs = bpp * n_dense / (ms * 1e-3) / 1e9
        out[name] = {"params_per_s": world * n_dense / (ms * 1e-3), "ms_per_step": ms, "bytes_per_param": bpp,
                     "achieved_gbs_per_gpu": gbs, "frac_of_measured_peak": gbs / peak}
    # post-burn-in steady state of BASELINE.json configs[2] (thin = 1): every step is followed by a moment capture.
    # Fused (bdl_step_capture, 40 B/param) vs the two launches it replaces (24 + 20 = 44 B/param).
    sc = make_scalars(_lib.SGHMC)
    mom1, mom2 = m, s2

    def timed(fn):
        for i in range(3):
            fn(i)
        barrier(world)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(3 + i)
        b.record()
        torch.cuda.synchronize()
        return allmax(a.elapsed_time(b), world, device) / steps

    def fused(i):
        ops.step(_lib.SGHMC, theta, g, theta0, v, None, None, None, runs_dev, nruns, sc,
                 ops.make_noise(seed=seed, subseq=3000 + i), capture=ops.make_capture("avg", mom1, mom2, 7 + i))

    def separate(i):
        ops.step(_lib.SGHMC, theta, g, theta0, v, None, None, None, runs_dev, nruns, sc, ops.make_noise(seed=seed, subseq=3000 + i))
        ops.moments_avg(theta, mom1, mom2, 7 + i)
    ms_f, ms_s = timed(fused), timed(separate)
    gbs = 40 * n_dense / (ms_f * 1e-3) / 1e9
    out["sghmc_step_with_fused_moment_capture"] = {
        "params_per_s": world * n_dense / (ms_f * 1e-3), "ms_per_step": ms_f, "bytes_per_param": 40,
        "achieved_gbs_per_gpu": gbs, "frac_of_measured_peak": gbs / peak, "separate_launches_ms": ms_s,
        "speedup_vs_separate": ms_s / ms_f,
        "note": "bdl_step_capture: step + running-moment update of the new theta in one pass (thin=1 steady state)"}
    # streaming draws (12 B/param each: two reads, one write): the posterior draw of evaluate() (a9) and the per-step
    # reparameterisation draws of the VI / MC-Dropout families
    dr_tab, dr_n = ops.upload_runs(lay.dropout_run_table("gaussian"), device)
    draws = [("posterior_draw", lambda i: ops.draw(mom1, mom2, buf, ops.VAR_FROM_MOMENTS, 1.25,
                                                   ops.make_noise(seed=seed, subseq=4000 + i, stream_id=_lib.STREAM_DRAW))),
             ("vi_reparam_draw", lambda i: ops.draw(mom1, mom2, buf, ops.STD_GIVEN, 1.0,
                                                    ops.make_noise(seed=seed, subseq=5000 + i, stream_id=_lib.STREAM_DRAW))),
             ("mc_dropout_mix_bias_gaussian", lambda i: ops.dropout_mix(mom1, theta0, buf, 0.1, ops.make_noise(
                 seed=seed, subseq=6000 + i, stream_id=_lib.STREAM_DRAW), dr_tab, dr_n)),
             ("mc_dropout_mix_spikymix", lambda i: ops.dropout_mix(mom1, theta0, buf, 0.1, ops.make_noise(
                 seed=seed, subseq=7000 + i, stream_id=_lib.STREAM_DRAW)))]
    for name, fn in draws:
        ms = timed(fn)
        gbs = 12 * n_dense / (ms * 1e-3) / 1e9
        out[name] = {"params_per_s": world * n_dense / (ms * 1e-3), "ms_per_launch": ms, "bytes_per_param": 12,
                     "achieved_gbs_per_gpu": gbs, "frac_of_measured_peak": gbs / peak}
    out["mc_dropout_mix_bias_gaussian"]["runs"] = dr_n
    # the draws' yardstick: 2 reads + 1 write per element with next to no arithmetic (bdl_probe_stream), same buffers
    ms = timed(lambda i: ops.probe_stream(buf, None, mom1, mom2, 2, 1, threads=128))
    out["bare_traffic_2r1w"] = {"ms_per_launch": ms, "bytes_per_param": 12, "achieved_gbs_per_gpu": 12 * n_dense / (ms * 1e-3) / 1e9,
                                "frac_of_measured_peak": 12 * n_dense / (ms * 1e-3) / 1e9 / peak,
                                "note": "bdl_probe_stream(2 reads, 1 write): what the draws' traffic costs without Philox / Box-Muller / sqrt"}
    # the shape Runner.train() actually launches: one run per tensor, each row pointing at that tensor's own
    # (separately allocated) autograd gradient -- no flat gradient buffer, no gather pass (step_table_kernel)
    del buf
    grads = [torch.randn(sg.numel, device=device) * 1e-2 for sg in lay.segments]
    ok = all(t.data_ptr() % 16 == 0 for t in grads)
    tab = lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in grads])
    rd, nr = ops.upload_runs(tab, device)
    sc_s = make_scalars(_lib.SGHMC)
    sc_a = ops.make_scalars(_lib.ADAM_CSGHMC, lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                            prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=0.05, t=10, div_mode=_lib.DIV_RECIP)
    s2.fill_(1e-6)
    m.zero_()
    ptr_cases = [("sghmc_per_tensor_gradient_pointers", 24,
                  lambda i: ops.step(_lib.SGHMC, theta, None, theta0, v, None, None, None, rd, nr, sc_s,
                                     ops.make_noise(seed=seed, subseq=2000 + i))),
                 ("adam_csghmc_per_tensor_gradient_pointers", 40,
                  lambda i: ops.step(_lib.ADAM_CSGHMC, theta, None, theta0, v, m, s2, None, rd, nr, sc_a,
                                     ops.make_noise(seed=seed, subseq=8000 + i)))]
    for name, bpp, fn in ptr_cases:
        ms = timed(fn)
        gbs = bpp * n_dense / (ms * 1e-3) / 1e9
        out[name] = {"params_per_s": world * n_dense / (ms * 1e-3), "ms_per_step": ms, "bytes_per_param": bpp,
                     "achieved_gbs_per_gpu": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_nominal_8TBps": gbs / 8000.0,
                     "runs": nr, "aligned": ok, "kernel": "bdl::step_ptable_kernel (run table in the kernel arguments)",
                     "note": "run table with one row per tensor carrying that tensor's p.grad address (the training-loop launch)"}
    return out


def sample_store_extra(lay, theta, step, device, peak, reps=20):
    """SURVEY 8f rows 2-3 at ViT-L/32 size: (i) raw-sample capture into the HBM ring (TMA bulk copy, 8 B/param);
    (ii) a 1.2 GB checkpoint vector written the reference's way (clone + synchronous torch.save: the training stream
    waits for all of it) vs through the asynchronous writer (snapshot launch + submit; D2H and the pickler run on a
    side stream / thread while sampler steps keep launching)."""
    from bayesdll_b200 import ops
    from bayesdll_b200.writer import AsyncWriter
    n, n_dense = lay.n_padded, lay.n_dense
    ring = torch.empty((2, n), dtype=torch.float32, device=device)
    for i in range(3):
        ops.capture_ring(theta, ring, i & 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.capture_ring(theta, ring, i & 1)
    e1.record()
    torch.cuda.synchronize()
    cap_ms = e0.elapsed_time(e1) / reps
    cap_gbs = 8 * n_dense / (cap_ms * 1e-3) / 1e9
    tmp = tempfile.mkdtemp(prefix="bdl_bench_store_")
    try:
        # reference way: the loop blocks until the file is written
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch.save({"last_theta": theta[:n_dense].clone()}, os.path.join(tmp, "sync.pt"))
        sync_s = time.perf_counter() - t0
        # asynchronous writer: stall = snapshot launch + submit; then sampler steps overlap the spill
        w = AsyncWriter(device)
        w.submit(os.path.join(tmp, "warm.pt"), {"x": theta[:1024].clone()})
        w.flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.capture_ring(theta, ring, 0)
        w.submit(os.path.join(tmp, "async.pt"), {"last_theta": ring[0, :n_dense]})
        stall_s = time.perf_counter() - t0
        e0.record()
        k = 0
        while w.is_pending(os.path.join(tmp, "async.pt")) and k < 20000:
            step(10_000 + k)
            k += 1
            if k % 50 == 0:
                torch.cuda.current_stream().synchronize()
        e1.record()
        torch.cuda.synchronize()
        w.flush()
        total_s = time.perf_counter() - t0
        overlapped_ms = e0.elapsed_time(e1) / max(k, 1)
        same = torch.equal(torch.load(os.path.join(tmp, "async.pt"))["last_theta"], ring[0, :n_dense].cpu())
        w.close()
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return {"capture_ms": cap_ms, "capture_gbs": cap_gbs, "capture_frac_of_measured_peak": cap_gbs / peak,
            "capture_kernel": "bdl::ring_copy_kernel (cp.async.bulk global->shared->global), 8 B/param",
            "bytes": 4 * n_dense, "sync_torch_save_stall_s": sync_s, "async_stall_s": stall_s,
            "async_complete_s": total_s, "sampler_steps_overlapped": k, "overlapped_step_ms": overlapped_ms,
            "file_identical": bool(same)}


def train_step_extra(device, rank, world, steps=6, batch=64, backbone="vit_l_32", method="sghmc"):
    """The call a user of the drop-in makes: Model.forward(x, y, net, net0, criterion, lrs, Ninflate, nd[, should_sample]).
    ``method``: "sghmc" (BASELINE.json configs[2], ViT-L/32) or "csghmc" (configs[1], ResNet-101, sampling phase)."""
    import argparse as ap
    import importlib
    import logging
    from bayesdll_b200 import shapes
    mod = importlib.import_module(f"bayesdll_b200.methods.{method}")
    torch.manual_seed(42 + rank)
    with torch.device(device):
        net = shapes.create_backbone(backbone, 37)
        net0 = shapes.create_backbone(backbone, 37)
    a = ap.Namespace(device=device, ND=HP["ND"], lr=HP["lr_body"], lr_head=HP["lr_head"], momentum=0.5, epochs=1,
                     pretrained="synthetic", num_classes=37, ece_num_bins=15, test_eval_freq=1, log_dir=tempfile.gettempdir(),
                     seed=42 + rank,
                     num_cycles=1, proportion_exploration=0.5, full_sample=False,
                     hparams=dict(prior_sig="1.0", Ninflate="1e3", nd="1.0", momentum_decay="0.18", burnin="5", thin="1",
                                  bias="informative", nst="5"))
    lg = logging.getLogger("bench")
    lg.addHandler(logging.NullHandler())
    runner = mod.Runner(net, net0, a, lg)
    extra = dict(should_sample=True) if method == "csghmc" else {}
    runner.net.train()
    x_host = torch.randn(batch, 3, 224, 224).pin_memory()
    y_host = torch.randint(0, 37, (batch,)).pin_memory()
    lrs = [pg["lr"] for pg in runner.optimizer.param_groups]

    def one():
        x, y = x_host.to(device, non_blocking=True), y_host.to(device, non_blocking=True)
        loss, _ = runner.model(x, y, runner.net, runner.net0, runner.criterion, lrs, runner.Ninflate, runner.nd, **extra)
        return loss
    default_mode = runner.model._opts["graph_train"]         # "auto": the Runner's default
    runner.model.configure(graph_train=False)                # first: every launch issued eagerly (hparams graph_train=0)
    for _ in range(3):
        one()
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = one()
    torch.cuda.synchronize()
    dt_eager = allmax(time.perf_counter() - t0, world, device)
    dt = dt_eager
    n = runner.model.chain.layout.n_dense
    # the default (hparams graph_train=auto): forward + loss + backward replayed as one CUDA graph for framework-provided
    # modules, then the fused step (bit-identical to eager, tested)
    graph_ms = None
    try:
        runner.model.configure(graph_train=default_mode)
        for _ in range(4):                                   # two eager warm-ups, the capture, one replay
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize()
        dt_default = allmax(time.perf_counter() - t0, world, device)
        graph_ms = dt_default / steps * 1e3
        if not any(isinstance(v, dict) for v in runner.model._train_graphs.values()):
            graph_ms = f"not captured: {list(runner.model._train_graphs.values())}"
        else:
            dt = dt_default                                  # the headline of this leg is what a user gets by default
    except Exception as e:
        graph_ms = f"failed: {type(e).__name__}: {e}"
    runner.model.configure(graph_train=False)
    runner.model._train_graphs.clear()
    # the same user call the reference's way: fwd + bwd, then the per-tensor eager update loop + SGD step
    # (baseline/eager_port.py restating methods/sghmc.py:482-510, :229) on the same network and GPU
    ref_ms = None
    try:
        if method != "sghmc":
            raise LookupError("the restatement covers SGHMC only; see reference_ms_per_step")
        from baseline import eager_port
        names = [nm for nm, _ in runner.net.named_parameters()]
        params = [p for _, p in runner.net.named_parameters()]
        params0 = [p.data for p in runner.net0.parameters()]
        mom = [torch.zeros_like(p) for p in params]
        hp = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"], prior_sig=HP["prior_sig"],
                  nd=HP["nd"], alpha=HP["alpha"])

        def one_ref():
            x, y = x_host.to(device, non_blocking=True), y_host.to(device, non_blocking=True)
            out = runner.net(x)
            loss_t = runner.criterion(out, y)
            runner.net.zero_grad()
            loss_t.backward()
            eager_port.sghmc_step_eager([p.data for p in params], [p.grad for p in params], params0, mom, names,
                                        runner.net.readout_name, **hp)
            return loss_t.item()
        for _ in range(2):
            one_ref()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_ref()
        torch.cuda.synchronize()
        ref_ms = allmax(time.perf_counter() - t0, world, device) / steps * 1e3
        del mom
    except LookupError as e:
        ref_ms = str(e)
    except Exception as e:                                   # context figure only
        ref_ms = f"failed: {type(e).__name__}: {e}"
    res = {"value": world * n * steps / dt, "unit": "params/s", "ms_per_step": dt / steps * 1e3, "steps": steps,
           "mode": f"hparams graph_train={default_mode} (the Runner default)", "eager_ms_per_step": dt_eager / steps * 1e3,
           "graph_train_ms_per_step": graph_ms, "reference_structure_ms_per_step": ref_ms,
           "images_per_s": world * batch * steps / dt, "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
           "d2h_bytes_per_step": 4, "last_loss": loss,
           "api": f"bayesdll_b200.methods.{method}.Model.forward on torchvision {backbone}, batch {batch}, fp32 fwd/bwd in PyTorch"}
    del runner, net, net0
    torch.cuda.empty_cache()
    # ... and the UNMODIFIED reference itself (baseline/_ref): its own backbone factory, Runner, Model.forward +
    # optimizer.step() on this GPU -- the user-visible step the drop-in replaces; rank 0 only
    if rank == 0:
        try:
            from baseline import reference_arm
            if reference_arm.available():
                hp_ref = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"],
                              prior_sig=HP["prior_sig"], nd=HP["nd"], alpha=HP["alpha"])
                res["reference_ms_per_step"] = reference_arm.train_step_ms(backbone, device, batch=batch, steps=steps, hp=hp_ref, method=method)
                res["reference_kind"] = f"reference (unmodified methods/{method}.py Runner from baseline/_ref on this GPU)"
            else:
                res["reference_ms_per_step"] = None
                res["reference_kind"] = "no reference tree found (baseline/install_ref.py)"
        except Exception as e:                                   # context figure only
            res["reference_ms_per_step"] = f"failed: {type(e).__name__}: {e}"
        torch.cuda.empty_cache()
    return res


def ensemble_extra(device, rank, world, args):
    """BASELINE.json configs[4] through the Runner API: ``csgld.Runner.evaluate()`` with hparams ``eval_shard=1`` on a
    ResNet-101 (K=37), 8 cycles x nst 5 = 40 posterior samples re-drawn for every batch (Appendix B.6), N = 3 669 rows
    (Pets test size) in batches of 64 from pinned host memory, then ECE / MCE / NLL with the rows sharded over the ranks
    and ONE all-reduce of the bin statistics.  preds/s = rows * samples / wall time (max over ranks); the per-phase split
    is device time from CUDA events on rank 0."""
    import argparse as ap
    import logging
    from bayesdll_b200 import calibration, ops, shapes
    from bayesdll_b200 import dist as bdist
    from bayesdll_b200.methods import csgld
    rows, batch, cycles, nst = args.ensemble_rows, 64, 8, 5
    torch.manual_seed(7)                                     # same weights and per-cycle statistics on every rank
    with torch.device(device):
        net = shapes.create_backbone("resnet101", 37)
    x_host = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(5)).pin_memory()
    # one training-mode pass with momentum 1 sets the BatchNorm running statistics to those of a synthetic batch, so the
    # random-init network yields finite O(1) logits in eval mode (running stats are buffers, not sampled)
    for mod in net.modules():
        if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.momentum = 1.0
    net.train()
    with torch.no_grad():
        net(x_host.to(device))
    net.eval()
    a = ap.Namespace(device=device, ND=3680, lr=1e-4, lr_head=1e-2, momentum=0.5, epochs=8, pretrained=None,
                     num_classes=37, ece_num_bins=15, test_eval_freq=1, log_dir=tempfile.gettempdir(), seed=42,
                     num_cycles=cycles, proportion_exploration=0.5, full_sample=False,
                     hparams=dict(prior_sig="1.0", Ninflate="1.0", nd="0.01", thin="10", bias="informative", nst=str(nst),
                                  seed="42", eval_shard="1"))
    lg = logging.getLogger("bench.ensemble")
    lg.addHandler(logging.NullHandler())
    lg.propagate = False
    runner = csgld.Runner(net, None, a, lg)
    n = sum(p.numel() for p in runner.net.parameters())
    gen = torch.Generator(device=device).manual_seed(11)
    with torch.no_grad():
        theta = torch.nn.utils.parameters_to_vector(runner.net.parameters())
    m1, m2 = {}, {}
    for c in range(1, cycles + 1):                           # synthetic per-cycle posterior statistics (dense, checkpoint layout)
        mean = theta + 1e-3 * torch.randn(n, device=device, generator=gen)
        m1[c] = mean
        m2[c] = mean * mean + 1e-6 * torch.rand(n, device=device, generator=gen)
    runner.cycle_theta_mom1, runner.cycle_theta_mom2 = m1, m2
    del m1, m2, theta
    runner.samples_per_cycle = {c: nst for c in range(1, cycles + 1)}
    runner.cycle_likelihoods = {c: [0.5] * nst for c in range(1, cycles + 1)}
    runner.current_cycle = cycles
    y_all = torch.randint(0, 37, (rows,), generator=torch.Generator().manual_seed(3))
    loader = [(x_host[:len(y_all[i:i + batch])], y_all[i:i + batch].pin_memory()) for i in range(0, rows, batch)]
    runner.evaluate(loader[:2] + loader[-1:])                # warm-up (cuDNN autotune, allocator, both batch shapes)
    barrier(world)
    torch.cuda.synchronize()
    runner.profile_eval = True
    t0 = time.perf_counter()
    loss, err, targets, logits, logits_all = runner.evaluate(loader)
    t_eval = time.perf_counter() - t0
    ece, mce, nll = bdist.calibrate_sharded(targets, logits, a.ece_num_bins, device, rank, world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt_max = allmax(dt, world, device)
    S = cycles * nst
    phases = dict(runner.eval_phases)
    phases["calibrate_ms"] = (dt - t_eval) * 1e3
    phases["wall_ms"] = dt * 1e3
    # the same numbers the unsharded host API gives (every rank holds the full logits): bin counts must be identical
    e1, m1_, n1 = calibration.analyze(targets, logits, a.ece_num_bins, None)
    _, _, _, _, sizes = calibration.calc_bins(targets, logits, a.ece_num_bins)
    # context: cost of materialising ONE posterior sample, ours (bdl_draw) vs the reference's structure on this GPU
    # (deepcopy(net) + randn_like / sqrt / mul / add / copy_ per tensor, methods/csgld.py:404-413)
    lay = runner.model.chain.layout
    c1, c2 = runner._cyc1[1], runner._cyc2[1]
    out = torch.empty_like(c1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        ops.draw(c1, c2, out, ops.VAR_FROM_MOMENTS, 1.25, ops.make_noise(seed=42, subseq=i, stream_id=2))
    torch.cuda.synchronize()
    ours_draw_ms = (time.perf_counter() - t0) / 10 * 1e3
    import copy
    mean_views, var_views = lay.views(c1), lay.views(torch.clamp(c2 - c1 ** 2, min=1e-12))
    t0 = time.perf_counter()
    for i in range(3):
        with torch.no_grad():
            net_sample = copy.deepcopy(runner.net)
            for p_, p_mean, p_var in zip(net_sample.parameters(), mean_views, var_views):
                p_.copy_(p_mean + p_var.sqrt() * torch.randn_like(p_))
    torch.cuda.synchronize()
    ref_draw_ms = (time.perf_counter() - t0) / 3 * 1e3
    del net_sample
    runner.flush_io()
    return {"metric": "ensemble preds/s (ResNet-101 cSGLD 40-sample posterior-predictive ensemble + ECE/MCE/NLL)",
            "value": rows * S / dt_max, "unit": "preds/s", "rows": rows, "samples": S, "batch": batch, "seconds": dt_max, "n_gpus": world,
            "api": "bayesdll_b200.methods.csgld.Runner.evaluate(loader) with hparams eval_shard=1 (5-tuple incl. logits_all "
                   f"{list(logits_all.shape)}), then row-sharded ECE/MCE/NLL",
            "sharding": "samples dealt round-robin (cycle*nst + s); one all-gather of the local sample logits, one "
                        "all-reduce of the 3*M+2 bin statistics",
            "phases_rank0_ms": phases, "h2d_bytes": sum(x.numel() * 4 + y.numel() * 8 for x, y in loader),
            "d2h_bytes": int(targets.nbytes + logits.nbytes + logits_all.nbytes),
            "ece": float(ece), "mce": float(mce), "nll": float(nll), "loss": float(loss), "err": float(err),
            "ece_unsharded": float(e1), "bin_sizes": [int(v) for v in sizes],
            "draw_ms": ours_draw_ms, "reference_structure_draw_ms": ref_draw_ms,
            "note": "per-sample cost = 1 draw kernel (12 B/param) + PyTorch fp32 forward of 64 images (CUDA-graph replay); "
                    "samples re-drawn for every batch as the reference does (Appendix B.6)"}


# ------------------------------------------------------------------------------------------------------------
def cpu_port_rate(lay, seconds=15.0, with_eager=False, n_limit=None, steps=None, warmup=1):
    """Time the fused C/OpenMP port of the reference path on the host cores (bounded sample)."""
    from bayesdll_b200 import _lib
    from oracle import c_oracle
    c_oracle.build()
    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers; report what is really used
    cores = c_oracle.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    n_full = lay.n_padded
    rng = np.random.default_rng(0)

    def alloc(n):
        th = torch.randn(n).mul_(0.02).numpy()
        g = torch.randn(n).mul_(0.01).numpy()
        th0 = torch.randn(n).mul_(0.02).numpy()
        v = np.zeros(n, np.float32)
        return th, g, th0, v

    def table(n):
        tab = (_lib.Run * 1)()
        tab[0].begin, tab[0].end, tab[0].valid_end, tab[0].cls = 0, n, n, _lib.CLS_PRIOR
        return tab
    sc = make_scalars(_lib.SGHMC)
    nz = _lib.Noise()
    nz.seed, nz.stream_id = 42, _lib.STREAM_STEP
    # probe on 16 Mi params to size the sample
    n_probe = min(n_full, 16 << 20)
    st = alloc(n_probe)
    tab = table(n_probe)
    c_oracle.step(_lib.SGHMC, st[0], st[1], st[2], st[3], None, None, None, tab, sc, nz)
    t0 = time.perf_counter()
    c_oracle.step(_lib.SGHMC, st[0], st[1], st[2], st[3], None, None, None, tab, sc, nz)
    rate = n_probe / (time.perf_counter() - t0)
    if steps is None:
        steps = 3
    n = n_limit or n_full
    n = int(min(n, max(8 << 20, rate * seconds / (steps + warmup)))) // 4 * 4
    if n != n_probe:
        st = alloc(n)
        tab = table(n)
    for i in range(warmup):
        nz.subseq = i
        c_oracle.step(_lib.SGHMC, st[0], st[1], st[2], st[3], None, None, None, tab, sc, nz)
    t0 = time.perf_counter()
    for i in range(steps):
        nz.subseq = warmup + i
        c_oracle.step(_lib.SGHMC, st[0], st[1], st[2], st[3], None, None, None, tab, sc, nz)
    dt = time.perf_counter() - t0
    out = {"value": n * steps / dt, "unit": "params/s", "cores": cores, "kind": "port",
           "sample": f"{steps} fused SGHMC steps over {n} of {lay.n_dense} ViT-L/32 parameters (oracle/bdl_oracle.c, "
                     f"OpenMP, in-loop Philox+Box-Muller)", "ms_per_step": dt / steps * 1e3, "steps": steps, "n": n}
    if with_eager:
        out["reference_structure_eager"] = eager_rate(lay, cores)
    return out


def eager_rate(lay, cores, max_params=64 << 20):
    """What the reference's own structure (per-tensor torch eager loop + SGD.step) achieves on these cores."""
    from baseline import eager_port
    torch.set_num_threads(cores)
    segs, tot = [], 0
    for s in lay.segments:
        segs.append(s)
        tot += s.numel
        if tot >= max_params:
            break
    names = [s.name for s in segs]
    params = [torch.randn(s.shape) * 0.02 for s in segs]
    params0 = [torch.randn(s.shape) * 0.02 for s in segs]
    grads = [torch.randn(s.shape) * 0.01 for s in segs]
    mom = [torch.zeros(s.shape) for s in segs]
    kw = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"], prior_sig=HP["prior_sig"],
              nd=HP["nd"], alpha=HP["alpha"])
    eager_port.sghmc_step_eager(params, grads, params0, mom, names, lay.readout_name, **kw)
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        eager_port.sghmc_step_eager(params, grads, params0, mom, names, lay.readout_name, **kw)
    dt = time.perf_counter() - t0
    return {"value": tot * reps / dt, "unit": "params/s", "cores": cores, "kind": "port",
            "sample": f"{reps} per-tensor torch-eager SGHMC steps over the first {len(segs)} ViT-L/32 tensors ({tot} params), "
                      f"baseline/eager_port.py restating methods/sghmc.py:482-510 + SGD.step"}


def eager_gpu_rate(lay, device, reps=3):
    """The reference's own structure on the SAME B200 ('reference on the same B200'): the unmodified
    methods/sghmc.py Model.forward + optimizer.step() over all 296 ViT-L/32 tensors when a reference tree is present,
    else its restatement baseline/eager_port.py.  This is the number the fused kernel replaces."""
    from baseline import reference_arm
    if reference_arm.available() is not None:
        named = [(sg.name, sg.shape) for sg in lay.segments]
        r = reference_arm.sghmc_update_rate(named, lay.readout_name, device, steps=reps, warmup=2, hp=HP)
        r["kernel_launches_per_step"] = "~12 eager kernels x 296 tensors"
        return r
    from baseline import eager_port
    names = [s.name for s in lay.segments]
    gen = torch.Generator(device=device).manual_seed(1)
    mk = lambda sc: [torch.randn(s.shape, device=device, generator=gen) * sc for s in lay.segments]
    params, params0, grads = mk(0.02), mk(0.02), mk(0.01)
    mom = [torch.zeros(s.shape, device=device) for s in lay.segments]
    kw = dict(lr_body=HP["lr_body"], lr_head=HP["lr_head"], ND=HP["ND"], Ninflate=HP["Ninflate"], prior_sig=HP["prior_sig"],
              nd=HP["nd"], alpha=HP["alpha"])
    for _ in range(2):
        eager_port.sghmc_step_eager(params, grads, params0, mom, names, lay.readout_name, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eager_port.sghmc_step_eager(params, grads, params0, mom, names, lay.readout_name, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": lay.n_dense / (ms * 1e-3), "unit": "params/s", "ms_per_step": ms, "steps": reps, "kind": "port",
            "kernel_launches_per_step": "~12 eager kernels x 296 tensors",
            "sample": "per-tensor torch-eager SGHMC loop + SGD step on cuda:0, all 296 ViT-L/32 tensors (baseline/eager_port.py)"}


def reference_cpu_rate(lay, seconds, steps, warmup):
    """The reference's OWN code on the host cores (kind "reference"), or None when no reference tree exists
    (neither /root/reference nor baseline/_ref)."""
    from baseline import reference_arm
    if reference_arm.available() is None:
        return None
    named = [(sg.name, sg.shape) for sg in lay.segments]
    return reference_arm.sghmc_update_rate(named, lay.readout_name, "cpu", steps=steps, warmup=warmup, hp=HP, seconds=seconds)


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on the host cores -- its unmodified
    methods/sghmc.py Model.forward + optimizer.step() (kind "reference") when a reference tree is present, the fused C port
    (kind "port") otherwise.  Rank 0 only; every step is a bounded sample of the workload."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    lay = build_layout()
    K, W = args.steps, max(args.warmup, 1)
    res = guarded("reference arm", reference_cpu_rate, lay, args.ref_seconds, K, W)
    if not res or "error" in res:
        res = cpu_port_rate(lay, seconds=args.ref_seconds, steps=K, warmup=W)
    world = env_int("WORLD_SIZE", 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "params/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights of the named shapes, g ~ N(0,1e-2^2))",
        "config": bench_config(world, lay.n_dense),
        "cpu_baseline": {"value": res["value"], "unit": "params/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "params/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_params_per_step": res["n"],
    }
    if not args.no_cfg1:
        line["cfg1_mlp_mnist_sgld"] = {"reference_cpu": guarded("cfg1", cfg1_line, True, "cpu")}
    print(json.dumps(line), flush=True)


def cfg1_line(reference, device, graph_train=False):
    from baseline import reference_arm
    if reference and reference_arm.available() is None:
        return {"unavailable": "no reference tree (/root/reference or baseline/_ref)"}
    return reference_arm.cfg1_mlp_mnist(device=device, reference=reference, graph_train=graph_train)


def main():
    # NCCL prints its version banner / debug lines to stdout by default; rank 0's stdout must carry the JSON line only
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=60.0)
    ap.add_argument("--ensemble-rows", type=int, default=3669, help="Pets test-set size (BASELINE.json configs[4])")
    ap.add_argument("--no-cfg1", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-gpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-sample-store", action="store_true")
    args = ap.parse_args()
    world = env_int("WORLD_SIZE", 1)
    if args.gpus != world and args.impl == "ours" and world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
