"""Drop-in for the reference's ``calibration`` module (calibration.py:24-259): same function names, arguments
and return values; the binning / ECE / MCE / NLL reductions run as one CUDA kernel (bdl_calibrate).

    calc_bins(labels, logits, num_bins, temperature=1)           calibration.py:24
    analyze(labels, logits, num_bins, plot_save_path, temperature=1) -> (ece, mce, nll)      :215
    find_optimal_temperature(labels, logits, plot_save_path, max_iter=10000) -> (Topt, ok)   :123
    draw_reliability_plot(...)                                                               :70

Inputs are the numpy arrays the Runners produce (``targets [N] int64``, ``logits [N,K] f32``); CUDA tensors are
accepted too and avoid the upload.  The temperature optimiser keeps scipy's scalar BFGS driver on the host and
evaluates its objective on the device (bdl_nll_temperature).  The plots are presentation (SURVEY.md section 2.1
row 8) and are drawn only when matplotlib is importable, exactly the reference's dependency.
"""
import numpy as np
import torch

from . import ops

_DEVICE = None


def _device():
    global _DEVICE
    if _DEVICE is None:
        if not torch.cuda.is_available():
            raise ops.BdlError("bayesdll_b200.calibration needs a CUDA device (no CPU fallback)")
        _DEVICE = torch.device("cuda", torch.cuda.current_device())
    return _DEVICE


def _to_dev(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(device=_device() if not a.is_cuda else a.device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(_device())


def bin_edges(num_bins):
    """Right bin boundaries, host fp64, exactly ``np.linspace(0, 1+1e-8, num_bins+1)[1:]`` (calibration.py:54)."""
    return np.linspace(0, 1 + 1e-8, num_bins + 1)[1:]


def _temperature(temperature):
    """(T as float, use_f64).  The reference's dtype behaviour: an int/float scalar
    keeps ``logits/T`` and the softmax in fp32; the fp64 ndarray returned by find_optimal_temperature promotes
    everything to fp64."""
    if isinstance(temperature, np.ndarray) or isinstance(temperature, np.floating):
        return float(np.asarray(temperature, dtype=np.float64).reshape(-1)[0]), True
    if isinstance(temperature, torch.Tensor):
        return float(temperature.reshape(-1)[0]), temperature.dtype == torch.float64
    return float(temperature), False


def _bin_stats(labels, logits, num_bins, temperature, want_binned):
    T, f64 = _temperature(temperature)
    lg = _to_dev(logits, torch.float32)
    lb = _to_dev(labels, torch.int64)
    bins = bin_edges(num_bins)
    edges = torch.from_numpy(bins).to(lg.device)
    size, acc, conf, nll, near, binned = ops.calibrate(lg, lb, edges, T, f64, want_binned)
    host = torch.cat([size, acc, conf, nll, near.double()]).cpu().numpy()     # one D2H read
    M = num_bins
    return bins, host[:M], host[M:2 * M], host[2 * M:3 * M], host[3 * M], int(host[3 * M + 1]), binned, lg.shape[0]


def calc_bins(labels, logits, num_bins, temperature=1):
    """-> (bins, binned, bin_accs, bin_confs, bin_sizes) as the reference (calibration.py:24-67)."""
    bins, sizes, acc_sum, conf_sum, _, _, binned, _ = _bin_stats(labels, logits, num_bins, temperature, True)
    nz = sizes > 0
    bin_accs = np.zeros(num_bins)
    bin_confs = np.zeros(num_bins)
    bin_accs[nz] = acc_sum[nz] / sizes[nz]
    bin_confs[nz] = conf_sum[nz] / sizes[nz]
    return bins, binned.cpu().numpy().astype(np.int64), bin_accs, bin_confs, sizes


last_near_edge = 0   # certification counter of the most recent analyze()/calc_bins() call (see bdl_calibrate)


def analyze(labels, logits, num_bins, plot_save_path, temperature=1):
    """-> (ece, mce, nll); draws the reliability plot when matplotlib is available (calibration.py:215-259)."""
    global last_near_edge
    bins, sizes, acc_sum, conf_sum, nll_sum, near, _, N = _bin_stats(labels, logits, num_bins, temperature, False)
    last_near_edge = near
    nz = sizes > 0
    bin_accs = np.zeros(num_bins)
    bin_confs = np.zeros(num_bins)
    bin_accs[nz] = acc_sum[nz] / sizes[nz]
    bin_confs[nz] = conf_sum[nz] / sizes[nz]
    ece = (np.abs(bin_accs - bin_confs) * (sizes / sizes.sum())).sum()
    mce = np.abs(bin_accs - bin_confs).max()
    nll = nll_sum / N
    if plot_save_path is not None:
        draw_reliability_plot(bins, bin_accs, plot_save_path, title=f"Temperature = {temperature}", ece=ece, mce=mce,
                              nll=nll)
    return ece, mce, nll


def draw_reliability_plot(bins, bin_accs, fig_name, title=None, ece=None, mce=None, nll=None):
    """Reliability diagram (calibration.py:70-120).  Presentation only; skipped when matplotlib is absent."""
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.patches as mpatches
        import matplotlib.pyplot as plt
    except Exception:
        return False
    centers = (np.insert(bins, 0, 0)[:-1] + bins) / 2
    width = centers[1] - centers[0] if len(centers) > 1 else 1.0
    fig = plt.figure(figsize=(8, 8))
    ax = fig.gca()
    ax.set_xlim(0, 1 + 1e-8)
    ax.set_ylim(0, 1)
    ax.set_xlabel("Confidence")
    ax.set_ylabel("Accuracy")
    ax.set_axisbelow(True)
    ax.grid(color="gray", linestyle="dashed")
    ideal = ax.bar(centers, centers, width=width, alpha=0.3, edgecolor="black", color="r", hatch="\\")
    model = ax.bar(centers, bin_accs, width=width, alpha=0.3, edgecolor="black", color="b")
    diag, = ax.plot([0, 1], [0, 1], "--", color="gray", linewidth=2)
    ax.set_aspect("equal", adjustable="box")
    first = ax.legend([diag, ideal, model], ["Y=X", "Ideal", "Model"], loc="upper left")
    if ece is not None and mce is not None and nll is not None:
        ax.legend(handles=[mpatches.Patch(color="green", label="ECE = {:.2f}%".format(ece * 100)),
                           mpatches.Patch(color="red", label="MCE = {:.2f}%".format(mce * 100)),
                           mpatches.Patch(color="blue", label="NLL = {:.4f}".format(nll))], loc="lower right")
        ax.add_artist(first)
    if title is not None:
        ax.set_title(title)
    fig.savefig(fig_name, bbox_inches="tight")
    plt.close(fig)
    return True


def find_optimal_temperature(labels, logits, plot_save_path, max_iter=10000):
    """Temperature scaling on the validation set (calibration.py:123-212): scipy's BFGS over the scalar T, exactly the
    reference's driver (``minimize(fun, np.ones(1), options={'maxiter': max_iter}, callback=...)``, numerical gradient),
    but every objective evaluation ``mean(logsumexp(logits/T) - (logits/T)[y])`` is ONE fp64 device reduction
    (bdl_nll_temperature) over logits uploaded once, instead of three host passes over [N,K] per evaluation.
    Returns ``(result.x, result.success)``: ``result.x`` is the fp64 array the reference returns."""
    import scipy.optimize
    lg = _to_dev(logits, torch.float32)
    lb = _to_dev(labels, torch.int64)
    row = torch.empty(lg.shape[0], dtype=torch.float64, device=lg.device)
    out = torch.empty(1, dtype=torch.float64, device=lg.device)

    def fun(T):
        ops.nll_temperature(lg, lb, float(np.asarray(T, dtype=np.float64).reshape(-1)[0]), row, out)
        return out.item()

    temps, losses = [], []

    def callback(x):
        temps.append(x)
        losses.append(fun(x))

    result = scipy.optimize.minimize(fun, np.ones(1), options={"maxiter": max_iter}, callback=callback)
    success = result.success
    try:
        Topt = result.x
        _draw_temperature_curve(temps, losses, plot_save_path)
    except Exception:
        Topt = 1
    return Topt, success


def _draw_temperature_curve(temps, losses, plot_save_path):
    """Optimisation curve (calibration.py:196-208).  Presentation only; skipped when matplotlib is absent."""
    if plot_save_path is None:
        return False
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.pyplot as plt
    except Exception:
        return False
    fig = plt.figure()
    plt.subplot(121)
    plt.plot(list(range(len(temps))), [float(np.asarray(t).reshape(-1)[0]) for t in temps])
    plt.gca().set_title("Temperature T")
    plt.gca().set_xlabel("Iterations")
    plt.subplot(122)
    plt.plot(list(range(len(losses))), losses)
    plt.gca().set_title("NLL on validation set")
    plt.gca().set_xlabel("Iterations")
    fig.savefig(plot_save_path)
    plt.close(fig)
    return True
