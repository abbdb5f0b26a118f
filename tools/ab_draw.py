"""Back-to-back timing of the streaming draws at ViT-L/32 size (for A/B builds: BDL_NVCC_EXTRA=... python -m
bayesdll_b200.build --force; python tools/ab_draw.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import _lib, ops, shapes  # noqa: E402
from bayesdll_b200.flat import FlatLayout  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    named, readout = shapes.named_shapes("vit_l_32")
    lay = FlatLayout(named, readout)
    n = lay.n_padded
    gen = torch.Generator(device=dev).manual_seed(0)
    theta = torch.randn(n, device=dev, generator=gen) * 0.02
    mom2 = theta * theta + 1e-6
    out = torch.empty(n, device=dev)
    cases = [("posterior_draw", lambda i: ops.draw(theta, mom2, out, ops.VAR_FROM_MOMENTS, 1.25, ops.make_noise(seed=1, subseq=i, stream_id=_lib.STREAM_DRAW))),
             ("welford_draw", lambda i: ops.draw(theta, mom2, out, ops.VAR_FROM_WELFORD, 6.0, ops.make_noise(seed=1, subseq=i, stream_id=_lib.STREAM_DRAW))),
             ("vi_draw", lambda i: ops.draw(theta, mom2, out, ops.STD_GIVEN, 1.0, ops.make_noise(seed=1, subseq=i, stream_id=_lib.STREAM_DRAW)))]
    if "--mix" in sys.argv:                          # the MC-Dropout mix, with and without the 294-run bias table
        theta0 = torch.randn(n, device=dev, generator=gen) * 0.02
        dr, dn = ops.upload_runs(lay.dropout_run_table("gaussian"), dev)
        cases = [("mix_bias_table", lambda i: ops.dropout_mix(theta, theta0, out, 0.1, ops.make_noise(seed=1, subseq=i, stream_id=_lib.STREAM_DRAW), dr, dn)),
                 ("mix_no_table", lambda i: ops.dropout_mix(theta, theta0, out, 0.1, ops.make_noise(seed=1, subseq=i, stream_id=_lib.STREAM_DRAW)))]
    res = {k: [] for k, _ in cases}
    for rnd in range(3):
        for name, fn in cases:
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 200)
    sums = []
    for name, fn in cases:                               # same bits from every build: checksum of one fixed draw per case
        fn(12345)
        sums.append(f"{int(out.view(torch.int32).sum(dtype=torch.int64)):x}")
    print("checksums " + " ".join(sums), flush=True)
    print("  ".join(f"{k}: " + "/".join(f"{x:.4f}" for x in v) + f" ms ({12 * lay.n_dense / min(v) / 1e6:.0f} GB/s)" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
