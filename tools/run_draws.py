"""Launch the streaming draws and the training-loop (gradient-pointer table) step a few times at ViT-L/32 size -- the
command the round-1 ncu captures of these kernels profile:

    ncu --set full --clock-control none --import-source on -k regex:'dropout_mix|draw_kernel|step_kernel' -s 8 -c 8 \
        python tools/run_draws.py
"""
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
    theta, theta0, v = (torch.randn(n, device=dev, generator=gen) * s for s in (0.02, 0.02, 0.001))
    mom2 = theta * theta + 1e-6
    out = torch.empty(n, device=dev)
    dr, dn = ops.upload_runs(lay.dropout_run_table("gaussian"), dev)
    grads = [torch.randn(sg.numel, device=dev, generator=gen) * 1e-2 for sg in lay.segments]
    rd, nr = ops.upload_runs(lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in grads]), dev)
    sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, prior_sig=1.0, nd=1.0, alpha=0.18)
    for i in range(4):          # 4 kernels per round -> with "-s 8 -c 8" ncu profiles rounds 3 and 4
        ops.draw(theta, mom2, out, ops.VAR_FROM_MOMENTS, 1.25, ops.make_noise(seed=1, subseq=i, stream_id=_lib.STREAM_DRAW))
        ops.draw(theta, mom2, out, ops.STD_GIVEN, 1.0, ops.make_noise(seed=1, subseq=10 + i, stream_id=_lib.STREAM_DRAW))
        ops.dropout_mix(theta, theta0, out, 0.1, ops.make_noise(seed=1, subseq=20 + i, stream_id=_lib.STREAM_DRAW), dr, dn)
        ops.step(_lib.SGHMC, theta, None, theta0, v, None, None, None, rd, nr, sc, ops.make_noise(seed=1, subseq=30 + i))
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
