"""Launch the step builds that matter at ViT-L/32 size a few times -- the command the round-2 ncu captures profile:

    ncu --set full --clock-control none -k regex:'step_' -s 12 -c 16 -o /tmp/r02_steps_full python tools/run_steps.py --generic

Per round, for SGHMC and for Adam-cSGHMC: flat gradient (step_kernel<kFast>, the headline launch), per-tensor gradient
pointers with the run table in the kernel arguments (step_ptable_kernel: the launch Runner.train() makes), the same table
searched in device memory (step_table_kernel: the fallback for > 512 rows).  `--generic` additionally launches the two
pointer cases through the generic build (explicit 64-thread shape), i.e. the round-1 training-loop launch, for an
instruction-count comparison.
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
    theta, theta0, g, v, m = (torch.randn(n, device=dev, generator=gen) * s for s in (0.02, 0.02, 0.01, 0.001, 0.001))
    s2 = torch.rand(n, device=dev, generator=gen) * 1e-4 + 1e-8
    flat_tab, flat_n = ops.upload_runs(lay.run_table("informative"), dev)
    grads = [torch.randn(sg.numel, device=dev, generator=gen) * 1e-2 for sg in lay.segments]
    rd, nr = ops.upload_runs(lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in grads]), dev)
    rd_dev_only, _ = ops.upload_runs(lay.run_table("informative", grad_ptrs=[t.data_ptr() for t in grads]), dev)
    del rd_dev_only._bdl_host          # no host copy -> the table is searched in device memory (step_table_kernel)
    kw = dict(lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, prior_sig=1.0, nd=1.0)
    sc_s = ops.make_scalars(_lib.SGHMC, alpha=0.18, **kw)
    sc_a = ops.make_scalars(_lib.ADAM_CSGHMC, alpha=0.05, t=10, **kw)
    rounds = 4                   # 6 kernels per round -> with "-s 12 -c 12" ncu profiles rounds 3 and 4 (+4 generic launches)
    for i in range(rounds):
        ops.step(_lib.SGHMC, theta, g, theta0, v, None, None, None, flat_tab, flat_n, sc_s, ops.make_noise(seed=1, subseq=i))
        ops.step(_lib.SGHMC, theta, None, theta0, v, None, None, None, rd, nr, sc_s, ops.make_noise(seed=1, subseq=10 + i))
        ops.step(_lib.SGHMC, theta, None, theta0, v, None, None, None, rd_dev_only, nr, sc_s, ops.make_noise(seed=1, subseq=60 + i))
        ops.step(_lib.ADAM_CSGHMC, theta, g, theta0, v, m, s2, None, flat_tab, flat_n, sc_a, ops.make_noise(seed=1, subseq=20 + i))
        ops.step(_lib.ADAM_CSGHMC, theta, None, theta0, v, m, s2, None, rd, nr, sc_a, ops.make_noise(seed=1, subseq=30 + i))
        ops.step(_lib.ADAM_CSGHMC, theta, None, theta0, v, m, s2, None, rd_dev_only, nr, sc_a, ops.make_noise(seed=1, subseq=70 + i))
    if "--generic" in sys.argv:
        ops.set_launch_config(0, 1, 64)
        for i in range(2):
            ops.step(_lib.SGHMC, theta, None, theta0, v, None, None, None, rd, nr, sc_s, ops.make_noise(seed=1, subseq=40 + i))
            ops.step(_lib.ADAM_CSGHMC, theta, None, theta0, v, m, s2, None, rd, nr, sc_a, ops.make_noise(seed=1, subseq=50 + i))
        ops.set_launch_config(0, 0, 0)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
