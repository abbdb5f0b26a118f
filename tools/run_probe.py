"""A few launches of the bare-traffic yardstick (bdl_probe_stream) at ViT-L/32 size, for an ncu capture next to the step
kernels':  ncu --set full --clock-control none -k regex:probe_stream -s 3 -c 3 -o /tmp/probe python tools/run_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import ops, shapes  # noqa: E402
from bayesdll_b200.flat import FlatLayout  # noqa: E402

named, readout = shapes.named_shapes("vit_l_32", 37)
n = FlatLayout(named, readout).n_padded
dev = torch.device("cuda:0")
a, b, c, d = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
for _ in range(3):
    ops.probe_stream(a, b, c, d, 4, 2, threads=64)
for _ in range(3):
    ops.probe_stream(a, None, c, d, 2, 1, threads=128)
for _ in range(3):
    ops.probe_stream(a, None, c, None, 1, 1, threads=128)
torch.cuda.synchronize()
print("ok")
