"""Chunk-size sweep of the host-buffer chain step (bdl_chain_step_host) at ViT-L/32 size (run under gpurun)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import _lib, ops, shapes  # noqa: E402
from bayesdll_b200.flat import FlatLayout  # noqa: E402

named, readout = shapes.named_shapes("vit_l_32", 37)
lay = FlatLayout(named, readout)
n = lay.n_padded
tab = lay.run_table("informative")
sc = ops.make_scalars(_lib.SGHMC, lr_body=1e-4, lr_head=1e-2, ND=1840, Ninflate=1e3, alpha=0.18)
g = (torch.randn(n) * 0.01).pin_memory()
out = torch.empty(n).pin_memory()
for chunk in (1 << 20, 4 << 20, 8 << 20, 16 << 20, 32 << 20, 64 << 20, n):
    ch = ops.HostChain(n, _lib.SGHMC, chunk_elems=chunk)
    ch.upload(_lib.BUF_THETA, g)
    ch.upload(_lib.BUF_THETA0, g)
    for i in range(2):
        ch.step_host(g, out, tab, sc, ops.make_noise(seed=1, subseq=i))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 8
    for i in range(K):
        ch.step_host(g, out, tab, sc, ops.make_noise(seed=1, subseq=2 + i))
    dt = (time.perf_counter() - t0) / K
    print(f"chunk {chunk >> 20:5d} Mi elems: {dt * 1e3:7.2f} ms/step  {lay.n_dense / dt / 1e9:6.2f} G params/s  "
          f"{4 * n / dt / 1e9:5.1f} GB/s each way", flush=True)
    ch.close()
# reference points: plain pinned copies
d = torch.empty(n, device="cuda")
for name, fn in (("H2D only", lambda: d.copy_(g, non_blocking=True)), ("D2H only", lambda: out.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt * 1e3:.2f} ms  {4 * n / dt / 1e9:.1f} GB/s")
