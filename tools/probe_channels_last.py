"""What bounds the ensemble leg: the PyTorch ResNet-101 eval forward (99 % of its wall time).  Probe, on one B200, of the
layouts cuDNN can be handed WITHOUT touching the flat parameter buffer (weights stay contiguous NCHW views of it):
plain NCHW input (what the reference does) vs a channels_last input, eager vs CUDA-graph replay, TF32 convolutions on
(torch's default, `torch.backends.cudnn.allow_tf32`) vs off.  Prints ms per 64-image forward and the largest logit
deviation from the NCHW result (run under gpurun)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import shapes  # noqa: E402


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def graphed(net, x):
    sx = x.clone()
    with torch.no_grad():
        for _ in range(2):
            net(sx)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = net(sx)
    return g, out


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with torch.device(dev):
        net = shapes.create_backbone("resnet101", 37)
    for mod in net.modules():
        if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.momentum = 1.0
    x = torch.randn(64, 3, 224, 224, device=dev)
    net.train()
    with torch.no_grad():
        net(x)
    net.eval()
    print(f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32} matmul.allow_tf32={torch.backends.cuda.matmul.allow_tf32} "
          f"cudnn.benchmark={torch.backends.cudnn.benchmark}", flush=True)
    ref = None
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        for bench in (False, True):
            torch.backends.cudnn.benchmark = bench
            for name, xin in (("nchw", x), ("channels_last_input", x.contiguous(memory_format=torch.channels_last))):
                with torch.no_grad():
                    eager = timed(lambda: net(xin))
                    g, out = graphed(net, xin)
                    rep = timed(g.replay)
                    g.replay()
                    o = out.float().clone()
                if ref is None:
                    ref = o
                print(f"tf32={int(tf32)} cudnn.benchmark={int(bench)} {name:22s} eager {eager:7.3f} ms  graph {rep:7.3f} ms  "
                      f"max|dlogit| vs first {float((o - ref).abs().max()):.3e}  (logit scale {float(ref.abs().max()):.2f})", flush=True)
                del g, out


if __name__ == "__main__":
    main()
