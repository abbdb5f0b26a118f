"""Eval-mode forward of the BASELINE backbones: PyTorch eager vs one CUDA-graph replay (same kernels, no host launch
cost).  python tools/probe_graph_forward.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import shapes  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for backbone, batch in (("resnet101", 64), ("resnet101", 16), ("vit_l_32", 64), ("vit_l_32", 16)):
        with torch.device(dev):
            net = shapes.create_backbone(backbone, 37).eval()
        x = torch.randn(batch, 3, 224, 224, device=dev)
        with torch.no_grad():
            for _ in range(3):
                ref = net(x)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                net(x)
            torch.cuda.synchronize()
            eager = (time.perf_counter() - t0) / 20 * 1e3
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    net(x)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = net(x)
            g.replay()
            torch.cuda.synchronize()
            same = torch.equal(out, ref)
            t0 = time.perf_counter()
            for _ in range(20):
                g.replay()
            torch.cuda.synchronize()
            graphed = (time.perf_counter() - t0) / 20 * 1e3
        print(f"{backbone:10s} batch {batch:3d}: eager {eager:7.3f} ms  graph {graphed:7.3f} ms  ({eager / graphed:.2f}x)  "
              f"bit-identical {same}", flush=True)
        del net, g, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
