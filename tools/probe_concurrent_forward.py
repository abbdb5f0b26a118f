"""Would running the posterior samples of one batch CONCURRENTLY help the ensemble leg?  ResNet-101 eval forward, batch 64
(the 99 % of `Runner.evaluate`'s time): S graph-captured copies of the network, each on its own stream, replayed together
vs one after the other.  Same kernels either way (outputs compared bit for bit).  Run under gpurun."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesdll_b200 import shapes  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with torch.device(dev):
        net = shapes.create_backbone("resnet101", 37).eval()
    x = torch.randn(64, 3, 224, 224, device=dev)
    lanes = 4
    nets = [copy.deepcopy(net) for _ in range(lanes)]
    streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
    graphs, outs, xs = [], [], []
    pool = None
    for n_, s in zip(nets, streams):
        sx = x.clone()
        with torch.no_grad(), torch.cuda.stream(s):
            for _ in range(2):
                n_(sx)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g, stream=s):
            o = n_(sx)
        graphs.append(g)
        outs.append(o)
        xs.append(sx)
    torch.cuda.synchronize()

    def run(k, reps=40):
        """reps rounds; in each round k lanes replay at once (each on its own stream), then the main stream joins them."""
        main_s = torch.cuda.current_stream()
        for warm in (True, False):
            if not warm:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for _ in range(3 if warm else reps):
                for i in range(k):
                    streams[i].wait_stream(main_s)
                    with torch.cuda.stream(streams[i]):
                        graphs[i].replay()
                for i in range(k):
                    main_s.wait_stream(streams[i])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * k)

    graphs[0].replay()                                   # a capture records, it does not run
    torch.cuda.synchronize()
    ref = outs[0].clone()
    for k in (1, 2, 3, 4):
        ms = run(k)
        same = all(torch.equal(outs[i], ref) for i in range(k))
        print(f"{k} forward(s) in flight: {ms:.3f} ms per forward of 64 images  ({64 / ms * 1e3:.0f} images/s)  outputs identical: {same}",
              flush=True)


if __name__ == "__main__":
    main()
