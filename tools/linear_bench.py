"""Timing of the fused small-M linear kernels (csrc/linear_ops.cu) at the default generator's sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import multi_stylegan_b200.multi_stylegan_generator as G_mod


def timeit(fn, n=20):
    """Device time per call in microseconds: the calls are captured into one CUDA graph (host overhead excluded)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3


def main():
    dev = torch.device("cuda:0")
    net = G_mod.StyleMapping(512, 8).to(dev)
    for M in (8, 16, 32):
        z = torch.randn(M, 512, device=dev)
        gy = torch.randn(M, 512, device=dev)
        with torch.no_grad():
            f = timeit(lambda: net(z))
            r = timeit(lambda: net.layers(z))

        def fb(fused):
            out = net(z) if fused else net.layers(z)
            out.backward(gy)
            for q in net.parameters():
                q.grad = None
        print("style mapping M=%2d: forward fused %.1f us (module-by-module %.1f us), forward+backward fused %.1f us (%.1f us)"
              % (M, f, r, timeit(lambda: fb(True)), timeit(lambda: fb(False))))


if __name__ == "__main__":
    main()
