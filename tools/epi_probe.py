"""Timing decomposition of short-K pixgemm launches (1x1 convs): run under MSG_B200_TC_VARIANT=0 / 32 (no stores) /
64 (no epilogue reads) to separate main loop, accumulator read-out and store cost.  Diagnostic only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_stylegan_b200 import _C

CASES = [  # (B, C, H, O, k, with_bias_act)
    (8, 512, 128, 512, 1, False), (8, 512, 128, 512, 1, True), (8, 512, 64, 512, 1, False),
    (8, 8, 256, 512, 1, False), (16, 256, 256, 128, 1, False), (16, 128, 256, 256, 1, False), (8, 512, 256, 512, 3, True),
    (16, 128, 256, 128, 3, True), (16, 256, 256, 128, 3, True)]


def main():
    dev = torch.device("cuda:0")
    cases = [CASES[int(a)] for a in sys.argv[1:]] or CASES
    for B, C, H, O, k, ba in cases:
        xs = [torch.randn(B, C, H, H, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(3)]
        w = torch.randn(O, C, k, k, device=dev) * 0.05
        bias = torch.randn(O, device=dev) if ba else None
        f = lambda i: _C.conv2d_forward(xs[i % 3], w, 1, k // 2, bias=bias, act=ba)
        for i in range(3):
            f(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for i in range(n):
            f(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        px = B * H * H
        gb = px * (C + O) * 4 / 1e9
        print("B=%d C=%d H=%d O=%d k=%d bias_act=%d : %.1f us  %.0f TFLOP/s  %.2f TB/s in+out" % (
            B, C, H, O, k, ba, ms * 1e3, 2.0 * px * C * O * k * k / ms / 1e9, gb / ms), flush=True)


if __name__ == "__main__":
    main()
