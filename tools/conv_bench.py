"""Microbenchmark of the hot kernels at BASELINE config-2 sizes (batch 8, 512 channels): conv forward / dgrad /
wgrad on the tcgen05 engine, the channels-last FIR and the fused activation kernels.  CUDA-event timing, inputs
rotate over a pool larger than the 126 MB L2.  Diagnostic: prints a table, writes JSON lines with --out.

  python tools/conv_bench.py [--quick] [--out gpurun_out/conv_bench.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from multi_stylegan_b200 import _C  # noqa: E402

DEV = torch.device("cuda:0")


def timeit(fn, n_pool, iters=10, warm=3):
    for i in range(warm):
        fn(i % n_pool)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i % n_pool)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def conv_case(B, C, O, R, k, per_sample, stride=1, pad=None):
    pad = k // 2 if pad is None else pad
    nbytes = B * max(C, O) * R * R * 4
    n_pool = max(2, min(6, int(300e6 // max(nbytes, 1)) + 1))
    xs = [cl(torch.randn(B, C, R, R, device=DEV)) for _ in range(n_pool)]
    w = torch.randn((B, O, C, k, k) if per_sample else (O, C, k, k), device=DEV) / (C * k * k) ** 0.5
    y = _C.conv2d_forward(xs[0], w, stride, pad)
    dys = [cl(torch.randn_like(y)) for _ in range(n_pool)]
    flops = 2.0 * B * O * C * k * k * y.shape[2] * y.shape[3]
    out = {}
    out["fwd"] = timeit(lambda i: _C.conv2d_forward(xs[i], w, stride, pad), n_pool)
    out["dgrad"] = timeit(lambda i: _C.conv2d_dgrad(dys[i], w, (R, R), stride, pad), n_pool)
    out["wgrad"] = timeit(lambda i: _C.conv2d_wgrad(dys[i], xs[i], (k, k), stride, pad, per_sample), n_pool)
    return flops, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring filter on the conv case names")
    ap.add_argument("--no-bw", action="store_true", help="skip the bandwidth-bound kernels")
    args = ap.parse_args()
    rows = []
    # (name, B, C, O, R, k, per_sample, stride, pad)
    cases = [
        ("G conv 3x3 512->512 @256 shared W", 8, 512, 512, 256, 3, False, 1, None),
        ("G conv 3x3 512->512 @128 shared W", 8, 512, 512, 128, 3, False, 1, None),
        ("D conv 3x3 128->128 @256 B=16", 16, 128, 128, 256, 3, False, 1, None),
        ("G modconv 3x3 512->512 @256", 8, 512, 512, 256, 3, True, 1, None),
        ("G modconv 3x3 512->512 @128", 8, 512, 512, 128, 3, True, 1, None),
        ("G modconv 3x3 512->512 @64", 8, 512, 512, 64, 3, True, 1, None),
        ("D conv 3x3 128->128 @256", 8, 128, 128, 256, 3, False, 1, None),
        ("D conv 3x3 256->128 @256", 8, 256, 128, 256, 3, False, 1, None),
        ("D conv 3x3 256->256 @128", 8, 256, 256, 128, 3, False, 1, None),
        ("D conv 3x3 768->768 @32", 8, 768, 768, 32, 3, False, 1, None),
        ("D conv 1x1 256->128 @256", 8, 256, 128, 256, 1, False, 1, 0),
        ("D conv 3x3 s2 128->128 @256", 8, 128, 128, 256, 3, False, 2, 0),
        ("HBM-bound 1x1 256->128 @256 B=16", 16, 256, 128, 256, 1, False, 1, 0),
        ("HBM-bound 1x1 128->128 @256 B=16", 16, 128, 128, 256, 1, False, 1, 0),
        ("HBM-bound 1x1 128->256 @128 B=16", 16, 128, 256, 128, 1, False, 1, 0),
        ("HBM-bound 1x1 512->512 @128", 8, 512, 512, 128, 1, False, 1, 0),
        ("HBM-bound 1x1 384->256 @128 B=16", 16, 384, 256, 128, 1, False, 1, 0),
        ("D first conv 3x3 6->128 @256 B=16", 16, 6, 128, 256, 3, False, 1, None),
        ("D first conv 1x1 6->128 @256 B=16", 16, 6, 128, 256, 1, False, 1, 0),
    ]
    if args.quick:
        cases = [c for c in cases if c[0].startswith("D conv 3x3") and c[7] == 1] + cases[:1]
    if args.only:
        cases = [c for c in cases if args.only in c[0]]
    print("%-32s %9s %9s %9s   TFLOP/s fwd / dgrad / wgrad" % ("conv (tcgen05, TF32)", "fwd ms", "dgrad ms", "wgrad ms"))
    for name, B, C, O, R, k, per, s, p in cases:
        flops, t = conv_case(B, C, O, R, k, per, s, p)
        tf = {kk: flops / (v * 1e-3) / 1e12 for kk, v in t.items()}
        print("%-32s %9.3f %9.3f %9.3f   %6.0f / %6.0f / %6.0f" % (name, t["fwd"], t["dgrad"], t["wgrad"], tf["fwd"], tf["dgrad"], tf["wgrad"]))
        if name.startswith("HBM-bound"):
            nbytes = 4.0 * B * R * R * (C + O)          # algorithmic bytes: the two activation-sized tensors
            print("%-32s %9s %9s %9s   %6.0f / %6.0f / %6.0f GB/s" % ("", "", "", "", *(nbytes / (t[kk] * 1e-3) / 1e9 for kk in ("fwd", "dgrad", "wgrad"))))
        rows.append(dict(kind="conv", name=name, flops=flops, ms=t, tflops=tf))
        torch.cuda.empty_cache()

    # bandwidth-bound kernels: algorithmic bytes = 4 * (N_in + N_out)
    print("\n%-44s %9s %9s" % ("bandwidth-bound kernel (channels-last fp32)", "ms", "GB/s"))
    k4 = torch.tensor([1., 3., 3., 1.], device=DEV)
    k2d = (k4[None] * k4[:, None]) / 64
    for R in ([] if args.no_bw else [128] if args.quick else [64, 128, 256]):
        n_pool = 3
        xs = [cl(torch.randn(8, 512, R, R, device=DEV)) for _ in range(n_pool)]
        x4 = [x.permute(0, 2, 3, 1) for x in xs]
        for name, fn, bytes_ in [
            ("blur 4x4 pad(2,1)   [8,512,%d,%d]" % (R, R), lambda i: _C.upfirdn2d(x4[i], k2d * 4, 1, 1, 1, 1, 2, 1, 2, 1), 8 * xs[0].numel()),
            ("up2 4x4 pad(2,1)    [8,512,%d,%d]" % (R // 2, R // 2), lambda i: _C.upfirdn2d(x4[i][:, :R // 2, :R // 2].contiguous(), k2d, 2, 2, 1, 1, 2, 1, 2, 1), 4 * xs[0].numel() * 1.25),
            ("down2 4x4 pad(1,1)  [8,512,%d,%d]" % (R, R), lambda i: _C.upfirdn2d(x4[i], k2d, 1, 1, 2, 2, 1, 1, 1, 1), 4 * xs[0].numel() * 1.25),
            ("bias+lrelu fwd      [8,512,%d,%d]" % (R, R), lambda i: _C.fused_bias_act(xs[i], torch.zeros(512, device=DEV), xs[i].new_empty(0), 3, 0, 0.2, 1.0), 8 * xs[0].numel()),
            ("bias+lrelu bwd+dbias[8,512,%d,%d]" % (R, R), lambda i: _C.fused_bias_act_bwd(xs[i], xs[(i + 1) % n_pool], 0.2, 1.0, 512), 12 * xs[0].numel()),
        ]:
            if "up2" in name:
                # the slice+contiguous copy is not part of the op: pre-build inputs
                small = [x4[j][:, :R // 2, :R // 2].contiguous() for j in range(n_pool)]
                fn = (lambda i, small=small: _C.upfirdn2d(small[i], k2d, 2, 2, 1, 1, 2, 1, 2, 1))
            ms = timeit(fn, n_pool)
            print("%-44s %9.3f %9.0f" % (name, ms, bytes_ / (ms * 1e-3) / 1e9))
            rows.append(dict(kind="bw", name=name, ms=ms, gbs=bytes_ / (ms * 1e-3) / 1e9))
        del xs, x4
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
