"""Randomised tcgen05-vs-CUDA-core comparison of the conv primitives over small odd shapes (on GPU)."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_stylegan_b200 import _C, _lib

def main(n=300, seed=0):
    rnd = random.Random(seed)
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    bad = 0
    for i in range(n):
        B = rnd.choice([1, 2, 3]); C = rnd.choice([3, 4, 6, 8, 17, 24, 33, 48, 64, 70]); O = rnd.choice([1, 3, 8, 16, 24, 33, 48, 64, 130])
        k = rnd.choice([1, 2, 3]); s = rnd.choice([1, 2]); p = rnd.choice([0, 1]) if k > 1 else 0
        H = rnd.choice([1, 2, 3, 4, 7, 8, 15, 16, 31, 32, 33, 40]); W = rnd.choice([1, 2, 3, 4, 7, 8, 15, 16, 31, 32, 33, 40])
        per = rnd.random() < 0.5
        if H + 2 * p < k or W + 2 * p < k:
            continue
        x = torch.randn(B, C, H, W, device=dev)
        w = torch.randn((B, O, C, k, k) if per else (O, C, k, k), device=dev) / (C * k * k) ** 0.5
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        dy = torch.randn(B, O, OH, OW, device=dev)
        res = {}
        for name, flags in (("simt", _lib.CONV_FORCE_SIMT), ("tc", _lib.CONV_FORCE_TC)):
            _C.conv_flags = flags
            res[name] = (_C.conv2d_forward(x, w, s, p), _C.conv2d_dgrad(dy, w, (H, W), s, p), _C.conv2d_wgrad(dy, x, (k, k), s, p, per))
        torch.cuda.synchronize()
        errs = [((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item() for a, b in zip(res["tc"], res["simt"])]
        if max(errs) > 5e-3 or any(e != e for e in errs):
            bad += 1
            print("BAD", dict(B=B, C=C, O=O, H=H, W=W, k=k, s=s, p=p, per=per), ["%.4f" % e for e in errs], flush=True)
    print("done; bad =", bad, "of", n)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
