"""Diagnostic: which operation of the ADA pipeline breaks CUDA-graph capture (each step captured on its own)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from multi_stylegan_b200 import adaptive_discriminator_augmentation as A

dev = torch.device("cuda:0")
B, H, W = 8, 256, 256
x = torch.rand(B, 6, H, W, device=dev)
plan = A.build_plan(A.sample_draws(B, 0.5), B, H, W).to(dev)


def attempt(name, fn):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream()
        with torch.cuda.graph(g, stream=s):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        print("OK   ", name)
        return out
    except Exception as exc:
        print("FAIL ", name, type(exc).__name__, str(exc)[:160].replace("\n", " "))
        torch.cuda.synchronize()
        return None


fn_list = [
    ("flip-select", lambda: torch.where(plan[:B].view(B, 1, 1, 1) > 0.5, x.flip(dims=(-1,)), x)),
    ("theta view + clone", lambda: plan[2 * B + 2:].view(5, B, 2, 3)[0].clone()),
    ("warp", lambda: A.affine_warp(x, plan[2 * B + 2:].view(5, B, 2, 3)[1], mode=0)),
    ("shift", lambda: plan[2 * B:2 * B + 2].round().to(torch.long)),
    ("arange-roll", lambda: torch.remainder(torch.arange(H, device=dev) - plan[2 * B:2 * B + 2].round().to(torch.long)[0], H)),
    ("index_select", lambda: x.index_select(2, torch.remainder(torch.arange(H, device=dev) - 3, H))),
    ("apply_plan", lambda: A.apply_plan(x, plan)),
    ("pipeline(plan)", lambda: A.AugmentationPipeline()(x, 0.5, plan=plan)),
]
for name, fn in fn_list:
    fn()
    attempt(name, fn)
