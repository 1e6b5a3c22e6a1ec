"""Diagnostic: which operation of the ADA path breaks CUDA-graph capture (each step captured on its own)."""
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
        print("OK   ", name, flush=True)
        return out
    except Exception as exc:
        print("FAIL ", name, type(exc).__name__, str(exc)[:200].replace("\n", " "), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return None


def fwd_bwd(fn):
    def run():
        xi = x.clone().requires_grad_(True)
        out = fn(xi * 1.0)
        out.sum().backward()
        return xi.grad
    return run


th = plan[2 * B + 2:].view(5, B, 2, 3)
fn_list = [
    ("flip fwd+bwd", fwd_bwd(lambda t: torch.where(plan[:B].view(B, 1, 1, 1) > 0.5, t.flip(dims=(-1,)), t))),
    ("warp fwd+bwd", fwd_bwd(lambda t: A.affine_warp(t, th[1], mode=0))),
    ("index_select fwd+bwd", fwd_bwd(lambda t: t.index_select(2, torch.remainder(torch.arange(H, device=dev) - 3, H)))),
    ("index_select dim3 fwd+bwd", fwd_bwd(lambda t: t.index_select(3, torch.remainder(torch.arange(W, device=dev) - 3, W)))),
    ("apply_plan fwd+bwd", fwd_bwd(lambda t: A.apply_plan(t, plan))),
]
for name, fn in fn_list:
    fn()
    attempt(name, fn)
