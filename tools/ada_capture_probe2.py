"""Diagnostic: capture pieces of the ADA-wrapped train step at the benchmark size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from multi_stylegan_b200 import adaptive_discriminator_augmentation as A, config
import multi_stylegan_b200.multi_stylegan_generator as G_mod
import multi_stylegan_b200.u_net_2d_discriminator as D_mod

dev = torch.device("cuda:0")
B = 8
torch.manual_seed(0)
G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev)
D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True).to(dev)
ada = A.AdaptiveDiscriminatorAugmentation(D)
ada.p = 0.5
real = torch.rand(B, 2, 3, 256, 256, device=dev)


def attempt(name, fn, warm=True):
    if warm:
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream()
        ada.begin_plan_capture(2 * B, dev)
        with torch.cuda.graph(g, stream=s):
            out = fn()
        cap = ada.end_plan_capture()
        ada.refresh_plans(cap)
        g.replay()
        torch.cuda.synchronize()
        print("OK   ", name, flush=True)
    except Exception as exc:
        ada._capture = None
        print("FAIL ", name, type(exc).__name__, str(exc)[:200].replace("\n", " "), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception:
            pass


def gen(grad):
    with torch.set_grad_enabled(grad):
        return G(torch.randn(B, 512, device=dev))


def pair_nograd():
    with torch.no_grad():
        return ada.forward_pair(real.clone(), gen(False))


def pair_bwd():
    (a, b), (c, d) = ada.forward_pair(real.clone(), gen(False))
    (a.sum() + b.mean() + c.sum() + d.mean()).backward()


def g_step():
    for p in D.parameters():
        p.requires_grad_(False)
    try:
        s, px = ada(gen(True), is_real=False)
    finally:
        for p in D.parameters():
            p.requires_grad_(True)
    (s.sum() + px.mean()).backward()


def aug_only_cl():
    with torch.no_grad():
        f = gen(False)
        return ada._augment(f.flatten(start_dim=1, end_dim=2), None)


def aug_bwd_cl():
    f = gen(True)
    out = ada._augment(f.flatten(start_dim=1, end_dim=2), None)
    out.sum().backward()


for name, fn in [("augment(G output) no_grad", aug_only_cl), ("augment(G output) fwd+bwd", aug_bwd_cl),
                 ("forward_pair no_grad", pair_nograd), ("forward_pair fwd+bwd", pair_bwd), ("generator step", g_step)]:
    attempt(name, fn)
