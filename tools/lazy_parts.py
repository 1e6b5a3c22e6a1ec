"""Host / device time of the two lazy regularisation blocks of the train step (R1 on D, path length on G)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multi_stylegan_b200 import config
import multi_stylegan_b200.multi_stylegan_generator as G_mod
import multi_stylegan_b200.u_net_2d_discriminator as D_mod
from multi_stylegan_b200.model_wrapper import ModelWrapper
from multi_stylegan_b200 import higher_order_gradients

dev = torch.device("cuda:0")
torch.manual_seed(0)
G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev)
D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True).to(dev)
hp = dict(config.generation_hyperparameters)
opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-4, lr_style=2e-6), betas=hp["betas"], fused=True)
opt_d = torch.optim.Adam(D.parameters(), lr=6e-4, betas=hp["betas"], fused=True)
mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev)
real = torch.rand(8, 2, 3, 256, 256, device=dev)


def timed(name, fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    hs, ds = [], []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        hs.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        ds.append(a.elapsed_time(b))
    print("%-34s host %7.1f ms   device-span %7.1f ms" % (name, min(hs), min(ds)))


def r1_forward():
    mw._zero()
    x = real.detach().requires_grad_(True)
    with higher_order_gradients():          # what ModelWrapper does for the R1 step (_mode.py)
        s, p = D(x, is_real=False, is_cut_mix=True)
    return x, s, p


def r1_full():
    x, s, p = r1_forward()
    r1 = mw.discriminator_regularization_loss(s, x, p)
    (10.0 * r1).backward()


def r1_first_order_only():
    x, s, p = r1_forward()
    torch.autograd.grad((s.sum(), p.sum()), x, create_graph=True)


def pl_full():
    mw._zero()
    grads = G(input=mw._noise(4), return_path_length_grads=True)
    loss, _ = mw.path_length_regularization(grads)
    (2.0 * loss).backward()


def pl_first_order_only():
    mw._zero()
    G(input=mw._noise(4), return_path_length_grads=True)


def g_forward4():
    with torch.no_grad():
        G(input=mw._noise(4))


timed("D forward (B=8, grad)", lambda: r1_forward())
timed("R1: forward + grad wrt image", r1_first_order_only)
timed("R1: full (double backward)", r1_full)
timed("G forward (B=4, no grad)", g_forward4)
timed("PL: forward + grad wrt latent", pl_first_order_only)
timed("PL: full (double backward)", pl_full)
