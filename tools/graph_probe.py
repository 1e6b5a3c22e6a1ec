"""Experiment: how much of the plain train step is launch gaps?  Captures one plain iteration in a CUDA graph
(host-side random decisions frozen - diagnostic only, not a product path) and times replay against eager.

  python tools/graph_probe.py [--steps 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    from multi_stylegan_b200 import config
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    from multi_stylegan_b200.model_wrapper import ModelWrapper

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev)
    D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True).to(dev)
    hp = dict(config.generation_hyperparameters)
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-4, lr_style=2e-6), betas=hp["betas"], fused=True, capturable=True)
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-4, betas=hp["betas"], fused=True, capturable=True)
    mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev)
    real = torch.rand(args.batch, 2, 3, 256, 256, device=dev)

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def plain():
        mw.iteration = 0
        mw.train_step(real)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            plain()
    torch.cuda.current_stream().wait_stream(s)
    print("eager  %.2f ms/step" % timed(plain, args.steps), flush=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plain()
    print("graph  %.2f ms/step" % timed(g.replay, args.steps), flush=True)


if __name__ == "__main__":
    main()
