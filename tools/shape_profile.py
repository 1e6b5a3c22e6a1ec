"""Per-shape table of the tcgen05 conv launches of plain training iterations (msg_profile_*: CUDA events around every
launch, eager issue): time, TFLOP/s and the time "lost" against a 900 TFLOP/s target, sorted by the loss.

    python tools/shape_profile.py [--iters 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    from multi_stylegan_b200 import _C, _lib, config
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    _lib.lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev)
    D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True).to(dev)
    hp = dict(config.generation_hyperparameters)
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-4, lr_style=2e-6), betas=hp["betas"], fused=True, capturable=True)
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-4, betas=hp["betas"], fused=True, capturable=True)
    mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=False)
    real = torch.rand(args.batch, 2, 3, 256, 256, device=dev)
    for _ in range(3):
        mw.iteration = 0
        mw.train_step(real)
    torch.cuda.synchronize()
    _C.profile_enable(True)
    for _ in range(args.iters):
        mw.iteration = 0
        mw.train_step(real)
    torch.cuda.synchronize()
    rows = _C.profile_summary()
    _C.profile_enable(False)
    tot_ms = sum(r["ms_total"] for r in rows) / args.iters
    tot_fl = sum(r["flops_per_launch"] * r["launches"] for r in rows) / args.iters
    print("tcgen05 conv launches per iteration: %.2f ms, %.1f TFLOP -> %.0f TFLOP/s" % (tot_ms, tot_fl / 1e12, tot_fl / tot_ms / 1e9))
    print("%-8s %4s %6s %6s %9s %5s %8s %8s %8s" % ("kind", "taps", "K", "N", "pixels", "n/it", "ms/it", "TFLOP/s", "lost ms"))
    out = []
    for r in rows:
        ms = r["ms_total"] / args.iters
        fl = r["flops_per_launch"] * r["launches"] / args.iters
        out.append((ms - fl / 900e9, r, ms, fl))
    for lost, r, ms, fl in sorted(out, key=lambda t: -t[0])[:45]:
        print("%-8s %4d %6d %6d %9d %5.1f %8.3f %8.0f %8.3f" % (r["kind"], r["taps"], r["k_channels"], r["n_channels"], r["pixels"],
                                                              r["launches"] / args.iters, ms, fl / ms / 1e9, lost))


if __name__ == "__main__":
    main()
