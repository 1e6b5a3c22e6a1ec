"""Per-kernel time breakdown of the train step (torch.profiler / CUPTI; no nsys in this image).

  python tools/step_profile.py [--steps 2] [--lazy] [--batch 8] [--out gpurun_out/step_profile.txt]

Prints kernels grouped by name, sorted by total device time.  Diagnostic only (never a bench number)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch.profiler import ProfilerActivity, profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--lazy", action="store_true", help="profile an iteration that runs R1 + path length")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.txt"))
    ap.add_argument("--rows", type=int, default=70)
    ap.add_argument("--shapes", action="store_true", help="also list torch copy / elementwise ops grouped by input shape")
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
    # same Adam as train_multi_stylegan.py:53-57; fused=True only selects PyTorch's single-kernel implementation
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-4, lr_style=2e-6), betas=hp["betas"], fused=True)
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-4, betas=hp["betas"], fused=True)
    mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev)
    real = torch.rand(args.batch, 2, 3, 256, 256, device=dev)
    for i in range(3):
        mw.iteration = 14 if (i == 0 and args.lazy) else 0
        mw.train_step(real)
    torch.cuda.synchronize()
    import time as _time
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=args.shapes) as prof:
        for i in range(args.steps):
            mw.iteration = 15 if args.lazy else 0
            mw.train_step(real)
        torch.cuda.synchronize()
    events = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    if not events:
        events = [e for e in prof.key_averages() if e.device_time_total > 0]
    events.sort(key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in events)
    lines = ["total device time %.2f ms over %d step(s) (%s)" % (total / 1e3, args.steps, "lazy" if args.lazy else "plain"),
             "%10s %7s %6s %9s  %s" % ("total_ms", "share", "calls", "avg_us", "kernel")]
    for e in events[:args.rows]:
        lines.append("%10.3f %6.2f%% %6d %9.1f  %s" % (e.device_time_total / 1e3, 100.0 * e.device_time_total / total,
                                                       e.count, e.device_time_total / e.count, e.key[:150]))
    if args.shapes:
        ops = [e for e in prof.key_averages(group_by_input_shape=True)
               if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::cat", "aten::sum", "aten::div",
                            "aten::mul_", "aten::clone", "aten::repeat", "aten::sub", "aten::neg", "aten::fill_", "aten::zero_")
               and e.device_time_total > 0]
        ops.sort(key=lambda e: -e.device_time_total)
        lines.append("")
        lines.append("torch elementwise / copy ops by input shape (device time, %d step(s))" % args.steps)
        for e in ops[:45]:
            lines.append("%10.3f ms %5d calls  %-14s %s" % (e.device_time_total / 1e3, e.count, e.key, str(e.input_shapes)[:120]))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:45] if not args.shapes else lines[-46:]))


if __name__ == "__main__":
    main()
