"""Which Python call sites launch the non-library ("glue") kernels of one plain training iteration?

Runs the bench workload eagerly (no CUDA graphs), profiles ONE plain iteration with torch.profiler (with_stack) and prints
device time per (operator, innermost package frame).  A diagnostic: numbers taken under the profiler are not bench values.

    python tools/glue_profile.py [--batch 8] [--top 60]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--aten-only", action="store_true")
    ap.add_argument("--kernels", action="store_true", help="also print the device time per kernel name")
    ap.add_argument("--lazy", action="store_true", help="profile a lazy (R1 + path length) iteration instead")
    args = ap.parse_args()
    from multi_stylegan_b200 import _lib, config
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
    lazy_every = hp["lazy_generator_regularization"]
    for _ in range(3):
        mw.iteration = lazy_every - 1 if args.lazy else 0
        mw.train_step(real)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    mw.iteration = lazy_every - 1 if args.lazy else 0
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        mw.train_step(real)
        torch.cuda.synchronize()
    rows = collections.defaultdict(lambda: [0.0, 0])
    total = 0.0
    pkg = "multi_stylegan_b200"
    for ev in prof.key_averages(group_by_input_shape=True, group_by_stack_n=16):
        t = getattr(ev, "self_device_time_total", None)
        if t is None:
            t = getattr(ev, "self_cuda_time_total", 0.0)
        if not t:
            continue
        total += t
        site = "?"
        for fr in (ev.stack or []):
            if pkg in fr or "bench.py" in fr:
                site = fr.replace(ROOT + "/", "")
                break
        shp = str(ev.input_shapes)[:70] if ev.key.startswith('aten::') else ''
        key = ((ev.key + ' ' + shp)[:100], site[:80])
        rows[key][0] += t
        rows[key][1] += ev.count
    print("device time in the profiled iteration: %.2f ms" % (total / 1e3))
    if args.kernels:
        from torch.autograd import DeviceType
        kern = collections.defaultdict(lambda: [0.0, 0])
        for ev in prof.events():
            if ev.device_type == DeviceType.CUDA:
                kern[ev.name[:110]][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
                kern[ev.name[:110]][1] += 1
        ktotal = sum(v[0] for v in kern.values())
        print("\n== device time per kernel (%.2f ms) ==" % (ktotal / 1e3))
        for name, (t, n) in sorted(kern.items(), key=lambda kv: -kv[1][0])[:args.top]:
            print("%8.3f ms %5.1f%% %5d  %s" % (t / 1e3, 100 * t / ktotal, n, name))
    by_site = collections.defaultdict(float)
    for (name, site), (t, n) in rows.items():
        if name.startswith("aten::"):
            by_site[site] += t
    print("\n== ATen / library device time per call site ==")
    for site, t in sorted(by_site.items(), key=lambda kv: -kv[1])[:args.top]:
        print("%8.3f ms  %s" % (t / 1e3, site))
    print("\n== per (operator, call site) ==")
    for (name, site), (t, n) in sorted(((k, v) for k, v in rows.items() if k[0].startswith("aten::") or not args.aten_only),
                                       key=lambda kv: -kv[1][0])[:args.top]:
        print("%8.3f ms %5d  %-100s %s" % (t / 1e3, n, name, site))


if __name__ == "__main__":
    main()
