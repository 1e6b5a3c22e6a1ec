"""Does the data path stay off the train step's critical path?  

Writes a synthetic 16-bit TIFF tree in the reference's naming, then times the same iterations of the default-size train step
(a) on device-resident batches and (b) fed by dataset.TFLMDatasetGAN -> dataset.DeviceLoader (decode threads -> pinned slabs
-> copy stream).  CUDA-event timing over whole epochs; a diagnostic, prints one JSON line.

    python tools/loader_bench.py [--batch 8] [--epochs 6] [--frames 35]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--frames", type=int, default=35, help="time steps of the one synthetic trap (windows = frames - 2)")
    args = ap.parse_args()
    import cv2
    from multi_stylegan_b200 import _lib, config
    from multi_stylegan_b200.dataset import DeviceLoader, TFLMDatasetGAN
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
    mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=True)
    with tempfile.TemporaryDirectory(prefix="msgds") as root:
        os.makedirs(os.path.join(root, "pos01"))
        rng = np.random.default_rng(0)
        for t in range(args.frames):
            for channel, hi in (("BF0", 4000), ("GFP", 3000)):
                cv2.imwrite(os.path.join(root, "pos01", "x_trap0001-%s_000_w_e_%03d.tif" % (channel, t)),
                            rng.integers(0, hi, size=(256, 256), dtype=np.uint16))
        ds = TFLMDatasetGAN(path=root, no_rfp=True, z_position_indications=("_000_",))
        t0 = time.perf_counter()
        for i in range(16):
            ds[i]
        decode_ms = (time.perf_counter() - t0) / 16 * 1e3
        loader = DeviceLoader(ds, batch_size=args.batch, device=dev, workers=8, depth=3)
        steps = len(loader) * args.epochs
        resident = [b.clone() for b in loader]

        def run(feed):
            mw.iteration = 0
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            n = 0
            for e in range(args.epochs):
                for batch in feed(e):
                    mw.train_step(batch)
                    n += 1
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n, n

        def from_loader(e):
            loader.set_epoch(e)
            return loader
        for _ in range(2):                                        # warm-up: eager first occurrences, then the graph captures
            run(lambda e: resident)
        ms_res, ms_load = [], []
        for _ in range(2):                                        # alternate the two feeds
            ms_res.append(run(lambda e: resident)[0])
            t, n = run(from_loader)
            ms_load.append(t)
        ms_res, ms_load = min(ms_res), min(ms_load)
    print(json.dumps({"metric": "train_step_ms_per_step", "per_gpu_batch": args.batch, "steps": n,
                      "device_resident_batches_ms": ms_res, "device_loader_fed_ms": ms_load,
                      "loader_overhead": ms_load / ms_res - 1.0, "samples": len(ds),
                      "decode_plus_normalise_ms_per_sample_one_thread": decode_ms,
                      "note": "default 512-channel / 256x256 networks, CUDA-graph replay, %d iterations incl. the lazy ones; "
                              "the loader decodes 16-bit TIFFs on 8 threads into pinned slabs 2 batches ahead" % n}))


if __name__ == "__main__":
    main()
