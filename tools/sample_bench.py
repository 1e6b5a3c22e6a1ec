"""BASELINE config 5 (second half): EMA-generator sampling throughput (sequences/s), the call of
scripts/get_gan_samples.py:40-42 — `generator(noise)` under no_grad, single-style z, fresh per-layer noise — on the
default 512-channel / 256x256 generator with random-init weights.  CUDA-event timing; one JSON line per batch size.

  python tools/sample_bench.py [--batches 8 16 32 64] [--iters 5]
  torchrun --nproc-per-node N tools/sample_bench.py ...      (independent replicas, aggregate = sum over ranks)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[8, 16, 32, 64])
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    from multi_stylegan_b200 import _C, config, dist as mdist
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    local_rank = mdist.init_from_env()
    world, rank = mdist.world_size(), mdist.rank()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev).eval()
    for B in args.batches:
        z = torch.randn(B, 512, device=dev)
        with torch.no_grad():
            for _ in range(2):
                img = G(z)
            torch.cuda.synchronize()
            n0 = _C.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.iters):
                img = G(torch.randn(B, 512, device=dev))
            b.record()
            torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / args.iters], device=dev)
        if world > 1:
            import torch.distributed as tdist
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"metric": "ema_generator_sampling_sequences_per_sec", "value": world * B / (float(ms) * 1e-3),
                              "unit": "sequences/s", "n_gpus": world, "per_gpu_batch": B, "ms_per_batch": float(ms),
                              "output_shape": list(img.shape), "gpu_launches_per_batch": (_C.launch_count() - n0) // args.iters,
                              "dtype": "tf32", "data": "synthetic z, random-init weights"}), flush=True)
        del img
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
